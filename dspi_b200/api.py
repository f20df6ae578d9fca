"""ctypes binding of the C ABI (``include/dspi_b200.h``) — the call a Python host makes.

There is no CPU path in this package: if ``libdspi_b200.so`` is missing or no
sm_100 device is visible, construction fails loudly.
"""
import ctypes as C
import os

import numpy as np

from . import layouts as L

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdspi_b200.so")

ARITH_F32_FUSED, ARITH_F32_STRICT, ARITH_Q28 = 0, 1, 2
ARITH = {"f32f": ARITH_F32_FUSED, "f32s": ARITH_F32_STRICT, "q28": ARITH_Q28}

# every symbol include/dspi_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "dspi_last_error", "dspi_device_count", "dspi_compute_coefficients_f32", "dspi_compute_coefficients_q28",
    "dspi_eq_create", "dspi_eq_destroy", "dspi_eq_upload_biquads", "dspi_eq_download_biquads", "dspi_eq_set_param",
    "dspi_eq_process_device", "dspi_eq_process_host", "dspi_eq_sync", "dspi_eq_stream", "dspi_eq_launch_count",
    "dspi_eq_kernel_info", "dspi_eq_set_params_device", "dspi_chain_set_eq_params_device", "dspi_chainq_set_eq_params_device",
    "dspi_chain_state_size", "dspi_chain_state_export", "dspi_chain_state_import", "dspi_chainq_state_size", "dspi_chainq_state_export", "dspi_chainq_state_import",
    "dspi_host_alloc", "dspi_host_free",
    "dspi_chain_create", "dspi_chain_destroy", "dspi_chain_set_params", "dspi_chain_upload_biquads", "dspi_chain_download_biquads",
    "dspi_chain_reset_state", "dspi_chain_process_host", "dspi_chain_process_device", "dspi_chain_sync", "dspi_chain_stream",
    "dspi_chain_launch_count", "dspi_delay_samples",
    "dspi_crossfeed_compute_coefficients_f32", "dspi_leveller_compute_coefficients", "dspi_loudness_compute_table_f32", "dspi_host_volume",
    "dspi_chainq_create", "dspi_chainq_destroy", "dspi_chainq_set_params", "dspi_chainq_upload_biquads", "dspi_chainq_download_biquads",
    "dspi_chainq_reset_state", "dspi_chainq_process_host", "dspi_chainq_process_device", "dspi_chainq_sync", "dspi_chainq_launch_count",
    "dspi_crossfeed_compute_coefficients_q28", "dspi_loudness_compute_table_q28",
    "dspi_spdif_lookup_table", "dspi_spdif_encode_device", "dspi_spdif_encode_host",
    "dspi_bulk_state_defaults", "dspi_bulk_params_apply", "dspi_bulk_params_collect", "dspi_bulk_state_to_chain_f32", "dspi_bulk_state_to_chain_q28",
    "dspi_preset_slot_size", "dspi_crc32", "dspi_preset_slot_apply", "dspi_preset_slot_collect",
    "dspi_preamp", "dspi_master_volume", "dspi_preset_mute_arm", "dspi_preset_mute_step",
    "dspi_nccl_unique_id", "dspi_sg_create", "dspi_sg_destroy", "dspi_sg_process",
    "dspi_chainq_stream", "dspi_eq_process_device_range", "dspi_bind_host_to_device", "dspi_eqx_create", "dspi_eqx_destroy", "dspi_eqx_shard_range",
    "dspi_eqx_upload_biquads", "dspi_eqx_download_biquads", "dspi_eqx_process_host", "dspi_eqx_process_root", "dspi_eqx_launch_count",
    "dspi_chain_set_dynamics_device", "dspi_chainq_set_dynamics_device", "dspi_chain_sm_partition", "dspi_chainq_sm_partition",
    "dspi_chain_set_preset_mute", "dspi_chain_get_preset_mute", "dspi_chainq_set_preset_mute", "dspi_chainq_get_preset_mute",
]


class DspiError(RuntimeError):
    pass


class _ChainDesc(C.Structure):
    _fields_ = [("arith", C.c_uint32), ("n_instances", C.c_uint32), ("n_bands", C.c_uint32),
                ("device", C.c_int32), ("max_frames", C.c_uint32)]


class _EqDesc(C.Structure):
    _fields_ = [("arith", C.c_uint32), ("n_channels", C.c_uint32), ("n_bands", C.c_uint32),
                ("device", C.c_int32), ("flags", C.c_uint32)]


_lib = None


class _EqxDesc(C.Structure):
    _fields_ = [("arith", C.c_uint32), ("n_channels", C.c_uint32), ("n_bands", C.c_uint32), ("n_devices", C.c_uint32),
                ("devices", C.c_int32 * 8), ("flags", C.c_uint32)]


def lib():
    """The loaded shared library (never built implicitly here: see ``dspi_b200.build``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DspiError(f"{LIB_PATH} is missing - build it with `python -m dspi_b200.build` "
                            "(there is no CPU fallback)")
        h = C.CDLL(LIB_PATH)
        vp, u32 = C.c_void_p, C.c_uint32
        h.dspi_last_error.restype = C.c_char_p
        h.dspi_device_count.restype = C.c_int
        h.dspi_compute_coefficients_f32.argtypes = [vp, vp, C.c_float]
        h.dspi_compute_coefficients_q28.argtypes = [vp, vp, C.c_float]
        h.dspi_eq_create.argtypes = [C.POINTER(vp), C.POINTER(_EqDesc)]
        h.dspi_eq_destroy.argtypes = [vp]
        h.dspi_eq_upload_biquads.argtypes = [vp, u32, u32, vp]
        h.dspi_eq_download_biquads.argtypes = [vp, u32, u32, vp]
        h.dspi_eq_set_param.argtypes = [vp, u32, vp, C.c_float]
        h.dspi_eq_process_device.argtypes = [vp, vp, u32, u32]
        h.dspi_eq_process_host.argtypes = [vp, vp, u32]
        h.dspi_eq_sync.argtypes = [vp]
        h.dspi_eq_stream.argtypes = [vp]
        h.dspi_eq_stream.restype = vp
        h.dspi_eq_launch_count.argtypes = [vp]
        h.dspi_eq_launch_count.restype = C.c_uint64
        h.dspi_eq_kernel_info.argtypes = [vp, C.c_char_p, C.c_size_t]
        h.dspi_chain_create.argtypes = [C.POINTER(vp), C.POINTER(_ChainDesc)]
        h.dspi_chain_destroy.argtypes = [vp]
        h.dspi_chain_set_params.argtypes = [vp, u32, u32, vp]
        h.dspi_chain_upload_biquads.argtypes = [vp, u32, u32, vp]
        h.dspi_chain_download_biquads.argtypes = [vp, u32, u32, vp]
        h.dspi_chain_reset_state.argtypes = [vp]
        h.dspi_chain_process_host.argtypes = [vp, vp, u32, u32, u32, vp, vp, vp]
        h.dspi_chain_process_device.argtypes = [vp, vp, u32, u32, u32, vp, vp, vp]
        h.dspi_chain_sync.argtypes = [vp]
        h.dspi_chain_stream.argtypes = [vp]
        h.dspi_chain_stream.restype = vp
        h.dspi_chain_launch_count.argtypes = [vp]
        h.dspi_chain_launch_count.restype = C.c_uint64
        h.dspi_delay_samples.argtypes = [C.c_float, C.c_float, C.c_int]
        h.dspi_delay_samples.restype = C.c_int32
        h.dspi_chainq_create.argtypes = [C.POINTER(vp), C.POINTER(_ChainDesc)]
        h.dspi_chainq_destroy.argtypes = [vp]
        h.dspi_chainq_set_params.argtypes = [vp, u32, u32, vp]
        h.dspi_chainq_upload_biquads.argtypes = [vp, u32, u32, vp]
        h.dspi_chainq_download_biquads.argtypes = [vp, u32, u32, vp]
        h.dspi_chainq_reset_state.argtypes = [vp]
        h.dspi_chainq_process_host.argtypes = [vp, vp, u32, u32, u32, vp, vp, vp]
        h.dspi_chainq_process_device.argtypes = [vp, vp, u32, u32, u32, vp, vp, vp]
        h.dspi_chainq_sync.argtypes = [vp]
        h.dspi_chainq_stream.argtypes = [vp]
        h.dspi_chainq_stream.restype = vp
        h.dspi_chainq_launch_count.argtypes = [vp]
        h.dspi_chainq_launch_count.restype = C.c_uint64
        h.dspi_crossfeed_compute_coefficients_q28.argtypes = [vp, vp, C.c_float]
        h.dspi_loudness_compute_table_q28.argtypes = [vp, C.c_float, C.c_float, C.c_float]
        h.dspi_crossfeed_compute_coefficients_f32.argtypes = [vp, vp, C.c_float]
        h.dspi_leveller_compute_coefficients.argtypes = [vp, vp, C.c_float]
        h.dspi_loudness_compute_table_f32.argtypes = [vp, C.c_float, C.c_float, C.c_float]
        h.dspi_host_volume.argtypes = [C.c_int16, vp]
        h.dspi_host_volume.restype = C.c_int16
        h.dspi_preamp.argtypes = [C.c_float, vp, vp]
        h.dspi_master_volume.argtypes = [C.c_float, vp, vp]
        h.dspi_preset_mute_arm.argtypes = [vp, u32]
        h.dspi_preset_mute_step.argtypes = [vp, u32, u32]
        h.dspi_preset_mute_step.restype = C.c_float
        for pre in ("dspi_chain", "dspi_chainq"):
            getattr(h, pre + "_set_preset_mute").argtypes = [vp, u32, u32, vp, u32]
            getattr(h, pre + "_get_preset_mute").argtypes = [vp, u32, u32, vp]
            getattr(h, pre + "_set_dynamics_device").argtypes = [vp, u32, u32, vp, C.c_float]
            getattr(h, pre + "_sm_partition").argtypes = [vp, vp, vp]
        h.dspi_eq_process_device_range.argtypes = [vp, vp, u32, u32, u32, u32]
        h.dspi_bind_host_to_device.argtypes = [C.c_int]
        h.dspi_nccl_unique_id.argtypes = [vp]
        h.dspi_sg_create.argtypes = [C.POINTER(vp), vp, C.c_int, vp, C.c_int, C.c_int, C.c_int]
        h.dspi_sg_destroy.argtypes = [vp]
        h.dspi_sg_process.argtypes = [vp, vp, u32, u32, u32]
        h.dspi_eqx_create.argtypes = [C.POINTER(vp), C.POINTER(_EqxDesc)]
        h.dspi_eqx_destroy.argtypes = [vp]
        h.dspi_eqx_shard_range.argtypes = [u32, u32, u32, vp, vp]
        h.dspi_eqx_upload_biquads.argtypes = [vp, u32, u32, vp]
        h.dspi_eqx_download_biquads.argtypes = [vp, u32, u32, vp]
        h.dspi_eqx_process_host.argtypes = [vp, vp, u32]
        h.dspi_eqx_process_root.argtypes = [vp, vp, u32, u32]
        h.dspi_eqx_launch_count.argtypes = [vp]
        h.dspi_eqx_launch_count.restype = C.c_uint64
        h.dspi_host_alloc.argtypes = [C.c_size_t]
        h.dspi_host_alloc.restype = vp
        h.dspi_host_free.argtypes = [vp]
        _lib = h
    return _lib


def _check(rc):
    if rc != 0:
        raise DspiError(f"dspi error {rc}: {lib().dspi_last_error().decode()}")


def compute_coefficients(params, q28=False, fs=48000.0, biquads=None):
    """Host-side ``dsp_compute_coefficients`` over an array of recipes.

    ``params``: EQ_PARAM array (clamped in place like the reference);
    ``biquads``: matching BIQUAD_* array to update (state kept) or None for fresh zeros.
    """
    h = lib()
    dt = L.BIQUAD_Q28 if q28 else L.BIQUAD_F32
    if biquads is None:
        biquads = np.zeros(params.shape, dt)
    fn = h.dspi_compute_coefficients_q28 if q28 else h.dspi_compute_coefficients_f32
    pp, bb = params.reshape(-1), biquads.reshape(-1)
    bp, bbp = pp.ctypes.data, bb.ctypes.data
    for i in range(pp.shape[0]):
        fn(bp + i * pp.dtype.itemsize, bbp + i * bb.dtype.itemsize, fs)
    return biquads


class PinnedBuffer:
    """Page-locked host memory viewed as a numpy array."""

    def __init__(self, shape, dtype):
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self.ptr = lib().dspi_host_alloc(self.nbytes)
        if not self.ptr:
            raise DspiError("dspi_host_alloc failed")
        buf = (C.c_uint8 * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype).reshape(self.shape)

    def free(self):
        if self.ptr:
            self.array = None
            lib().dspi_host_free(self.ptr)
            self.ptr = None


class EqEngine:
    """Many independent 10-band cascades on one B200 (``dspi_eq_*``)."""

    def __init__(self, arith, n_channels, n_bands=L.NUM_BANDS, device=0):
        self.arith = ARITH[arith] if isinstance(arith, str) else int(arith)
        self.q28 = self.arith == ARITH_Q28
        self.n_channels, self.n_bands, self.device = int(n_channels), int(n_bands), int(device)
        self.biquad_dtype = L.BIQUAD_Q28 if self.q28 else L.BIQUAD_F32
        self.sample_dtype = np.int32 if self.q28 else np.float32
        self._h = C.c_void_p()
        desc = _EqDesc(self.arith, self.n_channels, self.n_bands, self.device, 0)
        _check(lib().dspi_eq_create(C.byref(self._h), C.byref(desc)))

    def close(self):
        if self._h:
            lib().dspi_eq_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, biquads, ch0=0):
        b = np.ascontiguousarray(biquads)
        assert b.dtype == self.biquad_dtype and b.ndim == 2 and b.shape[1] == L.MAX_BANDS
        _check(lib().dspi_eq_upload_biquads(self._h, ch0, b.shape[0], b.ctypes.data))

    def download(self, n=None, ch0=0):
        n = self.n_channels - ch0 if n is None else n
        out = np.zeros((n, L.MAX_BANDS), self.biquad_dtype)
        _check(lib().dspi_eq_download_biquads(self._h, ch0, n, out.ctypes.data))
        return out

    def set_param(self, channel, param, fs):
        p = np.array([param], dtype=L.EQ_PARAM) if not isinstance(param, np.ndarray) else param.reshape(1).copy()
        _check(lib().dspi_eq_set_param(self._h, channel, p.ctypes.data, fs))
        return p[0]

    def process_device(self, dev_ptr, T, ld=None):
        _check(lib().dspi_eq_process_device(self._h, C.c_void_p(int(dev_ptr)), T, T if ld is None else ld))

    def process_device_range(self, dev_ptr, T, ld, ch0, n):
        """Channels [ch0, ch0+n) only; ``dev_ptr`` is the row of channel ch0."""
        _check(lib().dspi_eq_process_device_range(self._h, C.c_void_p(int(dev_ptr)), int(T), int(ld), int(ch0), int(n)))

    def process_host(self, samples):
        """``samples``: C-contiguous [n_channels, T] numpy array (or PinnedBuffer.array); in place."""
        assert samples.flags["C_CONTIGUOUS"] and samples.dtype == self.sample_dtype and samples.shape[0] == self.n_channels
        _check(lib().dspi_eq_process_host(self._h, samples.ctypes.data, samples.shape[1]))

    def sync(self):
        _check(lib().dspi_eq_sync(self._h))

    @property
    def stream(self):
        return lib().dspi_eq_stream(self._h)

    @property
    def launch_count(self):
        return int(lib().dspi_eq_launch_count(self._h))

    def set_params_device(self, recipes, fs, ch0=0):
        """``dsp_compute_coefficients`` for every band of ``recipes`` (EQ_PARAM [n, 12]) on the GPU; returns the clamped recipes."""
        r = np.ascontiguousarray(recipes, L.EQ_PARAM).copy()
        assert r.ndim == 2 and r.shape[1] == L.MAX_BANDS
        _check(lib().dspi_eq_set_params_device(self._h, int(ch0), int(r.shape[0]), r.ctypes.data_as(C.c_void_p), C.c_float(fs)))
        return r

    def kernel_info(self):
        """Which kernel the next process call runs (triggers a pending run-time specialisation)."""
        buf = C.create_string_buffer(320)
        _check(lib().dspi_eq_kernel_info(self._h, buf, len(buf)))
        return buf.value.decode()


class EqGroup:
    """The EQ engine over several GPUs of one box, one process (``dspi_eqx_*``): contiguous channel shards, no exchange."""

    def __init__(self, arith, n_channels, devices, n_bands=L.NUM_BANDS):
        self.arith = ARITH[arith] if isinstance(arith, str) else int(arith)
        self.q28 = self.arith == ARITH["q28"]
        self.n_channels, self.devices = int(n_channels), list(devices)
        d = _EqxDesc(self.arith, self.n_channels, int(n_bands), len(self.devices), (C.c_int32 * 8)(*(self.devices + [0] * (8 - len(self.devices)))), 0)
        self._h = C.c_void_p()
        _check(lib().dspi_eqx_create(C.byref(self._h), C.byref(d)))

    def close(self):
        if self._h:
            lib().dspi_eqx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def shard_range(self, k):
        lo, hi = C.c_uint32(), C.c_uint32()
        _check(lib().dspi_eqx_shard_range(self.n_channels, len(self.devices), k, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def upload(self, biquads, ch0=0):
        b = np.ascontiguousarray(biquads)
        _check(lib().dspi_eqx_upload_biquads(self._h, int(ch0), int(b.shape[0]), b.ctypes.data))

    def download(self, n=None, ch0=0):
        n = self.n_channels - ch0 if n is None else n
        out = np.zeros((n, L.MAX_BANDS), L.BIQUAD_Q28 if self.q28 else L.BIQUAD_F32)
        _check(lib().dspi_eqx_download_biquads(self._h, int(ch0), int(n), out.ctypes.data))
        return out

    def process_host(self, samples):
        assert samples.flags["C_CONTIGUOUS"] and samples.shape[0] == self.n_channels and samples.dtype.itemsize == 4
        _check(lib().dspi_eqx_process_host(self._h, samples.ctypes.data, int(samples.shape[1])))

    def process_root(self, dev_ptr, T, ld=None):
        _check(lib().dspi_eqx_process_root(self._h, C.c_void_p(int(dev_ptr)), int(T), int(T if ld is None else ld)))

    @property
    def launch_count(self):
        return int(lib().dspi_eqx_launch_count(self._h))


def nccl_unique_id():
    """128-byte NCCL unique id (call on one rank, hand the bytes to the others)."""
    buf = (C.c_uint8 * 128)()
    _check(lib().dspi_nccl_unique_id(buf))
    return bytes(buf)


class ScatterGather:
    """One rank's end of the native NCCL scatter / process / gather pipeline (``dspi_sg_*``) around an :class:`EqEngine`."""

    def __init__(self, engine, device, unique_id, rank, world, root=0):
        self._h = C.c_void_p()
        self.engine = engine
        idb = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        _check(lib().dspi_sg_create(C.byref(self._h), engine._h, int(device), idb, int(rank), int(world), int(root)))

    def process(self, full_ptr, total_channels, T, n_chunks=0):
        _check(lib().dspi_sg_process(self._h, C.c_void_p(int(full_ptr)) if full_ptr else None, int(total_channels), int(T), int(n_chunks)))

    def close(self):
        if self._h:
            lib().dspi_sg_destroy(self._h)
            self._h = C.c_void_p()


def bind_host_to_device(device):
    """NUMA-bind this thread and its future allocations to the device's PCIe node; returns the node or -1."""
    return int(lib().dspi_bind_host_to_device(int(device)))


class ChainEngine:
    """Many independent DSPi device instances, whole signal chain (``dspi_chain_*``)."""
    _PRE = "dspi_chain"

    def __init__(self, arith, n_instances, max_frames, n_bands=L.NUM_BANDS, device=0):
        self.arith = ARITH[arith] if isinstance(arith, str) else int(arith)
        self.n_instances, self.max_frames, self.device = int(n_instances), int(max_frames), int(device)
        self._h = C.c_void_p()
        desc = _ChainDesc(self.arith, self.n_instances, int(n_bands), self.device, self.max_frames)
        _check(lib().dspi_chain_create(C.byref(self._h), C.byref(desc)))

    def close(self):
        if self._h:
            lib().dspi_chain_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, params, inst0=0):
        p = np.ascontiguousarray(params)
        assert p.dtype == L.CHAIN_PARAMS_F32 and p.ndim == 1
        _check(lib().dspi_chain_set_params(self._h, inst0, p.shape[0], p.ctypes.data))

    def upload_biquads(self, biquads, inst0=0):
        b = np.ascontiguousarray(biquads)
        assert b.dtype == L.BIQUAD_F32 and b.shape[1:] == (L.CHAIN_EQ_CHANNELS, L.MAX_BANDS)
        _check(lib().dspi_chain_upload_biquads(self._h, inst0, b.shape[0], b.ctypes.data))

    def set_eq_params_device(self, recipes, fs, inst0=0):
        """EQ_PARAM [n, 11, 12] -> coefficients of all bands computed on the GPU; returns the clamped recipes."""
        r = np.ascontiguousarray(recipes, L.EQ_PARAM).copy()
        assert r.shape[1:] == (L.CHAIN_EQ_CHANNELS, L.MAX_BANDS)
        _check(lib().dspi_chain_set_eq_params_device(self._h, int(inst0), int(r.shape[0]), r.ctypes.data_as(C.c_void_p), C.c_float(fs)))
        return r

    def download_biquads(self, n=None, inst0=0):
        n = self.n_instances - inst0 if n is None else n
        out = np.zeros((n, L.CHAIN_EQ_CHANNELS, L.MAX_BANDS), L.BIQUAD_F32)
        _check(lib().dspi_chain_download_biquads(self._h, inst0, n, out.ctypes.data))
        return out

    def state_export(self):
        """Checkpoint: filter / leveller / delay / modulator state (and coefficients) as one bytes-like blob."""
        fn = getattr(lib(), "dspi_chain_state_size")
        fn.restype = C.c_size_t
        n = int(fn(self._h))
        blob = np.zeros(n, np.uint8)
        _check(getattr(lib(), "dspi_chain_state_export")(self._h, blob.ctypes.data_as(C.c_void_p), C.c_size_t(n)))
        return blob

    def state_import(self, blob):
        b = np.ascontiguousarray(blob, np.uint8)
        _check(getattr(lib(), "dspi_chain_state_import")(self._h, b.ctypes.data_as(C.c_void_p), C.c_size_t(b.size)))

    def sm_partition(self):
        """(SMs reserved for the modulator, SMs for every other stage); (0, 0) without a partition."""
        a, b = C.c_uint32(), C.c_uint32()
        _check(getattr(lib(), self._PRE + "_sm_partition")(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def set_dynamics_device(self, cfgs, fs, inst0=0):
        """DYNAMICS_CONFIG [n]: crossfeed / leveller / loudness coefficients and the host volume generated on the GPU."""
        cf = np.ascontiguousarray(cfgs, L.DYNAMICS_CONFIG)
        _check(getattr(lib(), self._PRE + "_set_dynamics_device")(self._h, int(inst0), int(cf.shape[0]), cf.ctypes.data_as(C.c_void_p), C.c_float(fs)))

    def set_preset_mute(self, states, fs, inst0=0, n=None):
        """Envelope mode for instances [inst0, inst0+n): ``states`` PRESET_MUTE [n], or None to leave envelope mode."""
        if states is None:
            n = self.n_instances - inst0 if n is None else n
            _check(getattr(lib(), self._PRE + "_set_preset_mute")(self._h, int(inst0), int(n), None, int(fs)))
            return
        st = np.ascontiguousarray(states, L.PRESET_MUTE)
        _check(getattr(lib(), self._PRE + "_set_preset_mute")(self._h, int(inst0), int(st.shape[0]), st.ctypes.data_as(C.c_void_p), int(fs)))

    def get_preset_mute(self, n=None, inst0=0):
        n = self.n_instances - inst0 if n is None else n
        out = np.zeros(n, L.PRESET_MUTE)
        _check(getattr(lib(), self._PRE + "_get_preset_mute")(self._h, int(inst0), int(n), out.ctypes.data_as(C.c_void_p)))
        return out

    def reset_state(self):
        _check(lib().dspi_chain_reset_state(self._h))

    def process_host(self, pcm, bit_depth, n_packets, frames_per_packet, want_spdif=True, want_pdm=True, want_status=True):
        """``pcm``: uint8 [n_instances, n_frames * bytes_per_frame].  Returns (spdif, pdm, status)."""
        F = n_packets * frames_per_packet
        pcm = np.ascontiguousarray(pcm)
        assert pcm.dtype == np.uint8 and pcm.shape == (self.n_instances, F * (6 if bit_depth == 24 else 4))
        spdif = np.zeros((self.n_instances, 4, F, 2), np.int32) if want_spdif else None
        pdm = np.zeros((self.n_instances, F, 8), np.uint32) if want_pdm else None
        status = np.zeros(self.n_instances, L.STATUS) if want_status else None
        _check(lib().dspi_chain_process_host(self._h, pcm.ctypes.data, bit_depth, n_packets, frames_per_packet,
                                             spdif.ctypes.data if want_spdif else None, pdm.ctypes.data if want_pdm else None,
                                             status.ctypes.data if want_status else None))
        return spdif, pdm, status

    def process_device(self, pcm_ptr, bit_depth, n_packets, frames_per_packet, spdif_ptr=0, pdm_ptr=0, status_ptr=0):
        _check(lib().dspi_chain_process_device(self._h, C.c_void_p(int(pcm_ptr)), bit_depth, n_packets, frames_per_packet,
                                               C.c_void_p(int(spdif_ptr)) if spdif_ptr else None,
                                               C.c_void_p(int(pdm_ptr)) if pdm_ptr else None,
                                               C.c_void_p(int(status_ptr)) if status_ptr else None))

    def sync(self):
        _check(lib().dspi_chain_sync(self._h))

    @property
    def stream(self):
        return lib().dspi_chain_stream(self._h)

    @property
    def launch_count(self):
        return int(lib().dspi_chain_launch_count(self._h))


def delay_samples(delay_ms, fs, is_last=False):
    return int(lib().dspi_delay_samples(delay_ms, fs, 1 if is_last else 0))


class _XfeedCfg(C.Structure):
    _fields_ = [("enabled", C.c_uint8), ("itd_enabled", C.c_uint8), ("preset", C.c_uint8), ("custom_fc", C.c_float), ("custom_feed_db", C.c_float)]


class _LevCfg(C.Structure):
    _fields_ = [("enabled", C.c_uint8), ("amount", C.c_float), ("speed", C.c_uint8), ("max_gain_db", C.c_float),
                ("lookahead", C.c_uint8), ("gate_threshold_db", C.c_float)]


def crossfeed_coefficients(fs, enabled=True, itd=True, preset=0, custom_fc=700.0, custom_feed_db=4.5):
    """``crossfeed_compute_coefficients`` -> XFEED_F32 record (state cleared)."""
    st = np.zeros(1, L.XFEED_F32)
    cfg = _XfeedCfg(int(enabled), int(itd), int(preset), custom_fc, custom_feed_db)
    lib().dspi_crossfeed_compute_coefficients_f32(st.ctypes.data, C.byref(cfg), fs)
    return st[0]


def leveller_coefficients(fs, amount=50.0, speed=0, max_gain_db=15.0, gate_db=-96.0):
    out = np.zeros(1, L.LEV_COEFFS)
    cfg = _LevCfg(1, amount, int(speed), max_gain_db, 1, gate_db)
    lib().dspi_leveller_compute_coefficients(out.ctypes.data, C.byref(cfg), fs)
    return out[0]


def loudness_table(fs, ref_spl=83.0, intensity_pct=100.0):
    t = np.zeros((L.LOUD_STEPS, 2), L.LOUD_F32)
    lib().dspi_loudness_compute_table_f32(t.ctypes.data, ref_spl, intensity_pct, fs)
    return t


def bulk_state_defaults(platform=L.PLATFORM_RP2350):
    """Power-on values of the globals ``bulk_params_apply`` edits (one ``BULK_STATE`` record)."""
    st = np.zeros(1, L.BULK_STATE)
    lib().dspi_bulk_state_defaults(st.ctypes.data_as(C.c_void_p), int(platform))
    return st


def bulk_params_apply(wire, state, exact_db=False):
    """``bulk_params_apply`` (bulk_params.c:178-377) on ``state`` (in place).  Returns the firmware's code: 0, -1 .. -4."""
    w = np.ascontiguousarray(wire).view(np.uint8).reshape(-1)
    assert w.size == L.WIRE_BULK.itemsize and state.dtype == L.BULK_STATE
    return int(lib().dspi_bulk_params_apply(w.ctypes.data_as(C.c_void_p), state.ctypes.data_as(C.c_void_p), int(bool(exact_db))))


def bulk_params_collect(state):
    """``bulk_params_collect`` (bulk_params.c:62-172): one ``WIRE_BULK`` record."""
    out = np.zeros(1, L.WIRE_BULK)
    lib().dspi_bulk_params_collect(state.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return out


def bulk_state_to_chain(state, fs, host_volume_8_8=0, host_mute=False, biquads=None):
    """What the main loop derives after an apply: (chain params [1], biquads [1, roles, 12]) for the state's platform."""
    q28 = int(state["platform"][0]) == L.PLATFORM_RP2040
    P = np.zeros(1, L.CHAIN_PARAMS_Q28 if q28 else L.CHAIN_PARAMS_F32)
    roles = L.CHAINQ_EQ_CHANNELS if q28 else L.CHAIN_EQ_CHANNELS
    bq = np.zeros((1, roles, L.MAX_BANDS), L.BIQUAD_Q28 if q28 else L.BIQUAD_F32) if biquads is None else biquads
    fn = lib().dspi_bulk_state_to_chain_q28 if q28 else lib().dspi_bulk_state_to_chain_f32
    _check(fn(state.ctypes.data_as(C.c_void_p), C.c_float(fs), C.c_int16(int(host_volume_8_8)), int(bool(host_mute)),
              P.ctypes.data_as(C.c_void_p), bq.ctypes.data_as(C.c_void_p)))
    return P, bq


def preset_slot_size(platform=L.PLATFORM_RP2350):
    lib().dspi_preset_slot_size.restype = C.c_size_t
    return int(lib().dspi_preset_slot_size(int(platform)))


def crc32(data):
    lib().dspi_crc32.restype = C.c_uint32
    b = bytes(data)
    return int(lib().dspi_crc32(b, C.c_size_t(len(b))))


def preset_slot_collect(state, slot_index):
    """``collect_live_state`` (flash_storage.c:464-556): the slot image as bytes."""
    n = preset_slot_size(int(state["platform"][0]))
    out = np.zeros(n, np.uint8)
    _check(lib().dspi_preset_slot_collect(state.ctypes.data_as(C.c_void_p), int(slot_index), out.ctypes.data_as(C.c_void_p), C.c_size_t(n)))
    return out


def preset_slot_apply(image, slot_index, state, master_volume_mode=0, dir_master_volume_db=0.0):
    """The state part of ``preset_load`` on ``state`` (in place).  Returns 0 or 3 (PRESET_ERR_CRC)."""
    img = np.ascontiguousarray(image, np.uint8)
    return int(lib().dspi_preset_slot_apply(img.ctypes.data_as(C.c_void_p), C.c_size_t(img.size), int(slot_index), int(master_volume_mode),
                                           C.c_float(dir_master_volume_db), state.ctypes.data_as(C.c_void_p)))


SPDIF_CHANNEL_STATUS = bytes([0x04, 0x00, 0x00, 0x00, 0x0B])       # audio_spdif.c:82-88 (byte 3 = sample-rate code, set at run time)


def spdif_lookup_table():
    """The reference's 256-entry biphase-mark table (``audio_spdif.c:141-153``)."""
    t = np.zeros(256, np.uint32)
    lib().dspi_spdif_lookup_table(t.ctypes.data_as(C.c_void_p))
    return t


def spdif_encode_device(d_words, n_streams, frames, d_subframes, block_pos0=0, channel_status=SPDIF_CHANNEL_STATUS, device=0, stream=None):
    """Device pointers in/out: ``[n_streams][frames][2]`` int32 words -> ``[n_streams][frames][2][2]`` uint32 {l, h}."""
    cs = (C.c_uint8 * 5)(*channel_status)
    _check(lib().dspi_spdif_encode_device(int(device), C.c_void_p(int(d_words)), C.c_uint64(int(n_streams)), int(frames), int(block_pos0), cs,
                                          C.c_void_p(int(d_subframes)), C.c_void_p(int(stream) if stream else None)))


def spdif_encode_host(words, block_pos0=0, channel_status=SPDIF_CHANNEL_STATUS, device=0):
    """``words`` int32 ``[n_streams, frames, 2]`` (host) -> uint32 ``[n_streams, frames, 2, 2]`` ({l, h} per subframe)."""
    w = np.ascontiguousarray(words, np.int32)
    assert w.ndim == 3 and w.shape[2] == 2
    out = np.empty(w.shape + (2,), np.uint32)
    cs = (C.c_uint8 * 5)(*channel_status)
    _check(lib().dspi_spdif_encode_host(int(device), w.ctypes.data_as(C.c_void_p), C.c_uint64(w.shape[0]), int(w.shape[1]), int(block_pos0), cs,
                                        out.ctypes.data_as(C.c_void_p)))
    return out


def host_volume(volume_8_8):
    """``audio_set_volume``: returns (vol_mul as int16, loudness table row)."""
    idx = C.c_uint8()
    v = lib().dspi_host_volume(int(volume_8_8), C.byref(idx))
    return int(v), int(idx.value)


def preamp(db):
    """``update_preamp``: (linear float, Q28 int)."""
    lin, q = C.c_float(), C.c_int32()
    if lib().dspi_preamp(float(db), C.byref(lin), C.byref(q)) != 0:
        raise DspiError("preamp: NaN / Inf rejected")
    return lin.value, q.value


def master_volume(db):
    """``update_master_volume``: (linear float, Q15 int)."""
    lin, q = C.c_float(), C.c_int32()
    if lib().dspi_master_volume(float(db), C.byref(lin), C.byref(q)) != 0:
        raise DspiError("master volume: NaN / Inf rejected")
    return lin.value, q.value


class ChainEngineQ28:
    """Many independent RP2040-shape instances (2 in -> 5 out), Q28 arithmetic (``dspi_chainq_*``)."""
    _PRE = "dspi_chainq"

    def __init__(self, n_instances, max_frames, n_bands=L.NUM_BANDS, device=0):
        self.n_instances, self.max_frames, self.device = int(n_instances), int(max_frames), int(device)
        self._h = C.c_void_p()
        desc = _ChainDesc(ARITH_Q28, self.n_instances, int(n_bands), self.device, self.max_frames)
        _check(lib().dspi_chainq_create(C.byref(self._h), C.byref(desc)))

    def close(self):
        if self._h:
            lib().dspi_chainq_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, params, inst0=0):
        p = np.ascontiguousarray(params)
        assert p.dtype == L.CHAIN_PARAMS_Q28 and p.ndim == 1
        _check(lib().dspi_chainq_set_params(self._h, inst0, p.shape[0], p.ctypes.data))

    def upload_biquads(self, biquads, inst0=0):
        b = np.ascontiguousarray(biquads)
        assert b.dtype == L.BIQUAD_Q28 and b.shape[1:] == (L.CHAINQ_EQ_CHANNELS, L.MAX_BANDS)
        _check(lib().dspi_chainq_upload_biquads(self._h, inst0, b.shape[0], b.ctypes.data))

    def set_eq_params_device(self, recipes, fs, inst0=0):
        r = np.ascontiguousarray(recipes, L.EQ_PARAM).copy()
        assert r.shape[1:] == (L.CHAINQ_EQ_CHANNELS, L.MAX_BANDS)
        _check(lib().dspi_chainq_set_eq_params_device(self._h, int(inst0), int(r.shape[0]), r.ctypes.data_as(C.c_void_p), C.c_float(fs)))
        return r

    def download_biquads(self, n=None, inst0=0):
        n = self.n_instances - inst0 if n is None else n
        out = np.zeros((n, L.CHAINQ_EQ_CHANNELS, L.MAX_BANDS), L.BIQUAD_Q28)
        _check(lib().dspi_chainq_download_biquads(self._h, inst0, n, out.ctypes.data))
        return out

    def state_export(self):
        """Checkpoint: filter / leveller / delay / modulator state (and coefficients) as one bytes-like blob."""
        fn = getattr(lib(), "dspi_chainq_state_size")
        fn.restype = C.c_size_t
        n = int(fn(self._h))
        blob = np.zeros(n, np.uint8)
        _check(getattr(lib(), "dspi_chainq_state_export")(self._h, blob.ctypes.data_as(C.c_void_p), C.c_size_t(n)))
        return blob

    def state_import(self, blob):
        b = np.ascontiguousarray(blob, np.uint8)
        _check(getattr(lib(), "dspi_chainq_state_import")(self._h, b.ctypes.data_as(C.c_void_p), C.c_size_t(b.size)))

    def sm_partition(self):
        """(SMs reserved for the modulator, SMs for every other stage); (0, 0) without a partition."""
        a, b = C.c_uint32(), C.c_uint32()
        _check(getattr(lib(), self._PRE + "_sm_partition")(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def set_dynamics_device(self, cfgs, fs, inst0=0):
        """DYNAMICS_CONFIG [n]: crossfeed / leveller / loudness coefficients and the host volume generated on the GPU."""
        cf = np.ascontiguousarray(cfgs, L.DYNAMICS_CONFIG)
        _check(getattr(lib(), self._PRE + "_set_dynamics_device")(self._h, int(inst0), int(cf.shape[0]), cf.ctypes.data_as(C.c_void_p), C.c_float(fs)))

    def set_preset_mute(self, states, fs, inst0=0, n=None):
        """Envelope mode for instances [inst0, inst0+n): ``states`` PRESET_MUTE [n], or None to leave envelope mode."""
        if states is None:
            n = self.n_instances - inst0 if n is None else n
            _check(getattr(lib(), self._PRE + "_set_preset_mute")(self._h, int(inst0), int(n), None, int(fs)))
            return
        st = np.ascontiguousarray(states, L.PRESET_MUTE)
        _check(getattr(lib(), self._PRE + "_set_preset_mute")(self._h, int(inst0), int(st.shape[0]), st.ctypes.data_as(C.c_void_p), int(fs)))

    def get_preset_mute(self, n=None, inst0=0):
        n = self.n_instances - inst0 if n is None else n
        out = np.zeros(n, L.PRESET_MUTE)
        _check(getattr(lib(), self._PRE + "_get_preset_mute")(self._h, int(inst0), int(n), out.ctypes.data_as(C.c_void_p)))
        return out

    def reset_state(self):
        _check(lib().dspi_chainq_reset_state(self._h))

    def process_host(self, pcm, bit_depth, n_packets, frames_per_packet):
        F = n_packets * frames_per_packet
        pcm = np.ascontiguousarray(pcm)
        assert pcm.dtype == np.uint8 and pcm.shape == (self.n_instances, F * (6 if bit_depth == 24 else 4))
        spdif = np.zeros((self.n_instances, 2, F, 2), np.int32)
        pdm = np.zeros((self.n_instances, F, 8), np.uint32)
        status = np.zeros(self.n_instances, L.STATUS_Q28)
        _check(lib().dspi_chainq_process_host(self._h, pcm.ctypes.data, bit_depth, n_packets, frames_per_packet,
                                              spdif.ctypes.data, pdm.ctypes.data, status.ctypes.data))
        return spdif, pdm, status

    def process_device(self, pcm_ptr, bit_depth, n_packets, frames_per_packet, spdif_ptr=0, pdm_ptr=0, status_ptr=0):
        _check(lib().dspi_chainq_process_device(self._h, C.c_void_p(int(pcm_ptr)), bit_depth, n_packets, frames_per_packet,
                                                C.c_void_p(int(spdif_ptr)) if spdif_ptr else None,
                                                C.c_void_p(int(pdm_ptr)) if pdm_ptr else None,
                                                C.c_void_p(int(status_ptr)) if status_ptr else None))

    def sync(self):
        _check(lib().dspi_chainq_sync(self._h))

    @property
    def stream(self):
        return lib().dspi_chainq_stream(self._h)

    @property
    def launch_count(self):
        return int(lib().dspi_chainq_launch_count(self._h))


def crossfeed_coefficients_q28(fs, enabled=True, itd=True, preset=0, custom_fc=700.0, custom_feed_db=4.5):
    st = np.zeros(1, L.XFEED_Q28)
    cfg = _XfeedCfg(int(enabled), int(itd), int(preset), custom_fc, custom_feed_db)
    lib().dspi_crossfeed_compute_coefficients_q28(st.ctypes.data, C.byref(cfg), fs)
    return st[0]


def loudness_table_q28(fs, ref_spl=83.0, intensity_pct=100.0):
    t = np.zeros((L.LOUD_STEPS, 2), L.LOUD_Q28)
    lib().dspi_loudness_compute_table_q28(t.ctypes.data, ref_spl, intensity_pct, fs)
    return t
