"""ctypes access to the CPU oracle (``oracle/liborc.so``) and, when present, to
the reference sources compiled on the host (``oracle/_ref/*.so``).

Test infrastructure: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline legs import this module.  The product package
(``dspi_b200``) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

_vp = C.c_void_p
_u32 = C.c_uint32


def _ptr(a):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_vp)


def build_oracle(with_ref=True):
    """Compile the restatement (always) and, if /root/reference exists, oracle/_ref."""
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "all"])
    if with_ref and os.path.isdir("/root/reference/firmware/DSPi"):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "ref"])


class Oracle:
    """The restatement.  ``flavour``: 'f32f' (fused), 'f32s' (strict) or 'q28'."""

    def __init__(self):
        path = os.path.join(ORACLE_DIR, "liborc.so")
        if not os.path.exists(path):
            build_oracle(with_ref=False)
        self.lib = L = C.CDLL(path)
        L.orc_sizeof.restype = C.c_size_t
        L.orc_mul_q28.restype = C.c_int32
        L.orc_mul_q28.argtypes = [C.c_int32, C.c_int32]
        L.orc_mul_q15.restype = C.c_int32
        L.orc_mul_q15.argtypes = [C.c_int32, C.c_int32]
        L.orc_f2i_sat.restype = C.c_int32
        L.orc_f2i_sat.argtypes = [C.c_float]
        L.orc_delay_samples.restype = C.c_int32
        L.orc_delay_samples.argtypes = [C.c_float, C.c_float, C.c_int, C.c_int32]
        L.orc_host_vol_mul.restype = C.c_int16
        L.orc_host_vol_mul.argtypes = [C.c_int16, _vp]
        L.orc_eq_many_mt.restype = C.c_double
        L.orc_eq_many_mt.argtypes = [C.c_int, _vp, _vp, _u32, _u32, _u32, _u32, _u32]
        for fl in ("f32s", "f32f", "q28"):
            getattr(L, f"orc_{fl}_eq_many").argtypes = [_vp, _vp, _u32, _u32, _u32, _u32]
            getattr(L, f"orc_{fl}_eq_block").argtypes = [_vp, _vp, _u32, _u32]
            getattr(L, f"orc_{fl}_xfeed").argtypes = [_vp, _vp, _vp, _u32]
            getattr(L, f"orc_{fl}_leveller").argtypes = [_vp, _vp, C.c_int, _vp, _vp, _u32]
            f = getattr(L, f"orc_{fl}_chain_packet")
            f.restype = _u32
            f.argtypes = [_vp, _vp, _u32, _u32, _vp, _u32, _vp]
        L.orc_spdif_lookup_init.argtypes = [_vp]
        L.orc_spdif_update_subframe.argtypes = [_vp, _vp, _vp, C.c_int32]
        L.orc_spdif_encode.argtypes = [_vp, _vp, _u32, _u32, _vp, _vp]
        L.orc_eq_coeffs_f32.argtypes = [_vp, _vp, C.c_float]
        L.orc_eq_coeffs_q28.argtypes = [_vp, _vp, C.c_float]
        L.orc_xfeed_coeffs_f32.argtypes = [_vp, _vp, C.c_float]
        L.orc_xfeed_coeffs_q28.argtypes = [_vp, _vp, C.c_float]
        L.orc_lev_coeffs_compute.argtypes = [_vp, _vp, C.c_float]
        L.orc_loud_table_f32.argtypes = [_vp, C.c_float, C.c_float, C.c_float]
        L.orc_loud_table_q28.argtypes = [_vp, C.c_float, C.c_float, C.c_float]
        L.orc_pdm_modulate.argtypes = [_vp, C.c_int32, _vp]
        L.orc_mute_envelope.restype = C.c_float
        L.orc_mute_envelope.argtypes = [_vp, _vp, _vp, _u32, _u32]

    # -- knobs -------------------------------------------------------------
    def set_x86_cvt(self, on):
        self.lib.orc_set_x86_cvt(int(on))

    def set_libm_f64(self, on):
        self.lib.orc_set_libm_f64(int(on))

    # -- EQ cascades ---------------------------------------------------------
    def eq_many(self, flavour, bq, samples, nbands=10, packet=0):
        """bq: [C, MAX_BANDS] structured; samples: [C, T]; both updated in place."""
        Cn, T = samples.shape
        getattr(self.lib, f"orc_{flavour}_eq_many")(_ptr(bq), _ptr(samples), Cn, T, nbands, packet)

    def eq_many_mt(self, flavour, bq, samples, nbands, packet, nthreads):
        kind = {"f32f": 0, "f32s": 1, "q28": 2}[flavour]
        Cn, T = samples.shape
        return self.lib.orc_eq_many_mt(kind, _ptr(bq), _ptr(samples), Cn, T, nbands, packet, nthreads)

    def xfeed(self, flavour, st, l, r):
        getattr(self.lib, f"orc_{flavour}_xfeed")(_ptr(st), _ptr(l), _ptr(r), l.shape[0])

    def leveller(self, flavour, st, coeffs, lookahead, l, r):
        getattr(self.lib, f"orc_{flavour}_leveller")(_ptr(st), _ptr(coeffs), int(lookahead), _ptr(l), _ptr(r), l.shape[0])

    def pdm(self, st, samples_q28):
        """st: PDM_STATE[1]; returns uint32 [N, 8]."""
        out = np.zeros((len(samples_q28), 8), np.uint32)
        for i, s in enumerate(samples_q28):
            self.lib.orc_pdm_modulate(_ptr(st), int(s), out[i].ctypes.data_as(_vp))
        return out

    # -- parameters ----------------------------------------------------------
    def spdif_table(self):
        t = np.zeros(256, np.uint32)
        self.lib.orc_spdif_lookup_init(t.ctypes.data)
        return t

    def spdif_update(self, table, l, h, sample):
        a, b = C.c_uint32(l), C.c_uint32(h)
        self.lib.orc_spdif_update_subframe(table.ctypes.data, C.addressof(a), C.addressof(b), int(np.int32(sample)))
        return int(a.value), int(b.value)

    def spdif_encode(self, words, pos0=0, cs=bytes([0x04, 0, 0, 0, 0x0B])):
        """``words`` int32 [n_streams, frames, 2] -> uint32 [n_streams, frames, 2, 2]; every stream starts at pos0."""
        w = np.ascontiguousarray(words, np.int32)
        out = np.empty(w.shape + (2,), np.uint32)
        table = self.spdif_table()
        csb = np.frombuffer(bytes(cs), np.uint8).copy()
        for s_ in range(w.shape[0]):
            self.lib.orc_spdif_encode(table.ctypes.data, w[s_].ctypes.data, int(w.shape[1]), int(pos0), csb.ctypes.data, out[s_].ctypes.data)
        return out

    def eq_coeffs(self, q28, params, bq, fs):
        """params: EQ_PARAM[N] (mutated: clamps written back); bq: matching biquads."""
        fn = self.lib.orc_eq_coeffs_q28 if q28 else self.lib.orc_eq_coeffs_f32
        pp = params.reshape(-1)
        bb = bq.reshape(-1)
        base_p, base_b = pp.ctypes.data, bb.ctypes.data
        sp, sb = pp.dtype.itemsize, bb.dtype.itemsize
        for i in range(pp.shape[0]):
            fn(base_p + i * sp, base_b + i * sb, fs)


class RefSpdif:
    """The reference's own ``spdif_update_subframe`` compiled from its header (oracle/ref_spdif_shim.c)."""

    PATH = os.path.join(ORACLE_DIR, "_ref", "libdspi_ref_spdif.so")

    @staticmethod
    def available():
        return os.path.exists(RefSpdif.PATH)

    def __init__(self, table):
        self.lib = L = C.CDLL(RefSpdif.PATH)
        L.ref_spdif_set_lookup.argtypes = [_vp]
        L.ref_spdif_update_subframe.argtypes = [_vp, C.c_int32]
        L.ref_spdif_copy_s32.argtypes = [_vp, _vp, _u32]
        t = np.ascontiguousarray(table, np.uint32)
        L.ref_spdif_set_lookup(t.ctypes.data)

    def update(self, l, h, sample):
        lh = np.array([l, h], np.uint32)
        self.lib.ref_spdif_update_subframe(lh.ctypes.data, int(np.int32(sample)))
        return int(lh[0]), int(lh[1])

    def copy_s32(self, subframes, words):
        """In place: ``subframes`` uint32 [n, 2, 2] updated with ``words`` int32 [n, 2]."""
        self.lib.ref_spdif_copy_s32(subframes.ctypes.data, words.ctypes.data, int(words.shape[0]))


class Ref:
    """One of the three host builds of the UNMODIFIED reference sources."""

    NAMES = {"f32s": "libdspi_ref_f32_strict.so", "f32f": "libdspi_ref_f32_fused.so", "q28": "libdspi_ref_q28.so"}

    @staticmethod
    def available():
        return all(os.path.exists(os.path.join(ORACLE_DIR, "_ref", n)) for n in Ref.NAMES.values())

    def __init__(self, flavour):
        self.flavour = flavour
        self.lib = L = C.CDLL(os.path.join(ORACLE_DIR, "_ref", Ref.NAMES[flavour]))
        L.ref_sizeof.restype = C.c_size_t
        L.ref_sizeof.argtypes = [C.c_int]
        L.ref_offsetof.restype = C.c_size_t
        L.ref_offsetof.argtypes = [C.c_int]
        L.ref_eq_coeffs.argtypes = [_vp, _vp, C.c_float]
        L.ref_eq_block.argtypes = [_vp, _vp, _u32]
        L.ref_eq_many.argtypes = [_vp, _vp, _u32, _u32, _u32]
        L.ref_eq_many_mt.restype = C.c_double
        L.ref_eq_many_mt.argtypes = [_vp, _vp, _u32, _u32, _u32, _u32]
        L.ref_xfeed_coeffs.argtypes = [_vp, C.c_uint8, C.c_uint8, C.c_uint8, C.c_float, C.c_float, C.c_float]
        L.ref_xfeed.argtypes = [_vp, _vp, _vp, _u32]
        L.ref_lev_coeffs.argtypes = [_vp, C.c_float, C.c_uint8, C.c_float, C.c_float, C.c_float]
        L.ref_lev_reset.argtypes = [_vp]
        L.ref_leveller.argtypes = [_vp, _vp, C.c_int, _vp, _vp, _u32]
        L.ref_loud_table.argtypes = [_vp, C.c_float, C.c_float, C.c_float]
        L.ref_delay_samples.restype = C.c_int32
        L.ref_delay_samples.argtypes = [C.c_float, C.c_float, C.c_int]
        if flavour == "q28":
            L.ref_mul_q28.restype = C.c_int32
            L.ref_mul_q28.argtypes = [C.c_int32, C.c_int32]
            L.ref_mul_q15.restype = C.c_int32
            L.ref_mul_q15.argtypes = [C.c_int32, C.c_int32]

    def set_nbands(self, n):
        self.lib.ref_set_nbands(n)

    def eq_many(self, bq, samples, nbands=10, packet=0):
        self.set_nbands(nbands)
        Cn, T = samples.shape
        self.lib.ref_eq_many(_ptr(bq), _ptr(samples), Cn, T, packet)

    def eq_many_mt(self, bq, samples, nbands, packet, nthreads):
        self.set_nbands(nbands)
        Cn, T = samples.shape
        return self.lib.ref_eq_many_mt(_ptr(bq), _ptr(samples), Cn, T, packet, nthreads)

    def eq_coeffs(self, params, bq, fs):
        pp = params.reshape(-1)
        bb = bq.reshape(-1)
        base_p, base_b = pp.ctypes.data, bb.ctypes.data
        sp, sb = pp.dtype.itemsize, bb.dtype.itemsize
        for i in range(pp.shape[0]):
            self.lib.ref_eq_coeffs(base_p + i * sp, base_b + i * sb, fs)

    def xfeed(self, st, l, r):
        self.lib.ref_xfeed(_ptr(st), _ptr(l), _ptr(r), l.shape[0])

    def leveller(self, st, coeffs, lookahead, l, r):
        self.lib.ref_leveller(_ptr(st), _ptr(coeffs), int(lookahead), _ptr(l), _ptr(r), l.shape[0])


# ---- whole-instance records of the oracle (oracle/dspi_oracle.h) -----------------------------
class OrcCrosspoint(C.Structure):
    _pack_ = 1
    _fields_ = [("enabled", C.c_uint8), ("phase_invert", C.c_uint8), ("reserved", C.c_uint8 * 2),
                ("gain_db", C.c_float), ("gain_linear", C.c_float)]


class OrcOutput(C.Structure):
    _pack_ = 1
    _fields_ = [("enabled", C.c_uint8), ("mute", C.c_uint8), ("reserved", C.c_uint8 * 2), ("gain_db", C.c_float),
                ("gain_linear", C.c_float), ("delay_ms", C.c_float), ("delay_samples", C.c_int32)]


class OrcBiquadF32(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("b0", "b1", "b2", "a1", "a2", "s1", "s2", "sva1", "sva2", "sva3", "svm0", "svm1", "svm2",
                                         "svic1eq", "svic2eq")] + [("svf_type", C.c_uint32), ("use_svf", C.c_uint8), ("bypass", C.c_uint8)]


class OrcLoudF32(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("sva1", "sva2", "sva3", "svm0", "svm1", "svm2")] + [("bypass", C.c_uint8)]


class OrcSvfState(C.Structure):
    _fields_ = [("ic1eq", C.c_float), ("ic2eq", C.c_float)]


class OrcXfeedF32(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("lp_a0", "lp_b1", "lp_state_L", "lp_state_R", "ap_a", "ap_state_L", "ap_state_R")]


class OrcLevCoeffs(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("alpha_rms", "alpha_attack", "alpha_release", "threshold_db", "ratio", "knee_width_db",
                                         "makeup_db", "gate_threshold_db", "max_gain_db")]


class OrcLevStateF32(C.Structure):
    _fields_ = [("env_sq_l", C.c_float), ("env_sq_r", C.c_float), ("gain_smooth_db", C.c_float), ("gain_linear", C.c_float),
                ("gain_prev_linear", C.c_float), ("lookahead_buf", (C.c_float * 480) * 2), ("la_write_idx", C.c_uint32)]


class OrcPdm(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("err1", "err2", "x1", "x2", "y1", "y2", "err_acc")] + [("rng", C.c_uint32), ("fade_in_pos", C.c_uint32)]


class OrcChainF32(C.Structure):
    _fields_ = [("n_out", C.c_uint32), ("n_bands", C.c_uint32), ("max_delay", C.c_uint32),
                ("bypass_master_eq", C.c_uint8), ("loudness_on", C.c_uint8), ("crossfeed_on", C.c_uint8), ("leveller_on", C.c_uint8),
                ("host_mute", C.c_uint8), ("any_delay_active", C.c_uint8), ("lev_lookahead", C.c_uint8), ("pad0", C.c_uint8),
                ("host_vol_mul", C.c_int16), ("pad1", C.c_int16), ("preset_mute_gain", C.c_float), ("master_volume_linear", C.c_float),
                ("preamp_linear", C.c_float * 2), ("xp", (OrcCrosspoint * 9) * 2), ("out", OrcOutput * 9), ("delay_samples", C.c_int32 * 9),
                ("channel_bypassed", C.c_uint8 * 11), ("pad2", C.c_uint8), ("filters", (OrcBiquadF32 * 12) * 11), ("loud", OrcLoudF32 * 2),
                ("loud_state", (OrcSvfState * 2) * 2), ("xfeed", OrcXfeedF32), ("levc", OrcLevCoeffs), ("levs", OrcLevStateF32),
                ("delay_lines", (C.c_float * 4096) * 9), ("delay_widx", C.c_uint32), ("pdm", OrcPdm), ("peaks", C.c_uint16 * 11),
                ("clip_flags", C.c_uint16),
                ("mute_env_on", C.c_uint8), ("preset_loading", C.c_uint8), ("pad3", C.c_uint8 * 2), ("preset_mute_counter", C.c_uint32),
                ("preset_mute_smooth_gain", C.c_float), ("sample_rate_hz", C.c_uint32)]


def make_orc_chain(oracle, params, biquads):
    """orc_chain_f32 for one instance from a CHAIN_PARAMS_F32 record and its [11, 12] biquads."""
    assert C.sizeof(OrcChainF32) == oracle.lib.orc_sizeof(2)
    c = OrcChainF32()
    c.n_out, c.n_bands, c.max_delay = 9, 10, 4096
    c.bypass_master_eq = int(params["bypass_master_eq"])
    c.loudness_on = int(params["loudness_enabled"])
    c.crossfeed_on = int(params["crossfeed_enabled"])
    c.leveller_on = int(params["leveller_enabled"])
    c.host_mute = int(params["host_mute"])
    c.lev_lookahead = int(params["leveller_lookahead"])
    c.host_vol_mul = int(params["host_vol_mul"])
    c.preset_mute_gain = float(params["preset_mute_gain"])
    c.master_volume_linear = float(params["master_volume_linear"])
    c.preamp_linear[0], c.preamp_linear[1] = float(params["preamp_linear"][0]), float(params["preamp_linear"][1])
    m = params["matrix"]
    any_delay = False
    for o in range(9):
        for i in range(2):
            x = m["crosspoints"][i, o]
            c.xp[i][o].enabled, c.xp[i][o].phase_invert = int(x["enabled"]), int(x["phase_invert"])
            c.xp[i][o].gain_db, c.xp[i][o].gain_linear = float(x["gain_db"]), float(x["gain_linear"])
        oc = m["outputs"][o]
        c.out[o].enabled, c.out[o].mute = int(oc["enabled"]), int(oc["mute"])
        c.out[o].gain_db, c.out[o].gain_linear = float(oc["gain_db"]), float(oc["gain_linear"])
        c.out[o].delay_ms, c.out[o].delay_samples = float(oc["delay_ms"]), int(oc["delay_samples"])
        c.delay_samples[o] = int(oc["delay_samples"])
        any_delay = any_delay or int(oc["delay_samples"]) > 0
    c.any_delay_active = 1 if any_delay else 0
    C.memmove(C.addressof(c.filters), np.ascontiguousarray(biquads).ctypes.data, 11 * 12 * 68)
    for r in range(11):
        c.channel_bypassed[r] = 1 if all(int(biquads[r, b]["bypass"]) for b in range(10)) else 0
    for j in range(2):
        C.memmove(C.addressof(c.loud[j]), params["loudness"][j:j + 1].tobytes(), 28)
    C.memmove(C.addressof(c.xfeed), params["crossfeed"].tobytes(), 28)
    C.memmove(C.addressof(c.levc), params["leveller"].tobytes(), 36)
    c.levs.gain_linear = 1.0
    c.levs.gain_prev_linear = 1.0
    c.pdm.rng = 123456789
    return c


def orc_chain_run(oracle, flavour, chain, pcm_bytes, bit_depth, n_packets, fpp):
    """Runs n_packets packets through one oracle instance; returns (spdif [4, F, 2], pdm [F, 8])."""
    F = n_packets * fpp
    bpf = 6 if bit_depth == 24 else 4
    spdif = np.zeros((4, F, 2), np.int32)
    pdm = np.zeros((F, 8), np.uint32)
    fn = getattr(oracle.lib, f"orc_{flavour}_chain_packet")
    data = np.ascontiguousarray(pcm_bytes)
    for p in range(n_packets):
        fn(C.addressof(chain), data.ctypes.data + p * fpp * bpf, fpp * bpf, bit_depth,
           spdif.ctypes.data + p * fpp * 8, F * 2, pdm.ctypes.data + p * fpp * 32)
    return spdif, pdm


class OrcBiquadQ28(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("b0", "b1", "b2", "a1", "a2", "s1", "s2")] + [("bypass", C.c_uint8)]


class OrcLoudQ28(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("b0", "b1", "b2", "a1", "a2")] + [("bypass", C.c_uint8)]


class OrcXfeedQ28(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("lp_a0", "lp_b1", "lp_state_L", "lp_state_R", "ap_a", "ap_state_L", "ap_state_R")]


class OrcLevStateQ28(C.Structure):
    _fields_ = [("env_sq_l", C.c_int32), ("env_sq_r", C.c_int32), ("gain_smooth_db", C.c_float), ("gain_q28", C.c_int32),
                ("gain_prev_q28", C.c_int32), ("lookahead_buf", (C.c_int32 * 480) * 2), ("la_write_idx", C.c_uint32)]


class OrcChainQ28(C.Structure):
    _fields_ = [("n_out", C.c_uint32), ("n_bands", C.c_uint32), ("max_delay", C.c_uint32),
                ("bypass_master_eq", C.c_uint8), ("loudness_on", C.c_uint8), ("crossfeed_on", C.c_uint8), ("leveller_on", C.c_uint8),
                ("host_mute", C.c_uint8), ("any_delay_active", C.c_uint8), ("lev_lookahead", C.c_uint8), ("pad0", C.c_uint8),
                ("host_vol_mul", C.c_int16), ("pad1", C.c_int16), ("preset_mute_gain", C.c_float), ("master_volume_q15", C.c_int32),
                ("preamp_q28", C.c_int32 * 2), ("xp", (OrcCrosspoint * 9) * 2), ("out", OrcOutput * 9), ("delay_samples", C.c_int32 * 9),
                ("channel_bypassed", C.c_uint8 * 11), ("pad2", C.c_uint8), ("filters", (OrcBiquadQ28 * 12) * 11), ("loud", OrcLoudQ28 * 2),
                ("loud_state", (OrcBiquadQ28 * 2) * 2), ("xfeed", OrcXfeedQ28), ("levc", OrcLevCoeffs), ("levs", OrcLevStateQ28),
                ("delay_lines", (C.c_int32 * 4096) * 9), ("delay_widx", C.c_uint32), ("pdm", OrcPdm), ("peaks", C.c_uint16 * 11),
                ("clip_flags", C.c_uint16),
                ("mute_env_on", C.c_uint8), ("preset_loading", C.c_uint8), ("pad3", C.c_uint8 * 2), ("preset_mute_counter", C.c_uint32),
                ("preset_mute_smooth_gain", C.c_float), ("sample_rate_hz", C.c_uint32)]


def make_orc_chain_q28(oracle, params, biquads):
    """orc_chain_q28 for one instance from a CHAIN_PARAMS_Q28 record and its [7, 12] biquads."""
    assert C.sizeof(OrcChainQ28) == oracle.lib.orc_sizeof(3)
    c = OrcChainQ28()
    c.n_out, c.n_bands, c.max_delay = 5, 10, 2048
    c.bypass_master_eq = int(params["bypass_master_eq"])
    c.loudness_on = int(params["loudness_enabled"])
    c.crossfeed_on = int(params["crossfeed_enabled"])
    c.leveller_on = int(params["leveller_enabled"])
    c.host_mute = int(params["host_mute"])
    c.lev_lookahead = int(params["leveller_lookahead"])
    c.host_vol_mul = int(params["host_vol_mul"])
    c.preset_mute_gain = float(params["preset_mute_gain"])
    c.master_volume_q15 = int(params["master_volume_q15"])
    c.preamp_q28[0], c.preamp_q28[1] = int(params["preamp_q28"][0]), int(params["preamp_q28"][1])
    m = params["matrix"]
    any_delay = False
    for o in range(5):
        for i in range(2):
            x = m["crosspoints"][i, o]
            c.xp[i][o].enabled, c.xp[i][o].phase_invert = int(x["enabled"]), int(x["phase_invert"])
            c.xp[i][o].gain_db, c.xp[i][o].gain_linear = float(x["gain_db"]), float(x["gain_linear"])
        oc = m["outputs"][o]
        c.out[o].enabled, c.out[o].mute = int(oc["enabled"]), int(oc["mute"])
        c.out[o].gain_db, c.out[o].gain_linear = float(oc["gain_db"]), float(oc["gain_linear"])
        c.out[o].delay_ms, c.out[o].delay_samples = float(oc["delay_ms"]), int(oc["delay_samples"])
        c.delay_samples[o] = int(oc["delay_samples"])
        any_delay = any_delay or int(oc["delay_samples"]) > 0
    c.any_delay_active = 1 if any_delay else 0
    bq = np.ascontiguousarray(biquads)
    for r in range(7):
        C.memmove(C.addressof(c.filters[r]), bq[r].ctypes.data, 12 * 32)
        c.channel_bypassed[r] = 1 if all(int(biquads[r, b]["bypass"]) for b in range(10)) else 0
    for j in range(2):
        C.memmove(C.addressof(c.loud[j]), params["loudness"][j:j + 1].tobytes(), 24)
    C.memmove(C.addressof(c.xfeed), params["crossfeed"].tobytes(), 28)
    C.memmove(C.addressof(c.levc), params["leveller"].tobytes(), 36)
    c.levs.gain_q28 = 1 << 28
    c.levs.gain_prev_q28 = 1 << 28
    c.pdm.rng = 123456789
    return c


def orc_chain_run_q28(oracle, chain, pcm_bytes, bit_depth, n_packets, fpp):
    F = n_packets * fpp
    bpf = 6 if bit_depth == 24 else 4
    spdif = np.zeros((2, F, 2), np.int32)
    pdm = np.zeros((F, 8), np.uint32)
    data = np.ascontiguousarray(pcm_bytes)
    for p in range(n_packets):
        oracle.lib.orc_q28_chain_packet(C.addressof(chain), data.ctypes.data + p * fpp * bpf, fpp * bpf, bit_depth,
                                        spdif.ctypes.data + p * fpp * 8, F * 2, pdm.ctypes.data + p * fpp * 32)
    return spdif, pdm


def arm_mute_envelope(chain, fs, loading=True, counter=None, smooth_gain=1.0):
    """Puts one oracle instance into envelope mode the way the firmware arms a preset mute
    (flash_storage.c:272-276, 347-348: counter = max(512, ceil(fs * 10 ms)), preset_loading = true)."""
    chain.mute_env_on = 1
    chain.sample_rate_hz = int(fs)
    chain.preset_loading = 1 if loading else 0
    if counter is None:
        counter = max(512, (int(fs) * 10 + 999) // 1000)
    chain.preset_mute_counter = int(counter)
    chain.preset_mute_smooth_gain = float(smooth_gain)


class RefChain:
    """The reference's own process_audio_packet() (usb_audio.c, compiled unmodified by oracle/ref_chain_shim.c)."""

    NAMES = {"f32s": "libdspi_ref_chain_rp2350_strict.so", "f32f": "libdspi_ref_chain_rp2350_fused.so",
             "q28": "libdspi_ref_chain_rp2040.so"}

    @staticmethod
    def available():
        return all(os.path.exists(os.path.join(ORACLE_DIR, "_ref", n)) for n in RefChain.NAMES.values())

    def __init__(self, flavour):
        self.flavour = flavour
        self.q28 = flavour == "q28"
        self.lib = L = C.CDLL(os.path.join(ORACLE_DIR, "_ref", RefChain.NAMES[flavour]))
        L.ref_chain_sizeof.restype = C.c_size_t
        L.ref_chain_sizeof.argtypes = [C.c_int]
        L.ref_chain_packet.restype = _u32
        L.ref_chain_packet.argtypes = [_vp, _u32, _vp, _u32, _u32, _vp, _u32, _vp, _vp]
        L.ref_host_vol_mul.restype = C.c_int16
        L.ref_host_vol_mul.argtypes = [C.c_int16, _vp]
        L.ref_preamp.argtypes = [C.c_float, _vp, _vp]
        L.ref_master_volume.argtypes = [C.c_float, _vp, _vp]
        assert L.ref_chain_sizeof(0) == C.sizeof(OrcChainQ28 if self.q28 else OrcChainF32)
        self.n_pairs = int(L.ref_chain_sizeof(4))

    def run(self, chain, fs, pcm_bytes, bit_depth, n_packets, fpp):
        """n_packets packets through the reference; returns (spdif [pairs, F, 2], sub_q28 [n pushed])."""
        F = n_packets * fpp
        bpf = 6 if bit_depth == 24 else 4
        spdif = np.zeros((self.n_pairs, F, 2), np.int32)
        sub = np.zeros(F, np.int32)
        got = C.c_uint32()
        data = np.ascontiguousarray(pcm_bytes)
        tot = 0
        for p in range(n_packets):
            n = self.lib.ref_chain_packet(C.addressof(chain), int(fs), data.ctypes.data + p * fpp * bpf, fpp * bpf, bit_depth,
                                          spdif.ctypes.data + p * fpp * 8, F * 2, sub.ctypes.data + tot * 4, C.byref(got))
            assert n == fpp, f"reference returned {n:#x}"
            tot += got.value
        return spdif, sub[:tot]


class RefPdm:
    """The reference's own delta-sigma loop (pdm_generator.c, compiled unmodified by oracle/ref_pdm_shim.c)."""

    PATH = os.path.join(ORACLE_DIR, "_ref", "libdspi_ref_pdm.so")

    @staticmethod
    def available():
        return os.path.exists(RefPdm.PATH)

    def __init__(self):
        self.lib = L = C.CDLL(RefPdm.PATH)
        L.ref_pdm_run.restype = _u32
        L.ref_pdm_run.argtypes = [_vp, _u32, _vp, _vp, _vp]

    def run(self, samples_q28, seed=123456789):
        """One stream from the restart state; returns (words [n, 8], rng state after)."""
        x = np.ascontiguousarray(samples_q28, np.int32)
        out = np.zeros((len(x), 8), np.uint32)
        rng = np.array([seed], np.uint32)
        cnt = np.zeros(4, np.uint32)
        nw = self.lib.ref_pdm_run(x.ctypes.data, len(x), rng.ctypes.data, out.ctypes.data, cnt.ctypes.data)
        assert nw == 8 * len(x) and not cnt.any(), (nw, cnt)
        return out, int(rng[0])
