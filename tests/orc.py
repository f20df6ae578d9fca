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
        L.orc_eq_coeffs_f32.argtypes = [_vp, _vp, C.c_float]
        L.orc_eq_coeffs_q28.argtypes = [_vp, _vp, C.c_float]
        L.orc_xfeed_coeffs_f32.argtypes = [_vp, _vp, C.c_float]
        L.orc_xfeed_coeffs_q28.argtypes = [_vp, _vp, C.c_float]
        L.orc_lev_coeffs_compute.argtypes = [_vp, _vp, C.c_float]
        L.orc_loud_table_f32.argtypes = [_vp, C.c_float, C.c_float, C.c_float]
        L.orc_loud_table_q28.argtypes = [_vp, C.c_float, C.c_float, C.c_float]
        L.orc_pdm_modulate.argtypes = [_vp, C.c_int32, _vp]

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
    def eq_coeffs(self, q28, params, bq, fs):
        """params: EQ_PARAM[N] (mutated: clamps written back); bq: matching biquads."""
        fn = self.lib.orc_eq_coeffs_q28 if q28 else self.lib.orc_eq_coeffs_f32
        pp = params.reshape(-1)
        bb = bq.reshape(-1)
        base_p, base_b = pp.ctypes.data, bb.ctypes.data
        sp, sb = pp.dtype.itemsize, bb.dtype.itemsize
        for i in range(pp.shape[0]):
            fn(base_p + i * sp, base_b + i * sb, fs)


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
