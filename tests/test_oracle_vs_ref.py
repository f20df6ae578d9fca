"""Pins the restatement (oracle/liborc.so) against the UNMODIFIED reference
sources compiled on the host (oracle/_ref/*.so): bit-exact, every flavour."""
import numpy as np
import pytest

from dspi_b200 import layouts as L
from dspi_b200 import workloads as W


def _bits(a):
    """Bit pattern of every named field (padding bytes excluded)."""
    a = np.ascontiguousarray(a)
    if a.dtype.names is None:
        return a.view(np.uint8)
    return np.concatenate([np.ascontiguousarray(a[n]).view(np.uint8).reshape(-1) for n in a.dtype.names])


def test_record_sizes_match_reference(refs):
    f, q = refs["f32s"].lib, refs["q28"].lib
    assert f.ref_sizeof(0) == L.BIQUAD_F32.itemsize == 68
    assert q.ref_sizeof(0) == L.BIQUAD_Q28.itemsize == 32
    assert f.ref_sizeof(1) == L.EQ_PARAM.itemsize
    assert f.ref_sizeof(2) == L.CROSSPOINT.itemsize and f.ref_sizeof(3) == L.OUTPUT.itemsize
    assert f.ref_sizeof(4) == 396 and q.ref_sizeof(4) == 220          # MatrixMixer
    assert f.ref_sizeof(5) == 28 and f.ref_sizeof(6) == 36 and f.ref_sizeof(7) == 3864
    assert q.ref_sizeof(7) == 3864
    assert f.ref_sizeof(8) == L.LOUD_F32.itemsize and q.ref_sizeof(8) == L.LOUD_Q28.itemsize
    assert f.ref_sizeof(11) == L.MAX_BANDS
    assert (f.ref_sizeof(12), f.ref_sizeof(13), f.ref_sizeof(14)) == (11, 9, 4096)
    assert (q.ref_sizeof(12), q.ref_sizeof(13), q.ref_sizeof(14)) == (7, 5, 2048)
    offs = [f.ref_offsetof(i) for i in range(7)]
    assert offs == [20, 65, 28, 40, 52, 60, 64]
    assert q.ref_offsetof(1) == 28


def test_q28_q15_multiplies(oracle, refs):
    rng = np.random.default_rng(7)
    q = refs["q28"].lib
    vals = np.concatenate([rng.integers(-2**31, 2**31, 4000, dtype=np.int64),
                           np.array([0, 1, -1, 2**31 - 1, -2**31, 1 << 28, -(1 << 28), 0xFFFF, 0x10000, -0x10000])])
    a = rng.permutation(vals)[:2000]
    b = rng.permutation(vals)[:2000]
    for x, y in zip(a, b):
        x, y = int(x), int(y)
        assert oracle.lib.orc_mul_q28(x, y) == q.ref_mul_q28(x, y)
        assert oracle.lib.orc_mul_q15(x, y) == q.ref_mul_q15(x, y)


@pytest.mark.parametrize("variant", ["A", "B", "mixed"])
@pytest.mark.parametrize("fs", [48000.0, 96000.0])
def test_eq_coefficients(oracle, refs, variant, fs):
    params = W.eq_params(variant, 24, fs=fs, seed=3)
    # float store vs the strict reference build
    p1, p2 = params.copy(), params.copy()
    b1 = np.zeros(params.shape, L.BIQUAD_F32)
    b2 = np.zeros(params.shape, L.BIQUAD_F32)
    oracle.eq_coeffs(False, p1, b1, fs)
    refs["f32s"].eq_coeffs(p2, b2, fs)
    assert np.array_equal(_bits(b1), _bits(b2)) and np.array_equal(_bits(p1), _bits(p2))
    # Q28 store
    p1, p2 = params.copy(), params.copy()
    b1 = np.zeros(params.shape, L.BIQUAD_Q28)
    b2 = np.zeros(params.shape, L.BIQUAD_Q28)
    oracle.eq_coeffs(True, p1, b1, fs)
    refs["q28"].eq_coeffs(p2, b2, fs)
    assert np.array_equal(_bits(b1), _bits(b2))


@pytest.mark.parametrize("flavour", ["f32s", "f32f"])
@pytest.mark.parametrize("variant", ["A", "B", "mixed"])
def test_float_cascade_bit_exact(oracle, refs, flavour, variant):
    fs, Cn, T = 96000.0, 16, 2000
    params = W.eq_params(variant, Cn, fs=fs, seed=11)
    bq = np.zeros(params.shape, L.BIQUAD_F32)
    refs["f32s"].eq_coeffs(params, bq, fs)
    x = W.inputs_f32(Cn, T)
    x[3] = 0; x[3, 0] = 1.0                       # impulse -> decays through the denormal range (FTZ)
    x[4] *= 1e-30                                 # tiny signal
    b1, b2, y1, y2 = bq.copy(), bq.copy(), x.copy(), x.copy()
    oracle.eq_many(flavour, b1, y1, 10, 48)
    refs[flavour].eq_many(b2, y2, 10, 48)
    assert np.array_equal(y1.view(np.uint32), y2.view(np.uint32))
    assert np.array_equal(_bits(b1), _bits(b2))
    # packet size must not matter (block form == per-sample form, dsp_pipeline.c:256-279)
    b3, y3 = bq.copy(), x.copy()
    oracle.eq_many(flavour, b3, y3, 10, 7)
    assert np.array_equal(y1.view(np.uint32), y3.view(np.uint32))


def test_fused_differs_from_strict(oracle):
    fs, Cn, T = 96000.0, 4, 4000
    params = W.eq_params("B", Cn, fs=fs, seed=5)
    bq = np.zeros(params.shape, L.BIQUAD_F32)
    oracle.eq_coeffs(False, params, bq, fs)
    x = W.inputs_f32(Cn, T)
    ys, yf = x.copy(), x.copy()
    oracle.eq_many("f32s", bq.copy(), ys, 10, 96)
    oracle.eq_many("f32f", bq.copy(), yf, 10, 96)
    assert not np.array_equal(ys, yf)             # H2: contraction changes bits
    assert np.allclose(ys, yf, rtol=0, atol=1e-4)


@pytest.mark.parametrize("variant", ["A", "B", "mixed"])
def test_q28_cascade_bit_exact(oracle, refs, variant):
    fs, Cn, T = 96000.0, 16, 2000
    params = W.eq_params(variant, Cn, fs=fs, seed=13)
    bq = np.zeros(params.shape, L.BIQUAD_Q28)
    refs["q28"].eq_coeffs(params, bq, fs)
    x = W.inputs_q28(Cn, T)
    x[2] = np.random.default_rng(1).integers(-2**31, 2**31, T, dtype=np.int64).astype(np.int32)  # wrap-around stress
    b1, b2, y1, y2 = bq.copy(), bq.copy(), x.copy(), x.copy()
    oracle.eq_many("q28", b1, y1, 10, 48)
    refs["q28"].eq_many(b2, y2, 10, 48)
    assert np.array_equal(y1, y2) and np.array_equal(_bits(b1), _bits(b2))


@pytest.mark.parametrize("flavour", ["f32s", "f32f", "q28"])
@pytest.mark.parametrize("preset", [0, 1, 2, 3])
def test_crossfeed(oracle, refs, flavour, preset):
    fs, T = 48000.0, 3000
    q = flavour == "q28"
    dt = L.XFEED_Q28 if q else L.XFEED_F32
    st_ref = np.zeros(1, dt)
    refs[flavour].lib.ref_xfeed_coeffs(st_ref.ctypes.data, 1, 1, preset, 1234.0, 7.0, fs)
    # restated coefficient function against the same-arithmetic reference build
    import ctypes as C
    cfg = (C.c_uint8 * 12)()
    cfg_np = np.frombuffer(cfg, np.uint8)
    cfg_np[0], cfg_np[1], cfg_np[2] = 1, 1, preset
    cfg_np[4:8] = np.frombuffer(np.float32(1234.0).tobytes(), np.uint8)
    cfg_np[8:12] = np.frombuffer(np.float32(7.0).tobytes(), np.uint8)
    st_orc = np.zeros(1, dt)
    (oracle.lib.orc_xfeed_coeffs_q28 if q else oracle.lib.orc_xfeed_coeffs_f32)(st_orc.ctypes.data, C.addressof(cfg), fs)
    if flavour != "f32f":      # the fused reference build contracts the coefficient maths too
        assert np.array_equal(_bits(st_orc), _bits(st_ref))
    st_orc = st_ref.copy()
    if q:
        l, r = W.inputs_q28(2, T)
    else:
        l, r = W.inputs_f32(2, T)
    l1, r1, l2, r2 = l.copy(), r.copy(), l.copy(), r.copy()
    oracle.xfeed(flavour, st_orc, l1, r1)
    refs[flavour].xfeed(st_ref, l2, r2)
    assert np.array_equal(_bits(l1), _bits(l2)) and np.array_equal(_bits(r1), _bits(r2))
    assert np.array_equal(_bits(st_orc), _bits(st_ref))


@pytest.mark.parametrize("flavour", ["f32s", "f32f", "q28"])
@pytest.mark.parametrize("lookahead", [0, 1])
@pytest.mark.parametrize("count", [1, 2, 48, 96, 191])
def test_leveller(oracle, refs, flavour, lookahead, count):
    fs = 96000.0
    q = flavour == "q28"
    oracle.set_libm_f64(0)
    oracle.set_x86_cvt(1)          # the x86 reference objects use CVTTSS2SI for the gain-cap cast
    try:
        coeffs = np.zeros(1, L.LEV_COEFFS)
        refs[flavour].lib.ref_lev_coeffs(coeffs.ctypes.data, 70.0, 2, 15.0, -80.0, fs)
        st_ref = np.zeros(1, L.LEV_STATE_Q28 if q else L.LEV_STATE_F32)
        refs[flavour].lib.ref_lev_reset(st_ref.ctypes.data)
        st_orc = st_ref.copy()
        nblk = 40
        if q:
            l, r = W.inputs_q28(2, nblk * count)
            l >>= 3
            r >>= 4
        else:
            l, r = W.inputs_f32(2, nblk * count)
            l *= 0.1
            r *= 0.05
        l1, r1, l2, r2 = l.copy(), r.copy(), l.copy(), r.copy()
        for k in range(nblk):
            s = slice(k * count, (k + 1) * count)
            oracle.leveller(flavour, st_orc, coeffs, lookahead, l1[s], r1[s])
            refs[flavour].leveller(st_ref, coeffs, lookahead, l2[s], r2[s])
        assert np.array_equal(_bits(l1), _bits(l2)) and np.array_equal(_bits(r1), _bits(r2))
        assert np.array_equal(_bits(st_orc), _bits(st_ref))
        assert not np.array_equal(_bits(l1), _bits(l))       # the leveller did something
    finally:
        oracle.set_x86_cvt(0)


def test_leveller_coeffs_and_loudness_tables(oracle, refs):
    import ctypes as C
    fs = 96000.0
    for speed in (0, 1, 2, 7):
        c_ref = np.zeros(1, L.LEV_COEFFS)
        refs["f32s"].lib.ref_lev_coeffs(c_ref.ctypes.data, 55.0, speed, 40.0, -100.0, fs)
        cfg = np.zeros(24, np.uint8)
        cfg[0] = 1
        cfg[4:8] = np.frombuffer(np.float32(55.0).tobytes(), np.uint8)
        cfg[8] = speed
        cfg[12:16] = np.frombuffer(np.float32(40.0).tobytes(), np.uint8)
        cfg[16] = 1
        cfg[20:24] = np.frombuffer(np.float32(-100.0).tobytes(), np.uint8)
        assert refs["f32s"].lib.ref_sizeof(10) == 24 and refs["f32s"].lib.ref_offsetof(8) == 20
        c_orc = np.zeros(1, L.LEV_COEFFS)
        oracle.lib.orc_lev_coeffs_compute(c_orc.ctypes.data, cfg.ctypes.data, fs)
        assert np.array_equal(_bits(c_orc), _bits(c_ref))
    for ref_spl, inten in ((83.0, 100.0), (70.0, 50.0), (120.0, 150.0)):
        t_ref = np.zeros((L.LOUD_STEPS, 2), L.LOUD_F32)
        t_orc = np.zeros((L.LOUD_STEPS, 2), L.LOUD_F32)
        refs["f32s"].lib.ref_loud_table(t_ref.ctypes.data, ref_spl, inten, fs)
        oracle.lib.orc_loud_table_f32(t_orc.ctypes.data, ref_spl, inten, fs)
        assert np.array_equal(_bits(t_orc), _bits(t_ref))
        t_ref = np.zeros((L.LOUD_STEPS, 2), L.LOUD_Q28)
        t_orc = np.zeros((L.LOUD_STEPS, 2), L.LOUD_Q28)
        refs["q28"].lib.ref_loud_table(t_ref.ctypes.data, ref_spl, inten, fs)
        oracle.lib.orc_loud_table_q28(t_orc.ctypes.data, ref_spl, inten, fs)
        assert np.array_equal(_bits(t_orc), _bits(t_ref))


def test_delay_samples_and_volume_table(oracle, refs):
    for fs in (44100.0, 48000.0, 96000.0):
        for ms in (0.0, 0.5, 10.0, 42.6, 85.3, 85.4, 1000.0, -3.0):
            for last in (0, 1):
                assert oracle.lib.orc_delay_samples(ms, fs, last, 4096) == refs["f32s"].lib.ref_delay_samples(ms, fs, last)
                assert oracle.lib.orc_delay_samples(ms, fs, last, 2048) == refs["q28"].lib.ref_delay_samples(ms, fs, last)
    # KATs stated in the reference: SUB_ALIGN_SAMPLES = 128; 0 dB host volume -> 0x8000 -> int16 -32768 (quirk 1)
    assert oracle.lib.orc_delay_samples(0.0, 48000.0, 1, 4096) == 128
    assert oracle.lib.orc_host_vol_mul(0, None) == -32768
    assert oracle.lib.orc_host_vol_mul(-20 * 256, None) == 0x0CCD
    assert oracle.lib.orc_host_vol_mul(-32768, None) == 0
