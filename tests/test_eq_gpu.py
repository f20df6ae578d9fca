"""GPU parity: the CUDA EQ kernels (K1 float fused/strict, K2 Q28), called through the C ABI,
against the CPU oracle on the same seeded inputs and against the committed fixtures.

Bars (BASELINE.json north_star): Q28 bit-exact; float <= 1 ULP against the oracle run in the
same arithmetic flavour (observed and asserted: 0 ULP, i.e. bit-identical)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from dspi_b200 import api, layouts as L, workloads as W      # noqa: E402
from tests.util import load_golden, same_bits, ulp_diff      # noqa: E402

FLOAT_ULP_TOL = 1      # north_star: <= 1 ULP for the float biquad/SVF path


def _run_gpu(flavour, bq, x, n_bands=10, splits=None, ld=None):
    """Process x [C, T] through an engine; returns (y, final biquads)."""
    Cn, T = x.shape
    eng = api.EqEngine(flavour, Cn, n_bands)
    try:
        eng.upload(bq)
        ld = T if ld is None else ld
        buf = torch.zeros((Cn, ld), dtype=torch.int32 if eng.q28 else torch.float32, device="cuda")
        buf[:, :T] = torch.from_numpy(x).cuda()
        torch.cuda.synchronize()
        t0 = 0
        for n in (splits or [T]):
            eng.process_device(buf.data_ptr() + t0 * 4, n, ld)
            t0 += n
        eng.sync()
        y = buf[:, :T].cpu().numpy()
        return y, eng.download()
    finally:
        eng.close()


def _oracle(oracle, flavour, bq, x, n_bands=10, packet=96):
    b, y = bq.copy(), x.copy()
    oracle.eq_many(flavour, b, y, n_bands, packet)
    return y, b


def _coeffs(variant, Cn, fs, q28, seed=5):
    params = W.eq_params(variant, Cn, fs=fs, seed=seed)
    return api.compute_coefficients(params, q28=q28, fs=fs)


def _assert_float_equal(y, want):
    d = ulp_diff(y, want)
    if d:
        bad = np.argwhere(y.view(np.uint32) != want.view(np.uint32))
        c, t = bad[0]
        print(f"first mismatch at channel {c} sample {t}: gpu {y[c, t]!r} oracle {want[c, t]!r}; "
              f"{len(bad)} mismatching samples in channels {sorted(set(bad[:, 0].tolist()))[:16]}")
    assert d <= FLOAT_ULP_TOL, f"max ULP distance {d}"
    assert d == 0, "float path is expected to be bit-identical to the same-flavour oracle"


@pytest.mark.parametrize("flavour", ["f32f", "f32s"])
@pytest.mark.parametrize("variant", ["A", "B", "mixed"])
def test_float_cascade_matches_oracle(oracle, flavour, variant):
    fs, Cn, T = 96000.0, 256, 2048
    bq = _coeffs(variant, Cn, fs, False)
    x = W.inputs_f32(Cn, T)
    x[3] = 0; x[3, 0] = 1.0            # impulse: decays through the flush-to-zero range
    x[5] *= 1e-30
    y, st = _run_gpu(flavour, bq, x)
    want, wst = _oracle(oracle, flavour, bq, x)
    _assert_float_equal(y, want)
    assert same_bits(st, wst)


@pytest.mark.parametrize("variant", ["A", "B", "mixed"])
def test_q28_cascade_bit_exact(oracle, variant):
    fs, Cn, T = 96000.0, 256, 2048
    bq = _coeffs(variant, Cn, fs, True)
    x = W.inputs_q28(Cn, T)
    x[2] = np.random.default_rng(1).integers(-2**31, 2**31, T, dtype=np.int64).astype(np.int32)   # wrap-around stress
    y, st = _run_gpu("q28", bq, x)
    want, wst = _oracle(oracle, "q28", bq, x)
    assert np.array_equal(y, want)
    assert same_bits(st, wst)


@pytest.mark.parametrize("n_bands", [1, 7, 10, 12])
@pytest.mark.parametrize("shape", ["all_on", "mixed", "one_band_flat_everywhere"])
def test_q28_band_counts_and_bypass_paths(oracle, n_bands, shape):
    """K2 has three code paths per warp (eq_q28.cu): one straight-line block when every band of every lane runs, the same
    block with selects when some lanes bypass a band, and the per-band branches when a band is flat in the whole warp or the
    engine has fewer bands than the kernel's template.  Every (band count, bypass shape) pair lands on one of them; all must
    be the firmware's arithmetic bit for bit (dsp_process_rp2040.S:225-394, bypass test :246-248)."""
    fs, Cn, T = 48000.0, 96, 1000                       # 3 warps; T leaves a partial register tile at the end
    params = W.eq_params("mixed" if shape == "mixed" else "A", Cn, fs=fs, nbands=n_bands, seed=21)
    if shape == "one_band_flat_everywhere":
        params["gain_db"][:, n_bands // 2] = 0.0
        params["type"][:, n_bands // 2] = L.PEAKING
    bq = api.compute_coefficients(params, q28=True, fs=fs)
    if shape == "all_on":
        assert not bq["bypass"][:, :n_bands].any()
    if shape == "one_band_flat_everywhere":
        assert bq["bypass"][:, n_bands // 2].all()
    x = W.inputs_q28(Cn, T)
    x[5] = np.random.default_rng(3).integers(-2**31, 2**31, T, dtype=np.int64).astype(np.int32)          # wrap-around stress
    y, st = _run_gpu("q28", bq, x, n_bands=n_bands, splits=[504, 496])
    want, wst = _oracle(oracle, "q28", bq, x, n_bands=n_bands)
    assert np.array_equal(y, want)
    assert same_bits(st, wst)


@pytest.mark.parametrize("flavour", ["f32f", "q28"])
@pytest.mark.parametrize("Cn,T,ld", [(1, 1, 4), (2, 48, 48), (100, 1004, 1004), (65, 1003, 1003), (130, 96, 200), (64, 33, 36)])
def test_ragged_shapes(oracle, flavour, Cn, T, ld):
    """empty-ish, ragged and unaligned shapes: partial groups, partial tiles, non-TMA strides"""
    fs = 48000.0
    q = flavour == "q28"
    bq = _coeffs("mixed", Cn, fs, q, seed=9)
    x = W.inputs_q28(Cn, T) if q else W.inputs_f32(Cn, T)
    y, st = _run_gpu(flavour, bq, x, ld=ld)
    want, wst = _oracle(oracle, flavour, bq, x, packet=48)
    assert np.array_equal(y.view(np.uint32), want.view(np.uint32))
    assert same_bits(st, wst)


@pytest.mark.parametrize("flavour", ["f32f", "f32s", "q28"])
def test_state_carries_across_calls(oracle, flavour):
    """packets of 96/48/odd sizes in separate launches == one launch == the oracle"""
    fs, Cn, T = 96000.0, 128, 96 * 8 + 48 + 20
    q = flavour == "q28"
    bq = _coeffs("B", Cn, fs, q)
    x = W.inputs_q28(Cn, T) if q else W.inputs_f32(Cn, T)
    y1, st1 = _run_gpu(flavour, bq, x, ld=T)
    y2, st2 = _run_gpu(flavour, bq, x, splits=[96] * 8 + [48, 20], ld=T)
    want, wst = _oracle(oracle, flavour, bq, x)
    assert np.array_equal(y1.view(np.uint32), want.view(np.uint32))
    assert np.array_equal(y2.view(np.uint32), want.view(np.uint32))
    assert same_bits(st1, wst) and same_bits(st2, wst)


@pytest.mark.parametrize("flavour", ["f32f", "f32s", "q28"])
def test_golden_fixtures(flavour):
    g = load_golden("mix.npz")
    q = flavour == "q28"
    y, st = _run_gpu(flavour, g["bq_q28" if q else "bq_f32"], g["q28_x" if q else "x"])
    assert np.array_equal(y.view(np.uint32), g[f"{flavour}_y"].view(np.uint32))
    assert same_bits(st, g[f"{flavour}_state"])
    g = load_golden("cfg1.npz")
    for sig in ("impulse", "sine", "sweep", "noise"):
        y, st = _run_gpu(flavour, g["bq_q28" if q else "bq_f32"], g[f"{sig}_q28_x" if q else f"{sig}_x"], splits=[48] * 100)
        assert np.array_equal(y.view(np.uint32), g[f"{sig}_{flavour}_y"].view(np.uint32)), sig
        assert same_bits(st, g[f"{sig}_{flavour}_state"])


@pytest.mark.parametrize("flavour", ["f32f", "q28"])
def test_host_path_and_set_param(oracle, flavour):
    """dspi_eq_process_host (chunked H2D/compute/D2H) and the REQ_SET_EQ_PARAM path"""
    fs, Cn, T = 48000.0, 192, 4096
    q = flavour == "q28"
    bq = _coeffs("B", Cn, fs, q)
    x = W.inputs_q28(Cn, T) if q else W.inputs_f32(Cn, T)
    eng = api.EqEngine(flavour, Cn)
    pin = api.PinnedBuffer((Cn, T), np.int32 if q else np.float32)
    try:
        eng.upload(bq)
        pin.array[...] = x
        eng.process_host(pin.array)
        want, wst = _oracle(oracle, flavour, bq, x)
        assert np.array_equal(pin.array.view(np.uint32), want.view(np.uint32))
        # change one band between "packets" like main.c:826-857, then keep going
        p = np.zeros(1, L.EQ_PARAM)
        p[0] = (0, 4, L.PEAKING, 0, 2500.0, 2.0, -5.0)
        eng.set_param(7, p[0], fs)
        wst2 = wst.copy()
        row = wst2[7:8, 4]
        oracle.eq_coeffs(q, p.copy(), row, fs)
        wst2[7, 4] = row[0]
        pin.array[...] = x
        eng.process_host(pin.array)
        want2, wst3 = _oracle(oracle, flavour, wst2, x)
        assert np.array_equal(pin.array.view(np.uint32), want2.view(np.uint32))
        assert same_bits(eng.download(), wst3)
    finally:
        pin.free()
        eng.close()


@pytest.mark.parametrize("flavour,Cn", [("f32f", 65536), ("q28", 32768)])
def test_full_size_properties(flavour, Cn):
    """BASELINE configs 2 / 4 at full channel count, checked through size-independent properties:
    (a) every 512th channel against nothing but itself processed alone in a small engine
    (sharding / grouping must not change a bit), (b) all-bypass coefficients are the identity."""
    fs, T = 96000.0, 256
    q = flavour == "q28"
    params = W.eq_params_fast("A" if not q else "B", Cn, fs=fs, seed=3)
    bq = api.compute_coefficients(params[::512].copy(), q28=q, fs=fs)
    full = np.zeros((Cn, L.MAX_BANDS), L.BIQUAD_Q28 if q else L.BIQUAD_F32)
    full["bypass"] = 1
    full["b0"] = (1 << 28) if q else 1.0
    full[::512] = bq
    x = (W.inputs_q28(Cn // 512, T) if q else W.inputs_f32(Cn // 512, T))
    xf = np.zeros((Cn, T), x.dtype)
    xf[::512] = x
    xf[1::512] = x                      # neighbours are bypassed: must come back untouched
    y, _ = _run_gpu(flavour, full, xf)
    ysmall, _ = _run_gpu(flavour, bq, x)
    assert np.array_equal(y[::512].view(np.uint32), ysmall.view(np.uint32))
    assert np.array_equal(y[1::512].view(np.uint32), x.view(np.uint32))


@pytest.mark.parametrize("flavour,Cn,variant", [("f32f", 65536, "A"), ("f32f", 65536, "B"), ("f32s", 65536, "A"), ("q28", 32768, "B")])
def test_full_size_every_channel_against_the_oracle(oracle, flavour, Cn, variant):
    """BASELINE configs 2 / 4 at their full channel counts: every word of every channel against the CPU oracle
    (two 96-frame packets; the oracle runs multithreaded), filter state included."""
    import os
    fs, T = 96000.0, 192
    q = flavour == "q28"
    params = W.eq_params_fast(variant, Cn, fs=fs, seed=5)
    bq = api.compute_coefficients(params, q28=q, fs=fs)
    x = W.inputs_q28(Cn, T) if q else W.inputs_f32(Cn, T)
    y, st = _run_gpu(flavour, bq, x)
    want, wst = x.copy(), bq.copy()
    oracle.eq_many_mt(flavour, wst, want, 10, 96, min(32, os.cpu_count() or 1))
    assert np.array_equal(y.view(np.uint32), want.view(np.uint32))
    assert same_bits(st, wst)


# ---- run-time specialised K1 (eq_jit.cu): same bits as the generic kernels and the oracle ----------
def _run_gpu_info(flavour, bq, x, n_bands=10):
    Cn, T = x.shape
    eng = api.EqEngine(flavour, Cn, n_bands)
    try:
        eng.upload(bq)
        info = eng.kernel_info()
        buf = torch.from_numpy(x).cuda()
        eng.process_device(buf.data_ptr(), T, T)
        eng.sync()
        return buf.cpu().numpy(), eng.download(), info
    finally:
        eng.close()


@pytest.mark.parametrize("flavour", ["f32f", "f32s"])
def test_specialised_kernel_matches_oracle(oracle, flavour, monkeypatch):
    """Variant B (9 SVF + 1 TDF2 at 96 kHz): every channel shares one topology vector -> NVRTC kernel.
    200 channels (ragged last group) x 100 samples (3 register tiles + a ragged tail tile)."""
    monkeypatch.setenv("DSPI_JIT", "force")
    fs, Cn, T = 96000.0, 200, 100
    bq = _coeffs("B", Cn, fs, False)
    x = W.inputs_f32(Cn, T)
    x[3] = 0; x[3, 0] = 1.0
    y, st, info = _run_gpu_info(flavour, bq, x)
    assert info.startswith("jit sig=0x"), info
    want, wst = _oracle(oracle, flavour, bq, x)
    _assert_float_equal(y, want)
    assert same_bits(st, wst)
    monkeypatch.setenv("DSPI_JIT", "0")
    y2, st2, info2 = _run_gpu_info(flavour, bq, x)
    assert info2.startswith("aot generic"), info2
    assert same_bits(y2, y) and same_bits(st2, st)


def test_specialised_kernel_with_foreign_channels(oracle, monkeypatch):
    """A dominant topology vector plus channels that differ from it (other types, bypassed bands, 8 of
    the 10 bands active): warps that do not match fall back to the generic path inside the same kernel."""
    monkeypatch.setenv("DSPI_JIT", "force")
    fs, Cn, T = 96000.0, 320, 256
    bq = _coeffs("B", Cn, fs, False)
    other = _coeffs("mixed", Cn, fs, False, seed=11)
    for c in list(range(64, 128)) + [130, 200, 319]:
        bq[c] = other[c]
    x = W.inputs_f32(Cn, T)
    y, st, info = _run_gpu_info("f32f", bq, x, n_bands=8)
    assert info.startswith("jit sig=0x"), info
    want, wst = _oracle(oracle, "f32f", bq, x, n_bands=8)
    _assert_float_equal(y, want)
    assert same_bits(st, wst)


def test_kernel_choice_reporting(monkeypatch):
    monkeypatch.delenv("DSPI_JIT", raising=False)
    fs = 96000.0
    eng = api.EqEngine("f32f", 64, 10)
    try:
        eng.upload(_coeffs("A", 64, fs, False))
        assert eng.kernel_info().startswith("aot straight-line biquad")
        eng.upload(_coeffs("B", 64, fs, False))
        assert "below 1024 channels" in eng.kernel_info()
    finally:
        eng.close()
    eng = api.EqEngine("q28", 64, 10)
    try:
        assert eng.kernel_info().startswith("aot q28")
    finally:
        eng.close()


def test_host_path_many_small_chunks():
    """dspi_eq_process_host with the staging chunk forced to 1 MiB: 16 chunks through the ring of staging buffers, every buffer's
    kernel stream used twice.  Output and filter state must equal the device-resident call (run in a subprocess: the chunk
    size is read from the environment once per process)."""
    import os, subprocess, sys, textwrap
    code = textwrap.dedent('''
        import numpy as np, torch
        from dspi_b200 import api, workloads as W
        from tests.util import same_bits
        fs, Cn, T = 96000.0, 4096, 1024
        bq = api.compute_coefficients(W.eq_params_fast("B", Cn, fs=fs, seed=3), q28=False, fs=fs)
        x = W.inputs_f32(Cn, T)
        a = api.EqEngine("f32f", Cn); a.upload(bq)
        buf = torch.from_numpy(x).cuda(); a.process_device(buf.data_ptr(), T, T); a.process_device(buf.data_ptr(), T, T); a.sync()
        want, wst = buf.cpu().numpy(), a.download(); a.close()
        b = api.EqEngine("f32f", Cn); b.upload(bq)
        pin = api.PinnedBuffer((Cn, T), np.float32); pin.array[...] = x
        b.process_host(pin.array); b.process_host(pin.array)
        assert np.array_equal(pin.array.view(np.uint32), want.view(np.uint32)), "samples differ"
        assert same_bits(b.download(), wst), "filter state differs"
        assert b.launch_count >= 32, b.launch_count
        b.close(); print("ok")
    ''')
    env = dict(os.environ, DSPI_HOST_CHUNK_MB="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr
