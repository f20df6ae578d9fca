"""GPU: dsp_compute_coefficients() on the device (dspi_eq_set_params_device, coeff.cu) against the oracle under the
same libm policy (libm calls evaluated in double and rounded once): bit-exact; against the host-libm path
(glibc float functions): a handful of bands differ in a last bit of A or tan(), i.e. by <= 1e-6 in a coefficient -
the documented difference."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from dspi_b200 import api, layouts as L, workloads as W      # noqa: E402
from tests.util import same_bits, ulp_diff                   # noqa: E402

FS = [44100.0, 48000.0, 96000.0]


def _recipes(n, seed):
    """Every filter type, flat bands, out-of-range Q / frequency (clamped), tiny gains (treated as flat)."""
    rng = np.random.default_rng(seed)
    p = np.zeros((n, L.MAX_BANDS), L.EQ_PARAM)
    p["band"] = np.arange(L.MAX_BANDS, dtype=np.uint8)[None, :]
    p["type"] = rng.integers(0, 6, (n, L.MAX_BANDS))
    p["freq"] = (5.0 * (12000.0 ** rng.random((n, L.MAX_BANDS)))).astype(np.float32)        # 5 Hz .. 60 kHz
    p["Q"] = (0.05 * (600.0 ** rng.random((n, L.MAX_BANDS)))).astype(np.float32)           # 0.05 .. 30
    g = rng.uniform(-18, 18, (n, L.MAX_BANDS)).astype(np.float32)
    g[rng.random((n, L.MAX_BANDS)) < 0.1] = 0.005
    p["gain_db"] = g
    p["freq"][rng.random((n, L.MAX_BANDS)) < 0.03] = 0.0
    return p


@pytest.mark.parametrize("q28", [False, True])
@pytest.mark.parametrize("fs", FS)
def test_device_coefficients_match_oracle_policy(oracle, q28, fs):
    n = 600
    rec = _recipes(n, int(fs) + q28)
    eng = api.EqEngine("q28" if q28 else "f32f", n, 12)
    try:
        clamped = eng.set_params_device(rec, fs)
        got = eng.download()
    finally:
        eng.close()
    want_rec = rec.copy()
    want = np.zeros((n, L.MAX_BANDS), L.BIQUAD_Q28 if q28 else L.BIQUAD_F32)
    oracle.set_libm_f64(1)
    try:
        oracle.eq_coeffs(q28, want_rec, want, fs)
    finally:
        oracle.set_libm_f64(0)
    assert same_bits(clamped, want_rec), "recipe clamps written back"
    if not same_bits(got, want):
        msg = []
        for f in got.dtype.names:
            bad = np.argwhere(np.ascontiguousarray(got[f]).view(np.uint8).reshape(n, L.MAX_BANDS, -1).any(axis=2) !=
                              np.ascontiguousarray(want[f]).view(np.uint8).reshape(n, L.MAX_BANDS, -1).any(axis=2)) if False else \
                np.argwhere((np.ascontiguousarray(got[f]).view(np.uint8).reshape(n, L.MAX_BANDS, -1) !=
                             np.ascontiguousarray(want[f]).view(np.uint8).reshape(n, L.MAX_BANDS, -1)).any(axis=2))
            if len(bad):
                c, b = bad[0]
                msg.append(f"{f}: {len(bad)} bands, first ch {c} band {b}: gpu {got[f][c, b]!r} want {want[f][c, b]!r} recipe {rec[c, b]}")
        raise AssertionError("coefficients under the libm policy differ:\n" + "\n".join(msg))
    # the host-libm path (glibc float functions): last-bit differences only
    host = api.compute_coefficients(rec.copy(), q28=q28, fs=fs)
    if q28:
        for f in ("b0", "b1", "b2", "a1", "a2"):
            assert np.max(np.abs(got[f].astype(np.int64) - host[f].astype(np.int64))) <= 512, f          # 2e-6 in Q28
    else:
        for f in ("b0", "b1", "b2", "a1", "a2", "sva1", "sva2", "sva3", "svm0", "svm1", "svm2"):
            assert float(np.max(np.abs(got[f] - host[f]))) <= 1e-6, f
        differing = int(np.sum(np.any(np.stack([got[f].view(np.uint32) != host[f].view(np.uint32) for f in ("b0", "b1", "b2", "a1", "a2", "sva1", "svm1")]), axis=0)))
        assert differing <= 0.01 * got.size, differing
        for f in ("use_svf", "bypass", "svf_type"):
            assert np.array_equal(got[f], host[f]), f


def test_state_survives_unless_the_topology_flips(oracle):
    fs, n, T = 96000.0, 64, 512
    rec = W.eq_params("B", n, fs=fs, seed=3)
    eng = api.EqEngine("f32f", n, 10)
    try:
        eng.set_params_device(rec, fs)
        x = torch.from_numpy(W.inputs_f32(n, T)).cuda()
        eng.process_device(x.data_ptr(), T, T)
        eng.sync()
        rec2 = rec.copy()
        rec2["gain_db"][:, 1] += 1.0                      # same topology: state kept
        rec2["freq"][:, 2] = 15000.0                      # SVF -> TDF2 at 96 kHz (>= fs / 7.5): state cleared
        # a twin engine provides the state "before" so that the engine under test is reconfigured with nothing but
        # its live, packed state (no download in between)
        twin = api.EqEngine("f32f", n, 10)
        twin.set_params_device(rec, fs)
        y = torch.from_numpy(W.inputs_f32(n, T)).cuda()
        twin.process_device(y.data_ptr(), T, T)
        twin.sync()
        before = twin.download()
        twin.close()
        eng.set_params_device(rec2, fs)
        after = eng.download()
    finally:
        eng.close()
    want = before.copy()
    want_rec = rec2.copy()
    oracle.set_libm_f64(1)
    try:
        oracle.eq_coeffs(False, want_rec, want, fs)
    finally:
        oracle.set_libm_f64(0)
    assert same_bits(after, want)
    assert np.all(after["svic1eq"][:, 2] == 0) and np.any(before["svic1eq"][:, 2] != 0)
    assert np.array_equal(after["svic1eq"][:, 1].view(np.uint32), before["svic1eq"][:, 1].view(np.uint32))


def test_mass_reconfiguration_throughput():
    """65536 instances x 11 channels x 12 bands = 8.65 M coefficient sets in one call; printed for the record."""
    import time
    n = 65536 * 11
    rec = np.tile(W.eq_params("B", 11, fs=96000.0, nbands=12, seed=1), (65536, 1))
    eng = api.EqEngine("f32f", n, 12)
    try:
        eng.set_params_device(rec[:1024], 96000.0)
        t0 = time.perf_counter()
        eng.set_params_device(rec, 96000.0)
        dt = time.perf_counter() - t0
    finally:
        eng.close()
    print(f"\n8.65 M coefficient sets (recipes H2D, kernel, clamps D2H, pack): {dt * 1e3:.1f} ms = {n * 12 / dt / 1e6:.0f} M sets/s")
    assert dt < 5.0


@pytest.mark.parametrize("q28", [False, True])
def test_chain_recipes_on_device(oracle, q28):
    """dsp_recalculate_all_filters() for whole instances: recipes [n][roles][12] -> both internal EQ engines."""
    fs, n = 48000.0, 21
    roles = L.CHAINQ_EQ_CHANNELS if q28 else L.CHAIN_EQ_CHANNELS
    rec = _recipes(n * roles, 77 + q28).reshape(n, roles, L.MAX_BANDS)
    eng = api.ChainEngineQ28(n, max_frames=96) if q28 else api.ChainEngine("f32f", n, max_frames=96)
    try:
        clamped = eng.set_eq_params_device(rec, fs)
        got = eng.download_biquads()
    finally:
        eng.close()
    want_rec = rec.copy().reshape(-1, L.MAX_BANDS)
    want = np.zeros((n * roles, L.MAX_BANDS), L.BIQUAD_Q28 if q28 else L.BIQUAD_F32)
    if q28:
        want["bypass"] = 1                                # the chain starts with every band bypassed
    oracle.set_libm_f64(1)
    try:
        oracle.eq_coeffs(q28, want_rec, want, fs)
    finally:
        oracle.set_libm_f64(0)
    assert same_bits(clamped.reshape(-1, L.MAX_BANDS), want_rec)
    assert same_bits(got.reshape(-1, L.MAX_BANDS), want)
