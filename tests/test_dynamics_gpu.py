"""Parameter ingest on the device, second half: dspi_chain(q)_set_dynamics_device generates crossfeed,
leveller and loudness coefficients and applies the host volume on the GPU.  Checked against the oracle's coefficient
functions under the libm policy (bit-exact records) and end to end: a running engine reconfigured on the device must
continue exactly like the oracle instance whose records were replaced the way the firmware's main loop does
(main.c:868-895: new coefficients, crossfeed state cleared, running leveller / loudness state kept)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from dspi_b200 import api, layouts as L                                          # noqa: E402
from tests.chain_cases import chain_params, chain_params_q28, pcm_bytes           # noqa: E402
from tests.orc import make_orc_chain, make_orc_chain_q28, orc_chain_run, orc_chain_run_q28   # noqa: E402

FS = 96000.0


def random_configs(n, seed):
    rng = np.random.default_rng(seed)
    c = np.zeros(n, L.DYNAMICS_CONFIG)
    u = rng.random((n, 16))
    c["xf_enabled"] = u[:, 0] < 0.8
    c["xf_itd_enabled"] = u[:, 1] < 0.7
    c["xf_preset"] = (u[:, 2] * 4).astype(np.uint8)                 # 3 = custom
    c["xf_custom_fc"] = (300 + 2200 * u[:, 3]).astype(np.float32)    # beyond both clamps
    c["xf_custom_feed_db"] = (-2 + 20 * u[:, 4]).astype(np.float32)
    c["lev_enabled"] = u[:, 5] < 0.7
    c["lev_amount"] = (-10 + 130 * u[:, 6]).astype(np.float32)
    c["lev_speed"] = (u[:, 7] * 5).astype(np.uint8)                  # >= 3 falls back to medium
    c["lev_max_gain_db"] = (-5 + 45 * u[:, 8]).astype(np.float32)
    c["lev_lookahead"] = u[:, 9] < 0.5
    c["lev_gate_threshold_db"] = (-110 + 120 * u[:, 10]).astype(np.float32)
    c["loudness_ref_spl"] = (30 + 80 * u[:, 11]).astype(np.float32)  # beyond [40, 100]
    c["loudness_intensity_pct"] = (150 * u[:, 12]).astype(np.float32)
    c["loudness_enabled"] = u[:, 13] < 0.8
    c["host_mute"] = u[:, 14] < 0.05
    c["volume_8_8"] = (-70 * 256 * u[:, 15]).astype(np.int16) + 512  # from above 0 dB down to below -60 dB
    return c


def apply_to_oracle(oracle, chain, cfg, q28):
    """What the main loop does for one instance (main.c:868-895 + audio_set_volume), with the oracle's policy functions."""
    xcfg = (C.c_uint8 * 12)()
    xn = np.frombuffer(xcfg, np.uint8)
    xn[0], xn[1], xn[2] = int(cfg["xf_enabled"]), int(cfg["xf_itd_enabled"]), int(cfg["xf_preset"])
    xn[4:8] = np.frombuffer(np.float32(cfg["xf_custom_fc"]).tobytes(), np.uint8)
    xn[8:12] = np.frombuffer(np.float32(cfg["xf_custom_feed_db"]).tobytes(), np.uint8)
    (oracle.lib.orc_xfeed_coeffs_q28 if q28 else oracle.lib.orc_xfeed_coeffs_f32)(C.addressof(chain.xfeed), C.addressof(xcfg), FS)
    chain.crossfeed_on = int(cfg["xf_enabled"])
    lcfg = np.zeros(24, np.uint8)
    lcfg[0] = int(cfg["lev_enabled"])
    lcfg[4:8] = np.frombuffer(np.float32(cfg["lev_amount"]).tobytes(), np.uint8)
    lcfg[8] = int(cfg["lev_speed"])
    lcfg[12:16] = np.frombuffer(np.float32(cfg["lev_max_gain_db"]).tobytes(), np.uint8)
    lcfg[16] = int(cfg["lev_lookahead"])
    lcfg[20:24] = np.frombuffer(np.float32(cfg["lev_gate_threshold_db"]).tobytes(), np.uint8)
    oracle.lib.orc_lev_coeffs_compute(C.addressof(chain.levc), lcfg.ctypes.data, FS)
    chain.leveller_on, chain.lev_lookahead = int(cfg["lev_enabled"]), int(cfg["lev_lookahead"])
    tab = np.zeros((L.LOUD_STEPS, 2), L.LOUD_Q28 if q28 else L.LOUD_F32)
    (oracle.lib.orc_loud_table_q28 if q28 else oracle.lib.orc_loud_table_f32)(tab.ctypes.data, float(cfg["loudness_ref_spl"]), float(cfg["loudness_intensity_pct"]), FS)
    idx = C.c_uint8()
    chain.host_vol_mul = oracle.lib.orc_host_vol_mul(int(cfg["volume_8_8"]), C.byref(idx))
    chain.host_mute = int(cfg["host_mute"])
    for j in range(2):
        C.memmove(C.addressof(chain.loud[j]), tab[idx.value, j:j + 1].tobytes(), tab.dtype.itemsize)
    chain.loudness_on = int(cfg["loudness_enabled"])


@pytest.mark.parametrize("flavour", ["f32f", "q28"])
def test_device_dynamics_ingest_matches_the_main_loop(oracle, flavour):
    q = flavour == "q28"
    N, n_packets, fpp = 48, 6, 96
    F = n_packets * fpp
    P, bq = chain_params_q28(oracle, N, FS, 41) if q else chain_params(oracle, N, FS, 42)
    for i in range(N):
        P[i]["preset_mute_gain"] = 1.0
    pcm = pcm_bytes(N, 2 * F, 24, 43)
    cfgs = random_configs(N, 44)
    sub_o = 4 if q else 8
    oracle.set_libm_f64(1)
    eng = api.ChainEngineQ28(N, max_frames=F) if q else api.ChainEngine(flavour, N, max_frames=F)
    try:
        eng.set_params(P)
        eng.upload_biquads(bq)
        chains = [(make_orc_chain_q28 if q else make_orc_chain)(oracle, P[i], bq[i]) for i in range(N)]
        run = (lambda ch, data: orc_chain_run_q28(oracle, ch, data, 24, n_packets, fpp)) if q else \
              (lambda ch, data: orc_chain_run(oracle, flavour, ch, data, 24, n_packets, fpp))
        c0 = np.ascontiguousarray(pcm[:, :F * 6])
        eng.process_host(c0, 24, n_packets, fpp)                       # warm state
        for i in range(N):
            run(chains[i], c0[i])
        eng.set_dynamics_device(cfgs, FS)
        for i in range(N):
            apply_to_oracle(oracle, chains[i], cfgs[i], q)
        c1 = np.ascontiguousarray(pcm[:, F * 6:])
        spdif, pdm, status = eng.process_host(c1, 24, n_packets, fpp)
        for i in range(N):
            ws, wp = run(chains[i], c1[i])
            assert np.array_equal(spdif[i], ws), f"instance {i}: S/PDIF words after the device-side reconfiguration"
            if P[i]["matrix"]["outputs"][sub_o]["enabled"]:
                assert np.array_equal(pdm[i], wp), f"instance {i}: PDM bits"
            assert list(status[i]["peaks"]) == list(chains[i].peaks)[:len(status[i]["peaks"])]
    finally:
        eng.close()
        oracle.set_libm_f64(0)


def test_policy_coefficients_stay_close_to_the_host_libm(oracle):
    """The same generators with glibc's float routines (what oracle/_ref pins): the policy moves a coefficient by at most a
    few float ulps - the documented libm difference, not an algorithmic one."""
    rng = np.random.default_rng(7)
    worst = 0.0
    for _ in range(200):
        xcfg = (C.c_uint8 * 12)()
        xn = np.frombuffer(xcfg, np.uint8)
        xn[0], xn[1], xn[2] = 1, 1, 3
        xn[4:8] = np.frombuffer(np.float32(500 + 1500 * rng.random()).tobytes(), np.uint8)
        xn[8:12] = np.frombuffer(np.float32(15 * rng.random()).tobytes(), np.uint8)
        out = []
        for mode in (0, 1):
            oracle.set_libm_f64(mode)
            st = np.zeros(1, L.XFEED_F32)
            oracle.lib.orc_xfeed_coeffs_f32(st.ctypes.data, C.addressof(xcfg), FS)
            out.append(np.array([st["lp_a0"][0], st["lp_b1"][0], st["ap_a"][0]], np.float64))
        oracle.set_libm_f64(0)
        worst = max(worst, float(np.max(np.abs(out[0] - out[1]) / np.maximum(np.abs(out[0]), 1e-30))))
    assert worst < 4 * 2.0 ** -23
