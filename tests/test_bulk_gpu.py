"""GPU: a WireBulkParams packet taken all the way - dspi_bulk_params_apply -> dspi_bulk_state_to_chain_* ->
dspi_chain(q)_set_params / upload_biquads -> process - against the oracle's process_audio_packet()
restatement fed with the same derived records."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from dspi_b200 import api, layouts as L                                     # noqa: E402
from tests.bulk_cases import wire_packet                                     # noqa: E402
from tests.chain_cases import pcm_bytes                                      # noqa: E402
from tests.orc import make_orc_chain, make_orc_chain_q28, orc_chain_run, orc_chain_run_q28   # noqa: E402


def _instances(platform, n, fs, vol):
    q28 = platform == L.PLATFORM_RP2040
    Ps = np.zeros(n, L.CHAIN_PARAMS_Q28 if q28 else L.CHAIN_PARAMS_F32)       # (np.concatenate would re-pack the padded C layouts)
    bqs = np.zeros((n, L.CHAINQ_EQ_CHANNELS if q28 else L.CHAIN_EQ_CHANNELS, L.MAX_BANDS), L.BIQUAD_Q28 if q28 else L.BIQUAD_F32)
    for i in range(n):
        st = api.bulk_state_defaults(platform)
        w = wire_packet(platform, 300 + i)
        w["master_volume"][0]["master_volume_db"] = np.float32(-3.0 * (i % 4))
        w["outputs"][0]["enabled"][:] = 1                     # audible configuration: every output on, most unmuted
        w["outputs"][0]["mute"][:] = 0
        w["crosspoints"][0]["enabled"][:] = 1
        assert api.bulk_params_apply(w, st) == 0
        P, bq = api.bulk_state_to_chain(st, fs, vol)
        Ps[i] = P[0]
        bqs[i] = bq[0]
    return Ps, bqs


def test_float_chain_from_wire_packets(oracle):
    N, fs, npk, fpp = 24, 96000.0, 6, 96
    P, bq = _instances(L.PLATFORM_RP2350, N, fs, -12 * 256)
    pcm = pcm_bytes(N, npk * fpp, 24, 9)
    oracle.set_libm_f64(1)
    eng = api.ChainEngine("f32f", N, max_frames=npk * fpp)
    try:
        eng.set_params(P)
        eng.upload_biquads(bq)
        spdif, pdm, status = eng.process_host(pcm, 24, npk, fpp)
        for i in range(N):
            ch = make_orc_chain(oracle, P[i], bq[i])
            ws, wp = orc_chain_run(oracle, "f32f", ch, pcm[i], 24, npk, fpp)
            if not np.array_equal(spdif[i], ws):
                bad = np.argwhere(spdif[i] != ws)
                pair, frame, chn = bad[0]
                o = 2 * pair + chn
                raise AssertionError(f"instance {i}: S/PDIF words differ first at pair {pair} frame {frame} ch {chn} ({len(bad)} words, outputs "
                                     f"{sorted(set((2 * b[0] + b[2]) for b in bad))}); output {o}: {P[i]['matrix']['outputs'][o]} "
                                     f"xp {P[i]['matrix']['crosspoints'][:, o]} flags bypass={P[i]['bypass_master_eq']} loud={P[i]['loudness_enabled']} "
                                     f"xf={P[i]['crossfeed_enabled']} lev={P[i]['leveller_enabled']} gpu {spdif[i][pair, frame, chn]} want {ws[pair, frame, chn]}")
            if P[i]["matrix"]["outputs"][8]["enabled"]:
                assert np.array_equal(pdm[i], wp), f"instance {i}: PDM differs"
            assert list(status[i]["peaks"]) == list(ch.peaks)
    finally:
        eng.close()
        oracle.set_libm_f64(0)


def test_q28_chain_from_wire_packets(oracle):
    N, fs, npk, fpp = 24, 48000.0, 8, 48
    P, bq = _instances(L.PLATFORM_RP2040, N, fs, -20 * 256)
    pcm = pcm_bytes(N, npk * fpp, 16, 10)
    oracle.set_libm_f64(1)
    eng = api.ChainEngineQ28(N, max_frames=npk * fpp)
    try:
        eng.set_params(P)
        eng.upload_biquads(bq)
        spdif, pdm, status = eng.process_host(pcm, 16, npk, fpp)
        for i in range(N):
            ch = make_orc_chain_q28(oracle, P[i], bq[i])
            ws, wp = orc_chain_run_q28(oracle, ch, pcm[i], 16, npk, fpp)
            assert np.array_equal(spdif[i], ws), f"instance {i}: S/PDIF words differ"
            if P[i]["matrix"]["outputs"][4]["enabled"]:
                assert np.array_equal(pdm[i], wp), f"instance {i}: PDM differs"
    finally:
        eng.close()
        oracle.set_libm_f64(0)
