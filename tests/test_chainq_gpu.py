"""GPU parity of the Q28 (RP2040) signal chain: bit-exact S/PDIF words, PDM bits, peaks, clip flags
and filter state against the oracle's restatement of usb_audio.c:968-1283."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from dspi_b200 import api, layouts as L                                     # noqa: E402
from tests.chain_cases import chain_params_q28, pcm_bytes                    # noqa: E402
from tests.orc import make_orc_chain_q28, orc_chain_run_q28                  # noqa: E402


def _compare(oracle, N, fs, bit_depth, n_packets, fpp, seed, leveller=True, calls=1):
    P, bq = chain_params_q28(oracle, N, fs, seed, leveller=leveller)
    F = n_packets * fpp
    pcm = pcm_bytes(N, F * calls, bit_depth, seed + 1)
    bpf = 6 if bit_depth == 24 else 4
    oracle.set_libm_f64(1)
    eng = api.ChainEngineQ28(N, max_frames=F)
    try:
        eng.set_params(P)
        eng.upload_biquads(bq)
        chains = [make_orc_chain_q28(oracle, P[i], bq[i]) for i in range(N)]
        for call in range(calls):
            chunk = np.ascontiguousarray(pcm[:, call * F * bpf:(call + 1) * F * bpf])
            spdif, pdm, status = eng.process_host(chunk, bit_depth, n_packets, fpp)
            for i in range(N):
                ws, wp = orc_chain_run_q28(oracle, chains[i], chunk[i], bit_depth, n_packets, fpp)
                assert np.array_equal(spdif[i], ws), f"instance {i} call {call}: S/PDIF words differ"
                if P[i]["matrix"]["outputs"][4]["enabled"]:
                    assert np.array_equal(pdm[i], wp), f"instance {i} call {call}: PDM bitstream differs"
                assert list(status[i]["peaks"]) == list(chains[i].peaks)[:7], f"instance {i}: peaks"
                assert int(status[i]["clip_flags"]) == int(chains[i].clip_flags), f"instance {i}: clip flags"
        got = eng.download_biquads()
        for i in range(N):
            for r in range(7):
                want = np.frombuffer(bytes(chains[i].filters[r]), L.BIQUAD_Q28)
                assert np.array_equal(got[i, r]["s1"], want["s1"]) and np.array_equal(got[i, r]["s2"], want["s2"]), f"instance {i} role {r}: state"
    finally:
        eng.close()
        oracle.set_libm_f64(0)


@pytest.mark.parametrize("bit_depth", [16, 24])
def test_chainq_matches_oracle(oracle, bit_depth):
    _compare(oracle, N=70, fs=96000.0, bit_depth=bit_depth, n_packets=12, fpp=96, seed=300)


@pytest.mark.parametrize("fpp", [1, 2, 47, 48, 192])
def test_chainq_packet_sizes(oracle, fpp):
    _compare(oracle, N=33, fs=48000.0, bit_depth=16, n_packets=6, fpp=fpp, seed=311)


def test_chainq_state_carries_across_calls(oracle):
    _compare(oracle, N=40, fs=48000.0, bit_depth=24, n_packets=5, fpp=48, seed=321, calls=3)


def test_chainq_without_leveller(oracle):
    _compare(oracle, N=32, fs=96000.0, bit_depth=24, n_packets=4, fpp=96, seed=331, leveller=False)


def test_chainq_24bit_odd_frame_count(oracle):
    _compare(oracle, N=35, fs=44100.0, bit_depth=24, n_packets=7, fpp=45, seed=341)


def test_chainq_call_longer_than_the_delay_ring(oracle):
    """2880 frames per call > 2048-slot rings, two calls"""
    _compare(oracle, N=6, fs=96000.0, bit_depth=16, n_packets=15, fpp=192, seed=351, calls=2)
