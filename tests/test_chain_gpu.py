"""GPU parity of the whole signal chain (BASELINE config 3 shape): the CUDA chain kernels through
the C ABI against the oracle's process_audio_packet() restatement, instance by instance.

Bars: S/PDIF words, PDM bitstream, peaks and clip flags bit-exact.  The leveller's per-block
log10f/powf are evaluated in double and rounded to float on both sides (oracle `libm_f64`,
DESIGN.md "libm policy")."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from dspi_b200 import api, layouts as L                                     # noqa: E402
from tests.chain_cases import chain_params, pcm_bytes                        # noqa: E402
from tests.orc import make_orc_chain, orc_chain_run                          # noqa: E402


def _compare(oracle, flavour, N, fs, bit_depth, n_packets, fpp, seed, leveller=True, uniform=False, calls=1):
    P, bq = chain_params(oracle, N, fs, seed, leveller=leveller, uniform=uniform)
    F = n_packets * fpp
    pcm = pcm_bytes(N, F * calls, bit_depth, seed + 1)
    bpf = 6 if bit_depth == 24 else 4
    oracle.set_libm_f64(1)
    eng = api.ChainEngine(flavour, N, max_frames=F)
    try:
        eng.set_params(P)
        eng.upload_biquads(bq)
        chains = [make_orc_chain(oracle, P[i], bq[i]) for i in range(N)]
        for call in range(calls):
            chunk = np.ascontiguousarray(pcm[:, call * F * bpf:(call + 1) * F * bpf])
            spdif, pdm, status = eng.process_host(chunk, bit_depth, n_packets, fpp)
            for i in range(N):
                ws, wp = orc_chain_run(oracle, flavour, chains[i], chunk[i], bit_depth, n_packets, fpp)
                assert np.array_equal(spdif[i], ws), f"instance {i} call {call}: S/PDIF words differ"
                if P[i]["matrix"]["outputs"][8]["enabled"]:
                    assert np.array_equal(pdm[i], wp), f"instance {i} call {call}: PDM bitstream differs"
                assert list(status[i]["peaks"]) == list(chains[i].peaks), f"instance {i}: peaks"
                assert int(status[i]["clip_flags"]) == int(chains[i].clip_flags), f"instance {i}: clip flags"
        got = eng.download_biquads()
        for i in range(N):
            want = np.frombuffer(bytes(chains[i].filters), L.BIQUAD_F32).reshape(11, 12)
            for name in ("s1", "s2", "svic1eq", "svic2eq"):
                assert np.array_equal(got[i][name].view(np.uint32), want[name].view(np.uint32)), f"instance {i}: filter state {name}"
    finally:
        eng.close()
        oracle.set_libm_f64(0)


@pytest.mark.parametrize("flavour", ["f32f", "f32s"])
@pytest.mark.parametrize("bit_depth", [16, 24])
def test_chain_matches_oracle(oracle, flavour, bit_depth):
    _compare(oracle, flavour, N=70, fs=96000.0, bit_depth=bit_depth, n_packets=12, fpp=96, seed=100)


def test_chain_uniform_config3_shape(oracle):
    """config 3 topology (all stages on, same topology everywhere): the warp-uniform fast paths"""
    _compare(oracle, "f32f", N=64, fs=96000.0, bit_depth=24, n_packets=10, fpp=96, seed=7, uniform=True)


@pytest.mark.parametrize("fpp", [1, 2, 47, 48, 192])
def test_chain_packet_sizes(oracle, fpp):
    """leveller block gain, delay write index and peaks depend on the packet length"""
    _compare(oracle, "f32f", N=33, fs=48000.0, bit_depth=16, n_packets=6, fpp=fpp, seed=11)


def test_chain_state_carries_across_calls(oracle):
    _compare(oracle, "f32f", N=40, fs=48000.0, bit_depth=16, n_packets=5, fpp=48, seed=21, calls=3)


def test_chain_without_leveller_is_libm_free(oracle):
    """no libm anywhere on the path: parity does not depend on the libm policy"""
    P, bq = chain_params(oracle, 32, 96000.0, 5, leveller=False)
    pcm = pcm_bytes(32, 4 * 96, 24, 6)
    eng = api.ChainEngine("f32f", 32, max_frames=4 * 96)
    try:
        eng.set_params(P)
        eng.upload_biquads(bq)
        spdif, pdm, status = eng.process_host(pcm, 24, 4, 96)
        for i in range(32):
            ch = make_orc_chain(oracle, P[i], bq[i])
            ws, wp = orc_chain_run(oracle, "f32f", ch, pcm[i], 24, 4, 96)
            assert np.array_equal(spdif[i], ws)
            if P[i]["matrix"]["outputs"][8]["enabled"]:
                assert np.array_equal(pdm[i], wp)
    finally:
        eng.close()


def test_chain_24bit_odd_frame_count(oracle):
    """44.1 kHz-like packets: 45 frames x 7 packets = 315 frames of packed 24-bit PCM - instance streams are not
    word-aligned (byte path of the pre stage) and slices do not start on 16-byte boundaries (K1's non-TMA path)"""
    _compare(oracle, "f32f", N=35, fs=44100.0, bit_depth=24, n_packets=7, fpp=45, seed=41)


def test_chain_call_longer_than_the_delay_ring(oracle):
    """4608 frames per call > 4096-slot delay rings, two calls: delayed samples come from the output rows inside a call
    and from the ring across calls; the ring keeps exactly the last 4096 samples"""
    _compare(oracle, "f32f", N=6, fs=96000.0, bit_depth=16, n_packets=24, fpp=192, seed=51, calls=2)
