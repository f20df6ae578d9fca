"""The oracle against the committed fixtures (generated from the compiled reference by
tests/golden/make_golden.py).  Runs anywhere, including the GPU box without /root/reference."""
import numpy as np
import pytest

from dspi_b200 import layouts as L
from tests.util import load_golden, same_bits


@pytest.mark.parametrize("flavour", ["f32s", "f32f", "q28"])
@pytest.mark.parametrize("signal", ["impulse", "sine", "sweep", "noise"])
def test_cfg1(oracle, flavour, signal):
    g = load_golden("cfg1.npz")
    q = flavour == "q28"
    bq = g["bq_q28" if q else "bq_f32"].copy()
    x = g[f"{signal}_q28_x" if q else f"{signal}_x"].copy()
    oracle.eq_many(flavour, bq, x, 10, 48)
    want = g[f"{signal}_{flavour}_y"]
    assert np.array_equal(x.view(np.uint32), want.view(np.uint32))
    assert same_bits(bq, g[f"{signal}_{flavour}_state"])


@pytest.mark.parametrize("flavour", ["f32s", "f32f", "q28"])
def test_mix(oracle, flavour):
    g = load_golden("mix.npz")
    q = flavour == "q28"
    bq = g["bq_q28" if q else "bq_f32"].copy()
    x = g["q28_x" if q else "x"].copy()
    oracle.eq_many(flavour, bq, x, 10, 96)
    assert np.array_equal(x.view(np.uint32), g[f"{flavour}_y"].view(np.uint32))
    assert same_bits(bq, g[f"{flavour}_state"])


def test_spdif_fixture(oracle):
    g = load_golden("spdif.npz")
    got = oracle.spdif_encode(g["words"], int(g["pos0"]), bytes(g["cs"]))
    assert np.array_equal(got, g["subframes"])


def test_cfg1_known_answers():
    """KATs the reference states in comments: flat band => bypass with b0 = 1; path split at fs/7.5."""
    g = load_golden("cfg1.npz")
    bq = g["bq_f32"]
    assert bq["use_svf"][0, 0] == 1 and bq["use_svf"][0, 1] == 1 and bq["use_svf"][0, 2] == 0   # 10 kHz >= 48k/7.5
    assert np.all(bq["bypass"][:, 3:] == 1) and np.all(bq["b0"][:, 3:] == 1.0)
    assert np.all(g["bq_q28"]["b0"][:, 3:] == 1 << 28)


@pytest.mark.parametrize("flavour", ["f32s", "f32f", "q28"])
def test_chain_fixture(oracle, flavour):
    """Vectors made by the reference's own process_audio_packet() + modulator loop (tests/golden/make_golden_chain.py)."""
    from tests.orc import make_orc_chain, make_orc_chain_q28, orc_chain_run, orc_chain_run_q28
    g = load_golden("chain.npz")
    P, bq, pcm = g[f"{flavour}_params"], g[f"{flavour}_biquads"], g[f"{flavour}_pcm"]
    npk, fpp = int(g["n_packets"]), int(g["fpp"])
    for i in range(len(P)):
        if flavour == "q28":
            ch = make_orc_chain_q28(oracle, P[i], bq[i])
            sp, pdm = orc_chain_run_q28(oracle, ch, pcm[i], 24, npk, fpp)
        else:
            ch = make_orc_chain(oracle, P[i], bq[i])
            sp, pdm = orc_chain_run(oracle, flavour, ch, pcm[i], 24, npk, fpp)
        assert np.array_equal(sp, g[f"{flavour}_spdif"][i]) and np.array_equal(pdm, g[f"{flavour}_pdm"][i])
        assert list(ch.peaks)[:g[f"{flavour}_peaks"].shape[1]] == list(g[f"{flavour}_peaks"][i]) and int(ch.clip_flags) == int(g[f"{flavour}_clip"][i])
