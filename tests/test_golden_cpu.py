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
