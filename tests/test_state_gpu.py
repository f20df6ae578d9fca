"""GPU: checkpoint / resume of the chain engines (dspi_chain(q)_state_export / _import): a run continued from an imported
blob - in the same engine after other work, and in a second engine of the same shape - produces the same bits."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from dspi_b200 import api                                                    # noqa: E402
from tests.chain_cases import chain_params, chain_params_q28, pcm_bytes      # noqa: E402


def _run(q28, oracle):
    N, fs, npk, fpp, bits = 20, 96000.0, 5, 96, 24
    F = npk * fpp
    P, bq = (chain_params_q28(oracle, N, fs, 61) if q28 else chain_params(oracle, N, fs, 60))
    a, b, cpcm = pcm_bytes(N, F, bits, 1), pcm_bytes(N, F, bits, 2), pcm_bytes(N, F, bits, 3)

    def make():
        e = api.ChainEngineQ28(N, max_frames=F) if q28 else api.ChainEngine("f32f", N, max_frames=F)
        e.set_params(P)
        e.upload_biquads(bq)
        return e
    e1 = make()
    try:
        e1.process_host(a, bits, npk, fpp)
        blob = e1.state_export()
        want = e1.process_host(b, bits, npk, fpp)                 # the continuation to reproduce
        e1.process_host(cpcm, bits, npk, fpp)                     # unrelated work changes every state
        e1.state_import(blob)
        again = e1.process_host(b, bits, npk, fpp)
        e2 = make()
        try:
            e2.state_import(blob)
            other = e2.process_host(b, bits, npk, fpp)
            with pytest.raises(api.DspiError):
                e2.state_import(blob[:-8])
        finally:
            e2.close()
    finally:
        e1.close()
    for got in (again, other):
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
        assert got[2].tobytes() == want[2].tobytes()
    assert not np.array_equal(want[0], 0 * want[0])


def test_float_chain_checkpoint(oracle):
    _run(False, oracle)


def test_q28_chain_checkpoint(oracle):
    _run(True, oracle)
