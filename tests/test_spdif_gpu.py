"""GPU parity of the S/PDIF subframe encoder (spdif.cu) through the C ABI: bit-exact against the oracle
and the committed fixture generated from the reference's own header function."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from dspi_b200 import api                                   # noqa: E402
from tests.util import load_golden                          # noqa: E402

CS = bytes([0x04, 0x00, 0x00, 0x02, 0x0B])


def test_fixture_from_reference_header():
    g = load_golden("spdif.npz")
    got = api.spdif_encode_host(g["words"], int(g["pos0"]), bytes(g["cs"]))
    assert np.array_equal(got, g["subframes"])


@pytest.mark.parametrize("n_streams,frames,pos0", [(1, 1, 0), (5, 191, 1), (7, 192, 0), (33, 1000, 191), (64, 4099, 77)])
def test_matches_oracle(oracle, n_streams, frames, pos0):
    rng = np.random.default_rng(n_streams * 1000 + frames)
    words = rng.integers(-2**31, 2**31, (n_streams, frames, 2), dtype=np.int64).astype(np.int32)
    got = api.spdif_encode_host(words, pos0, CS)
    assert np.array_equal(got, oracle.spdif_encode(words, pos0, CS))


def test_device_path_on_chain_layout(oracle):
    """Device pointers, the [4 N][F][2] layout the chain writes, block position carried across two calls."""
    N, F = 6, 300
    rng = np.random.default_rng(5)
    words = rng.integers(-2**23, 2**23, (4 * N, 2 * F, 2), dtype=np.int64).astype(np.int32)
    want = oracle.spdif_encode(words, 10, CS)
    out = torch.empty((4 * N, F, 2, 2), dtype=torch.int32, device="cuda")
    for part in range(2):
        w = torch.from_numpy(np.ascontiguousarray(words[:, part * F:(part + 1) * F])).cuda()
        api.spdif_encode_device(w.data_ptr(), 4 * N, F, out.data_ptr(), block_pos0=10 + part * F, channel_status=CS)
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy().view(np.uint32), want[:, part * F:(part + 1) * F])


def test_rejects_bad_arguments():
    with pytest.raises(api.DspiError):
        api.spdif_encode_device(0, 1, 1, 0)
    w = torch.zeros((1, 4, 2), dtype=torch.int32, device="cuda")
    o = torch.zeros((1, 4, 2, 2), dtype=torch.int32, device="cuda")
    with pytest.raises(api.DspiError):
        api.spdif_encode_device(w.data_ptr(), 1, 0, o.data_ptr())
    with pytest.raises(api.DspiError):
        api.spdif_encode_device(w.data_ptr() + 4, 1, 2, o.data_ptr())
