"""The EQ engine over several GPUs of one box (dspi_eqx_*, SURVEY §8b `devices[]`, §8e): sharding must not change a bit.
On-hardware form of the invariant that tests/test_sharding_cpu.py checks over gloo: outputs and filter state of a group of
1, 2, ... devices are byte-identical to one engine over all channels - for the host-block path (every device its own PCIe
pipeline) and for the root-block path (peers pull / push their rows over NVLink).  Cases with more than one device skip on
a single-GPU box."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from dspi_b200 import api, workloads as W                  # noqa: E402
from tests.util import same_bits                           # noqa: E402

FS = 96000.0


def _single(flavour, bq, x):
    Cn, T = x.shape
    eng = api.EqEngine(flavour, Cn)
    try:
        eng.upload(bq)
        buf = torch.from_numpy(x).cuda()
        eng.process_device(buf.data_ptr(), T, T)
        eng.sync()
        return buf.cpu().numpy(), eng.download()
    finally:
        eng.close()


def _case(flavour, Cn, T, seed):
    q = flavour == "q28"
    params = W.eq_params("mixed" if not q else "B", Cn, fs=FS, seed=seed)
    bq = api.compute_coefficients(params, q28=q, fs=FS)
    x = W.inputs_q28(Cn, T) if q else W.inputs_f32(Cn, T)
    return bq, x


@pytest.mark.parametrize("n_dev", [1, 2, 4, 8])
@pytest.mark.parametrize("flavour", ["f32f", "q28"])
def test_group_equals_single_engine(flavour, n_dev):
    if torch.cuda.device_count() < n_dev:
        pytest.skip(f"needs {n_dev} GPUs")
    Cn, T = 64 * 37, 480                                   # 37 warp groups: uneven shards
    bq, x = _case(flavour, Cn, T, 11)
    want, wst = _single(flavour, bq, x)
    grp = api.EqGroup(flavour, Cn, list(range(n_dev)))
    try:
        edges = [grp.shard_range(k) for k in range(n_dev)]
        assert edges[0][0] == 0 and edges[-1][1] == Cn and all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
        grp.upload(bq)
        # host block: pinned [C][T], in place
        pin = api.PinnedBuffer((Cn, T), x.dtype)
        pin.array[...] = x
        grp.process_host(pin.array)
        assert np.array_equal(pin.array.view(np.uint32), want.view(np.uint32)), "host path differs from the single engine"
        assert same_bits(grp.download(), wst)
        pin.free()
        # root block: resident on device 0, second pass continues from the state the first left
        want2, wst2 = want.copy(), None
        eng = api.EqEngine(flavour, Cn)
        eng.upload(wst)
        b2 = torch.from_numpy(want).cuda()
        eng.process_device(b2.data_ptr(), T, T)
        eng.sync()
        want2, wst2 = b2.cpu().numpy(), eng.download()
        eng.close()
        with torch.cuda.device(0):
            d = torch.from_numpy(want).cuda()
            torch.cuda.synchronize()
            grp.process_root(d.data_ptr(), T)
            got2 = d.cpu().numpy()
        assert np.array_equal(got2.view(np.uint32), want2.view(np.uint32)), "root path differs from the single engine"
        assert same_bits(grp.download(), wst2)
        assert grp.launch_count >= 2 * n_dev
    finally:
        grp.close()


def test_range_call_equals_whole_call():
    """dspi_eq_process_device_range over pieces == one call over all channels (what the NCCL pipeline relies on)."""
    Cn, T = 64 * 9, 384
    bq, x = _case("f32f", Cn, T, 5)
    want, wst = _single("f32f", bq, x)
    eng = api.EqEngine("f32f", Cn)
    try:
        eng.upload(bq)
        buf = torch.from_numpy(x).cuda()
        for a, b in [(128, 320), (0, 128), (320, Cn)]:
            eng.process_device_range(buf[a:b].data_ptr(), T, T, a, b - a)
        eng.sync()
        assert np.array_equal(buf.cpu().numpy().view(np.uint32), want.view(np.uint32))
        assert same_bits(eng.download(), wst)
        with pytest.raises(api.DspiError):
            eng.process_device_range(buf[32:].data_ptr(), T, T, 32, 64)     # not on a group boundary
    finally:
        eng.close()


def test_numa_binding_reports_a_node_or_declines():
    node = api.bind_host_to_device(0)
    assert node >= -1
