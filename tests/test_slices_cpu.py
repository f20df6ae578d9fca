"""The packet-slice plan of the chain engines (ChainStreams::plan_slices, dspi_b200/csrc/chain_streams.cuh) as a host-only
program: whatever the call length, the slices tile the packets exactly, there are at most kMaxSlices of them, and calls long
enough to matter start and end with short slices (the modulator starts early and has little left at the end).  Results never
depend on the plan (tests/test_chain_ref_gpu.py checks that on the GPU); this checks the plan itself."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def plans(tmp_path_factory):
    if not shutil.which("nvcc"):
        pytest.skip("nvcc not on PATH")
    exe = str(tmp_path_factory.mktemp("slices") / "slices")
    subprocess.run(["nvcc", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "dspi_b200", "csrc"), "-o", exe,
                    os.path.join(ROOT, "tests", "c_client", "slices.cu")], check=True, capture_output=True, timeout=600)
    env = {k: v for k, v in os.environ.items() if k != "DSPI_UNIFORM_SLICES"}
    out = subprocess.run([exe], check=True, capture_output=True, text=True, timeout=60, env=env).stdout
    res = {}
    for line in out.splitlines():
        n, rest = line.split(":")
        res[int(n)] = [int(v) for v in rest.split()]
    return res


def test_slices_tile_every_call_length(plans):
    assert sorted(plans) == list(range(1, 301))
    for n, b in plans.items():
        assert b[0] == 0 and b[-1] == n, (n, b)
        assert all(x < y for x, y in zip(b, b[1:])), (n, b)                  # no empty slice
        assert 1 <= len(b) - 1 <= 16, (n, b)


def test_long_calls_are_short_at_both_ends(plans):
    for n in (24, 64, 128, 256, 300):
        sizes = [y - x for x, y in zip(plans[n], plans[n][1:])]
        assert sizes[0] == 1 and sizes[-1] == 1, (n, sizes)
        assert sizes[1] <= 2 and sizes[-2] <= 2, (n, sizes)
        assert max(sizes) >= 4, (n, sizes)
    assert [y - x for x, y in zip(plans[64], plans[64][1:])][:3] == [1, 2, 4]
