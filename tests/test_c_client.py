"""include/dspi_b200.h must stay a valid C11 header and the library must be usable from plain C:
tests/c_client/client.c is compiled with gcc -std=c11 -pedantic -Werror against the header and linked to the shared
library.  CPU run: host-side parameter API + DSPI_ENODEV.  GPU run: BASELINE config 1 through the engine, bit-exact
against the cascade written out in the client (strict float, gcc -ffp-contract=off)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_client", "client.c")
EXE = os.path.join(ROOT, "tests", "c_client", "client")


@pytest.fixture(scope="module")
def client():
    from dspi_b200 import api
    if not os.path.exists(api.LIB_PATH):
        from dspi_b200.build import build
        build()
    libdir = os.path.dirname(api.LIB_PATH)
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(SRC), os.path.getmtime(api.LIB_PATH)):
        subprocess.check_call(["gcc", "-std=c11", "-D_GNU_SOURCE", "-O1", "-ffp-contract=off", "-Wall", "-Wextra", "-Werror", "-pedantic",
                               "-I", os.path.join(ROOT, "include"), SRC, "-o", EXE, "-L", libdir, "-ldspi_b200", "-lm",
                               "-Wl,-rpath," + libdir])
    return EXE


def test_c_client_host_api(client):
    out = subprocess.run([client], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert "host api ok" in out.stdout


@pytest.mark.gpu
def test_c_client_runs_config1_on_the_gpu(client):
    out = subprocess.run([client, "gpu"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert "gpu ok" in out.stdout
