"""Builds ``dspi_b200/libdspi_b200.so`` in-tree: hand-written sm_100a kernels + the C ABI.

nvcc cross-compiles without a GPU; the built library travels to the GPU box with the
repository snapshot (it is git-ignored, not gpurun-ignored).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdspi_b200.so")
CU_SOURCES = ["engine.cu", "eq_f32.cu", "eq_q28.cu", "chain_f32.cu", "chain_q28.cu"]
C_SOURCES = ["host_params.c"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-ftz=true",            # firmware runs FPSCR FZ (main.c:593-600)
    "-prec-div=true", "-prec-sqrt=true",
    "-fmad=false",          # never contract; fusing is spelled out with explicit intrinsics
    "-Xcompiler", "-fPIC",
]


def _newer(target, deps):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force=False, verbose=False):
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "dspi_b200.h"), os.path.abspath(__file__)]
    if not force and _newer(LIB, deps):
        return LIB
    inc = ["-I", os.path.join(ROOT, "include"), "-I", CSRC]
    objs = []
    bdir = os.path.join(HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    procs = []
    for src in C_SOURCES:
        obj = os.path.join(bdir, src + ".o")
        cmd = ["gcc", "-std=gnu11", "-O2", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-Wall"] + inc + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd)))
        objs.append(obj)
    for src in CU_SOURCES:
        obj = os.path.join(bdir, src + ".o")
        cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + inc + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd)))
        objs.append(obj)
    for cmd, p in procs:
        if p.wait() != 0:
            raise RuntimeError("build failed: " + " ".join(cmd))
    cmd = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-lm"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
