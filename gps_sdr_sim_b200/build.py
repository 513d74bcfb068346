"""Build gps_sdr_sim_b200/libgpusim.so (CUDA, sm_100a only) in-tree with nvcc."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgpusim.so")
SOURCES = ["gpusim_api.cu", "gpusim_kernels.cu", "gpusim_tables.cpp"]
HEADERS = ["gpusim_core.h", "gpusim_kernels.h", "gpusim_tables.h", os.path.join(ROOT, "include", "gpusim.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",   # Blackwell B200 only
    "-O3", "-lineinfo", "-std=c++17",
    "--fmad=false",                                 # never contract the reference's mul+add
    "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall,-Wno-unknown-pragmas",
    "-Xptxas", "-v",
    "-cudart", "static",
    "-shared",
]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *NVCC_FLAGS, "-I", os.path.join(ROOT, "include"), "-I", CSRC,
           *[os.path.join(CSRC, s) for s in SOURCES], "-o", LIB + ".tmp"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libgpusim.so")
    os.replace(LIB + ".tmp", LIB)      # atomically: a snapshot of the tree never sees a half-written library
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
