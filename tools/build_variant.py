#!/usr/bin/env python3
"""Build an experimental copy of libgpusim.so with extra -D flags: build/variants/libgpusim_<name>.so
(load it with GPUSIM_LIB=<path>).  usage: tools/build_variant.py <name> [-DGS_X=1 ...]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gps_sdr_sim_b200 import build as b

name, flags = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(ROOT, "variants")
os.makedirs(out_dir, exist_ok=True)
out = os.path.join(out_dir, f"libgpusim_{name}.so")
cmd = ["/usr/local/cuda/bin/nvcc", *b.NVCC_FLAGS, *flags, "-I", os.path.join(ROOT, "include"), "-I", b.CSRC,
       *[os.path.join(b.CSRC, s) for s in b.SOURCES], "-o", out]
r = subprocess.run(cmd, capture_output=True, text=True)
log = r.stdout + r.stderr
open(out + ".log", "w").write(log)
if r.returncode != 0:
    sys.stderr.write(log)
    raise SystemExit(1)
print(out)
