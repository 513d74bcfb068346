#!/usr/bin/env python3
"""One short K1+K2 job for ncu: python tools/profile_one.py [fmt] [accum] [epochs] [chunk] [carrier_mode] [pipeline]

pipeline defaults to 2: the synthesis kernel is the build that shares the SM with the next call's chain
kernel (112 registers) - the one bench.py's back-to-back steps run."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gps_sdr_sim_b200 as gs

fmt = int(sys.argv[1]) if len(sys.argv) > 1 else 8
accum = int(sys.argv[2]) if len(sys.argv) > 2 else 1
E = int(sys.argv[3]) if len(sys.argv) > 3 else 600
chunk = int(sys.argv[4]) if len(sys.argv) > 4 else 0
mode = int(sys.argv[5]) if len(sys.argv) > 5 else 0
pipeline = int(sys.argv[6]) if len(sys.argv) > 6 else 2
t = gs.synthetic_table(E, 260000, 13, fmt, carrier_mode=mode)
out = torch.empty(t.n_epochs * t.epoch_bytes, dtype=torch.uint8, device="cuda")
with gs.GpuSim.for_table(t) as sim:
    sim.set_option("accum", accum)
    sim.set_option("chunk", chunk)
    sim.set_option("pipeline", pipeline)
    sim.upload_table(t)
    for _ in range(3):
        sim.generate_device(0, E, out.data_ptr(), out.numel())
        tm = sim.timing()
    print(f"pipeline={pipeline} carrier_mode={mode} fmt={fmt} accum={accum} epochs={E} chunk={chunk} k1={tm.chain_ms:.3f} ms k2={tm.synth_ms:.3f} ms "
          f"{E * 260000 / tm.synth_ms / 1e6:.1f} GS/s")
