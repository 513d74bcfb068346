#!/usr/bin/env python3
"""K1 (chain kernel) time against the number of checkpoints per chain: plain layout, chunk = 128 .. 8192.
usage: python tools/k1_probe.py [epochs] [carrier_mode]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gps_sdr_sim_b200 as gs

E = int(sys.argv[1]) if len(sys.argv) > 1 else 2999
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
t = gs.synthetic_table(E, 260000, 13, 8, carrier_mode=mode)
out = torch.empty(E * t.epoch_bytes, dtype=torch.uint8, device="cuda")
for chunk in (0, 128, 256, 512, 1024, 2048, 4096, 8192):
    with gs.GpuSim.for_table(t) as sim:
        sim.set_option("pipeline", 0)
        if chunk:
            sim.set_option("layout", 1)
            sim.set_option("chunk", chunk)
        sim.upload_table(t)
        best = 1e9
        for _ in range(4):
            sim.generate_device(0, E, out.data_ptr(), out.numel())
            tm = sim.timing()
            best = min(best, tm.chain_ms)
    print(f"chunk={chunk or 'aligned(520)':>12}  k1={best:.3f} ms  k2={tm.synth_ms:.3f} ms", flush=True)
