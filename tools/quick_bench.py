#!/usr/bin/env python3
"""Kernel timing sweep (CUDA events inside the library) over accumulator / chunk / format options.
usage: python tools/quick_bench.py [epochs]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gps_sdr_sim_b200 as gs

E = int(sys.argv[1]) if len(sys.argv) > 1 else 600
N = 260000
print(torch.cuda.get_device_name(0), "epochs", E, flush=True)
for fmt in (16, 8, 1):
    t = gs.synthetic_table(E, N, 13, fmt)
    out = torch.empty(t.n_epochs * t.epoch_bytes, dtype=torch.uint8, device="cuda")
    for accum in (1, 0):
        for layout, chunk, slow in ((0, 0, 0), (0, 0, 1), (1, 128, 0), (1, 256, 0), (1, 512, 0)):
            if accum == 0 and layout == 1:
                continue
            with gs.GpuSim.for_table(t) as sim:
                sim.set_option("accum", accum)
                sim.set_option("layout", layout)
                sim.set_option("chunk", chunk)
                sim.set_option("force_slow", slow)
                sim.upload_table(t)
                best = None
                for _ in range(4):
                    sim.generate_device(0, E, out.data_ptr(), out.numel())
                    tm = sim.timing()
                    if best is None or tm.synth_ms < best[1]:
                        best = (tm.chain_ms, tm.synth_ms)
                samples = E * N
                print(f"fmt={fmt:2d} accum={accum} layout={layout} chunk={chunk:4d} force_slow={slow}  k1={best[0]:7.3f} ms  k2={best[1]:7.3f} ms  "
                      f"{samples / best[1] / 1e6:8.1f} GS/s  {t.n_epochs * t.epoch_bytes / best[1] / 1e6:7.1f} GB/s", flush=True)
# generic kernel and replay chain for scale
t = gs.synthetic_table(60, N, 13, 16)
out = torch.empty(t.n_epochs * t.epoch_bytes, dtype=torch.uint8, device="cuda")
for key in ("force_generic", "chain_replay"):
    with gs.GpuSim.for_table(t) as sim:
        sim.set_option(key, 1)
        sim.upload_table(t)
        for _ in range(2):
            sim.generate_device(0, 60, out.data_ptr(), out.numel())
            tm = sim.timing()
        print(f"{key}: 60 epochs k1={tm.chain_ms:.3f} ms k2={tm.synth_ms:.3f} ms", flush=True)
