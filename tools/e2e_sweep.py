#!/usr/bin/env python3
"""e2e (host table in, host bytes out) time of the bench workload for several sub-batch schedules of the
direct-to-caller-buffer path.  usage: python tools/e2e_sweep.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gps_sdr_sim_b200 as gs

E, N = 2999, 260000
t = gs.synthetic_table(E, N, 13, gs.SC08)
nbytes = t.n_epochs * t.epoch_bytes
host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
for _ in range(2):
    host.copy_(dev, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    host.copy_(dev, non_blocking=True)
torch.cuda.synchronize()
raw = (time.perf_counter() - t0) / 5
print(f"bare pinned D2H of {nbytes/1e9:.2f} GB: {raw*1e3:.2f} ms = {nbytes/raw/1e9:.1f} GB/s", flush=True)
for first_mb, mb in ((64, 64), (16, 64), (8, 128), (8, 256), (4, 512), (16, 2048), (2048, 2048)):
    with gs.GpuSim.for_table(t) as sim:
        sim.set_option("direct_first_mb", first_mb)
        sim.set_option("direct_mb", mb)
        for _ in range(2):
            sim.generate_epochs(t, out_ptr=host.data_ptr(), out_capacity=host.numel())
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            sim.generate_epochs(t, out_ptr=host.data_ptr(), out_capacity=host.numel())
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 10
        t1 = time.perf_counter()
        for _ in range(10):
            sim.upload_table(t)
        up = (time.perf_counter() - t1) / 10
    print(f"first {first_mb:5d} MiB, then {mb:5d} MiB: {dt*1e3:7.2f} ms/step = {nbytes/dt/1e9:5.1f} GB/s ({raw/dt*100:.1f} % of the bare copy); upload_table alone {up*1e3:.2f} ms", flush=True)
