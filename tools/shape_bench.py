#!/usr/bin/env python3
"""K1 / K2 device time for the epoch shapes of the BASELINE configs (synthetic rows, tables resident in HBM).
usage: python tools/shape_bench.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gps_sdr_sim_b200 as gs

SHAPES = [  # label, samples per epoch, epochs, channels, format, carrier mode
    ("config 1  static 2.6 MS/s 16-bit, 30 s", 260000, 299, 11, 16, 0),
    ("config 2  circle 2.6 MS/s 8-bit, 300 s", 260000, 2999, 13, 8, 0),
    ("config 3  spacecraft 2.6 MS/s 16-bit, 300 s", 260000, 2999, 13, 16, 0),
    ("config 4  NMEA 1 MS/s 1-bit, 156 s", 100000, 1560, 10, 1, 0),
    ("config 5  static 20 MS/s 16-bit, 102.4 s batch", 2000000, 1024, 11, 16, 0),
    ("config 1  as shipped (double carrier)", 260000, 299, 11, 16, 1),
    ("config 2  as shipped (double carrier)", 260000, 2999, 13, 8, 1),
    ("config 4  as shipped (double carrier)", 100000, 1560, 10, 1, 1),
]
for label, n, e, c, fmt, mode in SHAPES:
    t = gs.synthetic_table(e, n, c, fmt, carrier_mode=mode)
    out = torch.empty(t.n_epochs * t.epoch_bytes, dtype=torch.uint8, device="cuda")
    with gs.GpuSim.for_table(t) as sim:
        sim.set_option("pipeline", 0)
        sim.upload_table(t)
        best = None
        for _ in range(5):
            sim.generate_device(0, e, out.data_ptr(), out.numel())
            tm = sim.timing()
            if best is None or tm.synth_ms + tm.chain_ms < best[0] + best[1]:
                best = (tm.chain_ms, tm.synth_ms)
    tot = best[0] + best[1]
    print(f"{label:48s} K1 {best[0]:7.3f} ms  K2 {best[1]:8.3f} ms  {e * n / tot / 1e6:7.1f} GS/s  "
          f"{e * 0.1 / (tot / 1e3):9.0f} x real time  {e * n * c / best[1] / 1e6:8.1f} G sample-channels/s in K2", flush=True)
