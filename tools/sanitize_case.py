#!/usr/bin/env python3
"""A small job that touches every kernel variant, for compute-sanitizer:
   compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import gps_sdr_sim_b200 as gs

for mode in (0, 1):
    for fmt, n, nact in ((16, 26000, 13), (8, 26000, 16), (1, 10000, 9), (16, 12345, 5)):
        t = gs.synthetic_table(3, n, nact, fmt, seed=fmt + n, carrier_mode=mode)
        for opts in ({}, {"force_slow": 1}, {"layout": 1, "chunk": 128}, {"force_generic": 1}, {"accum": 0}):
            if mode == 1 and "accum" in opts:
                continue
            with gs.GpuSim.for_table(t) as sim:
                for k, v in opts.items():
                    sim.set_option(k, v)
                out = sim.generate_epochs(t)
                got = []
                sim.generate_epochs_to_sink(t, lambda mv: got.append(bytes(mv)))
                assert np.array_equal(out, np.frombuffer(b"".join(got), dtype=np.uint8))
print("sanitize_case ok")
