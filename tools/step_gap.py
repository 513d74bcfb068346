#!/usr/bin/env python3
"""Back-to-back generate_device steps on the bench shape: ms per step next to the K2 duration, for a caller's
stream of default and of high priority (is the gap between steps the chain kernel's blocks taking SM slots
before the next synthesis kernel's blocks are placed?).  usage: python tools/step_gap.py [steps] [carrier_mode]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gps_sdr_sim_b200 as gs

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0      # 0 integer carrier, 1 double carrier (FLOAT_CARR_PHASE hosts)
E, N = 2999, 260000
t = gs.synthetic_table(E, N, 13, 8, carrier_mode=mode)
out = torch.empty(t.n_epochs * t.epoch_bytes, dtype=torch.uint8, device="cuda")
for prio in (0, -1):
    for pipeline in (1, 0, 2):
        stream = torch.cuda.Stream(priority=prio)
        with gs.GpuSim.for_table(t) as sim:
            sim.set_option("pipeline", pipeline)
            sim.upload_table(t)
            for _ in range(3):
                sim.generate_device(0, E, out.data_ptr(), out.numel(), stream=stream.cuda_stream)
            stream.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(steps):
                sim.generate_device(0, E, out.data_ptr(), out.numel(), stream=stream.cuda_stream)
            e1.record(stream)
            stream.synchronize()
            tm = sim.timing()
            print(f"priority={prio:2d} pipeline={pipeline} step={e0.elapsed_time(e1) / steps:.3f} ms  k1={tm.chain_ms:.3f} k2={tm.synth_ms:.3f}", flush=True)
