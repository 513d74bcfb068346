#!/usr/bin/env python3
"""One short K1+K2 job for ncu.

    python tools/profile_one.py [fmt] [accum] [epochs] [chunk] [carrier_mode] [pipeline] [samples_per_epoch] [channels]   synthetic rows
    python tools/profile_one.py --table <table.npz> [pipeline]                                 a saved EpochTable
    python tools/profile_one.py --scenario <name> [pipeline]                                   rows recorded from the reference host

Three generate_device calls on resident tables (ncu: --launch-skip past the first ones).  pipeline defaults
to 0: every call runs its chain kernel and then the synthesis kernel on the same stream."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gps_sdr_sim_b200 as gs

argv = sys.argv[1:]
accum, chunk = 1, 0
if argv and argv[0] == "--table":
    from gps_sdr_sim_b200.table import EpochTable
    t = EpochTable.load_npz(argv[1])
    pipeline = int(argv[2]) if len(argv) > 2 else 0
elif argv and argv[0] == "--scenario":
    from gps_sdr_sim_b200 import scenarios
    t, _ = scenarios.load(argv[1])
    pipeline = int(argv[2]) if len(argv) > 2 else 0
else:
    fmt = int(argv[0]) if len(argv) > 0 else 8
    accum = int(argv[1]) if len(argv) > 1 else 1
    E = int(argv[2]) if len(argv) > 2 else 600
    chunk = int(argv[3]) if len(argv) > 3 else 0
    mode = int(argv[4]) if len(argv) > 4 else 0
    pipeline = int(argv[5]) if len(argv) > 5 else 0
    n_samples = int(argv[6]) if len(argv) > 6 else 260000
    channels = int(argv[7]) if len(argv) > 7 else 13
    t = gs.synthetic_table(E, n_samples, channels, fmt, carrier_mode=mode)
E = t.n_epochs
out = torch.empty(E * t.epoch_bytes, dtype=torch.uint8, device="cuda")
with gs.GpuSim.for_table(t) as sim:
    sim.set_option("accum", accum)
    sim.set_option("chunk", chunk)
    sim.set_option("pipeline", pipeline)
    sim.upload_table(t)
    for _ in range(3):
        sim.generate_device(0, E, out.data_ptr(), out.numel())
        tm = sim.timing()
    print(f"pipeline={pipeline} carrier_mode={t.carrier_mode} fmt={t.data_format} epochs={E} N={t.samples_per_epoch} "
          f"k1={tm.chain_ms:.3f} ms k2={tm.synth_ms:.3f} ms {E * t.samples_per_epoch / tm.synth_ms / 1e6:.1f} GS/s")
