#!/usr/bin/env python3
"""Time K2 of several builds of the library (tools/build_variant.py) on the bench shape and check that they
all write the same bytes.  usage: python tools/variant_bench.py name[,name...] [epochs] [c5]
Each build runs in its own process (the library is loaded once per process)."""
import hashlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT)
    import torch
    import gps_sdr_sim_b200 as gs
    E = int(sys.argv[2])
    N, C, CASES = 260000, 13, ((0, 8), (0, 16), (0, 1), (1, 8))
    QUICK = os.environ.get("VARIANT_QUICK") == "1"      # integer carrier, shipped options only
    if QUICK:
        CASES = ((0, 8), (0, 16), (0, 1))
    if len(sys.argv) > 3 and sys.argv[3] == "c5":      # config 5 shape: 20 MS/s, 16-bit, 11 channels (low-chip-rate path)
        N, C, CASES = 2000000, 11, ((0, 16),)
    for mode, fmt in CASES:
        t = gs.synthetic_table(E, N, C, fmt, carrier_mode=mode)
        out = torch.zeros(t.n_epochs * t.epoch_bytes, dtype=torch.uint8, device="cuda")
        for pipeline, lean in ((0, 1), (2, 1), (0, 0)) if mode == 0 and fmt == 8 and not QUICK else ((0, 1),):
            with gs.GpuSim.for_table(t) as sim:
                sim.set_option("pipeline", pipeline)
                sim.set_option("lean", lean)
                sim.upload_table(t)
                best = 1e9
                for _ in range(6):
                    sim.generate_device(0, E, out.data_ptr(), out.numel())
                    best = min(best, sim.timing().synth_ms)
            torch.cuda.synchronize()
            h = hashlib.sha256(out.cpu().numpy().tobytes()).hexdigest()[:16]
            print(f"  mode={'FLOAT' if mode else 'INT'} fmt={fmt:2d} pipeline={pipeline} lean={lean} k2={best:7.3f} ms "
                  f"{E * N / best / 1e6:7.1f} GS/s sha={h}", flush=True)
    sys.exit(0)

names = sys.argv[1].split(",")
E = sys.argv[2] if len(sys.argv) > 2 else "2999"
SHAPE = sys.argv[3:4]
for n in names:
    lib = os.path.join(ROOT, "variants", f"libgpusim_{n}.so")
    print(f"== {n}", flush=True)
    env = dict(os.environ, GPUSIM_LIB=lib)
    subprocess.run([sys.executable, os.path.abspath(__file__), "--child", E, *SHAPE], env=env, check=False)
