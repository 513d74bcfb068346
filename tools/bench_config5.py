#!/usr/bin/env python3
"""BASELINE config 5 at full size: static receiver, 86 400 s at 20 MS/s, 16-bit (863 999 epochs of
2 000 000 samples = 1.728e12 samples, 6.9 TB of output that cannot be stored), time-sharded across
the GPUs of one box.  Device-timed; the samples are generated into a device buffer that is reused
from batch to batch and checksummed, never copied back.

    python tools/bench_config5.py [--epochs E] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_config5.py

Strong scaling: the job is fixed, rank r generates a contiguous epoch range (no collective on the data
path; torch.distributed only gathers the timings).  Rows are the seeded synthetic table of
gps_sdr_sim_b200.synthetic_table (11 satellites), generated per rank for its own range.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import gps_sdr_sim_b200 as gs
from gps_sdr_sim_b200.shard import batches, epoch_range

ap = argparse.ArgumentParser()
ap.add_argument("--epochs", type=int, default=863999)
ap.add_argument("--batch", type=int, default=1024)
args = ap.parse_args()

rank, local_rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

N, FMT, NCH = 2000000, 16, 11
first, count = epoch_range(rank, world, args.epochs)
t0 = time.perf_counter()
table = gs.synthetic_table(count, N, NCH, FMT, seed=20141220 + rank)
t_table = time.perf_counter() - t0

sim = gs.GpuSim(N, table.delt, FMT, gs.CARRIER_INT, max_batch_epochs=args.batch, device=local_rank)
out = torch.empty(args.batch * table.epoch_bytes, dtype=torch.uint8, device="cuda")
stream = torch.cuda.current_stream()
chk = torch.zeros((), dtype=torch.int64, device="cuda")
k1 = k2 = 0.0
launches = 0
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
w0 = time.perf_counter()
for bf, bn in batches(0, count, args.batch):
    sim.upload_table(table.slice(bf, bn))
    sim.generate_device(0, bn, out.data_ptr(), out.numel(), stream=stream.cuda_stream)
    tm = sim.timing()                      # synchronises on the batch
    k1 += tm.chain_ms
    k2 += tm.synth_ms
    launches += tm.launches
    chk += out[: bn * table.epoch_bytes : 65521].to(torch.int64).sum()
torch.cuda.synchronize()
wall = time.perf_counter() - w0

vals = torch.tensor([k1 + k2, wall * 1000.0, float(count)], dtype=torch.float64, device="cuda")
if world > 1:
    allv = [torch.zeros_like(vals) for _ in range(world)]
    dist.all_gather(allv, vals)
else:
    allv = [vals]
if rank == 0:
    dev_ms = max(float(v[0]) for v in allv)
    wall_ms = max(float(v[1]) for v in allv)
    samples = float(args.epochs) * N
    print(json.dumps({
        "workload": f"config5: static, {args.epochs} epochs x {N} samples (20 MS/s), 16-bit, {NCH} satellites, synthetic rows",
        "n_gpus": world, "epochs_per_gpu": [int(v[2]) for v in allv], "batch_epochs": args.batch,
        "device_ms_max_over_ranks": dev_ms, "wall_ms_incl_table_upload_max_over_ranks": wall_ms,
        "samples_per_s_device_timed": samples / (dev_ms / 1000.0),
        "x_realtime_device_timed": samples / (dev_ms / 1000.0) / 20.0e6,
        "samples_per_s_wall": samples / (wall_ms / 1000.0),
        "x_realtime_wall": samples / (wall_ms / 1000.0) / 20.0e6,
        "rank0": {"k1_chain_ms": k1, "k2_synth_ms": k2, "launches": launches, "table_build_s": t_table,
                  "checksum": int(chk)},
        "output_bytes_total": samples * 4,
    }), flush=True)
sim.close()
if world > 1:
    dist.destroy_process_group()
