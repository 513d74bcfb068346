#!/usr/bin/env python3
"""bench.py - throughput of the GPS L1 C/A sample-synthesis hot path (gpssim.c:2190-2288).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): the circle.csv scenario's shape - 2999 epochs of 0.1 s,
13 visible satellites, 2.6 MS/s, 8-bit IQ - as a seeded synthetic epoch table
(gps_sdr_sim_b200.synthetic_table; there is no way to run the reference's host on the GPU
box inside this script).  One step = one pass of the hot path over all 2999 epochs
(779.74 M samples, 1.56 GB of output) on each GPU; with N GPUs the job is N times longer in
simulated time and time-sharded (weak scaling, no collective on the data path).

`value`  whole-job samples/s with the tables resident in HBM, CUDA events on the launch stream,
         max over ranks.
`e2e`    the same metric through gpusim_generate_epochs(): host table in, host bytes out,
         table compaction + H2D + kernels + D2H inside the timed region.
`--impl reference` times the reference's own single-threaded CPU build (oracle/_ref) instead.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

EPOCHS = 2999            # circle.csv: 3000 rows -> 2999 epochs written (gpssim.c:2154)
N_SAMPLES = 260000       # 2.6 MS/s
N_ACTIVE = 13
FMT = 8
METRIC = "IQ samples/sec (device-timed)"
WORKLOAD = "config2-shape: 2999 epochs x 13 channels, 2.6 MS/s, 8-bit IQ (circle.csv scenario shape)"


def rank_env():
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)),
            int(os.environ.get("WORLD_SIZE", 1)))


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the unmodified reference binary on the host cores
# ---------------------------------------------------------------------------------------------
def ref_paths():
    ref = os.path.join(ROOT, "oracle", "_ref", "gps-sdr-sim-int")
    data = os.path.join(ROOT, "oracle", "_ref", "data")
    ok = os.path.exists(ref) and os.path.exists(os.path.join(data, "circle.csv"))
    return ref, data, ok


def run_reference_once(duration_s: float):
    """One run of oracle/_ref/gps-sdr-sim-int on the circle.csv scenario, 8-bit, 2.6 MS/s.
    -> (samples, wall seconds).  Single-threaded program: 1 core."""
    ref, data, _ = ref_paths()
    epochs = int(duration_s * 10 + 0.5) - 1
    cmd = [ref, "-e", os.path.join(data, "brdc3540.14n"), "-u", os.path.join(data, "circle.csv"),
           "-s", "2600000", "-b", "8", "-d", f"{duration_s:.1f}", "-o", "/dev/null"]
    t0 = time.perf_counter()
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return epochs * N_SAMPLES, time.perf_counter() - t0


def run_port_once(n_epochs: int):
    """Fallback when oracle/_ref did not travel: the oracle port (plain C restatement), 1 thread."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    from gps_sdr_sim_b200 import synthetic_table
    t = synthetic_table(n_epochs, N_SAMPLES, N_ACTIVE, FMT)
    t0 = time.perf_counter()
    oracle_lib.generate(t, nthreads=1)
    return n_epochs * N_SAMPLES, time.perf_counter() - t0


def cpu_baseline(duration_s: float = 45.0):
    _, _, ok = ref_paths()
    if ok:
        samples, sec = run_reference_once(duration_s)
        kind, what = "reference", (f"oracle/_ref/gps-sdr-sim-int (unmodified gpssim.c, gcc -O3, integer carrier) "
                                   f"-u circle.csv -b 8 -s 2600000 -d {duration_s:.0f} -o /dev/null")
    else:
        samples, sec = run_port_once(60)
        kind, what = "port", "oracle/liboracle.so on 60 epochs of the synthetic config2-shape table"
    return {"value": samples / sec, "unit": "samples/s", "cores": 1, "kind": kind, "sample": what,
            "x_realtime": samples / sec / (10.0 * N_SAMPLES), "host_cores_available": os.cpu_count()}


def main_reference(args):
    rank, _, world = rank_env()
    if rank != 0:
        return 0
    _, _, ok = ref_paths()
    dur = 20.0
    runner = (lambda: run_reference_once(dur)) if ok else (lambda: run_port_once(40))
    for _ in range(args.warmup):
        runner()
    samples = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        s, _ = runner()
        samples += s
    sec = time.perf_counter() - t0
    v = samples / sec
    kind = "reference" if ok else "port"
    sample = (f"each step = oracle/_ref/gps-sdr-sim-int -u circle.csv -b 8 -s 2600000 -d {dur:.0f} -o /dev/null "
              f"({int(dur * 10) - 1} epochs of the same scenario shape)") if ok else "oracle port, 40 synthetic epochs per step"
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * sec / max(1, args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32/f64",
            "data": "reference scenario files (circle.csv, brdc3540.14n)" if ok else "synthetic",
            "config": {"workload": WORKLOAD, "bounded_sample": sample},
            "cpu_baseline": {"value": v, "unit": "samples/s", "cores": 1, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "x_realtime": v / (10.0 * N_SAMPLES), "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------
# clocks during the timed region
# ---------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons every few ms from before the warm-up until after the
    timed region; `window(t0, t1)` then reports what was seen DURING the timed region."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.max_mhz, self.error = index, [], None, None
        self._halt = threading.Event()
        self.ready = threading.Event()

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {pynvml.nvmlClocksEventReasonHwSlowdown: "hw_slowdown",
                     pynvml.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     pynvml.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     pynvml.nvmlClocksEventReasonSwPowerCap: "sw_power_cap"}
            self.ready.set()
            while not self._halt.is_set():
                mhz = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                self.samples.append((time.perf_counter(), mhz, tuple(n for b, n in names.items() if r & b)))
                time.sleep(0.004)
        except Exception as exc:  # noqa: BLE001 - clocks are evidence, not a dependency
            self.error = f"{type(exc).__name__}: {exc}"
            self.ready.set()

    def stop(self):
        self._halt.set()
        self.join(timeout=2)

    def window(self, t0: float, t1: float):
        inside = [s for s in self.samples if t0 <= s[0] <= t1]
        mhz = sorted(s[1] for s in inside)
        reasons = sorted({r for s in inside for r in s[2]})
        out = {"sm_mhz": mhz[len(mhz) // 2] if mhz else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
               "samples": len(inside)}
        if self.error:
            out["sampler_error"] = self.error
        return out


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_capture():
    """figures of the committed ncu --set full capture of the synthesis kernel (profiles/traffic.json)"""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(p) as f:
            return json.load(f)
    except Exception:  # noqa: BLE001
        return {}


def ncu_traffic():
    """dram bytes per launch of the synthesis kernel from the committed ncu --set full capture."""
    return ncu_capture().get("k2_synth_sc08_dram_bytes_per_launch")


def bind_to_gpu_cpus(index: int):
    """Pin this rank to the CPUs closest to its GPU (NVML's affinity mask) so that the page-locked
    output buffer is allocated on the NUMA node the GPU's PCIe link hangs off.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} cpus ({cpus[0]}..{cpus[-1]})"
    except Exception as exc:  # noqa: BLE001
        return f"unbound ({type(exc).__name__})"
    return "unbound"


def main_b200(args):
    import torch
    import torch.distributed as dist
    import gps_sdr_sim_b200 as gs

    rank, local_rank, world = rank_env()
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - this path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_cpus(local_rank)    # before any pinned allocation: keep staging memory NUMA-local
    if world > 1:
        # stdout of rank 0 is ONE JSON line: NCCL_DEBUG=VERSION (set on some boxes) makes NCCL print its version there
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # weak scaling: the job is `world` times longer in simulated time; rank r owns a contiguous
    # epoch range.  Every rank builds the same seeded table and keeps its slice.
    from gps_sdr_sim_b200.shard import epoch_range
    table_all = gs.synthetic_table(EPOCHS * world, N_SAMPLES, N_ACTIVE, FMT)
    first, count = epoch_range(rank, world, table_all.n_epochs)
    table = table_all.slice(first, count)
    del table_all
    eb = table.epoch_bytes
    out_bytes = count * eb

    sim = gs.GpuSim(N_SAMPLES, table.delt, FMT, gs.CARRIER_INT, max_batch_epochs=count, device=local_rank)
    if args.no_pipeline:
        sim.set_option("pipeline", 0)
    sim.upload_table(table)                      # tables resident in HBM before the timed region
    out = torch.empty(out_bytes, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.Stream()                 # a real stream: calls are asynchronous and pipeline

    def step():
        sim.generate_device(0, count, out.data_ptr(), out.numel(), stream=stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.ready.wait(timeout=20)
    for _ in range(args.warmup):
        step()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k1_ms = []
    w0 = time.perf_counter()
    ev0.record(stream)
    marks = [ev0]                                # one event between consecutive steps, on the launch stream
    for i in range(args.steps):
        step()
        if i + 1 < args.steps:
            marks.append(torch.cuda.Event(enable_timing=True))
            marks[-1].record(stream)
    ev1.record(stream)
    marks.append(ev1)
    barrier()
    w1 = time.perf_counter()
    clocks = sampler.window(w0, w1)
    ms = ev0.elapsed_time(ev1)
    # per-kernel durations of the last step (CUDA events recorded by the library on the same stream)
    t = sim.timing()
    k1_ms.append(t.chain_ms)
    k2_last_ms = t.synth_ms
    # Average duration of the synthesis kernel's launches over the timed region.  With the chain kernel
    # of the next step on the library's own stream, the launch stream carries nothing but the K2
    # launches (they follow each other within 3 us, tools/step_gap.py), so the time between two marks
    # is one K2 launch - including what the co-running chain kernel costs it, which the library's own
    # event pair of the LAST step (no chain kernel beside it any more) does not show.  Without the
    # overlap the chain kernel runs on the launch stream as well and is subtracted.
    per_step = [marks[i].elapsed_time(marks[i + 1]) for i in range(len(marks) - 1)]
    k2_ms = [m - (t.chain_ms if args.no_pipeline else 0.0) for m in per_step]
    launches_per_step = t.launches
    fast_path = t.fast_path

    # ---- e2e: host table in, host bytes out, through the public C-ABI call ----------------------
    host_out = torch.empty(out_bytes, dtype=torch.uint8, pin_memory=True)
    def e2e_step():
        sim.generate_epochs(table, out_ptr=host_out.data_ptr(), out_capacity=host_out.numel())
    for _ in range(min(2, args.warmup)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_sec = time.perf_counter() - t0
    checksum = int(host_out[:: max(1, out_bytes // 65536)].to(torch.int64).sum())   # result is really on the host
    sampler.stop()

    if world > 1:
        tm = torch.tensor([ms, e2e_sec * 1000.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(tm[0]), float(tm[1])
    else:
        e2e_ms = e2e_sec * 1000.0

    if rank == 0:
        total_samples = float(EPOCHS) * world * N_SAMPLES * args.steps
        value = total_samples / (ms / 1000.0)
        e2e_value = total_samples / (e2e_ms / 1000.0)
        peak, peak_src = measured_peak_hbm()
        k2 = sum(k2_ms) / len(k2_ms)
        achieved = (out_bytes / 1e9) / (k2 / 1000.0)
        h2d = count * 16 * 40 + count       # DevRow 32 B + x0 8 B per slot, + active-channel count
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32/f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "epochs_per_gpu": count, "samples_per_epoch": N_SAMPLES,
                       "channels": N_ACTIVE, "iq_bits": FMT, "carrier": "integer (gpssim.h:4 disabled)",
                       "cache": "each step writes 1.56 GB per GPU, 12x the 126 MB L2; inputs are tables of 2 MB",
                       "sharding": "time (epoch ranges), no collective"},
            "x_realtime": value / (10.0 * N_SAMPLES),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": out_bytes, "ms_per_step": e2e_ms / args.steps,
                    "x_realtime": e2e_value / (10.0 * N_SAMPLES), "api": "gpusim_generate_epochs (C ABI), pinned host output",
                    "rank0_cpu_binding": numa,
                    "host_checksum": checksum},
            "gpu_launches": launches_per_step * args.steps,
            "kernels": {"k1_chain_ms": sum(k1_ms) / len(k1_ms), "k2_synth_ms": k2, "k2_synth_last_step_ms": k2_last_ms,
                        "tuned_kernel": bool(fast_path),
                        "chain_overlaps_previous_synth": not args.no_pipeline},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(), "peak_source": peak_src,
                         "kernel": "k2_synth<8,32>", "algorithmic_bytes_per_launch": out_bytes,
                         "issue": {"active_frac_ncu": (ncu_capture().get("issue_active_pct") or 0.0) / 100.0 or None,
                                   "thread_instructions_per_sample_channel_ncu":
                                       ncu_capture().get("thread_instructions_per_sample_channel"),
                                   "source": "profiles/r01_k2_synth_sc08_ncu.md (not measured live)"},
                         "note": "2 B/sample (SC08) x samples per launch / average CUDA-event duration of the K2 launches of the timed region; "
                                 "the kernel is instruction-issue bound (INT/FP64/LDS per sample and channel), see DESIGN.md"},
        }
        if world == 1:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line), flush=True)

    sim.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-pipeline", action="store_true", help="A/B: chain kernel on the caller's stream, no overlap between steps")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3          # timing rule: at least three untimed passes
    return main_reference(args) if args.impl == "reference" else main_b200(args)


if __name__ == "__main__":
    sys.exit(main())
