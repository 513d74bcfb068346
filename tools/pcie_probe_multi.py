#!/usr/bin/env python3
"""Device-to-host copy bandwidth of every GPU of the box, one at a time and all at once.

Answers one question: is the end-to-end rate at N GPUs bounded by the platform (PCIe switches,
host memory, the container's CPU set) or by how this library stages its output?  One process per
GPU; each reports its CPU affinity, the NUMA node of its page-locked buffer, its bandwidth alone and
its bandwidth while all other GPUs copy too.

usage: python tools/pcie_probe_multi.py [n_gpus]        (parent)
"""
import ctypes
import os
import subprocess
import sys
import time

GB = 1 << 30
SIZE = 1 * GB
REPS = 8


def numa_node_of(addr):
    """NUMA node of the page holding addr (move_pages with nodes=NULL queries)."""
    libc = ctypes.CDLL(None, use_errno=True)
    pages = (ctypes.c_void_p * 1)(addr & ~4095)
    status = (ctypes.c_int * 1)(-1)
    SYS_move_pages = 279  # x86_64
    r = libc.syscall(SYS_move_pages, 0, ctypes.c_ulong(1), pages, None, status, 0)
    return status[0] if r == 0 else f"err{ctypes.get_errno()}"


def child(idx, n, sync_dir, bind):
    import torch
    aff0 = sorted(os.sched_getaffinity(0))
    note = ""
    if bind:
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
            cpus = [64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1]
            ok = [c for c in cpus if c in aff0]
            note = f"nvml cpus {cpus[0]}..{cpus[-1]} ({len(cpus)}), usable {len(ok)}"
            if ok:
                os.sched_setaffinity(0, ok)
        except Exception as exc:  # noqa: BLE001
            note = f"nvml: {type(exc).__name__}"
    torch.cuda.set_device(idx)
    d = torch.empty(SIZE, dtype=torch.uint8, device="cuda")
    hbuf = torch.empty(SIZE, dtype=torch.uint8, pin_memory=True)
    hbuf.fill_(1)
    node = numa_node_of(hbuf.data_ptr())

    def run():
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(REPS):
            hbuf.copy_(d, non_blocking=True)
        torch.cuda.synchronize()
        return SIZE * REPS / (time.perf_counter() - t0) / 1e9

    def wait(tag, count):
        open(os.path.join(sync_dir, f"{tag}.{idx}"), "w").close()
        while sum(1 for f in os.listdir(sync_dir) if f.startswith(tag + ".")) < count:
            time.sleep(0.002)

    run()
    wait("ready", n)
    alone = None
    for turn in range(n):           # one GPU at a time
        if turn == idx:
            alone = run()
        wait(f"turn{turn}", n)
    together = run()                # everybody
    wait("done", n)
    print(f"gpu {idx}: affinity {aff0[0]}..{aff0[-1]} ({len(aff0)}) {note}; pinned buffer on node {node}; "
          f"alone {alone:.1f} GB/s, all together {together:.1f} GB/s", flush=True)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5] == "1")
        return
    import tempfile
    import torch
    n = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
    print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout)
    for p in ("/sys/devices/system/node/online", "/sys/fs/cgroup/cpuset.cpus.effective", "/sys/fs/cgroup/cpuset.mems.effective"):
        try:
            print(p, open(p).read().strip())
        except OSError as exc:
            print(p, exc)
    for bind in ("0", "1"):
        print(f"--- bind to NVML cpu affinity: {bind}", flush=True)
        with tempfile.TemporaryDirectory() as sync_dir:
            procs = [subprocess.Popen([sys.executable, __file__, "--child", str(i), str(n), sync_dir, bind]) for i in range(n)]
            for p in procs:
                p.wait(timeout=300)


if __name__ == "__main__":
    main()
