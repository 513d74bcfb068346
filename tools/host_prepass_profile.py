#!/usr/bin/env python3
"""Where the reference host's per-epoch row pre-pass spends its time (is a device-side builder of the
navigation message / channel refresh worth having?).  Builds the bound host with -pg in a scratch directory, runs it
in dry-run mode (GPUSIM_DRYRUN=1: every epoch's row is produced and recorded, nothing is generated, no GPU) on
config 5 - static receiver, 20 MS/s, 86 400 s = 863 999 epochs - single-threaded, and prints gprof's flat profile.
usage: python tools/host_prepass_profile.py [seconds]        (needs /root/reference; CPU only)"""
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "integration"))
import build_host as bh  # noqa: E402

secs = sys.argv[1] if len(sys.argv) > 1 else "86400"
tmp = tempfile.mkdtemp(prefix="gpusim_prof_")
try:
    bh.CFLAGS = bh.CFLAGS + ["-pg", "-fno-inline-functions"]
    bh.OUT = tmp
    exe = bh.build("int", tmp)
    data = os.path.join(ROOT, "oracle", "_ref", "data")
    env = dict(os.environ, GPUSIM_DRYRUN="1", GPUSIM_DUMP=os.path.join(tmp, "rows.tab"), GPUSIM_HOST_THREADS="1",
               LD_LIBRARY_PATH=os.path.join(ROOT, "gps_sdr_sim_b200"))
    t0 = time.perf_counter()
    subprocess.run([exe, "-e", os.path.join(data, "brdc3540.14n"), "-l", "30.286502,120.032669,100", "-d", secs,
                    "-s", "20000000", "-b", "16", "-o", "/dev/null"], cwd=tmp, env=env, check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    wall = time.perf_counter() - t0
    epochs = int(float(secs) * 10) - 1
    print(f"dry run of config 5 (-d {secs} -s 20000000 -b 16, static): {epochs} epochs of rows in {wall:.2f} s "
          f"= {wall / epochs * 1e6:.2f} us per epoch, one host thread, table dump {os.path.getsize(env['GPUSIM_DUMP']) / 1e6:.0f} MB")
    prof = subprocess.run(["gprof", "-b", "-p", exe, os.path.join(tmp, "gmon.out")], capture_output=True, text=True).stdout
    print("\n".join(prof.splitlines()[:28]))
finally:
    shutil.rmtree(tmp, ignore_errors=True)
