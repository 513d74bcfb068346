"""The device algorithms (gpusim_core.h, as compiled for the CPU by tests/emu) under AddressSanitizer and
UndefinedBehaviorSanitizer: out-of-bounds reads of the chip-window / carrier tables, the checkpoint arrays and
the row arrays, shifts by out-of-range counts, signed overflow.  compute-sanitizer is not available on the GPU
pool; this is the same arithmetic and the same indexing, thread by thread, on the CPU."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

CHILD = r'''
import hashlib, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
import numpy as np
import emu_lib
from conftest import load_golden
import gps_sdr_sim_b200 as gs
def digests(buf, t):
    eb = t.epoch_bytes
    return [hashlib.sha256(buf[e * eb:(e + 1) * eb].tobytes()).hexdigest() for e in range(t.n_epochs)]
n = 0
for name, kw in (("static_int_b16", {}), ("static_int_b1", {"force_wrap": True}), ("static_int_b8", {"accum": 3}),
                 ("nmea_int_1msps_b1", {"kernel": emu_lib.TUNED16}), ("satellite_int_b16", {"chunk": 128}),
                 ("static_float_b16", {}), ("satellite_float_b16", {"force_wrap": True}),
                 ("odd_rate_int_b16", {"kernel": emu_lib.GENERIC}), ("static_int_b16", {"accum": 0})):
    t, want, _ = load_golden(name)
    t, want = t.slice(0, 2), want[:2]
    out = emu_lib.generate(t, **{"chunk": 512, "kernel": emu_lib.TUNED32, **kw})
    assert digests(out, t) == want, name
    n += 1
# ragged synthetic tables: empty slots, 16 channels, every format
for fmt in (1, 8, 16):
    t = gs.synthetic_table(2, 104000, 16, fmt, seed=5)
    t.cols["prn"][1, ::3] = 0
    emu_lib.generate(t, chunk=256)
    n += 1
# chains: tiny and huge steps, both signs
for d in (0.3935, -0.73, 1e-4, -1e-4, 1.0230003):
    emu_lib.phase_chain(100.0, d, 512.0 if abs(d) < 1 else 1023.0, 20000, 512)
print("sanitized cases ok:", n)
'''


def test_device_algorithms_are_clean_under_asan_and_ubsan():
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("no AddressSanitizer runtime on this machine")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import emu_lib
    emu_lib.build(sanitized=True)
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:halt_on_error=1:abort_on_error=0",
               UBSAN_OPTIONS="halt_on_error=1:print_stacktrace=1", GPUSIM_EMU_SANITIZED="1")
    r = subprocess.run([sys.executable, "-c", CHILD, ROOT], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0 and "sanitized cases ok" in r.stdout, (r.stdout[-500:], r.stderr[-3000:])
