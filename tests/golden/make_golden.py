#!/usr/bin/env python3
"""Regenerate tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py

For every scenario below it
  1. runs oracle/_ref/gps-sdr-sim-{int,float} (the reference's gpssim.c, built by
     oracle/build_ref.sh) with the given argv and hashes every epoch of its output file;
  2. runs the reference host with the libgpusim binding in dry-run mode
     (integration/_build/gps-sdr-sim-gpu-*, GPUSIM_DRYRUN=1 GPUSIM_DUMP=...) with the SAME argv
     to record the per-epoch rows that cross the C ABI;
  3. stores rows + per-epoch SHA-256 digests (+ the raw bytes of one short window) in an .npz.

The fixtures pin the oracle (tests/test_golden.py, CPU) and the CUDA path (-m gpu) to bytes
produced by the reference itself, on machines where /root/reference does not exist.
Only epochs listed in `keep` are stored, so the files stay small.
"""
import hashlib
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gps_sdr_sim_b200.table import EpochTable, epoch_bytes  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
DATA = os.path.join(REF, "data")
HOST = os.path.join(ROOT, "integration", "_build")
BRDC = os.path.join(DATA, "brdc3540.14n")
STATIC = ["-l", "30.286502,120.032669,100"]

# name: (carrier mode, argv after "-e brdc", epochs kept (None = all))
SCENARIOS = {
    # BASELINE config 1 (static, 2.6 MS/s), first 2 s, in all three output formats
    "static_int_b16": ("int", STATIC + ["-d", "2", "-s", "2600000", "-b", "16"], None),
    "static_int_b8": ("int", STATIC + ["-d", "2", "-s", "2600000", "-b", "8"], None),
    "static_int_b1": ("int", STATIC + ["-d", "2", "-s", "2600000", "-b", "1"], None),
    # the reference exactly as shipped (FLOAT_CARR_PHASE): pins the oracle's double-carrier branch
    "static_float_b16": ("float", STATIC + ["-d", "1", "-s", "2600000", "-b", "16"], None),
    "circle_float_b8": ("float", ["-u", os.path.join(DATA, "circle.csv"), "-s", "2600000", "-b", "8", "-d", "1.0"], None),
    "nmea_float_1msps_b1": ("float", ["-g", os.path.join(DATA, "triumphv3.txt"), "-s", "1000000", "-b", "1", "-d", "1.5"], None),
    # spacecraft: +-40 kHz Doppler, i.e. both signs of a large double carrier step
    "satellite_float_b16": ("float", ["-u", os.path.join(DATA, "satellite.csv"), "-i", "-s", "2600000", "-b", "16", "-d", "31"],
                            [0, 1, 2, 299, 300, 301]),
    # BASELINE config 4: NMEA trajectory, 1 MS/s (more than one chip per sample), 1-bit
    "nmea_int_1msps_b1": ("int", ["-g", os.path.join(DATA, "triumphv3.txt"), "-s", "1000000", "-b", "1", "-d", "3"], None),
    # BASELINE config 2: circle.csv user motion, 8-bit
    "circle_int_b8": ("int", ["-u", os.path.join(DATA, "circle.csv"), "-s", "2600000", "-b", "8", "-d", "1.5"], None),
    # BASELINE config 3: spacecraft, no iono, +-40 kHz Doppler; 30 s nav-message roll at epoch 300,
    # PRN 15 allocated at 90 s (epoch 900)
    "satellite_int_b16": ("int", ["-u", os.path.join(DATA, "satellite.csv"), "-i", "-s", "2600000", "-b", "16", "-d", "91"],
                          [0, 1, 298, 299, 300, 301, 898, 899, 900, 901, 908]),
    # ephemeris set hop (ieph++, gpssim.c:2311-2314) at 01:00:30 = epoch 900 of this run
    "ephhop_int_b16": ("int", STATIC + ["-t", "2014/12/20,00:59:00", "-d", "91", "-s", "2600000", "-b", "16"],
                       [0, 298, 299, 300, 899, 900, 901, 905]),
    # a sample rate whose epoch length is not a multiple of 32 (generic kernel), 20 MS/s short run
    "odd_rate_int_b16": ("int", STATIC + ["-d", "0.5", "-s", "1234570", "-b", "16"], None),
    "odd_rate_int_b1": ("int", STATIC + ["-d", "0.5", "-s", "1234570", "-b", "1"], None),
    # 2.5 MS/s: epochs of 250 000 samples - a multiple of 8 but not of 32, and a code period (2500
    # samples) that cannot be cut into chunks of a multiple of 8: tuned kernel, plain layout, 8-sample tails
    "rate2500k_int_b16": ("int", STATIC + ["-d", "0.4", "-s", "2500000", "-b", "16"], None),
    "rate2500k_int_b1": ("int", STATIC + ["-d", "0.4", "-s", "2500000", "-b", "1"], None),
    "rate2500k_float_b8": ("float", STATIC + ["-d", "0.4", "-s", "2500000", "-b", "8"], None),
    "static_int_20msps_b16": ("int", STATIC + ["-d", "0.3", "-s", "20000000", "-b", "16"], None),
}


def run(cmd, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run(cmd, env=e, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
    if r.returncode != 0:
        raise SystemExit(f"{' '.join(cmd)} failed:\n{r.stderr[-2000:]}")


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
    only = set(sys.argv[1:])
    with tempfile.TemporaryDirectory() as tmp:
        for name, (mode, argv, keep) in SCENARIOS.items():
            if only and name not in only:
                continue
            ref_bin = os.path.join(tmp, name + ".bin")
            dump = os.path.join(tmp, name + ".tab")
            run([os.path.join(REF, f"gps-sdr-sim-{mode}"), "-e", BRDC, *argv, "-o", ref_bin])
            run([os.path.join(HOST, f"gps-sdr-sim-gpu-{mode}"), "-e", BRDC, *argv, "-o", os.path.join(tmp, "none.bin")],
                env={"GPUSIM_DRYRUN": "1", "GPUSIM_DUMP": dump})
            table = EpochTable.load_dump(dump)
            eb = epoch_bytes(table.samples_per_epoch, table.data_format)
            raw = np.fromfile(ref_bin, dtype=np.uint8)
            assert raw.size == table.n_epochs * eb, (name, raw.size, table.n_epochs, eb)
            epochs = list(range(table.n_epochs)) if keep is None else keep
            digests = [hashlib.sha256(raw[e * eb:(e + 1) * eb].tobytes()).hexdigest() for e in epochs]
            cols = {k: np.ascontiguousarray(v[epochs]) for k, v in table.cols.items()}
            # raw bytes of the first 256 samples' worth of the first kept epoch, for eyeballing / exact diffs
            head = raw[epochs[0] * eb: epochs[0] * eb + min(eb, 1024)].copy()
            path = os.path.join(out_dir, name + ".npz")
            np.savez_compressed(path, samples_per_epoch=table.samples_per_epoch, delt=table.delt,
                                data_format=table.data_format, carrier_mode=table.carrier_mode,
                                epochs=np.array(epochs, dtype=np.int32), sha256=np.array(digests),
                                head=head, argv=np.array(" ".join(a if not a.startswith(DATA) else os.path.basename(a) for a in argv)),
                                **cols)
            print(f"{name}: {len(epochs)} of {table.n_epochs} epochs, N={table.samples_per_epoch}, "
                  f"fmt={table.data_format}, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
