"""The BASELINE.json scenarios as epoch tables: rows recorded from the reference's own host.

`record(name)` runs the reference host with the libgpusim binding (integration/_build/gps-sdr-sim-gpu-*,
the reference's argv unchanged) in dry-run mode - GPUSIM_DRYRUN=1 GPUSIM_DUMP=<file>: the host does
everything it always does (RINEX, orbits, computeRange/computeCodePhase, generateNavMsg, 30 s channel
refresh, gpssim.c:1738-2188 and :2294-2352) and the shim writes the rows that would cross the C ABI, without
touching a GPU.  That takes a fraction of a second per scenario, so bench.py and the tools record the
real tables on the box they run on instead of timing synthetic rows.

Needs integration/_build/ and oracle/_ref/data/ (built by __graft_entry__.build() where /root/reference
exists; both travel to the GPU box).  `bench_data/*.npz` holds the same tables for a checkout that has
neither (made by tools/make_bench_tables.py).
"""
from __future__ import annotations

import os
import subprocess
import tempfile

from .table import EpochTable

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST_DIR = os.path.join(ROOT, "integration", "_build")
DATA_DIR = os.path.join(ROOT, "oracle", "_ref", "data")
NPZ_DIR = os.path.join(ROOT, "bench_data")

_STATIC = ["-l", "30.286502,120.032669,100"]

# name -> (carrier, argv after "-e brdc3540.14n", label).  File arguments are relative to oracle/_ref/data.
SCENARIOS = {
    "config1": ("int", _STATIC + ["-d", "30", "-s", "2600000", "-b", "16"],
                "config 1: static, 2.6 MS/s, 16-bit, 30 s"),
    "config2": ("int", ["-u", "circle.csv", "-s", "2600000", "-b", "8"],
                "config 2: circle.csv user motion, 2.6 MS/s, 8-bit, 300 s"),
    "config3_satellite": ("int", ["-u", "satellite.csv", "-i", "-s", "2600000", "-b", "16"],
                          "config 3: satellite.csv -i, 2.6 MS/s, 16-bit, 300 s"),
    "config3_rocket": ("int", ["-u", "rocket.csv", "-i", "-s", "2600000", "-b", "16"],
                       "config 3: rocket.csv -i, 2.6 MS/s, 16-bit, 300 s"),
    "config4": ("int", ["-g", "triumphv3.txt", "-s", "1000000", "-b", "1"],
                "config 4: triumphv3.txt NMEA, 1 MS/s, 1-bit, 156 s"),
    "config5_batch": ("int", _STATIC + ["-d", "51.3", "-s", "20000000", "-b", "16"],
                      "config 5: static, 20 MS/s, 16-bit - one 512-epoch batch of the 86 400 s job"),
    "config2_float": ("float", ["-u", "circle.csv", "-s", "2600000", "-b", "8"],
                      "config 2 as shipped (FLOAT_CARR_PHASE): circle.csv, 2.6 MS/s, 8-bit, 300 s"),
    "config1_float": ("float", _STATIC + ["-d", "30", "-s", "2600000", "-b", "16"],
                      "config 1 as shipped (FLOAT_CARR_PHASE): static, 2.6 MS/s, 16-bit, 30 s"),
}
_FILE_ARGS = ("circle.csv", "satellite.csv", "rocket.csv", "triumphv3.txt")


def argv(name: str, out: str = "/dev/null") -> list[str]:
    """The reference's argv for a scenario (without the program name)."""
    _, args, _ = SCENARIOS[name]
    full = ["-e", os.path.join(DATA_DIR, "brdc3540.14n")]
    full += [os.path.join(DATA_DIR, a) if a in _FILE_ARGS else a for a in args]
    return full + ["-o", out]


def host_binary(name: str) -> str:
    return os.path.join(HOST_DIR, f"gps-sdr-sim-gpu-{SCENARIOS[name][0]}")


def can_record(name: str) -> bool:
    return os.path.exists(host_binary(name)) and os.path.exists(os.path.join(DATA_DIR, "brdc3540.14n"))


def record(name: str) -> EpochTable:
    """Rows of the scenario, recorded from the reference host in dry-run mode (no GPU involved)."""
    if not can_record(name):
        raise FileNotFoundError(f"{host_binary(name)} or {DATA_DIR} missing: run __graft_entry__.build() where "
                                f"the reference sources exist")
    with tempfile.TemporaryDirectory(prefix="gpusim_rows_") as tmp:
        dump = os.path.join(tmp, "rows.bin")
        env = dict(os.environ, GPUSIM_DRYRUN="1", GPUSIM_DUMP=dump)
        subprocess.run([host_binary(name), *argv(name)], check=True, env=env,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        return EpochTable.load_dump(dump)


def load(name: str) -> tuple[EpochTable, str]:
    """(table, provenance): recorded on this machine if possible, else the committed copy."""
    if can_record(name):
        return record(name), "rows recorded on this machine from the reference host (dry run of " \
                             f"integration/_build/gps-sdr-sim-gpu-{SCENARIOS[name][0]} {' '.join(SCENARIOS[name][1])})"
    p = os.path.join(NPZ_DIR, f"{name}.npz")
    if os.path.exists(p):
        return EpochTable.load_npz(p), f"rows of the reference host, committed copy bench_data/{name}.npz"
    raise FileNotFoundError(f"no rows for scenario {name}: neither the bound host nor bench_data/{name}.npz")
