"""Run the UNMODIFIED reference (oracle/_ref, built from /root/reference by oracle/build_ref.sh)
and compare its output file with the oracle fed by rows recorded from the reference host.
Skipped where oracle/_ref or integration/_build do not exist."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib
from conftest import ROOT
from gps_sdr_sim_b200.table import EpochTable

HOST = os.path.join(ROOT, "integration", "_build")


def _need(mode):
    ref = oracle_lib.ref_binary(mode)
    host = os.path.join(HOST, f"gps-sdr-sim-gpu-{mode}")
    if ref is None or not os.path.exists(host):
        pytest.skip("oracle/_ref or integration/_build not built (needs /root/reference)")
    return ref, host


CASES = [
    ("int", ["-l", "30.286502,120.032669,100", "-d", "1.2", "-s", "2600000", "-b", "16"]),
    ("float", ["-l", "30.286502,120.032669,100", "-d", "0.8", "-s", "2600000", "-b", "8"]),
    ("int", ["-u", "circle.csv", "-d", "1.0", "-s", "2600000", "-b", "1"]),
    ("int", ["-g", "triumphv3.txt", "-d", "1.0", "-s", "1000000", "-b", "8"]),
    ("int", ["-u", "rocket.csv", "-i", "-d", "0.7", "-s", "4000000", "-b", "16"]),
]


@pytest.mark.parametrize("mode,argv", CASES)
def test_oracle_equals_reference_run(mode, argv, tmp_path):
    ref, host = _need(mode)
    argv = [oracle_lib.ref_data(a) if a.endswith((".csv", ".txt")) else a for a in argv]
    common = ["-e", oracle_lib.ref_data("brdc3540.14n"), *argv]
    out = tmp_path / "ref.bin"
    dump = tmp_path / "rows.tab"
    subprocess.run([ref, *common, "-o", str(out)], check=True, capture_output=True)
    env = dict(os.environ, GPUSIM_DRYRUN="1", GPUSIM_DUMP=str(dump))
    subprocess.run([host, *common, "-o", str(tmp_path / "none.bin")], check=True, capture_output=True, env=env)
    table = EpochTable.load_dump(str(dump))
    want = np.fromfile(out, dtype=np.uint8)
    got = oracle_lib.generate(table)
    assert got.size == want.size and np.array_equal(got, want)


PREPASS_CASES = [
    # 30 s channel refresh at epoch 300, ephemeris hop at epoch 900, a satellite set that changes
    ("int", ["-l", "30.286502,120.032669,100", "-t", "2014/12/20,00:59:00", "-d", "95", "-s", "2600000", "-b", "16"]),
    ("int", ["-u", "satellite.csv", "-i", "-d", "95", "-s", "2600000", "-b", "16"]),
    ("float", ["-u", "circle.csv", "-d", "62", "-s", "2600000", "-b", "8"]),
    ("float", ["-g", "triumphv3.txt", "-s", "1000000", "-b", "1"]),
]


@pytest.mark.parametrize("mode,argv", PREPASS_CASES)
def test_parallel_host_prepass_records_the_serial_rows(mode, argv, tmp_path):
    """with GPUSIM_HOST_THREADS > 1 the hook computes computeRange() ahead of time in
    parallel over epochs and (FLOAT hosts) walks the carrier chains per batch, one channel slot per thread.
    The rows must be the ones the reference's serial order of calls produces, bit for bit, and every
    range must come from the look-ahead window (windows end at the 30 s refreshes)."""
    _, host = _need(mode)
    argv = [oracle_lib.ref_data(a) if a.endswith((".csv", ".txt")) else a for a in argv]
    common = [host, "-e", oracle_lib.ref_data("brdc3540.14n"), *argv, "-o", str(tmp_path / "none.bin")]
    dumps = {}
    logs = {}
    for threads, batch in (("1", "256"), ("6", "256"), ("3", "7")):
        dump = tmp_path / f"rows_{threads}.tab"
        env = dict(os.environ, GPUSIM_DRYRUN="1", GPUSIM_DUMP=str(dump), GPUSIM_HOST_THREADS=threads,
                   GPUSIM_BATCH_EPOCHS=batch, GPUSIM_VERBOSE="1")
        r = subprocess.run(common, check=True, capture_output=True, text=True, env=env)
        dumps[threads] = open(dump, "rb").read()
        logs[threads] = [ln for ln in r.stderr.splitlines() if ln.startswith("gpusim hook:") and "look-ahead" in ln][-1]
    assert dumps["1"] == dumps["6"] == dumps["3"] and len(dumps["1"]) > 1000
    assert "0 windows" in logs["1"]
    assert " 0 direct calls" in logs["6"] and " 0 windows" not in logs["6"]
