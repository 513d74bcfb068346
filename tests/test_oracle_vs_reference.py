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
