"""End to end on the GPU box: the reference's own host and CLI with the libgpusim binding
(integration/_build/gps-sdr-sim-gpu-int) against the unmodified reference CPU build
(oracle/_ref/gps-sdr-sim-int), identical argv, output files compared byte for byte (cmp)."""
import filecmp
import os
import subprocess

import pytest

import oracle_lib
from conftest import ROOT

pytestmark = pytest.mark.gpu

HOST = os.path.join(ROOT, "integration", "_build", "gps-sdr-sim-gpu-int")
HOST_FLOAT = os.path.join(ROOT, "integration", "_build", "gps-sdr-sim-gpu-float")
D = oracle_lib.ref_data

CASES = {
    "config1_static_b16": ["-l", "30.286502,120.032669,100", "-d", "30", "-s", "2600000", "-b", "16"],
    "config2_circle_b8": ["-u", "circle.csv", "-s", "2600000", "-b", "8", "-d", "40"],
    "config3_satellite_b16": ["-u", "satellite.csv", "-i", "-s", "2600000", "-b", "16", "-d", "35"],
    "config3_rocket_b16": ["-u", "rocket.csv", "-i", "-s", "2600000", "-b", "16", "-d", "35"],
    "config4_nmea_b1": ["-g", "triumphv3.txt", "-s", "1000000", "-b", "1"],
    "config5_prefix_20msps": ["-l", "30.286502,120.032669,100", "-d", "2", "-s", "20000000", "-b", "16"],
    "odd_rate_generic": ["-l", "30.286502,120.032669,100", "-d", "3", "-s", "1234570", "-b", "8"],
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_host_with_gpu_binding_equals_reference_cli(name, gpu_required, tmp_path):
    ref = oracle_lib.ref_binary("int")
    if ref is None or not os.path.exists(HOST):
        pytest.skip("oracle/_ref / integration/_build were not shipped to this box")
    argv = [D(a) if a.endswith((".csv", ".txt")) else a for a in CASES[name]]
    common = ["-e", D("brdc3540.14n"), *argv]
    a, b = tmp_path / "ref.bin", tmp_path / "gpu.bin"
    r = subprocess.run([ref, *common, "-o", str(a)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-500:]
    env = dict(os.environ, GPUSIM_BATCH_EPOCHS="128")
    g = subprocess.run([HOST, *common, "-o", str(b)], capture_output=True, text=True, env=env)
    assert g.returncode == 0, g.stderr[-800:]
    assert os.path.getsize(a) == os.path.getsize(b) > 0
    assert filecmp.cmp(a, b, shallow=False)
    # the CLI's own messages are untouched (gpssim.c:2037-2039, :2357)
    assert "Start time" in g.stderr and "Done!" in g.stderr


FLOAT_CASES = {
    "asshipped_static_b16": ["-l", "30.286502,120.032669,100", "-d", "10", "-s", "2600000", "-b", "16"],
    "asshipped_circle_b8": ["-u", "circle.csv", "-s", "2600000", "-b", "8", "-d", "35"],
    "asshipped_rocket_b1": ["-u", "rocket.csv", "-i", "-s", "2600000", "-b", "1", "-d", "8"],
}


@pytest.mark.parametrize("name", sorted(FLOAT_CASES))
def test_asshipped_float_host_with_gpu_binding_equals_reference_cli(name, gpu_required, tmp_path):
    """The reference exactly as shipped (FLOAT_CARR_PHASE, what `make` builds) against the same host with
    the libgpusim binding: double carrier phase on the device, exact host-side carrier advance."""
    ref = oracle_lib.ref_binary("float")
    if ref is None or not os.path.exists(HOST_FLOAT):
        pytest.skip("oracle/_ref / integration/_build were not shipped to this box")
    argv = [D(a) if a.endswith((".csv", ".txt")) else a for a in FLOAT_CASES[name]]
    common = ["-e", D("brdc3540.14n"), *argv]
    a, b = tmp_path / "ref.bin", tmp_path / "gpu.bin"
    r = subprocess.run([ref, *common, "-o", str(a)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-500:]
    g = subprocess.run([HOST_FLOAT, *common, "-o", str(b)], capture_output=True, text=True,
                       env=dict(os.environ, GPUSIM_BATCH_EPOCHS="100"))
    assert g.returncode == 0, g.stderr[-800:]
    assert os.path.getsize(a) == os.path.getsize(b) > 0
    assert filecmp.cmp(a, b, shallow=False)


def test_multi_gpu_time_sharded_host_writes_the_same_file(gpu_required, tmp_path):
    """GPUSIM_DEVICE_LIST: batches are dealt round-robin to one worker per listed device and written
    in order by one writer.  Uses every GPU of the box (two workers on one GPU if there is only one)."""
    import torch
    ref = oracle_lib.ref_binary("int")
    if ref is None or not os.path.exists(HOST):
        pytest.skip("oracle/_ref / integration/_build were not shipped to this box")
    n = torch.cuda.device_count()
    devices = ",".join(str(d) for d in range(n)) if n > 1 else "0,0"
    common = ["-e", D("brdc3540.14n"), "-u", D("circle.csv"), "-s", "2600000", "-b", "16", "-d", "25"]
    a, b = tmp_path / "ref.bin", tmp_path / "gpu.bin"
    r = subprocess.run([ref, *common, "-o", str(a)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-500:]
    g = subprocess.run([HOST, *common, "-o", str(b)], capture_output=True, text=True,
                       env=dict(os.environ, GPUSIM_BATCH_EPOCHS="24", GPUSIM_DEVICE_LIST=devices))
    assert g.returncode == 0, g.stderr[-800:]
    assert os.path.getsize(a) == os.path.getsize(b) > 0
    assert filecmp.cmp(a, b, shallow=False)


def test_stdout_streaming_keeps_the_format_contract(gpu_required, tmp_path):
    """`-o -` (gpssim.c:2103-2111): the samples go to stdout in epoch order, e.g. into a player's pipe.
    The ordered sink writes exactly the file bytes (output sink)."""
    ref = oracle_lib.ref_binary("int")
    if ref is None or not os.path.exists(HOST):
        pytest.skip("oracle/_ref / integration/_build were not shipped to this box")
    common = ["-e", D("brdc3540.14n"), "-u", D("circle.csv"), "-s", "2600000", "-b", "8", "-d", "12"]
    a = tmp_path / "ref.bin"
    r = subprocess.run([ref, *common, "-o", str(a)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-500:]
    g = subprocess.run([HOST, *common, "-o", "-"], capture_output=True, env=dict(os.environ, GPUSIM_BATCH_EPOCHS="40"))
    assert g.returncode == 0, g.stderr[-800:]
    assert g.stdout == a.read_bytes()
