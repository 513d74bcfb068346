"""End to end on the GPU box: the reference's own host and CLI with the libgpusim binding
(integration/_build/gps-sdr-sim-gpu-int) against the unmodified reference CPU build
(oracle/_ref/gps-sdr-sim-int), identical argv, output files compared byte for byte (cmp)."""
import concurrent.futures
import filecmp
import hashlib
import os
import subprocess

import pytest

import oracle_lib
from conftest import ROOT

pytestmark = pytest.mark.gpu

HOST = os.path.join(ROOT, "integration", "_build", "gps-sdr-sim-gpu-int")
HOST_FLOAT = os.path.join(ROOT, "integration", "_build", "gps-sdr-sim-gpu-float")
D = oracle_lib.ref_data


def shipped(mode: str = "int"):
    """The unmodified reference build and the bound host for `mode`.  Both are prebuilt binaries that travel
    with the tree (git-ignored, built by __graft_entry__.build() where /root/reference exists).  On a GPU box
    their absence is a FAILURE, not a skip: these are the strongest parity tests of the suite."""
    ref = oracle_lib.ref_binary(mode)
    host = HOST if mode == "int" else HOST_FLOAT
    missing = [p for p in (ref or os.path.join(ROOT, "oracle", "_ref", f"gps-sdr-sim-{mode}"), host,
                           D("brdc3540.14n")) if not os.path.exists(p)]
    if missing:
        pytest.fail("not shipped to this GPU box: " + ", ".join(os.path.relpath(m, ROOT) for m in missing) +
                    " - run __graft_entry__.build() in a container that has /root/reference before gpurun")
    return ref, host

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
    ref, _ = shipped("int")
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
    ref, _ = shipped("float")
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
    ref, _ = shipped("int")
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
    ref, _ = shipped("int")
    common = ["-e", D("brdc3540.14n"), "-u", D("circle.csv"), "-s", "2600000", "-b", "8", "-d", "12"]
    a = tmp_path / "ref.bin"
    r = subprocess.run([ref, *common, "-o", str(a)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-500:]
    g = subprocess.run([HOST, *common, "-o", "-"], capture_output=True, env=dict(os.environ, GPUSIM_BATCH_EPOCHS="40"))
    assert g.returncode == 0, g.stderr[-800:]
    assert g.stdout == a.read_bytes()


# ---- full-length runs of the BASELINE configurations -------------------------------------------------------
# Both programs write to stdout (`-o -`, gpssim.c:2103-2111) and the bytes are hashed as they arrive, so the
# 1.5 - 3 GB outputs never touch a disk.  The reference needs ~20 s of one host core per 300 s scenario; all
# reference runs are started together on the box's idle cores the first time one of them is needed.
FULL = {
    # name: (carrier mode, argv, what it crosses)
    "config2_circle_300s": ("int", ["-u", "circle.csv", "-s", "2600000", "-b", "8"],
                            "3000 rows = USER_MOTION_SIZE, nine 30 s nav-message rolls"),
    "config3_satellite_300s": ("int", ["-u", "satellite.csv", "-i", "-s", "2600000", "-b", "16"],
                               "3001 rows truncated to 3000 (gpssim.h:19-21); PRN added @90 s, dropped @150 s, slot reused @180 s"),
    "config3_rocket_300s": ("int", ["-u", "rocket.csv", "-i", "-s", "2600000", "-b", "16"],
                            "3901 rows truncated to 3000"),
    "config5_20msps_31s": ("int", ["-l", "30.286502,120.032669,100", "-d", "31", "-s", "20000000", "-b", "16"],
                           "20 MS/s across the 30 s refresh (gpssim.c:2294-2345)"),
    "config1_asshipped_30s": ("float", ["-l", "30.286502,120.032669,100", "-d", "30", "-s", "2600000", "-b", "16"],
                              "the reference exactly as `make` builds it (FLOAT_CARR_PHASE)"),
    "config4_asshipped_156s": ("float", ["-g", "triumphv3.txt", "-s", "1000000", "-b", "1"],
                               "as shipped, NMEA trajectory, 1-bit, PRN added @90 s"),
}
_pool = concurrent.futures.ThreadPoolExecutor(max_workers=8)
_ref_runs = {}


def _stream_digest(cmd, env=None):
    """run cmd, hash its stdout in 8 MiB pieces -> (returncode, sha256 hex, bytes, stderr tail)"""
    import tempfile
    with tempfile.TemporaryFile() as errf:      # the reference prints a progress line per epoch: never a pipe nobody drains
        p = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=errf, env=env)
        h, n = hashlib.sha256(), 0
        while True:
            buf = p.stdout.read(8 << 20)
            if not buf:
                break
            h.update(buf)
            n += len(buf)
        rc = p.wait()
        errf.seek(max(0, errf.tell() - 800))
        err = errf.read().decode(errors="replace")
    return rc, h.hexdigest(), n, err


def _full_argv(name):
    argv = [D(a) if a.endswith((".csv", ".txt")) else a for a in FULL[name][1]]
    return ["-e", D("brdc3540.14n"), *argv, "-o", "-"]


def _reference_digest(name):
    if not _ref_runs:          # start every reference run now; they are single-threaded
        for n, (mode, _, _) in FULL.items():
            ref, _ = shipped(mode)
            _ref_runs[n] = _pool.submit(_stream_digest, [ref, *_full_argv(n)])
    return _ref_runs[name].result()


@pytest.mark.parametrize("name", sorted(FULL))
def test_full_length_run_equals_reference(name, gpu_required):
    mode = FULL[name][0]
    _, host = shipped(mode)
    rc_r, sha_r, n_r, err_r = _reference_digest(name)
    assert rc_r == 0, err_r
    rc_g, sha_g, n_g, err_g = _stream_digest([host, *_full_argv(name)])
    assert rc_g == 0, err_g
    assert n_g == n_r > 0
    assert sha_g == sha_r, f"{name}: output differs from the reference ({FULL[name][2]})"


@pytest.mark.parametrize("level", ["1", "2"])
@pytest.mark.parametrize("name", ["config3_satellite_300s", "config2_circle_300s", "config4_asshipped_156s"])
def test_full_length_run_with_device_built_nav_words_equals_reference(name, level, gpu_required):
    """GPUSIM_NAV_DEVICE=1 - the shim hands the library the subframes (chan[i].sbf) of every
    generateNavMsg() call instead of data bits; TOW counts, week number and parity (gpssim.c:1467-1547, :693-756) are
    computed on the device and every row takes its bits from there.  GPUSIM_NAV_DEVICE=2: the subframes themselves are
    made on the device too, from the broadcast ephemerides (eph2sbf, gpssim.c:490-665).  GPUSIM_NAV_CHECK=1 also compares
    the device's words with the host's chan[i].dwrd after every build.  Same bytes as the unmodified reference."""
    mode = FULL[name][0]
    _, host = shipped(mode)
    rc_r, sha_r, n_r, err_r = _reference_digest(name)
    assert rc_r == 0, err_r
    env = dict(os.environ, GPUSIM_NAV_DEVICE=level, GPUSIM_NAV_CHECK="1")
    rc_g, sha_g, n_g, err_g = _stream_digest([host, *_full_argv(name)], env=env)
    assert rc_g == 0, err_g
    assert n_g == n_r > 0
    assert sha_g == sha_r, f"{name}: output with device-built navigation words differs from the reference"
