"""The C-ABI library: loads, exports every symbol include/gpusim.h declares, and fails loudly
(no CPU fallback) when there is no GPU.  No compute calls here."""
import ctypes
import os
import re

import pytest

import gps_sdr_sim_b200 as gs
from gps_sdr_sim_b200 import api
from conftest import ROOT, has_gpu


def declared_functions():
    hdr = open(os.path.join(ROOT, "include", "gpusim.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = re.findall(r"\b(gpusim_[a-z_0-9]+)\s*\(", hdr)
    return sorted(set(n for n in names if n not in ("gpusim_sink_fn",)))


def test_header_and_binding_agree():
    assert declared_functions() == sorted(api.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = gs.load_library()
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} not exported by libgpusim.so"
    assert lib.gpusim_abi_version() == 2
    assert lib.gpusim_strerror(0) == b"ok"


def test_library_is_the_in_tree_cuda_build():
    path = gs.library_path()
    assert os.path.dirname(path) == os.path.join(ROOT, "gps_sdr_sim_b200")
    # the sm_100a cubin is embedded in the shared object
    blob = open(path, "rb").read()
    assert b"sm_100a" in blob and b"k2_synth" in blob and b"k1_chain" in blob and b"k0_navmsg" in blob


def test_struct_layouts_match_header():
    assert ctypes.sizeof(api._Config) == 32
    assert ctypes.sizeof(api.Timing) == 24
    from gps_sdr_sim_b200.table import CEpochTable
    assert ctypes.sizeof(CEpochTable) == 8 + 13 * 8
    from gps_sdr_sim_b200.table import NAV_FRAME
    assert NAV_FRAME.itemsize == 256 and NAV_FRAME.fields["first"][1] == 200 and NAV_FRAME.fields["tow"][1] == 244


def test_bad_config_rejected_before_touching_cuda():
    with pytest.raises(gs.GpuSimError) as e:
        gs.GpuSim(260000, 1 / 2.6e6, data_format=12)
    assert e.value.status == 1
    with pytest.raises(gs.GpuSimError) as e:
        gs.GpuSim(0, 1 / 2.6e6)
    assert e.value.status == 1


def test_exact_host_carrier_advance_matches_sequential_replay():
    # FLOAT_CARR_PHASE hosts: gpusim_advance_carrier_f64 == N executions of gpssim.c:2245-2250
    import numpy as np
    import oracle_lib
    rng = np.random.default_rng(5)
    for f_carr in (3712.5, -3712.5, 43917.0, -40316.0, 0.0, 7.25):
        x0 = float(rng.uniform(0, 1))
        for fs, n in ((2.6e6, 260000), (1.0e6, 100000)):
            _, want = oracle_lib.carrier_phase_checkpoints(x0, f_carr, 1.0 / fs, n, n)
            got = gs.advance_carrier_f64(x0, f_carr, 1.0 / fs, n)
            assert np.float64(got).view(np.uint64) == np.float64(want).view(np.uint64), (f_carr, fs)


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_no_gpu_means_loud_failure_not_fallback():
    with pytest.raises(gs.GpuSimError) as e:
        gs.GpuSim(260000, 1 / 2.6e6)
    assert e.value.status == 2 and "no CPU path" in str(e.value)
