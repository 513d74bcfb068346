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
    assert lib.gpusim_abi_version() == 1
    assert lib.gpusim_strerror(0) == b"ok"


def test_library_is_the_in_tree_cuda_build():
    path = gs.library_path()
    assert os.path.dirname(path) == os.path.join(ROOT, "gps_sdr_sim_b200")
    # the sm_100a cubin is embedded in the shared object
    blob = open(path, "rb").read()
    assert b"sm_100a" in blob and b"k2_synth" in blob and b"k1_chain" in blob


def test_struct_layouts_match_header():
    assert ctypes.sizeof(api._Config) == 32
    assert ctypes.sizeof(api.Timing) == 20
    from gps_sdr_sim_b200.table import CEpochTable
    assert ctypes.sizeof(CEpochTable) == 8 + 10 * 8


def test_bad_config_rejected_before_touching_cuda():
    with pytest.raises(gs.GpuSimError) as e:
        gs.GpuSim(260000, 1 / 2.6e6, data_format=12)
    assert e.value.status == 1
    with pytest.raises(gs.GpuSimError) as e:
        gs.GpuSim(0, 1 / 2.6e6)
    assert e.value.status == 1


def test_float_carrier_hosts_are_refused_not_approximated():
    with pytest.raises(gs.GpuSimError) as e:
        gs.GpuSim(260000, 1 / 2.6e6, carrier_mode=gs.CARRIER_FLOAT)
    assert e.value.status == 5 and "FLOAT_CARR_PHASE" in str(e.value)


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_no_gpu_means_loud_failure_not_fallback():
    with pytest.raises(gs.GpuSimError) as e:
        gs.GpuSim(260000, 1 / 2.6e6)
    assert e.value.status == 2 and "no CPU path" in str(e.value)
