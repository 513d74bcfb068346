"""ctypes access to the CPU checker (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")

_lib = None


def build_oracle() -> str:
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    src = os.path.join(ORACLE_DIR, "gpssim_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", ORACLE_DIR, "liboracle.so"], check=True, capture_output=True)
    return so


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_oracle())
        _lib.oracle_generate_epochs.restype = ctypes.c_int
        _lib.oracle_generate_epochs.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                ctypes.c_void_p]
        _lib.oracle_code_phase_checkpoints.restype = None
        _lib.oracle_code_phase_checkpoints.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                                       ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                                       ctypes.c_void_p]
        _lib.oracle_ca_code.restype = ctypes.c_int
        _lib.oracle_carrier_phase_checkpoints.restype = ctypes.c_double
        _lib.oracle_carrier_phase_checkpoints.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                                          ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    return _lib


def generate(table, first: int = 0, count: int | None = None, nthreads: int | None = None) -> np.ndarray:
    """Oracle bytes for epochs [first, first+count) of an EpochTable."""
    count = table.n_epochs - first if count is None else count
    nthreads = (os.cpu_count() or 1) if nthreads is None else nthreads
    out = np.empty(count * table.epoch_bytes, dtype=np.uint8)
    c = table.as_c()
    rc = lib().oracle_generate_epochs(ctypes.addressof(c), first, count, table.samples_per_epoch,
                                      table.delt, table.data_format, table.carrier_mode,
                                      nthreads, out.ctypes.data)
    if rc != 0:
        raise RuntimeError("oracle_generate_epochs failed")
    return out


def code_phase_checkpoints(code_phase: float, f_code: float, delt: float, n: int, every: int):
    k = (n + every - 1) // every
    x = np.empty(k, dtype=np.float64)
    w = np.empty(k, dtype=np.int32)
    lib().oracle_code_phase_checkpoints(code_phase, f_code, delt, n, every, x.ctypes.data, w.ctypes.data)
    return x, w


def carrier_phase_checkpoints(carr_phase: float, f_carr: float, delt: float, n: int, every: int):
    k = (n + every - 1) // every
    x = np.empty(k, dtype=np.float64)
    end = lib().oracle_carrier_phase_checkpoints(carr_phase, f_carr, delt, n, every, x.ctypes.data)
    return x, end


def carrier_lut():
    s = (ctypes.c_int * 512)()
    c = (ctypes.c_int * 512)()
    lib().oracle_carrier_lut(s, c)
    return np.array(s[:], dtype=np.int32), np.array(c[:], dtype=np.int32)


def ca_code(prn: int) -> np.ndarray:
    ca = (ctypes.c_int * 1023)()
    if lib().oracle_ca_code(prn, ca) != 0:
        raise ValueError("bad prn")
    return np.array(ca[:], dtype=np.int32)


# ---- the unmodified reference, when oracle/_ref was built (oracle/build_ref.sh) ----------
def ref_binary(mode: str) -> str | None:
    p = os.path.join(REF_DIR, f"gps-sdr-sim-{mode}")
    return p if os.path.exists(p) else None


def ref_data(name: str) -> str:
    return os.path.join(REF_DIR, "data", name)


def ref_lib(mode: str = "int"):
    p = os.path.join(REF_DIR, f"libgpssim_ref_{mode}.so")
    return ctypes.CDLL(p) if os.path.exists(p) else None
