"""ctypes access to tests/emu/libgpusim_emu.so: the device algorithms run on the CPU.
TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "emu")
CSRC = os.path.join(ROOT, "gps_sdr_sim_b200", "csrc")
_lib = None

TUNED32, TUNED16, GENERIC = 0, 1, 2


def build(sanitized: bool = False) -> str:
    """sanitized: the same code with AddressSanitizer + UndefinedBehaviorSanitizer (tests/test_emu_sanitized.py
    runs it in a subprocess with the ASan runtime preloaded)."""
    so = os.path.join(EMU_DIR, "libgpusim_emu_asan.so" if sanitized else "libgpusim_emu.so")
    deps = [os.path.join(EMU_DIR, "emu.cpp"), os.path.join(CSRC, "gpusim_core.h"),
            os.path.join(CSRC, "gpusim_tables.cpp"), os.path.join(CSRC, "gpusim_tables.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        flags = ["-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all"] if sanitized else ["-O2"]
        subprocess.run(["g++", *flags, "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wall",
                        "-Wno-unknown-pragmas", "-I", os.path.join(ROOT, "include"), "-I", CSRC,
                        deps[0], deps[2], "-o", so], check=True, capture_output=True)
    return so


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build(sanitized=os.environ.get("GPUSIM_EMU_SANITIZED") == "1"))
        _lib.emu_generate.restype = ctypes.c_int
        _lib.emu_generate.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        _lib.emu_code_chain.restype = None
        _lib.emu_code_chain.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_void_p, ctypes.c_void_p]
        _lib.emu_phase_chain.restype = ctypes.c_double
        _lib.emu_phase_chain.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_void_p, ctypes.c_void_p]
        _lib.emu_phase_chain_signed.restype = ctypes.c_double
        _lib.emu_phase_chain_signed.argtypes = _lib.emu_phase_chain.argtypes
        _lib.emu_lin_runs.restype = ctypes.c_long
        _lib.emu_fast_runs.restype = ctypes.c_long
        _lib.emu_phase_chain_tab.restype = ctypes.c_double
        _lib.emu_phase_chain_tab.argtypes = _lib.emu_phase_chain.argtypes
        _lib.emu_nav_build.restype = None
        _lib.emu_nav_build.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        _lib.emu_nav_row_bits.restype = ctypes.c_uint32
        _lib.emu_nav_row_bits.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    return _lib


def generate(table, chunk: int = 512, kernel: int = TUNED32, force_wrap: bool = False,
             chain_replay: bool = False, accum: int = 1) -> np.ndarray:
    """accum: 1 = what the product launches (k2_lean for integer carrier, k2_synth<AccF32x2> for double carrier),
    3 = k2_synth<AccF32x2> for integer carrier (option lean=0), 0 = k2_synth<AccWide> (option accum=0),
    5 = the lean kernel without the linear low-chip-rate path (option lowrate=0)."""
    out = np.zeros(table.n_epochs * table.epoch_bytes, dtype=np.uint8)
    c = table.as_c()
    rc = lib().emu_generate(ctypes.addressof(c), table.samples_per_epoch, table.delt, table.data_format, chunk,
                            kernel, int(force_wrap), int(chain_replay), accum, int(table.carrier_mode), out.ctypes.data)
    if rc != 0:
        raise ValueError("table outside the selected kernel's ranges")
    return out


def lin_matches_fast(x0: float, d: float, phs0: int, steps: int, prn: int, nb: int, rinv: float):
    """One run of 32 samples of one channel through synth_lin<nb> and synth_fast_g from the same state:
    (everything agrees bit for bit, code phase after the run [linear model, per-sample chain])."""
    xe = (ctypes.c_double * 2)()
    ok = lib().emu_lin_matches_fast(ctypes.c_double(x0), ctypes.c_double(d), ctypes.c_uint32(phs0), ctypes.c_int32(steps),
                                    prn, nb, ctypes.c_double(rinv), xe)
    return bool(ok), (xe[0], xe[1])


def path_counts():
    """(run, channel) pairs of the last generate() that took (synth_lin, synth_fast_g)."""
    return int(lib().emu_lin_runs()), int(lib().emu_fast_runs())


def code_chain(x0: float, d: float, n: int, every: int, replay: bool = False):
    k = (n + every - 1) // every
    x = np.empty(k, dtype=np.float64)
    w = np.empty(k, dtype=np.int32)
    lib().emu_code_chain(x0, d, n, every, int(replay), x.ctypes.data, w.ctypes.data)
    return x, w


def phase_chain(x0: float, d: float, modulus: float, n_end: int, every: int):
    """(checkpoints at j*every <= n_end, wraps, final value) of the generic exact walk."""
    k = n_end // every + 1
    x = np.empty(k, dtype=np.float64)
    w = np.empty(k, dtype=np.int32)
    end = lib().emu_phase_chain(x0, d, modulus, n_end, every, x.ctypes.data, w.ctypes.data)
    # the sign-specialised instantiations (what the device and the host advance use) must walk identically
    xs, ws = np.empty_like(x), np.empty_like(w)
    end_s = lib().emu_phase_chain_signed(x0, d, modulus, n_end, every, xs.ctypes.data, ws.ctypes.data)
    assert np.array_equal(x.view(np.uint64), xs.view(np.uint64)) and np.array_equal(w, ws)
    assert np.float64(end).view(np.uint64) == np.float64(end_s).view(np.uint64)
    # and so must the tabulated walk (phase_chain_tab: k1_chain on the device, gpusim_advance_carrier_f64 on the host)
    if modulus <= 1024.0:
        xt, wt = np.empty_like(x), np.empty_like(w)
        end_t = lib().emu_phase_chain_tab(x0, d, modulus, n_end, every, xt.ctypes.data, wt.ctypes.data)
        assert np.array_equal(x.view(np.uint64), xt.view(np.uint64)) and np.array_equal(w, wt)
        assert np.float64(end).view(np.uint64) == np.float64(end_t).view(np.uint64)
    return x, w, end


def nav_build(frames: np.ndarray) -> np.ndarray:
    """k0_navmsg on the CPU: the 60 data words of every frame (NAV_FRAME array)."""
    from gps_sdr_sim_b200.table import NAV_FRAME
    f = np.ascontiguousarray(frames, dtype=NAV_FRAME)
    out = np.zeros((f.size, 60), dtype=np.uint32)
    lib().emu_nav_build(f.ctypes.data, f.size, out.ctypes.data)
    return out


def nav_row_bits(dwrd60: np.ndarray, iword: int, ibit: int) -> int:
    a = np.ascontiguousarray(dwrd60, dtype=np.uint32)
    return int(lib().emu_nav_row_bits(a.ctypes.data, iword, ibit))


def eph2sbf(eph: np.ndarray, iono: np.ndarray) -> np.ndarray:
    """k0_eph2sbf on the CPU: [n, 5, 10] source words from a NAV_EPH array and one NAV_IONO."""
    from gps_sdr_sim_b200.table import NAV_EPH, NAV_IONO
    e = np.ascontiguousarray(eph, dtype=NAV_EPH)
    io = np.ascontiguousarray(iono, dtype=NAV_IONO).reshape(1)
    out = np.zeros((e.size, 5, 10), dtype=np.uint32)
    f = lib().emu_eph2sbf
    f.restype = None
    f.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    f(e.ctypes.data, e.size, io.ctypes.data, out.ctypes.data)
    return out
