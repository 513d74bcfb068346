"""ctypes binding of libgpusim.so (include/gpusim.h) for tests and bench.py.

The product is the C-ABI library; this module is the thin Python host above it.  torch
is used only as plumbing - device buffers, streams, torch.distributed - and never for
the computation.  There is no CPU path: without the CUDA library and a B200 every
compute call raises.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import build as _build
from .table import NAV_EPH, NAV_FRAME, NAV_FRAME_REF, NAV_IONO, CEpochTable, EpochTable, epoch_bytes

_HERE = os.path.dirname(os.path.abspath(__file__))
# GPUSIM_LIB: load another build of the same library (kernel experiments); default = the in-tree build
_LIB_PATH = os.environ.get("GPUSIM_LIB") or os.path.join(_HERE, "libgpusim.so")
_lib = None


class GpuSimError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"gpusim status {status}: {message}")
        self.status = status


class _Config(ctypes.Structure):
    _fields_ = [("abi_version", ctypes.c_int32), ("device", ctypes.c_int32),
                ("samples_per_epoch", ctypes.c_int32), ("data_format", ctypes.c_int32),
                ("carrier_mode", ctypes.c_int32), ("max_batch_epochs", ctypes.c_int32),
                ("delt", ctypes.c_double)]


class Timing(ctypes.Structure):
    _fields_ = [("chain_ms", ctypes.c_float), ("synth_ms", ctypes.c_float), ("total_ms", ctypes.c_float),
                ("launches", ctypes.c_int32), ("fast_path", ctypes.c_int32), ("chain_overlapped", ctypes.c_int32)]


SINK_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t)

# every symbol include/gpusim.h declares (tests check the library exports all of them)
EXPORTS = (
    "gpusim_abi_version", "gpusim_strerror", "gpusim_last_error", "gpusim_create", "gpusim_destroy",
    "gpusim_epoch_bytes", "gpusim_generate_epochs", "gpusim_generate_epochs_to_sink",
    "gpusim_upload_table", "gpusim_generate_device", "gpusim_get_timing", "gpusim_set_option",
    "gpusim_carrier_lut", "gpusim_ca_code", "gpusim_pack_nav_bits", "gpusim_advance_carrier_f64",
    "gpusim_host_alloc", "gpusim_host_free", "gpusim_debug_guard_violations",
    "gpusim_nav_build", "gpusim_nav_read", "gpusim_nav_build_eph", "gpusim_nav_read_sbf",
)


def library_path() -> str:
    return _LIB_PATH


def load_library(build_if_missing: bool = True) -> ctypes.CDLL:
    """Load the in-tree CUDA library; fails loudly when it is absent (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        if not build_if_missing:
            raise FileNotFoundError(f"{_LIB_PATH} is missing - run `python -m gps_sdr_sim_b200.build`")
        _build.build()
    lib = ctypes.CDLL(_LIB_PATH)
    vp, i32, i64, sz = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_size_t
    lib.gpusim_abi_version.restype = ctypes.c_int
    lib.gpusim_strerror.restype = ctypes.c_char_p
    lib.gpusim_strerror.argtypes = [ctypes.c_int]
    lib.gpusim_last_error.restype = ctypes.c_char_p
    lib.gpusim_last_error.argtypes = [vp]
    lib.gpusim_create.restype = ctypes.c_int
    lib.gpusim_create.argtypes = [ctypes.POINTER(_Config), ctypes.POINTER(vp)]
    lib.gpusim_destroy.restype = None
    lib.gpusim_destroy.argtypes = [vp]
    lib.gpusim_epoch_bytes.restype = sz
    lib.gpusim_epoch_bytes.argtypes = [vp]
    lib.gpusim_generate_epochs.restype = ctypes.c_int
    lib.gpusim_generate_epochs.argtypes = [vp, ctypes.POINTER(CEpochTable), vp, sz]
    lib.gpusim_generate_epochs_to_sink.restype = ctypes.c_int
    lib.gpusim_generate_epochs_to_sink.argtypes = [vp, ctypes.POINTER(CEpochTable), SINK_FN, vp]
    lib.gpusim_upload_table.restype = ctypes.c_int
    lib.gpusim_upload_table.argtypes = [vp, ctypes.POINTER(CEpochTable)]
    lib.gpusim_generate_device.restype = ctypes.c_int
    lib.gpusim_generate_device.argtypes = [vp, i32, i32, vp, sz, vp]
    lib.gpusim_get_timing.restype = ctypes.c_int
    lib.gpusim_get_timing.argtypes = [vp, ctypes.POINTER(Timing)]
    lib.gpusim_set_option.restype = ctypes.c_int
    lib.gpusim_set_option.argtypes = [vp, ctypes.c_char_p, i64]
    lib.gpusim_carrier_lut.restype = None
    lib.gpusim_carrier_lut.argtypes = [vp, vp]
    lib.gpusim_ca_code.restype = ctypes.c_int
    lib.gpusim_ca_code.argtypes = [i32, vp]
    lib.gpusim_pack_nav_bits.restype = ctypes.c_uint32
    lib.gpusim_pack_nav_bits.argtypes = [vp, i32, i32, i32]
    lib.gpusim_host_alloc.restype = vp
    lib.gpusim_host_alloc.argtypes = [sz]
    lib.gpusim_host_free.restype = None
    lib.gpusim_host_free.argtypes = [vp]
    lib.gpusim_debug_guard_violations.restype = i64
    lib.gpusim_debug_guard_violations.argtypes = [vp]
    lib.gpusim_nav_build.restype = ctypes.c_int
    lib.gpusim_nav_build.argtypes = [vp, vp, i32]
    lib.gpusim_nav_read.restype = ctypes.c_int
    lib.gpusim_nav_read.argtypes = [vp, i32, i32, vp]
    lib.gpusim_nav_build_eph.restype = ctypes.c_int
    lib.gpusim_nav_build_eph.argtypes = [vp, vp, i32, vp, vp, i32]
    lib.gpusim_nav_read_sbf.restype = ctypes.c_int
    lib.gpusim_nav_read_sbf.argtypes = [vp, i32, i32, vp]
    lib.gpusim_advance_carrier_f64.restype = ctypes.c_double
    lib.gpusim_advance_carrier_f64.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_double, i32]
    _lib = lib
    return lib


# ---- host-only table helpers (no GPU needed) ------------------------------------------------
def carrier_lut():
    """(sin512, cos512) as the device uses them; equals sinTable512/cosTable512 (gpssim.c:15-83)."""
    s = np.empty(512, dtype=np.int32)
    c = np.empty(512, dtype=np.int32)
    load_library().gpusim_carrier_lut(s.ctypes.data, c.ctypes.data)
    return s, c


def ca_code(prn: int) -> np.ndarray:
    """1023 chips in {0,1}; equals codegen(ca, prn) (gpssim.c:132-171)."""
    ca = np.empty(1023, dtype=np.int32)
    rc = load_library().gpusim_ca_code(prn, ca.ctypes.data)
    if rc != 0:
        raise GpuSimError(rc, f"no C/A code for PRN {prn}")
    return ca


def pack_nav_bits(dwrd, iword: int, ibit: int) -> int:
    a = np.ascontiguousarray(np.asarray(dwrd, dtype=np.uint64))
    return int(load_library().gpusim_pack_nav_bits(a.ctypes.data, a.size, iword, ibit))


def advance_carrier_f64(carr_phase: float, f_carr: float, delt: float, n_samples: int) -> float:
    """chan[i].carr_phase after n_samples FLOAT_CARR_PHASE updates (gpssim.c:2245-2250), exactly."""
    return float(load_library().gpusim_advance_carrier_f64(carr_phase, f_carr, delt, n_samples))


class GpuSim:
    """One generator context on one GPU (gpusim_create ... gpusim_destroy)."""

    def __init__(self, samples_per_epoch: int, delt: float, data_format: int = 16, carrier_mode: int = 0,
                 max_batch_epochs: int = 256, device: int = 0):
        self._lib = load_library()
        self._ctx = ctypes.c_void_p()
        cfg = _Config(self._lib.gpusim_abi_version(), device, samples_per_epoch, data_format, carrier_mode,
                      max_batch_epochs, delt)
        rc = self._lib.gpusim_create(ctypes.byref(cfg), ctypes.byref(self._ctx))
        if rc != 0:
            raise GpuSimError(rc, self._lib.gpusim_last_error(None).decode())
        self.samples_per_epoch = samples_per_epoch
        self.data_format = data_format
        self.max_batch_epochs = max_batch_epochs
        self.device = device
        self.epoch_bytes = epoch_bytes(samples_per_epoch, data_format)
        assert self.epoch_bytes == self._lib.gpusim_epoch_bytes(self._ctx)

    @classmethod
    def for_table(cls, table: EpochTable, max_batch_epochs: int | None = None, device: int = 0) -> "GpuSim":
        return cls(table.samples_per_epoch, table.delt, table.data_format, table.carrier_mode,
                   max_batch_epochs or max(1, table.n_epochs), device)

    def close(self) -> None:
        if getattr(self, "_ctx", None) is not None and self._ctx:
            self._lib.gpusim_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int) -> None:
        if rc != 0:
            raise GpuSimError(rc, self._lib.gpusim_last_error(self._ctx).decode() or
                              self._lib.gpusim_strerror(rc).decode())

    def set_option(self, key: str, value: int) -> None:
        self._check(self._lib.gpusim_set_option(self._ctx, key.encode(), int(value)))

    # ---- host table in, host bytes out ---------------------------------------------------------
    def generate_epochs(self, table: EpochTable, out: np.ndarray | None = None, out_ptr: int | None = None,
                        out_capacity: int | None = None) -> np.ndarray | None:
        """gpusim_generate_epochs: replaces gpssim.c:2190-2288 for the table's epochs."""
        c = table.as_c()
        if out_ptr is not None:
            self._check(self._lib.gpusim_generate_epochs(self._ctx, ctypes.byref(c), out_ptr, out_capacity))
            return None
        need = table.n_epochs * self.epoch_bytes
        if out is None:
            out = np.empty(need, dtype=np.uint8)
        self._check(self._lib.gpusim_generate_epochs(self._ctx, ctypes.byref(c), out.ctypes.data, out.nbytes))
        return out[:need]

    def generate_epochs_to_sink(self, table: EpochTable, sink) -> None:
        """`sink(memoryview)` receives consecutive byte ranges in epoch order."""
        def _cb(_user, ptr, n):
            try:
                sink((ctypes.c_ubyte * n).from_address(ptr))
                return 0
            except Exception:  # noqa: BLE001 - reported through the status code
                return 1
        cb = SINK_FN(_cb)
        c = table.as_c()
        self._check(self._lib.gpusim_generate_epochs_to_sink(self._ctx, ctypes.byref(c), cb, None))

    # ---- device-resident path -----------------------------------------------------------------
    def upload_table(self, table: EpochTable) -> None:
        c = table.as_c()
        self._check(self._lib.gpusim_upload_table(self._ctx, ctypes.byref(c)))

    def generate_device(self, first_epoch: int, n_epochs: int, out_ptr: int, out_capacity: int,
                        stream: int | None = None) -> None:
        """out_ptr: 16-byte aligned device pointer (e.g. torch tensor .data_ptr()); stream: cudaStream_t."""
        self._check(self._lib.gpusim_generate_device(self._ctx, first_epoch, n_epochs, out_ptr, out_capacity,
                                                     stream))

    # ---- navigation data words on the device -------------------------------------
    def nav_build(self, frames: np.ndarray) -> None:
        """gpusim_nav_build: frames is a NAV_FRAME array, one element per generateNavMsg() call (gpssim.c:1467-1547)."""
        f = np.ascontiguousarray(frames, dtype=NAV_FRAME)
        self._check(self._lib.gpusim_nav_build(self._ctx, f.ctypes.data, f.size))

    def nav_build_eph(self, eph: np.ndarray, iono: np.ndarray, frames: np.ndarray) -> None:
        """gpusim_nav_build_eph: eph2sbf() (gpssim.c:490-665) on the device too - eph is a NAV_EPH array, iono one NAV_IONO,
        frames a NAV_FRAME_REF array naming their subframes by index into eph."""
        e = np.ascontiguousarray(eph, dtype=NAV_EPH)
        io = np.ascontiguousarray(iono, dtype=NAV_IONO).reshape(1)
        f = np.ascontiguousarray(frames, dtype=NAV_FRAME_REF)
        self._check(self._lib.gpusim_nav_build_eph(self._ctx, e.ctypes.data, e.size, io.ctypes.data, f.ctypes.data, f.size))

    def nav_read_sbf(self, first_eph: int, n_eph: int) -> np.ndarray:
        """The 5 x 10 source words (chan->sbf) the device made of ephemerides [first_eph, first_eph + n_eph)."""
        out = np.empty((n_eph, 5, 10), dtype=np.uint32)
        self._check(self._lib.gpusim_nav_read_sbf(self._ctx, first_eph, n_eph, out.ctypes.data))
        return out

    def nav_read(self, first_frame: int, n_frames: int) -> np.ndarray:
        """The 60 data words (chan->dwrd) of frames [first_frame, first_frame + n_frames), as built on the device."""
        out = np.empty((n_frames, 60), dtype=np.uint32)
        self._check(self._lib.gpusim_nav_read(self._ctx, first_frame, n_frames, out.ctypes.data))
        return out

    def guard_violations(self) -> int:
        """gpusim_debug_guard_violations: 0 = nothing wrote outside the context's device buffers (needs GPUSIM_GUARD=1)."""
        return int(self._lib.gpusim_debug_guard_violations(self._ctx))

    def timing(self) -> Timing:
        t = Timing()
        self._check(self._lib.gpusim_get_timing(self._ctx, ctypes.byref(t)))
        return t
