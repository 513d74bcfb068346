"""Epoch tables: the rows that cross the C ABI (include/gpusim.h, gpusim_epoch_table).

One row per (0.1 s epoch, channel slot) holding the state the reference's sample
loop starts from at gpssim.c:2190 (see Appendix A of SURVEY.md for the exact host
expression behind every column).  Structure-of-arrays, row index e*16 + slot.
"""
from __future__ import annotations

import ctypes
import struct
from dataclasses import dataclass, field

import numpy as np

MAX_CHAN = 16          # gpssim.h:16
CA_SEQ_LEN = 1023      # gpssim.h:36
SC01, SC08, SC16 = 1, 8, 16          # gpssim.h:77-79
CARRIER_INT, CARRIER_FLOAT = 0, 1    # gpssim.h:4 off / on

_COLUMNS = (
    ("prn", np.int32), ("f_code", np.float64), ("code_phase", np.float64), ("icode", np.int32),
    ("nav_bits", np.uint32), ("gain", np.int32), ("carr_phasestep", np.int32),
    ("carr_phase", np.uint32), ("f_carr", np.float64), ("carr_phase_f", np.float64),
)
_DUMP_EXTRA = (("iword", np.int32), ("ibit", np.int32))
# rows may reference device-built navigation frames instead of carrying their data bits
_NAV_REF = (("nav_frame", np.int32), ("iword", np.int32), ("ibit", np.int32))

# struct gpusim_nav_frame (include/gpusim.h): the input of one generateNavMsg() call, 256 bytes
NAV_FRAME = np.dtype([("sbf", np.uint32, (5, 10)), ("first", np.uint32, (10,)), ("tow_first", np.uint32),
                      ("tow", np.uint32), ("wn", np.uint32), ("reserved", np.uint32)])
assert NAV_FRAME.itemsize == 256
# struct gpusim_nav_eph / gpusim_nav_iono / gpusim_nav_frame_ref: what eph2sbf() reads of ephem_t / ionoutc_t (gpssim.h:101-146)
_EPH_DOUBLES = ("toe_sec", "toc_sec", "deltan", "cuc", "cus", "cic", "cis", "crc", "crs", "ecc", "sqrta", "m0", "omg0", "inc0",
                "aop", "omgdot", "idot", "af0", "af1", "af2", "tgd")
NAV_EPH = np.dtype([(n, np.float64) for n in _EPH_DOUBLES] +
                   [(n, np.int32) for n in ("toe_week", "iodc", "iode", "svhlth", "codeL2", "reserved")])
NAV_IONO = np.dtype([(n, np.float64) for n in ("alpha0", "alpha1", "alpha2", "alpha3", "beta0", "beta1", "beta2", "beta3", "A0", "A1")] +
                    [(n, np.int32) for n in ("vflg", "dtls", "tot", "wnt")])
NAV_FRAME_REF = np.dtype([("eph", np.int32), ("eph_first", np.int32), ("tow_first", np.uint32), ("tow", np.uint32),
                          ("wn", np.uint32), ("reserved", np.uint32)])
assert NAV_EPH.itemsize == 192 and NAV_IONO.itemsize == 96 and NAV_FRAME_REF.itemsize == 24


class CEpochTable(ctypes.Structure):
    """ctypes mirror of struct gpusim_epoch_table."""
    _fields_ = ([("n_epochs", ctypes.c_int32)] + [(name, ctypes.c_void_p) for name, _ in _COLUMNS] +
                [(name, ctypes.c_void_p) for name, _ in _NAV_REF])


def epoch_bytes(samples_per_epoch: int, data_format: int) -> int:
    """Bytes the reference writes per epoch: gpssim.c:2276 / :2283 / :2287."""
    if data_format == SC01:
        return samples_per_epoch // 4
    if data_format == SC08:
        return 2 * samples_per_epoch
    if data_format == SC16:
        return 4 * samples_per_epoch
    raise ValueError("data_format must be 1, 8 or 16")


@dataclass
class EpochTable:
    samples_per_epoch: int                 # iq_buff_size, gpssim.c:1878
    delt: float                            # gpssim.c:1881
    data_format: int = SC16                # -b
    carrier_mode: int = CARRIER_INT
    cols: dict = field(default_factory=dict)   # name -> np.ndarray [n_epochs, 16]
    nav_by_reference: bool = False             # rows carry nav_frame / iword / ibit, not nav_bits (gpusim_nav_build first)

    @property
    def n_epochs(self) -> int:
        return int(self.cols["prn"].shape[0])

    @property
    def epoch_bytes(self) -> int:
        return epoch_bytes(self.samples_per_epoch, self.data_format)

    def __getattr__(self, name):
        cols = self.__dict__.get("cols", {})
        if name in cols:
            return cols[name]
        raise AttributeError(name)

    def validate(self) -> None:
        n = self.n_epochs
        for name, dt in _COLUMNS + (_NAV_REF if self.nav_by_reference else ()):
            if name not in self.cols:
                raise ValueError(f"column {name} is missing")
            a = self.cols[name]
            if a.dtype != dt or a.shape != (n, MAX_CHAN) or not a.flags.c_contiguous:
                raise ValueError(f"column {name}: want C-contiguous {dt} [{n},{MAX_CHAN}]")

    def slice(self, first: int, count: int) -> "EpochTable":
        """Rows of epochs [first, first+count) - epochs are independent given their rows."""
        cols = {k: np.ascontiguousarray(v[first:first + count]) for k, v in self.cols.items()}
        return EpochTable(self.samples_per_epoch, self.delt, self.data_format, self.carrier_mode, cols, self.nav_by_reference)

    def with_format(self, data_format: int) -> "EpochTable":
        return EpochTable(self.samples_per_epoch, self.delt, data_format, self.carrier_mode, self.cols, self.nav_by_reference)

    def with_nav_references(self, nav_frame: np.ndarray) -> "EpochTable":
        """The same rows, their data bits to be taken on the device from frame nav_frame[e, slot] of the last
        GpuSim.nav_build() at (iword, ibit) - needs the iword / ibit columns a host dump carries."""
        cols = dict(self.cols)
        cols["nav_frame"] = np.ascontiguousarray(nav_frame, dtype=np.int32)
        return EpochTable(self.samples_per_epoch, self.delt, self.data_format, self.carrier_mode, cols, True)

    def as_c(self) -> CEpochTable:
        """The ctypes struct; keeps referencing self.cols, so keep `self` alive."""
        self.validate()
        c = CEpochTable()
        c.n_epochs = self.n_epochs
        for name, _ in _COLUMNS:
            setattr(c, name, self.cols[name].ctypes.data)
        if self.nav_by_reference:
            c.nav_bits = None
            for name, _ in _NAV_REF:
                setattr(c, name, self.cols[name].ctypes.data)
        return c

    def max_active(self) -> int:
        return int((self.cols["prn"] > 0).sum(axis=1).max()) if self.n_epochs else 0

    # ---- the dump written by integration/gpusim_hook.c (GPUSIM_DUMP=...) ----------------
    @staticmethod
    def load_dump(path: str) -> "EpochTable":
        with open(path, "rb") as f:
            raw = f.read()
        if raw[:8] != b"GPSTAB01":
            raise ValueError(f"{path}: not a gpusim table dump")
        n, N, fmt, mode, max_chan, _ = struct.unpack_from("<6i", raw, 8)
        (delt,) = struct.unpack_from("<d", raw, 32)
        if max_chan != MAX_CHAN:
            raise ValueError("dump has a different MAX_CHAN")
        off = 40
        cols = {}
        for name, dt in _COLUMNS + _DUMP_EXTRA:
            cnt = n * MAX_CHAN
            a = np.frombuffer(raw, dtype=np.dtype(dt).newbyteorder("<"), count=cnt, offset=off)
            off += cnt * np.dtype(dt).itemsize
            cols[name] = np.ascontiguousarray(a.reshape(n, MAX_CHAN).astype(dt))
        return EpochTable(N, delt, fmt, mode, cols)

    def save_npz(self, path: str) -> None:
        np.savez_compressed(path, samples_per_epoch=self.samples_per_epoch, delt=self.delt,
                            data_format=self.data_format, carrier_mode=self.carrier_mode, **self.cols)

    @staticmethod
    def load_npz(path: str) -> "EpochTable":
        z = np.load(path)
        cols = {}
        for name, dt in _COLUMNS + _DUMP_EXTRA:
            if name in z.files:
                cols[name] = np.ascontiguousarray(z[name].astype(dt))
        return EpochTable(int(z["samples_per_epoch"]), float(z["delt"]), int(z["data_format"]),
                          int(z["carrier_mode"]), cols)


def synthetic_table(n_epochs: int, samples_per_epoch: int = 260000, n_active: int = 13,
                    data_format: int = SC16, seed: int = 20141220, carrier_mode: int = CARRIER_INT) -> EpochTable:
    """Seeded synthetic rows inside the envelopes measured on the reference's own
    scenarios: every slot keeps its PRN for the whole
    table, the code phase and carrier phase are continuous from epoch to epoch the
    way the host produces them, Doppler drifts slowly.  carrier_mode=CARRIER_FLOAT fills the double
    carrier columns the way a FLOAT_CARR_PHASE host does (exact per-epoch advance)."""
    rng = np.random.default_rng(seed)
    N = int(samples_per_epoch)
    fs = 10.0 * N
    delt = 1.0 / fs
    E = int(n_epochs)
    cols = {name: np.zeros((E, MAX_CHAN), dtype=dt) for name, dt in _COLUMNS}
    prns = rng.permutation(32)[:n_active] + 1
    f_carr0 = rng.uniform(-3700.0, 3700.0, n_active)          # terrestrial Doppler, Hz
    f_drift = rng.uniform(-0.9, 0.9, n_active)                # Hz per second
    gain = rng.integers(36, 128, n_active)
    cp = rng.uniform(0.0, 1023.0, n_active)                   # chips
    ms_total = rng.integers(9 * 600, 58 * 600, n_active) + rng.integers(0, 20, n_active)
    phase = rng.integers(0, 2**32, n_active, dtype=np.uint64)
    nav_words = rng.integers(0, 2**30, (n_active, 64), dtype=np.uint64)
    t = np.arange(E) * 0.1
    for s in range(n_active):
        f_carr = f_carr0[s] + f_drift[s] * t + rng.normal(0.0, 0.05, E)
        f_code = 1.023e6 + f_carr / 1540.0
        step = np.rint(512.0 * 65536.0 * f_carr * delt).astype(np.int64)
        # chips advanced per epoch, tracked in exact integers of 1e-6 chip to stay in [0,1023)
        adv = f_code * 0.1
        code = (cp[s] + np.concatenate(([0.0], np.cumsum(adv[:-1])))) % 1023.0
        periods = np.floor((cp[s] + np.concatenate(([0.0], np.cumsum(adv[:-1])))) / 1023.0).astype(np.int64)
        ms = ms_total[s] + periods                       # code periods since the nav frame start
        icode = (ms % 20).astype(np.int32)
        bitno = ms // 20
        ph = (phase[s] + np.concatenate(([0], np.cumsum(step[:-1] * N)))) % (2**32)
        bits = np.zeros(E, dtype=np.uint32)
        flat = np.unpackbits(nav_words[s].astype(">u8").view(np.uint8))  # plenty of random data bits
        for k in range(32):
            bits |= (flat[(bitno + k) % flat.size].astype(np.uint32) << np.uint32(31 - k))
        cols["prn"][:, s] = prns[s]
        cols["f_code"][:, s] = f_code
        cols["code_phase"][:, s] = np.minimum(code, np.nextafter(1023.0, 0.0))
        cols["icode"][:, s] = icode
        cols["nav_bits"][:, s] = bits
        cols["gain"][:, s] = gain[s]
        cols["carr_phasestep"][:, s] = step.astype(np.int32)
        cols["carr_phase"][:, s] = ph.astype(np.uint32)
        cols["f_carr"][:, s] = f_carr
        if carrier_mode == CARRIER_FLOAT:
            from .api import advance_carrier_f64
            cph = float(rng.uniform(0.0, 1.0))
            for e in range(E):
                cols["carr_phase_f"][e, s] = cph
                cph = advance_carrier_f64(cph, float(f_carr[e]), delt, N)
    return EpochTable(N, delt, data_format, carrier_mode, cols)
