"""ctypes mirrors of the reference's ephem_t / ionoutc_t / channel_t (gpssim.h:81-183, integer-carrier build) and helpers
that call the reference's own readRinexNavAll(), eph2sbf() and generateNavMsg() through oracle/_ref/libgpssim_ref_int.so.
TEST INFRASTRUCTURE ONLY."""
import ctypes

import numpy as np

import oracle_lib
from gps_sdr_sim_b200 import NAV_EPH, NAV_IONO

MAX_SAT, EPHEM_ARRAY_SIZE = 32, 13


class GpsTime(ctypes.Structure):
    _fields_ = [("week", ctypes.c_int), ("sec", ctypes.c_double)]


class DateTime(ctypes.Structure):
    _fields_ = [("y", ctypes.c_int), ("m", ctypes.c_int), ("d", ctypes.c_int), ("hh", ctypes.c_int), ("mm", ctypes.c_int),
                ("sec", ctypes.c_double)]


_EPH_D = ("deltan", "cuc", "cus", "cic", "cis", "crc", "crs", "ecc", "sqrta", "m0", "omg0", "inc0", "aop", "omgdot", "idot",
          "af0", "af1", "af2", "tgd")


class Ephem(ctypes.Structure):
    _fields_ = ([("vflg", ctypes.c_int), ("t", DateTime), ("toc", GpsTime), ("toe", GpsTime), ("iodc", ctypes.c_int),
                 ("iode", ctypes.c_int)] + [(n, ctypes.c_double) for n in _EPH_D] +
                [("svhlth", ctypes.c_int), ("codeL2", ctypes.c_int), ("n", ctypes.c_double), ("sq1e2", ctypes.c_double),
                 ("A", ctypes.c_double), ("omgkdot", ctypes.c_double)])


class IonoUtc(ctypes.Structure):
    _fields_ = ([("enable", ctypes.c_int), ("vflg", ctypes.c_int)] +
                [(n, ctypes.c_double) for n in ("alpha0", "alpha1", "alpha2", "alpha3", "beta0", "beta1", "beta2", "beta3", "A0", "A1")] +
                [(n, ctypes.c_int) for n in ("dtls", "tot", "wnt", "dtlsf", "dn", "wnlsf")])


def ref_lib():
    lib = oracle_lib.ref_lib("int")
    if lib is None:
        return None
    lib.readRinexNavAll.restype = ctypes.c_int
    lib.readRinexNavAll.argtypes = [ctypes.c_void_p, ctypes.POINTER(IonoUtc), ctypes.c_char_p]
    lib.eph2sbf.restype = None
    lib.eph2sbf.argtypes = [Ephem, IonoUtc, ctypes.c_void_p]
    return lib


def broadcast_ephemerides(lib, rinex=None):
    """-> (list of valid Ephem, IonoUtc) as the reference reads them from its RINEX file (gpssim.c:818-1160)"""
    table = ((Ephem * MAX_SAT) * EPHEM_ARRAY_SIZE)()
    iono = IonoUtc()
    n = lib.readRinexNavAll(ctypes.byref(table), ctypes.byref(iono), (rinex or oracle_lib.ref_data("brdc3540.14n")).encode())
    assert n > 0
    ephs = [table[i][sv] for i in range(n) for sv in range(MAX_SAT) if table[i][sv].vflg == 1]
    return ephs, iono


def as_nav_eph(ephs):
    out = np.zeros(len(ephs), dtype=NAV_EPH)
    for i, e in enumerate(ephs):
        out[i]["toe_sec"], out[i]["toc_sec"], out[i]["toe_week"] = e.toe.sec, e.toc.sec, e.toe.week
        for n in _EPH_D:
            out[i][n] = getattr(e, n)
        for n in ("iodc", "iode", "svhlth", "codeL2"):
            out[i][n] = getattr(e, n)
    return out


def as_nav_iono(iono):
    out = np.zeros((), dtype=NAV_IONO)
    for n in NAV_IONO.names:
        out[n] = getattr(iono, n)
    return out


def ref_subframes(lib, eph, iono):
    """chan->sbf as the reference's eph2sbf() fills it: [5, 10] uint32"""
    sbf = ((ctypes.c_ulong * 10) * 5)()
    lib.eph2sbf(eph, iono, ctypes.byref(sbf))
    return np.array([[sbf[a][b] for b in range(10)] for a in range(5)], dtype=np.uint64).astype(np.uint32)
