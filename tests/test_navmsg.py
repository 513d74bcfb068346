"""SURVEY 8 f4 - navigation data words: the device algorithm (k0_navmsg / k0_navbits, run on the CPU through
tests/emu) and the oracle's restatement against the reference's own generateNavMsg() (gpssim.c:1467-1547) and
computeChecksum() (gpssim.c:693-756), called through oracle/_ref/libgpssim_ref_int.so; the host shim's frame
book-keeping (GPUSIM_NAV_DEVICE) against what the reference host really transmits in its own scenarios."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import emu_lib
import oracle_lib
from gps_sdr_sim_b200 import NAV_FRAME, pack_nav_bits

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- the reference's channel_t (gpssim.h:160-183, integer carrier build) -----------------------------
class GpsTime(ctypes.Structure):
    _fields_ = [("week", ctypes.c_int), ("sec", ctypes.c_double)]


class Range(ctypes.Structure):
    _fields_ = [("g", GpsTime), ("range", ctypes.c_double), ("rate", ctypes.c_double), ("d", ctypes.c_double),
                ("azel", ctypes.c_double * 2), ("iono_delay", ctypes.c_double)]


class Channel(ctypes.Structure):
    _fields_ = [("prn", ctypes.c_int), ("ca", ctypes.c_int * 1023), ("f_carr", ctypes.c_double),
                ("f_code", ctypes.c_double), ("carr_phase", ctypes.c_uint), ("carr_phasestep", ctypes.c_int),
                ("code_phase", ctypes.c_double), ("g0", GpsTime), ("sbf", (ctypes.c_ulong * 10) * 5),
                ("dwrd", ctypes.c_ulong * 60), ("iword", ctypes.c_int), ("ibit", ctypes.c_int), ("icode", ctypes.c_int),
                ("dataBit", ctypes.c_int), ("codeCA", ctypes.c_int), ("azel", ctypes.c_double * 2), ("rho0", Range)]


@pytest.fixture(scope="module")
def ref():
    lib = oracle_lib.ref_lib("int")
    if lib is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    lib.generateNavMsg.restype = ctypes.c_int
    lib.generateNavMsg.argtypes = [GpsTime, ctypes.POINTER(Channel), ctypes.c_int]
    lib.computeChecksum.restype = ctypes.c_ulong
    lib.computeChecksum.argtypes = [ctypes.c_ulong, ctypes.c_int]
    return lib


def random_sbf(rng):
    """Source words the way eph2sbf() leaves them: 24 bits in bits 29..6, TOW / week fields still empty."""
    sbf = (rng.integers(0, 1 << 24, (5, 10), dtype=np.uint64) << np.uint64(6)).astype(np.uint32)
    sbf[:, 1] &= np.uint32(~(0x1FFFF << 13) & 0xFFFFFFFF)      # hand-over word: TOW count ORed in later
    sbf[0, 2] &= np.uint32(~(0x3FF << 20) & 0xFFFFFFFF)        # subframe 1 word 3: week number ORed in later
    return sbf


def set_sbf(chan, sbf):
    for a in range(5):
        for b in range(10):
            chan.sbf[a][b] = int(sbf[a, b])


def oracle_frame(f):
    out = np.zeros(60, dtype=np.uint32)
    sbf = np.ascontiguousarray(f["sbf"], dtype=np.uint32)
    first = np.ascontiguousarray(f["first"], dtype=np.uint32)
    lib = oracle_lib.lib()
    lib.oracle_nav_frame.restype = None
    lib.oracle_nav_frame.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                     ctypes.c_void_p]
    lib.oracle_nav_frame(sbf.ctypes.data, first.ctypes.data, int(f["tow_first"]), int(f["tow"]), int(f["wn"]), out.ctypes.data)
    return out


def frame_request(sbf, first, tow_first, g0_week, g0_sec):
    f = np.zeros((), dtype=NAV_FRAME)
    f["sbf"] = sbf
    f["first"] = first
    f["tow_first"] = tow_first
    f["tow"] = int(g0_sec) // 6
    f["wn"] = g0_week % 1024
    return f


def test_every_word_against_computeChecksum(ref):
    """One word at a time: random source bits, all four D29*/D30* states, with and without the solved tail."""
    rng = np.random.default_rng(4)
    lib = oracle_lib.lib()
    for src in rng.integers(0, 1 << 32, 4000, dtype=np.uint64):
        for nib in (0, 1):
            want = ref.computeChecksum(int(src), nib) & 0xFFFFFFFF
            got = _emu_word(int(src), nib)
            assert got == want, hex(int(src))
    # a word with the solved tail ends in two zero bits whatever came before: subframes are independent
    for src in rng.integers(0, 1 << 32, 2000, dtype=np.uint64):
        assert ref.computeChecksum(int(src), 1) & 3 == 0


def _emu_word(src, nib):
    lib = emu_lib.lib()
    if not hasattr(lib, "_nav_word_ready"):
        lib.emu_nav_word.restype = ctypes.c_uint32
        lib.emu_nav_word.argtypes = [ctypes.c_uint32, ctypes.c_int]
        lib._nav_word_ready = True
    return int(lib.emu_nav_word(src & 0xFFFFFFFF, nib))


@pytest.mark.parametrize("week,sec", [(1823, 345600.0), (1823, 12.3), (2047, 604790.0), (1024, 0.0), (1823, 604769.9)])
def test_first_frame_of_a_channel(ref, week, sec):
    """generateNavMsg(init=1), what allocateChannel() calls for a new satellite (gpssim.c:1604-1608)."""
    rng = np.random.default_rng(int(sec) + week)
    for _ in range(40):
        chan = Channel()
        sbf = random_sbf(rng)
        set_sbf(chan, sbf)
        assert ref.generateNavMsg(GpsTime(week, sec), ctypes.byref(chan), 1) == 1
        want = np.array(chan.dwrd[:], dtype=np.uint64).astype(np.uint32)
        f = frame_request(sbf, sbf[4], int(chan.g0.sec) // 6, chan.g0.week, chan.g0.sec)
        assert np.array_equal(emu_lib.nav_build(f.reshape(1))[0], want)
        assert np.array_equal(oracle_frame(f), want)


def test_refresh_sequence_with_an_ephemeris_switch(ref):
    """generateNavMsg(init=0) every 30 s (gpssim.c:2296-2304): words 0..9 are the previous frame's words 50..59, also
    when eph2sbf() replaced the subframes in between (gpssim.c:2318-2330); across the end of the week too."""
    rng = np.random.default_rng(77)
    for start in (345600.0, 604680.0, 17.0):
        chan = Channel()
        sbf = random_sbf(rng)
        set_sbf(chan, sbf)
        g = GpsTime(1823, start)
        ref.generateNavMsg(g, ctypes.byref(chan), 1)
        built_with = sbf.copy()
        tow = int(chan.g0.sec) // 6
        for step in range(1, 9):
            sec, week = g.sec + 30.0, g.week
            if sec >= 604800.0:
                sec, week = sec - 604800.0, week + 1
            g = GpsTime(week, sec)
            ref.generateNavMsg(g, ctypes.byref(chan), 0)     # reads the subframes as they are NOW
            want = np.array(chan.dwrd[:], dtype=np.uint64).astype(np.uint32)
            f = frame_request(sbf, built_with[4], tow + 5, chan.g0.week, chan.g0.sec)
            assert np.array_equal(emu_lib.nav_build(f.reshape(1))[0], want), (start, step)
            assert np.array_equal(oracle_frame(f), want), (start, step)
            built_with = sbf.copy()
            tow = int(chan.g0.sec) // 6
            if step in (2, 5):                               # a new ephemeris set AFTER the call, as in the reference
                sbf = random_sbf(rng)
                set_sbf(chan, sbf)


def test_row_bits_equal_the_host_packing():
    rng = np.random.default_rng(9)
    dwrd = rng.integers(0, 1 << 30, 60, dtype=np.uint64)
    for iword in range(60):
        for ibit in range(30):
            assert emu_lib.nav_row_bits(dwrd.astype(np.uint32), iword, ibit) == pack_nav_bits(dwrd, iword, ibit)


SCENARIOS = {
    # satellite add @90 s, drop @150 s, slot reuse @180 s
    "satellite": ["-u", "satellite.csv", "-i", "-s", "2600000", "-b", "16"],
    "static_long": ["-l", "30.286502,120.032669,100", "-d", "400", "-s", "1000000", "-b", "1"],
}


@pytest.mark.parametrize("name", list(SCENARIOS))
def test_host_shim_frames_are_what_the_reference_transmits(name, tmp_path):
    """The reference host with the binding, GPUSIM_NAV_DEVICE=1 in dry-run mode: every frame request the shim derives
    from chan[i].sbf / chan[i].g0, built by the device algorithm (on the CPU here), equals chan[i].dwrd as the host's
    own generateNavMsg() left it - including the refresh copies and newly allocated channels."""
    host = os.path.join(ROOT, "integration", "_build", "gps-sdr-sim-gpu-int")
    if not os.path.exists(host) or oracle_lib.ref_binary("int") is None:
        pytest.skip("integration/_build or oracle/_ref not built (needs /root/reference)")
    dump = str(tmp_path / "nav.bin")
    argv = [a if not a.endswith((".csv", ".txt")) else oracle_lib.ref_data(a) for a in SCENARIOS[name]]
    env = dict(os.environ, GPUSIM_DRYRUN="1", GPUSIM_NAV_DEVICE="1", GPUSIM_NAV_DUMP=dump)
    subprocess.run([host, "-e", oracle_lib.ref_data("brdc3540.14n"), *argv, "-o", "/dev/null"], env=env, check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=600)
    rec = np.dtype([("frame", NAV_FRAME), ("dwrd", np.uint32, (60,))])
    recs = np.fromfile(dump, dtype=rec)
    assert recs.size >= 100                                     # ~10 refreshes x ~10 satellites (+ one per batch)
    got = emu_lib.nav_build(recs["frame"])
    assert np.array_equal(got, recs["dwrd"])
    assert all(np.array_equal(oracle_frame(f), w) for f, w in zip(recs["frame"][:200], recs["dwrd"][:200]))
    # (a refresh without an ephemeris switch asks for the same words a first frame would: the two kinds only differ
    # after eph2sbf() replaced the subframes, which test_refresh_sequence_with_an_ephemeris_switch constructs)


# ---- eph2sbf (gpssim.c:490-665) -------------------------------------------------------------------------------------
def test_subframes_of_every_broadcast_ephemeris_match_eph2sbf():
    """Every valid ephemeris of the reference's RINEX file (as its own readRinexNavAll() parses it), with and without
    ionosphere / UTC parameters, plus perturbed copies that reach negative and large field values."""
    import ref_nav
    lib = ref_nav.ref_lib()
    if lib is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    ephs, iono = ref_nav.broadcast_ephemerides(lib)
    assert len(ephs) > 300 and iono.vflg == 1
    rng = np.random.default_rng(12)
    for variant in range(3):
        io = ref_nav.IonoUtc.from_buffer_copy(iono)
        cases = list(ephs)
        if variant == 1:
            io.vflg = 0                                       # no ionosphere data: page 25 of subframe 4
        if variant == 2:                                      # sign changes and other magnitudes in every scaled field
            cases = []
            for e in ephs[:120]:
                p = ref_nav.Ephem.from_buffer_copy(e)
                for n in ref_nav._EPH_D:
                    setattr(p, n, getattr(p, n) * float(rng.choice([-1.0, 1.0])) * float(rng.uniform(0.2, 1.9)))
                p.ecc, p.sqrta = abs(p.ecc), abs(p.sqrta)     # unsigned fields (a negative value is undefined in the reference)
                p.iodc, p.iode = int(rng.integers(0, 1024)), int(rng.integers(0, 256))
                cases.append(p)
            for n in ("alpha0", "alpha1", "alpha2", "alpha3", "beta0", "beta1", "beta2", "beta3", "A0", "A1"):
                setattr(io, n, -getattr(io, n) * 1.37)
        want = np.stack([ref_nav.ref_subframes(lib, e, io) for e in cases])
        got = emu_lib.eph2sbf(ref_nav.as_nav_eph(cases), ref_nav.as_nav_iono(io))
        assert np.array_equal(got, want), variant


def test_host_shim_names_the_right_ephemerides(tmp_path):
    """GPUSIM_NAV_DEVICE=2 in dry-run mode on the satellite scenario: for every frame the shim registers, the subframes
    the device algorithm makes of the ephemerides it names equal chan[i].sbf as the host's eph2sbf() left it (the frame
    request of mode 1 carries a copy), and the words built from them equal chan[i].dwrd."""
    host = os.path.join(ROOT, "integration", "_build", "gps-sdr-sim-gpu-int")
    if not os.path.exists(host) or oracle_lib.ref_binary("int") is None:
        pytest.skip("integration/_build or oracle/_ref not built (needs /root/reference)")
    from gps_sdr_sim_b200 import NAV_EPH, NAV_IONO
    dump = str(tmp_path / "nav2.bin")
    env = dict(os.environ, GPUSIM_DRYRUN="1", GPUSIM_NAV_DEVICE="2", GPUSIM_NAV_DUMP=dump)
    subprocess.run([host, "-e", oracle_lib.ref_data("brdc3540.14n"), "-u", oracle_lib.ref_data("satellite.csv"), "-i",
                    "-s", "2600000", "-b", "16", "-o", "/dev/null"], env=env, check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=600)
    rec = np.dtype([("frame", NAV_FRAME), ("dwrd", np.uint32, (60,)), ("eph", NAV_EPH), ("eph_first", NAV_EPH), ("iono", NAV_IONO)])
    recs = np.fromfile(dump, dtype=rec)
    assert recs.size >= 100
    for r in recs:
        sbf = emu_lib.eph2sbf(r["eph"].reshape(1), r["iono"])[0]
        first = emu_lib.eph2sbf(r["eph_first"].reshape(1), r["iono"])[0][4]
        assert np.array_equal(sbf, r["frame"]["sbf"]) and np.array_equal(first, r["frame"]["first"])
        f = r["frame"].copy()
        f["sbf"], f["first"] = sbf, first
        assert np.array_equal(emu_lib.nav_build(f.reshape(1))[0], r["dwrd"])
