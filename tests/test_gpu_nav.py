"""SURVEY 8 f4 on the GPU: navigation data words built by k0_navmsg (gpusim_nav_build) against the oracle, and
tables whose rows reference those words (k0_navbits) against the same tables carrying their data bits."""
import ctypes

import numpy as np
import pytest

import oracle_lib
import gps_sdr_sim_b200 as gs
from gps_sdr_sim_b200 import NAV_FRAME

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _gpu(gpu_required):
    gs.load_library(build_if_missing=False)


def random_frames(n, seed):
    rng = np.random.default_rng(seed)
    f = np.zeros(n, dtype=NAV_FRAME)
    f["sbf"] = (rng.integers(0, 1 << 24, (n, 5, 10), dtype=np.uint64) << np.uint64(6)).astype(np.uint32)
    f["sbf"][:, :, 1] &= np.uint32(~(0x1FFFF << 13) & 0xFFFFFFFF)
    f["sbf"][:, 0, 2] &= np.uint32(~(0x3FF << 20) & 0xFFFFFFFF)
    f["first"] = (rng.integers(0, 1 << 24, (n, 10), dtype=np.uint64) << np.uint64(6)).astype(np.uint32)
    f["first"][:, 1] &= np.uint32(~(0x1FFFF << 13) & 0xFFFFFFFF)
    f["tow"] = rng.integers(0, 100800, n)
    f["tow_first"] = rng.integers(0, 100800, n)
    f["wn"] = rng.integers(0, 1024, n)
    return f


def oracle_words(frames):
    lib = oracle_lib.lib()
    lib.oracle_nav_frame.restype = None
    lib.oracle_nav_frame.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                     ctypes.c_void_p]
    out = np.zeros((frames.size, 60), dtype=np.uint32)
    for i, f in enumerate(frames):
        sbf = np.ascontiguousarray(f["sbf"])
        first = np.ascontiguousarray(f["first"])
        lib.oracle_nav_frame(sbf.ctypes.data, first.ctypes.data, int(f["tow_first"]), int(f["tow"]), int(f["wn"]),
                             out[i].ctypes.data)
    return out


@pytest.mark.parametrize("n", [1, 7, 64, 65, 3000])
def test_device_words_equal_the_oracle(n):
    frames = random_frames(n, 100 + n)
    with gs.GpuSim(260000, 1.0 / 2.6e6, 16, 0, max_batch_epochs=4) as sim:
        sim.nav_build(frames)
        got = sim.nav_read(0, n)
        assert np.array_equal(got, oracle_words(frames))
        if n > 10:      # a sub-range, and a second (smaller, then larger) build replaces the set
            assert np.array_equal(sim.nav_read(5, 3), got[5:8])
            sim.nav_build(frames[:3])
            assert np.array_equal(sim.nav_read(0, 3), got[:3])
            with pytest.raises(gs.GpuSimError):
                sim.nav_read(0, 4)
            more = random_frames(2 * n, 7)
            sim.nav_build(more)
            assert np.array_equal(sim.nav_read(0, 2 * n), oracle_words(more))


@pytest.mark.parametrize("fmt,mode", [(16, 0), (8, 0), (1, 0), (16, 1)])
def test_rows_by_reference_generate_the_same_bytes(fmt, mode):
    """The same synthetic scenario twice: rows carrying 32 data bits (packed on the host from the oracle's words), and
    rows carrying (frame, iword, ibit) resolved on the device.  Bytes must be identical, and equal to the oracle's."""
    E = 40
    table = gs.synthetic_table(E, 26000, 9, fmt, carrier_mode=mode)
    rng = np.random.default_rng(fmt + mode)
    frames = random_frames(37, 5)
    words = oracle_words(frames)
    nav_frame = rng.integers(0, frames.size, (E, gs.MAX_CHAN)).astype(np.int32)
    iword = rng.integers(0, 60, (E, gs.MAX_CHAN)).astype(np.int32)
    iword[:, 0] = 59                                     # runs off the end of the frame: the bits past word 59 read 0
    ibit = rng.integers(0, 30, (E, gs.MAX_CHAN)).astype(np.int32)
    bits = np.zeros((E, gs.MAX_CHAN), dtype=np.uint32)
    for e in range(E):
        for s in range(gs.MAX_CHAN):
            bits[e, s] = gs.pack_nav_bits(words[nav_frame[e, s]].astype(np.uint64), int(iword[e, s]), int(ibit[e, s]))
    by_value = gs.EpochTable(table.samples_per_epoch, table.delt, fmt, mode, dict(table.cols, nav_bits=bits))
    by_ref = gs.EpochTable(table.samples_per_epoch, table.delt, fmt, mode,
                           dict(table.cols, nav_frame=nav_frame, iword=iword, ibit=ibit), nav_by_reference=True)
    with gs.GpuSim.for_table(by_value) as sim:
        want = sim.generate_epochs(by_value)
        with pytest.raises(gs.GpuSimError):             # no frames built yet
            sim.generate_epochs(by_ref)
        sim.nav_build(frames)
        got = sim.generate_epochs(by_ref)
        assert np.array_equal(got, want)
        assert np.array_equal(want, oracle_lib.generate(by_value))
        # out-of-range references are refused, not read
        bad = gs.EpochTable(table.samples_per_epoch, table.delt, fmt, mode,
                            dict(by_ref.cols, nav_frame=np.full_like(nav_frame, frames.size)), nav_by_reference=True)
        with pytest.raises(gs.GpuSimError):
            sim.generate_epochs(bad)


def test_device_subframes_equal_the_reference_eph2sbf():
    """k0_eph2sbf on every valid ephemeris of the reference's RINEX file against the reference's own eph2sbf()
    (gpssim.c:490-665, through oracle/_ref/libgpssim_ref_int.so), and frames built from them against the oracle."""
    import ref_nav
    lib = ref_nav.ref_lib()
    if lib is None:
        pytest.fail("oracle/_ref/libgpssim_ref_int.so was not shipped to this GPU box")
    ephs, iono = ref_nav.broadcast_ephemerides(lib)
    want = np.stack([ref_nav.ref_subframes(lib, e, iono) for e in ephs])
    eph_arr, io = ref_nav.as_nav_eph(ephs), ref_nav.as_nav_iono(iono)
    rng = np.random.default_rng(8)
    n = 500
    refs = np.zeros(n, dtype=gs.NAV_FRAME_REF)
    refs["eph"] = rng.integers(0, len(ephs), n)
    refs["eph_first"] = rng.integers(0, len(ephs), n)
    refs["tow"] = rng.integers(0, 100800, n)
    refs["tow_first"] = rng.integers(0, 100800, n)
    refs["wn"] = rng.integers(0, 1024, n)
    with gs.GpuSim(260000, 1.0 / 2.6e6, 16, 0, max_batch_epochs=4) as sim:
        sim.nav_build_eph(eph_arr, io, refs)
        assert np.array_equal(sim.nav_read_sbf(0, len(ephs)), want)
        got = sim.nav_read(0, n)
        frames = np.zeros(n, dtype=NAV_FRAME)
        frames["sbf"] = want[refs["eph"]]
        frames["first"] = want[refs["eph_first"], 4]
        for k in ("tow", "tow_first", "wn"):
            frames[k] = refs[k]
        assert np.array_equal(got, oracle_words(frames))
        bad = refs.copy()
        bad["eph"][3] = len(ephs)
        with pytest.raises(gs.GpuSimError):
            sim.nav_build_eph(eph_arr, io, bad)
