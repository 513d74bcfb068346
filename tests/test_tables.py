"""Constant tables of the path: product (libgpusim host helpers) vs oracle vs the reference itself."""
import ctypes

import numpy as np
import pytest

import oracle_lib
import gps_sdr_sim_b200 as gs


def test_carrier_lut_product_equals_oracle():
    s, c = gs.carrier_lut()
    so, co = oracle_lib.carrier_lut()
    assert np.array_equal(s, so) and np.array_equal(c, co)


def test_carrier_lut_structure():
    # properties the kernels rely on (gpssim.c:15-83): amplitude 250, cos = sin advanced a quarter
    # cycle, half-cycle antisymmetry (the chip sign is folded into bit 8 of the table index)
    s, c = gs.carrier_lut()
    i = np.arange(512)
    assert s.max() == 250 and s.min() == -250
    assert np.array_equal(c, s[(i + 128) % 512])
    assert np.array_equal(s[(i + 256) % 512], -s)
    assert np.array_equal(c[(i + 256) % 512], -c)
    assert list(s[:8]) == [2, 5, 8, 11, 14, 17, 20, 23] and s[35] == 105


def test_carrier_lut_equals_reference_arrays():
    ref = oracle_lib.ref_lib("int")
    if ref is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    rs = np.array((ctypes.c_int * 512).in_dll(ref, "sinTable512")[:])
    rc = np.array((ctypes.c_int * 512).in_dll(ref, "cosTable512")[:])
    s, c = gs.carrier_lut()
    assert np.array_equal(s, rs) and np.array_equal(c, rc)


# first ten chips of PRN n in octal, ICD-GPS-200 Table 3-I
ICD_FIRST10 = dict(zip(range(1, 33), [
    0o1440, 0o1620, 0o1710, 0o1744, 0o1133, 0o1455, 0o1131, 0o1454, 0o1626, 0o1504, 0o1642, 0o1750,
    0o1764, 0o1772, 0o1775, 0o1776, 0o1156, 0o1467, 0o1633, 0o1715, 0o1746, 0o1763, 0o1063, 0o1706,
    0o1743, 0o1761, 0o1770, 0o1774, 0o1127, 0o1453, 0o1625, 0o1712]))


@pytest.mark.parametrize("prn", sorted(ICD_FIRST10))
def test_ca_code_known_answers(prn):
    ca = gs.ca_code(prn)
    first10 = int("".join(str(int(b)) for b in ca[:10]), 2)
    assert first10 == ICD_FIRST10[prn]
    assert ca.sum() == 512  # balanced Gold code: 512 ones, 511 zeros


def test_ca_code_product_equals_oracle_all_prn():
    for prn in range(1, 33):
        assert np.array_equal(gs.ca_code(prn), oracle_lib.ca_code(prn))
    with pytest.raises(gs.GpuSimError):
        gs.ca_code(33)
    with pytest.raises(gs.GpuSimError):
        gs.ca_code(0)


def test_ca_code_equals_reference_codegen():
    ref = oracle_lib.ref_lib("int")
    if ref is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    for prn in range(1, 33):
        ca = (ctypes.c_int * 1023)()
        ref.codegen(ca, prn)
        assert np.array_equal(np.array(ca[:]), gs.ca_code(prn)), prn


def test_pack_nav_bits():
    rng = np.random.default_rng(7)
    dwrd = rng.integers(0, 2**30, 60, dtype=np.uint64)
    for iword, ibit in [(9, 0), (10, 29), (33, 17), (59, 0), (59, 25)]:
        got = gs.pack_nav_bits(dwrd, iword, ibit)
        want = 0
        w, b = iword, ibit
        for k in range(32):
            bit = int((int(dwrd[w]) >> (29 - b)) & 1) if w < 60 else 0   # gpssim.c:2236
            want |= bit << (31 - k)
            b += 1
            if b >= 30:                                                   # gpssim.c:2225-2228
                b, w = 0, w + 1
        assert got == want, (iword, ibit)
