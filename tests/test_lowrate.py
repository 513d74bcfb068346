"""Low chip rates (>= ~8 samples per chip, e.g. config 5 at 20 MS/s): the linear-model path of the integer-carrier
kernel (synth_lin, gpusim_core.h) writes the bytes of the reference's per-sample loop (gpssim.c:2192-2263)."""
import numpy as np
import pytest

import emu_lib
import oracle_lib
import gps_sdr_sim_b200 as gs


@pytest.mark.parametrize("n,chunk,fmt,nch", [
    (2000000, 4000, gs.SC16, 11),     # config 5: 20 MS/s, 19.5 samples per chip: <= 2 boundaries per run
    (2000000, 512, gs.SC08, 13),
    (1600000, 3200, gs.SC01, 16),     # 16 MS/s
    (1000000, 2000, gs.SC16, 12),     # 10 MS/s: 4 boundaries per run
    (840000, 1024, gs.SC08, 9),       # 8.4 MS/s, just inside the 4-boundary build
])
def test_linear_path_equals_oracle(n, chunk, fmt, nch):
    t = gs.synthetic_table(2, n, nch, fmt, seed=n // 1000 + fmt)
    want = oracle_lib.generate(t)
    got = emu_lib.generate(t, chunk, emu_lib.TUNED32)
    lin, fast = emu_lib.path_counts()
    assert lin > 8 * fast > 0, (lin, fast)           # the linear path does the bulk; binade edges / chips < 32 do not
    assert np.array_equal(got, want)
    assert np.array_equal(emu_lib.generate(t, chunk, emu_lib.TUNED32, accum=5), want)   # option lowrate=0
    assert emu_lib.path_counts()[0] == 0


def test_rates_outside_the_linear_range_do_not_take_it():
    t = gs.synthetic_table(1, 260000, 13, gs.SC08, seed=3)      # 2.6 MS/s: 12.6 boundaries per run
    emu_lib.generate(t, 520, emu_lib.TUNED32)
    assert emu_lib.path_counts()[0] == 0


def test_exact_tie_rows_take_the_per_sample_loop():
    """A step f_code*delt whose significand makes it an exact half-ulp tie in one of the binades [2^5, 2^10): there
    the chain is NOT linear (ties-to-even depends on the parity of the running sum), so such rows must never
    take the linear path - and the output must still be the reference's."""
    n, delt = 2000000, 1.0 / 20.0e6
    t = gs.synthetic_table(1, n, 8, gs.SC16, seed=11)
    # find f_code values near the real ones whose rounded product has 9..13 trailing zero bits in its significand
    found = 0
    act = np.argwhere(t.prn > 0)
    for (e, i) in act:
        base = t.f_code[e, i]
        for k in range(1, 400000):
            f = np.float64(base) + np.float64(k) * 2.0 ** -20
            d = np.float64(f) * np.float64(delt)
            m = int(np.float64(d).view(np.uint64)) & ((1 << 52) - 1)
            tz = (m & -m).bit_length() - 1 if m else 99
            ex = (int(np.float64(d).view(np.uint64)) >> 52) - 1023
            if 5 <= ex + tz + 1 <= 9:
                t.f_code[e, i] = f
                found += 1
                break
        if found >= 3:
            break
    assert found >= 3
    want = oracle_lib.generate(t)
    assert np.array_equal(emu_lib.generate(t, 4000, emu_lib.TUNED32), want)


def _tie_step(delt, lo_binade, hi_binade, rng):
    """f_code near 1.023 MHz whose rounded product f_code*delt is an exact half-ulp tie in a binade of the range."""
    while True:
        f = np.float64(1.023e6 + rng.uniform(-28.0, 28.0))
        d = np.float64(f) * np.float64(delt)
        bits = int(d.view(np.uint64))
        m = bits & ((1 << 52) - 1)
        tz = (m & -m).bit_length() - 1 if m else 99
        if lo_binade <= (bits >> 52) - 1023 + tz + 1 <= hi_binade:
            return float(d)


def test_one_run_linear_model_equals_the_per_sample_chain():
    """Property: from any state with floor(x) >= 32 that lin_ok() admits, one run through synth_lin leaves the
    accumulators, the code phase (bit for bit) and the carrier phase of the per-sample chain."""
    rng = np.random.default_rng(2014)
    for fs, nb in ((20.0e6, 2), (16.4e6, 2), (10.0e6, 4), (8.2e6, 4)):
        delt = 1.0 / fs
        for _ in range(3000):
            f = np.float64(1.023e6 + rng.uniform(-28.5, 28.5))
            d = float(np.float64(f) * np.float64(delt))
            bits = int(np.float64(d).view(np.uint64))
            m = bits & ((1 << 52) - 1)
            tz = (m & -m).bit_length() - 1 if m else 99
            if 5 <= (bits >> 52) - 1023 + tz + 1 <= 9:
                continue                                   # tie rows never reach synth_lin (next test)
            c0 = int(rng.integers(32, 1000))
            if (c0 ^ (c0 + nb + 1)) > c0:
                continue                                   # lin_ok(): a power of two within reach
            # fractions on and next to the chip boundaries as well as anywhere
            frac = rng.choice([0.0, 2.0 ** -40, 1.0 - 2.0 ** -40, rng.uniform(0.0, 1.0), (1.0 - d * float(rng.integers(1, 30))) % 1.0])
            x0 = float(np.float64(c0) + np.float64(frac % 1.0))
            rinv = 1.0 / (1.023e6 * delt) * (1.0 + rng.uniform(-3e-5, 3e-5))
            ok, (xl, xf) = emu_lib.lin_matches_fast(x0, d, int(rng.integers(0, 2 ** 32)), int(rng.integers(-2 ** 20, 2 ** 20)) << 7,
                                                   int(rng.integers(1, 33)), nb, rinv)
            assert ok, (fs, x0, d, xl, xf)


def test_a_tie_step_from_an_odd_significand_is_where_the_linear_model_breaks():
    """Why upload routes tie rows away from synth_lin: with x odd and d an exact tie in x's binade the reference's
    first sum rounds to even, the linear model adds RN(d) - the code phases differ by one ulp for the rest of
    the chunk."""
    rng = np.random.default_rng(7)
    delt = 1.0 / 20.0e6
    d = _tie_step(delt, 9, 9, rng)                            # tie in [512, 1024)
    broke = 0
    for _ in range(200):
        x0 = float(np.float64(rng.integers(520, 1000)) + np.float64(rng.uniform(0, 1)))
        xb = int(np.float64(x0).view(np.uint64)) | 1           # odd significand
        x0 = float(np.uint64(xb).view(np.float64))
        ok, (xl, xf) = emu_lib.lin_matches_fast(x0, d, 0, 1 << 10, 5, 2, 1.0 / (1.023e6 * delt))
        broke += not ok
    assert broke > 0
