"""K1: the binade-jump walk of the code-phase chain equals the reference's N sequential adds."""
import numpy as np
import pytest

import emu_lib
import oracle_lib


def _cases():
    rng = np.random.default_rng(20141220)
    cases = []
    for fs, n in ((1.0e6, 100000), (2.6e6, 260000), (4.0e6, 400000), (20.0e6, 2000000)):
        delt = 1.0 / fs
        for _ in range(6 if n < 1000000 else 2):
            f_code = 1.023e6 + rng.uniform(-30.0, 30.0)
            x0 = rng.uniform(0.0, 1023.0)
            cases.append((x0, f_code, delt, n))
        # start right below the wrap, at zero, and in the tiny binades just above zero
        cases.append((np.nextafter(1023.0, 0.0), 1.023e6 + 3.0, delt, n))
        cases.append((0.0, 1.023e6 - 28.5, delt, n))
        cases.append((2.0 ** -30, 1.023e6 + 28.5, delt, n))
    return cases


@pytest.mark.parametrize("x0,f_code,delt,n", _cases())
def test_jump_chain_equals_sequential_replay(x0, f_code, delt, n):
    d = float(np.float64(f_code) * np.float64(delt))       # the reference's rounded product
    for every in (128, 512, 4096):
        xo, wo = oracle_lib.code_phase_checkpoints(x0, f_code, delt, n, every)
        xj, wj = emu_lib.code_chain(x0, d, n, every, replay=False)
        assert np.array_equal(xo.view(np.uint64), xj.view(np.uint64)), every   # bit-exact doubles
        assert np.array_equal(wo, wj)


def test_tie_increments_are_handled():
    # d with a significand that is an exact half-ulp tie in the binades above it
    rng = np.random.default_rng(3)
    for shift in range(1, 12):
        mant = (int(rng.integers(1 << 51, 1 << 52)) >> shift << shift) | (1 << (shift - 1)) | (1 << 52)
        d = float(np.ldexp(np.float64(mant), -54))          # in [0.25, 0.5)
        for x0 in (0.3, 1.5, 700.0):
            n = 20000
            # oracle takes f_code and delt; use delt = 1 so f_code*delt == d exactly
            xo, wo = oracle_lib.code_phase_checkpoints(x0, d, 1.0, n, 256)
            xj, wj = emu_lib.code_chain(x0, d, n, 256)
            assert np.array_equal(xo.view(np.uint64), xj.view(np.uint64)), (shift, x0)
            assert np.array_equal(wo, wj)


def test_replay_variant_matches_too():
    x0, f_code, delt, n = 511.25, 1.023e6 + 1.5, 1 / 2.6e6, 260000
    d = float(np.float64(f_code) * np.float64(delt))
    xo, wo = oracle_lib.code_phase_checkpoints(x0, f_code, delt, n, 512)
    xr, wr = emu_lib.code_chain(x0, d, n, 512, replay=True)
    assert np.array_equal(xo.view(np.uint64), xr.view(np.uint64)) and np.array_equal(wo, wr)


def _carrier_cases():
    rng = np.random.default_rng(99)
    cases = []
    for fs, n in ((1.0e6, 100000), (2.6e6, 260000), (20.0e6, 2000000)):
        for f_carr in (3712.5, -3712.5, 41234.25, -40316.0, 12.75, -0.5, 1500.0 * rng.uniform(0.1, 2.0)):
            cases.append((rng.uniform(0.0, 1.0), f_carr, 1.0 / fs, n))
        cases.append((0.0, 2500.0, 1.0 / fs, n))
        cases.append((np.nextafter(1.0, 0.0), -2500.0, 1.0 / fs, n))
        cases.append((2.0 ** -40, -3000.0, 1.0 / fs, n))
    return cases


@pytest.mark.parametrize("x0,f_carr,delt,n", _carrier_cases())
def test_carrier_chain_equals_sequential_replay(x0, f_carr, delt, n):
    """FLOAT_CARR_PHASE: x += f_carr*delt with wrap into [0,1) both ways (gpssim.c:2245-2250); the device
    walks 512*x with step 512*RN(f_carr*delt) and modulus 512 - an exact power-of-two rescaling."""
    d = float(np.float64(f_carr) * np.float64(delt))
    for every in (200, 520):
        xo, end_o = oracle_lib.carrier_phase_checkpoints(x0, f_carr, delt, n, every)
        last = ((n - 1) // every) * every
        xj, wj, end_j = emu_lib.phase_chain(512.0 * x0, 512.0 * d, 512.0, last, every)
        assert np.array_equal((xo * 512.0).view(np.uint64), xj[:xo.size].view(np.uint64)), every
    # whole-epoch advance (what the host needs for the next epoch's row)
    _, end_o = oracle_lib.carrier_phase_checkpoints(x0, f_carr, delt, n, n)
    _, _, end_j = emu_lib.phase_chain(512.0 * x0, 512.0 * d, 512.0, n, 1 << 30)
    assert np.float64(end_o * 512.0).view(np.uint64) == np.float64(end_j).view(np.uint64)


def test_chain_walk_property_random_steps_and_moduli():
    """Property (hypothesis): for random start, step (either sign, 2^-8 .. 2 chips per sample) and sample
    count, the jump walk equals the plain sequential recurrence bit for bit - checkpoints and final value."""
    from hypothesis import given, settings, strategies as st

    def replay(x, d, mod, n, every):
        out = []
        for i in range(n + 1):
            if i % every == 0:
                out.append(x)
            if i == n:
                break
            x = np.float64(x) + np.float64(d)
            if x >= mod:
                x = x - np.float64(mod)
            elif x < 0.0:
                x = x + np.float64(mod)
        return np.array(out, dtype=np.float64), float(x)

    @settings(max_examples=250, deadline=None)
    @given(frac=st.floats(0.0, 1.0, exclude_max=True), mant=st.floats(1.0, 2.0, exclude_max=True),
           expo=st.integers(-8, 0), neg=st.booleans(), mod=st.sampled_from([1023.0, 512.0]),
           n=st.integers(1, 6000), every=st.sampled_from([8, 200, 520, 4096]))
    def check(frac, mant, expo, neg, mod, n, every):
        d = float(np.ldexp(mant, expo)) * (-1.0 if neg else 1.0)
        x0 = float(np.float64(frac) * np.float64(mod))
        want, end_w = replay(x0, d, mod, n, every)
        got, _, end_g = emu_lib.phase_chain(x0, d, mod, n, every)
        assert np.array_equal(want.view(np.uint64), got[:want.size].view(np.uint64))
        assert np.float64(end_w).view(np.uint64) == np.float64(end_g).view(np.uint64)

    check()


def test_falling_chain_never_jumps_onto_the_binade_edge():
    """A falling chain that would land exactly on 2^e must take that step as a real addition: below 2^e the
    grid is twice as fine, so the reference's sum can round to 2^e - ulp/2 where a jump on the coarse grid
    says 2^e.  x0 = 256 + 3*q*u, d = -(q*u + 0.375*u): three steps, the third one lands on 256."""
    u = 2.0 ** -44                                   # ulp in [256, 512)
    for q in (1 << 20, (1 << 30) + 12345, 3):
        d = -(q * u + 0.375 * u)
        x0 = 256.0 + 3 * q * u
        want = np.float64(x0)
        for _ in range(3):
            want = want + np.float64(d)
        assert want < 256.0                          # the reference ends half a (coarse) ulp below the edge
        got, _, end = emu_lib.phase_chain(x0, d, 512.0, 3, 1)
        assert np.float64(end).view(np.uint64) == np.float64(want).view(np.uint64)
        assert np.float64(got[3]).view(np.uint64) == np.float64(want).view(np.uint64)
