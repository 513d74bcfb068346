"""The seeded synthetic table bench.py uses: valid rows, and the device algorithms (run on the
CPU) agree with the oracle on it."""
import numpy as np

import emu_lib
import oracle_lib
import gps_sdr_sim_b200 as gs


def test_synthetic_table_is_inside_the_measured_envelopes():
    t = gs.synthetic_table(40, 260000, 13, gs.SC08)
    t.validate()
    act = t.prn > 0
    assert t.max_active() == 13 and (act.sum(axis=1) == 13).all()
    assert ((t.code_phase[act] >= 0) & (t.code_phase[act] < 1023)).all()
    assert ((t.icode[act] >= 0) & (t.icode[act] < 20)).all()
    assert ((t.gain[act] >= 36) & (t.gain[act] <= 127)).all()
    assert np.abs(t.f_code[act] - 1.023e6).max() < 30.0
    assert np.abs(t.carr_phasestep[act]).max() < 570000
    # the carrier phase column is the exact uint32 prefix sum of N*step (gpssim.c:2252 summed)
    N = t.samples_per_epoch
    ph = t.carr_phase.astype(np.int64)
    st = t.carr_phasestep.astype(np.int64)
    assert np.array_equal((ph[:-1] + N * st[:-1]) % 2**32, ph[1:])
    # deterministic
    t2 = gs.synthetic_table(40, 260000, 13, gs.SC08)
    assert all(np.array_equal(t.cols[k], t2.cols[k]) for k in t.cols)


def test_device_algorithms_equal_oracle_on_synthetic_rows():
    for fmt in (gs.SC16, gs.SC08, gs.SC01):
        t = gs.synthetic_table(3, 260000, 13, fmt, seed=5)
        want = oracle_lib.generate(t)
        assert np.array_equal(emu_lib.generate(t, 512, emu_lib.TUNED32), want)
        assert np.array_equal(emu_lib.generate(t, 256, emu_lib.TUNED32, force_wrap=True), want)
    t = gs.synthetic_table(2, 100000, 16, gs.SC16, seed=9)      # 1 MS/s: more than one chip per sample
    assert np.array_equal(emu_lib.generate(t, 512, emu_lib.TUNED16), oracle_lib.generate(t))
