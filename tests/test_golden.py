"""Pin the oracle - and the device algorithms run on the CPU - to bytes the reference produced.

tests/golden/*.npz hold per-epoch rows recorded from the reference host and the SHA-256 of
the corresponding epochs of the UNMODIFIED reference's output file (tests/golden/make_golden.py).
"""
import hashlib

import numpy as np
import pytest

import emu_lib
import oracle_lib
from conftest import golden_names, load_golden
from gps_sdr_sim_b200.table import CARRIER_FLOAT


def digests(buf, table):
    eb = table.epoch_bytes
    return [hashlib.sha256(buf[e * eb:(e + 1) * eb].tobytes()).hexdigest() for e in range(table.n_epochs)]


def test_fixtures_present():
    names = golden_names()
    assert len(names) >= 10
    for fmt in ("b16", "b8", "b1"):
        assert any(n.endswith(fmt) for n in names)


@pytest.mark.parametrize("name", golden_names())
def test_oracle_reproduces_reference_bytes(name):
    table, want, head = load_golden(name)
    out = oracle_lib.generate(table)
    assert digests(out, table) == want
    assert np.array_equal(out[:head.size], head)


def _kernel_for(table):
    d_max = float((table.f_code * table.delt).max())
    if table.samples_per_epoch % 8 != 0:
        return emu_lib.GENERIC
    return emu_lib.TUNED32 if d_max <= 0.9999 else emu_lib.TUNED16


@pytest.mark.parametrize("name", golden_names())
def test_device_algorithms_on_cpu_reproduce_reference_bytes(name):
    table, want, _ = load_golden(name)
    assert (table.carrier_mode == CARRIER_FLOAT) == ("float" in name)
    if table.samples_per_epoch > 300000:
        table = table.slice(0, 1)
        want = want[:1]
    out = emu_lib.generate(table, chunk=512, kernel=_kernel_for(table))
    assert digests(out, table) == want


@pytest.mark.parametrize("name", ["static_int_b16", "static_int_b8", "static_int_b1", "satellite_int_b16",
                                  "static_float_b16", "satellite_float_b16", "circle_float_b8"])
@pytest.mark.parametrize("variant", ["wrap_path", "generic", "tuned16", "replay_chain", "chunk128", "chunk2048",
                                     "acc_wide", "acc_wide_wrap", "round1_kernel", "round1_kernel_wrap"])
def test_device_algorithm_variants_agree(name, variant):
    table, want, _ = load_golden(name)
    table, want = table.slice(0, 4), want[:4]
    kw = dict(chunk=512, kernel=emu_lib.TUNED32)
    if variant == "wrap_path":
        kw["force_wrap"] = True
    elif variant == "generic":
        kw["kernel"] = emu_lib.GENERIC
    elif variant == "tuned16":
        kw["kernel"] = emu_lib.TUNED16
    elif variant == "replay_chain":
        kw["chain_replay"] = True
    elif variant == "chunk128":
        kw["chunk"] = 128
    elif variant == "chunk2048":
        kw["chunk"] = 2048
    elif variant.startswith("acc_wide"):
        kw["accum"] = 0
        kw["force_wrap"] = variant.endswith("wrap")
    elif variant.startswith("round1_kernel"):
        kw["accum"] = 3
        kw["force_wrap"] = variant.endswith("wrap")
    out = emu_lib.generate(table, **kw)
    assert digests(out, table) == want


def test_formats_are_consistent_views_of_the_same_samples():
    # gpssim.c:2266-2288: sc8 = (signed char)(sc16 >> 4), bit = sc16 > 0, MSB first
    t16, _, _ = load_golden("static_int_b16")
    t16 = t16.slice(0, 2)
    s16 = oracle_lib.generate(t16).view(np.int16)
    s8 = oracle_lib.generate(t16.with_format(8)).view(np.int8)
    s1 = oracle_lib.generate(t16.with_format(1))
    assert np.array_equal(s8, (s16 >> 4).astype(np.int8))
    assert np.array_equal(np.unpackbits(s1), (s16 > 0).astype(np.uint8))
