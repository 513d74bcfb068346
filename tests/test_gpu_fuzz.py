"""Randomised parity (tools/fuzz_parity.py): seeded random tables well outside the envelopes of the
reference's scenarios, every entry point and kernel option, CUDA path vs oracle, bit-exact."""
import os
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [101, 202])
def test_random_tables_match_the_oracle(seed, gpu_required):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import fuzz_parity
    assert fuzz_parity.sweep(60, seed, verbose=False) is None
