"""CPU: the oracle's restatement of Sammon's mapping (oracle.c orc_remove_identicals / orc_sammon,
sammon.c:83-262) against the positions the unmodified reference produced
(tests/golden/make_golden_sammon.py) and, when oracle/_ref is present, the compiled reference."""
import os
import sys

import numpy as np
import pytest

from conftest import assert_bits_equal
from oracle.pyoracle import Reference

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from make_golden_sammon import CASES, make_case  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "sammon.npz"))


def oracle_flow(oracle, codes, mask, length, seed):
    from som_lvq_pak_b200 import engine                 # sammon_init is host arithmetic (no GPU)
    keep = oracle.remove_identicals(codes, mask)
    x0, y0 = engine.sammon_init(len(keep), seed)
    return oracle.sammon(codes[keep], length, x0, y0, None if mask is None else mask[keep])


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_sammon_golden(oracle, name):
    codes, mask, length, seed = make_case(name)
    x, y = oracle_flow(oracle, codes, mask, length, seed)
    assert_bits_equal(x, GOLD[name + "_x"], name + " x")
    assert_bits_equal(y, GOLD[name + "_y"], name + " y")


@pytest.mark.skipif(not Reference.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_sammon_vs_compiled_reference(oracle):
    ref = Reference()
    rng = np.random.default_rng(8)
    for M, D, length, seed in [(17, 2, 30, 1), (64, 12, 12, 77), (150, 4, 8, 31000)]:
        codes = rng.random((M, D), dtype=np.float32)
        codes[M // 2] = codes[1]
        x, y = oracle_flow(oracle, codes, None, length, seed)
        rx, ry = ref.sammon(codes, length, seed)
        assert_bits_equal(x, rx)
        assert_bits_equal(y, ry)
