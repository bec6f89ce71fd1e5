"""CPU: the oracle's min_distances / med_distances restatement (oracle.c orc_class_dists) against
the golden values the unmodified reference produced (tests/golden/make_golden_classdist.py) and,
when oracle/_ref is present, against the compiled reference on more seeds."""
import os
import sys

import numpy as np
import pytest

from conftest import assert_bits_equal
from oracle.pyoracle import Reference

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from make_golden_classdist import CASES, make_case  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "classdist.npz"))


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("median", [0, 1])
def test_oracle_class_dists_golden(oracle, name, median):
    codes, labels, mask = make_case(name)
    cls, noe, dists = oracle.class_dists(codes, labels, bool(median), mask)
    assert np.array_equal(cls, GOLD["%s_m%d_class" % (name, median)])
    assert np.array_equal(noe, GOLD["%s_m%d_noe" % (name, median)])
    assert_bits_equal(dists, GOLD["%s_m%d_dists" % (name, median)], name)


@pytest.mark.skipif(not Reference.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_class_dists_vs_compiled_reference(oracle):
    ref = Reference()
    rng = np.random.default_rng(5)
    for M, D, ncls in [(64, 3, 2), (333, 17, 9), (90, 40, 90)]:
        codes = rng.random((M, D), dtype=np.float32)
        labels = (rng.integers(0, ncls, M) + 1).astype(np.int32)
        for median in (False, True):
            a = oracle.class_dists(codes, labels, median)
            b = ref.class_dists(codes, labels, median)
            for x, y in zip(a, b):
                assert_bits_equal(x, y)


def test_hitlist_order_matches_oracle(oracle):
    """engine.hitlist_order is host logic (no GPU): same class order as the oracle's add_hit"""
    from som_lvq_pak_b200 import engine
    rng = np.random.default_rng(2)
    for n, ncls in [(50, 4), (400, 13), (7, 7)]:
        labels = (rng.integers(0, ncls, n) + 1).astype(np.int32)
        cls, noe, _ = oracle.class_dists(np.zeros((n, 2), np.float32), labels, True)
        lab, freq = engine.hitlist_order(labels)
        assert lab == list(cls) and freq == list(noe)
