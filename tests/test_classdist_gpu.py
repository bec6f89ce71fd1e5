"""GPU parity of K5 (bmu_class_nearest: the pair loops of min_distances / med_distances,
lvq_rout.c:280-492) against the oracle per entry and against the reference's golden class values.
Bar: bit-exact."""
import os
import sys

import numpy as np
import pytest

from conftest import assert_bits_equal

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from make_golden_classdist import CASES, make_case  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "classdist.npz"))


@pytest.mark.parametrize("name", sorted(CASES))
def test_class_distances_golden(engine, oracle, name):
    codes, labels, mask = make_case(name)
    _, _, _, near, found = oracle.class_dists(codes, labels, True, mask, per_entry=True)
    d, f = engine.class_nearest(codes, labels, mask)
    assert_bits_equal(f, found, name + " found")
    assert_bits_equal(d, near, name + " near")
    for median in (0, 1):
        cls, noe, dists = engine.class_distances(codes, labels, bool(median), mask)
        assert np.array_equal(cls, GOLD["%s_m%d_class" % (name, median)])
        assert np.array_equal(noe, GOLD["%s_m%d_noe" % (name, median)])
        assert_bits_equal(dists, GOLD["%s_m%d_dists" % (name, median)], "%s median=%d" % (name, median))


def test_class_nearest_random_and_nonfinite(engine, oracle):
    rng = np.random.default_rng(77)
    for M, D, ncls in [(1000, 64, 3), (777, 33, 50), (65, 1, 2), (2, 4, 1), (1, 4, 1)]:
        codes = rng.random((M, D), dtype=np.float32)
        labels = (rng.integers(0, ncls, M) + 1).astype(np.int32)
        if M > 100:
            codes[3, 0] = np.nan          # NaN distance never lowers dissf (dist < dissf is false)
            codes[8, 1] = np.inf
            codes[11] *= np.float32(1e-25)  # subnormal squares
            codes[12] = codes[11]
            labels[[11, 12]] = labels[11]
        _, _, _, near, found = oracle.class_dists(codes, labels, True, None, per_entry=True)
        d, f = engine.class_nearest(codes, labels)
        assert_bits_equal(f, found)
        assert_bits_equal(d, near)
