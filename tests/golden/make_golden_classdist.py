#!/usr/bin/env python
"""Golden per-class values of min_distances / med_distances (lvq_rout.c:280-492) computed by the
UNMODIFIED reference (oracle/_ref/libref_driver.so -> ref_class_dists) on seeded codebooks.
Run in the build container (needs /root/reference for `make -C oracle ref`):
    python tests/golden/make_golden_classdist.py    ->  tests/golden/classdist.npz"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

CASES = {  # name: (M, D, classes, quantised, masked)
    "lvqdemo": (200, 20, 10, False, False),
    "ties": (300, 8, 4, True, False),
    "masked": (150, 6, 5, True, True),
    "oneclass": (130, 7, 1, False, False),
    "singles": (40, 5, 37, False, False),
    "wide": (257, 100, 3, False, False),
}


def make_case(name):
    M, D, ncls, quant, masked = CASES[name]
    rng = np.random.default_rng(sum(map(ord, name)))
    if quant:
        codes = (rng.integers(0, 3, (M, D)) / 2).astype(np.float32)
    else:
        codes = rng.random((M, D), dtype=np.float32)
    labels = (rng.integers(0, ncls, M) + 1).astype(np.int32)        # label 0 is the empty label
    mask = None
    if masked:
        mask = (rng.random((M, D)) < 0.35).astype(np.uint8)
        mask[5] = 1                                                  # entries that share no component:
        mask[9, : D // 2] = 1                                        # vector_dist_euc returns -1
        mask[9, D // 2:] = 0
        mask[20, : D // 2] = 0
        mask[20, D // 2:] = 1
        labels[[5, 9, 20]] = labels[9]
        codes[mask != 0] = 0.0
    return codes, labels, mask


def main():
    from oracle.pyoracle import Reference
    if not Reference.available():
        sys.exit("build the reference first: make -C oracle ref (needs /root/reference)")
    ref = Reference()
    out = {}
    for name in CASES:
        codes, labels, mask = make_case(name)
        for median in (0, 1):
            cls, noe, dists = ref.class_dists(codes, labels, bool(median), mask)
            out["%s_m%d_class" % (name, median)] = cls
            out["%s_m%d_noe" % (name, median)] = noe
            out["%s_m%d_dists" % (name, median)] = dists
    np.savez_compressed(os.path.join(HERE, "classdist.npz"), **out)
    print("wrote classdist.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
