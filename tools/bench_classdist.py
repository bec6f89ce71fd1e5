#!/usr/bin/env python
"""K5 timing: bmu_class_nearest (host pointers in, host pointers out) on an LVQ codebook of M vectors,
next to the oracle's restatement of the same pair loops on one host core.
    python tools/bench_classdist.py [M] [D] [classes]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from som_lvq_pak_b200 import engine  # noqa: E402


def main():
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    D = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    ncls = int(sys.argv[3]) if len(sys.argv) > 3 else 32
    rng = np.random.default_rng(1)
    codes = rng.random((M, D), dtype=np.float32)
    labels = (rng.integers(0, ncls, M) + 1).astype(np.int32)
    engine.class_nearest(codes[:256], labels[:256])
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        d, f = engine.class_nearest(codes, labels)
        ts.append(time.perf_counter() - t0)
    out = {"M": M, "D": D, "classes": ncls, "gpu_ms_e2e": round(1e3 * min(ts), 3),
           "pair_elements_per_s": M * (M - 1) / 2 * D / min(ts)}
    if os.environ.get("WITH_ORACLE", "1") == "1":
        from oracle.pyoracle import Oracle           # checker + CPU timing only
        o = Oracle()
        m = min(M, 4000)
        t0 = time.perf_counter()
        _, _, _, near, found = o.class_dists(codes[:m], labels[:m], True, None, per_entry=True)
        t = time.perf_counter() - t0
        d2, f2 = engine.class_nearest(codes[:m], labels[:m])
        out["oracle_ms_at_M%d" % m] = round(1e3 * t, 1)
        out["oracle_same_class_pair_elements_per_s"] = float(
            sum(c * (c - 1) / 2 for c in np.bincount(labels[:m])) * D / t)
        out["bit_exact"] = bool(np.array_equal(d2.view(np.uint32), near.view(np.uint32)) and np.array_equal(f2, found))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
