#!/usr/bin/env python
"""K6 timing: bmu_sammon (host pointers in and out) for `sweeps` sweeps over M code vectors, next to
the oracle's restatement of sammon_iterate on one host core at a smaller M (cost is O(M^2) per sweep).
    python tools/bench_sammon.py [M] [D] [sweeps]"""
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
    sweeps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    rng = np.random.default_rng(1)
    codes = rng.random((M, D), dtype=np.float32)
    x0, y0 = engine.sammon_init(M, 5)
    engine.sammon(codes[:256], 2, x0[:256], y0[:256])
    def run(n):
        t0 = time.perf_counter()
        r = engine.sammon(codes, n, x0, y0)
        return time.perf_counter() - t0, r

    t_setup = min(run(0)[0] for _ in range(3))                   # distance matrix + copies only
    t1 = min(run(sweeps)[0] for _ in range(2))
    t5 = min(run(5 * sweeps)[0] for _ in range(2))
    per = (t5 - t1) / (4 * sweeps)                               # setup and launch ramp cancel
    out = {"M": M, "D": D, "sweeps": sweeps, "gpu_setup_ms": round(1e3 * t_setup, 2),
           "gpu_ms_per_sweep": round(1e3 * per, 3), "pairs_per_s": M * (M - 1) / per}
    if os.environ.get("WITH_ORACLE", "1") == "1":
        from oracle.pyoracle import Oracle                       # checker + CPU timing only
        o = Oracle()
        m = min(M, 1500)
        t0 = time.perf_counter()
        ox, oy = o.sammon(codes[:m], 3, x0[:m], y0[:m])
        t3 = time.perf_counter() - t0
        t0 = time.perf_counter()
        o.sammon(codes[:m], 0, x0[:m], y0[:m])
        tz = time.perf_counter() - t0
        gx, gy = engine.sammon(codes[:m], 3, x0[:m], y0[:m])
        out["oracle_ms_per_sweep_at_M%d" % m] = round(1e3 * (t3 - tz) / 3, 1)
        out["oracle_pairs_per_s"] = m * (m - 1) / ((t3 - tz) / 3)
        out["bit_exact"] = bool(np.array_equal(gx.view(np.uint32), ox.view(np.uint32)) and
                                np.array_equal(gy.view(np.uint32), oy.view(np.uint32)))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
