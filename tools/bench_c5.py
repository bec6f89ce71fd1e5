#!/usr/bin/env python
"""vsom steps/s on BASELINE.json configs[4]: 256x256 hexa gaussian map, 128-dim, rlen 1e6
(data 100 000 x 128 seed 4, map U[0,1) seed 5, alpha 0.05 linear, radius 100,
sample order of `-rand 3`).  Runs `--steps` of the 1e6-step schedule on one GPU and reports
steps/s from the CUDA events inside the library; optional parity check of a prefix against
the oracle."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import synth_numpy  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20000)
    ap.add_argument("--xdim", type=int, default=256)
    ap.add_argument("--ydim", type=int, default=256)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--neigh", default="gaussian")
    ap.add_argument("--radius", type=float, default=100.0)
    ap.add_argument("--check", type=int, default=0, help="compare the first CHECK steps with the oracle")
    a = ap.parse_args()
    import som_lvq_pak_b200 as bmu
    bmu.init(0)
    N, D, M, length = 100_000, a.dim, a.xdim * a.ydim, 1_000_000
    data = synth_numpy(4, 0, N * D).reshape(N, D)
    codes = synth_numpy(5, 0, M * D).reshape(M, D)
    order = bmu.rand_order(N, 3)
    neigh = bmu.NEIGH_GAUSSIAN if a.neigh == "gaussian" else bmu.NEIGH_BUBBLE
    tr = bmu.Trainer(codes, data)
    tr.set_som(a.xdim, a.ydim, bmu.TOPOL_HEXA, neigh)
    s, ta, trd = bmu.som_schedule(0, a.steps, length, 0.05, a.radius, bmu.ALPHA_LINEAR, N, order)
    tr.steps(s[:64], ta[:64], trd[:64])                       # warm-up launch
    tr2 = bmu.Trainer(codes, data)
    tr2.set_som(a.xdim, a.ydim, bmu.TOPOL_HEXA, neigh)
    t0 = time.perf_counter()
    tr2.steps(s, ta, trd)
    wall = time.perf_counter() - t0
    ms = tr2.last_ms()
    out = {"metric": "vsom steps/s", "value": a.steps / (ms * 1e-3), "unit": "steps/s", "steps": a.steps,
           "kernel_ms": ms, "us_per_step": 1e3 * ms / a.steps, "wall_s_incl_schedule_upload": wall,
           "config": {"map": "%dx%d hexa %s" % (a.xdim, a.ydim, a.neigh), "dim": D, "rlen": length,
                      "alpha": 0.05, "radius": a.radius, "order": "-rand 3"},
           "flop_per_step": 6.0 * M * D, "lane_tops": 6.0 * M * D * a.steps / (ms * 1e-3) / 1e12}
    if a.check:
        from oracle.pyoracle import Oracle
        o = Oracle()
        tr3 = bmu.Trainer(codes, data)
        tr3.set_som(a.xdim, a.ydim, bmu.TOPOL_HEXA, neigh)
        tr3.steps(s[:a.check], ta[:a.check], trd[:a.check])
        got = tr3.codes()
        # the oracle runs the same prefix of the SAME 1e6-step schedule: emulate by giving it the
        # schedule implicitly (length = 1e6, first `check` steps) -> needs a prefix-capable call
        t0 = time.perf_counter()
        exp = o.som_train_prefix(codes, data, a.xdim, a.ydim, 3, 2 if a.neigh == "gaussian" else 1,
                                 length, a.check, 0.05, a.radius, 1, order)
        cpu_s = time.perf_counter() - t0
        nbad = int((got.view(np.int32) != exp.view(np.int32)).sum())
        rel = float(np.max(np.abs(got - exp) / np.maximum(np.abs(exp), 1e-30)))
        out["check"] = {"steps": a.check, "floats_differing": nbad, "max_rel_diff": rel,
                        "oracle_steps_per_s_1core": a.check / cpu_s}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
