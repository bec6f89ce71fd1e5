#!/usr/bin/env python
"""Wall time of the two demo recipes (BASELINE.json configs[0] and [1]; reference Makefile:195-212)
through the reference's own binaries (oracle/_ref/bin, CPU) and through bmu_pak (B200), program by
program.  At these sizes (96 / 200 code vectors) every bmu_pak call is dominated by creating the CUDA
context, so this is a statement of where the GPU path does NOT pay, next to bench.py's C3-C5 numbers.
    python tools/bench_demo_cli.py"""
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "bin")
PAK = os.path.join(ROOT, "som_lvq_pak_b200", "host", "bmu_pak")

SOM = [("randinit", "-din ex.dat -cout ex.cod -xdim 12 -ydim 8 -topol hexa -neigh bubble -rand 123"),
       ("vsom", "-din ex.dat -cin ex.cod -cout ex.cod -rlen 1000 -alpha 0.05 -radius 10"),
       ("vsom", "-din ex.dat -cin ex.cod -cout ex.cod -rlen 10000 -alpha 0.02 -radius 3"),
       ("qerror", "-din ex.dat -cin ex.cod"),
       ("vcal", "-din ex_fts.dat -cin ex.cod -cout ex.cod"),
       ("visual", "-din ex_ndy.dat -cin ex.cod -dout ex.nvs")]
LVQ = [("eveninit", "-din ex1.dat -cout ex1e.cod -noc 200"),
       ("balance", "-din ex1.dat -cin ex1e.cod -cout ex1b.cod"),
       ("olvq1", "-din ex1.dat -cin ex1b.cod -cout ex1o.cod -rlen 5000"),
       ("lvq1", "-din ex1.dat -cin ex1o.cod -cout ex1l.cod -alpha 0.05 -rlen 50000"),
       ("accuracy", "-din ex2.dat -cin ex1l.cod")]


def run_chain(chain, ours, td):
    out = []
    for prog, args in chain:
        cmd = ([PAK, prog] if ours else [os.path.join(REF, prog)]) + args.split()
        t0 = time.perf_counter()
        p = subprocess.run(cmd, cwd=td, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        dt = time.perf_counter() - t0
        if p.returncode:
            sys.exit("%s failed: %s" % (cmd, p.stderr.decode()[-300:]))
        out.append((prog, round(dt, 3)))
    return out


def main():
    demo = np.load(os.path.join(ROOT, "tests", "golden", "demo.npz"))
    res = {}
    for name, chain in (("som_demo", SOM), ("lvq_demo", LVQ)):
        for ours in (False, True):
            if not ours and not os.path.isdir(REF):
                continue
            td = tempfile.mkdtemp()
            for f in ("ex.dat", "ex_fts.dat", "ex_ndy.dat", "ex1.dat", "ex2.dat"):
                open(os.path.join(td, f), "w").write(str(demo["in_" + f]))
            if ours:
                run_chain(chain[:1], True, td)                  # first CUDA start of the box is not the program's cost
            steps = run_chain(chain, ours, td)
            res["%s_%s" % (name, "bmu_pak_b200" if ours else "reference_cpu")] = {
                "total_s": round(sum(t for _, t in steps), 3), "steps": steps}
            if ours:                                            # the same recipe as ONE process
                recipe = "".join("%s %s\n" % (p, a) for p, a in chain)
                t0 = time.perf_counter()
                subprocess.run([PAK, "batch"], input=recipe.encode(), cwd=td, stdout=subprocess.PIPE,
                               stderr=subprocess.PIPE, check=True)
                res["%s_bmu_pak_b200_batch" % name] = {"total_s": round(time.perf_counter() - t0, 3)}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
