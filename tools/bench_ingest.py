#!/usr/bin/env python
"""Ingest throughput of the C host's loader next to the reference's.
CPU only.  The reference has no load-only program: `qerror` against a 1x1 map is timed instead
(its search is one distance per row, negligible next to sscanf), minus nothing -- an upper bound
on its loader's speed.   python tools/bench_ingest.py [rows] [dim]"""
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAK = os.path.join(ROOT, "som_lvq_pak_b200", "host", "bmu_pak")
REF_QERROR = os.path.join(ROOT, "oracle", "_ref", "bin", "qerror")

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 64
td = tempfile.mkdtemp()
src = os.path.join(td, "big.dat")
rng = np.random.default_rng(0)
with open(src, "w") as f:
    f.write("%d\n" % dim)
    for r0 in range(0, rows, 10000):
        block = rng.random((min(10000, rows - r0), dim), dtype=np.float32)
        f.write("\n".join(" ".join("%g" % v for v in row) for row in block) + "\n")
size = os.path.getsize(src)
print("file: %d rows x %d, %.1f MB" % (rows, dim, size / 1e6))
for threads in (1, 2, 4, 8, os.cpu_count()):
    out = subprocess.run([PAK, "pakstat", "-din", src], stdout=subprocess.PIPE, text=True, check=True,
                         env=dict(os.environ, BMU_PAK_THREADS=str(threads))).stdout.split()
    sec = float(out[out.index("load_seconds") + 1])
    print("bmu_pak loader, %2d threads: %.3f s  %.0f MB/s  %.2f M values/s  (sum %s)" % (
        threads, sec, size / sec / 1e6, rows * dim / sec / 1e6, out[out.index("sum") + 1]))
if os.path.exists(REF_QERROR):
    cod = os.path.join(td, "one.cod")
    open(cod, "w").write("%d hexa 1 1 bubble\n%s\n" % (dim, " ".join(["0.5"] * dim)))
    t0 = time.perf_counter()
    subprocess.run([REF_QERROR, "-din", src, "-cin", cod], stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True)
    sec = time.perf_counter() - t0
    print("reference qerror with a 1x1 map (load + 1 distance per row): %.3f s  %.0f MB/s" % (sec, size / sec / 1e6))
