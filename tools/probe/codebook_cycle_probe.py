"""repeated codebook create / search / destroy cycles: where do slow calls spend their time"""
import sys, time, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np
import som_lvq_pak_b200 as bmu
rng = np.random.default_rng(0)
bmu.init(0)
for M, D, N in ((200, 20, 2000), (10000, 64, 20000)):
    codes = rng.random((M, D), dtype=np.float32); data = rng.random((N, D), dtype=np.float32)
    slow = 0
    tot = [0.0, 0.0, 0.0]
    for rep in range(80):
        t0 = time.perf_counter()
        cb = bmu.Codebook(codes); t1 = time.perf_counter()
        cb.find_winners(data, 1); t2 = time.perf_counter()
        cb.close(); t3 = time.perf_counter()
        d = [(t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3]
        if rep >= 2:
            tot = [a + b for a, b in zip(tot, d)]
        if rep < 2 or max(d) > 20:
            slow += rep >= 2
            print("M=%d rep %2d: create %.1f search %.1f destroy %.1f ms" % (M, rep, *d))
    print("M=%d: mean create %.2f search %.2f destroy %.2f ms, %d slow cycles of 78" % (M, *(x / 78 for x in tot), slow))
