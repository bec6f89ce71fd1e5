"""where does the occasional slow training call spend its time: kernel (CUDA events) or host calls"""
import sys, time, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np
import som_lvq_pak_b200 as bmu
rng = np.random.default_rng(0)
bmu.init(0)
M, D, N, L = 200, 20, 1962, 50000
codes = rng.random((M, D), dtype=np.float32); data = rng.random((N, D), dtype=np.float32)
cl = rng.integers(1, 6, M).astype(np.int32); dl = rng.integers(1, 6, N).astype(np.int32)
s, ta = bmu.lvq_schedule(0, L, L, 0.05, 1, N, None)
for rep in range(25):
    t0 = time.perf_counter()
    tr = bmu.Trainer(codes, data, None)
    t1 = time.perf_counter()
    tr.set_lvq(1, cl, dl, 0.5, 0.1, 0.05, None)
    t2 = time.perf_counter()
    tr.steps(s, ta, None)
    t3 = time.perf_counter()
    out = tr.codes()
    t4 = time.perf_counter()
    ms = tr.last_ms()
    tr.close()
    t5 = time.perf_counter()
    print("rep %2d: create %.1f set %.1f steps %.1f (kernel %.1f) codes %.1f close %.1f ms" % (
        rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, ms, (t4 - t3) * 1e3, (t5 - t4) * 1e3))
print("-- through engine.lvq_training (schedule built per call)")
for rep in range(25):
    t0 = time.perf_counter()
    s2, ta2 = bmu.lvq_schedule(0, L, L, 0.05, 1, N, None)
    t1 = time.perf_counter()
    out = bmu.lvq_training(1, codes, cl, data, dl, L, 0.05, 1, 0.3, 0.1)
    t2 = time.perf_counter()
    print("rep %2d: schedule %.1f  lvq_training %.1f ms" % (rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3))
