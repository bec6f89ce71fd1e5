"""where does an occasional slow training call spend its time: kernel (CUDA events) or host calls"""
import sys, time, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np
import som_lvq_pak_b200 as bmu
rng = np.random.default_rng(0)
bmu.init(0)
for M, D, N, L in ((200, 20, 1962, 50000), (16384, 64, 5000, 20000)):
    codes = rng.random((M, D), dtype=np.float32); data = rng.random((N, D), dtype=np.float32)
    cl = rng.integers(1, 6, M).astype(np.int32); dl = rng.integers(1, 6, N).astype(np.int32)
    worst = None
    for rep in range(60):
        t = [time.perf_counter()]
        s, ta = bmu.lvq_schedule(0, L, L, 0.05, 1, N, None); t.append(time.perf_counter())
        tr = bmu.Trainer(codes, data, None); t.append(time.perf_counter())
        tr.set_lvq(1, cl, dl, 0.5, 0.1, 0.05, None); t.append(time.perf_counter())
        tr.steps(s, ta, None); t.append(time.perf_counter())
        out = tr.codes(); t.append(time.perf_counter())
        ms = tr.last_ms()
        tr.close(); t.append(time.perf_counter())
        d = [(b - a) * 1e3 for a, b in zip(t[:-1], t[1:])]
        line = "M=%d rep %2d: schedule %.1f create %.1f set %.1f steps %.1f (kernel %.1f) codes %.1f close %.1f ms" % (
            M, rep, d[0], d[1], d[2], d[3], ms, d[4], d[5])
        if rep < 2 or sum(d) > 1.5 * ms + 5:
            print(line)
    print("M=%d done" % M)
