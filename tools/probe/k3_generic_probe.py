import sys, time, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np
from som_lvq_pak_b200 import engine
rng = np.random.default_rng(0)
engine.init(0)
for (M, D, N, L, what) in [(2000, 64, 5000, 20000, "lvq1"), (16384, 64, 5000, 20000, "lvq1"), (16384, 64, 5000, 20000, "lvq3"), (4096, 16, 5000, 20000, "som_bubble"), (65536, 32, 5000, 20000, "som_gauss")]:
    codes = rng.random((M, D), dtype=np.float32); data = rng.random((N, D), dtype=np.float32)
    cl = rng.integers(1, 6, M).astype(np.int32); dl = rng.integers(1, 6, N).astype(np.int32)
    for rep in range(3):
        t0 = time.perf_counter()
        if what == "lvq1": engine.lvq_training(1, codes, cl, data, dl, L, 0.05, 1, 0.3, 0.1)
        elif what == "lvq3": engine.lvq_training(3, codes, cl, data, dl, L, 0.05, 1, 0.3, 0.1)
        elif what == "som_bubble": engine.som_training(codes, data, 64, M // 64, 3, 1, L, 0.05, 3.0, 1)
        else: engine.som_training(codes, data, 256, M // 256, 3, 2, L, 0.05, 30.0, 1)
        dt = time.perf_counter() - t0
    print("%-10s M=%d D=%d steps=%d: %.1f ms  %.2f us/step" % (what, M, D, L, dt * 1e3, dt * 1e6 / L))
