import sys, time, ctypes, os
sys.path.insert(0, os.getcwd())
import numpy as np
import som_lvq_pak_b200 as bmu
from som_lvq_pak_b200 import _lib
from bench import synth_numpy
bmu.init(0)
lib = _lib.load()
N, D, M = 10_000_000, 64, 10_000
codes = synth_numpy(2, 0, M * D).reshape(M, D)
data = np.empty((N, D), np.float32)
for r in range(0, N, 1 << 20):
    n = min(1 << 20, N - r)
    data[r:r + n] = synth_numpy(1, r * D, n * D).reshape(n, D)
cb = lib.bmu_codebook_create(codes.ctypes.data, M, D)
idx = np.empty((N, 1), np.int32); diff = np.empty((N, 1), np.float32); nf = np.empty(N, np.int32)
for label, thr in (("default", 0), ("threads 6", 6), ("threads 12", 12), ("default", 0)):
    lib.bmu_shutdown(); bmu.init(0); cb = lib.bmu_codebook_create(codes.ctypes.data, M, D)
    lib.bmu_set_copy_threads(thr)
    ts = []
    for it in range(5):
        t0 = time.perf_counter()
        _lib.check(lib.bmu_search(cb, data.ctypes.data, None, N, 1, idx.ctypes.data, diff.ctypes.data, nf.ctypes.data))
        ts.append(time.perf_counter() - t0)
    print(label, "pieces", os.environ.get("SOMLVQ_RING_PIECE_MB", "8"), ["%.1f" % (N / t / 1e6) for t in ts], flush=True)
