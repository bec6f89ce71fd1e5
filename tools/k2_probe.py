"""quick device-resident search timing probe (GPU box):
   python tools/k2_probe.py [rows] [M] [D] [k] [path]      path: 0 auto, 1 exact, 2 filter"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import som_lvq_pak_b200 as b
from bench import synth_rows_torch

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
M = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
D = int(sys.argv[3]) if len(sys.argv) > 3 else 64
k = int(sys.argv[4]) if len(sys.argv) > 4 else 1
path = int(sys.argv[5]) if len(sys.argv) > 5 else 2
b.init(0)
dev = torch.device("cuda:0")
codes = synth_rows_torch(2, 0, M, D, dev)
data = synth_rows_torch(1, 0, rows, D, dev)
idx = torch.empty((rows, k), dtype=torch.int32, device=dev)
diff = torch.empty((rows, k), dtype=torch.float32, device=dev)
nf = torch.empty(rows, dtype=torch.int32, device=dev)
b.set_search_path(path)
cb = b.Codebook(codes.cpu().numpy())
iters = int(os.environ.get('PROBE_ITERS', '4'))
for it in range(iters):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    cb.search_dev(data.data_ptr(), rows, k, idx.data_ptr(), diff.data_ptr(), nf.data_ptr())
    e1.record()
    torch.cuda.synchronize()
    if iters > 4:
        print('  it %d: %.3f ms  gemm %.3f' % (it, e0.elapsed_time(e1), b.last_search_kernel_ms()['k2_gemm']))
ms = e0.elapsed_time(e1)
print("rows %d M %d D %d k %d path %d: %.3f ms  %.1f M searches/s" % (rows, M, D, k, path, ms, rows / ms / 1e3))
print("  kernels:", {n: round(v, 3) for n, v in b.last_search_kernel_ms().items() if v})
print("  breakdown:", b.last_search_breakdown())
