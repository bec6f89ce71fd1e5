#!/usr/bin/env python
"""One-off sanity of 64-bit indexing: N rows x 64 (> 4 GiB of input, > 2^31 elements) through the filter
path and the exact path, all rows compared bit for bit.   python tools/big_n_check.py [rows]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import som_lvq_pak_b200 as b
from bench import synth_rows_torch

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 60_000_000
M, D = 10000, 64
b.init(0)
dev = torch.device("cuda:0")
codes = synth_rows_torch(2, 0, M, D, dev)
data = torch.empty((rows, D), dtype=torch.float32, device=dev)
step = 10_000_000
for r0 in range(0, rows, step):                       # generated in slices: the generator's temporaries are large
    n = min(step, rows - r0)
    data[r0:r0 + n] = synth_rows_torch(1, r0, n, D, dev)
cb = b.Codebook(codes.cpu().numpy())
res = []
for path in (2, 1):
    idx = torch.empty((rows, 1), dtype=torch.int32, device=dev)
    diff = torch.empty((rows, 1), dtype=torch.float32, device=dev)
    nf = torch.empty(rows, dtype=torch.int32, device=dev)
    b.set_search_path(path)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    cb.search_dev(data.data_ptr(), rows, 1, idx.data_ptr(), diff.data_ptr(), nf.data_ptr())
    e1.record()
    torch.cuda.synchronize()
    print("path %d: %.1f ms, %.1f M searches/s" % (path, e0.elapsed_time(e1), rows / e0.elapsed_time(e1) / 1e3), b.last_search_breakdown())
    res.append((idx, diff, nf))
b.set_search_path(0)
same = bool((res[0][0] == res[1][0]).all()) and bool((res[0][1].view(torch.int32) == res[1][1].view(torch.int32)).all())
print("rows %d: filter path == exact path on every row: %s; all found: %s; last rows idx %s" %
      (rows, same, bool((res[0][2] == 1).all()), res[0][0][-3:, 0].tolist()))
sys.exit(0 if same else 1)
