#!/usr/bin/env python
"""vfind over several GPUs: one process per GPU, every rank runs its share of the
trials, the best map wins.   torchrun --nproc-per-node N tools/vfind_multi.py [trials]
Prints the winning trial and checks that it is the one a single process finds."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import som_lvq_pak_b200 as bmu  # noqa: E402
from som_lvq_pak_b200 import distributed as Dm  # noqa: E402

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 8
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
bmu.init(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rng = np.random.default_rng(1)
data = rng.random((4000, 16), dtype=np.float32)
args = (data, data, 12, 8, bmu.TOPOL_HEXA, bmu.NEIGH_BUBBLE, trials, 1000, 0.05, 6.0, 4000, 0.02, 2.0)
codes, q, n = Dm.vfind(*args)
if dist.get_rank() == 0:
    solo = None
    for t in range(trials, 0, -1):
        c = bmu.randinit_codes(data, 12, 8, t)
        c = bmu.som_training(c, data, 12, 8, 3, 1, 1000, 0.05, 6.0)
        c = bmu.som_training(c, data, 12, 8, 3, 1, 4000, 0.02, 2.0)
        qq = bmu.find_qerror(c, data)
        if solo is None or qq < solo[0]:
            solo = (qq, t, c)
    ok = solo[1] == n and np.array_equal(solo[2].view(np.int32), codes.view(np.int32))
    print("vfind over %d GPUs: best trial %d, qerror/sample %.6f, identical to the single-process search: %s"
          % (dist.get_world_size(), n, q / len(data), ok))
    assert ok
dist.destroy_process_group()
