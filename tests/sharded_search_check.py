#!/usr/bin/env python
"""Process-per-GPU mode of the sharded search against the oracle (run under torchrun; used by
tests/test_multi_gpu.py::test_process_per_gpu_mode_vs_oracle).  Every rank binds its GPU, the library's NCCL
communicator is made from an id that torch.distributed carries (bmu_comm_unique_id -> bmu_comm_init_rank), rank 0's
codebook is replicated with bmu_comm_broadcast_dev, every rank searches its contiguous shard (bmu_search_dev +
bmu_search_stats_dev) and the statistics are combined by bmu_comm_allreduce_stats_dev.  Rank 0 gathers the per-row
results in shard order and compares everything with the oracle's unsharded search."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)                                # test infrastructure: may use the oracle
import som_lvq_pak_b200 as bmu  # noqa: E402
from som_lvq_pak_b200 import _lib  # noqa: E402
from som_lvq_pak_b200 import distributed as D  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    bmu.init(local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = D.comm_init()
    lib = _lib.load()
    M, Dm, N, k = 1500, 64, 50_000, 1
    rng = np.random.default_rng(77)
    codes_h = rng.random((M, Dm), dtype=np.float32)
    data_h = rng.random((N, Dm), dtype=np.float32)
    stream = torch.cuda.current_stream().cuda_stream
    codes = torch.from_numpy(codes_h).to(dev) if rank == 0 else torch.zeros((M, Dm), device=dev)
    _lib.check(lib.bmu_comm_broadcast_dev(codes.data_ptr(), M * Dm * 4, 0, stream))
    torch.cuda.synchronize()
    assert np.array_equal(codes.cpu().numpy(), codes_h)                 # every rank holds rank 0's codebook
    cb = lib.bmu_codebook_create_dev(codes.data_ptr(), M, Dm)
    lo, hi = D.shard_bounds(N, rank, world)
    data = torch.from_numpy(data_h[lo:hi]).to(dev)
    ss = D.ShardedSearch(cb, M, hi - lo, k, dev)
    for _ in range(2):                                                  # twice: the buffers are zeroed per step
        ss.step(data.data_ptr())
    qsum, nfound, hist = ss.totals()
    parts = [None] * world
    dist.all_gather_object(parts, (lo, ss.idx.cpu().numpy(), ss.diff.cpu().numpy()))
    ok = True
    if rank == 0:
        from oracle.pyoracle import Oracle
        eidx, ediff, _ = Oracle().search(codes_h, data_h, k)
        parts.sort(key=lambda t: t[0])
        idx = np.concatenate([p[1] for p in parts])
        diff = np.concatenate([p[2] for p in parts])
        esum = np.sqrt(ediff[:, 0].astype(np.float64)).sum()
        ok = (np.array_equal(idx, eidx) and np.array_equal(diff.view(np.int32), ediff.view(np.int32)) and nfound == N and
              np.array_equal(hist, np.bincount(eidx[:, 0], minlength=M)) and abs(qsum - esum) <= N * np.spacing(esum))
        print("process-per-GPU sharded search over %d ranks identical to the oracle: %s" % (world, ok))
    lib.bmu_codebook_destroy(cb)
    lib.bmu_comm_destroy()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
