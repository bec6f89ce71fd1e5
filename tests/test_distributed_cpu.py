"""CPU, world_size 2 over gloo: the launcher-side logic of the sharded batch search in its
process-per-GPU mode -- the library's shard bounds, the broadcast of the 128-byte communicator id,
per-shard results landing in data order, int64 counts summed exactly.  The per-shard winners come
from the oracle here (no GPU); on the GPU box the same split runs inside libbmu_b200.so
(tests/test_multi_gpu.py) and the id exchange feeds bmu_comm_init_rank (bench.py --gpus N)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def test_shard_bounds_cover_rows():
    from som_lvq_pak_b200.distributed import shard_bounds
    for n in (0, 1, 127, 128, 129, 1000, 10_000_000, 1962, 1_000_000, 40_000):
        for world in (1, 2, 3, 4, 8):
            b = [shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            for (l0, h0), (l1, h1) in zip(b, b[1:]):
                assert h0 == l1 and l0 <= h0
            sizes = [h - l for l, h in b]
            assert max(sizes) - min(sizes) <= 1024
            if n >= 4096 * world:
                assert all(l % 512 == 0 for l, _ in b)          # whole CTA passes


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.pyoracle import Oracle
    from som_lvq_pak_b200 import distributed as D
    o = Oracle()
    # the communicator id: made on rank 0 only, identical bytes on every rank afterwards
    uid = D.exchange_unique_id(lambda: bytes(range(128)) if rank == 0 else b"\0" * 128)
    assert uid == bytes(range(128))
    rng = np.random.default_rng(5)
    M, dim, N, L = 40, 6, 1000, 4
    codes = rng.random((M, dim), dtype=np.float32)
    data = rng.random((N, dim), dtype=np.float32)
    cl = rng.integers(0, L, M)
    dl = rng.integers(0, L, N)
    ct = torch.from_numpy(codes.copy() if rank == 0 else np.zeros_like(codes))
    dist.broadcast(ct, 0)                                  # codebook replicated from rank 0
    lo, hi = D.shard_bounds(N, rank, world)
    idx, diff, ret = o.search(ct.numpy(), data[lo:hi], 1)
    # {double sum} + {int64 n_found, hist, confusion}: the two vectors of the grouped all-reduce
    counts = np.zeros(1 + M + L * L, np.int64)
    counts[0] = len(idx)
    counts[1:1 + M] = np.bincount(idx[:, 0], minlength=M)
    np.add.at(counts[1 + M:].reshape(L, L), (dl[lo:hi], cl[idx[:, 0]]), 1)
    tsum = torch.tensor([np.sqrt(diff[:, 0].astype(np.float64)).sum()], dtype=torch.float64)
    tcnt = torch.from_numpy(counts)
    dist.all_reduce(tsum)
    dist.all_reduce(tcnt)
    # per-row results: every rank owns rows [lo, hi) of the caller's arrays (no gather in the product;
    # the test collects them to compare with the unsharded search)
    parts = [None] * world
    dist.all_gather_object(parts, (lo, hi, idx))
    if rank == 0:
        gidx, gdiff, _ = o.search(codes, data, 1)
        whole = np.concatenate([p[2] for p in sorted(parts, key=lambda t: t[0])])
        assert np.array_equal(whole, gidx)
        tot = tcnt.numpy()
        assert tot[0] == N
        assert np.array_equal(tot[1:1 + M], np.bincount(gidx[:, 0], minlength=M))
        gconf = np.zeros((L, L), np.int64)
        np.add.at(gconf, (dl, cl[gidx[:, 0]]), 1)
        assert np.array_equal(tot[1 + M:].reshape(L, L), gconf)
        assert abs(float(tsum[0]) - np.sqrt(gdiff[:, 0].astype(np.float64)).sum()) < 1e-9
        open(os.path.join(tmp, "ok"), "w").write("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo(tmp_path):
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()


def test_replay_qerror_is_the_sequential_float_sum():
    import som_lvq_pak_b200 as b
    from som_lvq_pak_b200 import _lib
    rng = np.random.default_rng(1)
    diff = rng.random((5000, 1), dtype=np.float32) * 30
    nf = np.ones(5000, np.int32)
    nf[::17] = 0
    q = np.float32(0)
    for d, f in zip(diff[:, 0], nf):
        if f:
            q = np.float32(np.float64(q) + np.sqrt(np.float64(d)))
    got = _lib.load().bmu_replay_qerror(diff.ctypes.data, nf.ctypes.data, 5000, 1)
    assert np.float32(got) == q


def test_vfind_trial_split_and_selection():
    from som_lvq_pak_b200.distributed import select_best, trial_numbers
    for trials in (1, 4, 7, 16):
        for world in (1, 2, 3, 8):
            parts = [trial_numbers(trials, r, world) for r in range(world)]
            assert sorted(sum(parts, []), reverse=True) == list(range(trials, 0, -1))
            assert all(p == sorted(p, reverse=True) for p in parts)
    # strictly-smaller rule while counting down: ties go to the larger trial number
    assert select_best([(np.float32(3.0), 1), (np.float32(2.0), 2), (np.float32(2.0), 4), (np.float32(5.0), 3)]) == (np.float32(2.0), 4)
    assert select_best([(np.float32(1.5), 1)]) == (np.float32(1.5), 1)
