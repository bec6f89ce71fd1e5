"""GPU parity of the data-parallel split behind the C ABI: the sharded search
(bmu_multi_search: contiguous row shards, codebook replicated, per-row results straight into the
caller's arrays, statistics combined by one grouped all-reduce) and the per-shard statistics kernel
(bmu_search_stats_dev), against the oracle.  On one GPU the shards are LOGICAL (several device
contexts on the same device); with two or more visible GPUs the same tests also run
over NCCL.  Reference accumulators: find_qerror som_rout.c:710-721, compute_cmatr cmatr.c:84-109.
Bar: bit-exact idx / diff / nfound, exact int64 counts, the double sum within N ulp."""
import os

import numpy as np
import pytest

from conftest import assert_bits_equal

pytestmark = pytest.mark.gpu


def _inputs(seed, M, D, N, L, masked=True):
    rng = np.random.default_rng(seed)
    codes = (np.round(rng.random((M, D)) * 8) / 8).astype(np.float32)        # exact ties
    codes[M // 2] = codes[3]                                                 # duplicate code vector
    data = (np.round(rng.random((N, D)) * 8) / 8).astype(np.float32)
    mask = None
    if masked:
        mask = (rng.random((N, D)) < 0.05).astype(np.uint8)
        mask[7] = 1                                                          # an all-masked row
        data[mask != 0] = 0.0
    return codes, data, mask, rng.integers(0, L, M).astype(np.int32), rng.integers(0, L, N).astype(np.int32)


def _expected_stats(eidx, ediff, eret, M, L, cl, dl):
    found = eret != 0
    j = eidx[found, 0]
    conf = np.zeros((L, L), np.int64)
    np.add.at(conf, (dl[found], cl[j]), 1)
    return (np.sqrt(ediff[found, 0].astype(np.float64)).sum(), int(found.sum()),
            np.bincount(j, minlength=M).astype(np.int64), conf)


@pytest.mark.parametrize("nshards", [1, 3, 4])
@pytest.mark.parametrize("k", [1, 5])
def test_logical_shards_vs_oracle(engine, oracle, nshards, k):
    M, D, N, L = 300, 16, 20000, 5
    codes, data, mask, cl, dl = _inputs(11 + k, M, D, N, L)
    mc = engine.MultiCodebook(codes, nshards=nshards, code_label=cl)
    try:
        assert mc.shards() == nshards
        idx, diff, nf, st = mc.find_winners(data, k, mask, stats=True, hist=True, sample_label=dl, n_labels=L)
    finally:
        mc.close()
    eidx, ediff, eret = oracle.search(codes, data, k, mask)
    assert_bits_equal(idx, eidx, "idx")
    assert_bits_equal(diff, ediff, "diff")
    assert_bits_equal(nf, eret, "ret")
    esum, efound, ehist, econf = _expected_stats(eidx, ediff, eret, M, L, cl, dl)
    assert st["n_found"] == efound == N - 1
    assert np.array_equal(st["hist"], ehist)
    assert np.array_equal(st["confusion"], econf)
    assert abs(st["sum_sqrt"] - esum) <= N * np.spacing(esum)      # double sum, another order: N ulp


def test_sharded_filter_path_and_chunk_ring(engine, oracle, monkeypatch):
    """the tensor-core path under sharding, with the host pipeline forced into many small chunks so that
    ring slots are reused (pageable numpy buffers: staged through the pinned ring)"""
    M, D, N = 1200, 64, 30000
    rng = np.random.default_rng(3)
    codes = rng.random((M, D), dtype=np.float32)
    data = rng.random((N, D), dtype=np.float32)
    monkeypatch.setenv("SOMLVQ_CHUNK_ROWS", "2048")
    engine.set_search_path(engine.PATH_FILTER)
    mc = engine.MultiCodebook(codes, nshards=2)
    try:
        idx, diff, nf, st = mc.find_winners(data, 1, None, stats=True, hist=True)
        idx2, diff2, nf2, st2 = mc.find_winners(data, 1, None, stats=True, hist=True)
    finally:
        mc.close()
        engine.set_search_path(engine.PATH_AUTO)
    eidx, ediff, eret = oracle.search(codes, data, 1)
    assert_bits_equal(idx, eidx, "idx")
    assert_bits_equal(diff, ediff, "diff")
    assert np.array_equal(st["hist"], np.bincount(eidx[:, 0], minlength=M))
    assert st["n_found"] == N
    # run-to-run: the double sum is reduced in a fixed order
    assert st["sum_sqrt"] == st2["sum_sqrt"] and np.array_equal(idx, idx2)
    assert abs(st["sum_sqrt"] - np.sqrt(ediff[:, 0].astype(np.float64)).sum()) <= N * np.spacing(st["sum_sqrt"])


def test_single_gpu_host_search_many_chunks(engine, oracle, monkeypatch):
    """bmu_search (host pointers) with more chunks than ring slots, masks, k = 2"""
    M, D, N = 96, 5, 9000
    codes, data, mask, _, _ = _inputs(5, M, D, N, 3)
    monkeypatch.setenv("SOMLVQ_CHUNK_ROWS", "1000")
    idx, diff, nf = engine.find_winner_knn(codes, data, 2, mask)
    eidx, ediff, eret = oracle.search(codes, data, 2, mask)
    assert_bits_equal(idx, eidx, "idx")
    assert_bits_equal(diff, ediff, "diff")
    assert_bits_equal(nf, eret, "ret")


def test_search_stats_dev_vs_oracle(engine, oracle):
    """bmu_search_stats_dev on device buffers: exact counts, deterministic double sum"""
    import torch
    M, D, N, L = 500, 20, 40000, 7
    codes, data, mask, cl, dl = _inputs(21, M, D, N, L)
    idx, diff, nf = engine.find_winner_knn(codes, data, 1, mask)
    dev = torch.device("cuda", 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d_idx, d_diff, d_nf, d_cl, d_dl = t(idx), t(diff), t(nf), t(cl), t(dl)
    sums = []
    for rep in range(2):
        d_sum = torch.zeros(1, dtype=torch.float64, device=dev)
        d_cnt = torch.zeros(1 + M + L * L, dtype=torch.int64, device=dev)
        half = N // 2 + 13                   # two calls accumulate, as two chunks / shards would
        for lo, hi in ((0, half), (half, N)):
            engine.search_stats_dev(d_idx[lo:hi].data_ptr(), d_diff[lo:hi].data_ptr(), d_nf[lo:hi].data_ptr(),
                                    hi - lo, 1, M, d_sum.data_ptr(), d_cnt.data_ptr(), d_cnt.data_ptr() + 8,
                                    d_dl[lo:hi].data_ptr(), d_cl.data_ptr(), L, d_cnt.data_ptr() + 8 * (1 + M),
                                    torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        sums.append(float(d_sum[0]))
        cnt = d_cnt.cpu().numpy()
    eidx, ediff, eret = oracle.search(codes, data, 1, mask)
    esum, efound, ehist, econf = _expected_stats(eidx, ediff, eret, M, L, cl, dl)
    assert cnt[0] == efound
    assert np.array_equal(cnt[1:1 + M], ehist)
    assert np.array_equal(cnt[1 + M:].reshape(L, L), econf)
    assert sums[0] == sums[1]
    assert abs(sums[0] - esum) <= N * np.spacing(esum)


def test_searches_on_two_streams_do_not_race(engine, oracle):
    """two bmu_search_dev calls in flight on different streams share the device's scratch; the library
    orders them (event), so both results are right (scratch race)"""
    import torch
    from som_lvq_pak_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    M, D, N = 2000, 64, 20000
    rng = np.random.default_rng(9)
    codes = rng.random((M, D), dtype=np.float32)
    datas = [rng.random((N, D), dtype=np.float32) for _ in range(2)]
    cb = engine.Codebook(codes)
    streams = [torch.cuda.Stream(dev) for _ in range(2)]
    outs = []
    for d, s in zip(datas, streams):
        dd = torch.from_numpy(d).to(dev)
        o = (dd, torch.empty((N, 1), dtype=torch.int32, device=dev), torch.empty((N, 1), dtype=torch.float32, device=dev),
             torch.empty(N, dtype=torch.int32, device=dev))
        outs.append(o)
    torch.cuda.synchronize()
    for rep in range(3):
        for o, s in zip(outs, streams):
            _lib.check(lib.bmu_search_dev(cb._h, o[0].data_ptr(), None, N, 1, o[1].data_ptr(), o[2].data_ptr(),
                                          o[3].data_ptr(), s.cuda_stream))
    torch.cuda.synchronize()
    for d, o in zip(datas, outs):
        eidx, ediff, _ = oracle.search(codes, d, 1)
        assert_bits_equal(o[1].cpu().numpy(), eidx, "idx")
        assert_bits_equal(o[2].cpu().numpy(), ediff, "diff")
    cb.close()


def test_real_devices_nccl(engine, oracle):
    """every visible GPU (>= 2): codebook ncclBroadcast, shards on distinct devices, NCCL all-reduce"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    M, D, N, L = 1500, 64, 60000, 4
    codes, data, _, cl, dl = _inputs(31, M, D, N, L, masked=False)
    mc = engine.MultiCodebook(codes, nshards=0, code_label=cl)
    try:
        assert mc.shards() == torch.cuda.device_count()
        idx, diff, nf, st = mc.find_winners(data, 1, None, stats=True, hist=True, sample_label=dl, n_labels=L)
    finally:
        mc.close()
    eidx, ediff, eret = oracle.search(codes, data, 1)
    assert_bits_equal(idx, eidx, "idx")
    assert_bits_equal(diff, ediff, "diff")
    esum, efound, ehist, econf = _expected_stats(eidx, ediff, eret, M, L, cl, dl)
    assert st["n_found"] == efound and np.array_equal(st["hist"], ehist) and np.array_equal(st["confusion"], econf)
    assert abs(st["sum_sqrt"] - esum) <= N * np.spacing(esum)


def test_page_locked_and_pageable_buffers_agree(engine, oracle, monkeypatch):
    """bmu_search reads page-locked caller buffers (bmu_host_register / bmu_host_alloc) by DMA directly and stages
    pageable ones through its pinned ring: same results either way, several chunks each"""
    import ctypes as C
    from som_lvq_pak_b200 import _lib
    lib = _lib.load()
    M, D, N = 700, 64, 40000                      # 10 MB of rows: above the 1 MB "leave it to the driver" limit
    rng = np.random.default_rng(17)
    codes = rng.random((M, D), dtype=np.float32)
    data = rng.random((N, D), dtype=np.float32)
    monkeypatch.setenv("SOMLVQ_CHUNK_ROWS", "6000")
    cb = engine.Codebook(codes)
    out = []
    for pinned in (False, True):
        idx = np.empty((N, 1), np.int32)
        diff = np.empty((N, 1), np.float32)
        nf = np.empty(N, np.int32)
        bufs = [data, idx, diff, nf]
        if pinned:
            for b in bufs:
                _lib.check(lib.bmu_host_register(b.ctypes.data, b.nbytes))
        try:
            _lib.check(lib.bmu_search(cb._h, data.ctypes.data, None, N, 1, idx.ctypes.data, diff.ctypes.data,
                                      nf.ctypes.data))
        finally:
            if pinned:
                for b in bufs:
                    lib.bmu_host_unregister(b.ctypes.data)
        out.append((idx, diff, nf))
    # library-allocated page-locked memory
    p = lib.bmu_host_alloc(data.nbytes)
    assert p
    C.memmove(p, data.ctypes.data, data.nbytes)
    idx3 = np.empty((N, 1), np.int32)
    diff3 = np.empty((N, 1), np.float32)
    nf3 = np.empty(N, np.int32)
    _lib.check(lib.bmu_search(cb._h, p, None, N, 1, idx3.ctypes.data, diff3.ctypes.data, nf3.ctypes.data))
    lib.bmu_host_free(p)
    cb.close()
    eidx, ediff, eret = oracle.search(codes, data, 1)
    for idx, diff, nf in out + [(idx3, diff3, nf3)]:
        assert_bits_equal(idx, eidx, "idx")
        assert_bits_equal(diff, ediff, "diff")
        assert_bits_equal(nf, eret, "ret")


def test_vfind_trials_over_the_visible_gpus():
    """vfind: independent trials, one stream of trials per GPU, (error, trial) pairs gathered and the
    winner's map broadcast over NCCL -- the same map, bit for bit, as the single-process search
    (tools/vfind_multi.py under torchrun; needs two GPUs)"""
    import subprocess
    import sys
    import torch
    from conftest import ROOT
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    n = min(n, 4)
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
                        "--master-addr", "127.0.0.1", "--master-port", "29577",
                        os.path.join(ROOT, "tools", "vfind_multi.py"), "8"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:]
    assert "identical to the single-process search: True" in p.stdout


@pytest.mark.parametrize("N", [0, 1, 3, 130])
def test_fewer_rows_than_shards(engine, oracle, N):
    """empty and nearly empty shards: 4 shards for 0 / 1 / 3 / 130 rows (find_qerror on an empty file adds nothing,
    som_rout.c:710-721)"""
    M, D, L = 50, 7, 3
    codes, data, mask, cl, dl = _inputs(41, M, D, max(N, 8), L, masked=False)
    data, dl = data[:N], dl[:N]
    mc = engine.MultiCodebook(codes, nshards=4, code_label=cl)
    try:
        idx, diff, nf, st = mc.find_winners(data, 2, None, stats=True, hist=True, sample_label=dl, n_labels=L)
    finally:
        mc.close()
    assert idx.shape == (N, 2) and st["n_found"] == N and int(st["hist"].sum()) == N
    if N:
        eidx, ediff, eret = oracle.search(codes, data, 2)
        assert_bits_equal(idx, eidx, "idx")
        assert_bits_equal(diff, ediff, "diff")
        assert_bits_equal(nf, eret, "ret")
        assert int(st["confusion"].sum()) == N


def test_process_per_gpu_mode_vs_oracle():
    """the launcher-per-GPU mode bench.py --gpus N uses (bmu_comm_unique_id / bmu_comm_init_rank / bmu_comm_broadcast_dev /
    bmu_search_stats_dev / bmu_comm_allreduce_stats_dev), under torchrun on every visible GPU (at most 4), rows and
    statistics against the oracle (tests/sharded_search_check.py; needs two GPUs)"""
    import subprocess
    import sys
    import torch
    from conftest import ROOT
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    n = min(n, 4)
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
                        "--master-addr", "127.0.0.1", "--master-port", "29578",
                        os.path.join(ROOT, "tests", "sharded_search_check.py")],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:]
    assert "identical to the oracle: True" in p.stdout
