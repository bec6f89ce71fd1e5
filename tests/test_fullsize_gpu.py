"""GPU: BASELINE.json's full-size configurations (C3 10 M x 64 vs 100x100, C4 1 M x 512 vs 4096
units with k = 5, C5 256x256x128 map) checked through size-independent properties, plus direct
oracle comparison on samples the CPU finishes in seconds:
  * every row gets a winner, the BMU histogram sums to N;
  * the tensor-core filter path (K2) and the exact FP32 path (K1) agree bit for bit on every row (C3) / a sample;
  * a sample agrees bit for bit with the oracle;
  * searching the codebook against itself returns the identity with distance 0;
  * k-NN lists are sorted by the reference's rule and hold distinct codes;
  * a prefix of the C5 training schedule reproduces the oracle's codebook."""
import numpy as np
import pytest
import torch

from bench import synth_numpy, synth_rows_torch
from conftest import assert_bits_equal

pytestmark = pytest.mark.gpu


def dev_search(engine, cb, data, k, path):
    n = data.shape[0]
    idx = torch.empty((n, k), dtype=torch.int32, device=data.device)
    diff = torch.empty((n, k), dtype=torch.float32, device=data.device)
    nf = torch.empty(n, dtype=torch.int32, device=data.device)
    engine.set_search_path(path)
    try:
        cb.search_dev(data.data_ptr(), n, k, idx.data_ptr(), diff.data_ptr(), nf.data_ptr())
        torch.cuda.synchronize()
    finally:
        engine.set_search_path(0)
    return idx, diff, nf


def test_c3_full_size(engine, oracle):
    dev = torch.device("cuda:0")
    N, D, M = 10_000_000, 64, 10_000
    codes = synth_rows_torch(2, 0, M, D, dev)
    data = synth_rows_torch(1, 0, N, D, dev)
    cb = engine.Codebook(codes.cpu().numpy())
    idx, diff, nf = dev_search(engine, cb, data, 1, 2)                      # K2 filter path
    bd = engine.last_search_breakdown()
    assert bd["k2_certified"] + bd["k2_failed"] == N and bd["k2_certified"] > 0.99 * N, bd
    assert bool((nf == 1).all())
    hist = torch.bincount(idx[:, 0].long(), minlength=M)
    assert int(hist.sum()) == N and int(idx.min()) >= 0 and int(idx.max()) < M
    # the exact FP32 path (K1) on ALL 10 M rows: identical winners and distance bits, row for row
    eidx, ediff, _ = dev_search(engine, cb, data, 1, 1)
    assert bool((eidx == idx).all()) and bool((ediff.view(torch.int32) == diff.view(torch.int32)).all())
    del eidx, ediff
    # a second run of the filter path gives the same bits (no order dependence in the fused kernel)
    idx2, diff2, _ = dev_search(engine, cb, data, 1, 2)
    assert bool((idx2 == idx).all()) and bool((diff2.view(torch.int32) == diff.view(torch.int32)).all())
    del idx2, diff2
    g = torch.Generator(device="cpu").manual_seed(5)
    rows = torch.randperm(N, generator=g)[:200_000].to(dev)
    # oracle on a smaller sample
    h = rows[:1500]
    o = oracle.search(codes.cpu().numpy(), data[h].cpu().numpy(), 1)
    assert_bits_equal(idx[h].cpu().numpy(), o[0])
    assert_bits_equal(diff[h].cpu().numpy(), o[1])
    # the codebook against itself: identity, distance 0
    sidx, sdiff, _ = dev_search(engine, cb, codes, 1, 2)
    assert bool((sidx[:, 0] == torch.arange(M, device=dev, dtype=torch.int32)).all()) and float(sdiff.max()) == 0.0
    cb.close()


def test_c4_full_size_knn(engine, oracle):
    dev = torch.device("cuda:0")
    N, D, M, k = 1_000_000, 512, 4096, 5
    codes = synth_rows_torch(2, 0, M, D, dev)
    data = synth_rows_torch(3, 0, N, D, dev)
    cb = engine.Codebook(codes.cpu().numpy())
    idx, diff, nf = dev_search(engine, cb, data, k, 0)                      # AUTO: K2 streaming kernel
    bd = engine.last_search_breakdown()
    assert bd["k2_certified"] > 0.98 * N, bd
    assert bool((nf == k).all())
    # k-NN order (lvq_pak.c:197): distance ascending, ties by descending index; distinct codes
    d0, d1 = diff[:, :-1], diff[:, 1:]
    i0, i1 = idx[:, :-1], idx[:, 1:]
    assert bool(((d0 < d1) | ((d0 == d1) & (i0 > i1))).all())
    rows = torch.arange(0, N, 10, device=dev)                               # 100 000 rows through the exact k-NN kernel
    sub = data[rows].contiguous()
    eidx, ediff, _ = dev_search(engine, cb, sub, k, 1)                       # exact path
    assert bool((eidx == idx[rows]).all()) and bool((ediff.view(torch.int32) == diff[rows].view(torch.int32)).all())
    o = oracle.search(codes.cpu().numpy(), sub[:200].cpu().numpy(), k)
    assert_bits_equal(idx[rows[:200]].cpu().numpy(), o[0])
    assert_bits_equal(diff[rows[:200]].cpu().numpy(), o[1])
    # k = 1 (accuracy / classify): first neighbour of the k-NN list unless a tie changes the rule
    idx1, diff1, _ = dev_search(engine, cb, data, 1, 0)
    assert bool((diff1[:, 0] == diff[:, 0]).all())
    # ... and equal, on every one of the 1 M rows, to the exact FP32 kernel's winner and distance bits
    eidx1, ediff1, _ = dev_search(engine, cb, data, 1, 1)
    assert bool((eidx1 == idx1).all()) and bool((ediff1.view(torch.int32) == diff1.view(torch.int32)).all())
    cb.close()


def test_c5_schedule_prefix(engine, oracle):
    """256x256 hexa gaussian map, 128-dim, the first steps of the rlen 1e6 schedule (fused K3 kernel)"""
    N, D, xdim, ydim, length, steps = 100_000, 128, 256, 256, 1_000_000, 120
    data = synth_numpy(4, 0, N * D).reshape(N, D)
    codes = synth_numpy(5, 0, xdim * ydim * D).reshape(xdim * ydim, D)
    order = engine.rand_order(N, 3)
    s, ta, tr = engine.som_schedule(0, steps, length, 0.05, 100.0, engine.ALPHA_LINEAR, N, order)
    t = engine.Trainer(codes, data)
    t.set_som(xdim, ydim, engine.TOPOL_HEXA, engine.NEIGH_GAUSSIAN)
    t.steps(s, ta, tr)
    got = t.codes()
    t.close()
    exp = oracle.som_train_prefix(codes, data, xdim, ydim, 3, 2, length, steps, 0.05, 100.0, 1, order=order)
    np.testing.assert_allclose(got, exp, rtol=1e-6, atol=0)
    nbad = int((got.view(np.int32) != exp.view(np.int32)).sum())
    assert nbad <= got.size // 1_000_000 + 4, "%d of %d floats differ" % (nbad, got.size)


def test_c5_shape_bubble_bit_exact(engine, oracle):
    """the C5 shape (256x256 hexa, 128-dim, fused K3 kernel) with the BUBBLE neighbourhood: no exp(), so the
    codebook must be bit-identical to the oracle's after every step of the prefix (radius 100: most of the
    map moves at each step)"""
    N, D, xdim, ydim, length, steps = 100_000, 128, 256, 256, 1_000_000, 100
    data = synth_numpy(4, 0, N * D).reshape(N, D)
    codes = synth_numpy(5, 0, xdim * ydim * D).reshape(xdim * ydim, D)
    order = engine.rand_order(N, 3)
    s, ta, tr = engine.som_schedule(0, steps, length, 0.05, 100.0, engine.ALPHA_LINEAR, N, order)
    t = engine.Trainer(codes, data)
    t.set_som(xdim, ydim, engine.TOPOL_HEXA, engine.NEIGH_BUBBLE)
    t.steps(s, ta, tr)
    got = t.codes()
    t.close()
    exp = oracle.som_train_prefix(codes, data, xdim, ydim, 3, 1, length, steps, 0.05, 100.0, 1, order=order)
    assert_bits_equal(got, exp, "bubble C5 prefix")


@pytest.mark.parametrize("neigh", [1, 2])
def test_long_run_reaches_the_late_schedule(engine, oracle, neigh):
    """a COMPLETE run of 20 000 steps on a reduced map (32 x 24 hexa, 128-dim: the fused large-dimension
    kernel, several CTAs) over 3 000 samples: the sample list wraps six times (som_rout.c:602-610), the
    radius falls linearly from 12 to 1 (trad < 2 for the last 1 800 steps: single-unit neighbourhoods) and
    alpha to ~0.  Bubble: bit-exact.  Gaussian: the documented 1e-6 relative tolerance (double exp)."""
    N, D, xdim, ydim, length = 3000, 128, 32, 24, 20_000
    data = synth_numpy(14, 0, N * D).reshape(N, D)
    codes = synth_numpy(15, 0, xdim * ydim * D).reshape(xdim * ydim, D)
    order = engine.rand_order(N, 11)
    got = engine.som_training(codes, data, xdim, ydim, engine.TOPOL_HEXA, neigh, length, 0.05, 12.0, rand_seed=11)
    exp = oracle.som_train(codes, data, xdim, ydim, 3, neigh, length, 0.05, 12.0, 1, order=order)
    if neigh == 1:
        assert_bits_equal(got, exp, "bubble long run")
    else:
        np.testing.assert_allclose(got, exp, rtol=1e-6, atol=0)
        nbad = int((got.view(np.int32) != exp.view(np.int32)).sum())
        assert nbad <= got.size // 1000, "%d of %d floats differ" % (nbad, got.size)


def test_host_pointer_search_through_filter_path(engine, oracle):
    """bmu_search() with host buffers: 2.3 M x 64 rows are cut into eleven 58 MB chunks (pageable numpy
    memory: staged through the pinned ring) whose copies overlap the kernels; every chunk goes through the
    tensor-core filter.  A sample must agree with
    the oracle bit for bit and the chunk seams must not lose or duplicate rows."""
    rng = np.random.default_rng(9)
    N, D, M = 2_300_003, 64, 3000
    codes = rng.random((M, D), dtype=np.float32)
    data = rng.random((N, D), dtype=np.float32)
    idx, diff, nf = engine.find_winner_euc(codes, data)            # AUTO -> K2 (M >= 512, N >= 4096)
    bd = engine.last_search_breakdown()
    assert bd["k2_certified"] > 0
    assert (nf == 1).all() and idx.min() >= 0 and idx.max() < M
    seam = 148 * 512 * 3                                           # rows per chunk at D = 64: 64 MB in whole waves
    sub = np.r_[0:300, seam - 150:seam + 150, 2 * seam - 150:2 * seam + 150, 9 * seam - 150:9 * seam + 150, N - 300:N]
    e = oracle.search(codes, data[sub], 1)
    assert_bits_equal(idx[sub], e[0])
    assert_bits_equal(diff[sub], e[1])
    # every reported distance is the exact distance to the reported code (float64 check on a sample)
    s = rng.integers(0, N, 2000)
    d64 = ((data[s].astype(np.float64) - codes[idx[s, 0]].astype(np.float64)) ** 2).sum(-1)
    assert np.allclose(d64, diff[s, 0], rtol=1e-5)
