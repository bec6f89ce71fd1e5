"""GPU parity of the batch winner search (through the C ABI, host-pointer entry point)
against (a) the golden vectors generated from the unmodified reference and (b) the CPU
oracle on seeded inputs.  Bar: bit-exact indices, squared distances and return values."""
import numpy as np
import pytest

from conftest import assert_bits_equal

pytestmark = pytest.mark.gpu

SHAPES = ["lowdim", "c3like", "odd", "c4like", "tiny", "wide"]
TAGS = {"u": ("codes", "data", None), "q": ("qcodes", "qdata", None),
        "m": ("qcodes", "qdata", "mask"), "nf": ("nfcodes", "nfdata", None)}


def check(engine, codes, data, k, mask, exp, what):
    idx, diff, nf = engine.find_winner_knn(codes, data, k, mask)
    assert_bits_equal(idx, exp[0], what + " idx")
    assert_bits_equal(diff, exp[1], what + " diff")
    assert_bits_equal(nf, exp[2], what + " ret")


@pytest.mark.parametrize("path", [1, 0, 2])
@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("k", [1, 2, 5, 10])
def test_search_golden(engine, golden, shape, k, path):
    g = golden.search
    engine.set_search_path(path)
    try:
        for tag, (c, d, m) in TAGS.items():
            key = "%s_%s_k%d" % (shape, tag, k)
            if key + "_idx" not in g:
                continue
            check(engine, g[shape + "_" + c], g[shape + "_" + d], k,
                  None if m is None else g[shape + "_" + m],
                  (g[key + "_idx"], g[key + "_diff"], g[key + "_ret"]), key)
    finally:
        engine.set_search_path(0)


@pytest.mark.parametrize("path", [1, 0, 2])
@pytest.mark.parametrize("M,D,N", [(96, 5, 3840), (200, 20, 1962), (1000, 64, 4096), (10000, 64, 1500),
                                   (4096, 512, 300), (129, 7, 130), (1, 3, 10), (300, 100, 257)])
def test_search_vs_oracle_random(engine, oracle, M, D, N, path):
    rng = np.random.default_rng(M * 131 + D)
    codes = rng.random((M, D), dtype=np.float32)
    data = rng.random((N, D), dtype=np.float32)
    engine.set_search_path(path)
    try:
        for k in (1, 5):
            check(engine, codes, data, k, None, oracle.search(codes, data, k), "rand k=%d" % k)
    finally:
        engine.set_search_path(0)


@pytest.mark.parametrize("path", [1, 0, 2])
def test_search_adversarial(engine, oracle, path):
    """exact ties, duplicated code vectors, masks, subnormal-scale values, huge values"""
    rng = np.random.default_rng(7)
    M, D, N = 500, 64, 1000
    codes = (rng.integers(0, 3, (M, D)) / 2).astype(np.float32)
    data = (rng.integers(0, 3, (N, D)) / 2).astype(np.float32)
    codes[100:200] = codes[0:100]              # duplicates: lowest index must win for k=1
    data[:50] = codes[100:150]                 # zero distance to two codes each
    engine.set_search_path(path)
    try:
        for k in (1, 2, 7, 16):
            check(engine, codes, data, k, None, oracle.search(codes, data, k), "ties k=%d" % k)
        mask = (rng.random((N, D)) < 0.5).astype(np.uint8)
        mask[10] = 1
        mask[11] = 0
        for k in (1, 3):
            check(engine, codes, data, k, mask, oracle.search(codes, data, k, mask), "mask k=%d" % k)
        # magnitudes that make squares subnormal / overflow: must leave the packed (.ftz) kernel
        tiny = data.copy()
        tiny[::3] *= np.float32(1e-30)
        tc = codes.copy()
        tc[::5] *= np.float32(1e-30)
        for k in (1, 4):
            check(engine, tc, tiny, k, None, oracle.search(tc, tiny, k), "tiny k=%d" % k)
        huge = data.copy()
        huge[::4] *= np.float32(3e19)
        for k in (1, 4):
            check(engine, codes, huge, k, None, oracle.search(codes, huge, k), "huge k=%d" % k)
        # k larger than the codebook
        check(engine, codes[:3], data[:40], 8, None, oracle.search(codes[:3], data[:40], 8), "k>M")
    finally:
        engine.set_search_path(0)


def test_search_chunked_host_path(engine, oracle):
    """N large enough that bmu_search cuts the rows into several overlapped chunks"""
    rng = np.random.default_rng(3)
    M, D = 64, 512
    N = 300000                                  # 614 MB of input -> 3 chunks of 256 MB
    codes = rng.random((M, D), dtype=np.float32)
    data = rng.random((N, D), dtype=np.float32)
    idx, diff, nf = engine.find_winner_euc(codes, data)
    sub = np.r_[0:2000, 131000:133000, N - 2000:N]
    e = oracle.search(codes, data[sub], 1)
    assert_bits_equal(idx[sub], e[0])
    assert_bits_equal(diff[sub], e[1])
    assert (nf == 1).all()
    # size-independent property: every reported distance is the distance to the reported code
    # and no code is closer (checked in float64 on a sample with a margin)
    s = rng.integers(0, N, 500)
    d64 = ((data[s, None, :].astype(np.float64) - codes[None].astype(np.float64)) ** 2).sum(-1)
    assert np.allclose(d64[np.arange(500), idx[s, 0]], diff[s, 0], rtol=1e-5)
    assert (d64.min(1) >= diff[s, 0] * (1 - 1e-5)).all()


def test_filter_path_certifies_most_rows(engine, oracle):
    """the tensor-core filter must actually answer the rows (not silently fall back to K1)"""
    rng = np.random.default_rng(21)
    for M, D, N, k in [(10000, 64, 8192, 1), (4096, 512, 2048, 1), (4096, 512, 1024, 5), (2000, 20, 4096, 1)]:
        codes = rng.random((M, D), dtype=np.float32) + np.float32(3.0)      # offset: centring matters
        data = rng.random((N, D), dtype=np.float32) + np.float32(3.0)
        engine.set_search_path(2)
        try:
            idx, diff, nf = engine.find_winner_knn(codes, data, k)
            bd = engine.last_search_breakdown()
        finally:
            engine.set_search_path(0)
        e = oracle.search(codes, data, k)
        assert_bits_equal(idx, e[0], "k2 idx M=%d D=%d k=%d" % (M, D, k))
        assert_bits_equal(diff, e[1], "k2 diff")
        assert bd["k2_certified"] >= 0.97 * N, bd
        assert bd["k2_certified"] + bd["k2_failed"] == N, bd


@pytest.mark.parametrize("D", [5, 20, 40, 56, 64, 77, 88, 89])
def test_record_kernel_every_k_depth(engine, oracle, D):
    """k = 1 on the filter path: D <= 88 runs k2_rec_kernel with 1..6 unrolled MMAs per accumulation
    (operand K = 16..96), D = 89 is the first shape of the streaming kernel; ragged M and N"""
    rng = np.random.default_rng(100 + D)
    M, N = 777, 2100
    codes = rng.random((M, D), dtype=np.float32) - np.float32(0.5)
    data = rng.random((N, D), dtype=np.float32) - np.float32(0.5)
    data[7] = codes[300]                                    # an exact hit
    engine.set_search_path(2)
    try:
        check(engine, codes, data, 1, None, oracle.search(codes, data, 1), "rec D=%d" % D)
        bd = engine.last_search_breakdown()
    finally:
        engine.set_search_path(0)
    assert bd["k2_certified"] + bd["k2_failed"] == N and bd["k2_certified"] >= 0.9 * N, bd


@pytest.mark.parametrize("path", [0, 2])
@pytest.mark.parametrize("M,D,N,k", [(1500, 2048, 600, 1), (1500, 2048, 300, 5), (65536, 128, 3000, 1),
                                     (100000, 16, 4000, 1), (100000, 16, 2000, 3), (513, 3000, 257, 2),
                                     (40000, 8, 5000, 1)])
def test_search_extreme_shapes(engine, oracle, M, D, N, k, path):
    """large D (many operand K slices, streaming kernel), large M (C5's 256 x 256 map, more code tiles than
    one pass holds), very low D with a huge codebook; the oracle checks a subsample of the rows"""
    rng = np.random.default_rng(M + D + k)
    codes = rng.random((M, D), dtype=np.float32)
    data = rng.random((N, D), dtype=np.float32)
    data[3] = codes[M - 1]                                  # exact hit on the last code
    data[4] = codes[0]
    engine.set_search_path(path)
    try:
        idx, diff, nf = engine.find_winner_knn(codes, data, k)
    finally:
        engine.set_search_path(0)
    budget = 1.5e9                                          # element-ops for the oracle
    sub = np.arange(N)[: max(16, int(budget / (M * D)))]
    e = oracle.search(codes, data[sub], k)
    assert_bits_equal(idx[sub], e[0], "idx")
    assert_bits_equal(diff[sub], e[1], "diff")
    assert (nf == k).all()
    assert idx[3, 0] == M - 1 and diff[3, 0] == 0.0 and idx[4, 0] == 0


@pytest.mark.parametrize("path", [1, 0, 2])
@pytest.mark.parametrize("M,D", [(1000, 64), (700, 33), (600, 200)])
def test_search_dev_writes_stay_inside_outputs(engine, oracle, M, D, path):
    """device-pointer entry point with the three output arrays embedded in guard bands: every
    kernel of every path (tile / warp / list kernels, record and streaming filter, re-rank) must
    write rows [0, N) x k only — ragged N around the 128-row tile and 512-row pass sizes"""
    import torch
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(M + D + path)
    codes = rng.random((M, D), dtype=np.float32)
    cb = engine.Codebook(codes)
    G = 4096
    SENT = 0x5A5A5A5A
    engine.set_search_path(path)
    try:
        for N in (1, 127, 129, 511, 513, 2049):
            data = rng.random((N, D), dtype=np.float32)
            data[::7] = codes[rng.integers(0, M, len(data[::7]))]          # zero distances
            d_data = torch.from_numpy(data).to(dev)
            for k in (1, 2, 5):
                bufs = [torch.full((N * w + 2 * G,), SENT, dtype=torch.int32, device=dev) for w in (k, k, 1)]
                ptr = [b.data_ptr() + 4 * G for b in bufs]
                cb.search_dev(d_data.data_ptr(), N, k, ptr[0], ptr[1], ptr[2])
                torch.cuda.synchronize()
                host = [b.cpu().numpy() for b in bufs]
                for h, name in zip(host, ("idx", "diff", "nfound")):
                    assert (h[:G] == SENT).all() and (h[-G:] == SENT).all(), \
                        "%s written outside [0, N*k): N=%d k=%d path=%d" % (name, N, k, path)
                exp = oracle.search(codes, data, k)
                assert_bits_equal(host[0][G:-G].reshape(N, k), exp[0], "idx N=%d k=%d" % (N, k))
                assert_bits_equal(host[1][G:-G].view(np.float32).reshape(N, k), exp[1], "diff N=%d k=%d" % (N, k))
                assert_bits_equal(host[2][G:-G], exp[2], "ret N=%d k=%d" % (N, k))
    finally:
        engine.set_search_path(0)


def test_host_search_writes_stay_inside_outputs(engine, oracle):
    """bmu_search with caller arrays embedded in guard bands: the staged chunks' drain copies
    (several chunks, the last one ragged) write rows [0, N) x k only"""
    import ctypes as C
    from som_lvq_pak_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(11)
    M, D, G = 200, 256, 1024
    codes = rng.random((M, D), dtype=np.float32)
    cb = engine.Codebook(codes)
    for N, k in ((1, 1), (70001, 1), (130003, 5)):       # 58 MB chunks: 56 832 rows of 1 KB each
        data = rng.random((N, D), dtype=np.float32)
        outs = [np.full(N * w + 2 * G, 0x5A5A5A5A, np.int32) for w in (k, k, 1)]
        ptr = [C.c_void_p(o.ctypes.data + 4 * G) for o in outs]
        _lib.check(lib.bmu_search(cb._h, data.ctypes.data_as(C.c_void_p), None, N, k, ptr[0], ptr[1], ptr[2]))
        for o in outs:
            assert (o[:G] == 0x5A5A5A5A).all() and (o[-G:] == 0x5A5A5A5A).all()
        sub = np.unique(np.r_[0:min(N, 500), max(0, N - 500):N, rng.integers(0, N, 500)])
        e = oracle.search(codes, data[sub], k)
        assert_bits_equal(outs[0][G:-G].reshape(N, k)[sub], e[0])
        assert_bits_equal(outs[1][G:-G].view(np.float32).reshape(N, k)[sub], e[1])
        assert_bits_equal(outs[2][G:-G][sub], e[2])
