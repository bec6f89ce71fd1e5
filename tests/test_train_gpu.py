"""GPU parity of the online training kernel K3 against the golden codebooks produced by the
unmodified reference (som_training / lvq*_training) and against the CPU oracle.
Bar: bit-exact for bubble SOM and all LVQ variants; gaussian SOM within 1e-6 relative
(north_star: the double exp() of libm is not bit-portable), and bit-exact in practice."""
import numpy as np
import pytest

from conftest import assert_bits_equal

pytestmark = pytest.mark.gpu

GAUSS_RTOL = 1e-6


def close_or_equal(out, exp, neigh, what):
    if neigh == 1:
        assert_bits_equal(out, exp, what)
    else:
        np.testing.assert_allclose(out, exp, rtol=GAUSS_RTOL, atol=0, err_msg=what)


@pytest.mark.parametrize("topol", [3, 4])
@pytest.mark.parametrize("neigh", [1, 2])
def test_som_golden(engine, golden, topol, neigh):
    g = golden.som
    xdim, ydim = map(int, g["dims"])
    for at in (1, 2):
        for seed in (-1, 11):
            out = engine.som_training(g["codes"], g["data"], xdim, ydim, topol, neigh, 1500, 0.05, 4.0,
                                      at, rand_seed=None if seed < 0 else seed)
            close_or_equal(out, g["t%d_n%d_a%d_s%d" % (topol, neigh, at, seed)], neigh,
                           "som t%d n%d a%d s%d" % (topol, neigh, at, seed))
    out = engine.som_training(g["codes"], g["data"], xdim, ydim, topol, neigh, 900, 0.05, 3.0, 1,
                              mask=g["mask"], weight=g["weight"], fixed_xy=g["fixed"])
    close_or_equal(out, g["t%d_n%d_mwf" % (topol, neigh)], neigh, "som mask+weight+fixed")


def test_som_gaussian_bit_exact_in_practice(engine, golden):
    """stronger than the stated tolerance: on these sizes the trajectories are identical"""
    g = golden.som
    xdim, ydim = map(int, g["dims"])
    out = engine.som_training(g["codes"], g["data"], xdim, ydim, 3, 2, 1500, 0.05, 4.0, 1)
    exp = g["t3_n2_a1_s-1"]
    nbad = int((out.view(np.int32) != exp.view(np.int32)).sum())
    assert nbad == 0, "%d of %d floats differ" % (nbad, out.size)


@pytest.mark.parametrize("algo", [1, 2, 3, 4])
def test_lvq_golden(engine, golden, algo):
    g = golden.lvq
    for seed in (-1, 4):
        for at in (1, 2):
            alpha = 0.3 if algo == 4 else 0.05
            out = engine.lvq_training(algo, g["codes"], g["code_label"], g["data"], g["data_label"],
                                      4000, alpha, at, 0.3, 0.1, rand_seed=None if seed < 0 else seed)
            key = "algo%d_s%d_a%d" % (algo, seed, at)
            if algo == 4:
                out, ua = out
                assert ["%g" % v for v in ua] == list(g[key + "_lra"])
            assert_bits_equal(out, g[key], key)


@pytest.mark.parametrize("M,D,xdim", [(600, 16, 30), (4096, 8, 64), (150 * 150, 4, 150)])
def test_som_multi_cta_vs_oracle(engine, oracle, M, D, xdim):
    """maps large enough that the units are spread over many CTAs (grid exchange exercised)"""
    rng = np.random.default_rng(M)
    ydim = M // xdim
    codes = rng.random((M, D), dtype=np.float32)
    data = (np.round(rng.random((400, D)) * 16) / 16).astype(np.float32)   # ties in the search
    length = 300
    for neigh, topol in ((1, 3), (1, 4), (2, 3)):
        out = engine.som_training(codes, data, xdim, ydim, topol, neigh, length, 0.05, 8.0, 1, rand_seed=3)
        exp = oracle.som_train(codes, data, xdim, ydim, topol, neigh, length, 0.05, 8.0, 1,
                               order=oracle.shuffle_order(400, 3))
        close_or_equal(out, exp, neigh, "M=%d neigh=%d topol=%d" % (M, neigh, topol))


def test_lvq_multi_cta_vs_oracle(engine, oracle):
    rng = np.random.default_rng(11)
    M, D, N, L = 5000, 24, 900, 7
    codes = rng.random((M, D), dtype=np.float32)
    data = rng.random((N, D), dtype=np.float32)
    cl = rng.integers(1, L + 1, M).astype(np.int32)
    dl = rng.integers(1, L + 1, N).astype(np.int32)
    for algo in (1, 2, 3, 4):
        alpha = 0.3 if algo == 4 else 0.05
        out = engine.lvq_training(algo, codes, cl, data, dl, 1200, alpha, 1, 0.3, 0.1, rand_seed=8)
        exp, eua = oracle.lvq_train(algo, codes, cl, data, dl, 1200, alpha, 1, 0.3, 0.1,
                                    order=oracle.shuffle_order(N, 8))
        if algo == 4:
            out, ua = out
            assert_bits_equal(ua, eua, "unit alpha")
        assert_bits_equal(out, exp, "lvq algo %d" % algo)


def test_som_chunked_equals_whole(engine, golden):
    """stopping at snapshot boundaries (som_rout.c:650) must not change the trajectory"""
    g = golden.som
    xdim, ydim = map(int, g["dims"])
    snaps = []
    out = engine.som_training(g["codes"], g["data"], xdim, ydim, 3, 1, 1500, 0.05, 4.0, 1,
                              snapshot_interval=400, snapshot_cb=lambda le, c: snaps.append(le))
    assert snaps == [400, 800, 1200]
    assert_bits_equal(out, g["t3_n1_a1_s-1"])


@pytest.mark.parametrize("xdim,ydim,D,topol,neigh", [(20, 15, 70, 3, 1), (20, 15, 64, 4, 2), (64, 48, 64, 3, 2),
                                                       (64, 48, 100, 4, 1), (37, 29, 128, 3, 2)])
def test_som_large_dim_fused_kernel(engine, oracle, xdim, ydim, D, topol, neigh):
    """D >= 64 without masks / fixed points takes the fused update+search kernel (one unit per
    thread, half of it in registers); same bar as the generic kernel, -rand order and a radius
    that shrinks through the run"""
    rng = np.random.default_rng(xdim * 7 + D)
    M, N, rlen = xdim * ydim, 500, 400
    codes = rng.random((M, D), dtype=np.float32)
    data = rng.random((N, D), dtype=np.float32)
    for at, seed in ((1, None), (2, 9)):
        out = engine.som_training(codes, data, xdim, ydim, topol, neigh, rlen, 0.05, 6.0, at, rand_seed=seed)
        order = None if seed is None else oracle.shuffle_order(N, seed)
        exp = oracle.som_train(codes, data, xdim, ydim, topol, neigh, rlen, 0.05, 6.0, at, order=order)
        close_or_equal(out, exp, neigh, "fused som %dx%d D=%d t%d n%d" % (xdim, ydim, D, topol, neigh))
        if neigh == 2:
            nbad = int((out.view(np.int32) != exp.view(np.int32)).sum())
            assert nbad <= out.size // 100000 + 2, "%d of %d floats differ" % (nbad, out.size)


def test_vfind_python_orchestration(engine, golden):
    """distributed.vfind at world size 1 must find the reference vfind's map (tests/golden/demo_extra.npz)"""
    import datfile
    from som_lvq_pak_b200 import distributed as Dm
    data = datfile.parse(str(golden.demo["in_ex.dat"]))
    ref = datfile.parse(str(golden.demo_extra["vfind_cod"]))
    codes, q, n = Dm.vfind(data.points, data.points, 6, 4, 3, 1, 4, 300, 0.05, 4.0, 600, 0.02, 2.0)
    assert n == 4
    assert "%f" % (q / np.float32(data.points.shape[0])) == "9.632525"
    assert [["%g" % v for v in row] for row in codes] == [["%g" % v for v in row] for row in ref.points]


@pytest.mark.parametrize("xdim,ydim,D,neigh", [(30, 20, 700, 1), (12, 9, 2100, 2), (300, 300, 8, 1), (64, 48, 130, 2),
                                               (3, 2, 1, 1), (1, 1, 5, 2)])
def test_som_extreme_shapes(engine, oracle, xdim, ydim, D, neigh):
    """codebooks far from the benchmark shapes: very long vectors (the unit does not fit one register /
    shared-memory slice), 90 000 units, and degenerate 1-component / 1-unit maps"""
    rng = np.random.default_rng(xdim * 31 + D)
    M, N = xdim * ydim, 300
    rlen = 120 if M * D > 2_000_000 else 250
    codes = rng.random((M, D), dtype=np.float32)
    data = rng.random((N, D), dtype=np.float32)
    out = engine.som_training(codes, data, xdim, ydim, 3, neigh, rlen, 0.05, 5.0, 1, rand_seed=4)
    exp = oracle.som_train(codes, data, xdim, ydim, 3, neigh, rlen, 0.05, 5.0, 1, order=oracle.shuffle_order(N, 4))
    close_or_equal(out, exp, neigh, "som %dx%d D=%d n%d" % (xdim, ydim, D, neigh))


@pytest.mark.parametrize("M,D", [(3000, 600), (20, 2500), (30000, 12), (1, 4), (2, 1)])
def test_lvq_extreme_shapes(engine, oracle, M, D):
    rng = np.random.default_rng(M + D)
    N, L = 400, 5
    codes = rng.random((M, D), dtype=np.float32)
    data = rng.random((N, D), dtype=np.float32)
    cl = rng.integers(1, L + 1, M).astype(np.int32)
    dl = rng.integers(1, L + 1, N).astype(np.int32)
    for algo in ((1, 4) if M < 2 else (1, 2, 3, 4)):                # lvq2 / lvq3 need two neighbours
        alpha = 0.3 if algo == 4 else 0.05
        out = engine.lvq_training(algo, codes, cl, data, dl, 500, alpha, 1, 0.3, 0.1, rand_seed=2)
        exp, eua = oracle.lvq_train(algo, codes, cl, data, dl, 500, alpha, 1, 0.3, 0.1, order=oracle.shuffle_order(N, 2))
        if algo == 4:
            out, ua = out
            assert_bits_equal(ua, eua, "unit alpha")
        assert_bits_equal(out, exp, "lvq algo %d M=%d D=%d" % (algo, M, D))
