"""GPU parity of the quantization-error consumers (find_qerror, som_rout.c:678-731, and the
neighbourhood-weighted find_qerror2, som_rout.c:734-891) through the C ABI, against the golden
values produced by the unmodified reference and against the CPU oracle on seeded inputs.
Bar: bit-exact float for qetype 0 and for bubble qetype 1; gaussian qetype 1 within 1e-6
relative (libm's double exp() is not bit-portable; north_star) -- identical in practice."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GAUSS_RTOL = 1e-6


def same_float(a, b):
    return np.float32(a).view(np.int32) == np.float32(b).view(np.int32)


@pytest.mark.parametrize("topol", [3, 4])
@pytest.mark.parametrize("neigh", [1, 2])
def test_qerror_golden(engine, golden, topol, neigh):
    """maps trained by the reference with masks/weights/fixed points, then both qerror types"""
    g = golden.som
    xdim, ydim = map(int, g["dims"])
    key = "t%d_n%d_mwf" % (topol, neigh)
    codes = g[key]
    q0 = engine.find_qerror(codes, g["data"], g["mask"])
    assert same_float(q0, g[key + "_q0"]), (q0, g[key + "_q0"])
    q1, _ = engine.find_qerror2(codes, g["data"], xdim, ydim, topol, neigh, 2.0, g["mask"])
    if neigh == 1:
        assert same_float(q1, g[key + "_q1"]), (q1, g[key + "_q1"])
    else:
        np.testing.assert_allclose(q1, g[key + "_q1"], rtol=GAUSS_RTOL)


@pytest.mark.parametrize("topol,neigh,radius", [(3, 1, 1.0), (3, 1, 3.5), (4, 1, 2.0), (3, 2, 2.0), (4, 2, 5.0)])
def test_qerror2_vs_oracle(engine, oracle, topol, neigh, radius):
    rng = np.random.default_rng(17 + topol + neigh)
    xdim, ydim, D, N = 13, 9, 7, 700
    codes = rng.random((xdim * ydim, D), dtype=np.float32)
    data = rng.random((N, D), dtype=np.float32)
    mask = (rng.random((N, D)) < 0.2).astype(np.uint8)
    mask[5] = 1                                           # an all-masked sample: skipped by both
    for mk in (None, mask):
        exp = oracle.qerror(codes, data, xdim, ydim, topol, neigh, 1, radius, mk)
        got, per = engine.find_qerror2(codes, data, xdim, ydim, topol, neigh, radius, mk)
        if neigh == 1:
            assert same_float(got, exp), (got, exp)
        else:
            np.testing.assert_allclose(got, exp, rtol=GAUSS_RTOL)
        if mk is not None:
            assert per[5] == 0.0


def test_qerror2_larger_map_property(engine):
    """size-independent properties on a map too large for the oracle: a bubble of radius 0 is
    exactly sum (sqrt(diff))^2 over the winners, and the error grows with the radius"""
    rng = np.random.default_rng(3)
    xdim, ydim, D, N = 64, 48, 32, 4000
    codes = rng.random((xdim * ydim, D), dtype=np.float32)
    data = rng.random((N, D), dtype=np.float32)
    idx, diff, nf = engine.find_winner_euc(codes, data)
    _, per0 = engine.find_qerror2(codes, data, xdim, ydim, 3, 1, 0.0)
    d = np.sqrt(diff[:, 0].astype(np.float64)).astype(np.float32)
    assert np.array_equal(per0.view(np.int32), (d * d).view(np.int32))
    _, per2 = engine.find_qerror2(codes, data, xdim, ydim, 3, 1, 2.0)
    assert (per2 >= per0).all() and (per2 > per0).any()


def test_qerror2_many_chunks(engine, oracle, monkeypatch):
    """bmu_qerror2's double-buffered chunk loop (copy of chunk c+1 || search + weighted pass of chunk c || values of
    chunk c-1 back): seven chunks with masks, the same per-sample values and float sum as one chunk / the oracle"""
    rng = np.random.default_rng(23)
    xdim, ydim, D, N = 13, 9, 7, 700
    codes = rng.random((xdim * ydim, D), dtype=np.float32)
    data = rng.random((N, D), dtype=np.float32)
    mask = (rng.random((N, D)) < 0.2).astype(np.uint8)
    whole, per_whole = engine.find_qerror2(codes, data, xdim, ydim, 3, 1, 2.5, mask)
    monkeypatch.setenv("SOMLVQ_CHUNK_ROWS", "101")
    got, per = engine.find_qerror2(codes, data, xdim, ydim, 3, 1, 2.5, mask)
    assert np.array_equal(per.view(np.int32), per_whole.view(np.int32)) and same_float(got, whole)
    assert same_float(got, oracle.qerror(codes, data, xdim, ydim, 3, 1, 1, 2.5, mask))
