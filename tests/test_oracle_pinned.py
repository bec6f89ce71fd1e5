"""CPU: pin oracle/oracle.c (our restatement) against the golden vectors generated from the
unmodified reference (tests/golden/make_golden.py) and, when it is present in this container,
against the compiled reference itself (oracle/_ref/libref_driver.so)."""
import numpy as np
import pytest

from conftest import assert_bits_equal
from oracle.pyoracle import Reference

SHAPES = ["lowdim", "c3like", "odd", "c4like", "tiny", "wide"]
TAGS = {"u": ("codes", "data", None), "q": ("qcodes", "qdata", None),
        "m": ("qcodes", "qdata", "mask"), "nf": ("nfcodes", "nfdata", None)}


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("k", [1, 2, 5, 10])
def test_search_golden(oracle, golden, shape, k):
    g = golden.search
    for tag, (c, d, m) in TAGS.items():
        key = "%s_%s_k%d" % (shape, tag, k)
        if key + "_idx" not in g:
            continue
        idx, diff, ret = oracle.search(g[shape + "_" + c], g[shape + "_" + d], k,
                                       None if m is None else g[shape + "_" + m])
        assert_bits_equal(idx, g[key + "_idx"], key + " idx")
        assert_bits_equal(diff, g[key + "_diff"], key + " diff")
        assert_bits_equal(ret, g[key + "_ret"], key + " ret")


def test_scalars_golden(oracle, golden):
    g = golden.scalars
    for n, s in [(10, 1), (3840, 123), (1962, 7), (40000, 3)]:
        assert np.array_equal(oracle.shuffle_order(n, s), g["shuffle_%d_%d" % (n, s)])
    for a, h, r in zip(g["lattice_args"], g["hexa"], g["rect"]):
        assert np.float32(oracle.hexa_dist(*map(int, a))) == h
        assert np.float32(oracle.rect_dist(*map(int, a))) == r
    for (it, ln, al), lin, inv in zip(g["alpha_args"], g["linear"], g["inverse_t"]):
        assert np.float32(oracle.linear_alpha(int(it), int(ln), float(al))) == lin
        assert np.float32(oracle.inverse_t_alpha(int(it), int(ln), float(al))) == inv
    off = 0
    for ln, head in zip(g["vote_len"], g["vote_head"]):
        assert oracle.hitlist_vote(g["vote_flat"][off:off + ln]) == head
        off += ln


def test_som_golden(oracle, golden):
    g = golden.som
    xdim, ydim = map(int, g["dims"])
    N = g["data"].shape[0]
    for topol in (3, 4):
        for neigh in (1, 2):
            for at in (1, 2):
                for seed in (-1, 11):
                    order = None if seed < 0 else oracle.shuffle_order(N, seed)
                    out = oracle.som_train(g["codes"], g["data"], xdim, ydim, topol, neigh, 1500,
                                           0.05, 4.0, at, order=order)
                    assert_bits_equal(out, g["t%d_n%d_a%d_s%d" % (topol, neigh, at, seed)])
            key = "t%d_n%d_mwf" % (topol, neigh)
            out = oracle.som_train(g["codes"], g["data"], xdim, ydim, topol, neigh, 900, 0.05, 3.0,
                                   1, mask=g["mask"], weight=g["weight"], fixed_xy=g["fixed"])
            assert_bits_equal(out, g[key])
            for qt in (0, 1):
                q = oracle.qerror(out, g["data"], xdim, ydim, topol, neigh, qt, 2.0, g["mask"])
                assert np.float32(q) == g["%s_q%d" % (key, qt)]


def test_lvq_golden(oracle, golden):
    g = golden.lvq
    N = g["data"].shape[0]
    for algo in (1, 2, 3, 4):
        for seed in (-1, 4):
            for at in (1, 2):
                alpha = 0.3 if algo == 4 else 0.05
                order = None if seed < 0 else oracle.shuffle_order(N, seed)
                out, ua = oracle.lvq_train(algo, g["codes"], g["code_label"], g["data"],
                                           g["data_label"], 4000, alpha, at, 0.3, 0.1, order=order)
                key = "algo%d_s%d_a%d" % (algo, seed, at)
                assert_bits_equal(out, g[key], key)
                if algo == 4:
                    assert ["%g" % v for v in ua] == list(g[key + "_lra"])


@pytest.mark.skipif(not Reference.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_against_compiled_reference(oracle):
    ref = Reference()
    rng = np.random.default_rng(123)
    for (M, D, N, k) in [(50, 5, 200, 1), (50, 5, 200, 5), (200, 20, 100, 2), (7, 3, 50, 10)]:
        codes = (np.round(rng.random((M, D)) * 8) / 8).astype(np.float32)
        data = (np.round(rng.random((N, D)) * 8) / 8).astype(np.float32)
        mask = (rng.random((N, D)) < 0.3).astype(np.uint8)
        data[1, 0] = np.nan
        codes[2, D - 1] = np.inf
        a, b = oracle.search(codes, data, k, mask), ref.search(codes, data, k, mask)
        for x, y in zip(a, b):
            assert_bits_equal(x, y)
    xdim, ydim, D, N = 6, 5, 4, 120
    codes = rng.random((xdim * ydim, D), dtype=np.float32)
    data = rng.random((N, D), dtype=np.float32)
    for neigh in (1, 2):
        a = oracle.som_train(codes, data, xdim, ydim, 3, neigh, 500, 0.05, 3.0, 1,
                             order=oracle.shuffle_order(N, 9))
        b = ref.som_train(codes, data, xdim, ydim, 3, neigh, 500, 0.05, 3.0, 1, rand_seed=9)
        assert_bits_equal(a, b)
