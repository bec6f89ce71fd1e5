"""CPU: the C-ABI library loads, exports every symbol include/bmu.h declares, fails loudly
without a GPU, and its pure-C host helpers agree with the oracle / golden vectors."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, assert_bits_equal


def header_symbols():
    text = open(os.path.join(ROOT, "include", "bmu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bmu_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from som_lvq_pak_b200 import _lib
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), "missing export " + s
    assert sorted(_lib.PROTOTYPES) == syms


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import som_lvq_pak_b200 as b
    with pytest.raises(RuntimeError, match="no CUDA device|sm_100a"):
        b.Codebook(np.zeros((4, 3), np.float32))
    with pytest.raises(RuntimeError):
        b.som_training(np.zeros((4, 3), np.float32), np.zeros((5, 3), np.float32), 2, 2, 3, 1, 10, 0.05, 1.0)


def test_product_does_not_reference_oracle():
    """the product never imports, links or names the test oracle (tier rule 3)"""
    pkg = os.path.join(ROOT, "som_lvq_pak_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", "Makefile")):
                text = open(os.path.join(dp, f)).read().lower()
                assert "oracle" not in text and "/root/reference" not in text, (dp, f)


def test_rand_order_matches_reference(golden, oracle):
    import som_lvq_pak_b200 as b
    g = golden.scalars
    for n, s in [(10, 1), (3840, 123), (1962, 7), (40000, 3)]:
        assert np.array_equal(b.rand_order(n, s), g["shuffle_%d_%d" % (n, s)])
    assert np.array_equal(b.rand_order(70000, 99), oracle.shuffle_order(70000, 99))


def test_schedules_match_oracle(oracle):
    import som_lvq_pak_b200 as b
    N = 37
    order = b.rand_order(N, 5)
    w = (np.arange(N) % 4).astype(np.int16)
    for at in (1, 2):
        for length in (100, 1000, 12345):
            s, ta, tr = b.som_schedule(0, length, length, 0.05, 7.0, at, N, order, None)
            assert np.array_equal(s, order[np.arange(length) % N])
            f = oracle.linear_alpha if at == 1 else oracle.inverse_t_alpha
            exp = np.array([f(le, length, 0.05) for le in range(length)], np.float32)
            assert_bits_equal(ta, exp)
            exp_r = np.array([np.float32(1.0 + (np.float64(np.float32(7.0)) - 1.0) *
                                         np.float64(np.float32(length - le)) / np.float64(np.float32(length)))
                              for le in range(length)], np.float32)
            assert_bits_equal(tr, exp_r)
            # chunked == whole
            s2, ta2, tr2 = b.som_schedule(10, 60, length, 0.05, 7.0, at, N, order, None)
            assert np.array_equal(s2, s[10:60]) and np.array_equal(ta2, ta[10:60]) and np.array_equal(tr2, tr[10:60])
            sl, tl = b.lvq_schedule(0, length, length, 0.05, at, N, None)
            assert_bits_equal(tl, exp)
            assert np.array_equal(sl, np.arange(length) % N)
    # weights: 1 - (float)pow(1 - a, w)
    s, ta, _ = b.som_schedule(0, 50, 50, 0.05, 3.0, 1, N, None, w)
    for le in range(50):
        a = np.float32(oracle.linear_alpha(le, 50, 0.05))
        ww = int(w[le % N])
        if ww > 0:
            a = np.float32(1.0 - np.float64(np.float32(np.power(1.0 - np.float64(a), np.float64(ww)))))
        assert ta[le] == a


def test_randinit_codes_matches_reference_randinit(golden):
    """pure host helper (no device): the map `randinit -rand 123 -xdim 12 -ydim 8` wrote for ex.dat
    and `-rand 7 -xdim 10 -ydim 7` (tests/golden/demo.npz), compared through the %g formatting"""
    import numpy as np
    import datfile
    import som_lvq_pak_b200 as b
    g = golden.demo
    data = datfile.parse(str(g["in_ex.dat"]))
    for key, xdim, ydim, seed in (("som_init_cod", 12, 8, 123), ("som_g_init_cod", 10, 7, 7)):
        ref = datfile.parse(str(g[key]))
        got = b.randinit_codes(data.points, xdim, ydim, seed, data.mask)
        assert got.shape == ref.points.shape
        assert [["%g" % v for v in row] for row in got] == [["%g" % v for v in row] for row in ref.points]
