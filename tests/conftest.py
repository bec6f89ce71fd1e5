import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with -m gpu)")


def bits(a):
    """bit pattern view for exact float comparisons (-0 != +0).  NaNs are canonicalised:
    the payload/sign of a generated NaN is a property of the FPU (x86 SSE gives 0xFFC00000,
    the GPU 0x7FFFFFFF), not of the algorithm."""
    a = np.ascontiguousarray(a)
    if a.dtype != np.float32:
        return a
    b = a.view(np.int32).copy()
    b[np.isnan(a)] = 0x7FC00000
    return b


def assert_bits_equal(a, b, what=""):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if not np.array_equal(bits(a), bits(b)):
        bad = np.argwhere(bits(a) != bits(b))
        raise AssertionError("%s: %d mismatches, first at %s: %r vs %r" % (
            what, len(bad), bad[0], a[tuple(bad[0])], b[tuple(bad[0])]))


@pytest.fixture(scope="session")
def golden():
    class G:
        def __getattr__(self, name):
            return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return G()


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def engine():
    """the CUDA engine; fails (does not skip) when the library is missing"""
    import som_lvq_pak_b200 as b
    b.init(0)
    return b
