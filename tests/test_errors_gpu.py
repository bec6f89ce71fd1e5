"""GPU: the error convention of the C ABI: 0 = OK, a BMU_ERR_* code and a message in
bmu_last_error() otherwise; empty inputs are not errors; nothing is written on failure."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_argument_errors_and_empty_inputs(engine):
    from som_lvq_pak_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(0)
    codes = rng.random((50, 6), dtype=np.float32)
    data = rng.random((20, 6), dtype=np.float32)
    cb = engine.Codebook(codes)
    idx = np.full((20, 3), 77, np.int32)
    diff = np.full((20, 3), 5.0, np.float32)
    nf = np.full(20, 9, np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    # k out of range (1..BMU_KMAX = 16)
    for k in (0, -1, 17):
        rc = lib.bmu_search(cb._h, p(data), None, 20, k, p(idx), p(diff), p(nf))
        assert rc == 2, (k, rc)                                     # BMU_ERR_ARG
        assert lib.bmu_last_error().decode() != ""
    assert (idx == 77).all() and (diff == 5.0).all() and (nf == 9).all()
    # NULL pointers, negative N
    assert lib.bmu_search(None, p(data), None, 20, 1, p(idx), p(diff), p(nf)) == 2
    assert lib.bmu_search(cb._h, None, None, 20, 1, p(idx), p(diff), p(nf)) == 2
    assert lib.bmu_search(cb._h, p(data), None, -5, 1, p(idx), p(diff), p(nf)) == 2
    # N = 0 is fine and writes nothing
    assert lib.bmu_search(cb._h, p(data), None, 0, 1, p(idx), p(diff), p(nf)) == 0
    assert (idx == 77).all()
    cb.close()
    # codebook creation
    lib.bmu_codebook_create.restype = C.c_void_p
    assert not lib.bmu_codebook_create(p(codes), 0, 6)
    assert not lib.bmu_codebook_create(p(codes), 50, 0)
    assert not lib.bmu_codebook_create(None, 50, 6)
    # training: map size must match the codebook, topology / neighbourhood codes are checked
    with pytest.raises(RuntimeError):
        engine.som_training(codes, data, 7, 7, 3, 1, 10, 0.05, 2.0, 1)
    with pytest.raises(RuntimeError):
        engine.som_training(codes, data, 10, 5, 9, 1, 10, 0.05, 2.0, 1)
    with pytest.raises(RuntimeError):
        engine.find_qerror2(codes, data, 7, 7, 3, 1, 2.0)
    # zero training steps: the codebook comes back unchanged
    out = engine.som_training(codes, data, 10, 5, 3, 1, 0, 0.05, 2.0, 1)
    assert np.array_equal(out, codes)
