"""GPU parity of K6 (bmu_identical_pairs / bmu_sammon: remove_identicals + sammon_iterate,
sammon.c:83-262): bit-exact positions against the oracle and the reference's golden positions, the
per-sweep mapping error, and the `sammon` program's file, stdout and stderr byte for byte."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, assert_bits_equal

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from make_golden_sammon import CASES, make_case  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "sammon.npz"))
PAK = os.path.join(ROOT, "som_lvq_pak_b200", "host", "bmu_pak")


@pytest.mark.parametrize("name", sorted(CASES))
def test_sammon_golden(engine, oracle, name):
    codes, mask, length, seed = make_case(name)
    keep = engine.remove_identicals(codes, mask)
    assert np.array_equal(keep, oracle.remove_identicals(codes, mask))
    x0, y0 = engine.sammon_init(len(keep), seed)
    km = None if mask is None else mask[keep]
    x, y, err = engine.sammon(codes[keep], length, x0, y0, km, errors=True)
    assert_bits_equal(x, GOLD[name + "_x"], name + " x")
    assert_bits_equal(y, GOLD[name + "_y"], name + " y")
    _, _, oerr = oracle.sammon(codes[keep], length, x0, y0, km, errors=True)
    assert_bits_equal(err, oerr, name + " mapping error")
    x2, y2 = engine.sammon(codes[keep], length, x0, y0, km)        # same sweeps without the error read-back
    assert_bits_equal(x2, x)
    assert_bits_equal(y2, y)


def test_sammon_vs_oracle_larger(engine, oracle):
    """more points than one wave of warps per CTA row, M not a multiple of 32 or 64"""
    rng = np.random.default_rng(12)
    for M, D, length in [(1500, 16, 4), (333, 64, 6), (2, 3, 5), (1, 3, 2)]:
        codes = rng.random((M, D), dtype=np.float32)
        x0, y0 = engine.sammon_init(M, 17)
        x, y = engine.sammon(codes, length, x0, y0)
        ox, oy = oracle.sammon(codes, length, x0, y0)
        assert_bits_equal(x, ox, "M=%d x" % M)
        assert_bits_equal(y, oy, "M=%d y" % M)


def test_identical_pairs(engine, oracle):
    rng = np.random.default_rng(4)
    codes = (rng.integers(0, 2, (400, 6))).astype(np.float32)      # 64 distinct vectors: many zero pairs
    pairs = engine.identical_pairs(codes)
    eq = (codes[:, None, :] == codes[None, :, :]).all(-1)
    want = np.argwhere(np.triu(eq, 1))
    assert np.array_equal(pairs, want)
    assert np.array_equal(engine.remove_identicals(codes), oracle.remove_identicals(codes))
    with pytest.raises(RuntimeError):
        engine.identical_pairs(codes, cap=10)


@pytest.mark.parametrize("key,cod,rlen,seed", [("cli_map", "som_stage2_cod", 100, 5), ("cli_lvq", "lvq_l_cod", 40, 9)])
def test_sammon_program(tmp_path, golden, key, cod, rlen, seed):
    assert os.path.exists(PAK), "host programs not built"
    (tmp_path / "c.cod").write_text(str(golden.demo[cod]))
    p = subprocess.run([PAK, "sammon", "-cin", "c.cod", "-cout", "c.sam", "-rlen", str(rlen), "-rand", str(seed),
                        "-v", "2"], cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr
    assert (tmp_path / "c.sam").read_text() == str(GOLD[key + "_sam"])
    assert p.stdout == str(GOLD[key + "_stdout"])
    assert p.stderr == str(GOLD[key + "_stderr"])
