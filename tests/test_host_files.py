"""CPU: the file layer of the C host (som_lvq_pak_b200/host/entries.c) -- header and entry
grammar, masks, multiple labels, `%g` output -- through the `pakcat` program (load + save, no
GPU call).  Files written by the unmodified reference (tests/golden/demo.npz) must come back
byte for byte; raw data files must reach a fixed point after one pass."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

PAK = os.path.join(ROOT, "som_lvq_pak_b200", "host", "bmu_pak")

WRITTEN_BY_REFERENCE = ["som_init_cod", "som_stage2_cod", "som_vcal_cod", "som_nvs", "som_g_stage1_cod",
                        "lvq_e_cod", "lvq_b_cod", "lvq_l_cod", "lvq_l_classify_dout"]
RAW_INPUTS = ["in_ex.dat", "in_ex_fts.dat", "in_ex_ndy.dat", "in_ex_fdy.dat", "in_ex1.dat"]


def pakcat(tmp_path, text, extra=()):
    src, dst = tmp_path / "in.txt", tmp_path / "out.txt"
    src.write_text(text)
    subprocess.run([PAK, "pakcat", "-din", str(src), "-dout", str(dst), *extra], check=True, cwd=tmp_path)
    return dst.read_text()


@pytest.fixture(scope="module")
def demo(golden):
    if not os.path.exists(PAK):
        pytest.fail("host programs not built (python -c 'import __graft_entry__ as g; g.build()')")
    return golden.demo


@pytest.mark.parametrize("key", WRITTEN_BY_REFERENCE)
def test_reference_written_files_round_trip(tmp_path, demo, key):
    text = str(demo[key])
    # comment lines (randinit's "# random seed") are not kept by save_entries either
    kept = "".join(l + "\n" for l in text.splitlines() if not l.startswith("#"))
    assert pakcat(tmp_path, text) == kept


@pytest.mark.parametrize("key", RAW_INPUTS)
def test_raw_inputs_reach_fixed_point(tmp_path, demo, key):
    once = pakcat(tmp_path, str(demo[key]))
    assert pakcat(tmp_path, once) == once
    # same number of entries as non-comment, non-empty lines after the header
    lines = [l for l in str(demo[key]).splitlines() if l.strip() and not l.startswith("#")]
    assert len(once.splitlines()) == len(lines)


def test_grammar_details(tmp_path):
    text = ("# comment first\n3 hexa 2 1 bubble\n# another\n"
            "1 2.5e0 x A B weight=3 fixed=1,0\n"
            "\n"
            "x x x dropped\n"
            "0.1\t-7 1e10\r\n")
    out = pakcat(tmp_path, text)
    assert out == "3 hexa 2 1 bubble\n1 2.5 x A B \n0.1 -7 1e+10 \n"
    keep = pakcat(tmp_path, text, ["-noskip"])
    assert keep == "3 hexa 2 1 bubble\n1 2.5 x A B \nx x x dropped \n0.1 -7 1e+10 \n"
    alt = pakcat(tmp_path, "2\n1 NA\n", ["-mask_str", "NA"])
    assert alt == "2\n1 NA \n"
