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


def _big_file(rows, dim, seed=0):
    rng = np.random.default_rng(seed)
    vals = rng.normal(size=(rows, dim)).astype(np.float32)
    lines = ["# generated", "%d" % dim]
    for r in range(rows):
        toks = ["%g" % v for v in vals[r]]
        if r % 17 == 0:
            toks[r % dim] = "x"
        if r % 1001 == 0:
            toks = ["x"] * dim                     # dropped entry (all components masked)
        if r % 5 == 0:
            toks.append("L%d" % (r % 13))
        if r % 29 == 0:
            toks += ["second", "weight=%d" % (r % 7), "fixed=%d,%d" % (r % 9, r % 4)]
        lines.append(" ".join(toks))
        if r % 97 == 0:
            lines.append("")
        if r % 211 == 0:
            lines.append("# a comment in the middle")
    return "\n".join(lines) + "\n"


def test_parallel_loader_is_thread_count_independent(tmp_path):
    """the block-parallel parser (pthreads) must give the same entries, masks and labels as one thread"""
    text = _big_file(90_000, 24)                   # ~20 MB: enough for several 1 MiB blocks
    src = tmp_path / "big.dat"
    src.write_text(text)
    outs = []
    for threads in ("1", "3", "8"):
        dst = tmp_path / ("out%s.dat" % threads)
        env = dict(os.environ, BMU_PAK_THREADS=threads)
        subprocess.run([PAK, "pakcat", "-din", str(src), "-dout", str(dst)], check=True, env=env)
        outs.append(dst.read_bytes())
    assert outs[0] == outs[1] == outs[2]
    kept = [l for l in text.splitlines()[2:] if l and not l.startswith("#") and set(l.split()[:24]) != {"x"}]
    assert outs[0].count(b"\n") == len(kept) + 1


def test_streamed_reader_equals_whole_file_load(tmp_path):
    """-buffer N (datafile.c:237-344) through the streamed reader: chunks of exactly N entries (text blocks of
    64 MB parsed block-parallel, the next one on a helper thread), concatenated = the whole-file load;
    masks, multiple labels and skipped all-masked lines included"""
    text = _big_file(40_000, 24)
    src = tmp_path / "big.dat"
    src.write_text(text)
    whole = tmp_path / "whole.dat"
    subprocess.run([PAK, "pakcat", "-din", str(src), "-dout", str(whole)], check=True)
    for buf in ("1", "777", "5000", "40000", "100000"):
        if buf == "1" and len(text) > 1_000_000:
            continue                                # one entry per chunk: covered on the small file below
        dst = tmp_path / ("b%s.dat" % buf)
        for block in ("", "300000", "1000"):        # text blocks of 64 MB (default), 300 KB, 1000 bytes (~4 lines)
            env = dict(os.environ, BMU_PAK_TEXT_BLOCK=block) if block else dict(os.environ)
            if block == "1000" and buf != "777":
                continue
            subprocess.run([PAK, "pakcat", "-din", str(src), "-dout", str(dst), "-buffer", buf], check=True,
                           stderr=subprocess.PIPE, env=env)
            assert dst.read_bytes() == whole.read_bytes(), (buf, block)
    small = tmp_path / "small.dat"
    small.write_text("3\n# c\n1 2 3 a b\nx x x dropped\n4 x 6 c\n\n7 8 9\n")
    outs = []
    for extra in ((), ("-buffer", "1"), ("-buffer", "2")):
        dst = tmp_path / "s.out"
        subprocess.run([PAK, "pakcat", "-din", str(small), "-dout", str(dst), *extra], check=True, stderr=subprocess.PIPE)
        outs.append(dst.read_text())
    assert outs[0] == outs[1] == outs[2] == "3\n1 2 3 a b \n4 x 6 c \n7 8 9 \n"


def test_binary_cache_of_parsed_files(tmp_path):
    """$BMU_PAK_CACHE=1: `<file>.bmuc` next to the file after the first parse; a second load comes from the cache and
    gives the same entries, masks and labels (label table order included: two files loaded in a row); touching the
    source invalidates it; other loader flags do not share a cache"""
    text = _big_file(20_000, 12)
    src = tmp_path / "c.dat"
    src.write_text(text)
    env = dict(os.environ, BMU_PAK_CACHE="1")
    plain = tmp_path / "plain.out"
    subprocess.run([PAK, "pakcat", "-din", str(src), "-dout", str(plain)], check=True)
    assert not (tmp_path / "c.dat.bmuc").exists()
    outs = []
    for i in range(2):
        dst = tmp_path / ("o%d.out" % i)
        subprocess.run([PAK, "pakcat", "-din", str(src), "-dout", str(dst)], check=True, env=env)
        assert (tmp_path / "c.dat.bmuc").exists()
        outs.append(dst.read_bytes())
    assert outs[0] == outs[1] == plain.read_bytes()
    # -noskip keeps the all-masked lines: another flag set, the cache made above must not be used
    noskip_plain, noskip_cached = tmp_path / "n0.out", tmp_path / "n1.out"
    subprocess.run([PAK, "pakcat", "-din", str(src), "-dout", str(noskip_plain), "-noskip"], check=True)
    subprocess.run([PAK, "pakcat", "-din", str(src), "-dout", str(noskip_cached), "-noskip"], check=True, env=env)
    assert noskip_plain.read_bytes() == noskip_cached.read_bytes() != plain.read_bytes()
    # a changed source (other size and time) is parsed again
    src.write_text(text + "1 2 3 4 5 6 7 8 9 10 11 12 newlabel\n")
    dst = tmp_path / "o2.out"
    subprocess.run([PAK, "pakcat", "-din", str(src), "-dout", str(dst), "-noskip"], check=True, env=env)
    assert dst.read_bytes().endswith(b"1 2 3 4 5 6 7 8 9 10 11 12 newlabel \n")


def test_parallel_loader_reports_the_file_line(tmp_path):
    text = _big_file(60_000, 24).splitlines()
    bad_line = 41_234
    text[bad_line - 1] = "oops " + text[bad_line - 1]
    src = tmp_path / "bad.dat"
    src.write_text("\n".join(text) + "\n")
    p = subprocess.run([PAK, "pakcat", "-din", str(src), "-dout", str(tmp_path / "o")], stderr=subprocess.PIPE,
                       text=True, env=dict(os.environ, BMU_PAK_THREADS="6"))
    assert p.returncode != 0
    assert "on line %d, component 0" % bad_line in p.stderr


def test_fast_decimal_parser_equals_strtof(tmp_path):
    """every token form a .dat file can hold must convert exactly like libc's strtof (= scanf %f)"""
    import ctypes
    libc = ctypes.CDLL("libc.so.6")
    libc.strtof.restype = ctypes.c_float
    libc.strtof.argtypes = [ctypes.c_char_p, ctypes.c_void_p]
    rng = np.random.default_rng(11)
    toks = []
    for v in rng.normal(size=20000):
        toks += ["%g" % v, "%.9g" % v, "%.3f" % (v * 1000), "%e" % (v * 1e-20), "%.17g" % v]
    toks += ["%d" % i for i in rng.integers(-10**9, 10**9, 5000)]
    toks += ["0", "-0", "+.5", "5.", "1e5", "1E-5", "1e", "1e+", "3.4028235e38", "3.5e38", "1e-45", "1.17549435e-38",
             "1.4e-45", "123456789012345678901234567890", "0.000000000000000000000000000001", "16777217", "16777216.5",
             "8388608.5", "8388609.5", "0.1", "0.3", "1.0000001192092896", "1.00000017881393432617187500",
             "inf", "-inf", "nan", "0x1.8p1", "1.5abc", "7e2x"]
    # float rounding midpoints: (2k+1) * 2^-1 around 2^24, written exactly and just beside
    for k in range(200):
        m = (1 << 24) + 2 * k + 1
        toks += ["%d.0" % m, "%d.5" % (m >> 1), "%.1f" % (m / 2 + 1e-7), "%d" % (m * 4), "%de-1" % (m * 5)]
    dim = 8
    toks = toks[:len(toks) // dim * dim]
    rows = [toks[i:i + dim] for i in range(0, len(toks), dim)]
    src = tmp_path / "tok.dat"
    src.write_text("%d\n" % dim + "\n".join(" ".join(r) for r in rows) + "\n")
    raw = tmp_path / "tok.f32"
    subprocess.run([PAK, "pakstat", "-din", str(src), "-rawout", str(raw)], check=True, stdout=subprocess.PIPE)
    got = np.fromfile(raw, dtype=np.float32)
    exp = np.array([libc.strtof(t.encode(), None) for t in toks], dtype=np.float32)
    assert got.shape == exp.shape
    same = (got.view(np.int32) == exp.view(np.int32)) | (np.isnan(got) & np.isnan(exp))
    assert same.all(), [(toks[i], got[i], exp[i]) for i in np.nonzero(~same)[0][:5]]


def test_randinit_program(tmp_path, demo):
    """randinit needs no GPU: ex.dat -> the reference's `randinit -rand 123` map, comment line included"""
    (tmp_path / "ex.dat").write_text(str(demo["in_ex.dat"]))
    subprocess.run([PAK, "randinit", "-din", "ex.dat", "-cout", "ex.cod", "-xdim", "12", "-ydim", "8", "-topol", "hexa",
                    "-neigh", "bubble", "-rand", "123"], check=True, cwd=tmp_path)
    assert (tmp_path / "ex.cod").read_text() == str(demo["som_init_cod"])
    subprocess.run([PAK, "mapinit", "-init", "rand", "-din", "ex.dat", "-cout", "g.cod", "-xdim", "10", "-ydim", "7",
                    "-topol", "rect", "-neigh", "gaussian", "-rand", "7"], check=True, cwd=tmp_path)
    assert (tmp_path / "g.cod").read_text() == str(demo["som_g_init_cod"])



def test_compressed_and_piped_names(tmp_path, demo):
    """open_file (fileio.c:60-190): names ending in .gz go through gzip, names starting with '|' are commands"""
    import shutil
    if not shutil.which("gzip"):
        pytest.skip("no gzip here")
    text = str(demo["in_ex_fts.dat"])
    (tmp_path / "a.dat").write_text(text)
    subprocess.run([PAK, "pakcat", "-din", "a.dat", "-dout", "plain.dat"], check=True, cwd=tmp_path)
    subprocess.run([PAK, "pakcat", "-din", "a.dat", "-dout", "b.dat.gz"], check=True, cwd=tmp_path)
    subprocess.run([PAK, "pakcat", "-din", "b.dat.gz", "-dout", "c.dat"], check=True, cwd=tmp_path)
    assert (tmp_path / "c.dat").read_text() == (tmp_path / "plain.dat").read_text()
    subprocess.run([PAK, "pakcat", "-din", "|cat a.dat", "-dout", "|cat > d.dat"], check=True, cwd=tmp_path)
    assert (tmp_path / "d.dat").read_text() == (tmp_path / "plain.dat").read_text()


def test_lininit_program(tmp_path, demo, golden):
    """lininit / mapinit -init lin (som_rout.c:211-429) is host arithmetic: byte-identical maps, also from
    data with masked components"""
    x = golden.demo_extra
    (tmp_path / "ex.dat").write_text(str(demo["in_ex.dat"]))
    (tmp_path / "ex_fts.dat").write_text(str(demo["in_ex_fts.dat"]))
    (tmp_path / "masked.dat").write_text(str(x["lininit3_in"]))
    subprocess.run([PAK, "lininit", "-din", "ex.dat", "-cout", "lin.cod", "-xdim", "12", "-ydim", "8", "-topol", "hexa",
                    "-neigh", "bubble", "-rand", "123"], check=True, cwd=tmp_path)
    assert (tmp_path / "lin.cod").read_text() == str(x["lininit_cod"])
    subprocess.run([PAK, "mapinit", "-init", "lin", "-din", "ex_fts.dat", "-cout", "lin2.cod", "-xdim", "5", "-ydim", "9",
                    "-topol", "rect", "-neigh", "gaussian", "-rand", "7"], check=True, cwd=tmp_path)
    assert (tmp_path / "lin2.cod").read_text() == str(x["lininit2_cod"])
    subprocess.run([PAK, "lininit", "-din", "masked.dat", "-cout", "lin3.cod", "-xdim", "4", "-ydim", "3", "-topol", "hexa",
                    "-neigh", "bubble", "-rand", "11"], check=True, cwd=tmp_path)
    assert (tmp_path / "lin3.cod").read_text() == str(x["lininit3_cod"])
