"""GPU: the drop-in boundary against the reference's OWN structs.

glue/_build/bin/* are the reference's unmodified programs (qerror.c, visual.c, vcal.c,
accuracy.c, knntest.c, classify.c, cmatr.c, vsom.c, lvqtrain.c, compiled from /root/reference by
glue/Makefile) linked with bmu_glue.c -- entries_flatten -> bmu_multi_search / bmu_trainer_* ->
entries_scatter on struct entries / data_entry / winner_info / teach_params (lvq_pak.h:73-124,186-204)
-- and libbmu_b200.so.  Both demo recipes (BASELINE.json configs[0] and [1], reference Makefile:195-212)
run through them must reproduce, byte for byte, what the stock binaries produced (tests/golden/demo.npz).
The binaries are built in the build container (they need /root/reference) and travel to the GPU box."""
import os
import subprocess

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

BIN = os.path.join(ROOT, "glue", "_build", "bin")


class Box:
    def __init__(self, tmp_path, demo):
        self.dir = tmp_path
        for f in ("ex.dat", "ex_fts.dat", "ex_ndy.dat", "ex_fdy.dat", "ex1.dat", "ex2.dat"):
            (tmp_path / f).write_text(str(demo["in_" + f]))

    def put(self, name, text):
        (self.dir / name).write_text(str(text))

    def text(self, name):
        return (self.dir / name).read_text()

    def run(self, prog, *args):
        p = subprocess.run([os.path.join(BIN, prog), *args], cwd=self.dir, stdout=subprocess.PIPE,
                           stderr=subprocess.PIPE, text=True)
        assert p.returncode == 0, (prog, args, p.stderr[-2000:])
        return p.stdout


@pytest.fixture()
def box(tmp_path, golden):
    assert os.path.exists(os.path.join(BIN, "qerror")), \
        "glue programs not built (make -C glue in the build container)"
    return Box(tmp_path, golden.demo)


def test_reference_som_programs_on_the_engine(box, golden):
    """C1: the reference's vsom / qerror / vcal / visual, som_training and the winner slot displaced"""
    g = golden.demo
    box.put("ex.cod", g["som_init_cod"])
    box.run("vsom", "-din", "ex.dat", "-cin", "ex.cod", "-cout", "ex.cod", "-rlen", "1000", "-alpha", "0.05",
            "-radius", "10")
    assert box.text("ex.cod") == str(g["som_stage1_cod"])
    box.run("vsom", "-din", "ex.dat", "-cin", "ex.cod", "-cout", "ex.cod", "-rlen", "10000", "-alpha", "0.02",
            "-radius", "3")
    assert box.text("ex.cod") == str(g["som_stage2_cod"])
    assert box.run("qerror", "-din", "ex.dat", "-cin", "ex.cod") == str(g["som_qerror_stdout"])
    # -qetype 1 walks the map around the (batched) winner with the reference's own bubble_qerror
    assert box.run("qerror", "-din", "ex.dat", "-cin", "ex.cod", "-qetype", "1", "-radius", "2") == \
        str(g["som_qerror1_stdout"])
    box.run("vcal", "-din", "ex_fts.dat", "-cin", "ex.cod", "-cout", "ex.cod")
    assert box.text("ex.cod") == str(g["som_vcal_cod"])
    box.run("visual", "-din", "ex_ndy.dat", "-cin", "ex.cod", "-dout", "ex.nvs")
    box.run("visual", "-din", "ex_fdy.dat", "-cin", "ex.cod", "-dout", "ex.fvs")
    assert box.text("ex.nvs") == str(g["som_nvs"])
    assert box.text("ex.fvs") == str(g["som_fvs"])


def test_reference_som_variants_on_the_engine(box, golden):
    """gaussian neighbourhood, -rand order (the reference's own shuffle at load), inverse_t rate"""
    g = golden.demo
    box.put("g.cod", g["som_g_init_cod"])
    box.run("vsom", "-din", "ex.dat", "-cin", "g.cod", "-cout", "g1.cod", "-rlen", "2000", "-alpha", "0.05",
            "-radius", "5", "-rand", "3", "-alpha_type", "inverse_t")
    assert box.text("g1.cod") == str(g["som_g_stage1_cod"])


def test_reference_lvq_programs_on_the_engine(box, golden):
    """C2: the reference's olvq1 / lvq1 / lvq2 / lvq3 (lvqtrain.c), accuracy, knntest, classify, cmatr"""
    g, x = golden.demo, golden.demo_extra
    box.put("ex1b.cod", g["lvq_b_cod"])
    box.run("olvq1", "-din", "ex1.dat", "-cin", "ex1b.cod", "-cout", "ex1o.cod", "-rlen", "5000")
    assert box.text("ex1o.cod") == str(g["lvq_o_cod"])
    assert box.run("accuracy", "-din", "ex2.dat", "-cin", "ex1o.cod") == str(g["lvq_o_accuracy_stdout"])
    box.run("lvq1", "-din", "ex1.dat", "-cin", "ex1o.cod", "-cout", "ex1l.cod", "-alpha", "0.05", "-rlen", "50000")
    assert box.text("ex1l.cod") == str(g["lvq_l_cod"])
    assert box.run("accuracy", "-din", "ex2.dat", "-cin", "ex1l.cod") == str(g["lvq_l_accuracy_stdout"])
    assert box.run("knntest", "-din", "ex2.dat", "-cin", "ex1l.cod", "-knn", "5") == str(g["lvq_l_knntest_stdout"])
    box.run("classify", "-din", "ex2.dat", "-cin", "ex1l.cod", "-dout", "ex2.cls", "-cfout", "ex2.cf")
    assert box.text("ex2.cls") == str(g["lvq_l_classify_dout"])
    assert box.text("ex2.cf") == str(g["lvq_l_classify_cfout"])
    box.run("lvq2", "-din", "ex1.dat", "-cin", "ex1o.cod", "-cout", "ex1_2.cod", "-alpha", "0.03", "-rlen", "8000",
            "-win", "0.3")
    assert box.text("ex1_2.cod") == str(g["lvq_2_cod"])
    box.run("lvq3", "-din", "ex1.dat", "-cin", "ex1o.cod", "-cout", "ex1_3.cod", "-alpha", "0.03", "-rlen", "8000",
            "-win", "0.3", "-epsilon", "0.1", "-rand", "5")
    assert box.text("ex1_3.cod") == str(g["lvq_3_cod"])
    assert box.run("cmatr", "-din", "ex2.dat", "-cin", "ex1l.cod", "-cfout", "cm.cf") == str(x["cmatr_stdout"])
    assert box.text("cm.cf") == str(x["cmatr_cfout"])


def test_reference_buffered_reading_on_the_engine(box, golden):
    """-buffer N (datafile.c:237-344): the reference's own chunked reader feeds the batched winner slot one
    chunk at a time -- same output as the whole-file run"""
    g = golden.demo
    box.put("ex.cod", g["som_stage2_cod"])
    assert box.run("qerror", "-din", "ex.dat", "-cin", "ex.cod", "-buffer", "500") == str(g["som_qerror_stdout"])
    box.put("ex1l.cod", g["lvq_l_cod"])
    assert box.run("accuracy", "-din", "ex2.dat", "-cin", "ex1l.cod", "-buffer", "300") == str(g["lvq_l_accuracy_stdout"])


def test_reference_planes_trajectory_on_the_engine(box, golden, tmp_path):
    """scan_data_traj (planes.c:220-262): the per-sample winner trajectory and the hit density of the
    reference's own `planes`, once with its stock winner function (oracle/_ref/bin, CPU) and once with the
    engine behind the same slot -- every PostScript file must come out identical"""
    g = golden.demo
    stock = os.path.join(ROOT, "oracle", "_ref", "bin", "planes")
    assert os.path.exists(stock), "stock reference binaries not built (make -C oracle ref)"
    box.put("ex.cod", g["som_vcal_cod"])
    ref_dir = tmp_path / "stock"
    ref_dir.mkdir()
    for f in ("ex.dat", "ex.cod"):
        (ref_dir / f).write_text(box.text(f))
    subprocess.run([stock, "-cin", "ex.cod", "-din", "ex.dat", "-plane", "0"], cwd=ref_dir, check=True,
                   stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    box.run("planes", "-cin", "ex.cod", "-din", "ex.dat", "-plane", "0")
    names = sorted(p.name for p in ref_dir.glob("*.eps"))
    assert "ex_tr.eps" in names and len(names) >= 6
    for n in names:
        assert box.text(n) == (ref_dir / n).read_text(), n
