#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref, built by
`make -C oracle ref` from /root/reference).  Run in the build container only:

    python tests/golden/make_golden.py

The reference ships no test vectors of its own, so these files --
outputs of the compiled reference on seeded inputs and on its six ex*.dat demo files --
are what pins both oracle/oracle.c and the CUDA path on the GPU box, where
/root/reference does not exist.  Inputs are stored next to the outputs so that nothing
has to be regenerated at test time.
"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.pyoracle import Reference, REF_BIN, build  # noqa: E402
import datfile  # noqa: E402

REFDIR = "/root/reference"


def save(name, **kw):
    np.savez_compressed(os.path.join(HERE, name), **kw)
    print("wrote", name, len(kw), "arrays")


def search_cases(r):
    rng = np.random.default_rng(20261018)
    out = {}
    shapes = [("lowdim", 96, 5, 400), ("c3like", 300, 64, 256), ("odd", 37, 33, 101),
              ("c4like", 160, 512, 48), ("tiny", 3, 2, 17), ("wide", 50, 130, 64)]
    for name, M, D, N in shapes:
        codes = rng.random((M, D), dtype=np.float32)
        data = rng.random((N, D), dtype=np.float32)
        # quantised variant: forces exact distance ties, plus duplicated code vectors
        qc = (np.round(codes * 4) / 4).astype(np.float32)
        qd = (np.round(data * 4) / 4).astype(np.float32)
        qc[M // 2] = qc[0]
        qc[M - 1] = qc[1]
        mask = (rng.random((N, D)) < 0.25).astype(np.uint8)
        mask[min(3, N - 1)] = 1                      # one all-masked sample
        nf_d = data.copy()
        nf_c = codes.copy()
        nf_d[1, 0] = np.nan
        nf_d[2, D - 1] = np.inf
        nf_d[4 % N, D // 2] = -np.inf
        nf_c[1, D - 1] = np.inf
        nf_c[2, D // 2] = np.nan
        for k in (1, 2, 5, 10):
            if k > 1 and name in ("c4like",) and k == 10:
                continue
            for tag, c, d, m in (("u", codes, data, None), ("q", qc, qd, None),
                                 ("m", qc, qd, mask), ("nf", nf_c, nf_d, None)):
                idx, diff, ret = r.search(c, d, k, m)
                out["%s_%s_k%d_idx" % (name, tag, k)] = idx
                out["%s_%s_k%d_diff" % (name, tag, k)] = diff
                out["%s_%s_k%d_ret" % (name, tag, k)] = ret
        out[name + "_codes"], out[name + "_data"] = codes, data
        out[name + "_qcodes"], out[name + "_qdata"] = qc, qd
        out[name + "_mask"] = mask
        out[name + "_nfcodes"], out[name + "_nfdata"] = nf_c, nf_d
    save("search.npz", **out)


def scalar_cases(r):
    out = {}
    out["shuffle_10_1"] = r.shuffle_order(10, 1)
    out["shuffle_3840_123"] = r.shuffle_order(3840, 123)
    out["shuffle_1962_7"] = r.shuffle_order(1962, 7)
    out["shuffle_40000_3"] = r.shuffle_order(40000, 3)
    g = np.array([(bx, by, tx, ty) for bx in range(5) for by in range(5)
                  for tx in range(5) for ty in range(5)], np.int32)
    out["lattice_args"] = g
    out["hexa"] = np.array([r.hexa_dist(*map(int, a)) for a in g], np.float32)
    out["rect"] = np.array([r.rect_dist(*map(int, a)) for a in g], np.float32)
    a = np.array([(it, ln, al) for it in (0, 1, 17, 999, 54321) for ln in (7, 1000, 10000, 1000000)
                  for al in (0.05, 0.02, 0.3)], np.float64)
    out["alpha_args"] = a
    out["linear"] = np.array([r.linear_alpha(int(i), int(l), float(x)) for i, l, x in a], np.float32)
    out["inverse_t"] = np.array([r.inverse_t_alpha(int(i), int(l), float(x)) for i, l, x in a], np.float32)
    rng = np.random.default_rng(5)
    votes = [rng.integers(1, 5, rng.integers(1, 11)).astype(np.int64) for _ in range(64)]
    out["vote_flat"] = np.concatenate(votes)
    out["vote_len"] = np.array([len(v) for v in votes], np.int32)
    out["vote_head"] = np.array([r.hitlist_vote(v) for v in votes], np.int64)
    save("scalars.npz", **out)


def som_cases(r):
    rng = np.random.default_rng(77)
    out = {}
    xdim, ydim, D, N = 9, 6, 7, 250
    codes = rng.random((xdim * ydim, D), dtype=np.float32)
    data = rng.random((N, D), dtype=np.float32)
    mask = (rng.random((N, D)) < 0.2).astype(np.uint8)
    mask[5] = 1
    weight = rng.integers(0, 4, N).astype(np.int16)
    fixed = np.full((N, 2), -1, np.int16)
    fixed[::9] = [3, 2]
    out.update(codes=codes, data=data, mask=mask, weight=weight, fixed=fixed,
               dims=np.array([xdim, ydim], np.int32))
    for topol in (3, 4):
        for neigh in (1, 2):
            for at in (1, 2):
                for seed in (-1, 11):
                    key = "t%d_n%d_a%d_s%d" % (topol, neigh, at, seed)
                    out[key] = r.som_train(codes, data, xdim, ydim, topol, neigh, 1500, 0.05,
                                           4.0, at, rand_seed=seed)
            key = "t%d_n%d_mwf" % (topol, neigh)
            tr = r.som_train(codes, data, xdim, ydim, topol, neigh, 900, 0.05, 3.0, 1,
                             mask=mask, weight=weight, fixed_xy=fixed)
            out[key] = tr
            for qt in (0, 1):
                out["%s_q%d" % (key, qt)] = np.float32(
                    r.qerror(tr, data, xdim, ydim, topol, neigh, qt, 2.0, mask))
    save("som.npz", **out)


def lvq_cases(r):
    rng = np.random.default_rng(99)
    out = {}
    M, D, N, L = 60, 12, 700, 6
    codes = rng.random((M, D), dtype=np.float32)
    data = rng.random((N, D), dtype=np.float32)
    cl = rng.integers(1, L + 1, M).astype(np.int32)
    dl = rng.integers(1, L + 1, N).astype(np.int32)
    out.update(codes=codes, data=data, code_label=cl, data_label=dl)
    td = tempfile.mkdtemp()
    for algo in (1, 2, 3, 4):
        for seed in (-1, 4):
            for at in (1, 2):
                alpha = 0.3 if algo == 4 else 0.05
                key = "algo%d_s%d_a%d" % (algo, seed, at)
                out[key] = r.lvq_train(algo, codes, cl, data, dl, 4000, alpha, at, 0.3, 0.1,
                                       rand_seed=seed, lra_in=td + "/i.cod",
                                       lra_out=td + "/o.cod")
                if algo == 4:
                    out[key + "_lra"] = np.array(open(td + "/o.lra").read().split())
    shutil.rmtree(td)
    save("lvq.npz", **out)


def run(cmd, cwd):
    p = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if p.returncode:
        raise RuntimeError("%s failed: %s" % (cmd, p.stderr))
    return p.stdout


def demo_cases():
    """C1 / C2 of BASELINE.json through the reference's own command-line programs
    (recipes: reference Makefile:195-212, plus the lvq1/knntest steps BASELINE.json names)."""
    td = tempfile.mkdtemp()
    for f in ("ex.dat", "ex_fts.dat", "ex_ndy.dat", "ex_fdy.dat", "ex1.dat", "ex2.dat"):
        shutil.copy(os.path.join(REFDIR, f), td)
    b = lambda p: os.path.join(REF_BIN, p)  # noqa: E731
    out = {}

    def text(name):
        with open(os.path.join(td, name)) as f:
            return f.read()

    for f in ("ex.dat", "ex_fts.dat", "ex_ndy.dat", "ex_fdy.dat", "ex1.dat", "ex2.dat"):
        out["in_" + f] = np.array(text(f))
    # ---- C1: SOM demo
    run([b("randinit"), "-din", "ex.dat", "-cout", "ex.cod", "-xdim", "12", "-ydim", "8",
         "-topol", "hexa", "-neigh", "bubble", "-rand", "123"], td)
    out["som_init_cod"] = np.array(text("ex.cod"))
    run([b("vsom"), "-din", "ex.dat", "-cin", "ex.cod", "-cout", "ex.cod", "-rlen", "1000",
         "-alpha", "0.05", "-radius", "10"], td)
    out["som_stage1_cod"] = np.array(text("ex.cod"))
    run([b("vsom"), "-din", "ex.dat", "-cin", "ex.cod", "-cout", "ex.cod", "-rlen", "10000",
         "-alpha", "0.02", "-radius", "3"], td)
    out["som_stage2_cod"] = np.array(text("ex.cod"))
    out["som_qerror_stdout"] = np.array(run([b("qerror"), "-din", "ex.dat", "-cin", "ex.cod"], td))
    out["som_qerror1_stdout"] = np.array(run([b("qerror"), "-din", "ex.dat", "-cin", "ex.cod",
                                              "-qetype", "1", "-radius", "2"], td))
    run([b("vcal"), "-din", "ex_fts.dat", "-cin", "ex.cod", "-cout", "ex.cod"], td)
    out["som_vcal_cod"] = np.array(text("ex.cod"))
    run([b("visual"), "-din", "ex_ndy.dat", "-cin", "ex.cod", "-dout", "ex.nvs"], td)
    run([b("visual"), "-din", "ex_fdy.dat", "-cin", "ex.cod", "-dout", "ex.fvs"], td)
    out["som_nvs"] = np.array(text("ex.nvs"))
    out["som_fvs"] = np.array(text("ex.fvs"))
    # variants: -rand order, gaussian, rect, inverse_t
    run([b("randinit"), "-din", "ex.dat", "-cout", "g.cod", "-xdim", "10", "-ydim", "7",
         "-topol", "rect", "-neigh", "gaussian", "-rand", "7"], td)
    out["som_g_init_cod"] = np.array(text("g.cod"))
    run([b("vsom"), "-din", "ex.dat", "-cin", "g.cod", "-cout", "g1.cod", "-rlen", "2000",
         "-alpha", "0.05", "-radius", "5", "-rand", "3", "-alpha_type", "inverse_t"], td)
    out["som_g_stage1_cod"] = np.array(text("g1.cod"))
    # ---- C2: LVQ demo
    run([b("eveninit"), "-din", "ex1.dat", "-cout", "ex1e.cod", "-noc", "200"], td)
    out["lvq_e_cod"] = np.array(text("ex1e.cod"))
    run([b("balance"), "-din", "ex1.dat", "-cin", "ex1e.cod", "-cout", "ex1b.cod"], td)
    out["lvq_b_cod"] = np.array(text("ex1b.cod"))
    run([b("olvq1"), "-din", "ex1.dat", "-cin", "ex1b.cod", "-cout", "ex1o.cod", "-rlen", "5000"], td)
    out["lvq_o_cod"] = np.array(text("ex1o.cod"))
    # NB: lvqtrain.c:248-249 removes ex1o.lra again right after olvq1 wrote it (a reference quirk)
    out["lvq_o_accuracy_stdout"] = np.array(run([b("accuracy"), "-din", "ex2.dat", "-cin", "ex1o.cod"], td))
    run([b("lvq1"), "-din", "ex1.dat", "-cin", "ex1o.cod", "-cout", "ex1l.cod", "-alpha", "0.05",
         "-rlen", "50000"], td)
    out["lvq_l_cod"] = np.array(text("ex1l.cod"))
    out["lvq_l_accuracy_stdout"] = np.array(run([b("accuracy"), "-din", "ex2.dat", "-cin", "ex1l.cod"], td))
    out["lvq_l_knntest_stdout"] = np.array(run([b("knntest"), "-din", "ex2.dat", "-cin", "ex1l.cod",
                                                "-knn", "5"], td))
    run([b("classify"), "-din", "ex2.dat", "-cin", "ex1l.cod", "-dout", "ex2.cls", "-cfout", "ex2.cf"], td)
    out["lvq_l_classify_dout"] = np.array(text("ex2.cls"))
    out["lvq_l_classify_cfout"] = np.array(text("ex2.cf"))
    run([b("lvq2"), "-din", "ex1.dat", "-cin", "ex1o.cod", "-cout", "ex1_2.cod", "-alpha", "0.03",
         "-rlen", "8000", "-win", "0.3"], td)
    out["lvq_2_cod"] = np.array(text("ex1_2.cod"))
    run([b("lvq3"), "-din", "ex1.dat", "-cin", "ex1o.cod", "-cout", "ex1_3.cod", "-alpha", "0.03",
         "-rlen", "8000", "-win", "0.3", "-epsilon", "0.1", "-rand", "5"], td)
    out["lvq_3_cod"] = np.array(text("ex1_3.cod"))
    shutil.rmtree(td)
    save("demo.npz", **out)


if __name__ == "__main__":
    if not os.path.isdir(REFDIR):
        sys.exit("needs /root/reference (build container only)")
    build(ref=True)
    ref = Reference()
    search_cases(ref)
    scalar_cases(ref)
    som_cases(ref)
    lvq_cases(ref)
    demo_cases()
