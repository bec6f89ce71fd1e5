#!/usr/bin/env python
"""Golden outputs of three more programs whose loops sit on the BMU path -- cmatr, setlabel, elimin
 -- produced by the UNMODIFIED reference binaries in oracle/_ref/bin from the inputs
already stored in demo.npz.  Run in the build container (needs /root/reference for `make -C oracle ref`):
    python tests/golden/make_golden_extra.py        ->  tests/golden/demo_extra.npz"""
import os
import re
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "bin")


def run(cmd, cwd):
    p = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if p.returncode:
        raise RuntimeError("%s failed: %s" % (cmd, p.stderr))
    return p.stdout


def main():
    if not os.path.isdir(REF_BIN):
        sys.exit("build the reference first: make -C oracle ref (needs /root/reference)")
    demo = np.load(os.path.join(HERE, "demo.npz"))
    td = tempfile.mkdtemp()
    for f in ("ex1.dat", "ex2.dat"):
        open(os.path.join(td, f), "w").write(str(demo["in_" + f]))
    open(os.path.join(td, "ex1l.cod"), "w").write(str(demo["lvq_l_cod"]))
    b = lambda p: os.path.join(REF_BIN, p)  # noqa: E731
    out = {}
    out["cmatr_stdout"] = np.array(run([b("cmatr"), "-din", "ex2.dat", "-cin", "ex1l.cod", "-cfout", "cm.cf"], td))
    out["cmatr_cfout"] = np.array(open(os.path.join(td, "cm.cf")).read())
    run([b("setlabel"), "-din", "ex1.dat", "-cin", "ex1l.cod", "-cout", "sl.cod", "-knn", "5"], td)
    out["setlabel_cod"] = np.array(open(os.path.join(td, "sl.cod")).read())
    run([b("setlabel"), "-din", "ex2.dat", "-cin", "ex1l.cod", "-cout", "sl3.cod", "-knn", "3"], td)
    out["setlabel3_cod"] = np.array(open(os.path.join(td, "sl3.cod")).read())
    run([b("elimin"), "-din", "ex1.dat", "-cout", "el.cod", "-knn", "5"], td)
    out["elimin_cod"] = np.array(open(os.path.join(td, "el.cod")).read())
    run([b("elimin"), "-din", "ex2.dat", "-cout", "el10.cod", "-knn", "10"], td)
    out["elimin10_cod"] = np.array(open(os.path.join(td, "el10.cod")).read())
    # mindist on two codebooks, propinit (eveninit's output is demo.npz's lvq_e_cod)
    open(os.path.join(td, "ex1b.cod"), "w").write(str(demo["lvq_b_cod"]))
    out["mindist_b_stdout"] = np.array(run([b("mindist"), "-cin", "ex1b.cod"], td))
    out["mindist_l_stdout"] = np.array(run([b("mindist"), "-cin", "ex1l.cod"], td))
    run([b("propinit"), "-din", "ex1.dat", "-cout", "p.cod", "-noc", "150", "-knn", "3"], td)
    out["propinit_cod"] = np.array(open(os.path.join(td, "p.cod")).read())
    run([b("eveninit"), "-din", "ex2.dat", "-cout", "e2.cod", "-noc", "317"], td)
    out["eveninit2_cod"] = np.array(open(os.path.join(td, "e2.cod")).read())
    # lininit (mapinit -init lin): map on the plane of the two largest eigenvectors (som_rout.c:211-429)
    for f in ("ex.dat", "ex_fts.dat", "ex_ndy.dat"):
        open(os.path.join(td, f), "w").write(str(demo["in_" + f]))
    run([b("lininit"), "-din", "ex.dat", "-cout", "lin.cod", "-xdim", "12", "-ydim", "8", "-topol", "hexa", "-neigh", "bubble",
         "-rand", "123"], td)
    out["lininit_cod"] = np.array(open(os.path.join(td, "lin.cod")).read())
    run([b("mapinit"), "-init", "lin", "-din", "ex_fts.dat", "-cout", "lin2.cod", "-xdim", "5", "-ydim", "9", "-topol", "rect",
         "-neigh", "gaussian", "-rand", "7"], td)
    out["lininit2_cod"] = np.array(open(os.path.join(td, "lin2.cod")).read())
    lines = str(demo["in_ex_fts.dat"]).splitlines()
    for r, c in ((3, 0), (3, 2), (10, 4), (57, 1), (100, 3), (101, 3), (200, 0)):    # a few masked components
        t = lines[r].split()
        t[c] = "x"
        lines[r] = " ".join(t)
    open(os.path.join(td, "masked.dat"), "w").write("\n".join(lines) + "\n")
    run([b("lininit"), "-din", "masked.dat", "-cout", "lin3.cod", "-xdim", "4", "-ydim", "3", "-topol", "hexa", "-neigh", "bubble",
         "-rand", "11"], td)
    out["lininit3_in"] = np.array(open(os.path.join(td, "masked.dat")).read())
    out["lininit3_cod"] = np.array(open(os.path.join(td, "lin3.cod")).read())
    # balance (config 2 of BASELINE.json): the demo step, and a second codebook / neighbour count
    open(os.path.join(td, "ex1e.cod"), "w").write(str(demo["lvq_e_cod"]))
    out["balance_stdout"] = np.array(run([b("balance"), "-din", "ex1.dat", "-cin", "ex1e.cod", "-cout", "bal.cod"], td))
    assert open(os.path.join(td, "bal.cod")).read() == str(demo["lvq_b_cod"])
    out["balance_lra"] = np.array(open(os.path.join(td, "bal.lra")).read())
    out["balance2_stdout"] = np.array(run([b("balance"), "-din", "ex2.dat", "-cin", "e2.cod", "-cout", "bal2.cod", "-knn", "3"], td))
    out["balance2_cod"] = np.array(open(os.path.join(td, "bal2.cod")).read())
    out["balance2_lra"] = np.array(open(os.path.join(td, "bal2.lra")).read())
    # snapshots (som_rout.c:650-658, lvq_pak.c:663-764): one file per snapshot, and -snaptype keepopen
    open(os.path.join(td, "ex.dat"), "w").write(str(demo["in_ex.dat"]))
    open(os.path.join(td, "ex.cod"), "w").write(str(demo["som_init_cod"]))
    run([b("vsom"), "-din", "ex.dat", "-cin", "ex.cod", "-cout", "sn.cod", "-rlen", "1000", "-alpha", "0.05",
         "-radius", "10", "-snapinterval", "300", "-snapfile", "snap_%d.cod"], td)
    for it in (300, 600, 900):
        out["snap_%d" % it] = np.array(open(os.path.join(td, "snap_%d.cod" % it)).read())
    out["snap_final"] = np.array(open(os.path.join(td, "sn.cod")).read())
    open(os.path.join(td, "ex1o.cod"), "w").write(str(demo["lvq_o_cod"]))
    run([b("lvq1"), "-din", "ex1.dat", "-cin", "ex1o.cod", "-cout", "l.cod", "-alpha", "0.05", "-rlen", "2500",
         "-snapinterval", "1000", "-snapfile", "lsnap.txt", "-snaptype", "keepopen"], td)
    out["lvq_snap_keepopen"] = np.array(open(os.path.join(td, "lsnap.txt")).read())
    out["lvq_snap_final"] = np.array(open(os.path.join(td, "l.cod")).read())
    # vfind: 4 trials of a 6x4 map on ex.dat (answers on stdin, vfind.c:138-185), both qerror types
    open(os.path.join(td, "ex.dat"), "w").write(str(demo["in_ex.dat"]))
    for tag, extra in (("vfind", []), ("vfind_q1", ["-qetype", "1", "-alpha_type", "inverse_t"])):
        answers = "\n".join(["4", "ex.dat", "ex.dat", tag + ".cod", "hexa", "bubble" if tag == "vfind" else "gaussian",
                             "6", "4", "300", "0.05", "4", "600", "0.02", "2"]) + "\n"
        p = subprocess.run([b("vfind")] + extra, cwd=td, input=answers, stdout=subprocess.PIPE,
                           stderr=subprocess.PIPE, text=True)
        if p.returncode:
            raise RuntimeError("vfind failed: " + p.stderr)
        out[tag + "_cod"] = np.array(open(os.path.join(td, tag + ".cod")).read())
        out[tag + "_trials"] = np.array("".join(re.findall(r"(?m)^ *\d+: [0-9.]+\n", p.stderr.replace("\r", "\n"))))
        out[tag + "_summary"] = np.array(p.stdout.splitlines()[-1] + "\n")
    shutil.rmtree(td)
    np.savez_compressed(os.path.join(HERE, "demo_extra.npz"), **out)
    print({k: len(str(v)) for k, v in out.items()})


if __name__ == "__main__":
    main()
