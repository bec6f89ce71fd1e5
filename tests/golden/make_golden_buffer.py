"""Golden outputs of the UNMODIFIED reference binaries (oracle/_ref/bin, built by oracle/Makefile) for the
`-buffer N` option (datafile.c:237-344): chunk-wise reading, every chunk re-shuffled with `-rand` each
time it is read.  Run in the build container:  python tests/golden/make_golden_buffer.py"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
BIN = os.path.join(ROOT, "oracle", "_ref", "bin")


def run(d, prog, *args):
    p = subprocess.run([os.path.join(BIN, prog), *args], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, (prog, p.stderr)
    return p.stdout


def main():
    g = np.load(os.path.join(HERE, "demo.npz"))
    out = {}
    with tempfile.TemporaryDirectory() as d:
        for f in ("ex.dat", "ex1.dat", "ex2.dat"):
            open(os.path.join(d, f), "w").write(str(g["in_" + f]))
        open(os.path.join(d, "ex.cod"), "w").write(str(g["som_init_cod"]))
        open(os.path.join(d, "ex1o.cod"), "w").write(str(g["lvq_o_cod"]))
        # 3840 rows in chunks of 500 (last chunk 340), 3 passes and a bit; chunk shuffles with a running generator
        run(d, "vsom", "-din", "ex.dat", "-cin", "ex.cod", "-cout", "b1.cod", "-rlen", "9000", "-alpha", "0.05",
            "-radius", "6", "-rand", "3", "-buffer", "500")
        out["som_buffer_rand_cod"] = open(os.path.join(d, "b1.cod")).read()
        # buffer == number of rows: still buffered, the single chunk is re-shuffled on every pass
        run(d, "vsom", "-din", "ex.dat", "-cin", "ex.cod", "-cout", "b2.cod", "-rlen", "8000", "-alpha", "0.05",
            "-radius", "6", "-rand", "7", "-buffer", "3840")
        out["som_buffer_exact_cod"] = open(os.path.join(d, "b2.cod")).read()
        # buffer larger than the file: buffering is switched off, one shuffle
        run(d, "vsom", "-din", "ex.dat", "-cin", "ex.cod", "-cout", "b3.cod", "-rlen", "5000", "-alpha", "0.05",
            "-radius", "6", "-rand", "7", "-buffer", "5000")
        out["som_buffer_large_cod"] = open(os.path.join(d, "b3.cod")).read()
        # no -rand: chunks in file order
        run(d, "vsom", "-din", "ex.dat", "-cin", "ex.cod", "-cout", "b4.cod", "-rlen", "4500", "-alpha", "0.05",
            "-radius", "6", "-buffer", "700")
        out["som_buffer_norand_cod"] = open(os.path.join(d, "b4.cod")).read()
        run(d, "lvq1", "-din", "ex1.dat", "-cin", "ex1o.cod", "-cout", "l1.cod", "-alpha", "0.05", "-rlen", "6000",
            "-rand", "5", "-buffer", "300")
        out["lvq_buffer_rand_cod"] = open(os.path.join(d, "l1.cod")).read()
        out["qerror_buffer_stdout"] = run(d, "qerror", "-din", "ex.dat", "-cin", "b1.cod", "-buffer", "500")
        out["accuracy_buffer_stdout"] = run(d, "accuracy", "-din", "ex2.dat", "-cin", "l1.cod", "-buffer", "300")
    np.savez_compressed(os.path.join(HERE, "demo_buffer.npz"), **out)
    print("wrote demo_buffer.npz:", sorted(out))


if __name__ == "__main__":
    sys.exit(main())
