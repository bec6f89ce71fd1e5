#!/usr/bin/env python
"""Wall time a USER of the command line sees: `qerror -din big.dat -cin map.cod` with the file parse
included, this repo's C host (bmu_pak qerror, block-parallel parser + sharded GPU search) beside the
unmodified reference binary (oracle/_ref/bin/qerror, one core) on the same files.

    python tools/bench_cli_qerror.py [--rows 1000000] [--ref-rows 50000] [--buffer 0]

Synthetic 64-dim data against a 100x100 hexa map (BASELINE.json configs[2] shape), written by
`bmu_pak paksynth`.  The reference needs ~0.36 ms per row for the search alone (10 000 x 64 map), so it
is timed on the first --ref-rows rows and its full-size time is the linear extrapolation, stated as such.
Writes one JSON line; profiles/r02_cli_qerror_walltime.txt keeps the run."""
import argparse
import json
import os
import subprocess
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAK = os.path.join(ROOT, "som_lvq_pak_b200", "host", "bmu_pak")
REF = os.path.join(ROOT, "oracle", "_ref", "bin", "qerror")


def timed(cmd, cwd):
    t0 = time.perf_counter()
    p = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    dt = time.perf_counter() - t0
    assert p.returncode == 0, (cmd, p.stderr[-1000:])
    return dt, p.stdout.strip()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--ref-rows", type=int, default=50_000)
    ap.add_argument("--buffer", type=int, default=0)
    ap.add_argument("--dir", default=None)
    a = ap.parse_args()
    with tempfile.TemporaryDirectory(dir=a.dir) as d:
        t_gen, _ = timed([PAK, "paksynth", "-dout", "big.dat", "-rows", str(a.rows), "-dim", "64", "-seed", "1"], d)
        timed([PAK, "paksynth", "-dout", "map.cod", "-rows", "10000", "-dim", "64", "-seed", "2", "-xdim", "100"], d)
        timed(["head", "-n", str(a.ref_rows + 1), "big.dat"], d)
        with open(os.path.join(d, "small.dat"), "w") as f:
            subprocess.run(["head", "-n", str(a.ref_rows + 1), "big.dat"], cwd=d, stdout=f, check=True)
        size = os.path.getsize(os.path.join(d, "big.dat"))
        extra = ["-buffer", str(a.buffer)] if a.buffer else []
        t_ours_small, out_small = timed([PAK, "qerror", "-din", "small.dat", "-cin", "map.cod"], d)
        t_ours, out_ours = timed([PAK, "qerror", "-din", "big.dat", "-cin", "map.cod", *extra], d)
        t_ours2, _ = timed([PAK, "qerror", "-din", "big.dat", "-cin", "map.cod", *extra], d)      # page cache warm
        line = {"rows": a.rows, "file_bytes": size, "paksynth_s": t_gen,
                "bmu_pak_qerror_s": min(t_ours, t_ours2), "bmu_pak_qerror_first_run_s": t_ours,
                "bmu_pak_rows_per_s": a.rows / min(t_ours, t_ours2), "bmu_pak_stdout": out_ours,
                "bmu_pak_qerror_small_s": t_ours_small, "buffer": a.buffer}
        if os.path.exists(REF):
            t_ref, out_ref = timed([REF, "-din", "small.dat", "-cin", "map.cod"], d)
            line.update({"reference_rows": a.ref_rows, "reference_qerror_s": t_ref,
                         "reference_rows_per_s": a.ref_rows / t_ref,
                         "reference_full_size_s_extrapolated": t_ref * a.rows / a.ref_rows,
                         "same_stdout_on_the_reference_rows": out_ref == out_small,
                         "speedup_wall_extrapolated": (t_ref * a.rows / a.ref_rows) / min(t_ours, t_ours2)})
        print(json.dumps(line))


if __name__ == "__main__":
    main()
