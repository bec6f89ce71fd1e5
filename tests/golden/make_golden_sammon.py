#!/usr/bin/env python
"""Golden outputs of Sammon's mapping (sammon.c:83-262) from the UNMODIFIED reference:
 * positions after remove_identicals + sammon_iterate on seeded codebooks
   (oracle/_ref/libref_driver.so -> ref_sammon: the reference's own sammon.c object);
 * the `.sam` file and stderr of the `sammon` binary on the demo's trained map and LVQ codebook.
Run in the build container (needs /root/reference for `make -C oracle ref`):
    python tests/golden/make_golden_sammon.py    ->  tests/golden/sammon.npz"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "bin")

CASES = {  # name: (M, D, sweeps, seed, duplicates, masked)
    "small": (30, 5, 20, 7, False, False),
    "map": (96, 5, 60, 123, False, False),
    "dups": (200, 20, 15, 3, True, False),
    "masked": (120, 9, 25, 11, True, True),
    "odd": (257, 3, 10, 99, False, False),
}


def make_case(name):
    M, D, length, seed, dups, masked = CASES[name]
    rng = np.random.default_rng(sum(map(ord, name)) + 1000)
    codes = rng.random((M, D), dtype=np.float32)
    mask = None
    if masked:
        mask = (rng.random((M, D)) < 0.2).astype(np.uint8)
        mask[:, 0] = 0                                   # every pair shares component 0
        codes[mask != 0] = 0.0
    if dups:
        codes[50] = codes[10]
        codes[51] = codes[10]
        codes[110] = codes[3]
        if masked:
            mask[50] = mask[10]
            mask[51] = mask[10]
            mask[110] = mask[3]
    return codes, mask, length, seed


def main():
    from oracle.pyoracle import Reference
    if not Reference.available() or not os.path.isdir(REF_BIN):
        sys.exit("build the reference first: make -C oracle ref (needs /root/reference)")
    ref = Reference()
    out = {}
    for name in CASES:
        codes, mask, length, seed = make_case(name)
        x, y = ref.sammon(codes, length, seed, mask)
        out[name + "_x"], out[name + "_y"] = x, y
    demo = np.load(os.path.join(HERE, "demo.npz"))
    td = tempfile.mkdtemp()
    for key, cod, rlen, seed in (("cli_map", "som_stage2_cod", 100, 5), ("cli_lvq", "lvq_l_cod", 40, 9)):
        open(os.path.join(td, "c.cod"), "w").write(str(demo[cod]))
        p = subprocess.run([os.path.join(REF_BIN, "sammon"), "-cin", "c.cod", "-cout", "c.sam", "-rlen", str(rlen),
                            "-rand", str(seed), "-v", "2"], cwd=td, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                           text=True, check=True)
        out[key + "_sam"] = np.array(open(os.path.join(td, "c.sam")).read())
        out[key + "_stdout"] = np.array(p.stdout)
        out[key + "_stderr"] = np.array(p.stderr)
    np.savez_compressed(os.path.join(HERE, "sammon.npz"), **out)
    print("wrote sammon.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
