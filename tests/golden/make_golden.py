#!/usr/bin/env python3
"""Regenerates tests/golden/*.json from the compiled, unmodified reference (oracle/_ref/libssa_ref.so).

Run in the build container (needs /root/reference for `make -C oracle ref`):
    python tests/golden/make_golden.py
Outputs
  matrices.json   the reference's 32x32 score table after init_score_matrix / init_constant_scores,
                  read from its exported `score_matrix_64` (reference src/matrices.c:294-296)
  searches.json   hit lists (score, id[, cigar, region]) of the reference for a fixed list of cases on the
                  FASTA files in tests/golden/data (copies of the reference's tests/testdata and
                  benchmark/data inputs) and on seeded synthetic databases (libssa_b200/synth.py)
  synth/*.fas     are NOT stored: they are regenerated from the seed by tests/cases.py
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.refapi import *  # noqa
import cases  # tests/cases.py

HERE = os.path.dirname(os.path.abspath(__file__))


def dump_matrix(ref):
    p = C.POINTER(C.c_int64).in_dll(ref.lib, "score_matrix_64")
    return [int(p[i]) for i in range(1024)]


def main():
    ref = reference()
    assert ref is not None, "build the reference first: make -C oracle ref"
    ref.set_output_mode(OUTPUT_SILENT)
    mats = {}
    for name in ("blosum45", "blosum50", "blosum62", "blosum80", "blosum90", "pam30", "pam70", "pam250"):
        ref.init_score_matrix(MATRIX_BUILDIN, name)
        mats[name] = dump_matrix(ref)
    for match, mismatch in ((1, -1), (2, -2), (5, -4), (127, -1), (30, -1), (12, -4)):
        ref.init_constant_scores(match, mismatch)
        mats["const_%d_%d" % (match, mismatch)] = dump_matrix(ref)
    json.dump(mats, open(os.path.join(HERE, "matrices.json"), "w"))

    out = []
    for case in cases.CASES:
        res = cases.run_case(ref, case)
        out.append({"case": case, "result": res})
        print(case["name"], "->", len(res), "runs")
    json.dump(out, open(os.path.join(HERE, "searches.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
