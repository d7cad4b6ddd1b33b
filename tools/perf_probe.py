#!/usr/bin/env python3
"""Development aid: kernel-time GCUPS of the scoring path on a synthetic DB."""
import os, sys, time, tempfile, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import libssa_b200 as L
from libssa_b200 import synth
D = os.path.join(ROOT, "tests/golden/data")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
queries = sys.argv[2].split(",") if len(sys.argv) > 2 else ["P07327"]
modes = sys.argv[3].split(",") if len(sys.argv) > 3 else ["sw"]
lib = L.load()
lib.set_output_mode(L.OUTPUT_INFO)
q0 = open(D + "/P07327").read().split("\n", 1)[1].replace("\n", "")
t0 = time.time(); db = synth.protein_db(n, 20150402, query=q0); tg = time.time() - t0
path = os.path.join(tempfile.mkdtemp(), "s.fas"); synth.write_fasta(path, db)
lib.init_symbol_translation(L.AMINOACID); lib.init_score_matrix(L.MATRIX_BUILDIN, "blosum62"); lib.init_gap_penalties(-11, -1)
t0 = time.time(); lib.init_db(path); lib.prepare_db(); print("gen %.1fs init+prepare %.2fs" % (tg, time.time() - t0))
lib.set_output_mode(L.OUTPUT_SILENT)
for qn in queries:
    q = lib.init_sequence_fasta(L.READ_FROM_FILE, D + "/" + qn)
    for mode in modes:
        for bw in (16, 8):
            f = lib.sw_align if mode == "sw" else lib.nw_align
            best = None
            for it in range(6):
                t0 = time.time(); hits = f(q, 10, bw, L.COMPUTE_SCORE); wall = time.time() - t0
                s = lib.stats()
                if it >= 2 and (best is None or s.score_kernel_ms < best[0]):
                    best = (s.score_kernel_ms, s.device_ms, wall * 1e3, s.cells, s.launches, s.overflow_8, s.wide_tier_seqs)
            k, d, w, cells, la, o8, wide = best
            print(json.dumps({"query": qn, "mode": mode, "bw": bw, "kernel_ms": round(k, 3), "device_ms": round(d, 3), "wall_ms": round(w, 3),
                              "gcups_kernel": round(cells / k / 1e6, 1), "gcups_e2e": round(cells / w / 1e6, 1), "launches": la, "ovf8": o8, "wide": wide,
                              "top": hits[0].score}))
