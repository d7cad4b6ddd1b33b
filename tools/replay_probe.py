#!/usr/bin/env python3
"""Development aid: cost of the hit selection (top-k vs arrival-order replay with 8/16-bit overflow emulation) by hit count."""
import os, sys, tempfile, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import libssa_b200 as L
from libssa_b200 import synth
D = os.path.join(ROOT, "tests/golden/data")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500000
lib = L.load(); lib.set_output_mode(L.OUTPUT_SILENT)
q0 = open(D + "/P07327").read().split("\n", 1)[1].replace("\n", "")
path = os.path.join(tempfile.mkdtemp(), "s.fas"); synth.write_fasta(path, synth.protein_db(n, 20150402, query=q0))
lib.init_symbol_translation(L.AMINOACID); lib.init_score_matrix(L.MATRIX_BUILDIN, "blosum62"); lib.init_gap_penalties(-11, -1)
lib.init_db(path); lib.prepare_db()
q = lib.init_sequence_fasta(L.READ_FROM_FILE, D + "/P07327")
for mode in ("sw", "nw"):
    for bw in (64, 16, 8):
        for k in (10, 55, 100, 1000, 5000):
            best = None
            for it in range(4):
                t0 = time.time(); hits = (lib.sw_align if mode == "sw" else lib.nw_align)(q, k, bw, L.COMPUTE_SCORE); wall = (time.time() - t0) * 1e3
                s = lib.stats()
                if it and (best is None or wall < best[0]): best = (wall, s.device_ms, s.score_kernel_ms, s.launches, s.wide_tier_seqs, s.d2h_bytes)
            print(json.dumps({"mode": mode, "bw": bw, "k": k, "wall_ms": round(best[0], 3), "device_ms": round(best[1], 3), "kernel_ms": round(best[2], 3),
                              "launches": best[3], "wide": best[4], "d2h": best[5], "last": [hits[-1].score, hits[-1].db_id]}), flush=True)
