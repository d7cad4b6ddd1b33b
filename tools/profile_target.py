#!/usr/bin/env python3
"""Small, fixed workload for ncu: P07327 (375 aa) SW searches against a 100 000-record synthetic database."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import libssa_b200 as L
from libssa_b200 import synth
D = os.path.join(ROOT, "tests/golden/data")
qname = sys.argv[1] if len(sys.argv) > 1 else "P07327"
mode = sys.argv[2] if len(sys.argv) > 2 else "sw"
q0 = open(D + "/P07327").read().split("\n", 1)[1].replace("\n", "")
nseq = int(sys.argv[3]) if len(sys.argv) > 3 else 100000
path = os.path.join(tempfile.gettempdir(), "prof_S%d.fas" % nseq)
if not os.path.exists(path):
    synth.write_protein_db(path, nseq, 20150402, query=q0)
lib = L.load()
lib.set_output_mode(L.OUTPUT_SILENT)
lib.init_symbol_translation(L.AMINOACID); lib.init_score_matrix(L.MATRIX_BUILDIN, "blosum62"); lib.init_gap_penalties(-11, -1)
lib.init_db(path)
q = lib.init_sequence_fasta(L.READ_FROM_FILE, D + "/" + qname)
for it in range(4):
    hits = (lib.sw_align if mode == "sw" else lib.nw_align)(q, 10, 16, L.COMPUTE_SCORE)
s = lib.stats()
print("kernel_ms %.3f gcups %.1f top %d" % (s.score_kernel_ms, s.cells / s.score_kernel_ms / 1e6, hits[0].score))
