#!/usr/bin/env python3
"""Small mixed workload for compute-sanitizer (memcheck / racecheck): every kernel family of the library runs at
least once - packed and wide scoring tiers (single- and multi-pass), SW overflow tracking, NW saturating-tier
emulation, arrival order, top-k, traceback, strand / reading-frame searches."""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import libssa_b200 as L  # noqa: E402
from libssa_b200 import synth  # noqa: E402

D = os.path.join(ROOT, "tests/golden/data")
tmp = tempfile.mkdtemp()
lib = L.load()
lib.set_output_mode(L.OUTPUT_SILENT)

q0 = open(D + "/P07327").read().split("\n", 1)[1].replace("\n", "")
prot = os.path.join(tmp, "prot.fas")
synth.write_fasta(prot, synth.protein_db(1500, 7, query=q0))
lib.init_symbol_translation(L.AMINOACID)
lib.init_score_matrix(L.MATRIX_BUILDIN, "blosum62")
lib.init_gap_penalties(-11, -1)
lib.init_db(prot)
for name in ("O74807", "P07327", "P04775", "P33450"):
    q = lib.init_sequence_fasta(L.READ_FROM_FILE, D + "/" + name)
    for bits in (8, 16, 64):
        h = lib.sw_align(q, 10, bits, L.COMPUTE_SCORE)
    h = lib.nw_align(q, 10, 16, L.COMPUTE_SCORE)
    if name == "P07327":
        h = lib.nw_align(q, 1400, 8, L.COMPUTE_SCORE)      # ties straddle the cut-off: emulation + replay
        h = lib.sw_align(q, 1400, 8, L.COMPUTE_SCORE)
        h = lib.sw_align(q, 5, 16, L.COMPUTE_ALIGNMENT)
        h = lib.nw_align(q, 5, 64, L.COMPUTE_ALIGNMENT)
    print(name, h[0].score, flush=True)
    lib.free_sequence(q)

# round 2: the pipelined 32-bit tier (match 127: every pair longer than 257 residues is beyond the static int16 bound; SW
# speculates and recomputes, NW splits statically) and two queries per word (a batch against a length-matched database)
lib.init_constant_scores(127, -1)
lib.init_gap_penalties(-1, -1)
lib.init_db(prot)
q = lib.init_sequence_fasta(L.READ_FROM_STRING, synth.random_protein(__import__("numpy").random.default_rng(3), 700))
print("wide pipe", lib.sw_align(q, 10, 16, L.COMPUTE_SCORE)[0].score, lib.nw_align(q, 10, 16, L.COMPUTE_SCORE)[0].score, flush=True)
lib.free_sequence(q)
import numpy as np  # noqa: E402
lm = os.path.join(tmp, "lm.fas")
synth.write_fasta(lm, synth.protein_db(1200, 9, mu=float(np.log(300)), sigma=0.04, min_len=250, max_len=360, n_dup=10))
lib.init_score_matrix(L.MATRIX_BUILDIN, "blosum62")
lib.init_gap_penalties(-11, -1)
lib.init_db(lm)
rng = np.random.default_rng(4)
qs = [lib.init_sequence_fasta(L.READ_FROM_STRING, synth.random_protein(rng, n)) for n in (300, 310, 64, 350, 900, 1000, 17)]
for nw in (False, True):
    res = lib.align_batch(qs, 10, 16, L.COMPUTE_SCORE, nw=nw)
    print("dual batch", nw, [r[0].score for r in res], flush=True)
for q in qs:
    lib.free_sequence(q)
lib.init_db(prot)

# nucleotide, long query (multi-pass shapes) and both strands (several frames, host replay)
nuc = os.path.join(tmp, "nuc.fas")
synth.write_fasta(nuc, synth.nucleotide_db(300, 11, max_len=4000))
lib.init_symbol_translation(L.NUCLEOTIDE)
lib.init_constant_scores(5, -4)
lib.init_gap_penalties(-6, -2)
lib.init_db(nuc)
q = lib.init_sequence_fasta(L.READ_FROM_STRING, synth.random_nucleotide(5, 3000))
print("nuc", lib.sw_align(q, 10, 16, L.COMPUTE_SCORE)[0].score, lib.nw_align(q, 10, 16, L.COMPUTE_SCORE)[0].score, flush=True)
lib.free_sequence(q)
lib.init_symbol_translation(L.NUCLEOTIDE, 3)
q = lib.init_sequence_fasta(L.READ_FROM_STRING, synth.random_nucleotide(6, 400))
print("both", lib.sw_align(q, 10, 8, L.COMPUTE_ALIGNMENT)[0].score, flush=True)
lib.free_sequence(q)
lib.init_symbol_translation(4, 3)   # TRANS_BOTH, both strands: 6 x 6 frames
lib.init_score_matrix(L.MATRIX_BUILDIN, "blosum62")
lib.init_gap_penalties(-7, -1)
q = lib.init_sequence_fasta(L.READ_FROM_STRING, synth.random_nucleotide(6, 400))
print("tboth", lib.sw_align(q, 10, 8, L.COMPUTE_ALIGNMENT)[0].score, flush=True)
lib.free_sequence(q)
lib.ssa_exit()
print("SANITIZE_TARGET_OK")
