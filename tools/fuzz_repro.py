#!/usr/bin/env python3
"""Development aid: replays the random inputs of tests/test_gpu_parity.py::test_live_reference_random_inputs up to a given trial and
compares exact scores (score_all) with the oracle port.   python tools/fuzz_repro.py <seed> <trial>"""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cases
import libssa_b200 as L
from libssa_b200 import synth
from oracle import port
seed, want = int(sys.argv[1]), int(sys.argv[2])
lib = L.load(); lib.set_output_mode(0)
rng = np.random.default_rng(seed)
tmp = tempfile.mkdtemp()
for trial in range(want + 1):
    n = int(rng.integers(1, 3000))
    qlen = int(rng.choice([1, 2, 17, 33, 128, 129, 250, 700, 1300, 2100]))
    q = synth.random_protein(rng, qlen)
    seqs = synth.protein_db(n, int(rng.integers(1, 1 << 30)), query=q, homolog_frac=0.03, n_dup=min(20, n // 3),
                            mu=float(rng.choice([3.0, 5.0])), sigma=0.7, min_len=0, max_len=4000)
    scoring = ["blosum62", "blosum45", "pam30", [4, -3]][trial % 4]
    gaps = [[-11, -1], [-3, -1], [-5, -5], [0, -2]][trial % 4]
    nuc = trial % 5 == 4
    if nuc:
        q = synth.random_nucleotide(int(rng.integers(1, 1 << 30)), qlen, iupac_rate=0.03)
        seqs = synth.nucleotide_db(min(n, 800), int(rng.integers(1, 1 << 30)), query=q, n_homologs=min(5, n), mu=float(rng.choice([4.0, 6.0])),
                                   sigma=0.6, min_len=1, max_len=5000, iupac_rate=0.01)
        scoring = [[5, -4], [2, -2], [1, -1]][trial % 3]
        gaps = [[-4, -2], [-1, -1], [-10, -1]][(trial // 5) % 3]
path = os.path.join(tmp, "db.fas"); synth.write_fasta(path, seqs)
print("trial", want, "n", len(seqs), "qlen", qlen, scoring, gaps, "nuc", nuc, "query", q)
lib.init_symbol_translation(cases.NUCLEOTIDE if nuc else cases.AMINOACID, 1, 1, 1)
if isinstance(scoring, str): lib.init_score_matrix(cases.MATRIX_BUILDIN, scoring)
else: lib.init_constant_scores(*scoring)
lib.init_gap_penalties(*gaps)
lib.init_db(path)
qh = lib.init_sequence_fasta(cases.READ_FROM_STRING, q)
M = np.zeros((32, 32), dtype=np.int64)
buf = (np.ctypeslib.ctypes.c_int64 * 1024)() if False else None
import ctypes
m = (ctypes.c_int64 * 1024)(); lib.lib.ssa_b200_get_score_matrix(m); M = np.array(list(m), dtype=np.int64).reshape(32, 32)
codes = [port.map_db(s, nuc) for s in seqs]; qc = port.map_query(q, nuc)
for nw in (False, True):
    ids, sc = lib.score_all(qh, nw)
    exp = port.score_db(codes, qc, M, gaps[0], gaps[1], nw)
    bad = np.nonzero(sc != exp)[0]
    print("nw" if nw else "sw", "mismatches", len(bad), [(int(i), len(seqs[i]), int(sc[i]), int(exp[i])) for i in bad[:12]])
    for k, bw in ((30, 8), (30, 16)):
        hits = (lib.nw_align if nw else lib.sw_align)(qh, k, bw, 0)
        print("   hits bw", bw, [(h.score, h.db_id) for h in hits][20:28])
# the compiled reference on the same input (if built): full NW lists at 8 and 16 bit, differences against ours
from oracle import refapi
ref = refapi.reference()
if ref is not None:
    ref.set_output_mode(0); ref.set_thread_count(1)
    ref.init_symbol_translation(cases.NUCLEOTIDE if nuc else cases.AMINOACID, 1, 1, 1)
    if isinstance(scoring, str): ref.init_score_matrix(cases.MATRIX_BUILDIN, scoring)
    else: ref.init_constant_scores(*scoring)
    ref.init_gap_penalties(*gaps)
    ref.set_chunk_size([1000, 64][want % 2]); lib.set_chunk_size([1000, 64][want % 2])
    ref.init_db(path)
    rq = ref.init_sequence_fasta(cases.READ_FROM_STRING, q)
    _, sc = lib.score_all(qh, True)
    for bw in (8, 16, 64):
        k = len(seqs)
        r = [(h.score, h.db_id) for h in ref.nw_align(rq, k, bw, 0)]
        o = [(h.score, h.db_id) for h in lib.nw_align(qh, k, bw, 0)]
        rs = dict((i, s) for s, i in r)
        wrong = [(i, len(seqs[i]), rs[i], int(sc[i])) for i in rs if rs[i] != int(sc[i])]
        print("ref nw bw", bw, "records whose reference score differs from the exact score (id, len, ref, exact):", wrong[:10], "lists equal:", r == o)
