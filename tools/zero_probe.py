#!/usr/bin/env python3
"""Aligning a Smith-Waterman hit of score 0 (tests/test_gpu_parity.py::test_zero_score_alignment_fails_like_the_reference)."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import libssa_b200 as L
lib = L.load()
p = os.path.join(tempfile.mkdtemp(), "zero.fas")
open(p, "w").write(">a\nMKV\n>b\nWW\n>c\nAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA\n>d\nG\n")
lib.init_symbol_translation(L.AMINOACID); lib.init_score_matrix(L.MATRIX_BUILDIN, "pam30"); lib.init_gap_penalties(-5, -5)
lib.init_db(p)
q = lib.init_sequence_fasta(L.READ_FROM_STRING, "W")
h = lib.sw_align(q, 6, 16, L.COMPUTE_SCORE); print([(x.score, x.db_id) for x in h], flush=True)
h = lib.sw_align(q, 6, 16, L.COMPUTE_ALIGNMENT); print([(x.score, x.db_id, x.cigar, x.q_start, x.q_end, x.d_start, x.d_end) for x in h], flush=True)
