#!/usr/bin/env python3
"""Quick GPU smoke: product vs compiled reference on a few DBs (development aid, not a test)."""
import os, sys, time, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.refapi import *
from libssa_b200 import synth

D = os.path.join(ROOT, "tests/golden/data")
prod = LibSSA(os.path.join(ROOT, "libssa_b200/libssa_b200.so"))
ref = reference()
fails = 0

def setup(lib, sym, scoring, gaps, db, qfile):
    lib.set_output_mode(OUTPUT_SILENT)
    lib.set_thread_count(1); lib.set_chunk_size(1000)
    lib.init_symbol_translation(sym)
    if isinstance(scoring, tuple): lib.init_constant_scores(*scoring)
    else: lib.init_score_matrix(MATRIX_BUILDIN, scoring)
    lib.init_gap_penalties(*gaps)
    lib.init_db(db)
    return lib.init_sequence_fasta(READ_FROM_FILE, qfile)

def compare(name, sym, scoring, gaps, db, qfile, k=10, widths=(64, 16, 8), modes=("sw", "nw")):
    global fails
    qp = setup(prod, sym, scoring, gaps, db, qfile)
    qr = setup(ref, sym, scoring, gaps, db, qfile)
    for mode in modes:
        for bw in widths:
            t0 = time.time(); hp = getattr(prod, mode + "_align")(qp, k, bw, COMPUTE_SCORE); tp = time.time() - t0
            t0 = time.time(); hr = getattr(ref, mode + "_align")(qr, k, bw, COMPUTE_SCORE); tr = time.time() - t0
            a = [(h.score, h.db_id) for h in hp]; b = [(h.score, h.db_id) for h in hr]
            ok = a == b
            fails += (not ok)
            print("%-28s %s %2d  %s  prod %.3fs ref %.3fs" % (name, mode, bw, "OK" if ok else "MISMATCH", tp, tr))
            if not ok:
                print("   prod", a[:12]); print("   ref ", b[:12])

compare("AF091148 const(2,-2)", NUCLEOTIDE, (2, -2), (-4, -2), D + "/AF091148.fas", D + "/one_seq.fas")
compare("AF091148 const(5,-4)", NUCLEOTIDE, (5, -4), (-4, -2), D + "/AF091148.fas", D + "/one_seq.fas", k=5)
compare("short_nuc const(1,-1)", NUCLEOTIDE, (1, -1), (-1, -1), D + "/short_nuc_db.fas", D + "/one_seq.fas", k=8)
q = open(D + "/P07327").read().split("\n", 1)[1].replace("\n", "")
tmp = tempfile.mkdtemp()
for n, seed in ((3000, 7), (20000, 20150403)):
    db = synth.protein_db(n, seed, query=q)
    path = os.path.join(tmp, "s%d.fas" % n)
    synth.write_fasta(path, db)
    compare("synth%d blosum62 P07327" % n, AMINOACID, "blosum62", (-11, -1), path, D + "/P07327", k=25)
    compare("synth%d blosum62 O74807" % n, AMINOACID, "blosum62", (-11, -1), path, D + "/O74807", k=25, widths=(16, 8))
compare("synth20000 blosum50 P33450", AMINOACID, "blosum50", (-3, -1), path, D + "/P33450", k=10, widths=(16,), modes=("sw",))

def compare_align(name, sym, scoring, gaps, db, qfile, k=10, widths=(16,), modes=("sw", "nw")):
    global fails
    qp = setup(prod, sym, scoring, gaps, db, qfile)
    qr = setup(ref, sym, scoring, gaps, db, qfile)
    for mode in modes:
        for bw in widths:
            t0 = time.time(); hp = getattr(prod, mode + "_align")(qp, k, bw, COMPUTE_ALIGNMENT); tp = time.time() - t0
            t0 = time.time(); hr = getattr(ref, mode + "_align")(qr, k, bw, COMPUTE_ALIGNMENT); tr = time.time() - t0
            a = [h.key() for h in hp]; b = [h.key() for h in hr]
            ok = a == b
            fails += (not ok)
            print("ALIGN %-28s %s %2d  %s  prod %.3fs ref %.3fs" % (name, mode, bw, "OK" if ok else "MISMATCH", tp, tr))
            if not ok:
                for x, y in zip(a, b):
                    if x != y: print("   prod", x); print("   ref ", y); break

compare_align("AF091148 const(5,-4)", NUCLEOTIDE, (5, -4), (-4, -2), D + "/AF091148.fas", D + "/one_seq.fas", k=5, widths=(64, 16))
compare_align("AF091148 const(2,-2)", NUCLEOTIDE, (2, -2), (-4, -2), D + "/AF091148.fas", D + "/one_seq.fas", k=20)
compare_align("test.fas const(1,-1)", NUCLEOTIDE, (1, -1), (-1, -1), D + "/test.fas", D + "/one_seq.fas", k=5)
compare_align("synth20000 blosum62 P07327", AMINOACID, "blosum62", (-11, -1), path, D + "/P07327", k=25)
compare_align("synth20000 blosum62 O74807", AMINOACID, "blosum62", (-11, -1), path, D + "/O74807", k=25)
compare_align("synth20000 blosum50 P27895", AMINOACID, "blosum50", (-3, -1), path, D + "/P27895", k=10)
print("FAILS", fails)
sys.exit(1 if fails else 0)
