#!/usr/bin/env python3
"""Runs the five BASELINE.json configurations on one B200 (config 4 at N = 1) and writes
profiles/configs_<tag>.md: throughput of the product and - on bounded samples - of the compiled reference,
plus a parity check of the hit lists where the reference is run.  Development/evidence tool, not a bench line."""
import json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import libssa_b200 as L
from libssa_b200 import synth
from oracle import refapi
D = os.path.join(ROOT, "tests/golden/data")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
TMP = tempfile.mkdtemp(prefix="ssa_cfg_")
lib = L.load(); lib.set_output_mode(L.OUTPUT_SILENT)
ref = refapi.reference()
rows = []

def qtext(n): return "".join(l.strip() for l in open(os.path.join(D, n)).read().split("\n") if l and not l.startswith(">"))

def setup(api, sym, scoring, gaps, db):
    api.set_output_mode(0); api.set_chunk_size(1000)
    api.init_symbol_translation(sym, 1, 1, 1)
    if isinstance(scoring, tuple): api.init_constant_scores(*scoring)
    else: api.init_score_matrix(L.MATRIX_BUILDIN, scoring)
    api.init_gap_penalties(*gaps); api.init_db(db)

def timed(fn, reps):
    fn(); t0 = time.perf_counter()
    for _ in range(reps): r = fn()
    return (time.perf_counter() - t0) / reps, r

def keyed(hits, align): return [(h.score, h.db_id) + ((h.cigar, h.q_start, h.q_end, h.d_start, h.d_end) if align else ()) for h in hits]

def residues(path): return sum(len(l.strip()) for l in open(path) if not l.startswith(">"))

# ---- config 1: SW 16-bit, P07327 vs the protein files of tests/testdata (parity) and S-20k (timing) ----
p1 = os.path.join(D, "testdata_proteins.fas")
s20k = os.path.join(TMP, "s20k.fas"); synth.write_protein_db(s20k, 20000, 20150403, query=qtext("P07327"))
for name, db in (("tests/testdata proteins (6 records)", p1), ("S-20k", s20k)):
    res = residues(db)
    setup(lib, L.AMINOACID, "blosum62", (-11, -1), db); q = lib.init_sequence_fasta(L.READ_FROM_FILE, os.path.join(D, "P07327"))
    tg, hg = timed(lambda: lib.sw_align(q, 10, 16, L.COMPUTE_SCORE), 20)
    ref.set_thread_count(1); setup(ref, L.AMINOACID, "blosum62", (-11, -1), db); qr = ref.init_sequence_fasta(L.READ_FROM_FILE, os.path.join(D, "P07327"))
    t1, hr = timed(lambda: ref.sw_align(qr, 10, 16, L.COMPUTE_SCORE), 3)
    same = keyed(hg, False) == keyed(hr, False)
    ref.set_thread_count(os.cpu_count()); tn, _ = timed(lambda: ref.sw_align(qr, 10, 16, L.COMPUTE_SCORE), 5)
    rows.append(("1", "SW-16 P07327 (375 aa) vs %s, BLOSUM62 -11/-1" % name, 375 * res / tg / 1e9, 375 * res / t1 / 1e9, 375 * res / tn / 1e9, "identical" if same else "MISMATCH"))

# ---- config 3: NW 16-bit, many queries vs a length-matched 100k database; score pass, then alignments ----
rng = np.random.default_rng(20150405)
lens = np.clip(np.rint(rng.normal(350, 20, 100000)), 315, 385).astype(int)
seqs = [synth.random_protein(rng, n) for n in lens[:3000]] if False else None
db3 = os.path.join(TMP, "s100k_lm.fas")
allseq = synth.protein_db(100000, 20150405, mu=float(np.log(350)), sigma=0.03, min_len=315, max_len=385, n_dup=50)
synth.write_fasta(db3, allseq); res3 = sum(len(s) for s in allseq)
setup(lib, L.AMINOACID, "blosum62", (-11, -1), db3)
NQ = 200
queries = [lib.init_sequence_fasta(L.READ_FROM_STRING, allseq[i]) for i in range(NQ)]
lib.nw_align(queries[0], 10, 16, L.COMPUTE_SCORE)
t0 = time.perf_counter()
for q in queries: lib.nw_align(q, 10, 16, L.COMPUTE_SCORE)
t_score = time.perf_counter() - t0
t0 = time.perf_counter()
for q in queries: lib.nw_align(q, 10, 16, L.COMPUTE_ALIGNMENT)
t_align = time.perf_counter() - t0
cells3 = sum(len(allseq[i]) for i in range(NQ)) * res3
# the same queries through the batched entry point (ssa_b200_align_batch): host work of query i overlaps query i+1
single_lists = [keyed(lib.nw_align(q, 10, 16, L.COMPUTE_ALIGNMENT), True) for q in queries[:20]]
lib.align_batch(queries[:4], 10, 16, L.COMPUTE_SCORE, nw=True)
t0 = time.perf_counter(); lib.align_batch(queries, 10, 16, L.COMPUTE_SCORE, nw=True); t_bscore = time.perf_counter() - t0
t0 = time.perf_counter(); bl = lib.align_batch(queries, 10, 16, L.COMPUTE_ALIGNMENT, nw=True); t_balign = time.perf_counter() - t0
ok3b = [keyed(h, True) for h in bl[:20]] == single_lists
ref.set_thread_count(os.cpu_count()); setup(ref, L.AMINOACID, "blosum62", (-11, -1), db3)
ok3 = True; tr = 0
for i in range(3):
    qr = ref.init_sequence_fasta(L.READ_FROM_STRING, allseq[i]); t0 = time.perf_counter(); hr = ref.nw_align(qr, 10, 16, L.COMPUTE_ALIGNMENT); tr += time.perf_counter() - t0
    ref.set_thread_count(1); hr1 = ref.nw_align(qr, 10, 16, L.COMPUTE_ALIGNMENT); ref.set_thread_count(os.cpu_count())
    ok3 &= keyed(lib.nw_align(queries[i], 10, 16, L.COMPUTE_ALIGNMENT), True) == keyed(hr1, True)
cells_ref = sum(len(allseq[i]) for i in range(3)) * res3
rows.append(("3", "NW-16, %d queries (one call each) vs length-matched 100k db: score pass" % NQ, cells3 / t_score / 1e9, None, None, ""))
rows.append(("3", "same, COMPUTE_ALIGNMENT with hitcount 10 (%.2f ms per query)" % (t_align / NQ * 1e3), cells3 / t_align / 1e9, None, cells_ref / tr / 1e9, "identical (3 queries, CIGARs)" if ok3 else "MISMATCH"))
rows.append(("3", "score pass as ONE ssa_b200_align_batch call (%.2f ms per query)" % (t_bscore / NQ * 1e3), cells3 / t_bscore / 1e9, None, None, ""))
rows.append(("3", "COMPUTE_ALIGNMENT as ONE ssa_b200_align_batch call (%.2f ms per query)" % (t_balign / NQ * 1e3), cells3 / t_balign / 1e9, None, None, "identical to single calls (20 queries)" if ok3b else "MISMATCH"))

# ---- config 4 at N = 1: SW 8/16-bit vs S-570k ----
db4 = os.path.join(TMP, "s570k.fas"); res4 = synth.write_protein_db(db4, 570000, 20150406, query=qtext("P07327"))
setup(lib, L.AMINOACID, "blosum62", (-11, -1), db4)
for qn in ("P07327", "P33450"):
    q = lib.init_sequence_fasta(L.READ_FROM_FILE, os.path.join(D, qn)); ql = len(qtext(qn))
    for bw in (8, 16):
        tg, _ = timed(lambda: lib.sw_align(q, 10, bw, L.COMPUTE_SCORE), 3)
        rows.append(("4", "SW-%d %s (%d aa) vs S-570k (%d M residues), 1 GPU" % (bw, qn, ql, res4 // 10**6), ql * res4 / tg / 1e9, None, None, ""))

# ---- config 5: nucleotide, 10 kbp query, alignments of the top hits ----
q5 = synth.random_nucleotide(20150407, 10000, iupac_rate=0.02)
seq5 = synth.nucleotide_db(50000, 20150407, query=q5, n_homologs=200, mu=7.25, sigma=0.35, min_len=100, max_len=12000)
hom = np.random.default_rng(7).choice(50000, 200, replace=False)
for h in hom[:100]: seq5[h] = synth.mutate(np.random.default_rng(int(h)), q5[: 8000 + int(h) % 4000], 0.15, "ACGT")
db5 = os.path.join(TMP, "nt50k.fas"); synth.write_fasta(db5, seq5); res5 = sum(len(s) for s in seq5)
setup(lib, L.NUCLEOTIDE, (5, -4), (-4, -2), db5); q = lib.init_sequence_fasta(L.READ_FROM_STRING, q5)
for mode in ("sw", "nw"):
    f = lib.sw_align if mode == "sw" else lib.nw_align
    ts, _ = timed(lambda: f(q, 10, 16, L.COMPUTE_SCORE), 2)
    ta, ha = timed(lambda: f(q, 10, 16, L.COMPUTE_ALIGNMENT), 1)
    rows.append(("5", "%s-16 nucleotide (5,-4)/(-4,-2), 10 kbp query vs 50k records (%d M nt): score pass" % (mode.upper(), res5 // 10**6), 10000 * res5 / ts / 1e9, None, None, ""))
    rows.append(("5", "same, COMPUTE_ALIGNMENT of the top 10 (%.0f ms for the alignments; longest CIGAR %d chars)" % ((ta - ts) * 1e3, max(len(h.cigar) for h in ha)), 10000 * res5 / ta / 1e9, None, None, ""))
# parity of config 5 on a bounded sample (the reference's 64-bit fallback is slow on 10 kbp pairs)
small = os.path.join(TMP, "nt_small.fas"); idx = list(hom[:6]) + list(range(300)); synth.write_fasta(small, [seq5[i] for i in idx])
setup(lib, L.NUCLEOTIDE, (5, -4), (-4, -2), small); q = lib.init_sequence_fasta(L.READ_FROM_STRING, q5)
ref.set_thread_count(1); setup(ref, L.NUCLEOTIDE, (5, -4), (-4, -2), small); qr = ref.init_sequence_fasta(L.READ_FROM_STRING, q5)
ok5 = all(keyed(getattr(lib, m + "_align")(q, 5, 16, L.COMPUTE_ALIGNMENT), True) == keyed(getattr(ref, m + "_align")(qr, 5, 16, L.COMPUTE_ALIGNMENT), True) for m in ("sw", "nw"))
rows.append(("5", "parity on a 306-record sample incl. six 8-12 kbp homologs (SW + NW, CIGARs)", None, None, None, "identical" if ok5 else "MISMATCH"))

out = ["# BASELINE.json configurations on one B200 (%s)" % tag, "",
       "`python tools/configs_report.py` - wall clock around the `libssa.h` calls (e2e, database resident), GCUPS = real cells / time.",
       "Reference = the compiled, unmodified libssa (oracle/_ref) on the box's %d host cores; config 2 is `bench.py`." % os.cpu_count(), "",
       "| config | workload | B200 GCUPS (e2e) | reference 1 thread | reference all cores | hit lists vs reference |", "|---|---|---|---|---|---|"]
fmt = lambda v: "-" if v is None else "%.1f" % v
for c, w, g, r1, rn, par in rows: out.append("| %s | %s | %s | %s | %s | %s |" % (c, w, fmt(g), fmt(r1), fmt(rn), par))
open(os.path.join(ROOT, "profiles", "configs_%s.md" % tag), "w").write("\n".join(out) + "\n")
print("\n".join(out))
