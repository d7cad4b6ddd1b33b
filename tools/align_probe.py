#!/usr/bin/env python3
"""Development aid: wall time of COMPUTE_ALIGNMENT for long pairs (10 kbp query vs ten ~10 kbp homologs) and for
many short pairs; run under `ncu --metrics gpu__time_duration.sum` for the per-kernel split."""
import os, sys, time, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import libssa_b200 as L
from libssa_b200 import synth
qlen = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
nhits = int(sys.argv[2]) if len(sys.argv) > 2 else 10
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
rng = np.random.default_rng(3)
q = synth.random_nucleotide(5, qlen, iupac_rate=0.0)
recs = [synth.mutate(rng, q, 0.08, "ACGT") for _ in range(nhits)] + [synth.random_nucleotide(100 + i, 300) for i in range(40)]
path = os.path.join(tempfile.mkdtemp(), "db.fas")
synth.write_fasta(path, recs)
lib = L.load(); lib.set_output_mode(L.OUTPUT_SILENT)
lib.init_symbol_translation(L.NUCLEOTIDE); lib.init_constant_scores(5, -4); lib.init_gap_penalties(-4, -2)
lib.init_db(path)
qq = lib.init_sequence_fasta(L.READ_FROM_STRING, q)
for mode in ("sw", "nw"):
    f = lib.sw_align if mode == "sw" else lib.nw_align
    f(qq, nhits, 16, L.COMPUTE_SCORE)
    for it in range(reps):
        t0 = time.perf_counter(); f(qq, nhits, 16, L.COMPUTE_SCORE); t1 = time.perf_counter()
        h = f(qq, nhits, 16, L.COMPUTE_ALIGNMENT); t2 = time.perf_counter()
    print("%s qlen %d hits %d: score pass %.2f ms, with alignments %.2f ms, launches %d, cigar[0] %d chars"
          % (mode, qlen, nhits, (t1 - t0) * 1e3, (t2 - t1) * 1e3, lib.stats().launches, len(h[0].cigar)))
