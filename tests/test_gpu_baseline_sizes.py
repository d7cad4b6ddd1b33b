"""GPU parity at BASELINE.json sizes: the CUDA path through the C-ABI against the compiled, unmodified reference run with ONE
thread (its deterministic order) on the very databases bench.py measures.  Mirrors the reference's own large-database tests
(reference tests/test_bigger_databases.c:70-162: top-10 lists pinned on its largest databases; tests/test_libssa.c:102-178:
alignments).  Everything is integer work: every comparison is bit-exact."""
import os
import sys

import numpy as np
import pytest

import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _setup_protein(lib, path, chunk=1000):
    lib.set_output_mode(0)
    lib.set_thread_count(1)
    lib.set_chunk_size(chunk)
    lib.init_symbol_translation(cases.AMINOACID, 1, 1, 1)
    lib.init_score_matrix(cases.MATRIX_BUILDIN, "blosum62")
    lib.init_gap_penalties(-11, -1)
    lib.init_db(path)


def _keys(hits, align=False):
    return [(h.score, h.db_id) + ((h.cigar, h.q_start, h.q_end, h.d_start, h.d_end) if align else ()) for h in hits]


def test_config1_500k_hit_lists_equal_reference(prod, ref):
    """BASELINE configs[1]: the 500 000-record database of bench.py (same seed), queries O74807 and P07327, entry bit widths
    8 and 16, k = 10 and a k at which equal scores straddle the cut-off.  With more than 262 144 records the product takes
    the real sparse heap replay (first 65 536 arrivals + later records above the root) and keeps the arrival order on the
    device; at bit width 8 the records scoring >= 255 go through the first-crossing tracker and the emulated 8 -> 16 bit
    cascade.  None of this had met the reference at this size before."""
    import bench
    path, meta = bench.protein_db_file(500000, 20150404)
    assert meta["n_seqs"] == 500000
    _setup_protein(prod, path)
    _setup_protein(ref, path)
    for name in ("O74807", "P07327"):
        q = prod.init_sequence_fasta(cases.READ_FROM_FILE, os.path.join(cases.DATA, name))
        rq = ref.init_sequence_fasta(cases.READ_FROM_FILE, os.path.join(cases.DATA, name))
        _, sw = prod.score_all(q, False)
        assert len(sw) == 500000
        desc = np.sort(sw)[::-1]
        k_tie = next(k for k in range(40, 20000) if desc[k - 1] == desc[k])      # ties across the cut-off: replay path
        for bw in (8, 16):
            for k in (10, k_tie):
                got = _keys(prod.sw_align(q, k, bw, cases.COMPUTE_SCORE))
                st = prod.stats()
                want = _keys(ref.sw_align(rq, k, bw, cases.COMPUTE_SCORE))
                assert got == want, (name, bw, k, got[-3:], want[-3:])
                if k == k_tie:
                    # the replay ran in its sparse form: far less than 4 bytes per record crossed PCIe
                    assert st.d2h_bytes < 4 * 500000, (name, bw, k, st.d2h_bytes)
        prod.free_sequence(q)
        ref.free_sequence(rq)


def test_config4_10kbp_alignments_equal_reference(prod, ref, tmp_path):
    """BASELINE configs[4]: nucleotide, 10 kbp query with IUPAC codes, SW and NW, top-10 with CIGARs.  Database: six of the
    planted 8-12 kbp homologs and the first 300 records of bench.py's 50 000-record database (the reference needs its scalar
    64-bit tier for the long NW pairs, so the full database is out of reach for one CPU thread)."""
    import bench
    from libssa_b200 import synth
    _, seqs, qtext, long_hom = bench.nucleotide_db_file()
    assert len(qtext) == 10000
    idx = long_hom[:6] + list(range(300))
    path = tmp_path / "nt306.fas"
    synth.write_fasta(str(path), [seqs[i] for i in idx])
    assert max(len(seqs[i]) for i in idx) >= 8000
    results = []
    for lib in (ref, prod):
        lib.set_output_mode(0)
        lib.set_thread_count(1)
        lib.set_chunk_size(1000)
        lib.init_symbol_translation(cases.NUCLEOTIDE, 1, 1, 1)
        lib.init_constant_scores(5, -4)
        lib.init_gap_penalties(-4, -2)
        lib.init_db(str(path))
        q = lib.init_sequence_fasta(cases.READ_FROM_STRING, qtext)
        out = []
        for mode, bw in (("sw", 16), ("nw", 16), ("sw", 8), ("nw", 64)):
            f = lib.sw_align if mode == "sw" else lib.nw_align
            out.append(_keys(f(q, 10, bw, cases.COMPUTE_ALIGNMENT), True))
        lib.free_sequence(q)
        results.append(out)
    assert results[0] == results[1]
    assert max(len(h[2]) for h in results[1][0]) > 100       # real gapped alignments, not trivial ones


def test_config2_nw_batch_of_60_queries_equals_reference(prod, ref, tmp_path):
    """BASELINE configs[2]: NW entry bit width 16, a batch of 60 queries (the first records of the database) through
    ssa_b200_align_batch against a length-matched database (315-385 aa; 20 000 records so that one reference thread
    finishes in seconds): every list equals the reference's one-call-per-query result; the first 6 also with CIGARs."""
    from libssa_b200 import synth
    seqs = synth.protein_db(20000, 20150405, mu=float(np.log(350)), sigma=0.03, min_len=315, max_len=385, n_dup=50)
    path = tmp_path / "lm20k.fas"
    synth.write_fasta(str(path), seqs)
    _setup_protein(prod, str(path))
    _setup_protein(ref, str(path))
    NQ = 60
    queries = [prod.init_sequence_fasta(cases.READ_FROM_STRING, seqs[i]) for i in range(NQ)]
    got = [_keys(h) for h in prod.align_batch(queries, 10, 16, cases.COMPUTE_SCORE, nw=True)]
    got_al = [_keys(h, True) for h in prod.align_batch(queries[:6], 10, 16, cases.COMPUTE_ALIGNMENT, nw=True)]
    for i in range(NQ):
        rq = ref.init_sequence_fasta(cases.READ_FROM_STRING, seqs[i])
        assert got[i] == _keys(ref.nw_align(rq, 10, 16, cases.COMPUTE_SCORE)), i
        if i < 6:
            assert got_al[i] == _keys(ref.nw_align(rq, 10, 16, cases.COMPUTE_ALIGNMENT), True), i
        ref.free_sequence(rq)
    for q in queries:
        prod.free_sequence(q)
