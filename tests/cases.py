"""Search cases shared by tests/golden/make_golden.py (reference -> fixtures) and the parity tests.

A case names a database (a FASTA file under tests/golden/data or a seeded synthetic database), a query,
the scoring and a list of runs (mode, bit width, hit count, align type).  `run_case` drives any
implementation of the libssa.h API (oracle.refapi.LibSSA or libssa_b200.LibSSA) and returns plain lists.
"""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
DATA = os.path.join(ROOT, "tests", "golden", "data")

NUCLEOTIDE, AMINOACID = 0, 1
COMPUTE_SCORE, COMPUTE_ALIGNMENT = 0, 1
READ_FROM_FILE, READ_FROM_STRING, MATRIX_BUILDIN = 0, 1, 2

_tmp = None
_synth_cache = {}


def query_text(name):
    """Sequence of a one-record FASTA file in tests/golden/data."""
    txt = open(os.path.join(DATA, name)).read()
    return "".join(l.strip() for l in txt.split("\n") if l and not l.startswith(">"))


def synth_db_path(spec):
    """Materialise a synthetic database described by spec = (kind, n, seed, query_file_or_None)."""
    global _tmp
    key = tuple(spec)
    if key in _synth_cache:
        return _synth_cache[key]
    from libssa_b200 import synth
    if _tmp is None:
        _tmp = tempfile.mkdtemp(prefix="ssa_synth_")
    kind, n, seed, qfile = spec
    q = query_text(qfile) if qfile else None
    if kind == "protein":
        seqs = synth.protein_db(n, seed, query=q)
    elif kind == "protein_edge":   # with empty records, one-residue records and unknown symbols
        seqs = synth.protein_db(n, seed, query=q)
        for i in range(3, n, 97):
            seqs[i] = ""
        for i in range(5, n, 101):
            seqs[i] = "W"
        for i in range(7, n, 89):
            seqs[i] = seqs[i][:10] + "?x1-" + seqs[i][10:]
    elif kind == "nucleotide":
        seqs = synth.nucleotide_db(n, seed, query=q, n_homologs=max(2, n // 50))
    elif kind == "nucleotide_long":
        q = synth.random_nucleotide(5, 3000, iupac_rate=0.02)
        seqs = synth.nucleotide_db(n, seed, query=q, n_homologs=12, mu=6.5, sigma=0.6, min_len=20, max_len=6000)
        seqs[3] = q
        seqs[7] = q[:2990] + "ACGT" * 40
    else:
        raise ValueError(kind)
    path = os.path.join(_tmp, "%s_%d_%d.fas" % (kind, n, seed))
    synth.write_fasta(path, seqs)
    _synth_cache[key] = path
    return path


def db_path(db):
    return synth_db_path(db) if isinstance(db, (list, tuple)) else os.path.join(DATA, db)


def configure(lib, case):
    lib.set_output_mode(0)
    lib.set_thread_count(1)
    lib.set_chunk_size(case.get("chunk", 1000))
    lib.init_symbol_translation(case["symtype"], 1, 1, 1)
    sc = case["scoring"]
    if isinstance(sc, (list, tuple)):
        lib.init_constant_scores(sc[0], sc[1])
    else:
        lib.init_score_matrix(MATRIX_BUILDIN, sc)
    lib.init_gap_penalties(case["gaps"][0], case["gaps"][1])
    lib.init_db(db_path(case["db"]))
    if "query_string_fn" in case:
        from libssa_b200 import synth
        return lib.init_sequence_fasta(READ_FROM_STRING, synth.random_nucleotide(5, 3000, iupac_rate=0.02))
    if "query_string" in case:
        return lib.init_sequence_fasta(READ_FROM_STRING, case["query_string"])
    return lib.init_sequence_fasta(READ_FROM_FILE, os.path.join(DATA, case["query"]))


def run_case(lib, case):
    q = configure(lib, case)
    results = []
    for mode, bw, k, align in case["runs"]:
        hits = (lib.sw_align if mode == "sw" else lib.nw_align)(q, k, bw, align)
        if align == COMPUTE_ALIGNMENT:
            results.append([[h.score, h.db_id, h.cigar, h.q_start, h.q_end, h.d_start, h.d_end] for h in hits])
        else:
            results.append([[h.score, h.db_id] for h in hits])
    lib.free_sequence(q)
    return results


def _runs(modes=("sw", "nw"), widths=(64, 16, 8), k=10, align=COMPUTE_SCORE):
    return [[m, w, k, align] for m in modes for w in widths]


S, A = COMPUTE_SCORE, COMPUTE_ALIGNMENT

CASES = [
    # --- the reference's own golden vectors ---
    {"name": "bigger_db_AF091148", "symtype": NUCLEOTIDE, "scoring": [2, -2], "gaps": [-4, -2], "db": "AF091148.fas",
     "query": "one_seq.fas", "runs": _runs()},
    {"name": "libssa_AF091148_align", "symtype": NUCLEOTIDE, "scoring": [5, -4], "gaps": [-4, -2], "db": "AF091148.fas",
     "query": "one_seq.fas", "runs": [["sw", 64, 5, A], ["nw", 64, 5, A], ["sw", 16, 5, A], ["nw", 16, 5, A], ["sw", 8, 5, A]]},
    {"name": "searcher_short_nuc", "symtype": NUCLEOTIDE, "scoring": [1, -1], "gaps": [-1, -1], "db": "short_nuc_db.fas",
     "query_string": "ATGCAAATTT", "runs": _runs(k=8)},
    {"name": "search64_test_fas", "symtype": NUCLEOTIDE, "scoring": [1, -1], "gaps": [-1, -1], "db": "test.fas",
     "query": "one_seq.fas", "runs": _runs(widths=(64, 16), k=5) + [["sw", 8, 5, S], ["sw", 16, 5, A], ["nw", 16, 5, A]]},
    {"name": "searcher_AT_short_db", "symtype": NUCLEOTIDE, "scoring": [1, -1], "gaps": [-1, -1], "db": "short_db.fas",
     "query_string": "AT", "runs": _runs(k=1) + [["sw", 16, 1, A], ["nw", 16, 1, A]]},
    {"name": "simd_sw_tmp", "symtype": NUCLEOTIDE, "scoring": [1, -1], "gaps": [-1, -1], "db": "tmp.fas",
     "query_string": "ATGCAAA", "runs": _runs(modes=("sw",), k=1)},
    {"name": "aligner_one_seq_db", "symtype": NUCLEOTIDE, "scoring": [1, -1], "gaps": [-1, -1], "db": "one_seq_db.fas",
     "query_string": "ATGCAAAAA", "runs": [["sw", 16, 1, A], ["nw", 16, 1, A]]},
    {"name": "searcher_blosum62_short_AA", "symtype": AMINOACID, "scoring": "blosum62", "gaps": [-1, -1], "db": "short_AA.fas",
     "query_string": "HPEVYILIIPGFGIISHVVSTYSKKPVFGEISMVYAMASIGLLGFLVWSHHMYIVGLDADTRAYFTSATMIIAIPTGIKI", "runs": _runs(k=1)},
    {"name": "searcher_const_short_AA", "symtype": AMINOACID, "scoring": [1, -1], "gaps": [-1, -1], "db": "short_AA.fas",
     "query_string": "HPEVYILIIPGFGIISHVVSTYSKKPVFGEISMVYAMASIGLLGFLVWSHHMYIVGLDADTRAYFTSATMIIAIPTGIKI", "runs": _runs(k=1)},
    {"name": "overflow_8_selfhit", "symtype": AMINOACID, "scoring": [1, -1], "gaps": [-1, -1], "db": "NP_009305.1.fas",
     "query": "NP_009305.1.fas", "runs": _runs(modes=("sw",), k=1)},
    {"name": "overflow_8_16_64_cascade", "symtype": AMINOACID, "scoring": [127, -1], "gaps": [-1, -1], "db": "NP_009305.1.fas",
     "query": "NP_009305.1.fas", "runs": _runs(k=1)},
    {"name": "overflow_8_const30", "symtype": AMINOACID, "scoring": [30, -1], "gaps": [-1, -1], "db": "short_AA.fas",
     "query_string": "MVQRWLYSTN", "runs": _runs(modes=("sw",), k=1)},
    # --- BASELINE config 1: P07327 vs the protein files of tests/testdata (concatenated) ---
    {"name": "config1_P07327_testdata", "symtype": AMINOACID, "scoring": "blosum62", "gaps": [-11, -1], "db": "testdata_proteins.fas",
     "query": "P07327", "runs": _runs(k=10) + [["sw", 16, 10, A], ["nw", 16, 10, A]]},
    # --- seeded synthetic databases ---
    {"name": "synth_protein_2000_P07327", "symtype": AMINOACID, "scoring": "blosum62", "gaps": [-11, -1],
     "db": ["protein", 2000, 11, "P07327"], "query": "P07327", "runs": _runs(k=40) + [["sw", 8, 12, A], ["nw", 16, 12, A]]},
    {"name": "synth_protein_2000_chunk100", "symtype": AMINOACID, "scoring": "blosum62", "gaps": [-11, -1], "chunk": 100,
     "db": ["protein", 2000, 11, "P07327"], "query": "P07327", "runs": _runs(k=40)},
    {"name": "synth_protein_2000_chunk5000", "symtype": AMINOACID, "scoring": "blosum62", "gaps": [-11, -1], "chunk": 5000,
     "db": ["protein", 2000, 11, "P07327"], "query": "P07327", "runs": _runs(k=40)},   # > 2048: host scheduler emulation
    {"name": "synth_protein_edge_O74807", "symtype": AMINOACID, "scoring": "blosum50", "gaps": [-3, -1],
     "db": ["protein_edge", 1500, 5, "O74807"], "query": "O74807", "runs": _runs(k=30) + [["sw", 16, 2000, S], ["sw", 16, 8, A]]},
    {"name": "synth_protein_long_query", "symtype": AMINOACID, "scoring": "blosum62", "gaps": [-11, -1],
     "db": ["protein", 600, 3, "P33450"], "query": "P33450", "runs": _runs(widths=(16, 8), k=10) + [["sw", 16, 3, A]]},
    {"name": "synth_protein_pam250", "symtype": AMINOACID, "scoring": "pam250", "gaps": [-10, -2],
     "db": ["protein", 800, 9, "P27895"], "query": "P27895", "runs": _runs(widths=(16,), k=10) + [["nw", 16, 4, A]]},
    # 16-bit NW overflow: match 5 x 3 kbp query against planted near-copies leaves the int16 range
    {"name": "synth_nucleotide_long_query", "symtype": NUCLEOTIDE, "scoring": [12, -4], "gaps": [-4, -2],
     "db": ["nucleotide_long", 150, 6, None], "query_string_fn": "nucleotide_3000", "runs": _runs(k=20) + [["nw", 16, 4, A]]},
    {"name": "synth_nucleotide_800", "symtype": NUCLEOTIDE, "scoring": [5, -4], "gaps": [-4, -2],
     "db": ["nucleotide", 800, 21, None], "query": "one_seq.fas", "runs": _runs(k=20) + [["sw", 16, 6, A], ["nw", 16, 6, A]]},
]

# Every mode is compared bit-exactly.  (Until the saturating NW tiers were emulated on the GPU - nw_sat.cuh - NW
# searches entered at 8 bit could only be compared tie-aware; the hook is kept for the multi-GPU worker.)
TIE_AWARE = lambda mode, bw: False
