"""Workloads for searches that multiply the query and/or the database into strands and reading frames
(reference src/algo/searcher.c:42-90 query frames, src/db_adapter.c:46-139 database entries).

`workload()` builds a small database with homologs planted on both strands and in several frames, exact
duplicates (ties) and sequences long enough to overflow the 8-bit tier; `entries_via_reference()` has the
compiled reference translate / reverse-complement every record, so the CPU tests get the entry sequences
without restating the translation a second time.
"""
import ctypes as C

import numpy as np

import cases

NUCLEOTIDE, AMINOACID, TRANS_QUERY, TRANS_DB, TRANS_BOTH = 0, 1, 2, 3, 4
FORWARD, COMPLEMENTARY, BOTH = 1, 2, 3

MODES = [(NUCLEOTIDE, COMPLEMENTARY), (NUCLEOTIDE, BOTH),
         (TRANS_QUERY, FORWARD), (TRANS_QUERY, COMPLEMENTARY), (TRANS_QUERY, BOTH),
         (TRANS_DB, FORWARD), (TRANS_DB, COMPLEMENTARY), (TRANS_DB, BOTH),
         (TRANS_BOTH, FORWARD), (TRANS_BOTH, COMPLEMENTARY), (TRANS_BOTH, BOTH)]
MODE_IDS = ["%s-%s" % (["nt", "aa", "tq", "tdb", "tboth"][m], ["", "fwd", "compl", "both"][s]) for m, s in MODES]

_CODON = {}
for i, a in enumerate("TCAG"):
    for j, b in enumerate("TCAG"):
        for k, c in enumerate("TCAG"):
            _CODON[a + b + c] = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"[16 * i + 4 * j + k]


def revcomp(s):
    return s[::-1].translate(str.maketrans("ACGT", "TGCA"))


def translate(s, frame=0):
    """standard code, test-data generation only (what is checked comes from the reference)"""
    return "".join(_CODON[s[i:i + 3]] for i in range(frame, len(s) - 2, 3)).replace("*", "X")


def _rand_dna(rng, n):
    return "".join(rng.choice(list("ACGT"), size=n))


def _mutate(rng, s, rate, alphabet):
    out = list(s)
    for i in range(len(out)):
        if rng.random() < rate:
            out[i] = alphabet[rng.integers(len(alphabet))]
    return "".join(out)


def workload(symtype, seed=1):
    """-> (db records (text), query text).  DNA database unless the mode searches proteins (TRANS_QUERY)."""
    rng = np.random.default_rng(seed)
    qd = _rand_dna(rng, 330)
    qd = qd[:200] + "A" * 45 + qd[245:]                                 # a poly-K stretch in every forward frame
    aa = "ACDEFGHIKLMNPQRSTVWY"
    recs = []
    if symtype == TRANS_QUERY:
        for f in range(3):
            recs.append(translate(qd, f))
            recs.append(_mutate(rng, translate(revcomp(qd), f), 0.1, aa))
        recs.append(translate(qd, 1))                                  # exact duplicate -> tie
        for _ in range(40):
            recs.append("".join(rng.choice(list(aa), size=int(rng.integers(8, 160)))))
        recs.insert(5, "")                                              # empty record: skipped (db_adapter.c:231-233)
        for r in (3, 17, 17, 30):                                       # more ties among unrelated records
            recs.append(recs[r])
        query = qd
    else:
        recs.append(qd)
        recs.append(revcomp(qd))
        recs.append("A" + _mutate(rng, qd, 0.05, "ACGT"))               # other frames
        recs.append("AC" + _mutate(rng, revcomp(qd), 0.05, "ACGT") + "G")
        recs.append(_rand_dna(rng, 120) + qd[60:270] + _rand_dna(rng, 77))
        recs.append(qd)                                                 # exact duplicate -> tie
        for _ in range(40):
            recs.append(_rand_dna(rng, int(rng.integers(24, 420))))
        recs.insert(7, "")
        for r in (9, 12, 12, 31):
            recs.append(recs[r])
        recs.append(_mutate(rng, recs[2], 0.02, "ACGT"))
        # entries of ONE record that tie: a reverse-complement palindrome (both strands identical) and a homopolymer of
        # length 3n+2 (all three frames identical) - their order is decided by the reference's sorts alone
        recs.insert(20, qd[:100] + revcomp(qd[:100]))
        recs.insert(33, "A" * 62)
        query = translate(qd, 0) if symtype == TRANS_DB else qd
    return recs, query


def scoring(symtype):
    """(matrix spec, gapO, gapE; penalties are passed negative, as the reference's tests do): nucleotide searches score 5/-4, protein-space searches BLOSUM62"""
    if symtype == NUCLEOTIDE:
        return (5, -4), -6, -2
    return "blosum62", -7, -1


def configure(lib, symtype, strands, gencode=1):
    lib.set_output_mode(0)
    lib.set_thread_count(1)
    lib.init_symbol_translation(symtype, strands, gencode, gencode)
    spec, go, ge = scoring(symtype)
    if isinstance(spec, tuple):
        lib.init_constant_scores(*spec)
    else:
        lib.init_score_matrix(cases.MATRIX_BUILDIN, spec)
    lib.init_gap_penalties(go, ge)


class _Seq(C.Structure):
    _fields_ = [("seq", C.c_void_p), ("len", C.c_size_t)]


class _Query(C.Structure):  # reference src/libssa_datatypes.h:100-107
    _fields_ = [("nt", _Seq * 2), ("aa", _Seq * 6), ("header", C.c_char_p), ("headerlen", C.c_size_t),
                ("map", C.c_void_p), ("sym", C.c_void_p)]


def _codes(seq):
    return np.frombuffer(C.string_at(seq.seq, seq.len), dtype=np.uint8).copy()


def query_frames(lib, symtype, strands, text, gencode=1):
    """[(codes, strand, frame)] in the reference's frame order (src/algo/searcher.c:42-90), read through `lib`"""
    lib.init_symbol_translation(symtype, strands, gencode, gencode)
    q = lib.init_sequence_fasta(cases.READ_FROM_STRING, text)
    s = C.cast(q, C.POINTER(_Query)).contents
    out = []
    if symtype == NUCLEOTIDE:
        for st in range(2):
            if (st + 1) & strands:
                out.append((_codes(s.nt[st]), st, 0))
    elif symtype in (AMINOACID, TRANS_DB):
        out.append((_codes(s.aa[0]), 0, 0))
    else:
        for st in range(2):
            if (st + 1) & strands:
                for f in range(3):
                    out.append((_codes(s.aa[3 * st + f]), st, f))
    lib.free_sequence(q)
    return out


def buffer_max(symtype, strands):
    if symtype == NUCLEOTIDE:
        return 2 if strands & 2 else 1
    if symtype in (TRANS_DB, TRANS_BOTH):
        return 6 if strands == BOTH else 3
    return 1


def entry_fields(symtype, strands, b):
    """(strand, frame) the reference stores with entry b of a record (src/db_adapter.c:57-104)"""
    if symtype == NUCLEOTIDE:
        return b, 0
    if symtype in (TRANS_DB, TRANS_BOTH):
        return (b // 3, b % 3) if strands == BOTH else (strands, b)
    return 0, 0


def entries_via_reference(ref, symtype, strands, recs, gencode=1):
    """code strings of every database entry, slot = record * buffer_max + b, produced by the reference's own
    query reader (reverse complement: src/query.c, translation: src/util/util_sequence.c:332-382)"""
    bm = buffer_max(symtype, strands)
    out = []
    for r in recs:
        if not r:
            out.extend([np.zeros(0, dtype=np.uint8)] * bm)
            continue
        if symtype == NUCLEOTIDE:
            fr = query_frames(ref, NUCLEOTIDE, BOTH, r)
            out.extend([fr[b][0] for b in range(bm)])
        elif symtype in (TRANS_DB, TRANS_BOTH):
            fr = query_frames(ref, TRANS_QUERY, BOTH, r, gencode)
            if strands == BOTH:
                out.extend([fr[b][0] for b in range(6)])
            else:
                out.extend([fr[3 * (strands - 1) + b][0] for b in range(3)])
        else:
            fr = query_frames(ref, AMINOACID, FORWARD, r)
            # database symbols are mapped, not filtered; the workloads hold known symbols only, so both agree
            out.append(fr[0][0])
    return out


def hit_key(h):
    return (h.score, h.db_id, h.q_strand, h.q_frame, h.db_strand, h.db_frame)


def hit_record(h):
    """everything alignment_t holds for one hit, JSON-able (the database sequence as a checksum)"""
    import zlib
    return {"score": h.score, "db_id": h.db_id, "q_strand": h.q_strand, "q_frame": h.q_frame,
            "db_strand": h.db_strand, "db_frame": h.db_frame, "db_len": h.db_len,
            "db_crc": zlib.crc32(h.db_seq) if h.db_seq is not None else None, "q_len": h.q_len,
            "cigar": h.cigar, "region": [h.q_start, h.q_end, h.d_start, h.d_end]}


def run_reference_worker(symtype, strands, out, entries=False):
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, os.path.join(here, "frames_ref_worker.py"), str(symtype), str(strands), str(out)]
    if entries:
        cmd.append("--entries")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]


def frame_fields(symtype, strands):
    """(strand, frame) of every query frame, in the reference's order (src/algo/searcher.c:42-90)"""
    if symtype == NUCLEOTIDE:
        return [(s, 0) for s in range(2) if (s + 1) & strands]
    if symtype in (AMINOACID, TRANS_DB):
        return [(0, 0)]
    return [(s, f) for s in range(2) if (s + 1) & strands for f in range(3)]


def check_product_mode(prod, port, mode, tmp_path):
    """Runs every search of one mode through the product's libssa.h calls (GPU) and compares with the fixture
    tests/golden/frames.json; used one-GPU (tests/test_gpu_frames.py) and sharded (tests/dist_gpu_worker.py)."""
    import json
    import os
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "frames.json")))
    symtype, strands = mode
    gold = gold[MODE_IDS[MODES.index(mode)]]
    recs, query = workload(symtype)
    db = os.path.join(str(tmp_path), "db_%d_%d.fas" % (symtype, strands))
    with open(db, "w") as f:
        for i, r in enumerate(recs):
            f.write(">r%d\n%s\n" % (i, r))
    configure(prod, symtype, strands)
    prod.set_chunk_size(gold["chunk"])
    prod.init_db(db)
    q = prod.init_sequence_fasta(cases.READ_FROM_STRING, query)
    qfields = frame_fields(symtype, strands)
    qf = query_frames(prod, symtype, strands, query)
    M = prod.score_matrix()
    _, go, ge = scoring(symtype)
    try:
        for nw in (False, True):
            for bits in (64, 16, 8):
                want = gold["searches"]["%s%d" % ("nw" if nw else "sw", bits)]
                for align in (cases.COMPUTE_SCORE, cases.COMPUTE_ALIGNMENT):
                    hits = (prod.nw_align if nw else prod.sw_align)(q, gold["k"], bits, align)
                    got = [[h.score, h.db_id, h.db_strand, h.db_frame, qfields.index((h.q_strand, h.q_frame))] for h in hits]
                    assert got == want["phase"], (mode, nw, bits, align)
                st = prod.stats()
                if not nw or st.overflow_8 or st.overflow_16:   # NW: only known when the saturating tiers had to be emulated
                    assert (st.overflow_8, st.overflow_16) == (want["overflow_8"], want["overflow_16"]), (mode, nw, bits)
                if "hits" in want:
                    assert [hit_record(h) for h in hits] == want["hits"], (mode, nw, bits)
                else:
                    # no reference record to compare with: the alignment must be the oracle's for the sequences returned
                    for h in hits[:8]:
                        d = np.frombuffer(h.db_seq, dtype=np.uint8)
                        a = qf[qfields.index((h.q_strand, h.q_frame))][0]
                        exp = port.cigar(a, d, M, go, ge, nw)
                        assert ([h.q_start, h.q_end, h.d_start, h.d_end], h.cigar) == (list(exp[0]), exp[1]), (mode, nw, bits)
    finally:
        prod.free_sequence(q)
        prod.set_chunk_size(1000)
        prod.init_symbol_translation(AMINOACID, 1, 1, 1)
