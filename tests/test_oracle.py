"""CPU tests of the oracle: the C restatement (oracle/restate.c) against the reference's own golden
vectors, against the committed fixtures generated from the compiled reference
(tests/golden/searches.json) and, where oracle/_ref is built, against the compiled reference live."""
import json
import os

import numpy as np
import pytest

import cases

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MATS = json.load(open(os.path.join(GOLD, "matrices.json")))
SEARCHES = json.load(open(os.path.join(GOLD, "searches.json")))


def matrix_of(scoring):
    key = scoring if isinstance(scoring, str) else "const_%d_%d" % tuple(scoring)
    return np.array(MATS[key], dtype=np.int64).reshape(32, 32)


def load_case(port, case):
    nuc = case["symtype"] == cases.NUCLEOTIDE
    recs = port.read_fasta(cases.db_path(case["db"]))
    seqs = [port.map_db(s, nuc) for _, s in recs]
    if "query_string_fn" in case:
        from libssa_b200 import synth
        q = port.map_query(synth.random_nucleotide(5, 3000, iupac_rate=0.02), nuc)
    elif "query_string" in case:
        q = port.map_query(case["query_string"], nuc)
    else:
        q = port.map_query(port.read_fasta(os.path.join(cases.DATA, case["query"]))[0][1], nuc)
    return seqs, q, matrix_of(case["scoring"])


# ---- expected values typed from the reference's own tests ----

def test_golden_bigger_databases(port):
    """reference tests/test_bigger_databases.c:70-92"""
    case = SEARCHES[0]["case"]
    seqs, q, M = load_case(port, case)
    nw = [(-88, 908), (-92, 1050), (-94, 938), (-98, 378), (-102, 75), (-104, 23), (-106, 1229), (-110, 1148), (-110, 612), (-112, 1016)]
    sw = [(20, 1177), (18, 384), (18, 277), (18, 215), (18, 214), (18, 210), (18, 151), (18, 126), (18, 114), (18, 101)]
    for bw in (64, 16, 8):
        assert port.search(seqs, q, M, -4, -2, True, 10, bw) == nw
        assert port.search(seqs, q, M, -4, -2, False, 10, bw) == sw


def test_golden_libssa_alignments(port):
    """reference tests/test_libssa.c:48-93"""
    case = SEARCHES[1]["case"]
    seqs, q, M = load_case(port, case)
    hits = port.search(seqs, q, M, -4, -2, False, 5, 64)
    assert [h[1] for h in hits] == [877, 847, 753, 565, 398] and all(h[0] == 91 for h in hits)
    assert port.cigar(q, seqs[877], M, -4, -2, False)[1] == "2MD2MIM2I3MD3MD4M2D8M5I3M5IM3I7MI7M3I5MI3M"
    hits = port.search(seqs, q, M, -4, -2, True, 5, 64)
    assert hits == [(33, 1050), (28, 908), (24, 938), (21, 378), (12, 75)]
    assert port.cigar(q, seqs[1050], M, -4, -2, True)[1] == "2MI2MI6M3I6MD3M19I4M3IMI3M5I3M5IM3I7MI7M4IMIMI4MI2M7I"


def test_golden_searcher_short(port):
    """reference tests/algo/test_searcher.c:146-227, :64-135"""
    M = matrix_of([1, -1])
    seqs = [port.map_db(s, True) for _, s in port.read_fasta(os.path.join(cases.DATA, "short_nuc_db.fas"))]
    q = port.map_query("ATGCAAATTT", True)
    nw = [(-1, 0), (-3, 3), (-4, 7), (-4, 6), (-4, 4), (-5, 2), (-6, 5), (-6, 1)]
    sw = [(4, 3), (4, 1), (4, 0), (3, 7), (3, 5), (3, 4), (3, 2), (2, 6)]
    for bw in (64, 16, 8):
        assert port.search(seqs, q, M, -1, -1, True, 8, bw) == nw
        assert port.search(seqs, q, M, -1, -1, False, 8, bw) == sw
    d = port.map_db("AATG", True)
    a = port.map_query("AT", True)
    assert port.sw_score(d, a, M, -1, -1) == 2 and port.nw_score(d, a, M, -1, -1) == -2


def test_golden_test_fas(port):
    """reference tests/algo/64/test_search_64.c:68-114, tests/algo/8/test_8_simd_avx2_nw.c:78-95"""
    M = matrix_of([1, -1])
    seqs = [port.map_db(s, True) for _, s in port.read_fasta(os.path.join(cases.DATA, "test.fas"))]
    q = port.map_query(port.read_fasta(os.path.join(cases.DATA, "one_seq.fas"))[0][1], True)
    sw = port.search(seqs, q, M, -1, -1, False, 5, 64)
    assert [s for s, _ in sw] == [8] * 5 and sw[0][1] == 4
    nw = port.search(seqs, q, M, -1, -1, True, 5, 64)
    assert [s for s, _ in nw] == [-43, -50, -52, -52, -147] and nw[0][1] == 4
    ovf = [port.nw_simd_channel(s, q, M, -1, -1, 8)[1] >= 0 for s in seqs]
    assert sum(ovf) == 4            # "4 of 5 overflow" in the 8-bit NW tier


def test_golden_protein(port):
    """reference tests/algo/test_searcher.c:327-546, tests/algo/8/test_8_simd_avx2_sw.c:104-124"""
    q = port.map_query("HPEVYILIIPGFGIISHVVSTYSKKPVFGEISMVYAMASIGLLGFLVWSHHMYIVGLDADTRAYFTSATMIIAIPTGIKI", False)
    d = port.map_db(port.read_fasta(os.path.join(cases.DATA, "short_AA.fas"))[0][1], False)
    assert port.sw_score(d, q, matrix_of("blosum62"), -1, -1) == 103
    assert port.nw_score(d, q, matrix_of("blosum62"), -1, -1) == 82
    assert port.sw_score(d, q, matrix_of([1, -1]), -1, -1) == 3
    assert port.nw_score(d, q, matrix_of([1, -1]), -1, -1) == -87
    s = port.map_db(port.read_fasta(os.path.join(cases.DATA, "NP_009305.1.fas"))[0][1], False)
    assert port.sw_score(s, s, matrix_of([1, -1]), -1, -1) == 534
    assert port.sw_score(s, s, matrix_of([127, -1]), -1, -1) == 67818
    assert port.nw_score(s, s, matrix_of([127, -1]), -1, -1) == 67818
    sc, b8, b16 = port.sw_score_blocks(s, s, matrix_of([127, -1]), -1, -1)
    assert sc == 67818 and b8 >= 0 and b16 >= 0          # 8 -> 16 -> 64 cascade
    q2 = port.map_query("MVQRWLYSTN", False)
    assert port.sw_score(d, q2, matrix_of([30, -1]), -1, -1) == 300


def test_golden_cigar_units(port):
    """reference tests/algo/test_cigar.c:55-134, tests/algo/test_align.c:49-75, tests/algo/test_aligner.c:85-218"""
    M = matrix_of([1, -1])
    nt = lambda s: port.map_query(s, True)
    assert port.cigar(nt("AATG"), nt("AT"), M, -1, -1, True)[1] == "MDMD"
    assert port.cigar(nt("AT"), nt("AATG"), M, -1, -1, True)[1] == "MIMI"
    assert port.cigar(nt("TACGGGTAT"), nt("GGACGTACG"), M, -1, -1, True)[1] == "MI3M2D3MI"
    assert port.cigar(nt("AATG"), nt("AT"), M, -1, -1, False)[1] == "2M"
    assert port.cigar(nt("AT"), nt("AATG"), M, -1, -1, False) == ((0, 1, 1, 2), "2M")
    assert port.cigar(nt("TACGGGTAT"), nt("GGACGTACG"), M, -1, -1, False)[1] == "4M"
    assert port.region_local(nt("AATC"), nt("AT"), M, -1, -1) == (1, 2, 0, 1)
    db = port.map_db(port.read_fasta(os.path.join(cases.DATA, "one_seq_db.fas"))[0][1], True)
    assert port.cigar(nt("ATGCAAAAA"), db, M, -1, -1, False)[1] == "9M"
    q = nt(port.read_fasta(os.path.join(cases.DATA, "one_seq.fas"))[0][1])
    d0 = port.map_db(port.read_fasta(os.path.join(cases.DATA, "test.fas"))[0][1], True)
    assert port.cigar(q, d0, M, -1, -1, False) == ((10, 34, 31, 54), "20MD4M")


def test_heap_behaviour(port):
    """reference tests/util/test_minheap.c:63-90: 20 ascending inserts into k=5 -> 19..15"""
    scores = np.arange(20, dtype=np.int64)
    assert [s for s, _ in port.topk_replay(scores, np.arange(20), 5)] == [19, 18, 17, 16, 15]


# ---- the committed fixtures (generated from the compiled reference by tests/golden/make_golden.py) ----

@pytest.mark.parametrize("entry", SEARCHES, ids=[e["case"]["name"] for e in SEARCHES])
def test_port_matches_fixtures(port, entry):
    case, expected = entry["case"], entry["result"]
    seqs, q, M = load_case(port, case)
    if sum(len(s) for s in seqs) * len(q) > 3e8:
        pytest.skip("too large for the scalar port")
    gO, gE = case["gaps"]
    for (mode, bw, k, align), exp in zip(case["runs"], expected):
        got = port.search(seqs, q, M, gO, gE, mode == "nw", k, bw, case.get("chunk", 1000))
        assert got == [(e[0], e[1]) for e in exp], (case["name"], mode, bw)
        if align == cases.COMPUTE_ALIGNMENT:
            for e in exp:
                region, cig = port.cigar(q, seqs[e[1]], M, gO, gE, mode == "nw")
                assert cig == e[2] and region == (e[3], e[4], e[5], e[6]), (case["name"], mode, e[1])


# ---- live comparison with the compiled reference on fresh random inputs ----

def test_port_matches_live_reference(port, ref, tmp_path):
    from libssa_b200 import synth
    rng = np.random.default_rng(99)
    for trial in range(3):
        n = int(rng.integers(50, 400))
        q = synth.random_protein(rng, int(rng.integers(20, 300)))
        seqs = synth.protein_db(n, int(rng.integers(1, 1 << 30)), query=q, homolog_frac=0.05, n_dup=10, mu=4.5, sigma=0.6, min_len=1)
        path = tmp_path / ("t%d.fas" % trial)
        synth.write_fasta(str(path), seqs)
        case = {"symtype": cases.AMINOACID, "scoring": "blosum62", "gaps": [-5, -2], "db": str(path), "query_string": q,
                "runs": [[m, w, 15, cases.COMPUTE_ALIGNMENT] for m in ("sw", "nw") for w in (64, 16, 8)]}
        expected = cases.run_case(ref, case)
        codes = [port.map_db(s, False) for s in seqs]
        qc = port.map_query(q, False)
        M = matrix_of("blosum62")
        for (mode, bw, k, _), exp in zip(case["runs"], expected):
            got = port.search(codes, qc, M, -5, -2, mode == "nw", k, bw)
            assert got == [(e[0], e[1]) for e in exp], (trial, mode, bw)
            for e in exp[:5]:
                region, cig = port.cigar(qc, codes[e[1]], M, -5, -2, mode == "nw")
                assert (cig, region) == (e[2], (e[3], e[4], e[5], e[6]))


def test_port_nw_narrow_tier_scores_match_live_reference(port, ref, tmp_path):
    """NW entered at 8 bit with a 2-residue query: the reference's 8-bit tier accepts records of ~130-160 columns with a score
    that is too high (its injected top boundary saturates silently) and reports it; the port does the same (search())."""
    from libssa_b200 import synth
    seqs = synth.nucleotide_db(300, 4242, mu=4.9, sigma=0.25, min_len=60, max_len=400, iupac_rate=0.01)
    path = tmp_path / "nt.fas"
    synth.write_fasta(str(path), seqs)
    case = {"symtype": cases.NUCLEOTIDE, "scoring": [5, -4], "gaps": [-1, -1], "db": str(path), "query_string": "AA",
            "runs": [["nw", w, len(seqs), cases.COMPUTE_SCORE] for w in (64, 16, 8)]}
    expected = cases.run_case(ref, case)
    codes = [port.map_db(s, True) for s in seqs]
    qc = port.map_query("AA", True)
    M = matrix_of([5, -4])
    for (mode, bw, k, _), exp in zip(case["runs"], expected):
        assert port.search(codes, qc, M, -1, -1, True, k, bw) == [(e[0], e[1]) for e in exp], bw
    assert sorted(e[0] for e in expected[2]) != sorted(e[0] for e in expected[0])   # 8 bit really reports other scores than 64 bit
