"""GPU parity tests: the CUDA path, called through the C-ABI (libssa.h entry points of libssa_b200.so),
against (1) the committed fixtures generated from the compiled reference, (2) the compiled reference live
on fresh random inputs, (3) the oracle port's exact scores, and (4) size-independent properties on a
database of BASELINE scale.  Integer work: every comparison is bit-exact."""
import json
import os

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SEARCHES = json.load(open(os.path.join(GOLD, "searches.json")))
MATS = json.load(open(os.path.join(GOLD, "matrices.json")))


def matrix_of(scoring):
    key = scoring if isinstance(scoring, str) else "const_%d_%d" % tuple(scoring)
    return np.array(MATS[key], dtype=np.int64).reshape(32, 32)


def assert_same_hits(got, exp, mode, bw, what):
    if cases.TIE_AWARE(mode, bw):
        # scores identical; identical records above the cut-off score
        assert [g[0] for g in got] == [e[0] for e in exp], what
        if exp:
            cut = exp[-1][0]
            assert [g[:2] for g in got if g[0] > cut] == [e[:2] for e in exp if e[0] > cut], what
    else:
        assert got == exp, what


@pytest.mark.parametrize("entry", SEARCHES, ids=[e["case"]["name"] for e in SEARCHES])
def test_fixtures_through_c_abi(prod, entry):
    case, expected = entry["case"], entry["result"]
    got = cases.run_case(prod, case)
    for run, g, e in zip(case["runs"], got, expected):
        assert_same_hits(g, e, run[0], run[1], (case["name"], run))


def test_live_reference_random_inputs(prod, ref, tmp_path):
    """differential test against the compiled reference on random inputs; LIBSSA_B200_FUZZ_TRIALS=n runs a longer campaign"""
    from libssa_b200 import synth
    trials = int(os.environ.get("LIBSSA_B200_FUZZ_TRIALS", "8"))
    rng = np.random.default_rng(2024 if trials == 8 else int(os.environ.get("LIBSSA_B200_FUZZ_SEED", "1")))
    for trial in range(trials):
        n = int(rng.integers(1, 3000))
        qlen = int(rng.choice([1, 2, 17, 33, 128, 129, 250, 700, 1300, 2100]))
        q = synth.random_protein(rng, qlen)
        seqs = synth.protein_db(n, int(rng.integers(1, 1 << 30)), query=q, homolog_frac=0.03, n_dup=min(20, n // 3),
                                mu=float(rng.choice([3.0, 5.0])), sigma=0.7, min_len=0, max_len=4000)
        path = tmp_path / ("live%d.fas" % trial)
        synth.write_fasta(str(path), seqs)
        scoring = ["blosum62", "blosum45", "pam30", [4, -3]][trial % 4]
        gaps = [[-11, -1], [-3, -1], [-5, -5], [0, -2]][trial % 4]
        symtype = cases.AMINOACID
        if trial % 5 == 4:   # nucleotide searches (constant scoring, IUPAC codes in query and database)
            symtype = cases.NUCLEOTIDE
            q = synth.random_nucleotide(int(rng.integers(1, 1 << 30)), qlen, iupac_rate=0.03)
            seqs = synth.nucleotide_db(min(n, 800), int(rng.integers(1, 1 << 30)), query=q, n_homologs=min(5, n), mu=float(rng.choice([4.0, 6.0])),
                                       sigma=0.6, min_len=1, max_len=5000, iupac_rate=0.01)
            synth.write_fasta(str(path), seqs)
            scoring = [[5, -4], [2, -2], [1, -1]][trial % 3]
            gaps = [[-4, -2], [-1, -1], [-10, -1]][(trial // 5) % 3]
        case = {"symtype": symtype, "scoring": scoring, "gaps": gaps, "db": str(path), "query_string": q,
                "chunk": [1000, 64][trial % 2],
                "runs": [[m, w, 30, cases.COMPUTE_SCORE] for m in ("sw", "nw") for w in (64, 16, 8)]
                        + [["sw", 16, 6, cases.COMPUTE_ALIGNMENT], ["nw", 16, 6, cases.COMPUTE_ALIGNMENT]]}
        # The reference corrupts its heap when it is asked to ALIGN a Smith-Waterman hit of score 0 (no region: the -1 sentinels of
        # src/algo/align.c:86-139 become indices; seen with a 1-residue query) - there is no oracle for that, so the
        # alignment run is dropped when the score pass reports such a hit among the first six.
        score_runs = [r for r in case["runs"] if r[3] == cases.COMPUTE_SCORE]
        exp = cases.run_case(ref, dict(case, runs=score_runs))
        sw16 = exp[[i for i, r in enumerate(score_runs) if r[0] == "sw" and r[1] == 16][0]]
        align_runs = [r for r in case["runs"] if r[3] != cases.COMPUTE_SCORE and not (r[0] == "sw" and any(h[0] <= 0 for h in sw16[:6]))]
        exp = exp + cases.run_case(ref, dict(case, runs=align_runs))
        case["runs"] = score_runs + align_runs
        got = cases.run_case(prod, case)
        for run, g, e in zip(case["runs"], got, exp):
            assert_same_hits(g, e, run[0], run[1], ("trial", trial, n, qlen, scoring, gaps, run))


def test_nw_narrow_tier_reports_saturated_scores(prod, ref, tmp_path):
    """The reference's 8-bit NW tier injects a top boundary H(-1, j) that saturates silently: a 2-residue query against records of
    more than ~125 columns finishes the tier unflagged with a score that is too high (-118 instead of -139), and the reference
    reports it.  The library substitutes the tier's own result (nw_sat.cuh) for the records a narrow tier accepts.  Found by the
    differential fuzz campaign (seed 20261019, trial 9)."""
    from libssa_b200 import synth
    seqs = synth.nucleotide_db(400, 4242, mu=4.9, sigma=0.25, min_len=60, max_len=400, iupac_rate=0.01)
    path = tmp_path / "nt.fas"
    synth.write_fasta(str(path), seqs)
    differs = 0
    for q, gaps in (("AA", (-1, -1)), ("ACG", (-2, -1)), ("T", (0, -1))):
        lists = {}
        for lib in (ref, prod):
            lib.set_output_mode(0)
            if lib is ref:
                lib.set_thread_count(1)
            lib.init_symbol_translation(cases.NUCLEOTIDE, 1, 1, 1)
            lib.init_constant_scores(5, -4)
            lib.init_gap_penalties(*gaps)
            lib.set_chunk_size(1000)
            lib.init_db(str(path))
            qh = lib.init_sequence_fasta(cases.READ_FROM_STRING, q)
            lists[lib is ref] = {bw: [(h.score, h.db_id) for h in lib.nw_align(qh, len(seqs), bw, cases.COMPUTE_SCORE)] for bw in (8, 16, 64)}
            lib.free_sequence(qh)
        assert lists[False] == lists[True], (q, gaps)
        differs += sorted(lists[True][8]) != sorted(lists[True][64])
    assert differs > 0   # the case is real: at 8 bit the reference itself reports other scores than at 64 bit


@pytest.mark.parametrize("qlen", [1, 3, 64, 127, 128, 130, 255, 384, 385, 500, 768, 1024, 1025, 1100, 1500, 2048, 2100, 3100])
def test_exact_scores_vs_oracle_port(prod, port, tmp_path, qlen):
    """every kernel shape (rows per lane 4..32, one and several passes, remainder passes on narrow groups), both tiers, SW and NW"""
    from libssa_b200 import synth
    rng = np.random.default_rng(qlen)
    q = synth.random_protein(rng, qlen)
    seqs = synth.protein_db(300, qlen + 1, query=q, homolog_frac=0.1, n_dup=10, mu=4.8, sigma=0.9, min_len=1, max_len=3000)
    path = tmp_path / "db.fas"
    synth.write_fasta(str(path), seqs)
    codes = [port.map_db(s, False) for s in seqs]
    qc = port.map_query(q, False)
    for scoring, gaps in (("blosum62", (-11, -1)), ([127, -1], (-1, -1)), ([2, -100], (-120, -7))):
        prod.init_symbol_translation(cases.AMINOACID, 1, 1, 1)
        if isinstance(scoring, str):
            prod.init_score_matrix(cases.MATRIX_BUILDIN, scoring)
        else:
            prod.init_constant_scores(*scoring)
        prod.init_gap_penalties(*gaps)
        prod.init_db(str(path))
        qh = prod.init_sequence_fasta(cases.READ_FROM_STRING, q)
        M = matrix_of(scoring) if isinstance(scoring, str) or "const_%d_%d" % tuple(scoring) in MATS else port.constant_matrix(*scoring)
        for nw in (False, True):
            ids, sc = prod.score_all(qh, nw)
            assert ids.tolist() == list(range(len(seqs)))
            exp = port.score_db(codes, qc, M, gaps[0], gaps[1], nw)
            bad = np.nonzero(sc != exp)[0]
            assert len(bad) == 0, (qlen, scoring, gaps, nw, bad[:5], sc[bad[:5]], exp[bad[:5]])
        prod.free_sequence(qh)


@pytest.mark.parametrize("G", [32, 16])
@pytest.mark.parametrize("R", [4, 8, 12, 16, 20, 24, 28, 32])
def test_full_warp_kernel_every_shape(prod, port, tmp_path, monkeypatch, R, G):
    """sw32_kernel<R, MULTIPASS, NW, G>: every rows-per-lane instantiation forced with LIBSSA_B200_SHAPE=G,R - arrays of 32 and of 16
    lanes (two chains per warp) - one pass (query fills the array, ends in its last / first / a middle row) and several passes
    (always full warps), SW and NW, pairs of equal / odd / even lengths"""
    from libssa_b200 import synth
    rng = np.random.default_rng(100 + R + G)
    monkeypatch.setenv("LIBSSA_B200_SHAPE", "%d,%d" % (G, R))
    base = synth.random_protein(rng, 3 * 32 * R)
    seqs = synth.protein_db(200, 7 * R, query=base[:32 * R], homolog_frac=0.15, n_dup=8, mu=4.6, sigma=0.9, min_len=1, max_len=2500)
    seqs += [base[:40], base[:40], base[:41], base[5:46], base[:1], base[:2], base[:32 * R], base[:32 * R + 1]]   # equal, odd/even neighbours
    path = tmp_path / "db.fas"
    synth.write_fasta(str(path), seqs)
    codes = [port.map_db(s, False) for s in seqs]
    M = matrix_of("blosum62")
    prod.init_symbol_translation(cases.AMINOACID, 1, 1, 1)
    prod.init_score_matrix(cases.MATRIX_BUILDIN, "blosum62")
    prod.init_gap_penalties(-11, -1)
    prod.init_db(str(path))
    for qlen in (G * R, G * R - R, G * R - R + 1, (G // 2 + 1) * R + 3, G * R + 1, 2 * 32 * R + R // 2 + 1, 3 * 32 * R):
        q = base[:qlen]
        qc = port.map_query(q, False)
        qh = prod.init_sequence_fasta(cases.READ_FROM_STRING, q)
        for nw in (False, True):
            _, sc = prod.score_all(qh, nw)
            exp = port.score_db(codes, qc, M, -11, -1, nw)
            bad = np.nonzero(sc != exp)[0]
            assert len(bad) == 0, (R, qlen, nw, bad[:5], sc[bad[:5]], exp[bad[:5]])
        prod.free_sequence(qh)


@pytest.mark.parametrize("tiers", ["LIBSSA_B200_STATIC_TIERS", "LIBSSA_B200_DEFAULT_TIERS"])
def test_static_and_speculative_tier_split(prod, port, tmp_path, monkeypatch, tiers):
    """match 127: every pair longer than 257 residues is beyond the static int16 bound; both ways of handling them (32-bit tier
    up front / packed tier + recomputation of the sequences that came near the limit) give the oracle's scores"""
    from libssa_b200 import synth
    monkeypatch.setenv(tiers, "1")
    rng = np.random.default_rng(77)
    q = synth.random_protein(rng, 700)
    seqs = synth.protein_db(400, 78, query=q, homolog_frac=0.2, n_dup=10, mu=5.5, sigma=0.8, min_len=1, max_len=3000)
    seqs[7] = q
    seqs[8] = q[:300] + q[350:]
    path = tmp_path / "db.fas"
    synth.write_fasta(str(path), seqs)
    codes = [port.map_db(s, False) for s in seqs]
    qc = port.map_query(q, False)
    prod.init_symbol_translation(cases.AMINOACID, 1, 1, 1)
    prod.init_constant_scores(127, -1)
    prod.init_gap_penalties(-1, -1)
    prod.init_db(str(path))
    qh = prod.init_sequence_fasta(cases.READ_FROM_STRING, q)
    _, sc = prod.score_all(qh, False)
    exp = port.score_db(codes, qc, port.constant_matrix(127, -1), -1, -1, False)
    assert (sc == exp).all() and sc.max() == 127 * 700
    assert prod.stats().wide_tier_seqs > 0
    prod.free_sequence(qh)


def test_nucleotide_long_query_wide_tier(prod, port, tmp_path):
    """BASELINE config 5 in small: 3 kbp query, match 5: NW leaves the 16-bit range for the long homologs"""
    from libssa_b200 import synth
    q = synth.random_nucleotide(5, 3000, iupac_rate=0.02)
    seqs = synth.nucleotide_db(120, 6, query=q, n_homologs=12, mu=6.5, sigma=0.6, min_len=20, max_len=6000)
    seqs[3] = q                                   # exact copy: SW = NW = 5 * (number of unambiguous-match columns...)
    path = tmp_path / "nt.fas"
    synth.write_fasta(str(path), seqs)
    prod.init_symbol_translation(cases.NUCLEOTIDE, 1, 1, 1)
    prod.init_constant_scores(5, -4)
    prod.init_gap_penalties(-4, -2)
    prod.init_db(str(path))
    qh = prod.init_sequence_fasta(cases.READ_FROM_STRING, q)
    codes = [port.map_db(s, True) for s in seqs]
    qc = port.map_query(q, True)
    M = matrix_of([5, -4])
    for nw in (False, True):
        _, sc = prod.score_all(qh, nw)
        assert (sc == port.score_db(codes, qc, M, -4, -2, nw)).all()
    hits = prod.sw_align(qh, 3, 16, cases.COMPUTE_ALIGNMENT)
    assert hits[0].db_id == 3 and hits[0].cigar == "3000M" and hits[0].score == sc.max()
    for h in hits:
        region, cig = port.cigar(qc, codes[h.db_id], M, -4, -2, False)
        assert (cig, region) == (h.cigar, (h.q_start, h.q_end, h.d_start, h.d_end))
    prod.free_sequence(qh)


def test_zero_score_alignment_fails_like_the_reference(tmp_path):
    """A Smith-Waterman hit of score 0 has no alignment region.  The reference reads an uninitialised region there and ends in
    fatal( "Internal error in align function." ) -> exit(1) (src/algo/align.c:39-131), or corrupts its heap (seen with a
    1-residue query).  The library keeps the defined half of that: the same message on stderr and exit code 1."""
    import subprocess
    import sys
    script = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "zero_probe.py")
    r = subprocess.run([sys.executable, script], capture_output=True, text=True, timeout=300)
    assert r.returncode == 1, (r.returncode, r.stdout, r.stderr)
    assert "[(13, 1), (0, 3), (0, 2), (0, 0)]" in r.stdout          # the score-only call is fine
    assert "Internal error in align function." in (r.stderr + r.stdout)


def test_result_record_contract(prod):
    """reference src/algo/aligner.c:62-93, src/libssa.h:85-117: code strings, lengths, COMPUTE_SCORE leaves the alignment empty,
    hitcount larger than the database shortens the list, order = score desc then ID desc"""
    case = SEARCHES[3]["case"]          # test.fas, 5 records
    q = cases.configure(prod, case)
    hits = prod.sw_align(q, 50, 16, cases.COMPUTE_SCORE)
    assert len(hits) == 5
    assert all(h.cigar is None and h.cigar_len == 0 and h.q_start == h.q_end == h.d_start == h.d_end == 0 for h in hits)
    assert all(len(h.db_seq) == h.db_len and max(h.db_seq) < 16 for h in hits)
    keys = [(h.score, h.db_id) for h in hits]
    assert keys == sorted(keys, reverse=True)
    hits = prod.nw_align(q, 2, 64, cases.COMPUTE_ALIGNMENT)
    assert len(hits) == 2 and all(h.cigar and h.cigar_len == len(h.cigar) for h in hits)
    assert prod.sw_align(q, 0, 16, cases.COMPUTE_SCORE) == []
    prod.free_sequence(q)


@pytest.fixture(scope="module")
def big_db(tmp_path_factory):
    """100 000 records / 33 M residues: BASELINE-scale database for the property tests (larger than L2 once packed)."""
    from libssa_b200 import synth
    q = cases.query_text("P07327")
    seqs = synth.protein_db(100000, 20150402, query=q)
    path = tmp_path_factory.mktemp("big") / "s100k.fas"
    synth.write_fasta(str(path), seqs)
    return str(path), seqs, q


def test_full_size_properties(prod, port, big_db, tmp_path):
    path, seqs, qtext = big_db
    prod.init_symbol_translation(cases.AMINOACID, 1, 1, 1)
    prod.init_score_matrix(cases.MATRIX_BUILDIN, "blosum62")
    prod.init_gap_penalties(-11, -1)
    prod.init_db(path)
    q = prod.init_sequence_fasta(cases.READ_FROM_FILE, os.path.join(cases.DATA, "P07327"))
    ids, sw = prod.score_all(q, False)
    _, nw = prod.score_all(q, True)
    assert len(ids) == len(seqs) and (ids == np.arange(len(seqs))).all()
    # (1) a sample of the records, exact against the oracle port
    rng = np.random.default_rng(1)
    sample = np.concatenate([rng.choice(len(seqs), 300, replace=False), np.argsort(sw)[-20:]])
    codes = [port.map_db(seqs[i], False) for i in sample]
    qc = port.map_query(qtext, False)
    M = matrix_of("blosum62")
    assert (sw[sample] == port.score_db(codes, qc, M, -11, -1, False)).all()
    assert (nw[sample] == port.score_db(codes, qc, M, -11, -1, True)).all()
    # (2) bounds that hold for every record: 0 <= SW <= 11 * min(len), NW <= SW, SW >= score of the best diagonal start
    lens = np.array([len(s) for s in seqs])
    assert (sw >= 0).all() and (sw <= 11 * np.minimum(lens, 375)).all() and (nw <= sw).all()
    # (3) exact duplicates of a record score identically (the generator plants 50)
    first = {}
    dup = 0
    for i, s in enumerate(seqs):
        if s in first:
            dup += 1
            assert sw[i] == sw[first[s]] and nw[i] == nw[first[s]]
        else:
            first[s] = i
    assert dup >= 40
    # (4) hit lists: identical for every entry bit width above the cut-off score, and equal to a replay of the score vector
    lists = {bw: [(h.score, h.db_id) for h in prod.sw_align(q, 100, bw, cases.COMPUTE_SCORE)] for bw in (64, 16, 8)}
    cut = lists[64][-1][0]
    for bw in (16, 8):
        assert [x for x in lists[bw] if x[0] > cut] == [x for x in lists[64] if x[0] > cut]
        assert [x[0] for x in lists[bw]] == [x[0] for x in lists[64]]
    order = prod.host_arrival_order(lens, 1000, 64)
    assert lists[64] == prod.host_replay(sw[order], order, 100)
    assert prod.stats().overflow_8 == int((sw >= 255).sum())          # counted during the bw = 8 search above
    # (4b) ties AT the cut-off force the arrival-order replay; its sparse form (first m arrivals + later records above the
    # heap's root, what databases beyond 262 144 records use) must return the same lists as the full form
    desc = np.sort(sw)[::-1]
    k_tie = next(k for k in range(60, 5000) if desc[k - 1] == desc[k])
    for k in (k_tie, 1, 0):
        full = {bw: [(h.score, h.db_id) for h in prod.sw_align(q, k, bw, cases.COMPUTE_SCORE)] for bw in (64, 16, 8)}
        assert full[64] == prod.host_replay(sw[order], order, k)
        d2h_full = prod.stats().d2h_bytes
        os.environ["LIBSSA_B200_REPLAY_PREFIX"] = "300"
        try:
            sparse = {bw: [(h.score, h.db_id) for h in prod.sw_align(q, k, bw, cases.COMPUTE_SCORE)] for bw in (64, 16, 8)}
            d2h_sparse = prod.stats().d2h_bytes
        finally:
            del os.environ["LIBSSA_B200_REPLAY_PREFIX"]
        assert sparse == full, k
        if k == k_tie:
            assert d2h_sparse < d2h_full / 4, (d2h_sparse, d2h_full)      # the sparse replay really ran
    # (5) permutation invariance: the same records in another provider order get the same scores
    perm = rng.permutation(len(seqs))
    from libssa_b200 import synth
    p2 = tmp_path / "perm.fas"
    synth.write_fasta(str(p2), [seqs[i] for i in perm])
    prod.init_db(str(p2))
    _, sw2 = prod.score_all(q, False)
    assert (sw2 == sw[perm]).all()
    prod.free_sequence(q)


def test_zz_exit_and_reinitialise(prod):
    """ssa_exit() releases matrices, database and all device state (reference src/libssa.c:266-271); a later search must
    rebuild everything from scratch and give the same result"""
    case = SEARCHES[0]["case"]
    before = cases.run_case(prod, case)
    prod.ssa_exit()
    after = cases.run_case(prod, case)
    assert before == after == SEARCHES[0]["result"]


def test_c_host_program_drop_in(prod, tmp_path):
    """A C program that includes only libssa.h, linked against libssa_b200.so, prints the hit list the Python mirror returns"""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["make", "-s", "-C", os.path.join(root, "examples")])
    db = cases.db_path(["protein", 2000, 11, "P07327"])
    qf = os.path.join(cases.DATA, "P07327")
    for mode, bw, align in (("sw", 16, "align"), ("nw", 8, "score")):
        out = subprocess.run([os.path.join(root, "examples", "search_example"), db, qf, mode, str(bw), "7", align],
                             capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr
        got = [l.split("\t") for l in out.stdout.strip().split("\n") if l and not l.startswith("libssa")]
        case = {"symtype": cases.AMINOACID, "scoring": "blosum62", "gaps": [-11, -1], "db": ["protein", 2000, 11, "P07327"], "query": "P07327",
                "runs": [[mode, bw, 7, 1 if align == "align" else 0]]}
        exp = cases.run_case(prod, case)[0]
        assert [(int(g[0]), int(g[1])) for g in got] == [(e[0], e[1]) for e in exp]
        if align == "align":
            assert [g[4] for g in got] == [e[2] for e in exp]


def test_cli_and_benchmark_driver_match_reference_build(prod, ref, tmp_path):
    """examples/ssa_cli (the reference example's option letters, reference src/libssa_example.c:44-196) and
    examples/ssa_benchmark (reference benchmark/src/benchmark_queries.c:32-79) are single C sources written against
    libssa.h; built once against libssa_b200.so and once against the compiled reference they must print the same
    hit lines."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["make", "-s", "-C", os.path.join(root, "examples")])
    refdir = os.path.join(root, "oracle", "_ref")
    for prog in ("ssa_cli", "ssa_benchmark"):
        subprocess.check_call(["gcc", "-O2", "-std=gnu99", "-I" + os.path.join(root, "include"), "-o", str(tmp_path / prog),
                               os.path.join(root, "examples", prog + ".c"), "-L" + refdir, "-lssa_ref", "-Wl,-rpath," + refdir])
    db = cases.db_path(["protein", 2000, 11, "P07327"])
    qf = os.path.join(cases.DATA, "P07327")

    def hit_lines(text):
        return [l for l in text.split("\n") if l and not l.startswith("libssa")]

    for args in (["-N", "1", "-O", "-3", "-E", "-1", "-M", "BLOSUM50", "-c", "10", "-i", qf, "-d", db, "-t", "SW", "-b", "16", "-s", "AVX2"],
                 ["-N", "1", "-O", "-11", "-E", "-1", "-M", "BLOSUM62", "-c", "6", "-i", qf, "-d", db, "-t", "NW", "-b", "8", "-s", "SSE41", "-a"]):
        outs = []
        for exe in (os.path.join(root, "examples", "ssa_cli"), str(tmp_path / "ssa_cli")):
            r = subprocess.run([exe] + args, capture_output=True, text=True, timeout=600)
            assert r.returncode == 0, r.stderr
            outs.append(hit_lines(r.stdout))
        assert outs[0] == outs[1] and outs[0][0].startswith("Nr of alignments: ")
    rows = []
    for exe in (os.path.join(root, "examples", "ssa_benchmark"), str(tmp_path / "ssa_benchmark")):
        r = subprocess.run([exe, "-n", "2", "-N", "1", "-b", "8,16", "-t", "SW,NW", "-d", db, qf, os.path.join(cases.DATA, "O74807")],
                           capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stderr
        lines = [l.split(",") for l in hit_lines(r.stdout)]
        assert lines[0][:4] == ["query", "qlen", "type", "bit_width"] and len(lines) == 1 + 2 * 2 * 2
        rows.append([l[:5] + l[9:] for l in lines[1:]])     # query, qlen, type, width, iterations, top score, top id
    assert rows[0] == rows[1]


def _key(h, align):
    return (h.score, h.db_id) + ((h.cigar, h.q_start, h.q_end, h.d_start, h.d_end) if align else ())


def test_batched_queries_equal_single_calls(prod, ref):
    """ssa_b200_align_batch: the pipelined batch returns, query by query, the lists of single
    sw_align / nw_align calls - which in turn are the compiled reference's (checked here for a sample of the queries).
    The batch mixes query lengths (different kernel shapes, buffer growth mid-batch), queries whose hit list has ties
    across the cut-off (full replay path inside the pipeline) and a hit count larger than the database."""
    from libssa_b200 import synth
    case = {"symtype": cases.AMINOACID, "scoring": "blosum62", "gaps": [-11, -1], "db": ["protein", 3000, 31, "P07327"], "query": "P07327", "runs": []}
    q0 = cases.configure(prod, case)
    prod.free_sequence(q0)
    rng = np.random.default_rng(5)
    texts = [cases.query_text(n) for n in ("P07327", "O74807", "P27895", "P56980")]
    texts += [synth.random_protein(rng, int(n)) for n in (35, 64, 200, 333, 700, 1300, 90, 410)]
    queries = [prod.init_sequence_fasta(cases.READ_FROM_STRING, t) for t in texts]
    try:
        for nw, bits, k, align in ((False, 8, 10, 0), (True, 16, 10, 1), (False, 16, 60, 0), (True, 8, 25, 0), (False, 64, 5000, 0), (False, 16, 7, 1)):
            single = [[_key(h, align) for h in (prod.nw_align if nw else prod.sw_align)(q, k, bits, align)] for q in queries]
            batch = [[_key(h, align) for h in hits] for hits in prod.align_batch(queries, k, bits, align, nw=nw)]
            assert batch == single, (nw, bits, k, align)
            again = [[_key(h, align) for h in (prod.nw_align if nw else prod.sw_align)(q, k, bits, align)] for q in queries[:3]]
            assert again == single[:3]          # the pipeline leaves no state behind
        # sample against the reference
        cases.configure(ref, case)
        for t, q in list(zip(texts, queries))[:3]:
            rq = ref.init_sequence_fasta(cases.READ_FROM_STRING, t)
            want = [_key(h, 1) for h in ref.nw_align(rq, 10, 16, 1)]
            ref.free_sequence(rq)
            assert [_key(h, 1) for h in prod.align_batch([q, q], 10, 16, 1, nw=True)[1]] == want
    finally:
        for q in queries:
            prod.free_sequence(q)


def test_batch_two_queries_per_word(prod, tmp_path, monkeypatch):
    """ssa_b200_align_batch scores consecutive queries of one kind two per packed word (score_kernel<..., DUAL>, every shape:
    groups of 4 / 8 / 16 / 32 lanes on a length-matched database): lists and alignments equal single calls, for pairs of
    equal, similar and very different lengths, an odd query count, SW and NW; and the same batch with the pairing disabled."""
    from libssa_b200 import synth
    rng = np.random.default_rng(99)
    seqs = synth.protein_db(3001, 1234, mu=float(np.log(300)), sigma=0.04, min_len=250, max_len=360, n_dup=20)
    lens = [300, 310, 64, 350, 17, 5, 128, 128, 1000, 900, 33, 700, 250]
    texts = [synth.random_protein(rng, n) for n in lens]
    seqs[5] = texts[1]
    seqs[6] = texts[0][:150] + texts[3][100:300]
    path = tmp_path / "lm.fas"
    synth.write_fasta(str(path), seqs)
    prod.init_symbol_translation(cases.AMINOACID, 1, 1, 1)
    prod.init_score_matrix(cases.MATRIX_BUILDIN, "blosum62")
    prod.init_gap_penalties(-11, -1)
    prod.init_db(str(path))
    queries = [prod.init_sequence_fasta(cases.READ_FROM_STRING, t) for t in texts]
    try:
        for nw, bits, k, align in ((True, 16, 10, 0), (False, 16, 10, 0), (True, 16, 8, 1), (False, 8, 12, 1), (True, 8, 40, 0)):
            single = [[_key(h, align) for h in (prod.nw_align if nw else prod.sw_align)(q, k, bits, align)] for q in queries]
            batch = [[_key(h, align) for h in hits] for hits in prod.align_batch(queries, k, bits, align, nw=nw)]
            assert batch == single, (nw, bits, k, align, [i for i in range(len(lens)) if batch[i] != single[i]])
            monkeypatch.setenv("LIBSSA_B200_NO_DUAL", "1")
            plain = [[_key(h, align) for h in hits] for hits in prod.align_batch(queries, k, bits, align, nw=nw)]
            monkeypatch.delenv("LIBSSA_B200_NO_DUAL")
            assert plain == single
    finally:
        for q in queries:
            prod.free_sequence(q)


def test_batched_queries_multi_frame_fallback(prod, tmp_path):
    """batches of multi-frame searches run query by query (no pipelining) and must still equal single calls"""
    import frames as F
    recs, query = F.workload(F.TRANS_QUERY)
    db = tmp_path / "db.fas"
    with open(db, "w") as f:
        for i, r in enumerate(recs):
            f.write(">r%d\n%s\n" % (i, r))
    F.configure(prod, F.TRANS_QUERY, F.BOTH)
    prod.init_db(str(db))
    qs = [prod.init_sequence_fasta(cases.READ_FROM_STRING, query[i:] + query[:i]) for i in (0, 7, 50)]
    try:
        single = [[(h.score, h.db_id, h.q_strand, h.q_frame, h.cigar) for h in prod.sw_align(q, 12, 8, 1)] for q in qs]
        batch = [[(h.score, h.db_id, h.q_strand, h.q_frame, h.cigar) for h in hits] for hits in prod.align_batch(qs, 12, 8, 1)]
        assert batch == single
    finally:
        for q in qs:
            prod.free_sequence(q)
        prod.init_symbol_translation(cases.AMINOACID, 1, 1, 1)


def test_asymmetric_matrix_alignment_matches_reference(prod, ref, tmp_path):
    """A user matrix read from a string need not be symmetric.  The search scores with M[db][query], the region passes of the
    alignment read M[query][db] (reference src/algo/align.c:39-131), so the alignment's own forward score differs from the
    search score - the reference returns the alignment of ITS forward score, and so must the library (it used to abort)."""
    from libssa_b200 import synth
    rng = np.random.default_rng(77)
    letters = "ARNDCQEGHILKMFPSTWYV"
    vals = rng.integers(-4, 3, size=(20, 20))
    for i in range(20):
        vals[i, i] = int(rng.integers(4, 9))
    assert (vals != vals.T).any()
    text = "   " + "  ".join(letters) + "\n" + "".join(
        "%s %s\n" % (letters[i], " ".join("%2d" % v for v in vals[i])) for i in range(20))
    q = synth.random_protein(rng, 180)
    seqs = synth.protein_db(400, 78, query=q, homolog_frac=0.05, n_dup=4, mu=5.0, sigma=0.5, min_len=5, max_len=900)
    path = tmp_path / "asym.fas"
    synth.write_fasta(str(path), seqs)
    results = []
    for lib in (ref, prod):
        lib.set_output_mode(0)
        lib.set_thread_count(1)
        lib.set_chunk_size(1000)
        lib.init_symbol_translation(cases.AMINOACID, 1, 1, 1)
        lib.init_score_matrix(cases.READ_FROM_STRING, text)
        lib.init_gap_penalties(-7, -1)
        lib.init_db(str(path))
        qh = lib.init_sequence_fasta(cases.READ_FROM_STRING, q)
        out = []
        for mode, bw in (("sw", 16), ("sw", 8), ("nw", 16), ("sw", 64)):
            hits = (lib.sw_align if mode == "sw" else lib.nw_align)(qh, 8, bw, cases.COMPUTE_ALIGNMENT)
            out.append([(h.score, h.db_id, h.cigar, h.q_start, h.q_end, h.d_start, h.d_end) for h in hits])
        lib.free_sequence(qh)
        results.append(out)
    assert results[0] == results[1]
