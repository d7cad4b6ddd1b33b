"""CPU tests of the product's host side: the C-ABI library loads and exports every symbol the headers
declare, and its host logic (matrices, query parsing, arrival-order emulation, heap replay, shard plan)
agrees with the oracle.  No search is run here - that needs a GPU (tests/test_gpu_parity.py)."""
import ctypes as C
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
MATS = json.load(open(os.path.join(GOLD, "matrices.json")))


def declared_functions(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = re.findall(r"\b([a-z_0-9]+)\s*\([^;{}]*\)\s*;", txt)
    txt = re.sub(r"typedef[^;]*;", "", txt)
    names = re.findall(r"\b([a-z_0-9]+)\s*\([^;{}]*\)\s*;", txt)
    return sorted(set(n for n in names if n not in ("void", "int", "size_t")))


def test_library_exports_every_declared_symbol(prod_host):
    out = subprocess.check_output(["nm", "-D", "--defined-only", prod_host.path]).decode()
    exported = set(l.split()[-1] for l in out.splitlines() if l.strip())
    for header in ("libssa.h", "libssa_b200.h", "libssa_extern_db.h"):
        fns = declared_functions(header)
        assert len(fns) >= 4, header
        missing = [f for f in fns if f not in exported]
        assert not missing, (header, missing)
    # nothing but the API is exported (internal names of the reference like `fatal` or `xmalloc` stay hidden)
    stray = [s for s in exported if not re.match(r"^(ssa_|set_|init_|free_|sw_align|nw_align)", s)]
    assert not stray, stray


def test_sass_is_sm100a_and_uses_dpx(prod_host):
    out = subprocess.run(["cuobjdump", "-sass", prod_host.path], capture_output=True, text=True).stdout
    assert "EF_CUDA_SM100" in out
    for mnemonic in ("VIADDMNMX.S16x2", "VIMNMX3.S16x2", "VIADD.16x2", "SHFL.UP", "PRMT"):
        assert mnemonic in out, mnemonic


@pytest.mark.parametrize("name", ["blosum45", "blosum50", "blosum62", "blosum80", "blosum90", "pam30", "pam70", "pam250"])
def test_builtin_matrices_equal_reference(prod_host, name):
    prod_host.init_score_matrix(cases.MATRIX_BUILDIN, name)
    assert prod_host.score_matrix().reshape(-1).tolist() == MATS[name]


@pytest.mark.parametrize("pair", [(1, -1), (2, -2), (5, -4), (127, -1)])
def test_constant_matrices_equal_reference(prod_host, pair):
    """first argument = match score (reference src/libssa.c:115-117); row/column 0 stay -1 (tests/test_matrices.c:115-124)"""
    prod_host.init_constant_scores(*pair)
    m = prod_host.score_matrix()
    assert m.reshape(-1).tolist() == MATS["const_%d_%d" % pair]
    assert m[3, 3] == pair[0] and m[3, 4] == pair[1] and m[0, 5] == -1 and m[5, 0] == -1


def test_matrix_from_string_and_file(prod_host, tmp_path):
    text = "# comment\n   A  R  N\nA  4 -1 -2\nR -1  5  0\nN -2  0  6\n"
    prod_host.init_score_matrix(cases.READ_FROM_STRING, text)
    m = prod_host.score_matrix()
    A, R, N = 1, 16, 13
    assert (m[A, A], m[A, R], m[R, N], m[N, N], m[2, 2]) == (4, -1, 0, 6, -1)
    p = tmp_path / "m.txt"
    p.write_text(text)
    prod_host.init_score_matrix(cases.READ_FROM_FILE, str(p))
    assert (prod_host.score_matrix() == m).all()


class _Seq(C.Structure):
    _fields_ = [("seq", C.POINTER(C.c_char)), ("len", C.c_size_t)]


class _Query(C.Structure):  # reference src/libssa_datatypes.h:100-107
    _fields_ = [("nt", _Seq * 2), ("aa", _Seq * 6), ("header", C.c_char_p), ("headerlen", C.c_size_t),
                ("map", C.c_void_p), ("sym", C.c_void_p)]


def _codes(seq):
    return list(C.string_at(seq.seq, seq.len))


def test_query_mapping_matches_oracle(prod_host, port):
    prod_host.set_output_mode(0)
    prod_host.init_symbol_translation(cases.AMINOACID, 1, 1, 1)
    text = "MKV-LA?x b*UOJz\n"
    q = prod_host.init_sequence_fasta(cases.READ_FROM_STRING, text)
    got = _codes(C.cast(q, C.POINTER(_Query)).contents.aa[0])
    assert got == port.map_query(text, False).tolist()      # unknown symbols removed (query.c:102-130)
    prod_host.free_sequence(q)
    prod_host.init_symbol_translation(cases.NUCLEOTIDE, 1, 1, 1)
    q = prod_host.init_sequence_fasta(cases.READ_FROM_STRING, "ACGTUacgtu-NRYKM#")
    got = _codes(C.cast(q, C.POINTER(_Query)).contents.nt[0])
    assert got == port.map_query("ACGTUacgtu-NRYKM#", True).tolist()
    prod_host.free_sequence(q)


def test_query_from_fasta_file(prod_host, port):
    prod_host.init_symbol_translation(cases.AMINOACID, 1, 1, 1)
    q = prod_host.init_sequence_fasta(cases.READ_FROM_FILE, os.path.join(cases.DATA, "P07327"))
    s = C.cast(q, C.POINTER(_Query)).contents
    assert s.aa[0].len == 375                               # reference benchmark/results/query_length:16
    assert _codes(s.aa[0]) == port.map_query(cases.query_text("P07327"), False).tolist()
    assert s.header.startswith(b"sp|P07327")
    prod_host.free_sequence(q)
    assert prod_host.init_sequence_fasta(cases.READ_FROM_FILE, "/nonexistent/file") is None


def test_default_fasta_provider_contract(prod_host):
    """reference tests/test_libssa_extern_db.c:12-58"""
    L = prod_host.lib

    class SeqInfo(C.Structure):
        _fields_ = [("ID", C.c_size_t), ("seqlen", C.c_size_t), ("seq", C.c_char_p)]
    L.ssa_db_get_sequence.restype = C.POINTER(SeqInfo)
    L.ssa_db_get_sequence.argtypes = [C.c_size_t]
    L.ssa_db_get_sequence_count.restype = C.c_size_t
    L.ssa_db_close()
    L.ssa_db_close()                                         # tolerates "nothing open"
    prod_host.init_db(os.path.join(cases.DATA, "AF091148.fas"))
    assert L.ssa_db_get_sequence_count() == 1403
    for i in (0, 1, 700, 1402):
        assert L.ssa_db_get_sequence(i).contents.ID == i
    assert not L.ssa_db_get_sequence(1403) and not L.ssa_db_get_sequence(1404)
    prod_host.init_db(os.path.join(cases.DATA, "one_seq.fas"))
    assert L.ssa_db_get_sequence(0).contents.seqlen == 54 and not L.ssa_db_get_sequence(1)


def _random_lens(rng, n):
    lens = np.clip(np.rint(rng.lognormal(4.5, 0.8, n)), 0, 3000).astype(np.uint64)
    lens[rng.random(n) < 0.03] = 0                          # skipped records
    return lens


@pytest.mark.parametrize("bit_width", [8, 16, 64])
@pytest.mark.parametrize("chunk", [1000, 100, 7])
def test_arrival_order_matches_oracle(prod_host, port, bit_width, chunk):
    rng = np.random.default_rng(bit_width * 1000 + chunk)
    for n in (1, 5, 31, 33, 257, 2500):
        lens = _random_lens(rng, n)
        ovf8 = np.where(rng.random(n) < 0.1, rng.integers(0, 40, n), -1).astype(np.int64)
        ovf16 = np.where((ovf8 >= 0) & (rng.random(n) < 0.3), ovf8 + rng.integers(0, 5, n), -1).astype(np.int64)
        for o8, o16 in ((None, None), (ovf8, ovf16)):
            a = prod_host.host_arrival_order(lens, chunk, bit_width, o8, o16)
            b = port.arrival_order(lens, chunk, bit_width, o8, o16)
            assert a.tolist() == b.tolist()


def test_arrival_order_stops_at_empty_chunk(prod_host):
    """a chunk without a non-empty record ends the reference's search loop (src/algo/16/search_16.c:123)"""
    lens = np.array([5, 6, 0, 0, 7, 8], dtype=np.uint64)
    assert prod_host.host_arrival_order(lens, 2, 64).tolist() == [0, 1]
    assert prod_host.host_arrival_order(lens, 3, 64).tolist() == [0, 1, 4, 5]


def test_replay_matches_oracle_with_ties(prod_host, port):
    rng = np.random.default_rng(5)
    for n, k in ((1, 1), (10, 20), (500, 10), (5000, 50), (5000, 1)):
        scores = rng.integers(-5, 12, n).astype(np.int64)       # many ties at the cut-off
        order = rng.permutation(n).astype(np.uint64)
        got = prod_host.host_replay(scores[order], order, k)
        assert got == port.topk_replay(scores, order, k)
    assert prod_host.host_replay(np.arange(20), np.arange(20), 5) == [(19, 19), (18, 18), (17, 17), (16, 16), (15, 15)]
    assert prod_host.host_replay(np.zeros(0), np.zeros(0), 5) == []


def test_pass_plan_covers_every_query_row(prod_host):
    """packed tier: the passes tile the query rows without gaps; long queries = full 32 x 32 passes + one remainder pass"""
    for qlen in list(range(1, 140)) + [255, 256, 257, 375, 511, 512, 513, 1000, 1023, 1024, 1025, 1030, 1100, 1500, 2005, 2047, 2048,
                                        2049, 3100, 5119, 5120, 5121, 5147, 10000, 35213]:
        passes = prod_host.host_plan_passes(qlen)
        assert passes, qlen
        row = 0
        for i, (G, R, base) in enumerate(passes):
            assert G in (4, 8, 16, 32) and R in (4, 8, 12, 16, 20, 24, 28, 32), (qlen, passes)
            assert base == row, (qlen, passes)
            if i + 1 < len(passes):
                assert G == 32, (qlen, passes)          # only the last pass may run on narrow groups
            row += G * R
        assert row >= qlen, (qlen, passes)
        assert row - passes[-1][0] * passes[-1][1] < qlen, (qlen, passes)     # no pass beyond the query
        if qlen > 1024:
            assert row - qlen < 128, (qlen, passes)   # at most a few padded rows (a uniform plan pads up to a whole lane block)
    # the 5147-row benchmark query: 5 x (32 x 32) + a small remainder pass
    p = prod_host.host_plan_passes(5147)
    assert [x[:2] for x in p[:5]] == [(32, 32)] * 5 and len(p) == 6 and p[5][0] * p[5][1] >= 27
    # a very long pair in the packed class keeps the remainder pass off the narrowest groups
    wide = prod_host.host_plan_passes(5147, total_cols=2_000_000, longest_cols=400_000)
    assert wide[-1][0] > p[-1][0]


def test_work_batches_partition_the_stream(prod_host):
    """work queue: whole pairs, contiguous, `cols` columns per batch, shrinking towards the end but never below min_cols"""
    rng = np.random.default_rng(5)
    lens = np.sort(rng.integers(30, 3000, size=20000))[::-1]
    cols_of = (lens + 1) & ~1
    pair_col = np.concatenate([[0], np.cumsum(cols_of)]).astype(np.uint64)
    total = int(pair_col[-1])
    for first_pair, cols, groups, min_cols in ((0, 4096, 0, 512), (0, 4096, 300, 512), (17, 2048, 1000, 128), (0, 256, 5000, 128)):
        first = prod_host.host_partition_batches(pair_col, first_pair, cols, groups, min_cols)
        assert first[0] == first_pair and (np.diff(first.astype(np.int64)) > 0).all()
        bounds = np.concatenate([first, [len(lens)]])
        size = (pair_col[bounds[1:]] - pair_col[bounds[:-1]]).astype(np.int64)
        assert size.sum() == total - int(pair_col[first_pair])
        left = total - pair_col[bounds[:-1]].astype(np.int64)
        want = np.full(len(size), cols) if groups == 0 else np.minimum(cols, np.maximum(min_cols, left // (2 * groups)))
        assert (size[:-1] >= want[:-1]).all()                                        # at least the wanted columns ...
        assert (size[:-1] - cols_of[bounds[1:-1] - 1] < want[:-1]).all()             # ... and not a pair more than needed
        if groups:
            uniform = prod_host.host_partition_batches(pair_col, first_pair, cols, 0, min_cols)
            assert len(first) >= len(uniform)
            if total // cols > 4 * groups:
                assert size[0] >= cols and size[-2] < cols                           # big batches first, small ones last


@pytest.mark.parametrize("seed", range(6))
def test_sparse_replay_equals_full_replay(prod_host, seed):
    """large databases: only the first m arrivals and the later arrivals above the heap's root leave the device"""
    rng = np.random.default_rng(seed)
    n = int(rng.integers(2000, 60000))
    hi = int(rng.choice([3, 40, 100000]))                    # few distinct scores: ties everywhere, also at the cut-off
    scores = rng.integers(0, hi, size=n).astype(np.int32)
    if seed % 2:
        scores = np.sort(scores)                              # worst case: every later record beats the root
    order = rng.permutation(n).astype(np.uint32)
    for k in (1, 7, 100):
        full = prod_host.host_replay(scores, order, k)
        for prefix in (k, k + 1, 500, 5000, n, n + 10):
            assert prod_host.host_replay_sparse(scores, order, k, prefix) == full, (seed, k, prefix)


def test_sharded_merge_two_ranks_gloo(tmp_path):
    """World-size-2 run of the shard/all-gather/replay logic on CPU (gloo): each rank scores its shard
    with the oracle standing in for the GPU kernels, the shard scores are all-gathered exactly as
    libssa_b200 does it (rank-major blocks of `stride` scores), and both ranks must reproduce the
    one-process hit list."""
    script = os.path.join(ROOT, "tests", "dist_merge_worker.py")
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.path.join(ROOT, "tests"))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29613", script, str(tmp_path)],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert out.stdout.count("MERGE_OK") == 2, out.stdout


@pytest.mark.parametrize("gencode", [1, 2, 4, 11, 16, 22, 23])
def test_query_translation_matches_reference(prod_host, ref, gencode):
    """6-frame translation of a nucleotide query (reference src/util/util_sequence.c:200-382, src/query.c:146-159)"""
    from libssa_b200 import synth
    text = synth.random_nucleotide(gencode, 311, iupac_rate=0.15)
    for strands in (1, 2, 3):
        frames = []
        for lib in (prod_host, ref):
            lib.set_output_mode(0)
            lib.init_symbol_translation(cases.TRANS_QUERY if hasattr(cases, "TRANS_QUERY") else 2, strands, gencode, gencode)
            q = lib.init_sequence_fasta(cases.READ_FROM_STRING, text)
            s = C.cast(q, C.POINTER(_Query)).contents
            frames.append([_codes(s.aa[i]) if s.aa[i].seq else None for i in range(6)])
            lib.free_sequence(q)
        assert frames[0] == frames[1], (gencode, strands)
        assert sum(f is not None for f in frames[0]) == (6 if strands == 3 else 3)
    for lib in (prod_host, ref):
        lib.init_symbol_translation(cases.AMINOACID, 1, 1, 1)


@pytest.mark.parametrize("mode", __import__("frames").MODES, ids=__import__("frames").MODE_IDS)
def test_frames_replay_matches_reference(prod_host, ref, port, tmp_path, mode):
    """Strand / reading-frame searches: the host-side replay of the reference's schedule (chunks x frames x
    overflow lists, heaps, both sorts), fed with oracle scores of the reference's own entry sequences, must give
    the hit list the reference's search phase produces (src/algo/8/search_8.c:94-124, src/algo/manager.c:141-176),
    and the committed fixture must be what the reference produces."""
    import frames as F
    symtype, strands = mode
    out = tmp_path / "ref.json"
    F.run_reference_worker(symtype, strands, out, entries=True)   # own process: see tests/frames_ref_worker.py
    R = json.load(open(out))
    gold = json.load(open(os.path.join(GOLD, "frames.json")))[F.MODE_IDS[F.MODES.index(mode)]]
    assert gold["searches"] == R["searches"]
    chunk, k = R["chunk"], R["k"]
    recs, query = F.workload(symtype)
    F.configure(prod_host, symtype, strands)
    M = prod_host.score_matrix()
    spec, go, ge = F.scoring(symtype)
    qf = [(np.array(c, dtype=np.uint8), s, f) for c, s, f in R["query_frames"]]
    assert [(c.tolist(), s, f) for c, s, f in F.query_frames(prod_host, symtype, strands, query)] == \
           [(c.tolist(), s, f) for c, s, f in qf]
    entries = [np.array(e, dtype=np.uint8) for e in R["entries"]]
    bm = F.buffer_max(symtype, strands)
    assert len(entries) == bm * len(recs)
    lens = np.array([len(e) for e in entries], dtype=np.uint64)
    Q, n = len(qf), len(entries)
    for nw in (False, True):
        score = np.zeros((Q, n), dtype=np.int32)
        fn = port.nw_score if nw else port.sw_score
        for q in range(Q):
            for s in range(n):
                if lens[s]:
                    score[q, s] = fn(entries[s], qf[q][0], M, go, ge)
        for bits in (64, 16, 8):
            o8 = np.full((Q, n), -1, dtype=np.int32)
            o16 = np.full((Q, n), -1, dtype=np.int32)
            if bits != 64:
                for q in range(Q):
                    for s in range(n):
                        if not lens[s]:
                            continue
                        if nw:
                            if bits == 8:
                                o8[q, s] = port.nw_simd_channel(entries[s], qf[q][0], M, go, ge, 8)[1]
                            o16[q, s] = port.nw_simd_channel(entries[s], qf[q][0], M, go, ge, 16)[1]
                        elif score[q, s] >= 255:
                            _, b8, b16 = port.sw_score_blocks(entries[s], qf[q][0], M, go, ge)
                            o8[q, s], o16[q, s] = (b8 if bits == 8 else -1), b16
            got = prod_host.host_frames_replay(lens, chunk * bm, bits, bm, score, k, o8, o16)
            got = [[sc, slot // bm] + list(reversed(F.entry_fields(symtype, strands, slot % bm))) + [q] for sc, slot, q in got]
            want = R["searches"]["%s%d" % ("nw" if nw else "sw", bits)]
            assert got == want["phase"], (mode, nw, bits)
            if not nw and bits == 8:
                assert want["overflow_8"] > 0, "workload must exercise the overflow lists"
    prod_host.init_symbol_translation(cases.AMINOACID, 1, 1, 1)


def test_reference_translated_query_vectors_oracle(prod_host, port):
    """The reference's known answers for a six-frame query (tests/algo/test_searcher.c:247-325) hold for the oracle's
    scores of the frames the product's query reader translates (genetic code 3)."""
    import frames as F
    prod_host.set_output_mode(0)
    prod_host.init_constant_scores(1, -1)
    M = prod_host.score_matrix()
    qf = F.query_frames(prod_host, F.TRANS_QUERY, F.BOTH, "ATGCCCAAAATTGACTTGAATAGGGGCGTGAGGAGGGGTTTTCAATACTATTTTTTG", 3)
    assert [len(c) for c, _, _ in qf] == [19, 18, 18, 19, 18, 18]
    rec = port.map_db("".join(l.strip() for l in open(os.path.join(cases.DATA, "short_AA.fas")) if not l.startswith(">")), False)
    assert sorted((int(port.nw_score(rec, c, M, -1, -1)) for c, _, _ in qf), reverse=True) == [-109, -109, -110, -110, -110, -113]
    assert sorted((int(port.sw_score(rec, c, M, -1, -1)) for c, _, _ in qf), reverse=True) == [2, 2, 2, 2, 2, 1]
    prod_host.init_symbol_translation(cases.AMINOACID, 1, 1, 1)


def test_nw_virtual_rows_carry_the_boundary(port):
    """The packed NW kernels shift the query down by `shift` virtual rows so that its last row is the last row of a lane
    (search.cu:fill_profile_packed, score_kernel.cuh:nw_start).  Scalar model of exactly that array - virtual rows score gapE against
    every residue, start at gapO (the last one at 0), real rows start at gapO + (i + 1) gapE, E starts at H + gapO + gapE, the top
    boundary gapO + (j + 1) gapE with F = H + gapO + gapE is fed in above the first row, the first diagonal is gapO (0 without
    virtual rows) - against the oracle's exact NW score, for every shift and several gap settings (including gapO = 0)."""
    rng = np.random.default_rng(11)
    M = rng.integers(-4, 6, size=(32, 32)).astype(np.int64)
    M = (M + M.T) // 2

    def nw_start(pos, shift, gO, gE):
        if pos < shift - 1:
            return gO
        if pos == shift - 1:
            return 0
        return gO + (pos - shift + 1) * gE

    def shifted_score(q, d, gO, gE, shift):
        qr = gO + gE
        rows = shift + len(q)
        H = [nw_start(p, shift, gO, gE) for p in range(rows)]       # column -1
        E = [h + qr for h in H]                                      # E entering column 0
        diag_top = nw_start(-1, shift, gO, gE)                       # H of the row above the array, column -1
        for j in range(len(d)):
            h_up = gO + (j + 1) * gE                                 # top boundary, fed in above the first row
            f = h_up + qr
            hd = diag_top
            diag_top = h_up
            for p in range(rows):
                v = gE if p < shift else int(M[d[j], q[p - shift]])
                t = hd + v
                hd = H[p]
                h = max(t, E[p], f)
                hq = h + qr
                E[p] = max(E[p] + gE, hq)
                f = max(f + gE, hq)
                H[p] = h
        return H[rows - 1]

    for gO, gE in ((-11, -1), (-3, -2), (0, -2), (-5, -5)):
        for _ in range(6):
            q = rng.integers(1, 24, size=int(rng.integers(1, 40))).astype(np.uint8)
            d = rng.integers(1, 24, size=int(rng.integers(1, 60))).astype(np.uint8)
            exact = int(port.score_db([d], q, M, gO, gE, True)[0])
            for shift in (0, 1, 2, 5, 11):
                assert shifted_score(q, d, gO, gE, shift) == exact, (gO, gE, len(q), len(d), shift)
