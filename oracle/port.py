"""TEST INFRASTRUCTURE ONLY - python face of oracle/restate.c (the C restatement).

Never imported by the product.  Sequences are numpy uint8 code arrays; the
score matrix is the reference's 32x32 int64 table (db_symbol*32 + query_symbol).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "liboracle.so")

# reference src/util/util_sequence.c:41-88 (symbol order of the two alphabets)
AA_SYMBOLS = "-ABCDEFGHIKLMNPQRSTVWXYZU*OJ"
NT_SYMBOLS = "-ACMGRSVTWYHKDBN"


def _build_map(symbols, extra=()):
    m = np.full(256, -1, dtype=np.int16)
    for code, ch in enumerate(symbols):
        m[ord(ch.upper())] = code
        m[ord(ch.lower())] = code
    for ch, code in extra:
        m[ord(ch.upper())] = code
        m[ord(ch.lower())] = code
    return m


MAP_AA = _build_map(AA_SYMBOLS)
MAP_AA[ord("-")] = -1          # '-' is not mapped for proteins (util_sequence.c:47)
MAP_NT = _build_map(NT_SYMBOLS, extra=(("U", 8),))   # U -> T (util_sequence.c:77)


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "port"])


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            build()
        L = C.CDLL(SO)
        u8p, i64p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_int64), C.POINTER(C.c_uint64)
        for f in (L.orc_sw_score, L.orc_nw_score):
            f.argtypes = [u8p, C.c_size_t, u8p, C.c_size_t, i64p, C.c_int, C.c_int]
            f.restype = C.c_int64
        L.orc_sw_score_blocks.argtypes = [u8p, C.c_size_t, u8p, C.c_size_t, i64p, C.c_int, C.c_int,
                                          C.c_int64, C.c_int64, i64p, i64p]
        L.orc_sw_score_blocks.restype = C.c_int64
        L.orc_nw_simd_channel.argtypes = [u8p, C.c_size_t, u8p, C.c_size_t, i64p, C.c_int, C.c_int, C.c_int, i64p]
        L.orc_nw_simd_channel.restype = C.c_int64
        L.orc_score_db.argtypes = [C.c_int, u8p, u64p, C.c_size_t, u8p, C.c_size_t, i64p, C.c_int, C.c_int, i64p]
        L.orc_score_db.restype = None
        L.orc_topk_replay.argtypes = [i64p, u64p, u64p, C.c_size_t, C.c_size_t, u64p, i64p]
        L.orc_topk_replay.restype = C.c_size_t
        L.orc_arrival_order.argtypes = [u64p, i64p, i64p, C.c_size_t, C.c_size_t, C.c_int, u64p]
        L.orc_arrival_order.restype = C.c_size_t
        L.orc_region_local.argtypes = [u8p, C.c_size_t, u8p, C.c_size_t, i64p, C.c_int, C.c_int, i64p]
        L.orc_region_local.restype = C.c_int
        L.orc_cigar.argtypes = [C.c_int, u8p, C.c_size_t, u8p, C.c_size_t, i64p, C.c_int, C.c_int, i64p]
        L.orc_cigar.restype = C.c_void_p
        L.orc_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def map_db(seq, nucleotide):
    """DB-side mapping: unknown symbols become code 0 (util_sequence.c:298-318)."""
    raw = np.frombuffer(seq.encode() if isinstance(seq, str) else seq, dtype=np.uint8)
    m = (MAP_NT if nucleotide else MAP_AA)[raw]
    return np.where(m < 0, 0, m).astype(np.uint8)


def map_query(seq, nucleotide):
    """Query-side mapping: unknown symbols are removed (query.c:102-130)."""
    raw = np.frombuffer(seq.encode() if isinstance(seq, str) else seq, dtype=np.uint8)
    m = (MAP_NT if nucleotide else MAP_AA)[raw]
    return m[m >= 0].astype(np.uint8)


def read_fasta(path):
    """[(header, ascii_sequence)] with whitespace stripped, record index = ID."""
    recs, cur, hdr = [], None, None
    with open(path, "rb") as f:
        for line in f:
            if line.startswith(b">"):
                if cur is not None:
                    recs.append((hdr, b"".join(cur)))
                hdr, cur = line[1:].strip().decode(errors="replace"), []
            elif cur is not None:
                cur.append(b"".join(line.split()))
    if cur is not None:
        recs.append((hdr, b"".join(cur)))
    return recs


def constant_matrix(match, mismatch):
    """reference src/matrices.c:447-460: codes 1..31, row/column 0 stay -1."""
    M = np.full((32, 32), -1, dtype=np.int64)
    M[1:, 1:] = mismatch
    for a in range(1, 32):
        M[a, a] = match
    return M


def sw_score(d, q, M, gapO, gapE):
    M = np.ascontiguousarray(M, dtype=np.int64)
    return lib().orc_sw_score(_p(d, C.c_uint8), len(d), _p(q, C.c_uint8), len(q), _p(M, C.c_int64), gapO, gapE)


def nw_score(d, q, M, gapO, gapE):
    M = np.ascontiguousarray(M, dtype=np.int64)
    return lib().orc_nw_score(_p(d, C.c_uint8), len(d), _p(q, C.c_uint8), len(q), _p(M, C.c_int64), gapO, gapE)


def sw_score_blocks(d, q, M, gapO, gapE, thr8=255, thr16=65535):
    M = np.ascontiguousarray(M, dtype=np.int64)
    b8, b16 = C.c_int64(), C.c_int64()
    s = lib().orc_sw_score_blocks(_p(d, C.c_uint8), len(d), _p(q, C.c_uint8), len(q), _p(M, C.c_int64), gapO, gapE,
                                  thr8, thr16, C.byref(b8), C.byref(b16))
    return s, b8.value, b16.value


def nw_simd_channel(d, q, M, gapO, gapE, bits):
    """(narrow-tier score, overflow block or -1) of the reference's 8/16-bit NW kernel for one sequence."""
    M = np.ascontiguousarray(M, dtype=np.int64)
    b = C.c_int64()
    s = lib().orc_nw_simd_channel(_p(d, C.c_uint8), len(d), _p(q, C.c_uint8), len(q), _p(M, C.c_int64), gapO, gapE,
                                  bits, C.byref(b))
    return s, b.value


def concat(seqs):
    offs = np.zeros(len(seqs) + 1, dtype=np.uint64)
    if seqs:
        offs[1:] = np.cumsum([len(s) for s in seqs])
    db = np.concatenate(seqs).astype(np.uint8) if seqs and offs[-1] > 0 else np.zeros(1, dtype=np.uint8)
    return np.ascontiguousarray(db), offs


def score_db(seqs, q, M, gapO, gapE, nw):
    """Exact scores of every database sequence (list of uint8 code arrays)."""
    db, offs = concat(seqs)
    out = np.zeros(len(seqs), dtype=np.int64)
    M = np.ascontiguousarray(M, dtype=np.int64)
    q = np.ascontiguousarray(q, dtype=np.uint8)
    lib().orc_score_db(1 if nw else 0, _p(db, C.c_uint8), _p(offs, C.c_uint64), len(seqs), _p(q, C.c_uint8), len(q),
                       _p(M, C.c_int64), gapO, gapE, _p(out, C.c_int64))
    return out


def arrival_order(lens, chunk, bit_width, ovf8=None, ovf16=None):
    lens = np.ascontiguousarray(lens, dtype=np.uint64)
    n = len(lens)
    order = np.zeros(max(n, 1), dtype=np.uint64)
    o8 = None if ovf8 is None else np.ascontiguousarray(ovf8, dtype=np.int64)
    o16 = None if ovf16 is None else np.ascontiguousarray(ovf16, dtype=np.int64)
    cnt = lib().orc_arrival_order(_p(lens, C.c_uint64), None if o8 is None else _p(o8, C.c_int64),
                                  None if o16 is None else _p(o16, C.c_int64), n, chunk, bit_width,
                                  _p(order, C.c_uint64))
    return order[:cnt]


def topk_replay(scores, order, k, ids=None):
    scores = np.ascontiguousarray(scores, dtype=np.int64)
    order = np.ascontiguousarray(order, dtype=np.uint64)
    out_ids = np.zeros(max(k, 1), dtype=np.uint64)
    out_sc = np.zeros(max(k, 1), dtype=np.int64)
    idp = None
    if ids is not None:
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        idp = _p(ids, C.c_uint64)
    n = lib().orc_topk_replay(_p(scores, C.c_int64), idp, _p(order, C.c_uint64), len(order), k,
                              _p(out_ids, C.c_uint64), _p(out_sc, C.c_int64))
    return [(int(out_sc[i]), int(out_ids[i])) for i in range(n)]


def search(seqs, q, M, gapO, gapE, nw, k, bit_width=64, chunk=1000):
    """Top-k (score, id) list of a one-thread reference search, ties resolved as the reference does."""
    scores = score_db(seqs, q, M, gapO, gapE, nw)
    lens = np.array([len(s) for s in seqs], dtype=np.uint64)
    ovf8 = ovf16 = None
    if not nw and bit_width in (8, 16):
        ovf8 = np.full(len(seqs), -1, dtype=np.int64)
        ovf16 = np.full(len(seqs), -1, dtype=np.int64)
        for i in np.nonzero(scores >= 255)[0]:
            _, b8, b16 = sw_score_blocks(seqs[i], q, M, gapO, gapE)
            ovf8[i], ovf16[i] = b8, b16
    if nw and bit_width in (8, 16):
        ovf8 = np.full(len(seqs), -1, dtype=np.int64)
        ovf16 = np.full(len(seqs), -1, dtype=np.int64)
        for i in range(len(seqs)):
            if len(seqs[i]) == 0:
                continue
            # a record a saturating tier ACCEPTS is reported with that tier's own score: the injected top boundary saturates
            # silently, so it can differ from the exact one (search_simd_nw.c:423-426, 445-453; e.g. -118 for an exact -139)
            if bit_width == 8:
                s8, ovf8[i] = nw_simd_channel(seqs[i], q, M, gapO, gapE, 8)
                if ovf8[i] < 0:
                    scores[i] = s8
            if bit_width == 16 or ovf8[i] >= 0:
                s16, ovf16[i] = nw_simd_channel(seqs[i], q, M, gapO, gapE, 16)
                if ovf16[i] < 0:
                    scores[i] = s16
    order = arrival_order(lens, chunk, bit_width, ovf8, ovf16)
    return topk_replay(scores, order, k)


def region_local(a, b, M, gapO, gapE):
    M = np.ascontiguousarray(M, dtype=np.int64)
    reg = np.zeros(4, dtype=np.int64)
    rc = lib().orc_region_local(_p(a, C.c_uint8), len(a), _p(b, C.c_uint8), len(b), _p(M, C.c_int64), gapO, gapE,
                                _p(reg, C.c_int64))
    if rc != 0:
        raise RuntimeError("Internal error in align function.")
    return tuple(int(x) for x in reg)


def cigar(a, b, M, gapO, gapE, nw, region=None):
    """(region, cigar string) for query a vs database sequence b."""
    M = np.ascontiguousarray(M, dtype=np.int64)
    if region is None:
        region = (0, len(a) - 1, 0, len(b) - 1) if nw else region_local(a, b, M, gapO, gapE)
    reg = np.array(region, dtype=np.int64)
    p = lib().orc_cigar(1 if nw else 0, _p(a, C.c_uint8), len(a), _p(b, C.c_uint8), len(b), _p(M, C.c_int64), gapO,
                        gapE, _p(reg, C.c_int64))
    s = C.string_at(p).decode()
    lib().orc_free(p)
    return tuple(region), s
