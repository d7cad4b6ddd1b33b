"""TEST INFRASTRUCTURE ONLY - ctypes binding of the libssa.h API.

Works for any shared library that exports the libssa.h symbols
(reference: src/libssa.h:122-263): the compiled reference
(oracle/_ref/libssa_ref.so) and the product library alike, so the parity
tests drive both through exactly the same calls.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libssa_ref.so")

# constants, reference src/libssa.h:28-72
NUCLEOTIDE, AMINOACID, TRANS_QUERY, TRANS_DB, TRANS_BOTH = 0, 1, 2, 3, 4
FORWARD_STRAND, COMPLEMENTARY_STRAND, BOTH_STRANDS = 1, 2, 3
OUTPUT_SILENT, OUTPUT_ERROR, OUTPUT_WARNING, OUTPUT_INFO = 0, 1, 2, 3
COMPUTE_SCORE, COMPUTE_ALIGNMENT = 0, 1
READ_FROM_FILE, READ_FROM_STRING, MATRIX_BUILDIN = 0, 1, 2
COMPUTE_ON_SSE2, COMPUTE_ON_SSE41, COMPUTE_ON_AVX2 = 0, 1, 2


class DbSeq(C.Structure):  # db_seq_t, src/libssa.h:85-91
    _fields_ = [("seq", C.POINTER(C.c_char)), ("len", C.c_size_t), ("ID", C.c_size_t),
                ("strand", C.c_int), ("frame", C.c_int)]


class QSeq(C.Structure):  # q_seq_t, src/libssa.h:93-98
    _fields_ = [("seq", C.POINTER(C.c_char)), ("len", C.c_size_t), ("strand", C.c_int), ("frame", C.c_int)]


class Alignment(C.Structure):  # alignment_t, src/libssa.h:100-110
    _fields_ = [("db_seq", DbSeq), ("query", QSeq), ("alignment", C.c_char_p),
                ("alignment_len", C.c_size_t), ("score", C.c_long),
                ("align_q_start", C.c_size_t), ("align_q_end", C.c_size_t),
                ("align_d_start", C.c_size_t), ("align_d_end", C.c_size_t)]


class AlignmentList(C.Structure):  # alignment_list_t, src/libssa.h:113-116
    _fields_ = [("alignments", C.POINTER(C.POINTER(Alignment))), ("len", C.c_size_t)]


class Hit:
    """Plain-python copy of one alignment_t."""
    __slots__ = ("score", "db_id", "db_len", "db_seq", "q_len", "cigar", "cigar_len",
                 "q_start", "q_end", "d_start", "d_end", "db_strand", "db_frame", "q_strand", "q_frame")

    def key(self):
        return (self.score, self.db_id, self.cigar, self.q_start, self.q_end, self.d_start, self.d_end)

    def __repr__(self):
        return "Hit(score=%d, id=%d, cigar=%r, q=%d-%d, d=%d-%d)" % (
            self.score, self.db_id, self.cigar, self.q_start, self.q_end, self.d_start, self.d_end)


class LibSSA:
    """One loaded implementation of the libssa.h API (process-global state, as in the reference)."""

    def __init__(self, path):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.path = path
        self.lib = lib = C.CDLL(path, mode=os.RTLD_LOCAL | os.RTLD_NOW)
        lib.set_output_mode.argtypes = [C.c_int]
        lib.set_simd_compute_mode.argtypes = [C.c_int]
        lib.set_chunk_size.argtypes = [C.c_size_t]
        lib.set_thread_count.argtypes = [C.c_size_t]
        lib.init_score_matrix.argtypes = [C.c_int, C.c_char_p]
        lib.init_constant_scores.argtypes = [C.c_int8, C.c_int8]
        lib.init_gap_penalties.argtypes = [C.c_int8, C.c_int8]
        lib.init_symbol_translation.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
        lib.init_db.argtypes = [C.c_char_p]
        lib.init_sequence_fasta.argtypes = [C.c_int, C.c_char_p]
        lib.init_sequence_fasta.restype = C.c_void_p
        lib.free_sequence.argtypes = [C.c_void_p]
        for f in (lib.sw_align, lib.nw_align):
            f.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int]
            f.restype = C.POINTER(AlignmentList)
        lib.free_alignment.argtypes = [C.POINTER(AlignmentList)]
        lib.ssa_exit.argtypes = []
        for name in ("set_output_mode", "set_simd_compute_mode", "set_chunk_size", "set_thread_count",
                     "init_score_matrix", "init_constant_scores", "init_gap_penalties",
                     "init_symbol_translation", "init_db", "free_sequence", "free_alignment", "ssa_exit"):
            getattr(lib, name).restype = None

    # thin pass-throughs -------------------------------------------------
    def set_output_mode(self, m): self.lib.set_output_mode(m)
    def set_simd_compute_mode(self, m): self.lib.set_simd_compute_mode(m)
    def set_chunk_size(self, n): self.lib.set_chunk_size(n)
    def set_thread_count(self, n): self.lib.set_thread_count(n)
    def init_score_matrix(self, mode, s): self.lib.init_score_matrix(mode, s.encode())
    def init_constant_scores(self, match, mismatch): self.lib.init_constant_scores(match, mismatch)
    def init_gap_penalties(self, o, e): self.lib.init_gap_penalties(o, e)
    def init_symbol_translation(self, t, strands=FORWARD_STRAND, dg=1, qg=1):
        self.lib.init_symbol_translation(t, strands, dg, qg)
    def init_db(self, path): self.lib.init_db(os.fspath(path).encode())
    def init_sequence_fasta(self, mode, s):
        return self.lib.init_sequence_fasta(mode, os.fspath(s).encode())
    def free_sequence(self, q): self.lib.free_sequence(q)
    def ssa_exit(self): self.lib.ssa_exit()

    def _collect(self, plist):
        hits = []
        if not plist:
            return hits
        al = plist.contents
        for i in range(al.len):
            a = al.alignments[i].contents
            h = Hit()
            h.score = a.score
            h.db_id = a.db_seq.ID
            h.db_len = a.db_seq.len
            h.db_seq = C.string_at(a.db_seq.seq, a.db_seq.len) if a.db_seq.seq else None
            h.q_len = a.query.len
            h.cigar = a.alignment.decode() if a.alignment else None
            h.cigar_len = a.alignment_len if a.alignment else 0
            h.q_start, h.q_end = a.align_q_start, a.align_q_end
            h.d_start, h.d_end = a.align_d_start, a.align_d_end
            h.db_strand, h.db_frame = a.db_seq.strand, a.db_seq.frame
            h.q_strand, h.q_frame = a.query.strand, a.query.frame
            hits.append(h)
        self.lib.free_alignment(plist)
        return hits

    def sw_align(self, q, hitcount, bit_width, align_type):
        return self._collect(self.lib.sw_align(q, hitcount, bit_width, align_type))

    def nw_align(self, q, hitcount, bit_width, align_type):
        return self._collect(self.lib.nw_align(q, hitcount, bit_width, align_type))


class _Elem(C.Structure):  # elem_t, src/util/minheap.h:26-32
    _fields_ = [("db_id", C.c_size_t), ("db_frame", C.c_uint8), ("db_strand", C.c_uint8), ("query_id", C.c_uint8),
                ("score", C.c_long)]


class _MinHeap(C.Structure):  # minheap_t, src/util/minheap.h:34-38
    _fields_ = [("alloc", C.c_size_t), ("count", C.c_size_t), ("array", C.POINTER(_Elem))]


class _SearchResult(C.Structure):  # search_result_t, src/libssa_datatypes.h:40-46
    _fields_ = [("heap", C.POINTER(_MinHeap)), ("chunk_count", C.c_size_t), ("seq_count", C.c_size_t),
                ("overflow_8_bit_count", C.c_size_t), ("overflow_16_bit_count", C.c_size_t)]


def search_phase(ref, q, hitcount, bit_width, nw):
    """The reference's search phase alone, through its internal entry points: one searcher pass (s_search,
    src/algo/searcher.c:158-181), the merge into the result heap and minheap_sort exactly as m_run does
    (src/algo/manager.c:141-176), WITHOUT the result-record creation - which crashes for TRANS_DB / TRANS_BOTH
    (us_prepare_sequence reallocs an uninitialised pointer, src/util/util_sequence.c:385,408).
    -> ([(score, db_id, db_frame, db_strand, query_id)], overflow_8, overflow_16)"""
    L = ref.lib
    (L.init_for_nw if nw else L.init_for_sw).argtypes = [C.c_void_p, C.c_int, C.c_int]
    (L.init_for_nw if nw else L.init_for_sw)(q, bit_width, 0)
    L.s_search.restype = C.POINTER(_SearchResult)
    L.s_search.argtypes = [C.POINTER(C.c_size_t)]
    L.minheap_init.restype = C.POINTER(_MinHeap)
    L.minheap_init.argtypes = [C.c_size_t]
    L.minheap_add.argtypes = [C.POINTER(_MinHeap), C.POINTER(_Elem)]
    L.minheap_sort.argtypes = [C.POINTER(_MinHeap)]
    L.minheap_exit.argtypes = [C.POINTER(_MinHeap)]
    L.s_free.argtypes = [C.POINTER(_SearchResult)]
    k = C.c_size_t(hitcount)
    res = L.s_search(C.byref(k))
    r = res.contents
    merged = L.minheap_init(hitcount)
    for j in range(r.heap.contents.count):
        L.minheap_add(merged, C.byref(r.heap.contents.array[j]))
    L.minheap_sort(merged)
    out = []
    for j in range(merged.contents.count):
        e = merged.contents.array[j]
        out.append((int(e.score), int(e.db_id), int(e.db_frame), int(e.db_strand), int(e.query_id)))
    o8, o16 = int(r.overflow_8_bit_count), int(r.overflow_16_bit_count)
    L.adp_exit()
    L.minheap_exit(merged)
    L.s_free(res)
    return out, o8, o16


_ref = None


def reference():
    """The compiled, unmodified reference (oracle/_ref/libssa_ref.so); None if it was not built."""
    global _ref
    if _ref is None and os.path.exists(REF_SO):
        _ref = LibSSA(REF_SO)
        _ref.set_output_mode(OUTPUT_SILENT)
    return _ref
