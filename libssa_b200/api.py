"""ctypes mirror of include/libssa.h and include/libssa_b200.h (same names, same argument meaning).

Reference interface being mirrored: reference src/libssa.h:122-263.  State is process-global, exactly
as in the reference; one ``LibSSA`` object per process.
"""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))

# constants, reference src/libssa.h:28-72
NUCLEOTIDE, AMINOACID, TRANS_QUERY, TRANS_DB, TRANS_BOTH = 0, 1, 2, 3, 4
FORWARD_STRAND, COMPLEMENTARY_STRAND, BOTH_STRANDS = 1, 2, 3
OUTPUT_SILENT, OUTPUT_ERROR, OUTPUT_WARNING, OUTPUT_INFO = 0, 1, 2, 3
COMPUTE_SCORE, COMPUTE_ALIGNMENT = 0, 1
READ_FROM_FILE, READ_FROM_STRING, MATRIX_BUILDIN = 0, 1, 2
BIT_WIDTH_8, BIT_WIDTH_16, BIT_WIDTH_64 = 8, 16, 64


class LibraryMissing(RuntimeError):
    pass


def library_path():
    # LIBSSA_B200_LIB: development aid for A/B runs of two builds on one box
    return os.environ.get("LIBSSA_B200_LIB") or os.path.join(HERE, "libssa_b200.so")


def build_library(jobs=8):
    """Compile libssa_b200.so in-tree with nvcc for sm_100a (see csrc/Makefile)."""
    subprocess.check_call(["make", "-s", "-j%d" % jobs, "-C", os.path.join(HERE, "csrc")])
    return library_path()


class _DbSeq(C.Structure):
    _fields_ = [("seq", C.POINTER(C.c_char)), ("len", C.c_size_t), ("ID", C.c_size_t),
                ("strand", C.c_int), ("frame", C.c_int)]


class _QSeq(C.Structure):
    _fields_ = [("seq", C.POINTER(C.c_char)), ("len", C.c_size_t), ("strand", C.c_int), ("frame", C.c_int)]


class _Alignment(C.Structure):
    _fields_ = [("db_seq", _DbSeq), ("query", _QSeq), ("alignment", C.c_char_p),
                ("alignment_len", C.c_size_t), ("score", C.c_long),
                ("align_q_start", C.c_size_t), ("align_q_end", C.c_size_t),
                ("align_d_start", C.c_size_t), ("align_d_end", C.c_size_t)]


class _AlignmentList(C.Structure):
    _fields_ = [("alignments", C.POINTER(C.POINTER(_Alignment))), ("len", C.c_size_t)]


class Stats(C.Structure):
    """ssa_b200_stats_t (include/libssa_b200.h)."""
    _fields_ = [("n_seqs", C.c_uint64), ("n_residues", C.c_uint64), ("local_seqs", C.c_uint64),
                ("local_residues", C.c_uint64), ("cells", C.c_uint64), ("overflow_8", C.c_uint64),
                ("overflow_16", C.c_uint64), ("wide_tier_seqs", C.c_uint64), ("launches", C.c_uint64),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("score_kernel_ms", C.c_double),
                ("device_ms", C.c_double), ("call_ms", C.c_double), ("pack_ms", C.c_double),
                ("upload_ms", C.c_double), ("db_device_bytes", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Hit:
    """Python copy of one alignment_t of the result list."""
    __slots__ = ("score", "db_id", "db_len", "db_seq", "q_len", "cigar", "cigar_len",
                 "q_start", "q_end", "d_start", "d_end", "db_strand", "db_frame", "q_strand", "q_frame")

    def __repr__(self):
        return "Hit(score=%d, id=%d, cigar=%r, q=%d-%d, d=%d-%d)" % (
            self.score, self.db_id, self.cigar, self.q_start, self.q_end, self.d_start, self.d_end)


class LibSSA:
    def __init__(self, path=None):
        path = path or library_path()
        if not os.path.exists(path):
            raise LibraryMissing(
                "%s not found: build it with `make -C libssa_b200/csrc` (or __graft_entry__.build()). "
                "libssa_b200 has no CPU fallback." % path)
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
            f.restype = C.POINTER(_AlignmentList)
        lib.free_alignment.argtypes = [C.POINTER(_AlignmentList)]
        lib.ssa_exit.argtypes = []
        lib.ssa_b200_set_device.argtypes = [C.c_int]
        lib.ssa_b200_set_shard.argtypes = [C.c_int, C.c_int]
        lib.ssa_b200_prepare_db.argtypes = []
        lib.ssa_b200_get_stats.argtypes = [C.POINTER(Stats)]
        lib.ssa_b200_score_all.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_uint64)]
        lib.ssa_b200_score_all.restype = C.c_size_t
        lib.ssa_b200_version.restype = C.c_int
        for name in ("set_output_mode", "set_simd_compute_mode", "set_chunk_size", "set_thread_count",
                     "init_score_matrix", "init_constant_scores", "init_gap_penalties", "init_symbol_translation",
                     "init_db", "free_sequence", "free_alignment", "ssa_exit", "ssa_b200_set_device",
                     "ssa_b200_set_shard", "ssa_b200_prepare_db", "ssa_b200_get_stats"):
            getattr(lib, name).restype = None

    # ---- libssa.h ----
    def set_output_mode(self, mode): self.lib.set_output_mode(mode)
    def set_simd_compute_mode(self, mode): self.lib.set_simd_compute_mode(mode)
    def set_chunk_size(self, n): self.lib.set_chunk_size(n)
    def set_thread_count(self, n): self.lib.set_thread_count(n)
    def init_score_matrix(self, mode, matrix): self.lib.init_score_matrix(mode, matrix.encode())
    def init_constant_scores(self, match, mismatch): self.lib.init_constant_scores(match, mismatch)
    def init_gap_penalties(self, gap_open, gap_extend): self.lib.init_gap_penalties(gap_open, gap_extend)

    def init_symbol_translation(self, symtype, strands=FORWARD_STRAND, db_gencode=1, q_gencode=1):
        self.lib.init_symbol_translation(symtype, strands, db_gencode, q_gencode)

    def init_db(self, path): self.lib.init_db(os.fspath(path).encode())

    def init_sequence_fasta(self, mode, text_or_path):
        return self.lib.init_sequence_fasta(mode, os.fspath(text_or_path).encode())

    def free_sequence(self, query): self.lib.free_sequence(query)
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
            h.cigar_len = a.alignment_len
            h.q_start, h.q_end = a.align_q_start, a.align_q_end
            h.d_start, h.d_end = a.align_d_start, a.align_d_end
            h.db_strand, h.db_frame = a.db_seq.strand, a.db_seq.frame
            h.q_strand, h.q_frame = a.query.strand, a.query.frame
            hits.append(h)
        self.lib.free_alignment(plist)
        return hits

    def sw_align(self, query, hitcount, bit_width=BIT_WIDTH_16, align_type=COMPUTE_SCORE):
        return self._collect(self.lib.sw_align(query, hitcount, bit_width, align_type))

    def nw_align(self, query, hitcount, bit_width=BIT_WIDTH_16, align_type=COMPUTE_SCORE):
        return self._collect(self.lib.nw_align(query, hitcount, bit_width, align_type))

    def align_batch(self, queries, hitcount, bit_width=BIT_WIDTH_16, align_type=COMPUTE_SCORE, nw=False):
        """One hit list per query: the lists sw_align / nw_align would return, computed as one pipelined batch."""
        n = len(queries)
        qs = (C.c_void_p * n)(*queries)
        out = (C.POINTER(_AlignmentList) * n)()
        f = self.lib.ssa_b200_align_batch
        f.restype = None
        f.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_void_p]
        f(1 if nw else 0, qs, n, hitcount, bit_width, align_type, out)
        return [self._collect(out[i]) for i in range(n)]

    # ---- libssa_b200.h ----
    def set_device(self, device): self.lib.ssa_b200_set_device(device)
    def set_shard(self, rank, world_size): self.lib.ssa_b200_set_shard(rank, world_size)

    def set_device_count(self, n):
        """One process, n GPUs: shards over devices 0..n-1, merged with NCCL inside the library."""
        self.lib.ssa_b200_set_device_count.argtypes = [C.c_int]
        self.lib.ssa_b200_set_device_count.restype = None
        self.lib.ssa_b200_set_device_count(n)

    def device_count(self):
        self.lib.ssa_b200_get_device_count.restype = C.c_int
        return self.lib.ssa_b200_get_device_count()

    def nccl_unique_id(self):
        """128 bytes identifying a new NCCL communicator (rank 0 creates it, the host carries it to the other ranks)."""
        buf = C.create_string_buffer(128)
        self.lib.ssa_b200_nccl_unique_id.argtypes = [C.c_void_p]
        self.lib.ssa_b200_nccl_unique_id.restype = None
        self.lib.ssa_b200_nccl_unique_id(buf)
        return buf.raw

    def nccl_init(self, unique_id, rank, world_size):
        """Joins the communicator: from now on the library all-gathers shard scores itself (ncclAllGather on its stream)."""
        assert len(unique_id) == 128
        self.lib.ssa_b200_nccl_init.argtypes = [C.c_char_p, C.c_int, C.c_int]
        self.lib.ssa_b200_nccl_init.restype = None
        self.lib.ssa_b200_nccl_init(unique_id, rank, world_size)
    def prepare_db(self): self.lib.ssa_b200_prepare_db()

    def stats(self):
        s = Stats()
        self.lib.ssa_b200_get_stats(C.byref(s))
        return s

    def score_all(self, query, nw=False):
        """(ids, scores) numpy arrays of every local database record, provider order."""
        import numpy as np
        self.prepare_db()
        n = self.stats().local_seqs
        scores = np.zeros(max(n, 1), dtype=np.int64)
        ids = np.zeros(max(n, 1), dtype=np.uint64)
        got = self.lib.ssa_b200_score_all(query, 1 if nw else 0,
                                          scores.ctypes.data_as(C.POINTER(C.c_int64)),
                                          ids.ctypes.data_as(C.POINTER(C.c_uint64)))
        return ids[:got], scores[:got]

    def version(self): return self.lib.ssa_b200_version()

    def score_matrix(self):
        """The current 32x32 score table as a numpy int64 array ([db_symbol, query_symbol]), or None."""
        import numpy as np
        m = np.zeros(1024, dtype=np.int64)
        self.lib.ssa_b200_get_score_matrix.argtypes = [C.POINTER(C.c_int64)]
        self.lib.ssa_b200_get_score_matrix.restype = C.c_int
        ok = self.lib.ssa_b200_get_score_matrix(m.ctypes.data_as(C.POINTER(C.c_int64)))
        return m.reshape(32, 32) if ok else None

    # ---- host-only pieces of the hit selection (no GPU involved) ----
    def host_arrival_order(self, lens, chunk, bit_width, ovf8=None, ovf16=None):
        import numpy as np
        lens = np.ascontiguousarray(lens, dtype=np.uint64)
        out = np.zeros(max(len(lens), 1), dtype=np.uint32)
        f = self.lib.ssa_b200_host_arrival_order
        f.restype = C.c_size_t
        f.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        o8 = None if ovf8 is None else np.ascontiguousarray(ovf8, dtype=np.int64)
        o16 = None if ovf16 is None else np.ascontiguousarray(ovf16, dtype=np.int64)
        n = f(lens.ctypes.data, len(lens), chunk, bit_width, None if o8 is None else o8.ctypes.data,
              None if o16 is None else o16.ctypes.data, out.ctypes.data)
        return out[:n]

    def host_replay(self, scores_in_arrival_order, order, hitcount):
        import numpy as np
        sc = np.ascontiguousarray(scores_in_arrival_order, dtype=np.int32)
        order = np.ascontiguousarray(order, dtype=np.uint32)
        ids = np.zeros(max(hitcount, 1), dtype=np.uint64)
        out = np.zeros(max(hitcount, 1), dtype=np.int64)
        f = self.lib.ssa_b200_host_replay
        f.restype = C.c_size_t
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p]
        n = f(sc.ctypes.data, order.ctypes.data, len(order), hitcount, ids.ctypes.data, out.ctypes.data)
        return [(int(out[i]), int(ids[i])) for i in range(n)]

    def host_replay_sparse(self, scores_in_arrival_order, order, hitcount, prefix):
        import numpy as np
        sc = np.ascontiguousarray(scores_in_arrival_order, dtype=np.int32)
        order = np.ascontiguousarray(order, dtype=np.uint32)
        ids = np.zeros(max(hitcount, 1), dtype=np.uint64)
        out = np.zeros(max(hitcount, 1), dtype=np.int64)
        f = self.lib.ssa_b200_host_replay_sparse
        f.restype = C.c_size_t
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p]
        n = f(sc.ctypes.data, order.ctypes.data, len(order), hitcount, prefix, ids.ctypes.data, out.ctypes.data)
        return [(int(out[i]), int(ids[i])) for i in range(n)]

    def host_frames_replay(self, slot_lens, chunk_slots, bit_width, buffer_max, scores, hitcount, ovf8=None, ovf16=None):
        """Hit list [(score, slot, frame)] of a multi-frame search; scores/ovf8/ovf16 are [q_count][nslots]."""
        import numpy as np
        lens = np.ascontiguousarray(slot_lens, dtype=np.uint64)
        sc = np.ascontiguousarray(scores, dtype=np.int32)
        q_count, nslots = sc.shape
        o8 = None if ovf8 is None else np.ascontiguousarray(ovf8, dtype=np.int32)
        o16 = None if ovf16 is None else np.ascontiguousarray(ovf16, dtype=np.int32)
        slots = np.zeros(max(hitcount, 1), dtype=np.uint32)
        frames = np.zeros(max(hitcount, 1), dtype=np.uint8)
        out = np.zeros(max(hitcount, 1), dtype=np.int64)
        f = self.lib.ssa_b200_host_frames_replay
        f.restype = C.c_size_t
        f.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                      C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        n = f(lens.ctypes.data, nslots, chunk_slots, bit_width, q_count, buffer_max, sc.ctypes.data,
              None if o8 is None else o8.ctypes.data, None if o16 is None else o16.ctypes.data, hitcount,
              slots.ctypes.data, frames.ctypes.data, out.ctypes.data)
        return [(int(out[i]), int(slots[i]), int(frames[i])) for i in range(n)]

    def host_sorted_order(self, lens):
        import numpy as np
        lens = np.ascontiguousarray(lens, dtype=np.uint64)
        out = np.zeros(max(len(lens), 1), dtype=np.uint32)
        f = self.lib.ssa_b200_host_sorted_order
        f.restype = C.c_size_t
        f.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        n = f(lens.ctypes.data, len(lens), out.ctypes.data)
        return out[:n]

    def host_plan_passes(self, qlen, total_cols=80_000_000, longest_cols=5000, sms=148):
        """[(G, R, rowbase)] of the packed tier's passes for a query of qlen rows (host-only)."""
        import numpy as np
        G = np.zeros(4096, dtype=np.int32); R = np.zeros(4096, dtype=np.int32); base = np.zeros(4096, dtype=np.uint64)
        f = self.lib.ssa_b200_host_plan_passes
        f.restype = C.c_size_t
        f.argtypes = [C.c_size_t, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        n = f(qlen, total_cols, longest_cols, sms, G.ctypes.data, R.ctypes.data, base.ctypes.data, 4096)
        return [(int(G[i]), int(R[i]), int(base[i])) for i in range(min(n, 4096))]

    def host_partition_batches(self, pair_col, first_pair, cols, groups, min_cols):
        """first pair of every work batch over pair_col[npairs + 1] (host-only)."""
        import numpy as np
        pc = np.ascontiguousarray(pair_col, dtype=np.uint64)
        out = np.zeros(max(len(pc), 1), dtype=np.uint32)
        f = self.lib.ssa_b200_host_partition_batches
        f.restype = C.c_size_t
        f.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_size_t]
        n = f(pc.ctypes.data, len(pc) - 1, first_pair, cols, groups, min_cols, out.ctypes.data, len(out))
        return out[:n]

    def set_allgather(self, fn):
        """Register the collective that merges shard results: fn(send_ptr, recv_ptr, nbytes, cuda_stream_ptr)."""
        proto = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)
        self._allgather_cb = proto(lambda ctx, s, r, n, st: fn(s, r, n, st)) if fn else proto()
        self.lib.ssa_b200_set_allgather.argtypes = [proto, C.c_void_p]
        self.lib.ssa_b200_set_allgather.restype = None
        self.lib.ssa_b200_set_allgather(self._allgather_cb, None)


_loaded = None


def load(path=None):
    """The process-wide library object (the C library keeps global state, like the reference)."""
    global _loaded
    if _loaded is None:
        _loaded = LibSSA(path)
    return _loaded
