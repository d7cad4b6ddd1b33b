/*
 * libssa_b200.h - additive, GPU-specific entry points of libssa_b200.so.
 *
 * Nothing here is needed by a program written against libssa.h; these calls
 * expose what the reference has no notion of: device selection, sharding of
 * the database over several GPUs (one process per GPU), the collective used to
 * merge per-GPU results, and counters/timers of the last search (the
 * reference only prints its overflow counters at OUTPUT_INFO,
 * reference src/algo/manager.c:147-160).
 */
#ifndef LIBSSA_B200_H_
#define LIBSSA_B200_H_

#include <stddef.h>
#include <stdint.h>

#include "libssa.h"

#ifdef __cplusplus
extern "C" {
#endif

/* counters and timers of the last sw_align/nw_align call (and of the last database build) */
typedef struct {
    uint64_t n_seqs;          /* non-empty database records (all shards) */
    uint64_t n_residues;      /* their total length */
    uint64_t local_seqs;      /* records held by this process (its shard; all its GPUs' shards in single-process mode) */
    uint64_t local_residues;
    uint64_t cells;           /* query length x local_residues: real DP cells of the last search */
    uint64_t overflow_8;      /* records the reference would have re-aligned at 16 bit (manager.c:157-160) */
    uint64_t overflow_16;     /* ... and at 64 bit */
    uint64_t wide_tier_seqs;  /* records this library scored with its exact 32-bit kernel */
    uint64_t launches;        /* CUDA kernels launched by the last search */
    uint64_t h2d_bytes;       /* host->device bytes of the last search (query profile etc.) */
    uint64_t d2h_bytes;       /* device->host bytes of the last search (scores, hits) */
    double score_kernel_ms;   /* device time of the scoring kernels, CUDA events on the library stream */
    double device_ms;         /* device time of everything the search enqueued */
    double call_ms;           /* host wall-clock of the whole *_align call */
    double pack_ms;           /* last database build: host packing */
    double upload_ms;         /* last database build: host->device upload */
    uint64_t db_device_bytes; /* resident packed database size on this device */
} ssa_b200_stats_t;

/* CUDA device this process uses (default: $LOCAL_RANK if set, else 0).  Call before the first search. */
void ssa_b200_set_device( int device );

/*
 * ONE process, SEVERAL GPUs - the drop-in for a plain C host (the counterpart of the reference's thread fan-out inside
 * the library, reference src/algo/manager.c:120-146): the database is sharded by residue count over devices 0 .. n-1, one
 * host thread per device enqueues its shard's kernels, the shard scores are merged with ncclAllGather over NVLink/NVSwitch
 * (ncclCommInitAll inside the library, libnccl.so.2 loaded on demand), hit selection and traceback run on device 0.
 * Without this call the environment variable LIBSSA_B200_GPUS (a number or "all") decides; default 1 - so an unchanged
 * libssa.h program uses 8 GPUs with `LIBSSA_B200_GPUS=8 ./program`.  Call before the first search.
 */
void ssa_b200_set_device_count( int n );
int ssa_b200_get_device_count( void );

/*
 * ONE process PER GPU (torchrun / MPI): shard the database over `world_size` processes.  The records are sorted by length
 * and dealt round-robin, so every shard holds the same residue count to within one record.  Each process
 * scores its shard only; the shard scores are merged by a collective, either
 *   (a) inside the library: rank 0 calls ssa_b200_nccl_unique_id, the host program carries the 128 bytes to the other
 *       ranks over any channel it has, every rank calls ssa_b200_nccl_init( id, rank, world_size ) (which implies
 *       ssa_b200_set_shard) - the library then issues ncclAllGather on its own CUDA stream; or
 *   (b) supplied by the host: ssa_b200_set_shard + ssa_b200_set_allgather( fn ) with fn all-gathering `bytes` bytes from
 *       every rank's DEVICE buffer `d_send` into `d_recv` (world_size * bytes, rank order), ordered on CUDA stream `stream`.
 */
void ssa_b200_set_shard( int rank, int world_size );
void ssa_b200_nccl_unique_id( void * out_128_bytes );
void ssa_b200_nccl_init( const void * id_128_bytes, int rank, int world_size );
typedef void (*ssa_b200_allgather_fn)( void * ctx, const void * d_send, void * d_recv, size_t bytes, void * stream );
void ssa_b200_set_allgather( ssa_b200_allgather_fn fn, void * ctx );

/* builds (packs + uploads) the device-resident database now instead of lazily in the first search */
void ssa_b200_prepare_db( void );

void ssa_b200_get_stats( ssa_b200_stats_t * out );

/*
 * Exact score of every local database record against the query, without hit selection:
 * writes `local_seqs` entries in provider order (the records this process holds); ids[] receives the
 * provider IDs.  Either pointer may be NULL.  nw = 0 Smith-Waterman, 1 Needleman-Wunsch.
 * Returns local_seqs.
 */
size_t ssa_b200_score_all( p_query p, int nw, int64_t * scores, uint64_t * ids );

/*
 * Host-only pieces of the hit selection (no GPU involved): the one-thread heap-arrival order of the
 * reference scheduler, the bounded-heap replay, and the global length-sorted order that defines the
 * shards.  The multi-process merge is built from these; exposed so it can be tested on CPU.
 * lens[n] are record lengths in provider order (0 = skipped record); ovf8/ovf16 may be NULL.
 */
size_t ssa_b200_host_arrival_order( const uint64_t * lens, size_t n, size_t chunk, int bit_width, const int64_t * ovf8,
        const int64_t * ovf16, uint32_t * order_out );
size_t ssa_b200_host_replay( const int32_t * scores_in_arrival_order, const uint32_t * order, size_t n, size_t hitcount,
        uint64_t * ids_out, int64_t * scores_out );
size_t ssa_b200_host_sorted_order( const uint64_t * lens, size_t n, uint32_t * sorted_out );
/* the replay from the first `prefix` >= hitcount arrivals plus the later arrivals that beat the heap's root after them - what
 * the library does on large databases so that only those records leave the device; equals ssa_b200_host_replay */
size_t ssa_b200_host_replay_sparse( const int32_t * scores_in_arrival_order, const uint32_t * order, size_t n, size_t hitcount,
        size_t prefix, uint64_t * ids_out, int64_t * scores_out );
/*
 * Host-only pieces of the scoring schedule (no GPU involved), exposed so they can be tested on CPU.
 * ssa_b200_host_plan_passes: the passes of the packed tier for a query of qlen rows - pass i covers query rows
 * [rowbase[i], rowbase[i] + G[i] * R[i]) on groups of G[i] lanes x R[i] rows; queries beyond 1024 rows run as full
 * passes of 32 x 32 rows plus a remainder pass (total_cols / longest_cols: stream columns of the packed class and of
 * its longest pair; sms: multiprocessors).  Replaces the reference's "whole query per SIMD block" loop
 * (src/algo/simd/search_simd_sw.c:238-277).  Returns the number of passes (at most max_passes are written).
 * ssa_b200_host_partition_batches: first pair of every work batch of the persistent scoring kernels - the counterpart of
 * the reference's chunk cursor (src/db_adapter.c:212-239): whole pairs, `cols` stream columns per batch, shrinking
 * towards the end of the queue when groups > 0.  pair_col[npairs + 1] = first stream column of each pair.
 */
size_t ssa_b200_host_plan_passes( size_t qlen, uint64_t total_cols, uint64_t longest_cols, int sms, int * G_out, int * R_out,
        size_t * rowbase_out, size_t max_passes );
size_t ssa_b200_host_partition_batches( const uint64_t * pair_col, size_t npairs, size_t first_pair, uint32_t cols, uint32_t groups,
        uint32_t min_cols, uint32_t * first_pair_out, size_t max_batches );

/*
 * Hit selection of a search over several query frames and/or several database entries per record (both strands,
 * translated searches): the reference's per-chunk schedule  frames x entries -> overflow lists re-run under every
 * frame  (src/algo/8/search_8.c:94-124, src/algo/16/search_16.c:92-113, src/algo/64/search_64.c:44-56), the bounded
 * heaps and both final sorts.  slot_lens[nslots]: entry lengths in provider order, buffer_max entries per record
 * (0 = skipped); scores / ovf8 / ovf16 are [q_count][nslots] (ovf: overflow block of the entry under that frame in
 * the 8- resp. 16-bit tier, -1 = none; may be NULL).  Returns the number of hits written.
 */
size_t ssa_b200_host_frames_replay( const uint64_t * slot_lens, size_t nslots, size_t chunk_slots, int bit_width, int q_count,
        int buffer_max, const int32_t * scores, const int32_t * ovf8, const int32_t * ovf16, size_t hitcount,
        uint32_t * slots_out, uint8_t * frames_out, int64_t * scores_out );

/*
 * Batched searches (the reference API has one call per query, src/libssa.h:229-251): out[i] receives
 * exactly what sw_align / nw_align( queries[i], hitcount, bit_width, align_type ) returns (free each with
 * free_alignment).  The library overlaps the host's share of query i with the GPU's work on query i+1, scores
 * consecutive one-pass queries two per packed word, and aligns the hits of many queries together.
 * nw: 0 = Smith-Waterman, 1 = Needleman-Wunsch.
 */
void ssa_b200_align_batch( int nw, const p_query * queries, size_t nqueries, size_t hitcount, int bit_width, int align_type,
        p_alignment_list * out );

/* copies the current 32x32 score table (index (db_symbol << 5) + query_symbol); 0 if scoring is not initialised */
int ssa_b200_get_score_matrix( int64_t * out_1024 );

/* library/ABI version: major*10000 + minor*100 + patch */
int ssa_b200_version( void );

#ifdef __cplusplus
}
#endif

#endif /* LIBSSA_B200_H_ */
