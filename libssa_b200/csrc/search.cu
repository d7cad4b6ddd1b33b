// Device side of a search: resident packed database, query profile, tier selection, kernel launches,
// arrival-order scatter and the hand-off to the host replay.  Replaces reference
// src/db_adapter.c (per-search chunking and re-mapping), src/algo/searcher.c + src/algo/{8,16,64}/search_*.c
// (tier drivers) and the search half of src/algo/manager.c:120-172.
#include "common.h"
#include "score_launch.h"
#include "traceback.h"
#include "arrival.cuh"
#include "nw_sat.cuh"
#include "topk.cuh"

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <array>
#include <map>
#include <memory>
#include <mutex>
#include <numeric>
#include <thread>

#include <dlfcn.h>
#include <nccl.h>

using namespace ssa;

#define CUDA_OK( call )                                                                                     \
    do {                                                                                                    \
        cudaError_t e_ = ( call );                                                                          \
        if( e_ != cudaSuccess ) fatal( "CUDA error %s at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString( e_ ) ); \
    } while( 0 )

namespace {


template<typename T> struct DevBuf {  // grow-only device buffer
    T * p = nullptr;
    size_t cap = 0;
    void reserve( size_t n ) {
        if( n <= cap ) return;
        if( p ) {
            cudaDeviceSynchronize();   // a batch may still have work in flight that uses the old buffer
            cudaFree( p );
        }
        size_t want = n + n / 4 + 64;
        CUDA_OK( cudaMalloc( &p, want * sizeof( T ) ) );
        cap = want;
    }
    void release() {
        if( p ) cudaFree( p );
        p = nullptr;
        cap = 0;
    }
};

template<typename T> struct PinBuf {  // grow-only pinned host buffer
    T * p = nullptr;
    size_t cap = 0;
    void reserve( size_t n ) {
        if( n <= cap ) return;
        if( p ) {
            cudaDeviceSynchronize();
            cudaFreeHost( p );
        }
        size_t want = n + n / 4 + 64;
        CUDA_OK( cudaMallocHost( &p, want * sizeof( T ) ) );
        cap = want;
    }
    void release() {
        if( p ) cudaFreeHost( p );
        p = nullptr;
        cap = 0;
    }
};

struct PackedDb {
    bool valid = false;
    uint64_t epoch = 0;
    int symtype = -1, strands = -1, db_gencode = -1;
    int rank = 0, world = 1;
    // Database entries in provider order (all shards): lengths, 0 = skipped.  One entry per record, or buffer_max
    // entries per record (index record * buffer_max + b) when strands / translation frames multiply the database
    // (reference src/db_adapter.c:46-107,118-139).
    int buffer_max = 1;
    std::vector<uint64_t> lens_all;
    uint64_t n_seqs = 0, n_residues = 0;
    // this shard, sorted by length (descending, provider index ascending)
    std::vector<uint32_t> sorted_all;  // provider index of GLOBAL sorted position g (all shards); shard r holds g = s*world + r
    std::vector<uint32_t> sorted_id;   // provider index of local sorted position s
    std::vector<int32_t> sorted_pos_of_id;  // provider index -> GLOBAL sorted position (-1 = skipped record)
    std::vector<uint32_t> sorted_len;
    uint64_t local_residues = 0;
    size_t npairs = 0;
    uint64_t total_cols = 0;
    std::vector<uint64_t> pair_col;    // [npairs+1] first column of pair p in the stream
    struct BatchTable {                // packed-tier work batches (whole pairs) of at least `cols` columns
        std::vector<Batch> batches;
        std::vector<uint32_t> pair0;   // first pair of each batch
        DevBuf<Batch> d_batches;
    };
    std::map<std::array<uint32_t, 4>, BatchTable> batch_tables;  // keyed by (columns per batch, first pair, resident groups, smallest batch)
    DevBuf<uint16_t> d_stream;
    DevBuf<uint32_t> d_lens_all;           // record lengths, provider order (for the arrival-order kernel)
    DevBuf<uint64_t> d_pair_col;           // first stream column of every pair
    DevBuf<uint32_t> d_sorted_all;         // global sorted position -> provider index
    DevBuf<uint32_t> d_sorted_len;         // record length by local sorted position
    DevBuf<int32_t> d_sorted_pos_of_id;    // provider index -> global sorted position
    struct ChunkInfo {                     // per chunk size: where each chunk's records start in arrival order
        DevBuf<uint32_t> d_chunk_base;
        uint32_t nchunks = 0;
        size_t narr = 0;
    };
    std::map<size_t, ChunkInfo> chunk_info;
    double pack_ms = 0, upload_ms = 0;
    // arrival order caches, keyed by (bit_width, chunk): provider indices in heap-arrival order
    struct Arrival {
        std::vector<uint32_t> order;       // arrival position -> provider index (host copy)
        DevBuf<int32_t> d_pos_of_sorted;   // GLOBAL sorted position -> arrival position (-1 = never processed)
        bool uploaded = false;
    };
    std::map<std::pair<int, size_t>, Arrival> arrival;
};

// One context per GPU this process drives.  Multi-process mode (torchrun): one context, rank / world from ssa_b200_set_shard or
// ssa_b200_nccl_init.  Single-process mode (LIBSSA_B200_GPUS / ssa_b200_set_device_count): one context per device, rank = index.
struct Ctx {
    int device = -1;
    int rank = 0, world = 1;
    ncclComm_t comm = nullptr;          // in-library collective (NCCL over NVLink / NVSwitch), nullptr = none
    cudaStream_t stream = nullptr, stream_aux = nullptr;
    cudaEvent_t ev_begin = nullptr, ev_k0 = nullptr, ev_k1 = nullptr, ev_end = nullptr, ev_fork = nullptr, ev_join = nullptr;
    PackedDb db;
    DevBuf<uint32_t> d_profile, d_profile_long, d_out_packed, d_counter, d_bnd, d_bnd_wide;
    DevBuf<Batch> d_long_batches;
    // pinned upload staging exists twice (index ss): a batch of queries is enqueued one query ahead (gpu_search_batch)
    int ss = 0;
    PinBuf<Batch> h_long_batches[2];
    PinBuf<uint32_t> h_profile_long[2];
    DevBuf<int32_t> d_first_compact, d_scores, d_gathered, d_arrival_scores, d_first, d_pos_tmp;
    DevBuf<int32_t> d_all_slot[4];      // batch: the gathered scores of the (up to four) queries in flight
    DevBuf<int32_t> d_scores2;          // two queries per word: scores of the second query
    DevBuf<uint32_t> d_profile_dual;
    PinBuf<uint32_t> h_profile_dual[2];
    std::vector<uint32_t> order_tmp;
    DevBuf<Batch> d_wide_batches[2];
    DevBuf<uint32_t> d_profile_wide[2];
    PinBuf<uint32_t> h_profile_wide[2][2];
    DevBuf<uint32_t> d_order, d_wave_list, d_topk_ids;
    DevBuf<int32_t> d_topk_scores;
    DevBuf<SelectState> d_select;
    PinBuf<SelectState> h_select[4];      // top-k results, per result slot (query index & 3)
    PinBuf<uint32_t> h_topk_ids[4], h_counts[4];
    PinBuf<int32_t> h_topk_scores[4];
    cudaEvent_t ev_topk[4] = { nullptr, nullptr, nullptr, nullptr };
    DevBuf<int16_t> d_nw_scratch;
    DevBuf<uint8_t> d_query;
    DevBuf<int32_t> d_matrix32;
    uint64_t matrix32_version = ~0ull;
    PinBuf<uint32_t> h_order;
    DevBuf<uint32_t> d_flag_list;
    PinBuf<uint32_t> h_profile[2], h_list;
    PinBuf<int32_t> h_scores, h_first, h_pos, h_cand_score;
    PinBuf<uint32_t> h_cand_pos, h_cand_id;
    DevBuf<uint32_t> d_cand_pos, d_cand_id;
    DevBuf<int32_t> d_cand_score;
    PinBuf<Batch> h_wide_batches[2][2];
    DevBuf<uint32_t> d_pass_state;        // sw32_kernel, several passes per launch: [SW32_MAX_PASSES] queue heads | [nbatches] passes done
    DevBuf<int32_t> d_nw_tier;            // NW narrow tiers: overflow columns and tier scores of the records static bounds cannot clear
    DevBuf<uint32_t> d_wide_progress[2];  // pipelined 32-bit tier: [batch][pass] progress counters
    DevBuf<uint32_t> d_spec_list;         // speculative packed tier: sequences listed for the 32-bit tier
    PinBuf<uint32_t> h_spec_count[2];
    bool spec_pending[2] = { false, false };
    ssa_b200_stats_t stats;
    bool scored = false;                  // score_shard launched kernels in the current call (ev_k0 / ev_k1 are valid)
};

// ---- process-wide configuration (set before the first search creates the contexts) ----
struct Config {
    int device = -1;                      // ssa_b200_set_device: the (primary) device, -1 = $LOCAL_RANK or 0
    int rank = 0, world = 1;              // ssa_b200_set_shard / ssa_b200_nccl_init: this process' shard
    int ndev = 0;                         // ssa_b200_set_device_count: devices of this process, 0 = $LIBSSA_B200_GPUS or 1
    ssa_b200_allgather_fn allgather = nullptr;   // host-supplied collective (alternative to the in-library NCCL)
    void * allgather_ctx = nullptr;
} g_cfg;

std::vector<Ctx *> g_devs;                // contexts of this process, [0] = primary (hit selection, traceback, host results)
thread_local Ctx * t_ctx = nullptr;       // the context the calling thread works on
int g_stage_slot = 0;                     // which of the two pinned staging buffer sets the call in flight uses (gpu_search_batch)

inline Ctx & ctx() { return *( t_ctx ? t_ctx : g_devs[0] ); }
inline bool is_primary() { return &ctx() == g_devs[0]; }

// RAII: work on another context (and its device) from the calling thread
struct Activate {
    Ctx * prev;
    explicit Activate( Ctx * c ) : prev( t_ctx ) {
        t_ctx = c;
        CUDA_OK( cudaSetDevice( c->device ) );
    }
    ~Activate() {
        t_ctx = prev;
        if( prev ) cudaSetDevice( prev->device );
    }
};

// ---- NCCL, loaded on demand: a single-GPU process never needs it ----
struct NcclApi {
    void * handle = nullptr;
    ncclResult_t ( *GetUniqueId )( ncclUniqueId * ) = nullptr;
    ncclResult_t ( *CommInitRank )( ncclComm_t *, int, ncclUniqueId, int ) = nullptr;
    ncclResult_t ( *CommInitAll )( ncclComm_t *, int, const int * ) = nullptr;
    ncclResult_t ( *CommDestroy )( ncclComm_t ) = nullptr;
    ncclResult_t ( *AllGather )( const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t ) = nullptr;
    const char * ( *GetErrorString )( ncclResult_t ) = nullptr;
} g_nccl;

#define NCCL_OK( call )                                                                                        \
    do {                                                                                                       \
        ncclResult_t r_ = ( call );                                                                            \
        if( r_ != ncclSuccess ) fatal( "NCCL error %s at %s:%d: %s", #call, __FILE__, __LINE__, g_nccl.GetErrorString( r_ ) ); \
    } while( 0 )

void load_nccl() {
    if( g_nccl.handle ) return;
    const char * names[] = { getenv( "LIBSSA_B200_NCCL" ), "libnccl.so.2", "libnccl.so" };
    for( const char * n : names ) {
        if( !n || !*n ) continue;
        g_nccl.handle = dlopen( n, RTLD_NOW | RTLD_GLOBAL );
        if( g_nccl.handle ) break;
    }
    if( !g_nccl.handle ) fatal( "libssa_b200: several GPUs need NCCL, but libnccl.so.2 could not be loaded: %s", dlerror() );
    auto sym = [&]( const char * name ) {
        void * p = dlsym( g_nccl.handle, name );
        if( !p ) fatal( "libssa_b200: %s not found in the NCCL library", name );
        return p;
    };
    g_nccl.GetUniqueId = (decltype( g_nccl.GetUniqueId )) sym( "ncclGetUniqueId" );
    g_nccl.CommInitRank = (decltype( g_nccl.CommInitRank )) sym( "ncclCommInitRank" );
    g_nccl.CommInitAll = (decltype( g_nccl.CommInitAll )) sym( "ncclCommInitAll" );
    g_nccl.CommDestroy = (decltype( g_nccl.CommDestroy )) sym( "ncclCommDestroy" );
    g_nccl.AllGather = (decltype( g_nccl.AllGather )) sym( "ncclAllGather" );
    g_nccl.GetErrorString = (decltype( g_nccl.GetErrorString )) sym( "ncclGetErrorString" );
}

// ---- one host thread per additional device (the counterpart of the reference's pthread pool, src/util/thread_pool.c:50-86:
// there one thread per CPU core, here one per GPU; the calling thread drives the primary device itself) ----
struct Workers {
    std::vector<std::thread> threads;
    std::mutex m;
    std::condition_variable cv_go, cv_done;
    const std::function<void()> * job = nullptr;
    uint64_t generation = 0;
    int pending = 0;
    bool stop = false;
};
// never destroyed: a host program may exit without ssa_exit() while the device threads wait for work, and a joinable
// std::thread must not meet its destructor
Workers & g_workers = *new Workers;

void worker_main( int index ) {
    t_ctx = g_devs[index];
    CUDA_OK( cudaSetDevice( t_ctx->device ) );
    uint64_t seen = 0;
    for( ;; ) {
        const std::function<void()> * job = nullptr;
        {
            std::unique_lock<std::mutex> lk( g_workers.m );
            g_workers.cv_go.wait( lk, [&] { return g_workers.stop || g_workers.generation != seen; } );
            if( g_workers.stop ) return;
            seen = g_workers.generation;
            job = g_workers.job;
        }
        ( *job )();
        {
            std::lock_guard<std::mutex> lk( g_workers.m );
            if( --g_workers.pending == 0 ) g_workers.cv_done.notify_one();
        }
    }
}

// runs fn once per context of this process, each on its own host thread with ctx() bound to that context; returns when
// every one has ENQUEUED its work (the GPUs keep running)
void run_on_all( const std::function<void()> & fn ) {
    if( g_devs.size() == 1 ) {
        fn();
        return;
    }
    {
        std::lock_guard<std::mutex> lk( g_workers.m );
        g_workers.job = &fn;
        g_workers.pending = (int) g_devs.size() - 1;
        g_workers.generation++;
    }
    g_workers.cv_go.notify_all();
    fn();
    std::unique_lock<std::mutex> lk( g_workers.m );
    g_workers.cv_done.wait( lk, [&] { return g_workers.pending == 0; } );
}

void init_ctx( Ctx * c, int device, int rank, int world ) {
    c->device = device;
    c->rank = rank;
    c->world = world;
    CUDA_OK( cudaSetDevice( device ) );
    CUDA_OK( cudaStreamCreateWithFlags( &c->stream, cudaStreamNonBlocking ) );
    CUDA_OK( cudaStreamCreateWithFlags( &c->stream_aux, cudaStreamNonBlocking ) );
    CUDA_OK( cudaEventCreateWithFlags( &c->ev_fork, cudaEventDisableTiming ) );
    CUDA_OK( cudaEventCreateWithFlags( &c->ev_join, cudaEventDisableTiming ) );
    CUDA_OK( cudaEventCreate( &c->ev_begin ) );
    CUDA_OK( cudaEventCreate( &c->ev_k0 ) );
    CUDA_OK( cudaEventCreate( &c->ev_k1 ) );
    CUDA_OK( cudaEventCreate( &c->ev_end ) );
    for( int i = 0; i < 4; i++ ) CUDA_OK( cudaEventCreateWithFlags( &c->ev_topk[i], cudaEventDisableTiming ) );
    memset( &c->stats, 0, sizeof( c->stats ) );
}

// Creates the contexts at the first use; afterwards binds the calling thread to the primary context and its device.
void ensure_ctx() {
    if( !g_devs.empty() ) {
        if( !t_ctx ) {
            t_ctx = g_devs[0];
            CUDA_OK( cudaSetDevice( t_ctx->device ) );
        }
        return;
    }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount( &count );
    if( e != cudaSuccess || count == 0 )
        fatal( "libssa_b200: no CUDA device available (%s). This library has no CPU fallback.",
                e != cudaSuccess ? cudaGetErrorString( e ) : "device count is 0" );
    int ndev = g_cfg.ndev;
    if( ndev <= 0 ) {
        ndev = 1;
        if( const char * env = getenv( "LIBSSA_B200_GPUS" ) ) ndev = strcmp( env, "all" ) == 0 ? count : atoi( env );
        if( ndev < 1 ) ndev = 1;
    }
    if( ndev > count ) fatal( "libssa_b200: %d GPUs requested, %d visible.", ndev, count );
    if( ndev > 1 && g_cfg.world > 1 )
        fatal( "libssa_b200: several GPUs per process (LIBSSA_B200_GPUS / ssa_b200_set_device_count) cannot be combined with "
               "several processes (ssa_b200_set_shard / ssa_b200_nccl_init): use one or the other." );
    if( ndev == 1 ) {
        int dev = g_cfg.device;
        if( dev < 0 ) {
            const char * lr = getenv( "LOCAL_RANK" );
            dev = ( lr ? atoi( lr ) : 0 ) % count;
        }
        Ctx * c = new Ctx;
        init_ctx( c, dev, g_cfg.rank, g_cfg.world );
        g_devs.push_back( c );
    }
    else {
        // single process, several GPUs: the database is sharded over devices 0 .. ndev-1, one host thread per device, the
        // shard scores are all-gathered with NCCL (ncclCommInitAll) exactly as between processes
        std::vector<int> devlist( ndev );
        for( int i = 0; i < ndev; i++ ) {
            devlist[i] = i;
            Ctx * c = new Ctx;
            init_ctx( c, i, i, ndev );
            g_devs.push_back( c );
        }
        load_nccl();
        std::vector<ncclComm_t> comms( ndev );
        NCCL_OK( g_nccl.CommInitAll( comms.data(), ndev, devlist.data() ) );
        for( int i = 0; i < ndev; i++ ) g_devs[i]->comm = comms[i];
        g_workers.stop = false;
        for( int i = 1; i < ndev; i++ ) g_workers.threads.emplace_back( worker_main, i );
    }
    t_ctx = g_devs[0];
    CUDA_OK( cudaSetDevice( t_ctx->device ) );
}

uint64_t process_residues() {
    uint64_t n = 0;
    for( Ctx * c : g_devs ) n += c->db.local_residues;
    return n;
}

// ------------------------------------------------------------------------------------------------
// database packing: once per init_db / init_symbol_translation / shard change
// ------------------------------------------------------------------------------------------------

void parallel_for( size_t n, size_t grain, const std::function<void( size_t, size_t )> & fn );

// Provider-wide part of the database build, shared by the shards of all contexts of this process: the provider walk, the
// entries (strands / reading frames) as code strings, and the global length-sorted order that defines the shards.
struct GlobalDb {
    int symtype = -1, strands_key = -1, gencode_key = -1;
    int buffer_max = 1;
    bool general = false;                  // entries are materialised as code strings (several entries per record / translation)
    const signed char * map = nullptr;
    std::vector<const char *> text;        // provider-owned record texts
    std::vector<uint64_t> rec_len;
    std::vector<uint64_t> lens_all;        // entry lengths, provider order (0 = skipped)
    std::vector<uint8_t> entry_codes;
    std::vector<uint64_t> entry_off;
    std::vector<uint32_t> sorted;          // provider index by global sorted position
    std::vector<int32_t> sorted_pos_of_id;
    uint64_t n_seqs = 0, n_residues = 0;
    size_t unknown_mapped = 0;
    double ms = 0;
    void release() {
        std::vector<const char *>().swap( text );
        std::vector<uint64_t>().swap( rec_len );
        std::vector<uint64_t>().swap( lens_all );
        std::vector<uint8_t>().swap( entry_codes );
        std::vector<uint64_t>().swap( entry_off );
        std::vector<uint32_t>().swap( sorted );
        std::vector<int32_t>().swap( sorted_pos_of_id );
    }
} g_gdb;

bool shard_is_current( const Ctx & c ) {
    const bool translate_db = g_symtype == TRANS_DB || g_symtype == TRANS_BOTH;
    const int gencode_key = translate_db ? g_db_gencode : 0;
    const int strands_key = ( g_symtype == NUCLEOTIDE || translate_db ) ? g_query_strands : 0;
    const PackedDb & db = c.db;
    return db.valid && db.epoch == g_db_epoch && db.symtype == g_symtype && db.rank == c.rank && db.world == c.world &&
           db.strands == strands_key && db.db_gencode == gencode_key;
}

void build_global() {
    GlobalDb & g = g_gdb;
    const double t0 = wall_ms();
    const bool translate_db = g_symtype == TRANS_DB || g_symtype == TRANS_BOTH;
    g.symtype = g_symtype;
    g.gencode_key = translate_db ? g_db_gencode : 0;
    g.strands_key = ( g_symtype == NUCLEOTIDE || translate_db ) ? g_query_strands : 0;
    g.map = ( g_symtype == NUCLEOTIDE || translate_db ) ? map_ncbi_nt16 : map_ncbi_aa;
    const signed char * map = g.map;
    int bm = 1;  // reference adp_init, src/db_adapter.c:118-139
    if( g_symtype == NUCLEOTIDE )
        bm = ( g_query_strands & COMPLEMENTARY_STRAND ) ? 2 : 1;
    else if( translate_db )
        bm = ( g_query_strands ^ BOTH_STRANDS ) ? 3 : 6;
    g.general = bm > 1 || translate_db;
    g.buffer_max = bm;

    g.text.clear();
    g.rec_len.clear();
    g.lens_all.clear();
    g.entry_codes.clear();
    g.entry_off.clear();
    g.unknown_mapped = 0;
    std::vector<const char *> & text = g.text;
    std::vector<uint64_t> & rec_len = g.rec_len;
    for( size_t i = 0;; i++ ) {  // the reference walks ids until the provider returns NULL (db_adapter.c:224-229)
        p_seqinfo rec = ssa_db_get_sequence( i );
        if( !rec ) break;
        rec_len.push_back( rec->seqlen );
        text.push_back( rec->seq );
    }
    if( !g.general ) {
        g.lens_all = rec_len;
    }
    else {
        // entries of record r: nucleotide = forward, reverse complement; translated = frames 0-2 of one strand or of
        // both (db_adapter.c:46-107).  Entries that come out empty are skipped like empty records.
        const size_t nrec = rec_len.size();
        if( nrec * (size_t) bm > 0xfffffff0ull ) fatal( "libssa_b200: more than 2^32 database entries are not supported." );
        g.lens_all.assign( nrec * bm, 0 );
        g.entry_off.assign( nrec * bm + 1, 0 );
        for( size_t r = 0; r < nrec; r++ )
            for( int b = 0; b < bm; b++ ) {
                uint64_t len = rec_len[r];
                if( translate_db ) len = len >= (uint64_t) ( b % 3 ) ? ( len - ( b % 3 ) ) / 3 : 0;
                g.lens_all[r * bm + b] = len;
                g.entry_off[r * bm + b + 1] = g.entry_off[r * bm + b] + len;
            }
        g.entry_codes.resize( g.entry_off.back() + 1 );
        std::vector<size_t> unknown_per( ( nrec + 255 ) / 256 + 1, 0 );
        parallel_for( nrec, 256, [&]( size_t r0, size_t r1 ) {
            std::vector<uint8_t> nt, tmp;
            size_t unknown = 0;
            for( size_t r = r0; r < r1; r++ ) {
                if( rec_len[r] == 0 ) continue;
                nt.resize( rec_len[r] + 1 );
                unknown += map_db_sequence( text[r], rec_len[r], nt.data(), map );
                for( int b = 0; b < bm; b++ ) {
                    uint8_t * dst = g.entry_codes.data() + g.entry_off[r * bm + b];
                    if( !translate_db ) {
                        if( b == 0 )
                            memcpy( dst, nt.data(), rec_len[r] );
                        else
                            revcompl_codes( nt.data(), rec_len[r], dst );
                    }
                    else {
                        const int strand = ( bm == 6 ) ? b / 3 : g_query_strands - 1;
                        translate_codes( true, nt.data(), rec_len[r], strand, b % 3, tmp );
                        if( tmp.size() != g.lens_all[r * bm + b] ) fatal( "libssa_b200: internal error: translated length." );
                        if( !tmp.empty() ) memcpy( dst, tmp.data(), tmp.size() );
                    }
                }
            }
            unknown_per[r0 / 256] = unknown;
        } );
        g.unknown_mapped = std::accumulate( unknown_per.begin(), unknown_per.end(), (size_t) 0 );
    }
    const size_t nall = g.lens_all.size();
    if( nall > 0xfffffff0ull ) fatal( "libssa_b200: more than 2^32 database records are not supported." );

    std::vector<uint32_t> & sorted = g.sorted;
    sorted.clear();
    sorted.reserve( nall );
    g.n_residues = 0;
    for( size_t i = 0; i < nall; i++ )
        if( g.lens_all[i] > 0 ) {
            sorted.push_back( (uint32_t) i );
            g.n_residues += g.lens_all[i];
            if( g.lens_all[i] > 0x7fffffffull ) fatal( "libssa_b200: database record %ld is too long.", (long) i );
        }
    g.n_seqs = sorted.size();
    std::stable_sort( sorted.begin(), sorted.end(), [&]( uint32_t a, uint32_t b ) { return g.lens_all[a] > g.lens_all[b]; } );
    g.sorted_pos_of_id.assign( nall, -1 );
    for( size_t i = 0; i < sorted.size(); i++ ) g.sorted_pos_of_id[sorted[i]] = (int32_t) i;
    g.ms = wall_ms() - t0;
}

// Packs and uploads the shard of the calling thread's context from g_gdb.  Returns the number of unknown symbols it mapped to 0.
size_t build_shard() {
    PackedDb & db = ctx().db;
    const GlobalDb & g = g_gdb;
    const double t0 = wall_ms();
    const signed char * map = g.map;
    const bool general = g.general;
    db.buffer_max = g.buffer_max;
    for( auto & kv : db.arrival ) kv.second.d_pos_of_sorted.release();
    db.arrival.clear();
    db.lens_all = g.lens_all;
    db.n_residues = g.n_residues;
    db.n_seqs = g.n_seqs;
    const size_t nall = db.lens_all.size();
    const std::vector<uint32_t> & sorted = g.sorted;
    const std::vector<const char *> & text = g.text;
    const std::vector<uint8_t> & entry_codes = g.entry_codes;
    const std::vector<uint64_t> & entry_off = g.entry_off;

    // shard: deal the length-sorted records round-robin
    db.sorted_all = sorted;
    db.sorted_pos_of_id = g.sorted_pos_of_id;
    db.sorted_id.clear();
    db.sorted_len.clear();
    db.local_residues = 0;
    for( size_t s = (size_t) ctx().rank; s < sorted.size(); s += (size_t) ctx().world ) {
        db.sorted_id.push_back( sorted[s] );
        db.sorted_len.push_back( (uint32_t) db.lens_all[sorted[s]] );
        db.local_residues += db.lens_all[sorted[s]];
    }
    const size_t nloc = db.sorted_id.size();
    db.npairs = ( nloc + 1 ) / 2;
    db.pair_col.assign( db.npairs + 1, 0 );
    // every pair occupies an even number of stream columns (the kernels advance two columns per step)
    for( size_t p = 0; p < db.npairs; p++ ) db.pair_col[p + 1] = db.pair_col[p] + ( ( (uint64_t) db.sorted_len[2 * p] + 1 ) & ~1ull );
    db.total_cols = db.pair_col[db.npairs];

    // column stream: bits 0-4 residue of A, 5-9 residue of B (0 past its end, like the reference's padding,
    // src/algo/8/search_8_util.h:60-74), 10 first column, 11 last column of B, 12 last column of A, 15 valid column
    PinBuf<uint16_t> h_stream;
    h_stream.reserve( db.total_cols + 64 );
    std::vector<size_t> unknown_per_chunk( ( db.npairs + 1023 ) / 1024 + 1, 0 );
    parallel_for( db.npairs, 1024, [&]( size_t p0, size_t p1 ) {
        size_t unknown = 0;
        for( size_t p = p0; p < p1; p++ ) {
            const uint32_t la = db.sorted_len[2 * p];
            const bool has_b = 2 * p + 1 < nloc;
            const uint32_t lb = has_b ? db.sorted_len[2 * p + 1] : 0;
            const unsigned char * ta = general ? entry_codes.data() + entry_off[db.sorted_id[2 * p]]
                                               : (const unsigned char *) text[db.sorted_id[2 * p]];
            const unsigned char * tb = !has_b ? nullptr
                    : general ? entry_codes.data() + entry_off[db.sorted_id[2 * p + 1]]
                              : (const unsigned char *) text[db.sorted_id[2 * p + 1]];
            uint16_t * out = h_stream.p + db.pair_col[p];
            for( uint32_t c = 0; c < la; c++ ) {
                int a = general ? (int) ta[c] : map[ta[c]];
                if( a < 0 ) {
                    a = 0;
                    unknown++;
                }
                int b = 0;
                if( c < lb ) {
                    b = general ? (int) tb[c] : map[tb[c]];
                    if( b < 0 ) {
                        b = 0;
                        unknown++;
                    }
                }
                uint32_t v = (uint32_t) a | ( (uint32_t) b << 5 ) | COL_VALID;
                if( c == 0 ) v |= COL_FIRST;
                if( c + 1 == lb ) v |= COL_LAST_B;
                if( c + 1 == la ) v |= COL_LAST_A;
                out[c] = (uint16_t) v;
            }
            // padding column of an odd-length pair: a processed column of PAD_RESIDUE in both halves (score_kernel.cuh)
            if( la & 1 ) out[la] = (uint16_t) ( PAD_RESIDUE | ( PAD_RESIDUE << 5 ) | COL_VALID );
        }
        unknown_per_chunk[p0 / 1024] = unknown;
    } );
    const size_t unknown = std::accumulate( unknown_per_chunk.begin(), unknown_per_chunk.end(), (size_t) 0 );
    for( size_t i = db.total_cols; i < db.total_cols + 64; i++ ) h_stream.p[i] = 0;

    for( auto & kv : db.batch_tables ) kv.second.d_batches.release();
    db.batch_tables.clear();
    db.pack_ms = wall_ms() - t0;

    const double t1 = wall_ms();
    db.d_stream.reserve( db.total_cols + 64 );
    CUDA_OK( cudaMemcpyAsync( db.d_stream.p, h_stream.p, ( db.total_cols + 64 ) * sizeof( uint16_t ), cudaMemcpyHostToDevice,
            ctx().stream ) );
    {
        std::vector<uint32_t> lens32( nall );
        for( size_t i = 0; i < nall; i++ ) lens32[i] = (uint32_t) db.lens_all[i];
        db.d_lens_all.reserve( nall + 1 );
        db.d_sorted_pos_of_id.reserve( nall + 1 );
        if( nall ) {
            CUDA_OK( cudaMemcpy( db.d_lens_all.p, lens32.data(), nall * 4, cudaMemcpyHostToDevice ) );
            CUDA_OK( cudaMemcpy( db.d_sorted_pos_of_id.p, db.sorted_pos_of_id.data(), nall * 4, cudaMemcpyHostToDevice ) );
        }
        db.d_sorted_all.reserve( sorted.size() + 1 );
        if( !sorted.empty() ) CUDA_OK( cudaMemcpy( db.d_sorted_all.p, sorted.data(), sorted.size() * 4, cudaMemcpyHostToDevice ) );
        db.d_pair_col.reserve( db.npairs + 2 );
        db.d_sorted_len.reserve( nloc + 2 );
        CUDA_OK( cudaMemcpy( db.d_pair_col.p, db.pair_col.data(), ( db.npairs + 1 ) * sizeof( uint64_t ), cudaMemcpyHostToDevice ) );
        if( nloc ) CUDA_OK( cudaMemcpy( db.d_sorted_len.p, db.sorted_len.data(), nloc * sizeof( uint32_t ), cudaMemcpyHostToDevice ) );
        for( auto & kv : db.chunk_info ) kv.second.d_chunk_base.release();
        db.chunk_info.clear();
    }
    CUDA_OK( cudaStreamSynchronize( ctx().stream ) );
    h_stream.release();
    db.upload_ms = wall_ms() - t1;

    db.epoch = g_db_epoch;
    db.symtype = g.symtype;
    db.strands = g.strands_key;
    db.db_gencode = g.gencode_key;
    db.rank = ctx().rank;
    db.world = ctx().world;
    db.valid = true;
    ctx().stats.n_seqs = db.n_seqs;
    ctx().stats.n_residues = db.n_residues;
    ctx().stats.local_seqs = nloc;
    ctx().stats.local_residues = db.local_residues;
    ctx().stats.pack_ms = db.pack_ms;
    ctx().stats.upload_ms = db.upload_ms;
    ctx().stats.db_device_bytes = ( db.total_cols + 64 ) * sizeof( uint16_t );
    print_info( "GPU %d database: %ld sequences, %ld residues, %ld columns in %ld pairs, packed in %.1f ms, uploaded in %.1f ms\n",
            ctx().device, (long) nloc, (long) db.local_residues, (long) db.total_cols, (long) db.npairs, db.pack_ms, db.upload_ms );
    return unknown;
}

// Builds (once per init_db / init_symbol_translation / shard change) the device-resident shard of every context.
void build_db() {
    ensure_ctx();
    bool current = true;
    for( Ctx * c : g_devs ) current = current && shard_is_current( *c );
    if( current ) return;
    const double t0 = wall_ms();
    build_global();
    std::atomic<size_t> unknown( g_gdb.unknown_mapped );
    run_on_all( [&]() { unknown += build_shard(); } );   // the provider is no longer touched: shards pack side by side
    if( unknown > 0 ) print_warning( "%ld unknown symbols found and set to zero", (long) unknown.load() );
    g_gdb.release();
    // the primary's record describes the whole process: shard sizes summed, wall clock of the whole build
    ssa_b200_stats_t & st = g_devs[0]->stats;
    if( g_devs.size() > 1 ) {
        st.local_seqs = 0;
        st.local_residues = 0;
        st.db_device_bytes = 0;
        double upload = 0;
        for( Ctx * c : g_devs ) {
            st.local_seqs += c->db.sorted_id.size();
            st.local_residues += c->db.local_residues;
            st.db_device_bytes += ( c->db.total_cols + 64 ) * sizeof( uint16_t );
            upload = std::max( upload, c->db.upload_ms );
        }
        st.upload_ms = upload;
        st.pack_ms = wall_ms() - t0 - upload;
    }
    else {
        st.pack_ms += g_gdb.ms;
    }
}

void parallel_for( size_t n, size_t grain, const std::function<void( size_t, size_t )> & fn ) {
    size_t nchunks = ( n + grain - 1 ) / grain;
    unsigned hw = std::thread::hardware_concurrency();
    size_t nthreads = std::min<size_t>( hw ? hw : 4, nchunks );
    if( nthreads <= 1 ) {
        for( size_t c = 0; c < nchunks; c++ ) fn( c * grain, std::min( n, ( c + 1 ) * grain ) );
        return;
    }
    std::vector<std::thread> pool;
    std::atomic<size_t> next( 0 );
    for( size_t t = 0; t < nthreads; t++ )
        pool.emplace_back( [&]() {
            for( ;; ) {
                size_t c = next.fetch_add( 1 );
                if( c >= nchunks ) break;
                fn( c * grain, std::min( n, ( c + 1 ) * grain ) );
            }
        } );
    for( auto & th : pool ) th.join();
}

// Work batches of the packed tier: runs of whole pairs with at least `cols` columns, built on demand and cached.
// Small batches balance the persistent grid's work queue, large ones amortise the G-1 fill/drain steps.
// First pair of every work batch: whole pairs, at least `want` stream columns each (the last batch may be shorter), where
// `want` = cols, or with groups > 0 min( cols, max( min_cols, columns left / (2 groups) ) ).  pair_col[npairs + 1].
std::vector<uint32_t> partition_batches( const uint64_t * pair_col, size_t npairs, size_t first_pair, uint32_t cols, uint32_t groups,
        uint32_t min_cols ) {
    std::vector<uint32_t> first;
    const uint64_t total = pair_col[npairs];
    for( size_t p = first_pair; p < npairs; ) {
        uint64_t want = cols;
        if( groups > 0 ) want = std::min<uint64_t>( cols, std::max<uint64_t>( min_cols, ( total - pair_col[p] ) / ( 2ull * groups ) ) );
        size_t q = p;
        while( q < npairs && pair_col[q] - pair_col[p] < want ) q++;
        first.push_back( (uint32_t) p );
        p = q;
    }
    return first;
}

// Work batches of ~`cols` stream columns (whole pairs), starting at pair `first_pair`; cached per database.
// The queue is consumed in table order by `groups` resident lane groups.  Towards the end of the table the batches shrink
// (guided self-scheduling: at most 1/(2 groups) of the columns that are left, never below `min_cols`), so that the groups
// finish within a small batch of each other instead of a full one - with ~8 batches per group the idle tail of a launch
// drops from ~6 % to < 1 % of its duration, for a few more wavefront fills.  groups = 0: uniform batches.
PackedDb::BatchTable & batch_table( uint32_t cols, uint32_t first_pair = 0, uint32_t groups = 0, uint32_t min_cols = 512 ) {
    PackedDb & db = ctx().db;
    if( getenv( "LIBSSA_B200_NO_TAPER" ) ) groups = 0;
    const std::array<uint32_t, 4> key = { cols, first_pair, groups, min_cols };
    auto it = db.batch_tables.find( key );
    if( it != db.batch_tables.end() ) return it->second;
    PackedDb::BatchTable & bt = db.batch_tables[key];
    bt.pair0 = partition_batches( db.pair_col.data(), db.npairs, first_pair, cols, groups, min_cols );
    for( size_t i = 0; i < bt.pair0.size(); i++ ) {
        const size_t p = bt.pair0[i], q = i + 1 < bt.pair0.size() ? bt.pair0[i + 1] : db.npairs;
        Batch b;
        memset( &b, 0, sizeof( b ) );
        b.col0 = db.pair_col[p];
        b.bnd0 = db.pair_col[p];
        b.out0 = (uint32_t) p;
        if( db.pair_col[q] - db.pair_col[p] > 0x7fffffffull ) fatal( "libssa_b200: batch too long." );
        b.ncols = (uint32_t) ( db.pair_col[q] - db.pair_col[p] );
        bt.batches.push_back( b );
    }
    bt.d_batches.reserve( bt.batches.size() + 1 );
    if( !bt.batches.empty() )
        CUDA_OK( cudaMemcpyAsync( bt.d_batches.p, bt.batches.data(), bt.batches.size() * sizeof( Batch ), cudaMemcpyHostToDevice, ctx().stream ) );
    CUDA_OK( cudaStreamSynchronize( ctx().stream ) );
    return bt;
}

// ------------------------------------------------------------------------------------------------
// small device kernels around the scoring kernels
// ------------------------------------------------------------------------------------------------

// (A:B) int16 pairs, stored by stream position -> one int32 score per sorted position.  Speculative packed tier (Smith-Waterman): the pairs below
// `spec_pairs` are not provably inside int16; a sequence whose running maximum came within one substitution score of the
// int16 limit (>= spec_thr) may have wrapped and is listed for the exact 32-bit tier - the reference's own scheme of
// "run narrow, detect, recompute" (reference src/algo/simd/search_simd_sw.c:372-379,422-423).
__global__ void unpack_scores_kernel( const uint32_t * packed, const uint64_t * pair_col, int32_t * scores, uint32_t pair0, uint32_t npairs,
        uint32_t nseq, int32_t offset, uint32_t spec_pairs, int32_t spec_thr, uint32_t * spec_list, uint32_t * spec_count ) {
    uint32_t p = pair0 + blockIdx.x * blockDim.x + threadIdx.x;
    if( p >= npairs ) return;
    uint32_t w = packed[pair_col[p + 1] / 2 - 1];   // the kernels store a pair's result at the column pair in which it ends
    const int32_t sa = (int32_t) (int16_t) ( w >> 16 ), sb = (int32_t) (int16_t) ( w & 0xffffu );
    scores[2 * p] = sa + offset;
    if( 2 * p + 1 < nseq ) scores[2 * p + 1] = sb + offset;
    if( p < spec_pairs ) {
        if( sa >= spec_thr ) spec_list[atomicAdd( spec_count, 1u )] = 2 * p;
        if( 2 * p + 1 < nseq && sb >= spec_thr ) spec_list[atomicAdd( spec_count, 1u )] = 2 * p + 1;
    }
}

// Two queries per word: (q1 : q2) int16 pairs per SEQUENCE, stored by the stream position in which the sequence ends (one
// array per launch half: sequence A resp. B of every pair) -> one int32 score per sorted position and query.
__global__ void unpack_dual_kernel( const uint32_t * packed, uint32_t out_stride, const uint64_t * pair_col, const uint32_t * sorted_len,
        int32_t * scores1, int32_t * scores2, uint32_t npairs, uint32_t nseq, int32_t offset ) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if( p >= npairs ) return;
    const uint32_t wa = packed[pair_col[p + 1] / 2 - 1];
    scores1[2 * p] = (int32_t) (int16_t) ( wa >> 16 ) + offset;
    scores2[2 * p] = (int32_t) (int16_t) ( wa & 0xffffu ) + offset;
    if( 2 * p + 1 < nseq ) {
        const uint32_t wb = packed[out_stride + ( pair_col[p] + sorted_len[2 * p + 1] - 1 ) / 2];
        scores1[2 * p + 1] = (int32_t) (int16_t) ( wb >> 16 ) + offset;
        scores2[2 * p + 1] = (int32_t) (int16_t) ( wb & 0xffffu ) + offset;
    }
}

// the output slots of a list of 32-bit-tier batches set to zero (the passes of the pipelined tier combine by atomicMax)
__global__ void zero_slots_kernel( const Batch * batches, uint32_t n, int32_t * out ) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if( i < n ) out[batches[i].out0] = 0;
}

// work batches of the 32-bit tier for the sequences the speculative packed tier listed (at most `cap`, a host-side bound);
// zero_out: the listed sequences' scores are reset (pipelined tier)
__global__ void build_spec_batches_kernel( const uint32_t * list, const uint32_t * count, uint32_t cap, const uint64_t * pair_col,
        const uint32_t * sorted_len, Batch * batches, uint32_t * nbatches, int32_t * zero_out ) {
    if( blockIdx.x != 0 || threadIdx.x != 0 ) return;
    const uint32_t n = min( *count, cap );
    uint64_t bnd = 0;
    for( uint32_t i = 0; i < n; i++ ) {
        const uint32_t s = list[i];
        Batch b;
        b.col0 = pair_col[s / 2];
        b.bnd0 = bnd;
        b.ncols = ( sorted_len[s] + 1u ) & ~1u;
        b.out0 = s;
        b.half = s & 1u;
        b.pad = 0;
        bnd += b.ncols;
        batches[i] = b;
        if( zero_out ) zero_out[s] = 0;
    }
    *nbatches = n;
}

// wide-tier query profile (layout of fill_profile_wide) built on the device: [pass][row][residue][lane]
__global__ void fill_profile_wide_kernel( const uint8_t * query, uint32_t qlen, const int32_t * matrix, int G, int R, int npass,
        int32_t pad_score, uint32_t * out ) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t words = (uint32_t) R * 1024u;
    if( i >= words * (uint32_t) npass ) return;
    const uint32_t pass = i / words, w = i % words;
    const uint32_t lane = w & 31u, d = ( w >> 5 ) & 31u, r = w >> 10;
    const uint32_t row = ( pass * (uint32_t) G + ( lane % (uint32_t) G ) ) * (uint32_t) R + r;
    out[i] = (uint32_t) ( row < qlen && d != PAD_RESIDUE ? matrix[( d << 5 ) + query[row]] : pad_score );
}

// scores by (rank, local sorted position) -> scores by heap-arrival position.  `gathered` holds `world` blocks of
// `stride` scores (the all-gathered shard results; world = 1: the local scores); shard r's local position s is the
// global sorted position s*world + r.
__global__ void scatter_arrival_kernel( const int32_t * gathered, const int32_t * pos, int32_t * out, uint32_t stride,
        uint32_t world, uint32_t nglobal ) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if( i >= stride * world ) return;
    uint32_t r = i / stride, sl = i % stride;
    uint64_t g = (uint64_t) sl * world + r;
    if( g >= nglobal ) return;
    int32_t a = pos[g];
    if( a >= 0 ) out[a] = gathered[i];
}

// gathered indices (r*stride + s) of the records whose score reaches thr: candidates of the reference's overflow cascade
__global__ void flag_overflow_kernel( const int32_t * gathered, uint32_t stride, uint32_t world, uint32_t nglobal, int32_t thr,
        uint32_t * list, uint32_t * count, uint32_t cap ) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if( i >= stride * world ) return;
    uint32_t r = i / stride, sl = i % stride;
    if( (uint64_t) sl * world + r >= nglobal ) return;
    if( gathered[i] >= thr ) {
        uint32_t k = atomicAdd( count, 1u );
        if( k < cap ) list[k] = i;
    }
}

// first-crossing columns of the flagged records only: out[k] = first8[idx], out[count + k] = first16[idx]
__global__ void gather_first_kernel( const uint32_t * list, uint32_t count, const int32_t * first_all, uint32_t stride, int32_t * out ) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if( k >= count ) return;
    const uint32_t i = list[k], r = i / stride, sl = i % stride;
    out[k] = first_all[(size_t) r * 2 * stride + sl];
    out[count + k] = first_all[(size_t) r * 2 * stride + stride + sl];
}

__global__ void iota_kernel( uint32_t * p, uint32_t first, uint32_t n ) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if( i < n ) p[i] = first + i;
}

// counts the entries that are not INT_MAX (records that overflow a tier)
// NW entered at 8 / 16 bit: a record the narrow tier accepts (first == INT_MAX among the n emulated ones) is reported with THAT
// tier's score (nw_sat.cuh), not with the exact one
__global__ void nw_apply_sat_kernel( const int32_t * first, const int32_t * sat, uint32_t n, int32_t * scores ) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if( i < n && first[i] == INT_MAX ) scores[i] = sat[i];
}

__global__ void count_set_kernel( const int32_t * p, uint32_t n, uint32_t * out ) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if( i < n && p[i] != 0x7fffffff ) atomicAdd( out, 1u );
}

// sparse heap replay: arrival positions >= first whose score exceeds thr (any order; few)
__global__ void compact_greater_kernel( const int32_t * scores, uint32_t first, uint32_t n, int32_t thr, uint32_t * pos_out, int32_t * score_out,
        uint32_t * counter ) {
    const uint32_t i = first + blockIdx.x * blockDim.x + threadIdx.x;
    if( i < n && scores[i] > thr ) {
        const uint32_t j = atomicAdd( counter, 1u );
        pos_out[j] = i;
        score_out[j] = scores[i];
    }
}

__global__ void gather_u32_kernel( const uint32_t * table, const uint32_t * idx, uint32_t n, uint32_t * out ) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if( i < n ) out[i] = table[idx[i]];
}

__global__ void fill_i32_kernel( int32_t * p, int32_t v, uint32_t n ) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if( i < n ) p[i] = v;
}

// ------------------------------------------------------------------------------------------------
// query profile
// ------------------------------------------------------------------------------------------------

struct Plan {
    int G = 32, R = 4, npass = 1, qpad = 0;                      // packed tier: lanes per group, rows per lane, passes
    int wide_G = 32, wide_R = 8, wide_npass = 1, wide_qpad = 0;   // wide tier
    int pad_score = 0;
    int64_t lo = 0, hi = 0;  // matrix extremes
};

// Picks the systolic-array shape for a query: G lanes x R rows per pass.  Model of the instruction count per
// packed word-row: ~7.5 instructions of DP + lookup, plus ~60 instructions of per-step work (shuffles, flags,
// feed) shared by the R rows of a lane; rows beyond the query are wasted work, every pass re-streams the database.
void pick_shape( size_t qlen, const int * r_options, int n_options, int * G_out, int * R_out, int * npass_out ) {
    double best = 1e300;
    for( int G = 4; G <= 32; G *= 2 )
        for( int i = 0; i < n_options; i++ ) {
            const int R = r_options[i];
            const size_t rows = (size_t) G * R;
            const size_t npass = ( qlen + rows - 1 ) / rows;
            if( npass > 1 && G != 32 ) continue;  // several passes only with full-warp arrays
            const double cost = (double) npass * rows * ( 7.5 + 60.0 / R ) + ( npass > 1 ? 0.03 * npass * rows * 7.5 : 0.0 );
            if( cost < best ) {
                best = cost;
                *G_out = G;
                *R_out = R;
                *npass_out = (int) npass;
            }
        }
}

Plan make_plan( size_t qlen ) {
    Plan pl;
    pick_shape( qlen, PACKED_R, sizeof( PACKED_R ) / sizeof( int ), &pl.G, &pl.R, &pl.npass );
    if( const char * ov = getenv( "LIBSSA_B200_SHAPE" ) ) {  // development aid: "G,R" overrides the packed-tier shape
        int g = 0, r = 0;
        if( sscanf( ov, "%d,%d", &g, &r ) == 2 && ( g == 4 || g == 8 || g == 16 || g == 32 ) ) {
            for( int opt : PACKED_R )
                if( opt == r ) {
                    pl.G = g;
                    pl.R = r;
                    pl.npass = (int) ( ( qlen + (size_t) g * r - 1 ) / ( (size_t) g * r ) );
                }
        }
    }
    pl.qpad = pl.npass * pl.G * pl.R;
    pick_shape( qlen, WIDE_R, sizeof( WIDE_R ) / sizeof( int ), &pl.wide_G, &pl.wide_R, &pl.wide_npass );
    pl.wide_qpad = pl.wide_npass * pl.wide_G * pl.wide_R;
    pl.lo = g_matrix.lo;
    pl.hi = g_matrix.hi;
    pl.pad_score = (int) std::min<int64_t>( pl.lo, 0 );
    return pl;
}

// Packed-tier profile of one pass (layout: score_kernel.cuh): G lanes x R rows starting at query row `rowbase`; lane l of a
// group owns rows rowbase + (l mod G) * R ....  All 32 lanes have their own copy (groups repeat), so a warp-wide lookup is
// bank-conflict free whatever residues the lanes hold.  First MR rows of a lane: 128-bit entries [row quad][residue][lane] =
// four rows, one zero-extended int16 each.  Remaining rows: 64-bit entries [row quad][residue][lane] = four rows, two per
// word (low half = even row).  Rows beyond the query get pad_score.
size_t packed_profile_words( int R ) { return (size_t) merge_rows( R ) * 1024 + (size_t) ( ( R - merge_rows( R ) ) / 2 ) * 1024; }

// Packed Needleman-Wunsch shifts the query DOWN by `shift` rows so that its last row is the last row of a lane (the result is
// then the value the lane hands to the next one - no row has to be picked out of the registers, no column has to be run on its
// own; nw_row_shift).  The `shift` VIRTUAL rows on top score gapE against every residue: with start values gapO (0 in the last
// of them) they carry the matrix' top boundary H(-1, j) = gapO + (j + 1) gapE down unchanged - H = diagonal + gapE, E and F stay
// below - and hand row 0 exactly the boundary it expects (H, and F = H + gapO + gapE).  Position p of the array = query row p - shift.
void fill_profile_packed( uint32_t * out, const uint8_t * q, size_t qlen, int G, int R, size_t rowbase, int pad_score, size_t shift = 0 ) {
    const int MR = merge_rows( R );
    uint32_t * out16 = out + (size_t) MR * 1024;
    for( int d = 0; d < 32; d++ )
        for( int lane = 0; lane < 32; lane++ ) {
            const size_t row0 = rowbase + (size_t) ( lane % G ) * R;
            auto score = [&]( int r ) {
                const size_t pos = row0 + r;
                if( pos < shift ) return (int) g_gapE;
                return pos - shift < qlen && d != (int) PAD_RESIDUE ? (int) g_matrix.m[( d << 5 ) + q[pos - shift]] : pad_score;
            };
            for( int r = 0; r < MR; r++ )
                out[( ( ( r / 4 ) * 32 + d ) * 32 + lane ) * 4 + ( r & 3 )] = (uint32_t) (uint16_t) (int16_t) score( r );
            for( int r = MR; r < R; r += 2 ) {
                const int k = ( r - MR ) / 2;   // row pair within the 64-bit part
                out16[( ( ( k / 2 ) * 32 + d ) * 32 + lane ) * 2 + ( k & 1 )] =
                        ( (uint32_t) (uint16_t) (int16_t) score( r + 1 ) << 16 ) | (uint32_t) (uint16_t) (int16_t) score( r );
            }
        }
}

// Profile of the two-queries-per-word kernels: 128-bit entries [row quad][residue][lane] of four rows, each word = (score of
// query 1 : score of query 2) for that row and residue; rows beyond a query's end hold pad_score in its half.
// shift1 / shift2: NW, virtual rows on top of each query (fill_profile_packed)
void fill_profile_dual( uint32_t * out, const uint8_t * q1, size_t len1, const uint8_t * q2, size_t len2, int G, int R, int pad_score,
        size_t shift1 = 0, size_t shift2 = 0 ) {
    auto score = [&]( const uint8_t * q, size_t len, size_t shift, size_t pos, int d ) {
        if( pos < shift ) return (int) g_gapE;
        return pos - shift < len && d != (int) PAD_RESIDUE ? (int) g_matrix.m[( d << 5 ) + q[pos - shift]] : pad_score;
    };
    for( int d = 0; d < 32; d++ )
        for( int lane = 0; lane < 32; lane++ ) {
            const size_t row0 = (size_t) ( lane % G ) * R;
            for( int r = 0; r < R; r++ ) {
                const size_t row = row0 + r;
                const int s1 = score( q1, len1, shift1, row, d ), s2 = score( q2, len2, shift2, row, d );
                out[( ( ( r / 4 ) * 32 + d ) * 32 + lane ) * 4 + ( r & 3 )] = ( (uint32_t) (uint16_t) (int16_t) s1 << 16 ) | (uint32_t) (uint16_t) (int16_t) s2;
            }
        }
}

void fill_profile_wide( uint32_t * out, const uint8_t * q, size_t qlen, const Plan & pl, int pass ) {
    const int R = pl.wide_R, G = pl.wide_G;
    for( int r = 0; r < R; r++ )
        for( int d = 0; d < 32; d++ )
            for( int lane = 0; lane < 32; lane++ ) {
                size_t row = ( (size_t) pass * G + (size_t) ( lane % G ) ) * R + r;
                int v = row < qlen && d != (int) PAD_RESIDUE ? (int) g_matrix.m[( d << 5 ) + q[row]] : pl.pad_score;
                out[( r * 32 + d ) * 32 + lane] = (uint32_t) v;
            }
}

// Longest database sequence the packed (int16) tier is guaranteed exact for, and the NW offset.
// SW: 0 <= H <= hi+ * min(q, d); E, F >= gapO + gapE.  NW: with L(i,j) = 2 gapO + (i+j+2) gapE
// a lower bound of every H (all-gap path) and hi+ * min(i+1, j+1) an upper bound, all of H, E, F and the diagonal
// candidates stay inside [ L(q-1,d-1) + min(lo, gapO+gapE), hi+ * (min(q,d) + 1) ].
int64_t packed_safe_len( const Plan & pl, bool nw, int qrows, int32_t * offset ) {
    const int64_t hip = std::max<int64_t>( pl.hi, 0 );
    const int64_t lom = std::min<int64_t>( std::min<int64_t>( pl.lo, 0 ), (int64_t) g_gapO + g_gapE );
    *offset = 0;
    if( !nw ) {
        if( hip == 0 ) return INT64_MAX;
        // hi+ * (min(q,d) + 1) <= 32767
        int64_t m = 32767 / hip - 1;
        if( m >= qrows ) return INT64_MAX;
        return m < 0 ? 0 : m;
    }
    // find the largest d with hi(d) - lo(d) <= 65000
    auto span = [&]( int64_t d, int64_t * lo_out, int64_t * hi_out ) {
        int64_t hi = hip * ( std::min<int64_t>( qrows, d ) + 1 );
        int64_t lo = 2 * (int64_t) g_gapO + ( (int64_t) qrows + d + 2 ) * (int64_t) g_gapE + lom + std::min<int64_t>( 2 * (int64_t) g_gapO, 0 );
        if( lo > 0 ) lo = 0;
        *lo_out = lo;
        *hi_out = hi;
        return hi - lo;
    };
    int64_t lo, hi;
    int64_t d_lo = 0, d_hi = 1 << 30;
    if( span( 1, &lo, &hi ) > 65000 ) return 0;
    while( d_lo + 1 < d_hi ) {  // span is monotone in d
        int64_t mid = ( d_lo + d_hi ) / 2;
        if( span( mid, &lo, &hi ) <= 65000 )
            d_lo = mid;
        else
            d_hi = mid;
    }
    return d_lo;
}

int32_t nw_offset_for( const Plan & pl, int qrows, int64_t dmax ) {
    const int64_t hip = std::max<int64_t>( pl.hi, 0 );
    const int64_t lom = std::min<int64_t>( std::min<int64_t>( pl.lo, 0 ), (int64_t) g_gapO + g_gapE );
    int64_t hi = hip * ( std::min<int64_t>( qrows, dmax ) + 1 );
    int64_t lo = 2 * (int64_t) g_gapO + ( (int64_t) qrows + dmax + 2 ) * (int64_t) g_gapE + lom + std::min<int64_t>( 2 * (int64_t) g_gapO, 0 );
    if( lo > 0 ) lo = 0;
    return (int32_t) ( ( hi + lo ) / 2 );  // centre the value range on 0
}

// ------------------------------------------------------------------------------------------------
// scoring of the whole shard: scores by sorted position in ctx().d_scores
// ------------------------------------------------------------------------------------------------

void launch_fill( int32_t * p, int32_t v, size_t n ) {
    if( n == 0 ) return;
    fill_i32_kernel<<<(unsigned) ( ( n + 255 ) / 256 ), 256, 0, ctx().stream>>>( p, v, (uint32_t) n );
    ctx().stats.launches++;
}

// The pipelined 32-bit tier (wide_pipe.cu) serves every query of more than one of ITS passes (32 lanes x WIDE_PIPE_R rows),
// except the first-crossing tracker.  LIBSSA_B200_NO_WIDE_PIPE=1: one launch per pass as before (A/B measurements, tests).
bool wide_pipe_applies( size_t qlen ) { return qlen > (size_t) 32 * WIDE_PIPE_R && !getenv( "LIBSSA_B200_NO_WIDE_PIPE" ); }

Plan wide_pipe_plan( const Plan & pl, size_t qlen ) {
    Plan p = pl;
    p.wide_G = 32;
    p.wide_R = WIDE_PIPE_R;
    p.wide_npass = (int) ( ( qlen + (size_t) 32 * WIDE_PIPE_R - 1 ) / ( (size_t) 32 * WIDE_PIPE_R ) );
    p.wide_qpad = p.wide_npass * 32 * WIDE_PIPE_R;
    return p;
}

// batches in d_wide_batches[slot] (n of them, or *d_nb on the device), profiles in d_profile_wide[slot]; bnd_cols = boundary columns of all batches
void launch_wide_pipe_tier( int slot, uint32_t n, const uint32_t * d_nb, uint64_t bnd_cols, size_t qlen, int npass, bool nw, cudaStream_t stream ) {
    ctx().d_bnd_wide.reserve( 4 * bnd_cols + 256 );   // 2 buffers x ( bnd_cols / 2 + 1 ) entries x 4 words
    ctx().d_wide_progress[slot].reserve( (size_t) n * npass + 1 );
    ctx().d_counter.reserve( 64 );
    CUDA_OK( cudaMemsetAsync( ctx().d_wide_progress[slot].p, 0, ( (size_t) n * npass + 1 ) * 4, stream ) );
    CUDA_OK( cudaMemsetAsync( ctx().d_counter.p + 16 + 8 * slot, 0, 4, stream ) );
    WidePipeArgs a;
    memset( &a, 0, sizeof( a ) );
    a.stream = ctx().db.d_stream.p;
    a.batches = ctx().d_wide_batches[slot].p;
    a.nbatches = n;
    a.nbatches_dev = d_nb;
    a.counter = ctx().d_counter.p + 16 + 8 * slot;
    a.profile = ctx().d_profile_wide[slot].p;
    a.npass = (uint32_t) npass;
    a.gap_open = g_gapO;
    a.gap_extend = g_gapE;
    a.cap_pass = (int) ( ( qlen - 1 ) / ( (size_t) 32 * WIDE_PIPE_R ) );
    a.cap_lane = (int) ( ( ( qlen - 1 ) % ( (size_t) 32 * WIDE_PIPE_R ) ) / WIDE_PIPE_R );
    a.cap_row = (int) ( ( qlen - 1 ) % WIDE_PIPE_R );
    a.out = ctx().d_scores.p;
    a.bnd0 = (uint4 *) ctx().d_bnd_wide.p;
    a.bnd1 = a.bnd0 + ( bnd_cols / 2 + 1 );
    a.progress = ctx().d_wide_progress[slot].p;
    launch_wide_pipe( nw, a, stream );
    ctx().stats.launches++;
}

// runs the wide tier over `items` (slots = sorted positions); optional first-crossing tracking
// `slot` selects the staging buffers (0 = wide tier of a search, 1 = overflow tracking) so that a second use inside
// one search never overwrites pinned memory an earlier asynchronous copy has not consumed yet
void run_wide( const std::vector<uint32_t> & slots, const uint8_t * q, size_t qlen, const Plan & pl, bool nw, bool track,
        int32_t * d_first8, int32_t * d_first16, int slot, cudaStream_t stream ) {
    PackedDb & db = ctx().db;
    if( slots.empty() ) return;
    const size_t n = slots.size();
    ctx().h_wide_batches[g_stage_slot][slot].reserve( n );
    uint64_t bnd = 0;
    for( size_t i = 0; i < n; i++ ) {
        Batch b;
        memset( &b, 0, sizeof( b ) );
        uint32_t s = slots[i];
        b.col0 = db.pair_col[s / 2];
        b.bnd0 = bnd;
        b.ncols = ( db.sorted_len[s] + 1u ) & ~1u;   // whole column pairs
        b.out0 = s;
        b.half = s & 1;
        bnd += b.ncols;
        ctx().h_wide_batches[g_stage_slot][slot].p[i] = b;
    }
    ctx().d_wide_batches[slot].reserve( n );
    CUDA_OK( cudaMemcpyAsync( ctx().d_wide_batches[slot].p, ctx().h_wide_batches[g_stage_slot][slot].p, n * sizeof( Batch ), cudaMemcpyHostToDevice, stream ) );
    ctx().stats.h2d_bytes += n * sizeof( Batch );
    if( !track && wide_pipe_applies( qlen ) ) {   // several passes: all of them in one launch, as a pipeline per sequence
        const Plan plp = wide_pipe_plan( pl, qlen );
        const int npass = plp.wide_npass;
        const size_t words = (size_t) WIDE_PIPE_R * 1024;
        ctx().h_profile_wide[g_stage_slot][slot].reserve( words * npass );
        for( int p = 0; p < npass; p++ ) fill_profile_wide( ctx().h_profile_wide[g_stage_slot][slot].p + p * words, q, qlen, plp, p );
        ctx().d_profile_wide[slot].reserve( words * npass );
        CUDA_OK( cudaMemcpyAsync( ctx().d_profile_wide[slot].p, ctx().h_profile_wide[g_stage_slot][slot].p, words * npass * 4, cudaMemcpyHostToDevice, stream ) );
        ctx().stats.h2d_bytes += words * npass * 4;
        if( !nw ) {
            zero_slots_kernel<<<(unsigned) ( ( n + 255 ) / 256 ), 256, 0, stream>>>( ctx().d_wide_batches[slot].p, (uint32_t) n, ctx().d_scores.p );
            ctx().stats.launches++;
        }
        launch_wide_pipe_tier( slot, (uint32_t) n, nullptr, bnd, qlen, npass, nw, stream );
        ctx().stats.wide_tier_seqs += n;
        return;
    }
    const int npass = pl.wide_npass, R = pl.wide_R, G = pl.wide_G;
    const size_t words = (size_t) R * 1024;
    ctx().h_profile_wide[g_stage_slot][slot].reserve( words * npass );
    for( int p = 0; p < npass; p++ ) fill_profile_wide( ctx().h_profile_wide[g_stage_slot][slot].p + p * words, q, qlen, pl, p );
    ctx().d_profile_wide[slot].reserve( words * npass );
    CUDA_OK( cudaMemcpyAsync( ctx().d_profile_wide[slot].p, ctx().h_profile_wide[g_stage_slot][slot].p, words * npass * 4, cudaMemcpyHostToDevice, stream ) );
    ctx().stats.h2d_bytes += words * npass * 4;
    if( npass > 1 ) ctx().d_bnd_wide.reserve( 4 * bnd + 256 );   // 2 buffers x ( bnd / 2 + 1 ) entries x 4 words
    ctx().d_counter.reserve( 64 );
    for( int p = 0; p < npass; p++ ) {
        CUDA_OK( cudaMemsetAsync( ( ctx().d_counter.p + 16 + 8 * slot ), 0, 4, stream ) );
        ScoreArgs a;
        memset( &a, 0, sizeof( a ) );
        a.stream = db.d_stream.p;
        a.batches = ctx().d_wide_batches[slot].p;
        a.nbatches = (uint32_t) n;
        a.counter = ( ctx().d_counter.p + 16 + 8 * slot );
        a.profile = ctx().d_profile_wide[slot].p + p * words;
        a.group = (uint32_t) G;
        a.k65536 = 65536u;
        a.qr_word = (uint32_t) ( (int) g_gapO + (int) g_gapE );
        a.rg_word = (uint32_t) (int) g_gapE;
        a.gap_open_extend = (int) g_gapO + (int) g_gapE;
        a.gap_extend = g_gapE;
        a.gap_open = g_gapO;
        a.offset = 0;
        a.rowbase = p * G * R;
        a.cap_lane = -1;
        a.cap_row = 0;
        if( nw && (size_t) a.rowbase <= qlen - 1 && qlen - 1 < (size_t) a.rowbase + (size_t) G * R ) {
            a.cap_lane = (int) ( ( qlen - 1 - a.rowbase ) / R );
            a.cap_row = (int) ( ( qlen - 1 - a.rowbase ) % R );
        }
        a.first_pass = ( p == 0 );
        a.out = ctx().d_scores.p;
        a.first8 = d_first8;
        a.first16 = d_first16;
        a.thr8 = 255;
        a.thr16 = 65535;
        if( npass > 1 ) {   // two buffers of one 16-byte entry per column pair, written and read alternately
            uint4 * b0 = (uint4 *) ctx().d_bnd_wide.p, * b1 = b0 + ( bnd / 2 + 1 );
            a.bnd_in = p > 0 ? ( ( p & 1 ) ? b0 : b1 ) : nullptr;
            a.bnd_out = p + 1 < npass ? ( ( p & 1 ) ? b1 : b0 ) : nullptr;
        }
        launch_score_wide( R, nw, npass > 1, track, a, stream );
        ctx().stats.launches++;
    }
    ctx().stats.wide_tier_seqs += n;
}

void upload_matrix32() {
    if( ctx().matrix32_version == g_matrix.version ) return;
    int32_t m32[1024];
    for( int i = 0; i < 1024; i++ ) m32[i] = (int32_t) g_matrix.m[i];
    ctx().d_matrix32.reserve( 1024 );
    CUDA_OK( cudaMemcpy( ctx().d_matrix32.p, m32, sizeof( m32 ), cudaMemcpyHostToDevice ) );
    ctx().matrix32_version = g_matrix.version;
}

// The 32-bit tier over the sequences the speculative packed tier listed on the DEVICE (unpack_scores_kernel): batches,
// profile and batch count never visit the host, so the common case - nothing listed - costs a few empty launches and no
// synchronisation.  `cap` = sequences that can be listed at all (those the static bound does not clear), cap_cols their
// stream columns.
void run_wide_spec( const uint8_t * q, size_t qlen, const Plan & pl, uint32_t cap, uint64_t cap_cols, cudaStream_t stream ) {
    PackedDb & db = ctx().db;
    const bool pipe = wide_pipe_applies( qlen );
    const Plan plw = pipe ? wide_pipe_plan( pl, qlen ) : pl;
    const int npass = plw.wide_npass, R = plw.wide_R, G = plw.wide_G;
    const size_t words = (size_t) R * 1024;
    upload_matrix32();
    ctx().d_query.reserve( qlen + 1 );
    CUDA_OK( cudaMemcpyAsync( ctx().d_query.p, q, qlen, cudaMemcpyHostToDevice, stream ) );
    ctx().stats.h2d_bytes += qlen;
    ctx().d_profile_wide[0].reserve( words * npass );
    fill_profile_wide_kernel<<<(unsigned) ( ( words * npass + 255 ) / 256 ), 256, 0, stream>>>( ctx().d_query.p, (uint32_t) qlen,
            ctx().d_matrix32.p, G, R, npass, pl.pad_score, ctx().d_profile_wide[0].p );
    ctx().d_wide_batches[0].reserve( cap );
    uint32_t * d_nb = ctx().d_counter.p + 48;
    build_spec_batches_kernel<<<1, 32, 0, stream>>>( ctx().d_spec_list.p, ctx().d_counter.p + 44, cap, db.d_pair_col.p, db.d_sorted_len.p,
            ctx().d_wide_batches[0].p, d_nb, pipe ? ctx().d_scores.p : nullptr );
    ctx().stats.launches += 2;
    if( pipe ) launch_wide_pipe_tier( 0, cap, d_nb, cap_cols, qlen, npass, false, stream );
    if( npass > 1 ) ctx().d_bnd_wide.reserve( 4 * cap_cols + 256 );
    for( int p = 0; p < npass && !pipe; p++ ) {
        CUDA_OK( cudaMemsetAsync( ctx().d_counter.p + 16, 0, 4, stream ) );
        ScoreArgs a;
        memset( &a, 0, sizeof( a ) );
        a.stream = db.d_stream.p;
        a.batches = ctx().d_wide_batches[0].p;
        a.nbatches = cap;               // sizes the grid; the kernel reads the real count
        a.nbatches_dev = d_nb;
        a.counter = ctx().d_counter.p + 16;
        a.profile = ctx().d_profile_wide[0].p + p * words;
        a.group = (uint32_t) G;
        a.k65536 = 65536u;
        a.qr_word = (uint32_t) ( (int) g_gapO + (int) g_gapE );
        a.rg_word = (uint32_t) (int) g_gapE;
        a.gap_open_extend = (int) g_gapO + (int) g_gapE;
        a.gap_extend = g_gapE;
        a.gap_open = g_gapO;
        a.rowbase = p * G * R;
        a.cap_lane = -1;
        a.first_pass = ( p == 0 );
        a.out = ctx().d_scores.p;
        a.thr8 = 255;
        a.thr16 = 65535;
        if( npass > 1 ) {
            uint4 * b0 = (uint4 *) ctx().d_bnd_wide.p, * b1 = b0 + ( cap_cols / 2 + 1 );
            a.bnd_in = p > 0 ? ( ( p & 1 ) ? b0 : b1 ) : nullptr;
            a.bnd_out = p + 1 < npass ? ( ( p & 1 ) ? b1 : b0 ) : nullptr;
        }
        launch_score_wide( R, false, npass > 1, false, a, stream );
        ctx().stats.launches++;
    }
    // how many were recomputed: read back with the call's results (valid after the stream is synchronised)
    ctx().h_spec_count[g_stage_slot].reserve( 1 );
    CUDA_OK( cudaMemcpyAsync( ctx().h_spec_count[g_stage_slot].p, ctx().d_counter.p + 44, 4, cudaMemcpyDeviceToHost, stream ) );
    ctx().spec_pending[g_stage_slot] = true;
    ctx().stats.d2h_bytes += 4;
}

// One pass of the packed tier: G lanes x R rows from query row `rowbase` on, over its own list of work batches (the boundary
// buffers are indexed by stream column, so the passes of one query may cut the stream into different batches).
struct PassLaunch {
    int G = 32, R = 4;
    size_t rowbase = 0;
    const Batch * d_batches = nullptr;
    uint32_t nbatches = 0;
};

// LIBSSA_B200_NO_SW32=1: full-warp arrays through the generic kernel as well; LIBSSA_B200_NO_NW32=1: only Needleman-Wunsch
// (A/B measurements)
bool sw32_enabled( bool nw ) {
    static const bool on = getenv( "LIBSSA_B200_NO_SW32" ) == nullptr, on_nw = getenv( "LIBSSA_B200_NO_NW32" ) == nullptr;
    return on && ( !nw || on_nw );
}
// whether groups of G lanes run on the chained kernel (sw32_kernel.cuh): full warps, and half warps for one-pass queries
// (LIBSSA_B200_NO_G16=1: half warps on the generic kernel, A/B measurements)
// Measured [B200], 375-aa query, half-warp chains against the generic kernel: Smith-Waterman +1.6 % (570 k records) / +4 % (71 k
// records); Needleman-Wunsch -4 % while its kernel carried a one-column path (128 registers), +1.1 % / +2.2 % since the query's
// last row is the last row of a lane (nw_row_shift); two queries per word +4 %.
bool chained_shape( int G, bool nw, bool one_pass, bool dual = false ) {
    static const bool g16 = getenv( "LIBSSA_B200_NO_G16" ) == nullptr;
    (void) dual;
    return sw32_enabled( nw ) && ( G == 32 || ( G == 16 && one_pass && g16 ) );
}

// Packed NW: virtual rows on top of a query of qlen rows whose last pass starts at array position rowbase_last and has R rows per
// lane - as many as it takes for the query's last row to be the last row of a lane (fill_profile_packed)
size_t nw_row_shift( size_t qlen, size_t rowbase_last, int R ) { return ( (size_t) R - ( qlen - rowbase_last ) % (size_t) R ) % (size_t) R; }

// The packed tier over one class of pairs: all passes of the query, one launch each.
void launch_packed_class( cudaStream_t stream, const uint8_t * q, size_t qlen, bool nw, const Plan & pl, const std::vector<PassLaunch> & passes,
        uint32_t * d_counter, DevBuf<uint32_t> & d_profile_buf, PinBuf<uint32_t> & h_profile_buf, int32_t offset ) {
    PackedDb & db = ctx().db;
    const int npass = (int) passes.size();
    const bool multi = npass > 1 || getenv( "LIBSSA_B200_FORCE_MULTI" ) != nullptr;
    size_t words_total = 0;
    for( const PassLaunch & ps : passes ) words_total += packed_profile_words( ps.R );
    h_profile_buf.reserve( words_total );
    d_profile_buf.reserve( words_total );
    uint32_t * const h_profile = h_profile_buf.p, * const d_profile = d_profile_buf.p;
    size_t w0 = 0;
    size_t shift = 0;
    if( nw )   // from the pass that holds the query's last row (the last one, unless a forced shape added an empty pass)
        for( const PassLaunch & ps : passes )
            if( ps.rowbase <= qlen - 1 && qlen - 1 < ps.rowbase + (size_t) ps.G * ps.R ) shift = nw_row_shift( qlen, ps.rowbase, ps.R );
    const size_t cap_pos = qlen + shift - 1;   // array position of the query's last row: last row of a lane of the last pass
    for( const PassLaunch & ps : passes ) {
        fill_profile_packed( h_profile + w0, q, qlen, ps.G, ps.R, ps.rowbase, pl.pad_score, shift );
        w0 += packed_profile_words( ps.R );
    }
    CUDA_OK( cudaMemcpyAsync( d_profile, h_profile, words_total * 4, cudaMemcpyHostToDevice, stream ) );
    ctx().stats.h2d_bytes += words_total * 4;
    w0 = 0;
    for( int p = 0; p < npass; p++ ) {
        const PassLaunch & ps = passes[p];
        if( chained_shape( ps.G, nw, npass == 1 ) ) {
            // Full-warp arrays: the specialised kernel (sw32_kernel.cuh).  Consecutive passes of one shape
            // over the same batches - the full passes of a long query - are ONE launch: the kernel chains them itself, a batch
            // of pass p+1 waits only for the same batch of pass p.  (LIBSSA_B200_PASS_PER_LAUNCH=1: one launch per pass.)
            int run = 1;
            static const bool per_launch = getenv( "LIBSSA_B200_PASS_PER_LAUNCH" ) != nullptr;
            while( !per_launch && multi && p + run < npass && run < SW32_MAX_PASSES && passes[p + run].G == 32 && passes[p + run].R == ps.R &&
                    passes[p + run].d_batches == ps.d_batches && passes[p + run].nbatches == ps.nbatches )
                run++;
            // work queue heads: one per pass of the launch (the class' own counter when there is only one)
            uint32_t * counters = d_counter;
            if( run > 1 ) {
                ctx().d_pass_state.reserve( SW32_MAX_PASSES + ps.nbatches + 2 );
                counters = ctx().d_pass_state.p;
                CUDA_OK( cudaMemsetAsync( counters, 0, ( SW32_MAX_PASSES + (size_t) ps.nbatches ) * 4, stream ) );
            }
            else {
                CUDA_OK( cudaMemsetAsync( d_counter, 0, 4, stream ) );
            }
            Sw32Args s32;
            memset( &s32, 0, sizeof( s32 ) );
            s32.stream32 = (const uint32_t *) db.d_stream.p;
            s32.batches = ps.d_batches;
            s32.nbatches = ps.nbatches;
            s32.counters = counters;
            s32.profile = d_profile + w0;
            s32.k65536 = 65536u;
            s32.qr_word = (uint32_t) (uint16_t) (int16_t) ( (int) g_gapO + (int) g_gapE ) * 0x10001u;
            s32.rg_word = (uint32_t) (uint16_t) (int16_t) (int) g_gapE * 0x10001u;
            s32.npass = (uint32_t) run;
            s32.pass0 = (uint32_t) p;
            s32.more_passes = p + run < npass ? 1u : 0u;
            s32.out_by_col = ctx().d_out_packed.p;
            if( npass > 1 ) {
                s32.bnd0 = (uint4 *) ctx().d_bnd.p;
                s32.bnd1 = s32.bnd0 + ( db.total_cols / 2 + 1 );
            }
            s32.done = run > 1 ? counters + SW32_MAX_PASSES : nullptr;
            s32.cap_pass = -1;
            if( nw ) {
                s32.gap_open = g_gapO;
                s32.gap_extend = g_gapE;
                s32.offset = offset;
                s32.rowbase = (int32_t) ps.rowbase;
                s32.nw_shift = s32.nw_shift2 = (int32_t) shift;
                for( int c = p; c < p + run; c++ )   // the pass (of this launch) that holds the query's last row
                    if( passes[c].rowbase <= cap_pos && cap_pos < passes[c].rowbase + (size_t) ps.G * ps.R ) {
                        s32.cap_pass = c;
                        s32.cap_lane = (int) ( ( cap_pos - passes[c].rowbase ) / ps.R );
                        s32.cap_row = (int) ( ( cap_pos - passes[c].rowbase ) % ps.R );   // = R - 1
                    }
            }
            launch_sw32( ps.R, nw, multi && ps.G == 32, s32, stream, ps.G );
            ctx().stats.launches++;
            w0 += (size_t) run * packed_profile_words( ps.R );
            p += run - 1;
            continue;
        }
        CUDA_OK( cudaMemsetAsync( d_counter, 0, 4, stream ) );
        ScoreArgs a;
        memset( &a, 0, sizeof( a ) );
        a.stream = db.d_stream.p;
        a.batches = ps.d_batches;
        a.nbatches = ps.nbatches;
        a.counter = d_counter;
        a.profile = d_profile + w0;
        w0 += packed_profile_words( ps.R );
        a.group = (uint32_t) ps.G;
        a.k65536 = 65536u;
        a.qr_word = (uint32_t) (uint16_t) (int16_t) ( (int) g_gapO + (int) g_gapE ) * 0x10001u;
        a.rg_word = (uint32_t) (uint16_t) (int16_t) (int) g_gapE * 0x10001u;
        a.gap_open_extend = (int) g_gapO + (int) g_gapE;
        a.gap_extend = g_gapE;
        a.gap_open = g_gapO;
        a.offset = offset;
        a.rowbase = (int32_t) ps.rowbase;
        a.cap_lane = -1;
        a.nw_shift = a.nw_shift2 = (int32_t) shift;
        if( nw && ps.rowbase <= cap_pos && cap_pos < ps.rowbase + (size_t) ps.G * ps.R ) {
            a.cap_lane = (int) ( ( cap_pos - ps.rowbase ) / ps.R );
            a.cap_row = (int) ( ( cap_pos - ps.rowbase ) % ps.R );   // = R - 1
        }
        a.first_pass = ( p == 0 );
        a.out = ctx().d_out_packed.p;
        if( npass > 1 ) {   // two buffers of one 16-byte entry per stream column pair: pass p reads [(p - 1) & 1], writes [p & 1]
            uint4 * b0 = (uint4 *) ctx().d_bnd.p, * b1 = b0 + ( db.total_cols / 2 + 1 );
            a.bnd_in = p > 0 ? ( ( p & 1 ) ? b0 : b1 ) : nullptr;
            a.bnd_out = p + 1 < npass ? ( ( p & 1 ) ? b1 : b0 ) : nullptr;
        }
        launch_score_packed( ps.R, nw, multi, a, stream );
        ctx().stats.launches++;
    }
}

// Passes of a query longer than one 32 x 32 pass.  Uniform plan (pick_shape): npass passes of 32 x R.  Mixed plan: full
// passes of 32 x 32 rows - the shape with the least per-row cost - and a remainder pass whose shape (any G) fits the rows
// that are left, so that at most a few rows are padding (5147 rows: 5 x 1024 + 4 x 8 instead of 6 x 896).  Narrow groups
// process a pair from its first to its last column on G lanes: the longest pair must not outlast the rest of the pass.
// Costs in the units of pick_shape (instructions per stream column and warp).
// full_R: rows per lane of the full passes (32; LIBSSA_B200_NW_FULL_R sets another value for Needleman-Wunsch, A/B measurements)
std::vector<PassLaunch> plan_passes( size_t qlen, int R_uniform, int npass_uniform, uint64_t total_cols, uint64_t longest_cols, int sms, int full_R = 32 ) {
    std::vector<PassLaunch> v;
    const double uniform_cost = npass_uniform * ( 7.5 * R_uniform + 60 ) * 1.03;
    const size_t full_rows = (size_t) 32 * full_R;
    const size_t full = qlen / full_rows, rem = qlen % full_rows;
    double best = 1e300;
    int bg = 0, br = 0;
    if( rem > 0 )
        for( int G = 4; G <= 32; G *= 2 )
            for( int R : PACKED_R ) {
                if( (size_t) G * R < rem ) continue;
                const double step = 7.5 * R + 75;
                // columns per scheduler if the pass were perfectly spread vs the one group that owns the longest pair
                const double spread = (double) total_cols * G / 32.0, serial = (double) longest_cols * sms * 4 * 2.7;
                const double cost = step * std::max( spread, serial ) / (double) std::max<uint64_t>( total_cols, 1 );
                if( cost < best ) {
                    best = cost;
                    bg = G;
                    br = R;
                }
            }
    const double mixed_cost = full * ( 7.5 * full_R + 60 ) * 1.03 + ( rem > 0 ? best : 0.0 );
    if( getenv( "LIBSSA_B200_SHAPE" ) || getenv( "LIBSSA_B200_UNIFORM_PASSES" ) || full == 0 || mixed_cost >= uniform_cost ) {
        for( int p = 0; p < npass_uniform; p++ ) {
            PassLaunch ps;
            ps.G = 32;
            ps.R = R_uniform;
            ps.rowbase = (size_t) p * 32 * R_uniform;
            v.push_back( ps );
        }
        return v;
    }
    for( size_t p = 0; p < full; p++ ) {
        PassLaunch ps;
        ps.G = 32;
        ps.R = full_R;
        ps.rowbase = p * full_rows;
        v.push_back( ps );
    }
    if( rem > 0 ) {
        PassLaunch ps;
        ps.G = bg;
        ps.R = br;
        ps.rowbase = full * full_rows;
        v.push_back( ps );
    }
    return v;
}

// first pair (pairs are sorted by descending length of their longer sequence) whose length is <= len
size_t first_pair_not_longer( int64_t len ) {
    PackedDb & db = ctx().db;
    size_t lo = 0, hi = db.npairs;
    while( lo < hi ) {
        size_t mid = ( lo + hi ) / 2;
        if( (int64_t) db.sorted_len[2 * mid] > len )
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

void score_shard( const uint8_t * q, size_t qlen, bool nw ) {
    PackedDb & db = ctx().db;
    const size_t nloc = db.sorted_id.size();
    ctx().d_scores.reserve( ( db.sorted_all.size() + ctx().world - 1 ) / ctx().world + 2 );
    if( nloc == 0 ) return;
    if( qlen == 0 ) fatal( "libssa_b200: empty query." );
    if( qlen > ( 1u << 24 ) || ( qlen + db.sorted_len[0] ) > ( 1ull << 24 ) )
        fatal( "libssa_b200: query/database sequences beyond 2^24 symbols are not supported by the 32-bit tier." );
    Plan pl = make_plan( qlen );
    int sms = 148;
    cudaDeviceGetAttribute( &sms, cudaDevAttrMultiProcessorCount, ctx().device );
    if( db.batch_tables.size() >= 24 ) {  // bound the cache: a table's first pair depends on the query and the score matrix
        CUDA_OK( cudaDeviceSynchronize() );
        for( auto & kv : db.batch_tables ) kv.second.d_batches.release();
        db.batch_tables.clear();
    }

    // ---- tier split: pairs whose longer sequence exceeds the packed tier's guaranteed range go to the wide tier ----
    int32_t offset = 0;
    const int qrows_max = std::max( pl.qpad, 32 * 32 * ( ( (int) qlen + 1023 ) / 1024 ) );
    int64_t safe = packed_safe_len( pl, nw, qrows_max, &offset );
    const bool positive_gaps = (int) g_gapO + (int) g_gapE > 0 || g_gapE > 0;
    if( positive_gaps ) safe = 0;  // positive "penalties": only the wide tier is valid
    const size_t p_static = first_pair_not_longer( safe );   // pairs [0, p_static) are not provably inside int16
    // Speculative packed tier (Smith-Waterman): run those pairs in int16 all the same and recompute, in the 32-bit tier, only
    // the sequences whose running maximum came within one substitution score of the limit.  A wrap needs a diagonal
    // predecessor H > 32767 - hi+, every H is bounded by a diagonal candidate the running maximum has seen, and the maximum
    // never decreases - so a final maximum below the threshold proves that nothing wrapped.  E and F stay above
    // 2 (gapO + gapE).  (LIBSSA_B200_STATIC_TIERS=1: the static split, for A/B measurements.)
    // (The static split is not the cheaper choice even for a handful of long records: with all full passes of the bulk in ONE
    // persistent launch the 32-bit tier's launches on the auxiliary stream find no free multiprocessor slot after the first
    // and run behind the bulk - measured +10 ms on the 570 k-record database for 34 records [B200].
    // LIBSSA_B200_STATIC_TIERS=1 forces it for A/B measurements and tests.)
    const int64_t hip = std::max<int64_t>( pl.hi, 0 );
    const bool speculative = !nw && !positive_gaps && p_static > 0 && hip <= 1024 && pl.lo >= -1024 && !getenv( "LIBSSA_B200_STATIC_TIERS" );
    const int32_t spec_thr = (int32_t) ( 32767 - hip );
    const size_t p_safe = speculative ? 0 : p_static;

    // ---- shape of the packed tier ----
    // The bulk of the database runs on groups of G lanes x R rows (few padded rows, per-step work amortised over many
    // rows).  A pair is processed by ONE group from its first to its last column, so with narrow groups the longest
    // pairs would outlast everything else; pairs longer than `long_len` columns therefore run on full-warp arrays
    // (32 x long_R) in a second launch that overlaps the first on another stream.  Cost model in instructions:
    // a warp-step costs 7.5 R + 60 and advances 32/G groups by one column.
    int G = pl.G, R = pl.R, npass = pl.npass, long_R = 0;
    size_t p_long_end = p_safe;
    if( npass == 1 && qlen <= 1024 && !getenv( "LIBSSA_B200_SHAPE" ) ) {
        const int * opts = PACKED_R;
        const int nopt = sizeof( PACKED_R ) / sizeof( int );
        int r32 = opts[nopt - 1];
        for( int i = nopt - 1; i >= 0; i-- )
            if( (size_t) 32 * opts[i] >= qlen ) r32 = opts[i];
        double best = 1e300;
        for( int g = 4; g <= 32; g *= 2 ) {
            int r = 0;
            for( int i = nopt - 1; i >= 0; i-- )
                if( (size_t) g * opts[i] >= qlen ) r = opts[i];
            if( r == 0 ) continue;
            // pairs longer than this many columns would dominate the tail of a launch with groups of g lanes
            static const double long_div = getenv( "LIBSSA_B200_LONG_DIV" ) ? atof( getenv( "LIBSSA_B200_LONG_DIV" ) ) : 1165.0;
            const int64_t long_len = g == 32 ? INT64_MAX : std::max<int64_t>( 64, (int64_t) ( (double) db.total_cols * g / ( long_div * sms ) ) );
            const size_t pl_end = std::max( p_safe, first_pair_not_longer( long_len ) );
            const double cols_long = (double) ( db.pair_col[pl_end] - db.pair_col[p_safe] );
            const double cols_bulk = (double) ( db.total_cols - db.pair_col[pl_end] );
            // per-step work around the recurrence: ~90 instructions on the chained kernel (sw32_kernel.cuh), ~140 on the generic one
            const double frame = chained_shape( g, nw, true ) ? 90 : 140, frame32 = sw32_enabled( nw ) ? 90 : 140;
            const double cost = cols_bulk * ( g / 32.0 ) * ( 7.5 * r + frame ) + cols_long * ( 7.5 * r32 + frame32 );
            if( cost < best ) {
                best = cost;
                G = g;
                R = r;
                p_long_end = pl_end;
                long_R = g == 32 ? 0 : r32;
            }
        }
    }

    // batch size of the bulk class: at least ~3 batches per resident group of lanes, 256..4096 columns (powers of two, cached)
    // Smith-Waterman on full-warp arrays runs on sw32_kernel: its warps chain their batches (no wavefront fill per batch), so
    // the batches can be small - at least ~8 per warp, down to 128 columns at the end of the queue
    const bool chained = chained_shape( G, nw, npass == 1 );
    const int resident = chained ? sw32_blocks_per_sm( R, nw, npass > 1, G ) : packed_blocks_per_sm( R, nw, npass > 1 );
    const uint64_t groups = (uint64_t) sms * resident * ( SCORE_THREADS / 32 ) * ( 32 / G );
    uint32_t batch_cols = chained ? 2048 : 4096;
    // (with the shrinking tail of batch_table the choice is flat between 2 and 12 batches per group [B200]; fewer, longer batches
    // mean fewer wavefront fills on small databases)
    while( batch_cols > 256 && db.total_cols / batch_cols < ( chained ? 8 : 3 ) * groups ) batch_cols /= 2;
    PackedDb::BatchTable & bt = batch_table( batch_cols, 0, (uint32_t) groups, chained ? 128u : (uint32_t) std::max( 128, 16 * G ) );
    const size_t b0 = std::lower_bound( bt.pair0.begin(), bt.pair0.end(), (uint32_t) p_long_end ) - bt.pair0.begin();
    const size_t bulk_pair0 = b0 < bt.batches.size() ? bt.pair0[b0] : db.npairs;   // bulk class = pairs [bulk_pair0, npairs)
    size_t long_pair0 = p_safe;                                                  // long class = pairs [long_pair0, bulk_pair0)
    if( long_R == 0 ) long_pair0 = bulk_pair0;                                  // no long class: those pairs go to the wide tier
    const size_t wide_pairs = long_pair0;                                        // wide tier = pairs [0, long_pair0)
    if( nw && wide_pairs < db.npairs ) offset = nw_offset_for( pl, qrows_max, db.sorted_len[2 * wide_pairs] );

    CUDA_OK( cudaEventRecord( ctx().ev_k0, ctx().stream ) );
    ctx().scored = true;
    ctx().d_out_packed.reserve( db.total_cols / 2 + 64 );   // results by stream column pair (sparse: one word per pair of sequences)
    ctx().d_counter.reserve( 64 );
    // ---- auxiliary stream: the few sequences of the wide tier and the long class overlap the bulk launch ----
    const size_t n_long = bulk_pair0 - long_pair0;
    const bool use_aux = n_long > 0 || wide_pairs > 0;
    if( use_aux ) {
        CUDA_OK( cudaEventRecord( ctx().ev_fork, ctx().stream ) );
        CUDA_OK( cudaStreamWaitEvent( ctx().stream_aux, ctx().ev_fork, 0 ) );
    }
    if( wide_pairs > 0 ) {  // pairs beyond the packed tier's guaranteed range: exact int32 tier
        std::vector<uint32_t> slots;
        for( size_t s = 0; s < std::min( nloc, 2 * wide_pairs ); s++ ) slots.push_back( (uint32_t) s );
        run_wide( slots, q, qlen, pl, nw, false, nullptr, nullptr, 0, ctx().stream_aux );
    }
    if( n_long > 0 ) {
        ctx().h_long_batches[g_stage_slot].reserve( n_long );
        for( size_t i = 0; i < n_long; i++ ) {
            Batch b;
            memset( &b, 0, sizeof( b ) );
            const size_t p = long_pair0 + i;
            b.col0 = db.pair_col[p];
            b.bnd0 = db.pair_col[p];
            b.ncols = (uint32_t) ( db.pair_col[p + 1] - db.pair_col[p] );
            b.out0 = (uint32_t) p;
            ctx().h_long_batches[g_stage_slot].p[i] = b;
        }
        ctx().d_long_batches.reserve( n_long );
        CUDA_OK( cudaMemcpyAsync( ctx().d_long_batches.p, ctx().h_long_batches[g_stage_slot].p, n_long * sizeof( Batch ), cudaMemcpyHostToDevice, ctx().stream_aux ) );
        ctx().stats.h2d_bytes += n_long * sizeof( Batch );
        PassLaunch lp;
        lp.G = 32;
        lp.R = long_R;
        lp.d_batches = ctx().d_long_batches.p;
        lp.nbatches = (uint32_t) n_long;
        launch_packed_class( ctx().stream_aux, q, qlen, nw, pl, { lp }, ctx().d_counter.p + 8, ctx().d_profile_long, ctx().h_profile_long[g_stage_slot], offset );
    }
    if( use_aux ) CUDA_OK( cudaEventRecord( ctx().ev_join, ctx().stream_aux ) );
    // ---- bulk class ----
    if( b0 < bt.batches.size() ) {
        std::vector<PassLaunch> passes;
        if( npass > 1 ) {
            // (measured [B200]: NW full passes of 28 rows per lane +0.9 % for the 5147-aa protein query, -3.6 % for the 10 kbp
            // nucleotide query - the default stays 32)
            static const int nw_full_R = getenv( "LIBSSA_B200_NW_FULL_R" ) ? atoi( getenv( "LIBSSA_B200_NW_FULL_R" ) ) : 32;
            passes = plan_passes( qlen, R, npass, db.total_cols - db.pair_col[bulk_pair0], db.sorted_len[2 * bulk_pair0], sms,
                    nw && sw32_enabled( true ) ? nw_full_R : 32 );
            ctx().d_bnd.reserve( 4 * db.total_cols + 256 );   // 2 buffers x ( total_cols / 2 + 1 ) entries x 4 words
        }
        else {
            PassLaunch ps;
            ps.G = G;
            ps.R = R;
            passes.push_back( ps );
        }
        for( PassLaunch & ps : passes ) {
            if( ps.G == G && ps.R == R ) {  // the shape the bulk batch table was sized for
                ps.d_batches = bt.d_batches.p + b0;
                ps.nbatches = (uint32_t) ( bt.batches.size() - b0 );
                continue;
            }
            // another shape has another number of resident groups: its own batches over the same pairs
            const bool ps_chained = chained_shape( ps.G, nw, false );
            const int ps_resident = ps_chained ? sw32_blocks_per_sm( ps.R, nw, true ) : packed_blocks_per_sm( ps.R, nw, true );
            const uint64_t ps_groups = (uint64_t) sms * ps_resident * ( SCORE_THREADS / 32 ) * ( 32 / ps.G );
            uint32_t ps_cols = ps_chained ? 2048 : 4096;
            while( ps_cols > 256 && db.total_cols / ps_cols < ( ps_chained ? 8 : 6 ) * ps_groups ) ps_cols /= 2;
            PackedDb::BatchTable & pt = batch_table( ps_cols, (uint32_t) bulk_pair0, (uint32_t) ps_groups,
                    ps_chained ? 128u : (uint32_t) std::max( 128, 16 * ps.G ) );
            ps.d_batches = pt.d_batches.p;
            ps.nbatches = (uint32_t) pt.batches.size();
        }
        launch_packed_class( ctx().stream, q, qlen, nw, pl, passes, ctx().d_counter.p, ctx().d_profile, ctx().h_profile[g_stage_slot], offset );
    }
    if( use_aux ) CUDA_OK( cudaStreamWaitEvent( ctx().stream, ctx().ev_join, 0 ) );
    if( wide_pairs < db.npairs ) {
        const uint32_t np = (uint32_t) ( db.npairs - wide_pairs );
        const uint32_t spec_cap = speculative ? (uint32_t) std::min( nloc, 2 * p_static ) : 0;
        if( speculative ) {
            ctx().d_spec_list.reserve( spec_cap + 2 );
            CUDA_OK( cudaMemsetAsync( ctx().d_counter.p + 44, 0, 4, ctx().stream ) );
        }
        unpack_scores_kernel<<<( np + 255 ) / 256, 256, 0, ctx().stream>>>( ctx().d_out_packed.p, db.d_pair_col.p, ctx().d_scores.p, (uint32_t) wide_pairs,
                (uint32_t) db.npairs, (uint32_t) nloc, offset, speculative ? (uint32_t) p_static : 0u, spec_thr, ctx().d_spec_list.p,
                ctx().d_counter.p + 44 );
        ctx().stats.launches++;
        if( speculative ) {
            uint64_t cap_cols = 0;   // an upper bound for any spec_cap sequences: the longest ones
            for( uint32_t i = 0; i < spec_cap; i++ ) cap_cols += ( (uint64_t) db.sorted_len[i] + 1 ) & ~1ull;
            run_wide_spec( q, qlen, pl, spec_cap, cap_cols, ctx().stream );
        }
    }
    CUDA_OK( cudaEventRecord( ctx().ev_k1, ctx().stream ) );
}

// Two queries per packed word (score_kernel<..., DUAL>): both queries of a batch call against the whole shard, one pass, packed
// tier only - every sequence is scored by two launches (sequence A resp. B of every pair), each word holding one cell of
// query 1 and of query 2.  Applies when both queries fit one pass, every pair is statically inside int16 for the longer of
// them and no pair is long enough to need the full-warp class next to narrow groups (dual_plan; otherwise the caller
// scores the queries one after the other).  Scores by sorted position: ctx().d_scores (query 1), ctx().d_scores2 (query 2).
struct DualPlan {
    int G = 0, R = 0;
    int32_t offset = 0;
    int pad_score = 0;
    int qrows_max = 0;
    Plan pl;
};

// whether two queries of these lengths can share a scoring pass, and its shape
bool dual_plan( size_t len1, size_t len2, bool nw, DualPlan & dp ) {
    PackedDb & db = ctx().db;
    const size_t nloc = db.sorted_id.size();
    const size_t lmax = std::max( len1, len2 );
    if( getenv( "LIBSSA_B200_NO_DUAL" ) || nloc == 0 || len1 == 0 || len2 == 0 || lmax > 1024 || ctx().world != 1 ) return false;
    if( (int) g_gapO + (int) g_gapE > 0 || g_gapE > 0 ) return false;
    dp.pl = make_plan( lmax );
    if( dp.pl.npass != 1 ) return false;
    dp.qrows_max = std::max( dp.pl.qpad, 1024 );
    const int64_t safe = packed_safe_len( dp.pl, nw, dp.qrows_max, &dp.offset );
    if( first_pair_not_longer( safe ) != 0 ) return false;   // some pair is not provably inside int16: single-query path (speculation / 32-bit tier)
    int sms = 148;
    cudaDeviceGetAttribute( &sms, cudaDevAttrMultiProcessorCount, ctx().device );
    // shape: G lanes x R rows with few padded rows; narrow groups only if no pair would outlast the launch on them
    const int nopt = sizeof( PACKED_R ) / sizeof( int );
    double best = 1e300;
    dp.G = 0;
    for( int g = 4; g <= 32; g *= 2 ) {
        int r = 0;
        for( int i = nopt - 1; i >= 0; i-- )
            if( (size_t) g * PACKED_R[i] >= lmax ) r = PACKED_R[i];
        if( r == 0 ) continue;
        const int64_t long_len = g == 32 ? INT64_MAX : std::max<int64_t>( 64, (int64_t) ( (double) db.total_cols * g / ( 1165.0 * sms ) ) );
        if( first_pair_not_longer( long_len ) != 0 ) continue;
        double cost = ( g / 32.0 ) * ( 5.75 * r + 60 );
        if( r > 24 ) cost *= 1.6;   // 112-128 KB of profile: one resident CTA per multiprocessor
        if( cost < best ) {
            best = cost;
            dp.G = g;
            dp.R = r;
        }
    }
    if( dp.G == 0 ) return false;
    dp.pad_score = dp.pl.pad_score;
    if( nw ) dp.offset = nw_offset_for( dp.pl, dp.qrows_max, db.sorted_len[0] );
    return true;
}

void score_shard_dual( const uint8_t * q1, size_t len1, const uint8_t * q2, size_t len2, bool nw, const DualPlan & dp ) {
    PackedDb & db = ctx().db;
    const size_t nloc = db.sorted_id.size();
    const int G = dp.G, R = dp.R;
    const int32_t offset = dp.offset;
    const Plan & pl = dp.pl;
    int sms = 148;
    cudaDeviceGetAttribute( &sms, cudaDevAttrMultiProcessorCount, ctx().device );
    if( db.batch_tables.size() >= 24 ) {
        CUDA_OK( cudaDeviceSynchronize() );
        for( auto & kv : db.batch_tables ) kv.second.d_batches.release();
        db.batch_tables.clear();
    }
    // arrays of 16 / 32 lanes run on the chained kernel (small batches, at least ~8 per array), narrower ones on the generic kernel
    const bool chained = chained_shape( G, nw, true, true ) && !getenv( "LIBSSA_B200_NO_DUAL_CHAINED" );
    const int resident = chained ? sw32_blocks_per_sm( R, nw, false, G, true ) : dual_blocks_per_sm( R, nw );
    const uint64_t groups = (uint64_t) sms * resident * ( SCORE_THREADS / 32 ) * ( 32 / G );
    uint32_t batch_cols = chained ? 2048 : 4096;
    while( batch_cols > 256 && db.total_cols / batch_cols < ( chained ? 8 : 3 ) * groups ) batch_cols /= 2;
    PackedDb::BatchTable & bt = batch_table( batch_cols, 0, (uint32_t) groups, chained ? 128u : (uint32_t) std::max( 128, 16 * G ) );

    const size_t words = (size_t) R * 1024;
    ctx().h_profile_dual[g_stage_slot].reserve( words );
    ctx().d_profile_dual.reserve( words );
    const size_t shift1 = nw ? nw_row_shift( len1, 0, R ) : 0, shift2 = nw ? nw_row_shift( len2, 0, R ) : 0;
    fill_profile_dual( ctx().h_profile_dual[g_stage_slot].p, q1, len1, q2, len2, G, R, pl.pad_score, shift1, shift2 );
    CUDA_OK( cudaMemcpyAsync( ctx().d_profile_dual.p, ctx().h_profile_dual[g_stage_slot].p, words * 4, cudaMemcpyHostToDevice, ctx().stream ) );
    ctx().stats.h2d_bytes += words * 4;
    const uint32_t out_stride = (uint32_t) ( db.total_cols / 2 + 64 );
    ctx().d_out_packed.reserve( 2 * (size_t) out_stride );
    ctx().d_counter.reserve( 64 );
    ctx().d_scores.reserve( nloc + 2 );
    ctx().d_scores2.reserve( nloc + 2 );
    CUDA_OK( cudaEventRecord( ctx().ev_k0, ctx().stream ) );
    ctx().scored = true;
    for( uint32_t half = 0; half < 2; half++ ) {
        if( half == 1 && nloc < 2 ) break;
        CUDA_OK( cudaMemsetAsync( ctx().d_counter.p, 0, 4, ctx().stream ) );
        if( chained ) {
            Sw32Args s32;
            memset( &s32, 0, sizeof( s32 ) );
            s32.stream32 = (const uint32_t *) db.d_stream.p;
            s32.batches = bt.d_batches.p;
            s32.nbatches = (uint32_t) bt.batches.size();
            s32.counters = ctx().d_counter.p;
            s32.profile = ctx().d_profile_dual.p;
            s32.k65536 = 65536u;
            s32.qr_word = (uint32_t) (uint16_t) (int16_t) ( (int) g_gapO + (int) g_gapE ) * 0x10001u;
            s32.rg_word = (uint32_t) (uint16_t) (int16_t) (int) g_gapE * 0x10001u;
            s32.npass = 1;
            s32.out_by_col = ctx().d_out_packed.p;
            s32.gap_open = g_gapO;
            s32.gap_extend = g_gapE;
            s32.offset = offset;
            s32.cap_pass = nw ? 0 : -1;
            s32.cap_lane = s32.cap_lane2 = -1;
            if( nw ) {
                s32.cap_lane = (int) ( ( len1 + shift1 - 1 ) / R );
                s32.cap_lane2 = (int) ( ( len2 + shift2 - 1 ) / R );
                s32.cap_row = s32.cap_row2 = R - 1;
                s32.nw_shift = (int32_t) shift1;
                s32.nw_shift2 = (int32_t) shift2;
            }
            s32.dual_half = half;
            s32.out_stride = out_stride;
            launch_sw32( R, nw, false, s32, ctx().stream, G, true );
            ctx().stats.launches++;
            continue;
        }
        ScoreArgs a;
        memset( &a, 0, sizeof( a ) );
        a.stream = db.d_stream.p;
        a.batches = bt.d_batches.p;
        a.nbatches = (uint32_t) bt.batches.size();
        a.counter = ctx().d_counter.p;
        a.profile = ctx().d_profile_dual.p;
        a.group = (uint32_t) G;
        a.k65536 = 65536u;
        a.qr_word = (uint32_t) (uint16_t) (int16_t) ( (int) g_gapO + (int) g_gapE ) * 0x10001u;
        a.rg_word = (uint32_t) (uint16_t) (int16_t) (int) g_gapE * 0x10001u;
        a.gap_open_extend = (int) g_gapO + (int) g_gapE;
        a.gap_extend = g_gapE;
        a.gap_open = g_gapO;
        a.offset = offset;
        a.rowbase = 0;
        a.cap_lane = a.cap_lane2 = -1;
        if( nw ) {
            a.cap_lane = (int) ( ( len1 + shift1 - 1 ) / R );
            a.cap_lane2 = (int) ( ( len2 + shift2 - 1 ) / R );
            a.cap_row = a.cap_row2 = R - 1;
            a.nw_shift = (int32_t) shift1;
            a.nw_shift2 = (int32_t) shift2;
        }
        a.first_pass = 1;
        a.out = ctx().d_out_packed.p;
        a.dual_half = half;
        a.out_stride = out_stride;
        launch_score_dual( R, nw, a, ctx().stream );
        ctx().stats.launches++;
    }
    unpack_dual_kernel<<<(unsigned) ( ( db.npairs + 255 ) / 256 ), 256, 0, ctx().stream>>>( ctx().d_out_packed.p, out_stride, db.d_pair_col.p,
            db.d_sorted_len.p, ctx().d_scores.p, ctx().d_scores2.p, (uint32_t) db.npairs, (uint32_t) nloc, offset );
    ctx().stats.launches++;
    CUDA_OK( cudaEventRecord( ctx().ev_k1, ctx().stream ) );
}

// ---- order emulation of the reference's narrow NW tiers (nw_sat.cuh) ----

// Longest database sequence for which the saturating `bits`-wide NW kernel of the reference provably never
// saturates nor trips its overflow test (all of H - padded columns included -, E, F and the injected boundaries
// stay strictly inside the tier's range), so that no emulation is needed.  -1 = none.
int64_t nw_tier_safe_len( size_t qlen, int bits ) {
    const int64_t LO = bits == 8 ? -128 : -32768, HI = bits == 8 ? 127 : 32767;
    const int64_t gO = g_gapO, gE = g_gapE, QR = gO + gE;
    if( gO > 0 || gE > 0 || QR <= LO + 2 ) return -1;
    for( int i = 0; i < 1024; i++ )
        if( g_matrix.m[i] < LO || g_matrix.m[i] > HI ) return -1;   // the typed copies of the matrix would wrap
    const int64_t hip = std::max<int64_t>( g_matrix.hi, 0 ), lom = std::min<int64_t>( g_matrix.lo, 0 );
    auto safe = [&]( int64_t d ) {
        const int64_t d4 = ( d + 3 ) / 4 * 4;
        const int64_t hi = hip * ( std::min<int64_t>( (int64_t) qlen, d4 ) + 1 );
        const int64_t lo = 2 * gO + ( (int64_t) qlen + d4 + 2 ) * gE + lom + QR;
        return hi < HI - 1 && lo > LO - QR + 2;
    };
    if( !safe( 1 ) ) return -1;
    int64_t a = 1, b = 1 << 26;
    while( a + 1 < b ) {
        int64_t mid = ( a + b ) / 2;
        if( safe( mid ) )
            a = mid;
        else
            b = mid;
    }
    return a;
}

// number of leading local sorted positions (longest sequences first) whose length exceeds `len`
size_t sorted_prefix_longer_than( int64_t len ) {
    PackedDb & db = ctx().db;
    size_t lo = 0, hi = db.sorted_len.size();
    while( lo < hi ) {
        size_t mid = ( lo + hi ) / 2;
        if( (int64_t) db.sorted_len[mid] > len )
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

// emulates the `bits`-wide saturating NW kernel for the first `count` local sorted positions; first_col[s] receives
// 4 * (overflow block) for the records that overflow
void run_nw_sat( const uint8_t * q, size_t qlen, int bits, size_t count, int32_t * d_first_col, int32_t * d_sat_score = nullptr ) {
    PackedDb & db = ctx().db;
    if( count == 0 ) return;
    upload_matrix32();
    ctx().d_query.reserve( qlen + 1 );
    CUDA_OK( cudaMemcpyAsync( ctx().d_query.p, q, qlen, cudaMemcpyHostToDevice, ctx().stream ) );
    ctx().stats.h2d_bytes += qlen;
    const size_t wave = std::max<size_t>( 4096, ( 256u << 20 ) / ( 4 * qlen ) );
    ctx().d_nw_scratch.reserve( 2 * qlen * std::min( wave, count ) + 16 );
    ctx().d_wave_list.reserve( std::min( wave, count ) + 1 );
    for( size_t first = 0; first < count; first += wave ) {
        const size_t n = std::min( wave, count - first );
        iota_kernel<<<(unsigned) ( ( n + 255 ) / 256 ), 256, 0, ctx().stream>>>( ctx().d_wave_list.p, (uint32_t) first, (uint32_t) n );
        NwSatArgs a;
        a.stream = db.d_stream.p;
        a.pair_col = db.d_pair_col.p;
        a.sorted_len = db.d_sorted_len.p;
        a.list = ctx().d_wave_list.p;
        a.count = (uint32_t) n;
        a.query = ctx().d_query.p;
        a.qlen = (uint32_t) qlen;
        a.matrix = ctx().d_matrix32.p;
        a.gap_open = g_gapO;
        a.gap_extend = g_gapE;
        a.bits = bits;
        a.scratch_h = ctx().d_nw_scratch.p;
        a.scratch_e = ctx().d_nw_scratch.p + qlen * n;
        a.first_col = d_first_col;
        a.sat_score = d_sat_score;
        nw_sat_kernel<<<(unsigned) ( ( n + 127 ) / 128 ), 128, 0, ctx().stream>>>( a );
        ctx().stats.launches += 2;
    }
}

// where each chunk of `chunk` provider records starts in arrival order (the reference stops at the first chunk
// without a non-empty record, src/algo/16/search_16.c:123)
PackedDb::ChunkInfo & chunk_info_for( size_t chunk ) {
    PackedDb & db = ctx().db;
    auto it = db.chunk_info.find( chunk );
    if( it != db.chunk_info.end() ) return it->second;
    PackedDb::ChunkInfo & ci = db.chunk_info[chunk];
    std::vector<uint32_t> base;
    size_t pos = 0;
    const size_t n = db.lens_all.size();
    for( size_t start = 0; start < n; start += chunk ) {
        size_t cnt = 0;
        for( size_t i = start; i < std::min( n, start + chunk ); i++ ) cnt += db.lens_all[i] > 0;
        if( cnt == 0 ) break;
        base.push_back( (uint32_t) pos );
        pos += cnt;
    }
    ci.nchunks = (uint32_t) base.size();
    ci.narr = pos;
    ci.d_chunk_base.reserve( base.size() + 1 );
    if( !base.empty() ) CUDA_OK( cudaMemcpy( ci.d_chunk_base.p, base.data(), base.size() * 4, cudaMemcpyHostToDevice ) );
    return ci;
}

// Runs the scheduler emulation on the GPU (arrival.cuh): fills d_pos (global sorted position -> arrival position,
// -1 = never reached) and d_order / h_order (arrival position -> provider index).  first_all = gathered overflow
// columns of the tracking kernel or nullptr.  Returns the number of records in arrival order.
size_t gpu_arrival_order( int bit_width, size_t chunk, const int32_t * d_first_all, size_t stride, DevBuf<int32_t> & d_pos,
        uint32_t * h_order ) {
    PackedDb & db = ctx().db;
    PackedDb::ChunkInfo & ci = chunk_info_for( chunk );
    const size_t nglob = db.sorted_all.size();
    d_pos.reserve( nglob + 1 );
    ctx().d_order.reserve( ci.narr + 1 );
    CUDA_OK( cudaMemsetAsync( d_pos.p, 0xff, ( nglob + 1 ) * sizeof( int32_t ), ctx().stream ) );
    if( ci.nchunks > 0 ) {
        ArrivalArgs a;
        a.lens = db.d_lens_all.p;
        a.n_all = (uint32_t) db.lens_all.size();
        a.chunk = (uint32_t) chunk;
        a.nchunks = ci.nchunks;
        a.chunk_base = ci.d_chunk_base.p;
        a.bit_width = bit_width;
        a.sorted_pos_of_id = db.d_sorted_pos_of_id.p;
        a.first_all = d_first_all;
        a.stride = (uint32_t) stride;
        a.world = (uint32_t) ctx().world;
        a.order = ctx().d_order.p;
        a.pos_of_sorted = d_pos.p;
        arrival_kernel<<<ci.nchunks, 32, 0, ctx().stream>>>( a );
        ctx().stats.launches++;
        if( h_order ) {  // nullptr: the caller fetches the part of ctx().d_order it needs
            CUDA_OK( cudaMemcpyAsync( h_order, ctx().d_order.p, ci.narr * sizeof( uint32_t ), cudaMemcpyDeviceToHost, ctx().stream ) );
            ctx().stats.d2h_bytes += ci.narr * sizeof( uint32_t );
        }
    }
    return ci.narr;
}

PackedDb::Arrival & arrival_for( int bit_width, size_t chunk, bool on_gpu );
void upload_positions( const std::vector<uint32_t> & order, DevBuf<int32_t> & d_pos ) {
    PackedDb & db = ctx().db;
    const size_t nglob = db.sorted_all.size();
    ctx().h_pos.reserve( nglob + 1 );
    int32_t * pos = ctx().h_pos.p;
    if( order.size() != nglob ) std::fill( pos, pos + nglob, -1 );  // records the reference never reaches
    for( size_t a = 0; a < order.size(); a++ ) pos[db.sorted_pos_of_id[order[a]]] = (int32_t) a;
    d_pos.reserve( nglob + 1 );
    CUDA_OK( cudaMemcpyAsync( d_pos.p, pos, nglob * sizeof( int32_t ), cudaMemcpyHostToDevice, ctx().stream ) );
    CUDA_OK( cudaStreamSynchronize( ctx().stream ) );
    ctx().stats.h2d_bytes += nglob * sizeof( int32_t );
}

PackedDb::Arrival & arrival_for( int bit_width, size_t chunk, bool on_gpu ) {
    PackedDb & db = ctx().db;
    auto key = std::make_pair( bit_width, chunk );
    auto it = db.arrival.find( key );
    if( it != db.arrival.end() ) return it->second;
    PackedDb::Arrival & ar = db.arrival[key];
    if( on_gpu ) {
        ctx().h_order.reserve( db.sorted_all.size() + 1 );
        const size_t n = gpu_arrival_order( bit_width, chunk, nullptr, 1, ar.d_pos_of_sorted, ctx().h_order.p );
        CUDA_OK( cudaStreamSynchronize( ctx().stream ) );
        ar.order.assign( ctx().h_order.p, ctx().h_order.p + n );
    }
    else {
        arrival_order( db.lens_all, nullptr, nullptr, chunk, bit_width, ar.order );
        upload_positions( ar.order, ar.d_pos_of_sorted );
    }
    ar.uploaded = true;
    return ar;
}

// all-gather `count` int32 per rank from d_send into d_recv (rank-major) on the calling context's stream; world = 1: plain copy.
// The collective runs INSIDE the library: ncclAllGather on the communicator of ssa_b200_nccl_init (one process per GPU) or
// of ncclCommInitAll (one process, several GPUs: every context's host thread calls it for its own rank).  A host-supplied
// callback (ssa_b200_set_allgather) is the alternative for hosts that bring their own collective.
void gather_i32( const int32_t * d_send, int32_t * d_recv, size_t count ) {
    if( ctx().world == 1 ) {
        if( d_send != d_recv )
            CUDA_OK( cudaMemcpyAsync( d_recv, d_send, count * sizeof( int32_t ), cudaMemcpyDeviceToDevice, ctx().stream ) );
        return;
    }
    if( ctx().comm ) {
        NCCL_OK( g_nccl.AllGather( d_send, d_recv, count, ncclInt32, ctx().comm, ctx().stream ) );
        return;
    }
    if( !g_cfg.allgather )
        fatal( "libssa_b200: the database is sharded over %d processes but no collective is available "
               "(ssa_b200_nccl_init or ssa_b200_set_allgather).", ctx().world );
    g_cfg.allgather( g_cfg.allgather_ctx, d_send, d_recv, count * sizeof( int32_t ), (void *) ctx().stream );
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// entry points used by api.cpp
// ------------------------------------------------------------------------------------------------

void gpu_prepare_db() { build_db(); }

static void release_ctx( Ctx & c ) {
    cudaSetDevice( c.device );
    cudaDeviceSynchronize();
    PackedDb & db = c.db;
    db.d_stream.release();
    db.d_lens_all.release();
    db.d_pair_col.release();
    db.d_sorted_all.release();
    db.d_sorted_len.release();
    db.d_sorted_pos_of_id.release();
    for( auto & kv : db.chunk_info ) kv.second.d_chunk_base.release();
    db.chunk_info.clear();
    for( auto & kv : db.batch_tables ) kv.second.d_batches.release();
    db.batch_tables.clear();
    for( auto & kv : db.arrival ) kv.second.d_pos_of_sorted.release();
    db.arrival.clear();
    db.valid = false;
    c.d_profile.release();
    c.d_profile_long.release();
    c.d_out_packed.release();
    c.d_counter.release();
    c.d_bnd.release();
    c.d_bnd_wide.release();
    c.d_long_batches.release();
    c.d_first_compact.release();
    c.d_scores.release();
    c.d_gathered.release();
    c.d_arrival_scores.release();
    c.d_first.release();
    c.d_pos_tmp.release();
    c.d_order.release();
    c.d_wave_list.release();
    c.d_topk_ids.release();
    c.d_topk_scores.release();
    c.d_select.release();
    c.d_nw_scratch.release();
    c.d_query.release();
    c.d_matrix32.release();
    c.matrix32_version = ~0ull;
    c.d_flag_list.release();
    c.d_cand_pos.release();
    c.d_cand_id.release();
    c.d_cand_score.release();
    c.d_spec_list.release();
    c.d_nw_tier.release();
    c.d_wide_progress[0].release();
    c.d_wide_progress[1].release();
    c.d_pass_state.release();
    c.h_spec_count[0].release();
    c.h_spec_count[1].release();
    c.spec_pending[0] = c.spec_pending[1] = false;
    c.h_order.release();
    c.h_list.release();
    c.h_scores.release();
    c.h_first.release();
    c.h_pos.release();
    c.h_cand_pos.release();
    c.h_cand_id.release();
    c.h_cand_score.release();
    c.d_scores2.release();
    c.d_profile_dual.release();
    for( int i = 0; i < 4; i++ ) {
        c.d_all_slot[i].release();
        c.h_select[i].release();
        c.h_topk_ids[i].release();
        c.h_topk_scores[i].release();
        c.h_counts[i].release();
        if( c.ev_topk[i] ) cudaEventDestroy( c.ev_topk[i] );
        c.ev_topk[i] = nullptr;
    }
    for( int i = 0; i < 2; i++ ) {
        c.h_profile_dual[i].release();
        c.d_wide_batches[i].release();
        c.d_profile_wide[i].release();
        for( int j = 0; j < 2; j++ ) {
            c.h_wide_batches[i][j].release();
            c.h_profile_wide[i][j].release();
        }
        c.h_long_batches[i].release();
        c.h_profile_long[i].release();
        c.h_profile[i].release();
    }
    cudaEventDestroy( c.ev_begin );
    cudaEventDestroy( c.ev_k0 );
    cudaEventDestroy( c.ev_k1 );
    cudaEventDestroy( c.ev_end );
    cudaEventDestroy( c.ev_fork );
    cudaEventDestroy( c.ev_join );
    cudaStreamDestroy( c.stream_aux );
    cudaStreamDestroy( c.stream );
}

void gpu_release() {
    if( g_devs.empty() ) return;
    if( !g_workers.threads.empty() ) {
        {
            std::lock_guard<std::mutex> lk( g_workers.m );
            g_workers.stop = true;
        }
        g_workers.cv_go.notify_all();
        for( std::thread & th : g_workers.threads ) th.join();
        g_workers.threads.clear();
        g_workers.stop = false;
    }
    for( Ctx * c : g_devs ) {
        cudaSetDevice( c->device );
        cudaDeviceSynchronize();
    }
    for( Ctx * c : g_devs )
        if( c->comm ) {
            g_nccl.CommDestroy( c->comm );
            c->comm = nullptr;
        }
    cudaSetDevice( g_devs[0]->device );
    traceback_release();
    for( Ctx * c : g_devs ) {
        release_ctx( *c );
        delete c;
    }
    g_devs.clear();
    t_ctx = nullptr;
    g_gdb.release();
}

// per-call counters of every context -> zero; the primary's record is what ssa_b200_get_stats returns
static void reset_call_stats() {
    for( Ctx * c : g_devs ) {
        c->stats.launches = 0;
        c->stats.h2d_bytes = c->stats.d2h_bytes = 0;
        c->stats.wide_tier_seqs = 0;
        c->stats.overflow_8 = c->stats.overflow_16 = 0;
        c->scored = false;
    }
}

// sums the other contexts' counters into the primary's record; score kernel time = the slowest device's
static void merge_call_stats() {
    ssa_b200_stats_t & st = g_devs[0]->stats;
    float kernel_ms = 0;
    for( Ctx * c : g_devs ) {
        for( int slot = 0; slot < 2; slot++ )   // sequences the speculative packed tier sent to the 32-bit tier (copied back with the results)
            if( c->spec_pending[slot] ) {
                c->stats.wide_tier_seqs += *c->h_spec_count[slot].p;
                c->spec_pending[slot] = false;
            }
        if( c != g_devs[0] ) {
            st.launches += c->stats.launches;
            st.h2d_bytes += c->stats.h2d_bytes;
            st.d2h_bytes += c->stats.d2h_bytes;
            st.wide_tier_seqs += c->stats.wide_tier_seqs;
            c->stats.launches = 0;
            c->stats.h2d_bytes = c->stats.d2h_bytes = 0;
            c->stats.wide_tier_seqs = 0;
        }
        if( c->scored ) {
            float ms = 0;
            if( cudaEventElapsedTime( &ms, c->ev_k0, c->ev_k1 ) == cudaSuccess ) kernel_ms = std::max( kernel_ms, ms );
        }
    }
    cudaGetLastError();   // an event of another device may not have completed yet: not an error
    st.score_kernel_ms = kernel_ms;
}

// Exact scores of every record this PROCESS holds (all its contexts' shards), provider order.
size_t gpu_score_all( const uint8_t * query, size_t qlen, int search_type, int64_t * scores, uint64_t * ids ) {
    build_db();
    reset_call_stats();
    CUDA_OK( cudaEventRecord( ctx().ev_begin, ctx().stream ) );
    run_on_all( [&]() {
        PackedDb & db = ctx().db;
        const size_t nloc = db.sorted_id.size();
        score_shard( query, qlen, search_type == NEEDLEMAN_WUNSCH );
        ctx().h_scores.reserve( nloc + 1 );
        if( nloc ) CUDA_OK( cudaMemcpyAsync( ctx().h_scores.p, ctx().d_scores.p, nloc * sizeof( int32_t ), cudaMemcpyDeviceToHost, ctx().stream ) );
        if( is_primary() ) CUDA_OK( cudaEventRecord( ctx().ev_end, ctx().stream ) );
        CUDA_OK( cudaStreamSynchronize( ctx().stream ) );
        ctx().stats.d2h_bytes += nloc * sizeof( int32_t );
    } );
    merge_call_stats();
    float ms = 0;
    cudaEventElapsedTime( &ms, ctx().ev_begin, ctx().ev_end );
    ctx().stats.device_ms = ms;
    ctx().stats.cells = (uint64_t) qlen * process_residues();
    std::vector<std::pair<uint32_t, int32_t>> all;   // (provider index, score)
    for( Ctx * c : g_devs )
        for( size_t i = 0; i < c->db.sorted_id.size(); i++ ) all.emplace_back( c->db.sorted_id[i], c->h_scores.p[i] );
    std::sort( all.begin(), all.end() );
    for( size_t i = 0; i < all.size(); i++ ) {
        if( scores ) scores[i] = all[i].second;
        if( ids ) ids[i] = all[i].first;
    }
    return all.size();
}

// ---- top-k on the device ----
// Radix-select the k-th largest score T and count the ties at T (topk.cuh).  If no more than k records score >= T the
// reported set is {score >= T} whatever the arrival order, so the k hits are compacted on the device and nothing else
// crosses PCIe.  Only when the tie at T straddles the cut-off does the reference's tie behaviour matter; then the
// arrival-order replay of search_single runs.  Split in two so that a batch of queries can keep the GPU busy with the
// next query while the host picks up the results of this one.
static bool topk_applicable( PackedDb::Arrival & base, size_t k ) { return k > 0 && k < base.order.size() && !ctx().db.sorted_all.empty(); }

static void enqueue_topk( const int32_t * d_all, PackedDb::Arrival & base, size_t k, bool count_overflow, int slot ) {
    PackedDb & db = ctx().db;
    const size_t nglob = db.sorted_all.size();
    const size_t world = (size_t) ctx().world;
    const size_t stride = ( nglob + world - 1 ) / world;
    if( ctx().d_select.cap == 0 ) {   // the passes keep the state's histogram, ticket and accumulators zero from here on
        ctx().d_select.reserve( 1 );
        CUDA_OK( cudaMemsetAsync( ctx().d_select.p, 0, sizeof( SelectState ), ctx().stream ) );
    }
    ctx().h_select[slot].reserve( 1 );
    SelectArgs sa;
    sa.scores = d_all;
    sa.pos = base.d_pos_of_sorted.p;
    sa.stride = (uint32_t) stride;
    sa.world = (uint32_t) world;
    sa.nglobal = (uint32_t) nglob;
    sa.state = ctx().d_select.p;
    const uint32_t total = (uint32_t) ( stride * world );
    const unsigned grid = std::min<unsigned>( ( total + 255 ) / 256, 148 * 8 );
    const int shifts[3] = { 21, 10, 0 };
    const uint32_t masks[3] = { 0x7ff, 0x7ff, 0x3ff };
    for( int p = 0; p < 3; p++ ) select_pass_kernel<<<grid, 256, 0, ctx().stream>>>( sa, p, shifts[p], masks[p], (uint32_t) k, p == 2 );
    ctx().d_topk_scores.reserve( k + 1 );
    ctx().d_topk_ids.reserve( k + 1 );
    select_compact_kernel<<<grid, 256, 0, ctx().stream>>>( sa, db.d_sorted_all.p, ctx().d_topk_scores.p, ctx().d_topk_ids.p, (uint32_t) k );
    ctx().stats.launches += 4;
    ctx().h_topk_scores[slot].reserve( k + 1 );
    ctx().h_topk_ids[slot].reserve( k + 1 );
    ctx().h_counts[slot].reserve( 2 );
    ctx().h_counts[slot].p[0] = count_overflow ? 1 : 0;   // whether the reference's overflow counters apply (manager.c:157-160)
    CUDA_OK( cudaMemcpyAsync( ctx().h_select[slot].p, ctx().d_select.p, SELECT_HEADER_BYTES, cudaMemcpyDeviceToHost, ctx().stream ) );
    CUDA_OK( cudaMemcpyAsync( ctx().h_topk_scores[slot].p, ctx().d_topk_scores.p, k * sizeof( int32_t ), cudaMemcpyDeviceToHost, ctx().stream ) );
    CUDA_OK( cudaMemcpyAsync( ctx().h_topk_ids[slot].p, ctx().d_topk_ids.p, k * sizeof( uint32_t ), cudaMemcpyDeviceToHost, ctx().stream ) );
    CUDA_OK( cudaEventRecord( ctx().ev_end, ctx().stream ) );
    CUDA_OK( cudaEventRecord( ctx().ev_topk[slot], ctx().stream ) );
    ctx().stats.d2h_bytes += SELECT_HEADER_BYTES + k * 8;
}

// waits for the slot's results; true (and the sorted hit list) if the set does not depend on the arrival order
static bool collect_topk( size_t k, int bit_width, int slot, std::vector<Hit> & hits ) {
    CUDA_OK( cudaEventSynchronize( ctx().ev_topk[slot] ) );
    const SelectState & st = *ctx().h_select[slot].p;
    hits.clear();
    if( (size_t) st.count_gt + st.count_eq > k ) return false;
    const size_t n = (size_t) st.count_gt + st.count_eq;
    hits.resize( n );
    for( size_t i = 0; i < n; i++ ) hits[i] = Hit{ ctx().h_topk_ids[slot].p[i], ctx().h_topk_scores[slot].p[i] };
    std::sort( hits.begin(), hits.end(), []( const Hit & x, const Hit & y ) {
        if( x.score != y.score ) return x.score > y.score;
        return x.db_id > y.db_id;
    } );
    const bool counted = ctx().h_counts[slot].p[0] != 0;
    if( bit_width == 8 ) ctx().stats.overflow_8 = counted ? st.ge8 : 0;
    if( bit_width == 8 || bit_width == 16 ) ctx().stats.overflow_16 = counted ? st.ge16 : 0;
    return true;
}

// One query frame against the entry stream: top-k with the reference's one-thread tie behaviour.  Hit ids are entry
// (slot) indices.  Every key (score, id) is distinct, so the final order does not depend on the sort algorithm.
// scores of this shard's entries -> every rank; returns the gathered score vector (rank-major blocks of `stride`)
// whether the reference's saturating NW tiers can report something else than the exact score for some record of this database
// The emulation runs one thread per record from its first column to the block in which the tier flags it (or to its end), so
// its critical path is query length x longest record: bounded here to 5 M cells (~0.1 s).  Beyond that the tiers are taken
// to flag every record the static bounds cannot clear, i.e. the exact scores stand - a 10 kbp query against 20 kbp records
// (BASELINE configs[4]) would otherwise spend 9 s per search on a case in which a tier can only accept a record whose
// last columns match row after row behind the point where the boundary saturates.  Deviation recorded in DESIGN.md 2.
static bool nw_narrow_scores_needed( size_t qlen, bool nw, int bit_width ) {
    PackedDb & db = ctx().db;
    if( !nw || db.sorted_all.empty() || ( bit_width != 8 && bit_width != 16 ) ) return false;
    const int64_t longest = (int64_t) db.lens_all[db.sorted_all[0]];
    if( (double) qlen * (double) longest > 5.0e6 && !getenv( "LIBSSA_B200_NW_TIER_SCORES" ) ) return false;
    return ( bit_width == 8 && longest > nw_tier_safe_len( qlen, 8 ) ) || longest > nw_tier_safe_len( qlen, 16 );
}

static int32_t * score_query( const uint8_t * query, size_t qlen, bool nw, int bit_width ) {
    const size_t nglob = ctx().db.sorted_all.size();
    const size_t world = (size_t) ctx().world;
    const size_t stride = ( nglob + world - 1 ) / world;  // local positions per rank, padded
    const bool narrow = nw_narrow_scores_needed( qlen, nw, bit_width );   // the same on every rank (global longest record)
    run_on_all( [&]() {
        ctx().d_scores.reserve( stride + 2 );
        score_shard( query, qlen, nw );
        if( narrow ) {   // the scores the reference's 8- / 16-bit NW tiers report for the records they accept (nw_sat.cuh)
            const size_t n8 = bit_width == 8 ? sorted_prefix_longer_than( nw_tier_safe_len( qlen, 8 ) ) : 0;
            const size_t n16 = sorted_prefix_longer_than( nw_tier_safe_len( qlen, 16 ) );
            ctx().d_nw_tier.reserve( 4 * stride + 4 );   // first8 | first16 | score8 | score16
            int32_t * d_f = ctx().d_nw_tier.p, * d_s = d_f + 2 * stride;
            launch_fill( d_f, INT_MAX, 2 * stride );
            if( n16 ) {
                run_nw_sat( query, qlen, 16, n16, d_f + stride, d_s + stride );
                nw_apply_sat_kernel<<<(unsigned) ( ( n16 + 255 ) / 256 ), 256, 0, ctx().stream>>>( d_f + stride, d_s + stride, (uint32_t) n16, ctx().d_scores.p );
            }
            if( n8 ) {   // (after the 16-bit tier: a record the 8-bit tier accepts never reaches it)
                run_nw_sat( query, qlen, 8, n8, d_f, d_s );
                nw_apply_sat_kernel<<<(unsigned) ( ( n8 + 255 ) / 256 ), 256, 0, ctx().stream>>>( d_f, d_s, (uint32_t) n8, ctx().d_scores.p );
            }
            ctx().stats.launches += 2;
        }
        // ---- shard results -> every rank (ncclAllGather over NVLink / NVSwitch, on this context's stream) ----
        if( world > 1 ) {
            ctx().d_gathered.reserve( stride * world + 2 );
            gather_i32( ctx().d_scores.p, ctx().d_gathered.p, stride );
        }
    } );
    return world == 1 ? ctx().d_scores.p : ctx().d_gathered.p;
}

static void select_replay( const int32_t * d_all, const uint8_t * query, size_t qlen, const SearchRequest & req, std::vector<Hit> & hits );

// sparse heap replay (replay_topk_sparse): the first m arrivals are replayed in full
static size_t sparse_replay_prefix( size_t hitcount ) {
    size_t floor = 65536;
    if( const char * e = getenv( "LIBSSA_B200_REPLAY_PREFIX" ) ) floor = std::max<size_t>( 1, (size_t) atol( e ) );   // tests: small databases too
    return std::max<size_t>( floor, 64 * hitcount );
}
static bool sparse_replay_applies( size_t narr, size_t hitcount ) {
    return narr > 4 * sparse_replay_prefix( hitcount ) && !getenv( "LIBSSA_B200_FULL_REPLAY" );
}

static void search_single( const uint8_t * query, size_t qlen, const SearchRequest & req, std::vector<Hit> & hits ) {
    PackedDb & db = ctx().db;
    const bool nw = req.search_type == NEEDLEMAN_WUNSCH;
    ctx().stats.cells = (uint64_t) qlen * process_residues();
    const int32_t * d_all = score_query( query, qlen, nw, req.bit_width );
    // ---- top-k on the device (see enqueue_topk) ----
    const size_t chunk = g_chunk_size * (size_t) db.buffer_max;  // entries per reference chunk (db_adapter.c:196)
    PackedDb::Arrival & base = arrival_for( req.bit_width, chunk, chunk <= (size_t) ARRIVAL_MAX_CHUNK );  // validity: records the reference reaches
    hits.clear();
    if( topk_applicable( base, req.hitcount ) ) {
        enqueue_topk( d_all, base, req.hitcount, !nw && ( req.bit_width == 8 || req.bit_width == 16 ), 0 );
        if( collect_topk( req.hitcount, req.bit_width, 0, hits ) ) return;
    }
    select_replay( d_all, query, qlen, req, hits );
}

// The reference's tie behaviour in full: which records overflow a narrow tier and where, the heap-arrival order of the
// emulated one-thread run, and the heap replay (used when ties straddle the cut-off of the hit list).
static void select_replay( const int32_t * d_all, const uint8_t * query, size_t qlen, const SearchRequest & req, std::vector<Hit> & hits ) {
    PackedDb & db = ctx().db;
    const size_t nglob = db.sorted_all.size();
    const size_t world = (size_t) ctx().world;
    const size_t stride = ( nglob + world - 1 ) / world;
    const bool nw = req.search_type == NEEDLEMAN_WUNSCH;
    const size_t chunk = g_chunk_size * (size_t) db.buffer_max;
    const bool use_gpu_arrival = chunk <= (size_t) ARRIVAL_MAX_CHUNK;
    hits.clear();

    // ---- which records would the reference have pushed through its overflow cascade? (SW only) ----
    // A record overflows the 8-bit tier when its score reaches 255 and the 16-bit tier at 65535
    // (reference src/algo/simd/search_simd_sw.c:372-379,422-423); it then leaves its SIMD channel after the 4-column
    // block in which the running maximum got there, which changes the order in which scores reach the heap.
    std::vector<int64_t> ovf8, ovf16;
    bool have_overflow = false;
    const int32_t * d_first_gathered = nullptr;
    uint32_t n_flagged = 0;
    if( !nw && nglob > 0 && ( req.bit_width == 8 || req.bit_width == 16 ) ) {
        const int32_t thr = req.bit_width == 8 ? 255 : 65535;
        const uint32_t total = (uint32_t) ( stride * world );
        ctx().d_flag_list.reserve( total + 2 );
        uint32_t * d_count = ctx().d_flag_list.p + total;
        CUDA_OK( cudaMemsetAsync( d_count, 0, 4, ctx().stream ) );
        flag_overflow_kernel<<<( total + 255 ) / 256, 256, 0, ctx().stream>>>( d_all, (uint32_t) stride, (uint32_t) world,
                (uint32_t) nglob, thr, ctx().d_flag_list.p, d_count, total );
        ctx().stats.launches++;
        ctx().h_list.reserve( total + 2 );
        CUDA_OK( cudaMemcpyAsync( ctx().h_list.p + total, d_count, 4, cudaMemcpyDeviceToHost, ctx().stream ) );
        CUDA_OK( cudaStreamSynchronize( ctx().stream ) );
        const uint32_t count = ctx().h_list.p[total];
        ctx().stats.d2h_bytes += 4;
        if( count > 0 ) {  // every rank sees the same gathered scores, so every rank takes this branch together
            have_overflow = true;
            CUDA_OK( cudaMemcpyAsync( ctx().h_list.p, ctx().d_flag_list.p, count * 4, cudaMemcpyDeviceToHost, ctx().stream ) );
            CUDA_OK( cudaStreamSynchronize( ctx().stream ) );
            ctx().stats.d2h_bytes += count * 4;
            std::vector<uint32_t> flagged( ctx().h_list.p, ctx().h_list.p + count );  // gathered indices r*stride + s
            std::sort( flagged.begin(), flagged.end() );
            // first columns at which the running maximum reaches 255 / 65535, via the tracking wide kernel: every context
            // re-runs its own flagged records, the columns are all-gathered like the scores
            const Plan pl = make_plan( qlen );
            run_on_all( [&]() {
                std::vector<uint32_t> slots;  // this rank's own flagged local positions
                for( uint32_t i : flagged )
                    if( i / stride == (uint32_t) ctx().rank ) slots.push_back( (uint32_t) ( i % stride ) );
                ctx().d_first.reserve( 2 * stride * ( world + 1 ) + 4 );
                int32_t * d_f = ctx().d_first.p;  // [stride] first8 | [stride] first16
                launch_fill( d_f, INT_MAX, 2 * stride );
                run_wide( slots, query, qlen, pl, false, true, d_f, d_f + stride, 1, ctx().stream );  // rewrites the same exact scores
                if( world > 1 ) gather_i32( d_f, ctx().d_first.p + 2 * stride, 2 * stride );
            } );
            int32_t * d_f_all = world > 1 ? ctx().d_first.p + 2 * stride : ctx().d_first.p;
            d_first_gathered = d_f_all;
            // overflow counters (the reference prints them at OUTPUT_INFO) and, for the host fallback, the columns
            ctx().d_first_compact.reserve( 2 * (size_t) count + 2 );
            gather_first_kernel<<<( count + 255 ) / 256, 256, 0, ctx().stream>>>( ctx().d_flag_list.p, count, d_f_all, (uint32_t) stride,
                    ctx().d_first_compact.p );
            ctx().stats.launches++;
            ctx().h_first.reserve( 2 * (size_t) count + 2 );
            CUDA_OK( cudaMemcpyAsync( ctx().h_first.p, ctx().d_first_compact.p, 2 * (size_t) count * sizeof( int32_t ), cudaMemcpyDeviceToHost, ctx().stream ) );
            if( !use_gpu_arrival ) {
                CUDA_OK( cudaStreamSynchronize( ctx().stream ) );
                ovf8.assign( db.lens_all.size(), -1 );
                ovf16.assign( db.lens_all.size(), -1 );
                for( uint32_t k = 0; k < count; k++ ) {   // ctx().h_list still holds the (unsorted) device list order
                    const uint32_t i = ctx().h_list.p[k];
                    const size_t r = i / stride, sl = i % stride;
                    const uint32_t id = db.sorted_all[sl * world + r];
                    const int32_t c8 = ctx().h_first.p[k], c16 = ctx().h_first.p[count + k];
                    if( req.bit_width == 8 && c8 != INT_MAX ) ovf8[id] = c8 / 4;
                    if( c16 != INT_MAX ) ovf16[id] = c16 / 4;
                }
            }
            ctx().stats.d2h_bytes += 2 * (size_t) count * sizeof( int32_t );
            n_flagged = count;
        }
    }

    // ---- NW entered at 8 or 16 bit: which records overflow the reference's saturating tiers, and where ----
    // (only for the records static bounds cannot clear: usually none at 16 bit, everything at 8 bit)
    uint32_t * d_nw_counts = nullptr;
    if( nw && nglob > 0 && ( req.bit_width == 8 || req.bit_width == 16 ) ) {
        // every rank takes the same branch: whether any shard needs the emulation follows from the GLOBAL longest record
        const int64_t longest = nglob ? (int64_t) db.lens_all[db.sorted_all[0]] : 0;
        const bool need8 = req.bit_width == 8 && longest > nw_tier_safe_len( qlen, 8 );
        const bool need16 = longest > nw_tier_safe_len( qlen, 16 );
        if( need8 || need16 ) {
            have_overflow = true;
            run_on_all( [&]() {
                const size_t n8 = req.bit_width == 8 ? sorted_prefix_longer_than( nw_tier_safe_len( qlen, 8 ) ) : 0;
                const size_t n16 = sorted_prefix_longer_than( nw_tier_safe_len( qlen, 16 ) );
                ctx().d_first.reserve( 2 * stride * ( world + 1 ) + 4 );
                int32_t * d_f = ctx().d_first.p;  // [stride] 8-bit tier | [stride] 16-bit tier
                launch_fill( d_f, INT_MAX, 2 * stride );
                if( n8 ) run_nw_sat( query, qlen, 8, n8, d_f );
                if( n16 ) run_nw_sat( query, qlen, 16, n16, d_f + stride );
                if( world > 1 ) gather_i32( d_f, ctx().d_first.p + 2 * stride, 2 * stride );
            } );
            // overflow counters of this process' records, counted on the primary from the (gathered) columns
            d_first_gathered = world > 1 ? ctx().d_first.p + 2 * stride : ctx().d_first.p;
            ctx().d_counter.reserve( 64 );
            d_nw_counts = ctx().d_counter.p + 32;
            CUDA_OK( cudaMemsetAsync( d_nw_counts, 0, 8, ctx().stream ) );
            for( Ctx * c : g_devs ) {
                const int32_t * d_f = d_first_gathered + ( world > 1 ? (size_t) c->rank * 2 * stride : 0 );
                count_set_kernel<<<(unsigned) ( ( stride + 255 ) / 256 ), 256, 0, ctx().stream>>>( d_f, (uint32_t) stride, d_nw_counts );
                count_set_kernel<<<(unsigned) ( ( stride + 255 ) / 256 ), 256, 0, ctx().stream>>>( d_f + stride, (uint32_t) stride, d_nw_counts + 1 );
                ctx().stats.launches += 2;
            }
            if( !use_gpu_arrival ) {  // chunk sizes the arrival kernel does not cover: overflow blocks to the host emulation
                ctx().h_first.reserve( 2 * stride * world + 2 );
                CUDA_OK( cudaMemcpyAsync( ctx().h_first.p, d_first_gathered, 2 * stride * world * sizeof( int32_t ), cudaMemcpyDeviceToHost, ctx().stream ) );
                CUDA_OK( cudaStreamSynchronize( ctx().stream ) );
                ctx().stats.d2h_bytes += 2 * stride * world * sizeof( int32_t );
                ovf8.assign( db.lens_all.size(), -1 );
                ovf16.assign( db.lens_all.size(), -1 );
                for( size_t g = 0; g < nglob; g++ ) {
                    const size_t r = g % world, sl = g / world;
                    const int32_t c8 = ctx().h_first.p[r * 2 * stride + sl], c16 = ctx().h_first.p[r * 2 * stride + stride + sl];
                    if( req.bit_width == 8 && c8 != INT_MAX ) ovf8[db.sorted_all[g]] = c8 / 4;
                    if( c16 != INT_MAX ) ovf16[db.sorted_all[g]] = c16 / 4;
                }
            }
        }
    }

    // ---- heap-arrival order of the emulated one-thread reference run ----
    const uint32_t * order = nullptr;
    size_t narr = 0;
    DevBuf<int32_t> * d_pos = nullptr;
    if( !have_overflow ) {  // a function of the database alone: cached per (bit width, chunk size)
        PackedDb::Arrival & ar = arrival_for( req.bit_width, chunk, use_gpu_arrival );
        order = ar.order.data();
        narr = ar.order.size();
        d_pos = &ar.d_pos_of_sorted;
    }
    else if( use_gpu_arrival ) {
        ctx().h_order.reserve( nglob + 1 );
        const bool sparse = sparse_replay_applies( chunk_info_for( chunk ).narr, req.hitcount );
        narr = gpu_arrival_order( req.bit_width, chunk, d_first_gathered, stride, ctx().d_pos_tmp, sparse ? nullptr : ctx().h_order.p );
        order = sparse ? nullptr : ctx().h_order.p;   // complete once the stream is synchronised below; sparse: stays in ctx().d_order
        d_pos = &ctx().d_pos_tmp;
    }
    else {
        arrival_order( db.lens_all, ovf8.data(), ovf16.data(), chunk, req.bit_width, ctx().order_tmp );
        upload_positions( ctx().order_tmp, ctx().d_pos_tmp );
        order = ctx().order_tmp.data();
        narr = ctx().order_tmp.size();
        d_pos = &ctx().d_pos_tmp;
    }
    if( narr > 0 && nglob > 0 ) {
        ctx().d_arrival_scores.reserve( narr + 1 );
        const uint32_t total = (uint32_t) ( stride * world );
        scatter_arrival_kernel<<<( total + 255 ) / 256, 256, 0, ctx().stream>>>( d_all, d_pos->p, ctx().d_arrival_scores.p,
                (uint32_t) stride, (uint32_t) world, (uint32_t) nglob );
        ctx().stats.launches++;
        ctx().h_scores.reserve( narr + 1 );
        if( !sparse_replay_applies( narr, req.hitcount ) ) {
            CUDA_OK( cudaMemcpyAsync( ctx().h_scores.p, ctx().d_arrival_scores.p, narr * sizeof( int32_t ), cudaMemcpyDeviceToHost, ctx().stream ) );
            CUDA_OK( cudaEventRecord( ctx().ev_end, ctx().stream ) );
            CUDA_OK( cudaStreamSynchronize( ctx().stream ) );
            ctx().stats.d2h_bytes += narr * sizeof( int32_t );
            replay_topk( ctx().h_scores.p, order, narr, req.hitcount, hits );
        }
        else {
            // Large databases: only the first m arrivals and the later arrivals that beat the heap's root after those m can
            // touch the heap (replay.cpp: replay_topk_sparse) - 4 m bytes and a few hundred candidates cross PCIe instead of
            // 4 bytes per record of the WHOLE database on every rank, and the host looks at m records instead of all.
            const size_t m = sparse_replay_prefix( req.hitcount );
            const uint32_t * d_order = order ? nullptr : ctx().d_order.p;   // arrival position -> provider index, where the host has no copy
            CUDA_OK( cudaMemcpyAsync( ctx().h_scores.p, ctx().d_arrival_scores.p, m * sizeof( int32_t ), cudaMemcpyDeviceToHost, ctx().stream ) );
            if( d_order ) CUDA_OK( cudaMemcpyAsync( ctx().h_order.p, d_order, m * sizeof( uint32_t ), cudaMemcpyDeviceToHost, ctx().stream ) );
            CUDA_OK( cudaEventRecord( ctx().ev_end, ctx().stream ) );   // (recorded again behind the candidate fetch)
            CUDA_OK( cudaStreamSynchronize( ctx().stream ) );
            ctx().stats.d2h_bytes += m * sizeof( int32_t ) + ( d_order ? m * sizeof( uint32_t ) : 0 );
            auto fetch = [&]( int32_t thr, const int32_t ** cand_scores, const uint32_t ** cand_ids ) -> size_t {
                const size_t cap = narr - m;
                ctx().d_cand_pos.reserve( cap + 1 );
                ctx().d_cand_score.reserve( cap + 1 );
                ctx().d_counter.reserve( 64 );
                uint32_t * d_count = ctx().d_counter.p + 40;
                CUDA_OK( cudaMemsetAsync( d_count, 0, 4, ctx().stream ) );
                compact_greater_kernel<<<(unsigned) ( ( cap + 255 ) / 256 ), 256, 0, ctx().stream>>>( ctx().d_arrival_scores.p, (uint32_t) m, (uint32_t) narr, thr,
                        ctx().d_cand_pos.p, ctx().d_cand_score.p, d_count );
                ctx().stats.launches++;
                uint32_t count = 0;
                CUDA_OK( cudaMemcpyAsync( &count, d_count, 4, cudaMemcpyDeviceToHost, ctx().stream ) );
                CUDA_OK( cudaStreamSynchronize( ctx().stream ) );
                ctx().h_cand_pos.reserve( 2 * (size_t) count + 2 );   // [count] positions | [count] ids in arrival order
                ctx().h_cand_score.reserve( 2 * (size_t) count + 2 ); // [count] scores as compacted | [count] in arrival order
                ctx().h_cand_id.reserve( (size_t) count + 2 );
                if( count > 0 ) {
                    CUDA_OK( cudaMemcpyAsync( ctx().h_cand_pos.p, ctx().d_cand_pos.p, count * sizeof( uint32_t ), cudaMemcpyDeviceToHost, ctx().stream ) );
                    CUDA_OK( cudaMemcpyAsync( ctx().h_cand_score.p, ctx().d_cand_score.p, count * sizeof( int32_t ), cudaMemcpyDeviceToHost, ctx().stream ) );
                    if( d_order ) {
                        ctx().d_cand_id.reserve( count + 1 );
                        gather_u32_kernel<<<( count + 255 ) / 256, 256, 0, ctx().stream>>>( d_order, ctx().d_cand_pos.p, count, ctx().d_cand_id.p );
                        ctx().stats.launches++;
                        CUDA_OK( cudaMemcpyAsync( ctx().h_cand_id.p, ctx().d_cand_id.p, count * sizeof( uint32_t ), cudaMemcpyDeviceToHost, ctx().stream ) );
                    }
                }
                CUDA_OK( cudaEventRecord( ctx().ev_end, ctx().stream ) );
                CUDA_OK( cudaStreamSynchronize( ctx().stream ) );
                ctx().stats.d2h_bytes += 4 + (size_t) count * ( d_order ? 12 : 8 );
                // the compaction is unordered: back into arrival order
                std::vector<uint32_t> perm( count );
                std::iota( perm.begin(), perm.end(), 0u );
                std::sort( perm.begin(), perm.end(), [&]( uint32_t x, uint32_t y ) { return ctx().h_cand_pos.p[x] < ctx().h_cand_pos.p[y]; } );
                uint32_t * ids = ctx().h_cand_pos.p + count;
                int32_t * sc = ctx().h_cand_score.p + count;
                for( uint32_t j = 0; j < count; j++ ) {
                    const uint32_t src = perm[j];
                    sc[j] = ctx().h_cand_score.p[src];
                    ids[j] = d_order ? ctx().h_cand_id.p[src] : order[ctx().h_cand_pos.p[src]];
                }
                *cand_scores = sc;
                *cand_ids = ids;
                return count;
            };
            replay_topk_sparse( ctx().h_scores.p, order ? order : ctx().h_order.p, m, req.hitcount, fetch, hits );
        }
    }
    else {
        CUDA_OK( cudaEventRecord( ctx().ev_end, ctx().stream ) );
        CUDA_OK( cudaStreamSynchronize( ctx().stream ) );
    }

    if( d_nw_counts ) {  // records of this shard the reference would have re-aligned (NW tiers)
        uint32_t c[2] = { 0, 0 };
        CUDA_OK( cudaMemcpy( c, d_nw_counts, 8, cudaMemcpyDeviceToHost ) );
        if( req.bit_width == 8 ) ctx().stats.overflow_8 = c[0];
        ctx().stats.overflow_16 = c[1];
    }
    for( uint32_t k = 0; k < n_flagged; k++ ) {  // valid after the synchronisation above
        if( req.bit_width == 8 && ctx().h_first.p[k] != INT_MAX ) ctx().stats.overflow_8++;
        if( ctx().h_first.p[n_flagged + k] != INT_MAX ) ctx().stats.overflow_16++;
    }
}

// Searches over several query frames, or whose hit list holds tied entries of one record: every frame is scored on the
// GPU, the scores (and, for 8/16-bit entry, the overflow blocks) come back to the host, and the reference's chunk /
// frame / overflow-list schedule is replayed there (replay.cpp: general_search_order).
static void search_frames( const SearchRequest & req, std::vector<FrameHit> & fh ) {
    PackedDb & db = ctx().db;
    const size_t nloc = db.sorted_id.size();
    const size_t nglob = db.sorted_all.size();
    const size_t world = (size_t) ctx().world;
    const size_t stride = ( nglob + world - 1 ) / world;
    const size_t total = stride * world;
    const size_t nslots = db.lens_all.size();
    const bool nw = req.search_type == NEEDLEMAN_WUNSCH;
    const int Q = (int) req.frames.size();
    const bool tiers = req.bit_width == 8 || req.bit_width == 16;
    std::vector<std::vector<int32_t>> scores( Q ), ovf8( Q ), ovf16( Q );
    ctx().stats.cells = 0;
    for( int q = 0; q < Q; q++ ) {
        const QueryFrame & f = req.frames[q];
        scores[q].assign( nslots, 0 );
        ctx().stats.cells += (uint64_t) f.len * process_residues();
        if( nglob == 0 ) continue;
        const int32_t * d_all = score_query( f.seq, f.len, nw, req.bit_width );
        ctx().h_scores.reserve( total + 1 );
        CUDA_OK( cudaMemcpyAsync( ctx().h_scores.p, d_all, total * sizeof( int32_t ), cudaMemcpyDeviceToHost, ctx().stream ) );
        CUDA_OK( cudaStreamSynchronize( ctx().stream ) );
        ctx().stats.d2h_bytes += total * sizeof( int32_t );
        for( size_t g = 0; g < nglob; g++ ) scores[q][db.sorted_all[g]] = ctx().h_scores.p[( g % world ) * stride + g / world];
        if( !tiers ) continue;

        bool have = false;   // d_first of every context: [stride] 8-bit tier | [stride] 16-bit tier, first overflowing column
        if( !nw ) {  // SW: score >= 255 / 65535 (src/algo/simd/search_simd_sw.c:372-379)
            const int32_t thr = req.bit_width == 8 ? 255 : 65535;
            for( size_t g = 0; g < nglob && !have; g++ ) have = scores[q][db.sorted_all[g]] >= thr;
            if( have ) {
                const Plan pl = make_plan( f.len );
                const int32_t * h_all = ctx().h_scores.p;   // the gathered scores, on the host
                run_on_all( [&]() {
                    std::vector<uint32_t> slots;
                    const size_t nl = ctx().db.sorted_id.size();
                    for( size_t sl = 0; sl < nl; sl++ )
                        if( h_all[(size_t) ctx().rank * stride + sl] >= thr ) slots.push_back( (uint32_t) sl );
                    ctx().d_first.reserve( 2 * stride * ( world + 1 ) + 4 );
                    int32_t * d_fl = ctx().d_first.p;
                    launch_fill( d_fl, INT_MAX, 2 * stride );
                    run_wide( slots, f.seq, f.len, pl, false, true, d_fl, d_fl + stride, 1, ctx().stream );
                    if( world > 1 ) gather_i32( d_fl, ctx().d_first.p + 2 * stride, 2 * stride );
                } );
            }
        }
        else {  // NW: the saturating tiers, only for the entries static bounds cannot clear (nw_sat.cuh)
            const int64_t longest = (int64_t) db.lens_all[db.sorted_all[0]];
            have = ( req.bit_width == 8 && longest > nw_tier_safe_len( f.len, 8 ) ) || longest > nw_tier_safe_len( f.len, 16 );
            if( have ) {
                run_on_all( [&]() {
                    const size_t n8 = req.bit_width == 8 ? sorted_prefix_longer_than( nw_tier_safe_len( f.len, 8 ) ) : 0;
                    const size_t n16 = sorted_prefix_longer_than( nw_tier_safe_len( f.len, 16 ) );
                    ctx().d_first.reserve( 2 * stride * ( world + 1 ) + 4 );
                    int32_t * d_fl = ctx().d_first.p;
                    launch_fill( d_fl, INT_MAX, 2 * stride );
                    if( n8 ) run_nw_sat( f.seq, f.len, 8, n8, d_fl );
                    if( n16 ) run_nw_sat( f.seq, f.len, 16, n16, d_fl + stride );
                    if( world > 1 ) gather_i32( d_fl, ctx().d_first.p + 2 * stride, 2 * stride );
                } );
            }
        }
        if( !have ) continue;
        const int32_t * d_f_all = world > 1 ? ctx().d_first.p + 2 * stride : ctx().d_first.p;
        ctx().h_first.reserve( 2 * total + 2 );
        CUDA_OK( cudaMemcpyAsync( ctx().h_first.p, d_f_all, 2 * total * sizeof( int32_t ), cudaMemcpyDeviceToHost, ctx().stream ) );
        CUDA_OK( cudaStreamSynchronize( ctx().stream ) );
        ctx().stats.d2h_bytes += 2 * total * sizeof( int32_t );
        ovf8[q].assign( nslots, -1 );
        ovf16[q].assign( nslots, -1 );
        for( size_t g = 0; g < nglob; g++ ) {
            const size_t r = g % world, sl = g / world;
            const int32_t c8 = ctx().h_first.p[r * 2 * stride + sl], c16 = ctx().h_first.p[r * 2 * stride + stride + sl];
            if( req.bit_width == 8 && c8 != INT_MAX ) ovf8[q][db.sorted_all[g]] = c8 / 4;
            if( c16 != INT_MAX ) ovf16[q][db.sorted_all[g]] = c16 / 4;
        }
    }
    CUDA_OK( cudaEventRecord( ctx().ev_end, ctx().stream ) );
    CUDA_OK( cudaStreamSynchronize( ctx().stream ) );
    size_t n8 = 0, n16 = 0;
    general_search_order( db.lens_all, g_chunk_size * (size_t) db.buffer_max, req.bit_width, Q, db.buffer_max, scores, ovf8, ovf16,
            req.hitcount, fh, &n8, &n16 );
    ctx().stats.overflow_8 = n8;
    ctx().stats.overflow_16 = n16;
}
// ---- result records: reference src/algo/aligner.c:62-93 ----
// The reference rebuilds the database side from the record with us_prepare_sequence( seq, len, frame, strand )
// (src/util/util_sequence.c:384-417): nucleotide records always come back as the forward strand, and a translated
// record of a one-strand search is translated from the complementary strand whichever strand was scored.  The
// alignment is computed on that sequence; reproduced as is.
// defer: if set, the alignment is left to the caller (gpu_search_batch aligns the hits of many queries together) and
// *defer receives whether the alignment runs on the very sequence that was scored
static void finish_results( const SearchRequest & req, const std::vector<FrameHit> & fh, std::vector<HitAlignment> & out, cudaStream_t tb_stream,
        bool * defer = nullptr ) {
    const uint32_t bm = (uint32_t) ctx().db.buffer_max;
    const bool nw = req.search_type == NEEDLEMAN_WUNSCH;
    const bool translate_db = g_symtype == TRANS_DB || g_symtype == TRANS_BOTH;
    const signed char * map = ( g_symtype == NUCLEOTIDE || translate_db ) ? map_ncbi_nt16 : map_ncbi_aa;
    out.clear();
    out.resize( fh.size() );
    bool scored_sequence = true;  // the alignment runs on the very sequence that was scored
    for( size_t i = 0; i < fh.size(); i++ ) {
        const uint32_t rec_id = fh[i].slot / bm, b = fh[i].slot % bm;
        out[i].hit = Hit{ rec_id, fh[i].score };
        out[i].query_id = fh[i].query_id;
        if( g_symtype == NUCLEOTIDE ) {
            out[i].db_strand = (int) b;
            out[i].db_frame = 0;
            if( b ) scored_sequence = false;
        }
        else if( translate_db ) {
            out[i].db_strand = bm == 6 ? (int) ( b / 3 ) : g_query_strands;  // db_adapter.c:80,92
            out[i].db_frame = (int) ( b % 3 );
            if( bm == 3 && g_query_strands == FORWARD_STRAND ) scored_sequence = false;
        }
        p_seqinfo rec = ssa_db_get_sequence( rec_id );
        if( !rec ) fatal( "Could not get sequence from DB: %ld", (long) rec_id );  // aligner.c:66-68
        if( translate_db ) {
            std::vector<uint8_t> nt( rec->seqlen + 1 );
            map_db_sequence( rec->seq, rec->seqlen, nt.data(), map );
            translate_codes( true, nt.data(), rec->seqlen, out[i].db_strand, out[i].db_frame, out[i].db_codes );
        }
        else {
            out[i].db_codes.resize( rec->seqlen );
            map_db_sequence( rec->seq, rec->seqlen, out[i].db_codes.data(), map );
        }
    }
    if( defer ) {
        *defer = scored_sequence;
        return;
    }
    if( req.align_type == COMPUTE_ALIGNMENT && !out.empty() ) {
        if( req.frames.size() == 1 ) {
            traceback_hits( tb_stream, req.frames[0].seq, req.frames[0].len, nw, out, &ctx().stats, scored_sequence );
        }
        else {
            for( size_t q = 0; q < req.frames.size(); q++ ) {
                std::vector<HitAlignment> group;
                std::vector<size_t> where;
                for( size_t i = 0; i < out.size(); i++ )
                    if( (size_t) out[i].query_id == q ) {
                        where.push_back( i );
                        group.push_back( std::move( out[i] ) );
                    }
                if( group.empty() ) continue;
                traceback_hits( tb_stream, req.frames[q].seq, req.frames[q].len, nw, group, &ctx().stats, scored_sequence );
                for( size_t j = 0; j < where.size(); j++ ) out[where[j]] = std::move( group[j] );
            }
        }
    }
}

void gpu_search( const SearchRequest & req, std::vector<HitAlignment> & out ) {
    const double t_call = wall_ms();
    build_db();
    PackedDb & db = ctx().db;
    const uint32_t bm = (uint32_t) db.buffer_max;
    out.clear();
    reset_call_stats();
    CUDA_OK( cudaEventRecord( ctx().ev_begin, ctx().stream ) );

    std::vector<FrameHit> fh;
    bool done = false;
    if( req.frames.size() == 1 ) {
        std::vector<Hit> hits;
        search_single( req.frames[0].seq, req.frames[0].len, req, hits );
        // entries of one record that tie on the score are ordered by the C library's sort in the reference: replay then
        done = true;
        for( size_t i = 1; i < hits.size() && bm > 1; i++ )
            if( hits[i].score == hits[i - 1].score && hits[i].db_id / bm == hits[i - 1].db_id / bm ) done = false;
        if( done ) {
            fh.resize( hits.size() );
            for( size_t i = 0; i < hits.size(); i++ ) fh[i] = FrameHit{ (uint32_t) hits[i].db_id, 0, hits[i].score };
        }
    }
    if( !done ) search_frames( req, fh );

    if( ctx().stats.overflow_8 || ctx().stats.overflow_16 )  // reference src/algo/manager.c:157-160
        print_info( "Overflow occurred: %ld sequences were re-aligned with 16 bit, and %ld sequences with 64 bit\n",
                (long) ctx().stats.overflow_8, (long) ctx().stats.overflow_16 );

    finish_results( req, fh, out, ctx().stream );

    merge_call_stats();
    float ms = 0;
    cudaEventElapsedTime( &ms, ctx().ev_begin, ctx().ev_end );
    ctx().stats.device_ms = ms;
    ctx().stats.call_ms = wall_ms() - t_call;
}

// A batch of independent queries with the same search parameters: the results are those of one
// gpu_search per query.  Queries that take the device-side top-k path are software-pipelined: the scoring and
// selection kernels of query i+1 are enqueued before the host waits for the results of query i, so the GPU does not
// idle during the host's share of a call (result records, launches, the PCIe round trip); alignments of query i run on
// the auxiliary stream beside the scoring of query i+1.  Device buffers are shared (one in-order stream: the copies
// of query i leave before query i+1 overwrites them), pinned staging buffers exist twice.
void gpu_search_batch( const std::vector<SearchRequest> & reqs, std::vector<std::vector<HitAlignment>> & outs ) {
    const double t_call = wall_ms();
    build_db();
    PackedDb & db = ctx().db;
    const size_t n = reqs.size();
    outs.assign( n, std::vector<HitAlignment>() );
    bool pipelined = db.buffer_max == 1 && n > 1;
    for( const SearchRequest & r : reqs ) pipelined = pipelined && r.frames.size() == 1;
    if( !pipelined ) {
        for( size_t i = 0; i < n; i++ ) gpu_search( reqs[i], outs[i] );
        return;
    }
    reset_call_stats();
    const size_t nglob = db.sorted_all.size();
    const size_t world = (size_t) ctx().world;
    const size_t stride = ( nglob + world - 1 ) / world;
    uint64_t cells = 0;
    // Units of work: two consecutive queries of the same kind share one scoring pass (two queries per packed word,
    // score_shard_dual) where that applies, otherwise a unit is one query.  The scoring and device top-k of unit u+1 are enqueued
    // before the host waits for unit u; result buffers are per query (index & 3), pinned upload staging per unit (u & 1).
    struct Unit {
        size_t first;
        int count;
    };
    std::vector<Unit> units;
    std::vector<DualPlan> plans;
    for( size_t i = 0; i < n; ) {
        DualPlan dp;
        const bool nwq = reqs[i].search_type == NEEDLEMAN_WUNSCH;
        const bool pair = i + 1 < n && reqs[i].search_type == reqs[i + 1].search_type && world == 1 && g_devs.size() == 1 &&
                !nw_narrow_scores_needed( reqs[i].frames[0].len, nwq, reqs[i].bit_width ) &&
                !nw_narrow_scores_needed( reqs[i + 1].frames[0].len, nwq, reqs[i + 1].bit_width ) &&
                dual_plan( reqs[i].frames[0].len, reqs[i + 1].frames[0].len, reqs[i].search_type == NEEDLEMAN_WUNSCH, dp );
        units.push_back( Unit{ i, pair ? 2 : 1 } );
        plans.push_back( dp );
        i += pair ? 2 : 1;
    }
    bool armed[4] = { false, false, false, false };
    const size_t chunk = g_chunk_size;
    // Alignments: the hits of many queries are aligned together (traceback_hits_multi) - the tile launches of one query's ten
    // hits are latency-bound and, queued behind the persistent scoring launch of the next unit, would stall the pipeline once
    // per query.  Pending queries are flushed when ~2000 hits (or ~1.5 GB of traceback scratch) have accumulated, and at the end.
    struct PendingAlign {
        size_t query;
        bool scored_sequence;
    };
    std::vector<PendingAlign> pending;
    size_t pending_hits = 0, pending_cells = 0;
    auto flush_alignments = [&]() {
        if( pending.empty() ) return;
        const bool nw = reqs[pending[0].query].search_type == NEEDLEMAN_WUNSCH;
        std::vector<TbQuery> queries;
        std::vector<uint32_t> query_of_hit;
        std::vector<char> scored;
        std::vector<HitAlignment> all;
        for( const PendingAlign & pa : pending ) {
            const SearchRequest & r = reqs[pa.query];
            queries.push_back( TbQuery{ r.frames[0].seq, r.frames[0].len } );
            for( HitAlignment & h : outs[pa.query] ) {
                query_of_hit.push_back( (uint32_t) ( queries.size() - 1 ) );
                scored.push_back( pa.scored_sequence ? 1 : 0 );
                all.push_back( std::move( h ) );
            }
        }
        traceback_hits_multi( ctx().stream_aux, queries, query_of_hit, nw, all, &ctx().stats, scored );
        size_t k = 0;
        for( const PendingAlign & pa : pending )
            for( HitAlignment & h : outs[pa.query] ) h = std::move( all[k++] );
        pending.clear();
        pending_hits = pending_cells = 0;
    };
    auto enqueue_selection = [&]( size_t i, const int32_t * d_all ) {
        const SearchRequest & r = reqs[i];
        const int rslot = (int) ( i & 3 );
        const bool nw = r.search_type == NEEDLEMAN_WUNSCH;
        // the replay path of this query, should it be needed, runs after later scoring passes have overwritten d_all
        ctx().d_all_slot[rslot].reserve( stride * world + 2 );
        if( nglob ) CUDA_OK( cudaMemcpyAsync( ctx().d_all_slot[rslot].p, d_all, stride * world * sizeof( int32_t ), cudaMemcpyDeviceToDevice, ctx().stream ) );
        PackedDb::Arrival & base = arrival_for( r.bit_width, chunk, chunk <= (size_t) ARRIVAL_MAX_CHUNK );
        armed[rslot] = topk_applicable( base, r.hitcount );
        if( armed[rslot] ) enqueue_topk( d_all, base, r.hitcount, !nw && ( r.bit_width == 8 || r.bit_width == 16 ), rslot );
    };
    auto enqueue = [&]( size_t u ) {
        const Unit & un = units[u];
        g_stage_slot = (int) ( u & 1 );
        const SearchRequest & r0 = reqs[un.first];
        const bool nw = r0.search_type == NEEDLEMAN_WUNSCH;
        if( un.count == 2 ) {
            const SearchRequest & r1 = reqs[un.first + 1];
            score_shard_dual( r0.frames[0].seq, r0.frames[0].len, r1.frames[0].seq, r1.frames[0].len, nw, plans[u] );
            enqueue_selection( un.first, ctx().d_scores.p );
            enqueue_selection( un.first + 1, ctx().d_scores2.p );
        }
        else
            enqueue_selection( un.first, score_query( r0.frames[0].seq, r0.frames[0].len, nw, r0.bit_width ) );
    };
    CUDA_OK( cudaEventRecord( ctx().ev_begin, ctx().stream ) );
    enqueue( 0 );
    for( size_t u = 0; u < units.size(); u++ ) {
        if( u + 1 < units.size() ) enqueue( u + 1 );
        for( int j = 0; j < units[u].count; j++ ) {
            const size_t i = units[u].first + j;
            const SearchRequest & r = reqs[i];
            const int rslot = (int) ( i & 3 );
            std::vector<Hit> hits;
            ctx().stats.overflow_8 = ctx().stats.overflow_16 = 0;
            if( !( armed[rslot] && collect_topk( r.hitcount, r.bit_width, rslot, hits ) ) ) {
                // ties straddle the cut-off (or the hit count covers the database): the arrival-order replay, from this query's
                // own copy of the scores; it queues behind the next unit's scoring
                g_stage_slot = (int) ( u & 1 );
                select_replay( ctx().d_all_slot[rslot].p, r.frames[0].seq, r.frames[0].len, r, hits );
            }
            if( ctx().stats.overflow_8 || ctx().stats.overflow_16 )
                print_info( "Overflow occurred: %ld sequences were re-aligned with 16 bit, and %ld sequences with 64 bit\n",
                        (long) ctx().stats.overflow_8, (long) ctx().stats.overflow_16 );
            cells += (uint64_t) r.frames[0].len * process_residues();
            std::vector<FrameHit> fh( hits.size() );
            for( size_t k = 0; k < hits.size(); k++ ) fh[k] = FrameHit{ (uint32_t) hits[k].db_id, 0, hits[k].score };
            if( r.align_type == COMPUTE_ALIGNMENT ) {
                if( !pending.empty() && reqs[pending[0].query].search_type != r.search_type ) flush_alignments();
                bool scored_sequence = true;
                finish_results( r, fh, outs[i], ctx().stream_aux, &scored_sequence );
                if( !outs[i].empty() ) {
                    pending.push_back( PendingAlign{ i, scored_sequence } );
                    pending_hits += outs[i].size();
                    for( const HitAlignment & h : outs[i] ) pending_cells += r.frames[0].len * h.db_codes.size();
                    if( pending_hits >= 2000 || pending_cells >= ( 1ull << 30 ) ) flush_alignments();
                }
            }
            else
                finish_results( r, fh, outs[i], ctx().stream_aux );
        }
    }
    flush_alignments();
    CUDA_OK( cudaStreamSynchronize( ctx().stream ) );
    g_stage_slot = 0;
    merge_call_stats();
    float ms = 0;
    cudaEventElapsedTime( &ms, ctx().ev_begin, ctx().ev_end );
    ctx().stats.device_ms = ms;
    ctx().stats.score_kernel_ms = 0;
    ctx().stats.cells = cells;
    ctx().stats.call_ms = wall_ms() - t_call;
}

// ---- additive API (include/libssa_b200.h) ----
extern "C" {
__attribute__( ( visibility( "default" ) ) ) void ssa_b200_set_device( int device ) {
    if( !g_devs.empty() && ( g_devs.size() > 1 || g_devs[0]->device != device ) ) gpu_release();
    g_cfg.device = device;
}
__attribute__( ( visibility( "default" ) ) ) void ssa_b200_set_device_count( int n ) {
    if( n < 1 ) fatal( "libssa_b200: invalid device count %d", n );
    if( !g_devs.empty() && (int) g_devs.size() != n ) gpu_release();
    g_cfg.ndev = n;
}
__attribute__( ( visibility( "default" ) ) ) int ssa_b200_get_device_count( void ) {
    ensure_ctx();
    return (int) g_devs.size();
}
__attribute__( ( visibility( "default" ) ) ) void ssa_b200_set_shard( int rank, int world_size ) {
    if( world_size < 1 || rank < 0 || rank >= world_size ) fatal( "libssa_b200: invalid shard %d of %d", rank, world_size );
    if( g_devs.size() > 1 && world_size > 1 )
        fatal( "libssa_b200: this process already drives %d GPUs; sharding over processes as well is not supported.", (int) g_devs.size() );
    g_cfg.rank = rank;
    g_cfg.world = world_size;
    if( g_devs.size() == 1 ) {   // takes effect at the next search (the shard is rebuilt)
        g_devs[0]->rank = rank;
        g_devs[0]->world = world_size;
    }
}
__attribute__( ( visibility( "default" ) ) ) void ssa_b200_set_allgather( ssa_b200_allgather_fn fn, void * c ) {
    g_cfg.allgather = fn;
    g_cfg.allgather_ctx = c;
}
// ---- the collective inside the library, one process per GPU: rank 0 creates the NCCL id, the host program carries its
// 128 bytes to the other ranks (any channel: MPI, a file, torch.distributed), every rank joins ----
__attribute__( ( visibility( "default" ) ) ) void ssa_b200_nccl_unique_id( void * out_128_bytes ) {
    load_nccl();
    static_assert( sizeof( ncclUniqueId ) == 128, "ncclUniqueId is 128 bytes" );
    ncclUniqueId id;
    NCCL_OK( g_nccl.GetUniqueId( &id ) );
    memcpy( out_128_bytes, &id, sizeof( id ) );
}
__attribute__( ( visibility( "default" ) ) ) void ssa_b200_nccl_init( const void * id_128_bytes, int rank, int world_size ) {
    if( world_size < 1 || rank < 0 || rank >= world_size ) fatal( "libssa_b200: invalid rank %d of %d", rank, world_size );
    g_cfg.rank = rank;
    g_cfg.world = world_size;
    ensure_ctx();
    if( g_devs.size() > 1 ) fatal( "libssa_b200: ssa_b200_nccl_init is for one GPU per process." );
    Ctx & c = *g_devs[0];
    c.rank = rank;
    c.world = world_size;
    if( c.comm ) {
        g_nccl.CommDestroy( c.comm );
        c.comm = nullptr;
    }
    if( world_size == 1 ) return;
    load_nccl();
    ncclUniqueId id;
    memcpy( &id, id_128_bytes, sizeof( id ) );
    CUDA_OK( cudaSetDevice( c.device ) );
    NCCL_OK( g_nccl.CommInitRank( &c.comm, world_size, id, rank ) );
}
__attribute__( ( visibility( "default" ) ) ) void ssa_b200_prepare_db( void ) { build_db(); }
__attribute__( ( visibility( "default" ) ) ) void ssa_b200_get_stats( ssa_b200_stats_t * out ) {
    if( !out ) return;
    if( g_devs.empty() )
        memset( out, 0, sizeof( *out ) );
    else
        *out = g_devs[0]->stats;
}
__attribute__( ( visibility( "default" ) ) ) int ssa_b200_version( void ) { return 100; }
// host-only: the packed tier's passes for a query of qlen rows (see plan_passes)
__attribute__( ( visibility( "default" ) ) ) size_t ssa_b200_host_plan_passes( size_t qlen, uint64_t total_cols, uint64_t longest_cols, int sms,
        int * G_out, int * R_out, size_t * rowbase_out, size_t max_passes ) {
    if( qlen == 0 ) return 0;
    const Plan pl = make_plan( qlen );
    std::vector<PassLaunch> v;
    if( pl.npass > 1 )
        v = plan_passes( qlen, pl.R, pl.npass, total_cols, longest_cols, sms );
    else {
        PassLaunch ps;
        ps.G = pl.G;
        ps.R = pl.R;
        v.push_back( ps );
    }
    for( size_t i = 0; i < v.size() && i < max_passes; i++ ) {
        G_out[i] = v[i].G;
        R_out[i] = v[i].R;
        rowbase_out[i] = v[i].rowbase;
    }
    return v.size();
}
// host-only: the work-queue partition of a stream of pairs (see partition_batches / batch_table)
__attribute__( ( visibility( "default" ) ) ) size_t ssa_b200_host_partition_batches( const uint64_t * pair_col, size_t npairs, size_t first_pair,
        uint32_t cols, uint32_t groups, uint32_t min_cols, uint32_t * first_pair_out, size_t max_batches ) {
    const std::vector<uint32_t> first = partition_batches( pair_col, npairs, first_pair, cols, groups, min_cols );
    for( size_t i = 0; i < first.size() && i < max_batches; i++ ) first_pair_out[i] = first[i];
    return first.size();
}
}
