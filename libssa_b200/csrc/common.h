// Internal declarations shared by the host-side translation units of libssa_b200.so.
#pragma once

#include <cstddef>
#include <cstdint>
#include <string>
#include <functional>
#include <vector>

#include "../../include/libssa.h"
#include "../../include/libssa_extern_db.h"
#include "../../include/libssa_b200.h"

// reference src/util/util.h:34-35
constexpr int SMITH_WATERMAN = 0;
constexpr int NEEDLEMAN_WUNSCH = 1;

constexpr int SCORE_MATRIX_DIM = 32;  // reference src/matrices.h:26
constexpr size_t DEFAULT_CHUNK_SIZE = 1000;  // reference src/util/util.h:50

// ---- sequence_t / query_t: layout kept identical to reference src/libssa_datatypes.h:31-34,100-107
// because reference tests (and some callers) peek into the opaque p_query.
struct sequence_t {
    char * seq;
    size_t len;
};

struct _query {
    sequence_t nt[2];
    sequence_t aa[6];
    char * header;
    size_t headerlen;
    const char * map;
    const char * sym;
};

// ---- util.cpp (reference src/util/util.c)
extern int g_output_mode;
void * xmalloc( size_t size );
void * xrealloc( void * ptr, size_t size );
[[noreturn]] void fatal( const char * format, ... );
void print_info( const char * format, ... );
void print_warning( const char * format, ... );
void print_error( const char * format, ... );
double wall_ms();

// ---- sequence.cpp (reference src/util/util_sequence.c)
extern signed char map_ncbi_aa_storage[256];
#define map_ncbi_aa map_ncbi_aa_storage
extern signed char map_ncbi_nt16_storage[256];
#define map_ncbi_nt16 map_ncbi_nt16_storage
extern const char * sym_ncbi_aa;
extern const char * sym_ncbi_nt16;
extern int g_symtype;
extern int g_query_strands;
extern int g_db_gencode;  // genetic code the database translation table was built with
bool gencode_is_valid( int code );
// DB-side mapping: unknown symbols -> 0 (+ warning); returns number of unknown symbols
size_t map_db_sequence( const char * ascii, size_t len, uint8_t * out, const signed char * map );
void revcompl_codes( const uint8_t * in, size_t len, uint8_t * out );
void init_translation_tables( int q_gencode, int d_gencode );
void translate_codes( bool db_side, const uint8_t * nt, size_t len, int strand, int frame, std::vector<uint8_t> & out );

// ---- matrices.cpp (reference src/matrices.c)
struct ScoreMatrix {
    bool set = false;
    int64_t m[SCORE_MATRIX_DIM * SCORE_MATRIX_DIM];  // (db_symbol << 5) + query_symbol
    int64_t lo = 0, hi = 0;
    uint64_t version = 0;  // bumped on every change: device caches key on it
};
extern ScoreMatrix g_matrix;
bool is_constant_scoring();
void mat_init_constant_scoring( int8_t match, int8_t mismatch );
void mat_init_from_file( const char * path );
void mat_init_from_string( const char * text );
void mat_init_buildin( const char * name );
void mat_free();

// ---- query.cpp (reference src/query.c)
p_query query_read_from_file( const char * filename );
p_query query_read_from_string( const char * sequence );
void query_free( p_query q );

// ---- api.cpp globals (reference src/libssa.c:35-36, src/algo/manager.c:40, src/util/thread_pool.c)
extern int8_t g_gapO, g_gapE;
extern size_t g_chunk_size;
extern size_t g_thread_count;
extern uint64_t g_db_epoch;  // bumped by init_db / init_symbol_translation: invalidates the packed DB

// ---- replay.cpp: one-thread arrival order of the reference scheduler + bounded heap replay
struct Hit {
    uint64_t db_id;
    int64_t score;
};
// order of heap arrival for a search entered at `bit_width`; lens[] in provider order (0 = skipped record);
// ovf8/ovf16: nullptr or per-record overflow block (-1 = none)
void arrival_order( const std::vector<uint64_t> & lens, const int64_t * ovf8, const int64_t * ovf16, size_t chunk,
        int bit_width, std::vector<uint32_t> & order );
// replays thread heap + merge heap + final sort (reference minheap.c, manager.c:141-172)
void replay_topk( const int32_t * scores_in_arrival_order, const uint32_t * order, size_t n, size_t k,
        std::vector<Hit> & out );
// the same result from the first m >= k arrivals plus the later arrivals that score above the heap's root after those m
// (fetch( root, &scores, &ids ) returns them in arrival order): all other records are rejected by the reference's heap
void replay_topk_sparse( const int32_t * prefix_scores, const uint32_t * prefix_ids, size_t m, size_t k,
        const std::function<size_t( int32_t, const int32_t **, const uint32_t ** )> & fetch, std::vector<Hit> & out );
// General form for searches with several query frames and/or several database entries per record (strands, 6-frame
// translation).  `slots` = database entries in provider order (buffer_max per record, length 0 = absent);
// scores/ovf8/ovf16 are indexed [frame][slot].  Produces the reference's final, sorted hit list.
struct FrameHit {
    uint32_t slot;
    uint8_t query_id;
    int64_t score;
};
void general_search_order( const std::vector<uint64_t> & slot_lens, size_t chunk_slots, int bit_width, int q_count, int buffer_max,
        const std::vector<std::vector<int32_t>> & scores, const std::vector<std::vector<int32_t>> & ovf8,
        const std::vector<std::vector<int32_t>> & ovf16, size_t hitcount, std::vector<FrameHit> & out,
        size_t * n_ovf8 = nullptr, size_t * n_ovf16 = nullptr );

// ---- search.cu: the device side
struct QueryFrame {          // one entry of the reference's search_data_t::queries[] (src/libssa_datatypes.h:49-61)
    const uint8_t * seq;     // code string
    size_t len;
    int strand, frame;
};
struct SearchRequest {
    std::vector<QueryFrame> frames;  // 1 frame for AMINOACID / NUCLEOTIDE forward / TRANS_DB, up to 6 otherwise
    int search_type;   // SMITH_WATERMAN / NEEDLEMAN_WUNSCH
    int bit_width;     // 8 / 16 / 64: entry tier of the emulated reference scheduler
    size_t hitcount;
    int align_type;
};
struct HitAlignment {
    Hit hit;                        // db_id = provider record index
    int query_id = 0;               // index into SearchRequest::frames
    int db_strand = 0, db_frame = 0; // as stored by the reference's db_adapter for the scored database entry
    std::vector<uint8_t> db_codes;  // database sequence of the result record (reference us_prepare_sequence semantics)
    std::string cigar;              // empty for COMPUTE_SCORE
    size_t q_start = 0, q_end = 0, d_start = 0, d_end = 0;
};
void gpu_search( const SearchRequest & req, std::vector<HitAlignment> & out );
void gpu_search_batch( const std::vector<SearchRequest> & reqs, std::vector<std::vector<HitAlignment>> & outs );
size_t gpu_score_all( const uint8_t * query, size_t qlen, int search_type, int64_t * scores, uint64_t * ids );
int db_buffer_max();  // database entries per provider record for the current symbol type / strands (reference adp_init)
void gpu_prepare_db();
void gpu_release();
