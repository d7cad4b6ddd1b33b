// The libssa.h entry points (reference src/libssa.c:41-271) on top of the GPU search.
// Process-global state and "fatal = exit(1)" error handling are kept exactly as in the reference:
// the library is not re-entrant and one search runs at a time.
#include "common.h"

#include <cstdlib>
#include <cstring>
#include <algorithm>

#define API extern "C" __attribute__( ( visibility( "default" ) ) )

int8_t g_gapO = 0;  // reference src/libssa.c:35-36
int8_t g_gapE = 0;
size_t g_chunk_size = DEFAULT_CHUNK_SIZE;  // reference src/algo/manager.c:40
size_t g_thread_count = (size_t) -1;       // reference src/util/thread_pool.c:37
uint64_t g_db_epoch = 1;

API void set_output_mode( int mode ) { g_output_mode = mode; }

// reference: selects the CPU ISA (src/cpu_config.c:39-51).  There is one GPU code path; accepted and ignored.
API void set_simd_compute_mode( int ) {}

API void set_chunk_size( size_t size ) {
    if( size == 0 ) {
        print_error( "Only non zero chunk sizes are allowed. Using the default size of 1000 sequences.", size );
        size = DEFAULT_CHUNK_SIZE;
    }
    g_chunk_size = size;
}

// accepted; the result is always the reference's deterministic one-thread result
API void set_thread_count( size_t nr ) { g_thread_count = nr; }

API void init_score_matrix( int mode, const char * matrix ) {
    if( mode == READ_FROM_FILE )
        mat_init_from_file( matrix );
    else if( mode == READ_FROM_STRING )
        mat_init_from_string( matrix );
    else if( mode == MATRIX_BUILDIN )
        mat_init_buildin( matrix );
    else
        fatal( "Unknown mode for reading score matrices: %d", mode );
}

API void init_gap_penalties( const int8_t gap_open, const int8_t gap_extend ) {
    g_gapO = gap_open;
    g_gapE = gap_extend;
}

// first argument = match score, second = mismatch score (src/libssa.c:115-117 -> src/matrices.c:447)
API void init_constant_scores( const int8_t p, const int8_t m ) { mat_init_constant_scoring( p, m ); }

API void init_db( const char * fasta_db_file ) {
    ssa_db_close();
    ssa_db_init( fasta_db_file );
    g_db_epoch++;
    print_info( "DB read %lu sequences\n", (unsigned long) ssa_db_get_sequence_count() );
}

API void init_symbol_translation( int type, int strands, int d_gencode, int q_gencode ) {
    if( !gencode_is_valid( q_gencode ) ) fatal( "Illegal query genetic code specified." );
    if( !gencode_is_valid( d_gencode ) ) fatal( "Illegal database genetic code specified." );
    // The reference validates the PREVIOUS global values here (src/libssa.c:144-149), i.e. it accepts any
    // argument; we check the arguments themselves.
    if( type < 0 || type > 4 ) fatal( "Illegal symbol type specified." );
    if( strands < 1 || strands > 3 ) fatal( "Illegal strands specified." );
    g_symtype = type;
    g_query_strands = strands;
    init_translation_tables( q_gencode, d_gencode );  // reference src/libssa.c:156
    g_db_epoch++;
}

API p_query init_sequence_fasta( int mode, const char * fasta_sequence ) {
    if( mode == READ_FROM_FILE ) return query_read_from_file( fasta_sequence );
    if( mode == READ_FROM_STRING ) return query_read_from_string( fasta_sequence );
    fatal( "Unknown mode for reading query sequences: %d", mode );
}

API void free_sequence( p_query p ) { query_free( p ); }

static void test_configuration( p_query query ) {  // reference src/libssa.c:196-219
    if( !g_gapO && !g_gapE ) print_warning( "Gap opening and gap extension cost set to zero. Possible error." );
    if( !g_matrix.set ) fatal( "Scoring not initialized." );
    if( !query ) fatal( "Query not initialized." );
    if( ssa_db_get_sequence_count() == 0 ) print_warning( "Database contains zero sequences. Possible error." );
    if( !is_constant_scoring() && g_symtype == NUCLEOTIDE )
        fatal( "Nucleotide sequences can only be aligned using constant scores." );
}

// The query frames a search runs over (reference src/algo/searcher.c:42-90).
static std::vector<QueryFrame> query_frames( p_query query ) {
    std::vector<QueryFrame> frames;
    auto add = [&]( const sequence_t & s, int strand, int frame ) {
        frames.push_back( QueryFrame{ (const uint8_t *) s.seq, s.len, strand, frame } );
    };
    if( g_symtype == NUCLEOTIDE ) {
        for( int s = 0; s < 2; s++ )
            if( ( s + 1 ) & g_query_strands ) add( query->nt[s], s, 0 );
    }
    else if( g_symtype == AMINOACID || g_symtype == TRANS_DB ) {
        add( query->aa[0], 0, 0 );
    }
    else {
        for( int s = 0; s < 2; s++ )
            if( ( s + 1 ) & g_query_strands )
                for( int f = 0; f < 3; f++ ) add( query->aa[3 * s + f], s, f );
    }
    for( const QueryFrame & f : frames )
        if( !f.seq ) fatal( "Query was not read with the current symbol type / strands (re-read it after init_symbol_translation)." );
    return frames;
}

static SearchRequest make_request( p_query query, size_t hitcount, int bit_width, int align_type, int search_type ) {
    test_configuration( query );
    if( bit_width != BIT_WIDTH_64 && bit_width != BIT_WIDTH_16 && bit_width != BIT_WIDTH_8 )
        fatal( "\nunknown bit width provided: %d\n\n", bit_width );  // reference src/algo/searcher.c:115-117
    SearchRequest req;
    req.frames = query_frames( query );
    req.search_type = search_type;
    req.bit_width = bit_width;
    req.hitcount = hitcount;
    req.align_type = align_type;
    return req;
}

// result records: reference src/algo/aligner.c:62-93 (init_alignment), src/algo/manager.c:83-113
static p_alignment_list make_list( const SearchRequest & req, const std::vector<HitAlignment> & hits ) {
    p_alignment_list alist = (p_alignment_list) xmalloc( sizeof( alignment_list_t ) );
    alist->len = hits.size();
    alist->alignments = (p_alignment *) xmalloc( ( hits.size() ? hits.size() : 1 ) * sizeof( p_alignment ) );
    for( size_t i = 0; i < hits.size(); i++ ) {
        const HitAlignment & h = hits[i];
        const QueryFrame & qf = req.frames[h.query_id];
        p_alignment a = (p_alignment) xmalloc( sizeof( alignment_t ) );
        a->db_seq.seq = (char *) xmalloc( h.db_codes.size() + 1 );
        memcpy( a->db_seq.seq, h.db_codes.data(), h.db_codes.size() );
        a->db_seq.seq[h.db_codes.size()] = 0;
        a->db_seq.len = h.db_codes.size();
        a->db_seq.ID = h.hit.db_id;
        a->db_seq.strand = h.db_frame;   // the reference fills these two swapped (src/algo/aligner.c:76-77)
        a->db_seq.frame = h.db_strand;
        a->query.seq = (char *) qf.seq;  // borrowed, not freed by free_alignment (aligner.c:80,118)
        a->query.len = qf.len;
        a->query.strand = qf.strand;
        a->query.frame = qf.frame;
        a->score = (long) h.hit.score;
        a->alignment = nullptr;
        a->alignment_len = 0;  // left uninitialised by the reference for COMPUTE_SCORE; we define it as 0
        a->align_q_start = a->align_q_end = a->align_d_start = a->align_d_end = 0;
        if( req.align_type != COMPUTE_SCORE ) {
            a->alignment = (char *) xmalloc( h.cigar.size() + 1 );
            memcpy( a->alignment, h.cigar.c_str(), h.cigar.size() + 1 );
            a->alignment_len = h.cigar.size();
            a->align_q_start = h.q_start;
            a->align_q_end = h.q_end;
            a->align_d_start = h.d_start;
            a->align_d_end = h.d_end;
        }
        alist->alignments[i] = a;
    }
    return alist;
}

static p_alignment_list run_search( p_query query, size_t hitcount, int bit_width, int align_type, int search_type ) {
    SearchRequest req = make_request( query, hitcount, bit_width, align_type, search_type );
    std::vector<HitAlignment> hits;
    gpu_search( req, hits );
    return make_list( req, hits );
}

// Batched entry point (include/libssa_b200.h): out[i] is what sw_align / nw_align( queries[i], ... ) returns.
API void ssa_b200_align_batch( int nw, const p_query * queries, size_t nqueries, size_t hitcount, int bit_width, int align_type,
        p_alignment_list * out ) {
    std::vector<SearchRequest> reqs;
    reqs.reserve( nqueries );
    for( size_t i = 0; i < nqueries; i++ )
        reqs.push_back( make_request( queries[i], hitcount, bit_width, align_type, nw ? NEEDLEMAN_WUNSCH : SMITH_WATERMAN ) );
    std::vector<std::vector<HitAlignment>> hits;
    gpu_search_batch( reqs, hits );
    for( size_t i = 0; i < nqueries; i++ ) out[i] = make_list( reqs[i], hits[i] );
}

API p_alignment_list sw_align( p_query query, size_t hitcount, int bit_width, int align_type ) {
    return run_search( query, hitcount, bit_width, align_type, SMITH_WATERMAN );
}

API p_alignment_list nw_align( p_query query, size_t hitcount, int bit_width, int align_type ) {
    return run_search( query, hitcount, bit_width, align_type, NEEDLEMAN_WUNSCH );
}

API void free_alignment( p_alignment_list alist ) {  // reference src/algo/aligner.c:101-143
    if( !alist ) return;
    if( alist->alignments ) {
        for( size_t i = 0; i < alist->len; i++ ) {
            p_alignment a = alist->alignments[i];
            if( !a ) continue;
            free( a->db_seq.seq );
            free( a->alignment );
            free( a );
        }
        free( alist->alignments );
    }
    free( alist );
}

API void ssa_exit( void ) {  // reference src/libssa.c:266-271
    mat_free();
    ssa_db_close();
    g_db_epoch++;
    gpu_release();
}

API size_t ssa_b200_score_all( p_query query, int nw, int64_t * scores, uint64_t * ids ) {
    test_configuration( query );
    std::vector<QueryFrame> frames = query_frames( query );
    return gpu_score_all( frames[0].seq, frames[0].len, nw ? NEEDLEMAN_WUNSCH : SMITH_WATERMAN, scores, ids );
}

// ---- host-only helpers of the hit selection (no GPU needed): used by the multi-process merge and its CPU tests ----
API size_t ssa_b200_host_arrival_order( const uint64_t * lens, size_t n, size_t chunk, int bit_width, const int64_t * ovf8,
        const int64_t * ovf16, uint32_t * order_out ) {
    std::vector<uint64_t> l( lens, lens + n );
    std::vector<uint32_t> order;
    arrival_order( l, ovf8, ovf16, chunk, bit_width, order );
    for( size_t i = 0; i < order.size(); i++ ) order_out[i] = order[i];
    return order.size();
}

API size_t ssa_b200_host_replay( const int32_t * scores_in_arrival_order, const uint32_t * order, size_t n, size_t hitcount,
        uint64_t * ids_out, int64_t * scores_out ) {
    std::vector<Hit> hits;
    replay_topk( scores_in_arrival_order, order, n, hitcount, hits );
    for( size_t i = 0; i < hits.size(); i++ ) {
        ids_out[i] = hits[i].db_id;
        scores_out[i] = hits[i].score;
    }
    return hits.size();
}

// the sparse form of the same replay (replay_topk_sparse), candidates found by a host scan: must equal ssa_b200_host_replay
API size_t ssa_b200_host_replay_sparse( const int32_t * scores_in_arrival_order, const uint32_t * order, size_t n, size_t hitcount,
        size_t prefix, uint64_t * ids_out, int64_t * scores_out ) {
    std::vector<Hit> hits;
    const size_t m = std::min( prefix, n );
    std::vector<int32_t> cs;
    std::vector<uint32_t> ci;
    auto fetch = [&]( int32_t thr, const int32_t ** s, const uint32_t ** i ) -> size_t {
        for( size_t a = m; a < n; a++ )
            if( scores_in_arrival_order[a] > thr ) {
                cs.push_back( scores_in_arrival_order[a] );
                ci.push_back( order[a] );
            }
        *s = cs.data();
        *i = ci.data();
        return cs.size();
    };
    replay_topk_sparse( scores_in_arrival_order, order, m, hitcount, fetch, hits );
    for( size_t i = 0; i < hits.size(); i++ ) {
        ids_out[i] = hits[i].db_id;
        scores_out[i] = hits[i].score;
    }
    return hits.size();
}

API size_t ssa_b200_host_frames_replay( const uint64_t * slot_lens, size_t nslots, size_t chunk_slots, int bit_width, int q_count,
        int buffer_max, const int32_t * scores, const int32_t * ovf8, const int32_t * ovf16, size_t hitcount,
        uint32_t * slots_out, uint8_t * frames_out, int64_t * scores_out ) {
    std::vector<uint64_t> lens( slot_lens, slot_lens + nslots );
    std::vector<std::vector<int32_t>> sc( q_count ), o8( q_count ), o16( q_count );
    for( int q = 0; q < q_count; q++ ) {
        sc[q].assign( scores + (size_t) q * nslots, scores + (size_t) ( q + 1 ) * nslots );
        if( ovf8 ) o8[q].assign( ovf8 + (size_t) q * nslots, ovf8 + (size_t) ( q + 1 ) * nslots );
        if( ovf16 ) o16[q].assign( ovf16 + (size_t) q * nslots, ovf16 + (size_t) ( q + 1 ) * nslots );
    }
    std::vector<FrameHit> hits;
    general_search_order( lens, chunk_slots, bit_width, q_count, buffer_max, sc, o8, o16, hitcount, hits );
    for( size_t i = 0; i < hits.size(); i++ ) {
        slots_out[i] = hits[i].slot;
        frames_out[i] = hits[i].query_id;
        scores_out[i] = hits[i].score;
    }
    return hits.size();
}

// global length-sorted order (descending length, provider index ascending); shard r owns positions r, r+world, ...
API size_t ssa_b200_host_sorted_order( const uint64_t * lens, size_t n, uint32_t * sorted_out ) {
    std::vector<uint32_t> sorted;
    for( size_t i = 0; i < n; i++ )
        if( lens[i] > 0 ) sorted.push_back( (uint32_t) i );
    std::stable_sort( sorted.begin(), sorted.end(), [&]( uint32_t a, uint32_t b ) { return lens[a] > lens[b]; } );
    for( size_t i = 0; i < sorted.size(); i++ ) sorted_out[i] = sorted[i];
    return sorted.size();
}

// the current 32x32 score table, index (db_symbol << 5) + query_symbol; returns 0 if no scoring is set
API int ssa_b200_get_score_matrix( int64_t * out ) {
    if( !g_matrix.set ) return 0;
    memcpy( out, g_matrix.m, sizeof( g_matrix.m ) );
    return 1;
}
