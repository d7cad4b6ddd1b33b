/*
 * Command line search tool with the option letters of the reference's example program
 * (reference src/libssa_example.c:44-140: -N -O -E -M -i -d -c -t -b -s), so scripts written for it keep working.
 * Written against libssa.h only: it links unchanged against the reference's libssa.a + libsdb.a or against
 * libssa_b200.so.  Additions: -a prints region and CIGAR of every hit, -y selects the symbol type, -r the strands.
 *
 *   ssa_cli -N 1 -O -3 -E -1 -M BLOSUM50 -c 10 -i query.fasta -d db.fasta -t SW -b 16 -s AVX2
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "libssa.h"

static void usage( const char * prog ) {
    fprintf( stderr,
            "Usage: %s [options]\n"
            "  -N <n>     worker threads of a CPU build (accepted; the GPU build reproduces the one-thread result)\n"
            "  -O <n>     gap open penalty, negative (e.g. -11)\n"
            "  -E <n>     gap extension penalty, negative (e.g. -1)\n"
            "  -M <name>  built-in matrix: BLOSUM45|50|62|80|90, PAM30|70|250\n"
            "  -m <a,b>   constant match,mismatch scores instead of a matrix (required for nucleotides)\n"
            "  -i <file>  query (FASTA)\n"
            "  -d <file>  database (FASTA)\n"
            "  -c <n>     number of hits to report\n"
            "  -t SW|NW   local or global alignment\n"
            "  -b 8|16|64 bit width the search starts with\n"
            "  -s <isa>   AVX2|SSE41|SSE2 for a CPU build (accepted and ignored by the GPU build)\n"
            "  -y <type>  0 nucleotide, 1 amino acid (default), 2 translate query, 3 translate db, 4 translate both\n"
            "  -r <n>     strands: 1 forward (default), 2 complementary, 3 both\n"
            "  -a         compute alignments (region + CIGAR) instead of scores only\n",
            prog );
    exit( EXIT_FAILURE );
}

int main( int argc, char ** argv ) {
    const char * query_file = NULL, * db_file = NULL, * matrix = NULL, * constant = NULL, * type = NULL, * isa = "AVX2";
    int gap_open = 0, gap_extend = 0, bits = -1, symtype = AMINOACID, strands = FORWARD_STRAND, align = COMPUTE_SCORE;
    long hits = 0, threads = 0;

    if( argc == 1 ) usage( argv[0] );
    int c;
    while( ( c = getopt( argc, argv, "N:O:E:M:m:i:d:c:t:b:s:y:r:a" ) ) != -1 ) {
        switch( c ) {
        case 'N': threads = atol( optarg ); break;
        case 'O': gap_open = atoi( optarg ); break;
        case 'E': gap_extend = atoi( optarg ); break;
        case 'M': matrix = optarg; break;
        case 'm': constant = optarg; break;
        case 'i': query_file = optarg; break;
        case 'd': db_file = optarg; break;
        case 'c': hits = atol( optarg ); break;
        case 't': type = optarg; break;
        case 'b': bits = atoi( optarg ); break;
        case 's': isa = optarg; break;
        case 'y': symtype = atoi( optarg ); break;
        case 'r': strands = atoi( optarg ); break;
        case 'a': align = COMPUTE_ALIGNMENT; break;
        default: usage( argv[0] );
        }
    }
    if( !query_file || !db_file || !type || bits < 0 || ( !matrix && !constant ) ) usage( argv[0] );
    int nw;
    if( strcmp( type, "SW" ) == 0 )
        nw = 0;
    else if( strcmp( type, "NW" ) == 0 )
        nw = 1;
    else {
        fprintf( stderr, "Unknown search type, try --help.\n" );
        return EXIT_FAILURE;
    }

    set_output_mode( OUTPUT_INFO );
    set_chunk_size( 1000 );
    if( threads > 0 ) set_thread_count( (size_t) threads );
    if( strcmp( isa, "AVX2" ) == 0 )
        set_simd_compute_mode( COMPUTE_ON_AVX2 );
    else if( strcmp( isa, "SSE41" ) == 0 )
        set_simd_compute_mode( COMPUTE_ON_SSE41 );
    else if( strcmp( isa, "SSE2" ) == 0 )
        set_simd_compute_mode( COMPUTE_ON_SSE2 );
    else {
        fprintf( stderr, "Unknown SIMD option, try --help.\n" );
        return EXIT_FAILURE;
    }
    init_symbol_translation( symtype, strands, 3, 3 );  /* genetic code 3 on both sides, as the reference's tool */
    if( constant ) {
        int match = 0, mismatch = 0;
        if( sscanf( constant, "%d,%d", &match, &mismatch ) != 2 ) usage( argv[0] );
        init_constant_scores( (int8_t) match, (int8_t) mismatch );
    }
    else {
        init_score_matrix( MATRIX_BUILDIN, matrix );
    }
    init_gap_penalties( (int8_t) gap_open, (int8_t) gap_extend );
    init_db( db_file );
    p_query query = init_sequence_fasta( READ_FROM_FILE, query_file );
    if( !query ) return EXIT_FAILURE;

    p_alignment_list list = nw ? nw_align( query, (size_t) hits, bits, align ) : sw_align( query, (size_t) hits, bits, align );

    /* same two lines as the reference tool prints (src/libssa_example.c:128-140) */
    printf( "Nr of alignments: %ld\n", (long) list->len );
    printf( "(Score, DB-ID), " );
    for( size_t i = 0; i < list->len; i++ ) printf( "(%ld, %ld), ", list->alignments[i]->score, (long) list->alignments[i]->db_seq.ID );
    printf( "\n" );
    if( align == COMPUTE_ALIGNMENT )
        for( size_t i = 0; i < list->len; i++ ) {
            p_alignment a = list->alignments[i];
            printf( "%ld\t%ld\tq %ld-%ld\td %ld-%ld\t%s\n", a->score, (long) a->db_seq.ID, (long) a->align_q_start,
                    (long) a->align_q_end, (long) a->align_d_start, (long) a->align_d_end, a->alignment );
        }
    free_alignment( list );
    free_sequence( query );
    ssa_exit();
    return 0;
}
