/*
 * Minimal C host program written against libssa.h only - it compiles and links unchanged against the reference's
 * libssa.a + libsdb.a or against libssa_b200.so (drop-in check; counterpart of reference src/libssa_example.c:144-196).
 *
 *   search_example <db.fasta> <query.fasta> [sw|nw] [bit_width] [hitcount] [score|align] [matrix] [gapO] [gapE]
 *
 * Prints one line per hit: score, database ID, region and CIGAR.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "libssa.h"

int main( int argc, char ** argv ) {
    if( argc < 3 ) {
        fprintf( stderr, "usage: %s db.fasta query.fasta [sw|nw] [8|16|64] [hitcount] [score|align] [matrix] [gapO] [gapE]\n", argv[0] );
        return 2;
    }
    const char * mode = argc > 3 ? argv[3] : "sw";
    int bit_width = argc > 4 ? atoi( argv[4] ) : BIT_WIDTH_16;
    size_t hitcount = argc > 5 ? (size_t) atol( argv[5] ) : 10;
    int align = ( argc > 6 && strcmp( argv[6], "align" ) == 0 ) ? COMPUTE_ALIGNMENT : COMPUTE_SCORE;
    const char * matrix = argc > 7 ? argv[7] : BLOSUM62;
    int gapO = argc > 8 ? atoi( argv[8] ) : -11;
    int gapE = argc > 9 ? atoi( argv[9] ) : -1;

    set_output_mode( OUTPUT_WARNING );
    set_chunk_size( 1000 );
    set_thread_count( 1 );
    init_symbol_translation( AMINOACID, FORWARD_STRAND, 1, 1 );
    init_score_matrix( MATRIX_BUILDIN, matrix );
    init_gap_penalties( (int8_t) gapO, (int8_t) gapE );
    init_db( argv[1] );
    p_query query = init_sequence_fasta( READ_FROM_FILE, argv[2] );
    if( !query ) return 1;

    p_alignment_list hits = strcmp( mode, "nw" ) == 0 ? nw_align( query, hitcount, bit_width, align )
                                                       : sw_align( query, hitcount, bit_width, align );
    for( size_t i = 0; i < hits->len; i++ ) {
        p_alignment a = hits->alignments[i];
        printf( "%ld\t%lu\t%lu-%lu\t%lu-%lu\t%s\n", a->score, (unsigned long) a->db_seq.ID, (unsigned long) a->align_q_start,
                (unsigned long) a->align_q_end, (unsigned long) a->align_d_start, (unsigned long) a->align_d_end,
                a->alignment ? a->alignment : "-" );
    }
    free_alignment( hits );
    free_sequence( query );
    ssa_exit();
    return 0;
}
