/*
 * Timing driver in the style of the reference's benchmark programs (reference benchmark/src/benchmark_queries.c:32-79,
 * benchmark/src/benchmark_util.c:29-53): for every query file, search type and bit width it times `iterations`
 * repetitions of the *_align call with the database already loaded, and logs one CSV line per combination.
 * The reference logs seconds only and derives GCUPS in R (benchmark/scripts/evaluate_threads.r:111-127); this driver adds
 * the GCUPS column itself: query residues x database residues / median seconds / 1e9.
 * Written against libssa.h + libssa_extern_db.h only, so it links against the reference build or libssa_b200.so.
 *
 *   ssa_benchmark [-M BLOSUM62] [-O -11] [-E -1] [-c 10] [-n 10] [-b 8,16] [-t SW,NW] [-N threads] -d db.fasta query.fasta...
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <unistd.h>

#include "libssa.h"
#include "libssa_extern_db.h"

static double now_s( void ) {
    struct timeval t;
    gettimeofday( &t, NULL );
    return (double) t.tv_sec + (double) t.tv_usec / 1e6;
}

static int cmp_double( const void * a, const void * b ) {
    double x = *(const double *) a, y = *(const double *) b;
    return ( x > y ) - ( x < y );
}

/* residues of the first record of a FASTA file, counted the way the query reader does for proteins: letters only */
static size_t query_length( const char * path ) {
    FILE * f = fopen( path, "r" );
    if( !f ) return 0;
    size_t n = 0;
    int c, header = 0, records = 0;
    while( ( c = fgetc( f ) ) != EOF ) {
        if( c == '>' ) {
            if( ++records > 1 ) break;
            header = 1;
        }
        else if( c == '\n' )
            header = 0;
        else if( !header && ( ( c >= 'A' && c <= 'Z' ) || ( c >= 'a' && c <= 'z' ) || c == '*' ) )
            n++;
    }
    fclose( f );
    return n;
}

int main( int argc, char ** argv ) {
    const char * matrix = BLOSUM62, * db_file = NULL, * widths = "8,16", * types = "SW";
    int gap_open = -11, gap_extend = -1, iterations = 10;
    long hits = 10, threads = 0;
    int c;
    while( ( c = getopt( argc, argv, "M:O:E:c:n:b:t:N:d:" ) ) != -1 ) {
        switch( c ) {
        case 'M': matrix = optarg; break;
        case 'O': gap_open = atoi( optarg ); break;
        case 'E': gap_extend = atoi( optarg ); break;
        case 'c': hits = atol( optarg ); break;
        case 'n': iterations = atoi( optarg ); break;
        case 'b': widths = optarg; break;
        case 't': types = optarg; break;
        case 'N': threads = atol( optarg ); break;
        case 'd': db_file = optarg; break;
        default: db_file = NULL; optind = argc; break;
        }
    }
    if( !db_file || optind >= argc || iterations < 1 ) {
        fprintf( stderr, "usage: %s [-M matrix] [-O gapO] [-E gapE] [-c hits] [-n iterations] [-b 8,16,64] [-t SW,NW] "
                         "[-N threads] -d db.fasta query.fasta...\n", argv[0] );
        return EXIT_FAILURE;
    }

    set_output_mode( OUTPUT_SILENT );
    init_score_matrix( MATRIX_BUILDIN, matrix );
    init_gap_penalties( (int8_t) gap_open, (int8_t) gap_extend );
    init_symbol_translation( AMINOACID, FORWARD_STRAND, 3, 3 );
    if( threads > 0 ) set_thread_count( (size_t) threads );
    double t0 = now_s();
    init_db( db_file );
    size_t records = ssa_db_get_sequence_count(), residues = 0;
    for( size_t i = 0; i < records; i++ ) residues += ssa_db_get_sequence( i )->seqlen;
    fprintf( stderr, "database: %ld records, %ld residues, loaded in %.2f s\n", (long) records, (long) residues, now_s() - t0 );

    double * secs = (double *) malloc( sizeof( double ) * (size_t) iterations );
    printf( "query,qlen,type,bit_width,iterations,first_call_s,median_s,min_s,gcups_median,top_score,top_id\n" );
    for( int q = optind; q < argc; q++ ) {
        p_query query = init_sequence_fasta( READ_FROM_FILE, argv[q] );
        if( !query ) continue;
        size_t qlen = query_length( argv[q] );
        for( const char * t = types; t && *t; t = strchr( t, ',' ) ? strchr( t, ',' ) + 1 : NULL ) {
            int nw = strncmp( t, "NW", 2 ) == 0;
            for( const char * b = widths; b && *b; b = strchr( b, ',' ) ? strchr( b, ',' ) + 1 : NULL ) {
                int bits = atoi( b );
                double first = now_s();
                p_alignment_list r = nw ? nw_align( query, (size_t) hits, bits, COMPUTE_SCORE ) : sw_align( query, (size_t) hits, bits, COMPUTE_SCORE );
                first = now_s() - first;
                long top_score = r->len ? r->alignments[0]->score : 0, top_id = r->len ? (long) r->alignments[0]->db_seq.ID : -1;
                free_alignment( r );
                for( int i = 0; i < iterations; i++ ) {
                    double s = now_s();
                    free_alignment( nw ? nw_align( query, (size_t) hits, bits, COMPUTE_SCORE ) : sw_align( query, (size_t) hits, bits, COMPUTE_SCORE ) );
                    secs[i] = now_s() - s;
                }
                qsort( secs, (size_t) iterations, sizeof( double ), cmp_double );
                double median = iterations % 2 ? secs[iterations / 2] : 0.5 * ( secs[iterations / 2 - 1] + secs[iterations / 2] );
                printf( "%s,%ld,%s,%d,%d,%.6f,%.6f,%.6f,%.1f,%ld,%ld\n", argv[q], (long) qlen, nw ? "NW" : "SW", bits, iterations,
                        first, median, secs[0], (double) qlen * (double) residues / median / 1e9, top_score, top_id );
                fflush( stdout );
            }
        }
        free_sequence( query );
    }
    free( secs );
    ssa_exit();
    return 0;
}
