// Substitution matrices (reference src/matrices.c:290-568, src/matrices.h:26-44).
// One 32x32 int64 table, index (db_symbol << 5) + query_symbol, every cell -1 unless set.
// The typed int8/int16 copies of the reference are not needed: the device builds its own
// query profile from this table (search.cu).
#include "common.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <strings.h>

ScoreMatrix g_matrix;
static int constant_scoring = 0;  // sticky, exactly like the reference (matrices.c:298,448)

bool is_constant_scoring() { return constant_scoring != 0; }

static void prepare() {
    for( int i = 0; i < SCORE_MATRIX_DIM * SCORE_MATRIX_DIM; i++ ) g_matrix.m[i] = -1;  // matrices.c:344
    g_matrix.set = true;
}

static void finalize() {
    int64_t lo = g_matrix.m[0], hi = g_matrix.m[0];
    for( int i = 0; i < SCORE_MATRIX_DIM * SCORE_MATRIX_DIM; i++ ) {
        if( g_matrix.m[i] < lo ) lo = g_matrix.m[i];
        if( g_matrix.m[i] > hi ) hi = g_matrix.m[i];
    }
    g_matrix.lo = lo;
    g_matrix.hi = hi;
    g_matrix.version++;
}

struct Parser {
    int symbols = 0;
    signed char order[2048];
};

// One line of an NCBI-format matrix (reference matrices.c:388-437): '#'/blank = comment, a line that
// starts with blank/tab lists the column symbols, any other line is "<row symbol> <scores...>".
static void read_line( const char * line, Parser & ps ) {
    const char * p = line;
    char c = *p++;
    switch( c ) {
    case '\0':
    case '\n':
    case '#':
        break;
    case '\t':
    case ' ': {
        signed char * q = ps.order;
        while( ( c = *p++ ) ) {
            if( c != ' ' && c != '\t' && c != '\n' ) {
                *q++ = map_ncbi_aa[(unsigned char) c];
                ps.symbols++;
            }
        }
        break;
    }
    default: {
        int a = map_ncbi_aa[(unsigned char) c];
        int read = 0;
        for( int i = 0; i < ps.symbols; i++ ) {
            long sc = 0;
            if( sscanf( p, "%ld%n", &sc, &read ) == 0 ) fatal( "Problem parsing score matrix file." );
            int b = ps.order[i];
            if( a >= 0 && b >= 0 && a < SCORE_MATRIX_DIM && b < SCORE_MATRIX_DIM ) g_matrix.m[( a << 5 ) + b] = sc;
            p += read;
        }
        break;
    }
    }
}

void mat_init_constant_scoring( int8_t match, int8_t mismatch ) {
    constant_scoring = 1;
    prepare();
    for( int a = 1; a < SCORE_MATRIX_DIM; a++ )
        for( int b = 1; b < SCORE_MATRIX_DIM; b++ ) g_matrix.m[( a << 5 ) + b] = ( a == b ) ? match : mismatch;
    finalize();
}

void mat_init_from_string( const char * text ) {
    if( !text ) fatal( "Cannot read score matrix string." );
    prepare();
    Parser ps;
    const char * s = text;
    char line[2048];
    while( *s ) {
        const char * nl = strchr( s, '\n' );
        size_t len = nl ? (size_t) ( nl - s ) : strlen( s );
        if( len >= sizeof( line ) ) len = sizeof( line ) - 1;
        memcpy( line, s, len );
        line[len] = 0;
        read_line( line, ps );
        s = nl ? nl + 1 : s + strlen( s );
    }
    finalize();
}

void mat_init_from_file( const char * path ) {
    FILE * fp = fopen( path, "r" );
    if( !fp ) fatal( "Cannot open score matrix file." );
    prepare();
    Parser ps;
    char line[2048];
    while( fgets( line, sizeof( line ), fp ) ) read_line( line, ps );
    fclose( fp );
    finalize();
}

struct Builtin {
    const char * name;
    const char * order;
    int dim;
    signed char values[25 * 25];
};

static const Builtin builtins[] = {
#include "matrix_data.inc"
};

void mat_init_buildin( const char * name ) {
    for( const Builtin & b : builtins ) {
        if( strcasecmp( name, b.name ) == 0 ) {
            prepare();
            for( int r = 0; r < b.dim; r++ ) {
                int a = map_ncbi_aa[(unsigned char) b.order[r]];
                for( int c = 0; c < b.dim; c++ ) {
                    int d = map_ncbi_aa[(unsigned char) b.order[c]];
                    if( a >= 0 && d >= 0 ) g_matrix.m[( a << 5 ) + d] = b.values[r * b.dim + c];
                }
            }
            finalize();
            return;
        }
    }
    fatal( "Unknown matrix: %s", name );
}

void mat_free() {
    g_matrix.set = false;
    g_matrix.version++;
}
