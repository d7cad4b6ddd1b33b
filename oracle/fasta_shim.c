/*
 * TEST INFRASTRUCTURE ONLY - not part of the product.
 *
 * Minimal FASTA store that satisfies the four-symbol provider interface the
 * reference imports (reference: src/libssa_extern_db.h:30-47).  The reference
 * links against the author's separate `libsdb`, which is not vendored; its
 * contract is pinned by the reference's tests/test_libssa_extern_db.c:12-58
 * (IDs are 0-based record indices, NULL past the end, close is idempotent).
 * This shim is linked into oracle/_ref/libssa_ref.so so the unmodified
 * reference can be run as the parity oracle.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    size_t ID;
    size_t seqlen;
    char * seq;
} seqinfo_t;
typedef seqinfo_t * p_seqinfo;

static seqinfo_t * records = NULL;
static size_t record_count = 0;
static size_t record_cap = 0;

static void push_record( char * seq, size_t len ) {
    if( record_count == record_cap ) {
        record_cap = record_cap ? record_cap * 2 : 1024;
        records = realloc( records, record_cap * sizeof(seqinfo_t) );
    }
    seq[len] = 0;
    records[record_count].ID = record_count;
    records[record_count].seqlen = len;
    records[record_count].seq = seq;
    record_count++;
}

void ssa_db_close( void ) {
    for( size_t i = 0; i < record_count; i++ )
        free( records[i].seq );
    free( records );
    records = NULL;
    record_count = record_cap = 0;
}

int ssa_db_init( const char * db_name ) {
    ssa_db_close();
    FILE * fp = fopen( db_name, "r" );
    if( !fp ) {
        fprintf( stderr, "fasta_shim: cannot open %s\n", db_name );
        return -1;
    }
    char * line = NULL;
    size_t cap = 0;
    ssize_t n;
    char * cur = NULL;
    size_t cur_len = 0, cur_cap = 0;
    int in_record = 0;
    while( (n = getline( &line, &cap, fp )) >= 0 ) {
        if( n > 0 && line[0] == '>' ) {
            if( in_record )
                push_record( cur, cur_len );
            cur_cap = 256;
            cur = malloc( cur_cap + 1 );
            cur_len = 0;
            in_record = 1;
            continue;
        }
        if( !in_record )
            continue;
        for( ssize_t i = 0; i < n; i++ ) {
            char c = line[i];
            if( c == '\n' || c == '\r' || c == ' ' || c == '\t' )
                continue;
            if( cur_len == cur_cap ) {
                cur_cap *= 2;
                cur = realloc( cur, cur_cap + 1 );
            }
            cur[cur_len++] = c;
        }
    }
    if( in_record )
        push_record( cur, cur_len );
    free( line );
    fclose( fp );
    return 0;
}

size_t ssa_db_get_sequence_count( void ) {
    return record_count;
}

p_seqinfo ssa_db_get_sequence( size_t id ) {
    if( id >= record_count )
        return NULL;
    return &records[id];
}
