// Query reading (reference src/query.c:42-252).  A query is the FIRST record of a FASTA file (the
// '>' header is optional) or a raw sequence string; symbols that the current alphabet does not know
// are REMOVED (query.c:102-130) - unlike database sequences, where they become code 0.
// The alphabet is fixed when the query is read, from the symbol type set at that moment.
#include "common.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static p_query new_query() {
    p_query q = (p_query) xmalloc( sizeof( struct _query ) );
    memset( q, 0, sizeof( struct _query ) );
    if( g_symtype == AMINOACID || g_symtype == TRANS_DB ) {
        q->map = (const char *) map_ncbi_aa;
        q->sym = sym_ncbi_aa;
    }
    else {
        q->map = (const char *) map_ncbi_nt16;
        q->sym = sym_ncbi_nt16;
    }
    return q;
}

void query_free( p_query q ) {
    if( !q ) return;
    free( q->header );
    for( int s = 0; s < 2; s++ ) free( q->nt[s].seq );
    for( int f = 0; f < 6; f++ ) free( q->aa[f].seq );
    free( q );
}

static void fill_and_map( p_query q, const char * text, size_t len ) {
    sequence_t mapped = { (char *) xmalloc( len + 1 ), 0 };
    std::string unknown;
    for( size_t i = 0; i < len; i++ ) {
        signed char m = q->map[(unsigned char) text[i]];
        if( m >= 0 )
            mapped.seq[mapped.len++] = m;
        else if( text[i] != '\n' && text[i] != ' ' && text[i] != '\t' )
            unknown.push_back( text[i] );
    }
    if( !unknown.empty() )
        print_warning( "%ld unknown symbols found and removed: '%s'", (long) unknown.size(), unknown.c_str() );
    mapped.seq[mapped.len] = 0;

    if( g_symtype == NUCLEOTIDE ) {
        q->nt[0] = mapped;
        if( g_query_strands & 2 ) {
            q->nt[1] = { (char *) xmalloc( mapped.len + 1 ), mapped.len };
            revcompl_codes( (const uint8_t *) mapped.seq, mapped.len, (uint8_t *) q->nt[1].seq );
            q->nt[1].seq[mapped.len] = 0;
        }
    }
    else if( g_symtype == TRANS_QUERY || g_symtype == TRANS_BOTH ) {
        // reading frames of the selected strands (reference src/query.c:146-159); the nucleotide string itself is dropped
        std::vector<uint8_t> prot;
        for( int s = 0; s < 2; s++ ) {
            if( !( ( s + 1 ) & g_query_strands ) ) continue;
            for( int f = 0; f < 3; f++ ) {
                translate_codes( false, (const uint8_t *) mapped.seq, mapped.len, s, f, prot );
                sequence_t & dst = q->aa[3 * s + f];
                dst.seq = (char *) xmalloc( prot.size() + 1 );
                memcpy( dst.seq, prot.data(), prot.size() );
                dst.seq[prot.size()] = 0;
                dst.len = prot.size();
            }
        }
        free( mapped.seq );
    }
    else {
        q->aa[0] = mapped;
    }
}

p_query query_read_from_string( const char * sequence ) {
    p_query q = new_query();
    fill_and_map( q, sequence, strlen( sequence ) );
    return q;
}

p_query query_read_from_file( const char * filename ) {
    if( strcmp( filename, "-" ) == 0 ) {
        print_error( "Query not specified" );
        return nullptr;
    }
    FILE * fp = fopen( filename, "r" );
    if( !fp ) {
        print_error( "Cannot open query file: %s", filename );
        return nullptr;
    }
    char * line = nullptr;
    size_t cap = 0;
    ssize_t n = getline( &line, &cap, fp );
    if( n < 0 ) {
        print_error( "Could not initialise from query sequence" );
        free( line );
        fclose( fp );
        return nullptr;
    }
    p_query q = new_query();
    std::string text;
    bool have_line = true;
    if( line[0] == '>' ) {
        size_t len = (size_t) n;
        if( len && line[len - 1] == '\n' ) len--;
        q->header = (char *) xmalloc( len + 1 );
        memcpy( q->header, line + 1, len - 1 );
        q->header[len - 1] = 0;
        q->headerlen = len - 1;
        n = getline( &line, &cap, fp );
        if( n < 0 ) {
            print_error( "Could not read first line from query sequence" );
            query_free( q );
            free( line );
            fclose( fp );
            return nullptr;
        }
    }
    while( have_line && line[0] && line[0] != '>' ) {
        text.append( line, (size_t) n );
        n = getline( &line, &cap, fp );
        have_line = n >= 0;
    }
    free( line );
    fclose( fp );
    fill_and_map( q, text.data(), text.size() );
    return q;
}
