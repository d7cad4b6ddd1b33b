// Default database provider: a FASTA store for the four symbols the library imports
// (include/libssa_extern_db.h; reference src/libssa_extern_db.h:30-47).  The reference delegates this
// to the author's separate libsdb; its observable contract is pinned by the reference's
// tests/test_libssa_extern_db.c:12-58: record i has ID i, sequences are raw ASCII with line breaks
// removed, NULL past the end, close() tolerates "nothing open".
// The definitions are WEAK so a host program may link another provider over them.
#include "common.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {
struct Store {
    std::vector<char> text;         // all residues, NUL after each record
    std::vector<size_t> offset;     // start of record i in text
    std::vector<seqinfo_t> info;
};
Store store;
}  // namespace

extern "C" {

__attribute__( ( weak, visibility( "default" ) ) ) void ssa_db_close( void ) {
    store.text.clear();
    store.text.shrink_to_fit();
    store.offset.clear();
    store.info.clear();
}

__attribute__( ( weak, visibility( "default" ) ) ) int ssa_db_init( const char * db_name ) {
    ssa_db_close();
    FILE * fp = fopen( db_name, "rb" );
    if( !fp ) {
        print_error( "Cannot open database file: %s", db_name );
        return -1;
    }
    // Line by line over 4 MB blocks (a partial last line is carried into the next block): a line that starts with '>' opens a
    // record, every other line of a record is appended without its line break, carriage returns, blanks and tabs.
    std::vector<char> buf( 4u << 20 );
    std::vector<size_t> lens;
    if( fseek( fp, 0, SEEK_END ) == 0 ) {   // the residues cannot outgrow the file: one allocation instead of repeated growth
        const long bytes = ftell( fp );
        if( bytes > 0 ) store.text.reserve( (size_t) bytes + 1 );
        rewind( fp );
    }
    bool in_record = false;
    size_t have = 0;   // bytes in buf: [0, have)
    auto take_line = [&]( const char * p, size_t n ) {
        if( n > 0 && p[0] == '>' ) {
            if( in_record ) {
                lens.push_back( store.text.size() - store.offset.back() );
                store.text.push_back( 0 );
            }
            store.offset.push_back( store.text.size() );
            in_record = true;
            return;
        }
        if( !in_record || n == 0 ) return;
        const size_t at = store.text.size();
        store.text.resize( at + n );
        char * out = store.text.data() + at;
        size_t m = 0;
        for( size_t i = 0; i < n; i++ ) {
            const char c = p[i];
            out[m] = c;
            m += ( c != '\r' && c != ' ' && c != '\t' );
        }
        store.text.resize( at + m );
    };
    while( true ) {
        if( have == buf.size() ) buf.resize( 2 * buf.size() );   // a line longer than the block
        const size_t got = fread( buf.data() + have, 1, buf.size() - have, fp );
        have += got;
        const bool eof = got == 0;
        size_t pos = 0;
        while( pos < have ) {
            const char * nl = (const char *) memchr( buf.data() + pos, '\n', have - pos );
            if( !nl ) break;
            take_line( buf.data() + pos, (size_t) ( nl - ( buf.data() + pos ) ) );
            pos = (size_t) ( nl - buf.data() ) + 1;
        }
        if( eof ) {
            if( pos < have ) take_line( buf.data() + pos, have - pos );   // last line without a line break
            break;
        }
        memmove( buf.data(), buf.data() + pos, have - pos );
        have -= pos;
    }
    fclose( fp );
    if( in_record ) {
        lens.push_back( store.text.size() - store.offset.back() );
        store.text.push_back( 0 );
    }
    store.info.resize( store.offset.size() );
    for( size_t i = 0; i < store.offset.size(); i++ ) {
        store.info[i].ID = i;
        store.info[i].seqlen = lens[i];
        store.info[i].seq = store.text.data() + store.offset[i];
    }
    return 0;
}

__attribute__( ( weak, visibility( "default" ) ) ) size_t ssa_db_get_sequence_count( void ) { return store.info.size(); }

__attribute__( ( weak, visibility( "default" ) ) ) p_seqinfo ssa_db_get_sequence( size_t id ) {
    if( id >= store.info.size() ) return nullptr;
    return &store.info[id];
}

}  // extern "C"
