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
    std::vector<char> buf( 1 << 20 );
    std::vector<size_t> lens;
    bool in_header = false, in_record = false, line_start = true;
    size_t got;
    while( ( got = fread( buf.data(), 1, buf.size(), fp ) ) > 0 ) {
        for( size_t i = 0; i < got; i++ ) {
            char c = buf[i];
            if( in_header ) {
                if( c == '\n' ) {
                    in_header = false;
                    line_start = true;
                }
                continue;
            }
            if( line_start && c == '>' ) {
                if( in_record ) {
                    lens.push_back( store.text.size() - store.offset.back() );
                    store.text.push_back( 0 );
                }
                store.offset.push_back( store.text.size() );
                in_record = true;
                in_header = true;
                continue;
            }
            line_start = ( c == '\n' );
            if( !in_record || c == '\n' || c == '\r' || c == ' ' || c == '\t' ) continue;
            store.text.push_back( c );
        }
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
