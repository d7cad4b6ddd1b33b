// Alphabets and sequence mapping (reference src/util/util_sequence.c:36-88,160-190,298-318,427-437).
//   protein   : "-ABCDEFGHIKLMNPQRSTVWXYZU*OJ" -> codes 0..27; '-' itself is NOT a mapped input symbol
//   nucleotide: "-ACMGRSVTWYHKDBN"             -> codes 0..15; '-' maps to 0 and U maps like T
// Both are case-insensitive; every other byte is "unknown" (-1).  On the database side unknown
// symbols become code 0, on the query side they are dropped (query.cpp).
#include "common.h"

#include <cstring>

const char * sym_ncbi_aa = "-ABCDEFGHIKLMNPQRSTVWXYZU*OJ####";
const char * sym_ncbi_nt16 = "-acmgrsvtwyhkdbn################";

signed char map_ncbi_aa_storage[256];
signed char map_ncbi_nt16_storage[256];

namespace {
struct MapInit {
    MapInit() {
        memset( map_ncbi_aa_storage, -1, 256 );
        memset( map_ncbi_nt16_storage, -1, 256 );
        const char * aa = "-ABCDEFGHIKLMNPQRSTVWXYZU*OJ";
        for( int code = 1; aa[code]; code++ ) {
            unsigned char c = (unsigned char) aa[code];
            map_ncbi_aa_storage[c] = (signed char) code;
            if( c >= 'A' && c <= 'Z' ) map_ncbi_aa_storage[c + 32] = (signed char) code;
        }
        const char * nt = "-ACMGRSVTWYHKDBN";
        for( int code = 0; nt[code]; code++ ) {
            unsigned char c = (unsigned char) nt[code];
            map_ncbi_nt16_storage[c] = (signed char) code;
            if( c >= 'A' && c <= 'Z' ) map_ncbi_nt16_storage[c + 32] = (signed char) code;
        }
        map_ncbi_nt16_storage['U'] = map_ncbi_nt16_storage['T'];
        map_ncbi_nt16_storage['u'] = map_ncbi_nt16_storage['T'];
    }
} map_init;
}  // namespace

int g_symtype = AMINOACID;             // DEFAULT_SYMTYPE, reference src/util/util.h:43
int g_query_strands = FORWARD_STRAND;  // DEFAULT_STRAND, reference src/util/util.h:44

// genetic codes 1..23 known to the reference; 7, 8 and 17-20 do not exist
// (reference src/util/util_sequence.c:94-120, src/libssa.c:138-143)
bool gencode_is_valid( int code ) {
    if( code < 1 || code > 23 ) return false;
    return !( code == 7 || code == 8 || ( code >= 17 && code <= 20 ) );
}

size_t map_db_sequence( const char * ascii, size_t len, uint8_t * out, const signed char * map ) {
    size_t unknown = 0;
    for( size_t i = 0; i < len; i++ ) {
        signed char m = map[(unsigned char) ascii[i]];
        if( m >= 0 ) {
            out[i] = (uint8_t) m;
        }
        else {
            out[i] = 0;
            unknown++;
        }
    }
    return unknown;
}

// complement of a nucleotide code = bit-reversal of its 4-bit IUPAC set (A=1,C=2,G=4,T=8)
void revcompl_codes( const uint8_t * in, size_t len, uint8_t * out ) {
    for( size_t i = 0; i < len; i++ ) {
        uint8_t c = in[len - 1 - i] & 15;
        out[i] = (uint8_t) ( ( ( c & 1 ) << 3 ) | ( ( c & 2 ) << 1 ) | ( ( c & 4 ) >> 1 ) | ( ( c & 8 ) >> 3 ) );
    }
}
