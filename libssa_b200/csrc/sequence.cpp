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

// ------------------------------------------------------------------------------------------------
// 6-frame translation (reference src/util/util_sequence.c:120-160 tables, :200-288 table build, :332-382 translate)
// ------------------------------------------------------------------------------------------------

// NCBI genetic codes 1..25 (codon order TCAG x TCAG x TCAG); codes 7, 8 and 17-20 do not exist.
static const char * const genetic_codes[25] = {
    "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",  //  1 standard
    "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNKKSS**VVVVAAAADDEEGGGG",  //  2 vertebrate mitochondrial
    "FFLLSSSSYY**CCWWTTTTPPPPHHQQRRRRIIMMTTTTNNKKSSRRVVVVAAAADDEEGGGG",  //  3 yeast mitochondrial
    "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",  //  4 mold / protozoan mitochondrial
    "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNKKSSSSVVVVAAAADDEEGGGG",  //  5 invertebrate mitochondrial
    "FFLLSSSSYYQQCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",  //  6 ciliate nuclear
    nullptr,
    nullptr,
    "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNNKSSSSVVVVAAAADDEEGGGG",  //  9 echinoderm mitochondrial
    "FFLLSSSSYY**CCCWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",  // 10 euplotid nuclear
    "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",  // 11 bacterial / plastid
    "FFLLSSSSYY**CC*WLLLSPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",  // 12 alternative yeast nuclear
    "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNKKSSGGVVVVAAAADDEEGGGG",  // 13 ascidian mitochondrial
    "FFLLSSSSYYY*CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNNKSSSSVVVVAAAADDEEGGGG",  // 14 alternative flatworm mitochondrial
    "FFLLSSSSYY*QCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",  // 15 blepharisma nuclear
    "FFLLSSSSYY*LCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",  // 16 chlorophycean mitochondrial
    nullptr,
    nullptr,
    nullptr,
    nullptr,
    "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNNKSSSSVVVVAAAADDEEGGGG",  // 21 trematode mitochondrial
    "FFLLSS*SYY*LCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",  // 22 scenedesmus mitochondrial
    "FF*LSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",  // 23 thraustochytrium mitochondrial
    "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSSKVVVVAAAADDEEGGGG",  // 24 pterobranchia mitochondrial
    "FFLLSSSSYY**CCGWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",  // 25 candidate division SR1
};

static uint8_t q_translate[4096];  // query side: IUPAC codon (3 x 4 bit) -> amino-acid code
static uint8_t d_translate[4096];  // database side

// A codon of IUPAC codes translates to the amino acid all of its concrete codons agree on; D/N disagreement
// collapses to B, E/Q to Z, anything else (or a gap) to X.
static void build_translate_table( int code, uint8_t * table ) {
    static const int tcag_of_bit[4] = { 2, 1, 3, 0 };  // bit 0 = A, 1 = C, 2 = G, 3 = T  ->  index in TCAG order
    const char * aas = genetic_codes[code - 1];
    for( int a = 0; a < 16; a++ )
        for( int b = 0; b < 16; b++ )
            for( int c = 0; c < 16; c++ ) {
                char aa = '-';
                for( int i = 0; i < 4; i++ )
                    for( int j = 0; j < 4; j++ )
                        for( int k = 0; k < 4; k++ ) {
                            if( !( ( a >> i ) & 1 ) || !( ( b >> j ) & 1 ) || !( ( c >> k ) & 1 ) ) continue;
                            const char x = aas[tcag_of_bit[i] * 16 + tcag_of_bit[j] * 4 + tcag_of_bit[k]];
                            if( aa == '-' || aa == x )
                                aa = x;
                            else if( aa == 'B' && ( x == 'D' || x == 'N' ) ) {
                            }
                            else if( ( aa == 'D' && ( x == 'B' || x == 'N' ) ) || ( aa == 'N' && ( x == 'B' || x == 'D' ) ) )
                                aa = 'B';
                            else if( aa == 'Z' && ( x == 'Q' || x == 'E' ) ) {
                            }
                            else if( ( aa == 'E' && ( x == 'Z' || x == 'Q' ) ) || ( aa == 'Q' && ( x == 'Z' || x == 'E' ) ) )
                                aa = 'Z';
                            else
                                aa = 'X';
                        }
                if( aa == '-' ) aa = 'X';
                table[256 * a + 16 * b + c] = (uint8_t) map_ncbi_aa[(unsigned char) aa];
            }
}

int g_db_gencode = 1;

void init_translation_tables( int q_gencode, int d_gencode ) {
    g_db_gencode = d_gencode;
    build_translate_table( q_gencode, q_translate );
    build_translate_table( d_gencode, d_translate );
}

namespace {
struct TranslateInit {
    TranslateInit() { init_translation_tables( 1, 1 ); }  // DEFAULT_QUERY_GENCODE / DEFAULT_DB_GENCODE
} translate_init;
}  // namespace

// Protein of reading frame `frame` (0..2) of the forward (strand 0) or complementary strand (any other value);
// length (len - frame) / 3, as in the reference.
void translate_codes( bool db_side, const uint8_t * nt, size_t len, int strand, int frame, std::vector<uint8_t> & out ) {
    const uint8_t * table = db_side ? d_translate : q_translate;
    const size_t plen = len >= (size_t) frame ? ( len - frame ) / 3 : 0;
    out.resize( plen );
    if( strand == 0 ) {
        size_t pos = (size_t) frame;
        for( size_t p = 0; p < plen; p++, pos += 3 ) out[p] = table[( ( nt[pos] & 15 ) << 8 ) | ( ( nt[pos + 1] & 15 ) << 4 ) | ( nt[pos + 2] & 15 )];
    }
    else {
        auto compl4 = []( uint8_t c ) { return (uint8_t) ( ( ( c & 1 ) << 3 ) | ( ( c & 2 ) << 1 ) | ( ( c & 4 ) >> 1 ) | ( ( c & 8 ) >> 3 ) ); };
        size_t pos = len - 1 - (size_t) frame;
        for( size_t p = 0; p < plen; p++, pos -= 3 )
            out[p] = table[( compl4( nt[pos] ) << 8 ) | ( compl4( nt[pos - 1] ) << 4 ) | compl4( nt[pos - 2] )];
    }
}
