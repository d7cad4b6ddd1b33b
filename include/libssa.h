/*
 * libssa.h - the public C API of the libssa hot path, served by the B200-native
 * library libssa_b200.so.
 *
 * This header restates, declaration for declaration, the interface of the
 * reference (reference src/libssa.h:28-263) so that C host code written
 * against libssa links against this library unchanged.  Constants, struct
 * layouts and function signatures are ABI; the reference line each one
 * replaces is cited.  Everything below `m_run` in the reference
 * (src/algo/manager.c:120-185: SSE/AVX2 kernels, db_adapter chunking, pthread
 * dispatch, min-heap merge, aligner) is replaced by CUDA kernels for sm_100a.
 */
#ifndef LIBSSA_H_
#define LIBSSA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- constants: reference src/libssa.h:28-72 ---- */
#define VERBOSE_ON 1
#define VERBOSE_OFF 0

#define SIMD_ON 1
#define SIMD_OFF 0

#define BLOSUM45 "blosum45"
#define BLOSUM50 "blosum50"
#define BLOSUM62 "blosum62"
#define BLOSUM80 "blosum80"
#define BLOSUM90 "blosum90"
#define PAM30 "pam30"
#define PAM70 "pam70"
#define PAM250 "pam250"

#define NUCLEOTIDE 0  /* nucleotide query against a nucleotide DB */
#define AMINOACID 1   /* protein query against a protein DB */
#define TRANS_QUERY 2 /* 6-frame translated nucleotide query against a protein DB */
#define TRANS_DB 3    /* protein query against a 6-frame translated nucleotide DB */
#define TRANS_BOTH 4  /* both sides translated */

#define FORWARD_STRAND 1
#define COMPLEMENTARY_STRAND 2
#define BOTH_STRANDS 3

#define BIT_WIDTH_8 8
#define BIT_WIDTH_16 16
#define BIT_WIDTH_64 64

#define OUTPUT_SILENT 0
#define OUTPUT_ERROR 1
#define OUTPUT_WARNING 2
#define OUTPUT_INFO 3

#define COMPUTE_SCORE 0
#define COMPUTE_ALIGNMENT 1

#define READ_FROM_FILE 0
#define READ_FROM_STRING 1
#define MATRIX_BUILDIN 2

#define COMPUTE_ON_SSE2 0
#define COMPUTE_ON_SSE41 1
#define COMPUTE_ON_AVX2 2

/* ---- data types: reference src/libssa.h:78-117 ---- */
struct _query;
typedef struct _query * p_query;

typedef struct { /* reference src/libssa.h:85-91 */
    char * seq;  /* code string (0..31), NUL-terminated, owned by the result */
    size_t len;
    size_t ID;   /* 0-based record index of the DB provider */
    int strand;
    int frame;
} db_seq_t;

typedef struct { /* reference src/libssa.h:93-98 */
    char * seq;  /* borrowed from the p_query: the query must outlive the result */
    size_t len;
    int strand;
    int frame;
} q_seq_t;

typedef struct { /* reference src/libssa.h:100-110 */
    db_seq_t db_seq;
    q_seq_t query;
    char * alignment;       /* CIGAR ("2MD3MI"), NULL for COMPUTE_SCORE */
    size_t alignment_len;
    long score;
    size_t align_q_start;   /* 0-based, inclusive */
    size_t align_q_end;
    size_t align_d_start;
    size_t align_d_end;
} alignment_t;
typedef alignment_t * p_alignment;

typedef struct { /* reference src/libssa.h:112-116 */
    p_alignment * alignments;
    size_t len;
} alignment_list_t;
typedef alignment_list_t * p_alignment_list;

/* ---- technical initialisation: reference src/libssa.h:122-128, src/libssa.c:41-60 ---- */
void set_output_mode( int mode );
/* accepted for compatibility; the GPU path has a single code path (reference: CPU ISA selection) */
void set_simd_compute_mode( int mode );
/* chunk size of the emulated one-thread reference scheduler: only affects which of several
 * sequences tied at the cut-off score are reported; 0 -> 1000 with an error message */
void set_chunk_size( size_t size );
/* accepted; results always equal the reference's one-thread (deterministic) result */
void set_thread_count( size_t count );

/* ---- initialisations: reference src/libssa.h:133-215, src/libssa.c:80-180 ---- */
void init_score_matrix( int mode, const char * matrix );
/* NOTE: as in the reference implementation (src/libssa.c:115-117 -> src/matrices.c:447) the FIRST
 * argument is the match score and the second the mismatch score */
void init_constant_scores( const int8_t p, const int8_t m );
/* penalties are passed negative; a gap of length k costs gapO + k*gapE */
void init_gap_penalties( const int8_t gapO, const int8_t gapE );
void init_symbol_translation( int type, int strands, int db_gencode, int q_gencode );
void init_db( const char * db_file );
p_query init_sequence_fasta( int mode, const char * fasta_seq_file );
void free_sequence( p_query p );

/* ---- alignment: reference src/libssa.h:220-261, src/libssa.c:229-264 ---- */
p_alignment_list sw_align( p_query p, size_t hitcount, int bit_width, int align_type );
p_alignment_list nw_align( p_query p, size_t hitcount, int bit_width, int align_type );
void free_alignment( p_alignment_list alist );

void ssa_exit( void ); /* reference src/libssa.c:266-271; also releases all device state */

#ifdef __cplusplus
}
#endif

#endif /* LIBSSA_H_ */
