// Scoring kernels for sm_100a: affine-gap Smith-Waterman / Needleman-Wunsch score of one query
// against a stream of database sequences.  Replaces the reference's SIMD kernels
// (src/algo/simd/search_simd_{sw,nw}.c), its profile transposition (src/algo/{8,16}/search_*_util.c),
// the scalar 64-bit tier (src/algo/64/*.c) and the chunk loop around them.
//
// Decomposition:
//   * a warp is split into 32/G independent systolic arrays of G lanes (G = 4..32, chosen per query); lane l of an
//     array owns R consecutive query rows [rowbase + l*R, rowbase + (l+1)*R) whose H (previous column) and E state
//     live in registers for the whole launch;
//   * database columns stream through the lanes, TWO per step: at step t lane l works on column pair t-l and hands
//     the H and F of its last row (both columns) to lane l+1 with __shfl_up_sync - the anti-diagonal wavefront, one
//     shuffle hop per step; the two columns' F chains are independent, which doubles the instruction-level
//     parallelism of the otherwise serial row loop;
//   * the columns of MANY database sequences are concatenated into one stream per work batch; flag bits in the
//     stream mark first/last columns, where a lane resets its state / reports its result.  The wavefront
//     therefore fills and drains once per batch (G-1 steps per >= 256 columns), not once per sequence.  Every lane
//     reads the 32-bit word of ITS column pair straight from the stream, prefetched two steps ahead;
//   * PACKED tier: every 32-bit register holds the same DP cell of TWO database sequences (A in the high, B in the
//     low 16 bits): inter-sequence SIMD-in-word.  All arithmetic is one SASS instruction per packed word:
//     VIADD.16x2 (FMA-side pipe), VIMNMX3.S16x2[.RELU] and VIADDMNMX.S16x2 (ALU pipe, DPX);
//   * substitution scores come from a query profile in shared memory laid out [row quad][residue][lane], so a
//     warp-wide lookup is bank-conflict free whatever residues the lanes hold.  For the first MR rows of a lane an
//     entry is 128 bits - four rows, one ZERO-EXTENDED int16 each - and the (A:B) word of a row is one multiply-add
//     a * 65536 + b on the FMA-side pipe (IMAD); for the remaining rows an entry is 64 bits - four rows, two per
//     word - and a PRMT on the ALU pipe interleaves A and B (see merge_rows);
//   * WIDE tier (PACKED=false): the same systolic array on plain int32 for the (rare) sequences whose scores
//     cannot be guaranteed to fit 16 bits - the counterpart of the reference's 16->64 bit overflow cascade.
//     Optionally tracks the first column whose running maximum reaches 255 / 65535, which is what the
//     reference's 8/16-bit tiers would have overflowed at (needed to reproduce its hit order, replay.cpp);
//   * queries longer than 32*32 rows are processed in passes: full passes of 32 lanes x 32 rows and a remainder pass
//     whose shape (G lanes x R rows, any G) fits the rows that are left; the H/F of the last row of a pass is
//     written to a boundary buffer in global memory and read back by the top lane(s) of the next pass.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace ssa {

constexpr uint32_t COL_FIRST = 1u << 10;   // first column of a sequence pair
constexpr uint32_t COL_LAST_B = 1u << 11;  // last real column of sequence B (low half)
constexpr uint32_t COL_LAST_A = 1u << 12;  // last column of sequence A = last column of the pair
constexpr uint32_t COL_VALID = 1u << 15;   // column to process (words beyond a batch read as 0)
constexpr uint32_t PAD_RESIDUE = 31;       // residue code of the padding column that makes every pair an even number of columns:
                                           // no alphabet uses it, its profile rows hold a non-positive score, so a packed
                                           // Smith-Waterman step never needs a one-column path (the column can only repeat
                                           // values that exist already; the next pair's first column resets the lane)

constexpr int SCORE_THREADS = 256;

struct Batch {
    uint64_t col0;   // first column in the stream
    uint64_t bnd0;   // first entry in the pass-boundary buffers (multi-pass launches)
    uint32_t ncols;  // columns to process
    uint32_t out0;   // PACKED: first pair index; WIDE: output slot (sorted position)
    uint32_t half;   // WIDE: 0 = sequence A (bits 0-4), 1 = sequence B (bits 5-9)
    uint32_t pad;
};

struct ScoreArgs {
    const uint16_t * stream;
    const Batch * batches;
    uint32_t nbatches;
    uint32_t * counter;        // work queue head
    const uint32_t * nbatches_dev;   // if set: the number of batches is read from device memory (work lists built on the device)
    const uint32_t * profile;  // this pass' query profile, laid out for shared memory
    uint32_t group;            // lanes per systolic array: 4, 8, 16 or 32
    uint32_t k65536;           // the constant 65536 (see the profile merge in score_kernel)
    uint32_t qr_word;          // gapO + gapE as the tier's register word (packed: replicated in both halves)
    uint32_t rg_word;          // gapE likewise
    int32_t gap_open_extend;   // gapO + gapE (negative)
    int32_t gap_extend;        // gapE
    int32_t gap_open;          // gapO
    int32_t offset;            // NW: values are stored minus this offset
    int32_t rowbase;           // first query row of this pass
    int32_t cap_lane;          // NW: lane / row-in-lane holding query row qlen-1, -1 if not in this pass
    int32_t cap_row;
    uint32_t first_pass;       // SW: 1 = store, 0 = max-combine with the previous passes
    void * out;                // PACKED: uint32 (A:B), indexed by the stream column pair in which the pair of sequences ends
                               // (bnd0 of a packed batch = its first stream column); WIDE: int32 per slot
    int32_t * first8;          // TRACK: first column reaching thr8 / thr16, per slot (atomicMin)
    int32_t * first16;
    int32_t thr8, thr16;
    // DUAL (two QUERIES per packed word, one database sequence per launch half)
    uint32_t dual_half;        // 0 = the launch scores sequence A of every pair (bits 0-4), 1 = sequence B (bits 5-9)
    uint32_t out_stride;       // words between the result arrays of the two halves
    int32_t cap_lane2, cap_row2;   // NW: lane / row of the SECOND query's last row (cap_lane / cap_row: the first query's)
    int32_t nw_shift, nw_shift2;   // packed NW: virtual rows on top of the (first / second) query, so that the query's last row is
                                   // the LAST row of a lane (search.cu:fill_profile_packed); rowbase counts array positions
    const uint4 * bnd_in;       // boundary row from the previous pass (nullptr in pass 0): one entry per column pair,
                                // { H column A, H column B, F column A, F column B } of the pass' last query row
    uint4 * bnd_out;            // boundary row for the next pass (nullptr in the last pass)
};

template<bool PACKED> struct Arith;

template<> struct Arith<true> {
    typedef uint32_t T;
    static __device__ __forceinline__ T bcast( int v ) { return (uint32_t) (uint16_t) v * 0x10001u; }
    static __device__ __forceinline__ T add( T a, T b ) { return __vadd2( a, b ); }
    static __device__ __forceinline__ T max2( T a, T b ) { return __vmaxs2( a, b ); }
    static __device__ __forceinline__ T max3( T a, T b, T c ) { return __vimax3_s16x2( a, b, c ); }
    static __device__ __forceinline__ T max3_relu( T a, T b, T c ) { return __vimax3_s16x2_relu( a, b, c ); }
    static __device__ __forceinline__ T addmax( T a, T b, T c ) { return __viaddmax_s16x2( a, b, c ); }
};

template<> struct Arith<false> {
    typedef int32_t T;
    static __device__ __forceinline__ T bcast( int v ) { return v; }
    static __device__ __forceinline__ T add( T a, T b ) { return a + b; }
    static __device__ __forceinline__ T max2( T a, T b ) { return max( a, b ); }
    static __device__ __forceinline__ T max3( T a, T b, T c ) { return __vimax3_s32( a, b, c ); }
    static __device__ __forceinline__ T max3_relu( T a, T b, T c ) { return __vimax3_s32_relu( a, b, c ); }
    static __device__ __forceinline__ T addmax( T a, T b, T c ) { return __viaddmax_s32( a, b, c ); }
};

__device__ __forceinline__ uint32_t lds32( uint32_t addr ) {
    uint32_t v;
    asm volatile( "ld.shared.u32 %0, [%1];" : "=r"( v ) : "r"( addr ) );
    return v;
}

__device__ __forceinline__ uint2 lds64( uint32_t addr ) {
    uint2 v;
    asm volatile( "ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"( v.x ), "=r"( v.y ) : "r"( addr ) );
    return v;
}

__device__ __forceinline__ uint4 lds128( uint32_t addr ) {
    uint4 v;
    asm volatile( "ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"( v.x ), "=r"( v.y ), "=r"( v.z ), "=r"( v.w ) : "r"( addr ) );
    return v;
}

// Packed tier: rows of a lane whose profile entries hold one zero-extended int16 per 32-bit word (merged by IMAD);
// the other R - MR rows hold two scores per word (merged by PRMT).  The IMAD form pays where the ALU pipe throttles:
// the shapes with 3-4 resident CTAs per SM (R <= 16: +5 % [B200]).  The 2-CTA shapes (R >= 24) are issue-bound at
// 4 warps per scheduler and gain nothing from it (measured -1..+2 %), so they keep the smaller PRMT profile.
// (SSA_MR_FULL / SSA_MR_HI: build-time knobs of that measurement - `make EXTRA=-DSSA_MR_FULL=24 OUT=... OBJDIR=...` builds a
// variant library that LIBSSA_B200_LIB selects at run time; profiles/prof_r01_shapes.md.)
#ifndef SSA_MR_HI
#define SSA_MR_HI 0
#endif
#ifndef SSA_MR_FULL
#define SSA_MR_FULL 16
#endif
__host__ __device__ constexpr int merge_rows( int R ) { return R <= SSA_MR_FULL ? R : SSA_MR_HI; }

// words of shared memory the profile of one pass needs
template<bool PACKED, int R> __host__ __device__ constexpr int profile_words() {
    return PACKED ? merge_rows( R ) * 1024 + ( ( R - merge_rows( R ) ) / 2 ) * 1024 : R * 1024;
}

// Packed NW: H of array position `pos` in column -1 (query row pos - shift; `shift` virtual rows on top); pos = -1 is the row
// above the array
__host__ __device__ __forceinline__ int nw_start( int pos, int shift, int gap_open, int gap_extend ) {
    if( pos < shift - 1 ) return gap_open;                                // virtual rows (and the boundary row above them): gapO
    if( pos == shift - 1 ) return 0;                                       // last virtual row, or the boundary row above row 0: the corner
    return gap_open + ( pos - shift + 1 ) * gap_extend;
}

// dynamic shared memory of a kernel instance: the profile and, for packed Needleman-Wunsch, the start-value table
template<bool PACKED, int R> __host__ __device__ constexpr int score_smem_bytes( bool nw ) {
    return ( profile_words<PACKED, R>() + ( PACKED && nw ? 2 * R * 32 : 0 ) ) * 4;
}

// resident CTAs per SM the register allocator has to make room for (64K registers / 256 threads / blocks)
template<bool PACKED, int R> __host__ __device__ constexpr int min_blocks() { return PACKED ? ( R <= 12 ? 4 : ( R <= 20 ? 3 : 2 ) ) : 2; }

// MULTI: 0 = the whole query in one pass; 1 = one pass of several, full-warp groups; 2 = one pass of several, narrow groups
// DUAL (packed tier, one pass): the two halves of every word belong to two QUERIES scored against the SAME database
// sequence (sequence A of every pair in one launch, sequence B in another).  Both halves then see the same residue, so a
// profile entry IS the (q1 : q2) word of a row: one LDS.128 delivers four rows ready for use - no merge instruction, half
// the lookups: 5.75 (SW) / 5.25 (NW) instead of 7.5 / 7 instructions per word, 3.5 / 3 of them on the ALU pipe.  Used by
// ssa_b200_align_batch for consecutive queries (the reference takes one query per call, src/libssa.h:240-250).
template<bool PACKED, int R, bool NW, int MULTI, bool TRACK, bool DUAL = false>
__global__ void __launch_bounds__( SCORE_THREADS, min_blocks<PACKED, R>() ) score_kernel( const ScoreArgs a ) {
    typedef Arith<PACKED> A;
    typedef typename A::T T;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ uint32_t prof[];
    static_assert( !DUAL || ( PACKED && MULTI == 0 && !TRACK ), "DUAL: packed tier, one pass" );
    constexpr int PWORDS = DUAL ? R * 1024 : profile_words<PACKED, R>();

    for( int i = threadIdx.x; i < PWORDS; i += blockDim.x ) prof[i] = a.profile[i];
    // Packed Needleman-Wunsch: start values of every lane's rows - H(i, -1) = gapO + (i + 1) gapE and the E entering column 0,
    // H + gapO + gapE, both minus the offset - as 128-bit entries [H | E][row quad][lane] behind the profile: a sequence start
    // costs a lane R/2 shared loads instead of 3R additions (and the WARP runs that code once per lane and sequence pair)
    constexpr int NW_TAB = PWORDS;
    if( NW && PACKED ) {
        for( int i = threadIdx.x; i < R * 32; i += blockDim.x ) {
            const int r = ( i >> 7 ) * 4 + ( i & 3 ), l = ( i >> 2 ) & 31;
            const int pos = a.rowbase + ( l & ( (int) a.group - 1 ) ) * R + r;
            const int h1 = nw_start( pos, a.nw_shift, a.gap_open, a.gap_extend ) - a.offset;
            const int h2 = DUAL ? nw_start( pos, a.nw_shift2, a.gap_open, a.gap_extend ) - a.offset : h1;
            const int qr = a.gap_open + a.gap_extend;
            prof[NW_TAB + i] = ( (uint32_t) (uint16_t) (int16_t) h1 << 16 ) | (uint32_t) (uint16_t) (int16_t) h2;
            prof[NW_TAB + R * 32 + i] = ( (uint32_t) (uint16_t) (int16_t) ( h1 + qr ) << 16 ) | (uint32_t) (uint16_t) (int16_t) ( h2 + qr );
        }
    }
    __syncthreads();

    // A warp is split into 32/G independent systolic arrays of G lanes (G = 4, 8, 16 or 32, chosen per query so
    // that G*R covers the query with few padded rows while R stays large enough to amortise the per-step work).
    // Every step a lane advances by TWO database columns (A = even, B = odd column of the step): the two columns'
    // F chains are independent, which doubles the instruction-level parallelism of the otherwise serial row loop,
    // and the per-step work (shuffles, flags, loads) is shared by 2R cells.  Sequence pairs are padded to an even
    // number of columns in the stream, so a sequence always starts in column A.
    // The packed tier has no one-column path: SW pairs are padded with a harmless column (PAD_RESIDUE); packed NW reads its result
    // from the values the lane with the query's last row hands on (that row is the lane's last one: ScoreArgs::nw_shift)
    constexpr bool HAS_SLOW = !PACKED;   // (packed NW reads its result from the values a lane hands on: nw_shift)
    // passes of a long query on full-warp arrays: the group IS the warp, known at compile time
    constexpr bool FULLWARP = MULTI == 1;
    const int lane = threadIdx.x & 31;
    const int G = FULLWARP ? 32 : (int) a.group;
    const uint32_t nbatches = a.nbatches_dev ? *a.nbatches_dev : a.nbatches;
    const int gmask = G - 1;
    const int gl = lane & gmask;                                   // lane within the group
    const unsigned group_mask = ( G == 32 ? FULL : ( ( 1u << G ) - 1u ) << ( lane & ~gmask ) );
    const bool is_top = gl == 0;
    const bool is_bottom = gl == gmask;
    // packed tier: 64-bit entries [row quad][residue][lane] = two (row pair) words; wide tier: 32-bit entries [row][residue][lane]
    constexpr int MR = DUAL ? R : ( PACKED ? merge_rows( R ) : 0 );   // (DUAL: every row has a 128-bit entry of ready-made words)
    // packed tier: 128-bit entries for the first MR rows, then 64-bit entries; wide tier: 32-bit entries [row][residue][lane]
    const uint32_t prof_base = (uint32_t) __cvta_generic_to_shared( prof ) + lane * ( PACKED ? 16 : 4 );
    constexpr int NQ16 = PACKED ? ( R - MR ) / 4 : 0;   // 64-bit entries (4 rows, two scores per word) after the first MR rows
    const uint32_t prof_base16 = (uint32_t) __cvta_generic_to_shared( prof ) + MR * 4096 + lane * 8;
    const uint32_t K65536 = a.k65536;    // 65536 as a kernel argument: the merge stays an IMAD with a constant-bank operand
                                         // (FMA-side pipe) instead of being strength-reduced to a shift/LEA on the ALU pipe
    const T QR = (T) a.qr_word;   // gapO + gapE, replicated per half for the packed tier
    const T RG = (T) a.rg_word;   // gapE
    const int row0 = a.rowbase + gl * R;
    const bool gen_top = is_top && ( MULTI == 0 || a.rowbase == 0 );  // this lane synthesises the matrix' top boundary

    T Hp[R], E[R];   // H of the last processed column / E entering the next column, per row of this lane
#pragma unroll
    for( int r = 0; r < R; r++ ) {
        Hp[r] = A::bcast( 0 );
        E[r] = A::bcast( 0 );
    }
    T S = A::bcast( 0 );                // SW: running maximum of this lane's rows
    T diag = A::bcast( 0 );             // H of (row0-1, last processed column)
    T top_h = A::bcast( 0 ), top_f = A::bcast( 0 );
    T ha_out = A::bcast( 0 ), fa_out = A::bcast( 0 ), hb_out = A::bcast( 0 ), fb_out = A::bcast( 0 ), s_out = A::bcast( 0 );
    T cap_b = A::bcast( 0 );            // NW PACKED: score of sequence B, kept until the pair ends
    int jcol = 0;                       // TRACK: column index within the sequence
    int f8 = 0x7fffffff, f16 = 0x7fffffff, f8_out = 0x7fffffff, f16_out = 0x7fffffff;

    // current work batch of this group
    const uint32_t * st = (const uint32_t *) a.stream;   // two 16-bit column words per element
    uint32_t bnd_base = 0;              // MULTI: this batch's first column pair in the pass-boundary buffers
    int nsteps_cols = 0, nsteps = 0, t = 0, out_idx = 0;   // nsteps_cols = column pairs of the batch
    uint32_t half = 0;
    bool done = false;
    uint32_t w_n1 = 0, w_n2 = 0;        // column words of the next two steps (software prefetch)
    // MULTI, full-warp groups: the previous pass' boundary row (H and F of two columns per step) is fetched by the
    // whole warp, 32 steps at a time and one block of 32 ahead - lane l holds the entry of step 32*blk + l - and reaches
    // the top lane through an indexed shuffle: no divergent code, no global latency on the step's path.
    uint4 b_cur = make_uint4( 0, 0, 0, 0 ), b_nxt = make_uint4( 0, 0, 0, 0 );

    // one database column through this lane's R rows (used where the two-column loop does not apply)
    auto single_column = [&]( const uint32_t ci, const T h_in, const T f_in, T & h_last, T & f_last ) {
        T f = f_in;
        T hd = diag;
        diag = h_in;
        if( PACKED ) {
            const uint32_t ra = DUAL ? ( ci >> ( half * 5u ) ) & 31u : ci & 31u, rb = ( ci >> 5 ) & 31u;
            T hd2 = hd;
            // rows 2k, 2k+1 with their (A:B) substitution scores v0, v1
            auto dp_pair = [&]( const int k, const uint32_t v0, const uint32_t v1 ) {
                const T t0 = A::add( hd2, (T) v0 );
                const T t1 = A::add( Hp[2 * k], (T) v1 );
                hd2 = Hp[2 * k + 1];
                T hq;
                Hp[2 * k] = NW ? A::max3( t0, E[2 * k], f ) : A::max3_relu( t0, E[2 * k], f );
                hq = A::add( Hp[2 * k], QR );
                E[2 * k] = A::addmax( E[2 * k], RG, hq );
                f = A::addmax( f, RG, hq );
                Hp[2 * k + 1] = NW ? A::max3( t1, E[2 * k + 1], f ) : A::max3_relu( t1, E[2 * k + 1], f );
                hq = A::add( Hp[2 * k + 1], QR );
                E[2 * k + 1] = A::addmax( E[2 * k + 1], RG, hq );
                f = A::addmax( f, RG, hq );
                if( !NW ) S = A::max3( S, t0, t1 );
            };
#pragma unroll
            for( int q = 0; q < MR / 4; q++ ) {   // one zero-extended score per word: merge by multiply-add
                const uint4 la = lds128( prof_base + ( ra << 9 ) + q * 16384 );
                if( DUAL ) {
                    dp_pair( 2 * q, la.x, la.y );
                    dp_pair( 2 * q + 1, la.z, la.w );
                }
                else {
                    const uint4 lb = lds128( prof_base + ( rb << 9 ) + q * 16384 );
                    dp_pair( 2 * q, la.x * K65536 + lb.x, la.y * K65536 + lb.y );
                    dp_pair( 2 * q + 1, la.z * K65536 + lb.z, la.w * K65536 + lb.w );
                }
            }
#pragma unroll
            for( int q = 0; q < NQ16; q++ ) {      // two scores per word, four rows per load
                const uint2 la = lds64( prof_base16 + ( ra << 8 ) + q * 8192 ), lb = lds64( prof_base16 + ( rb << 8 ) + q * 8192 );
                dp_pair( MR / 2 + 2 * q, __byte_perm( lb.x, la.x, 0x5410 ), __byte_perm( lb.x, la.x, 0x7632 ) );
                dp_pair( MR / 2 + 2 * q + 1, __byte_perm( lb.y, la.y, 0x5410 ), __byte_perm( lb.y, la.y, 0x7632 ) );
            }
        }
        else {
            const uint32_t d = half ? ( ( ci >> 5 ) & 31u ) : ( ci & 31u );
            const uint32_t pa = prof_base + ( d << 7 );
            T colmax = A::bcast( 0 );
#pragma unroll
            for( int r = 0; r < R; r++ ) {
                const T v = (T) lds32( pa + r * 4096 );
                const T tt = A::add( hd, v );
                hd = Hp[r];
                const T h = NW ? A::max3( tt, E[r], f ) : A::max3_relu( tt, E[r], f );
                const T hq = A::add( h, QR );
                E[r] = A::addmax( E[r], RG, hq );
                f = A::addmax( f, RG, hq );
                Hp[r] = h;
                if( !NW ) colmax = A::max2( colmax, h );
            }
            if( !NW ) S = A::max2( S, colmax );
            if( TRACK ) {
                if( (int) colmax >= a.thr8 && f8 == 0x7fffffff ) f8 = jcol;
                if( (int) colmax >= a.thr16 && f16 == 0x7fffffff ) f16 = jcol;
                jcol++;
            }
        }
        h_last = Hp[R - 1];
        f_last = f;
    };

    // NW, 32-bit tier: the lane that owns the query's last row records H at the last column of its sequence (the packed tier reads
    // its result from the values a lane hands on, see the two-column path)
    auto nw_capture = [&]( const uint32_t ci ) {
        if( gl == a.cap_lane && ( ci & ( COL_LAST_A | COL_LAST_B ) ) ) {
            // pick H of the query's last row without turning Hp[] into an indexable (local-memory) array:
            // a chain of predicated moves, kept opaque to the optimiser
            T sel = Hp[0];
#pragma unroll
            for( int r = 1; r < R; r++ ) {
                T cand = Hp[r];
                asm volatile( "" : "+r"( cand ) );
                if( r == a.cap_row ) sel = cand;
                asm volatile( "" : "+r"( sel ) );
            }
            if( ci & ( half ? COL_LAST_B : COL_LAST_A ) ) ( (int32_t *) a.out )[out_idx] = (int32_t) sel + a.offset;
        }
    };

    while( true ) {
        if( t >= nsteps ) {  // group-uniform: this group's batch is finished, pull the next one from the queue
            uint32_t b = 0xffffffffu;
            if( is_top && !done ) b = atomicAdd( a.counter, 1u );
            b = __shfl_sync( group_mask, b, 0, G );
            if( b >= nbatches ) {
                done = true;
                nsteps_cols = 0;
                nsteps = 0x7fffffff;
                w_n1 = 0;
                w_n2 = 0;
            }
            else {
                const Batch bt = a.batches[b];
                st = (const uint32_t *) ( a.stream + bt.col0 );   // col0 is even: 4-byte aligned
                nsteps_cols = (int) ( bt.ncols >> 1 );
                nsteps = nsteps_cols + gmask;
                out_idx = (int) bt.out0;
                half = DUAL ? a.dual_half : bt.half;
                bnd_base = (uint32_t) ( bt.bnd0 >> 1 );
                // lane gl works on column pair t - gl at step t: preload the column words of steps 0 and 1
                w_n1 = ( gl == 0 ) ? st[0] : 0;
                w_n2 = ( 1 - gl >= 0 && 1 - gl < nsteps_cols ) ? st[1 - gl] : 0;
                if( MULTI == 1 && a.bnd_in ) {
                    if( lane < nsteps_cols ) b_cur = a.bnd_in[bnd_base + lane];
                    if( 32 + lane < nsteps_cols ) b_nxt = a.bnd_in[bnd_base + 32 + lane];
                }
                if( MULTI == 2 && a.bnd_in ) {
                    if( is_top ) {  // narrow groups (remainder pass): the top lane fetches its own entries, two steps ahead
                        if( 0 < nsteps_cols ) b_cur = a.bnd_in[bnd_base];
                        if( 1 < nsteps_cols ) b_nxt = a.bnd_in[bnd_base + 1];
                    }
                }
            }
            t = 0;
            if( FULLWARP && done ) break;   // warp-uniform: one group per warp, the exit test is off the per-step path
        }
        if( !FULLWARP && __all_sync( FULL, done ) ) break;   // also the point where the groups of a warp reconverge every step

        // Every lane reads the column words of ITS column pair (t - gl) straight from the stream, two steps ahead
        // of use: the G lanes of a group touch G consecutive 32-bit words, all groups run the same instructions.
        const uint32_t w = w_n1;
        w_n1 = w_n2;
        {
            const int c = t + 2 - gl;
            w_n2 = ( c >= 0 && c < nsteps_cols ) ? st[c] : 0;
        }
        const uint32_t cia = w & 0xffffu, cib = w >> 16;

        T ha_in = (T) __shfl_up_sync( FULL, ha_out, 1, G );
        T fa_in = (T) __shfl_up_sync( FULL, fa_out, 1, G );
        T hb_in = (T) __shfl_up_sync( FULL, hb_out, 1, G );
        T fb_in = (T) __shfl_up_sync( FULL, fb_out, 1, G );
        T s_in = A::bcast( 0 );
        int f8_in = 0x7fffffff, f16_in = 0x7fffffff;
        if( !NW ) s_in = (T) __shfl_up_sync( FULL, s_out, 1, G );
        if( TRACK ) {
            f8_in = __shfl_up_sync( FULL, f8_out, 1, G );
            f16_in = __shfl_up_sync( FULL, f16_out, 1, G );
        }
        if( is_top ) {
            s_in = A::bcast( 0 );
            f8_in = 0x7fffffff;
            f16_in = 0x7fffffff;
            if( gen_top ) {
                if( NW && ( cia & COL_FIRST ) ) {
                    top_h = A::bcast( a.gap_open + a.gap_extend - a.offset );
                    top_f = A::bcast( 2 * a.gap_open + 2 * a.gap_extend - a.offset );
                }
                ha_in = top_h;
                fa_in = top_f;
                if( NW ) {
                    top_h = A::add( top_h, RG );
                    top_f = A::add( top_f, RG );
                }
                hb_in = top_h;
                fb_in = top_f;
                if( NW ) {
                    top_h = A::add( top_h, RG );
                    top_f = A::add( top_f, RG );
                }
            }
        }
        if( MULTI == 2 && a.bnd_in ) {
            // Narrow groups: every group has its own batch and step counter, so the top lane of each group reads the
            // boundary entries of ITS column pair t itself (prefetched two steps ahead).  Only the remainder pass of a
            // long query runs this way (a few per cent of the work), the full passes use the warp-wide fetch below.
            if( is_top ) {
                ha_in = (T) b_cur.x;
                hb_in = (T) b_cur.y;
                fa_in = (T) b_cur.z;
                fb_in = (T) b_cur.w;
                b_cur = b_nxt;
                if( t + 2 < nsteps_cols ) b_nxt = a.bnd_in[bnd_base + t + 2];
            }
        }
        if( MULTI == 1 && a.bnd_in ) {  // warp-uniform (G = 32): t is the column pair of the top lane
            if( ( t & 31 ) == 0 && t > 0 ) {
                b_cur = b_nxt;
                if( t + 32 + lane < nsteps_cols ) b_nxt = a.bnd_in[bnd_base + t + 32 + lane];
            }
            const uint32_t bha = __shfl_sync( FULL, b_cur.x, t & 31 ), bhb = __shfl_sync( FULL, b_cur.y, t & 31 );
            const uint32_t bfa = __shfl_sync( FULL, b_cur.z, t & 31 ), bfb = __shfl_sync( FULL, b_cur.w, t & 31 );
            if( is_top ) {
                ha_in = (T) bha;
                hb_in = (T) bhb;
                fa_in = (T) bfa;
                fb_in = (T) bfb;
            }
        }

        T ha_last = A::bcast( 0 ), fa_last = A::bcast( 0 ), hb_last = A::bcast( 0 ), fb_last = A::bcast( 0 );
        if( cia & COL_VALID ) {
            if( cia & COL_FIRST ) {  // a new sequence (pair) starts in this lane (always in column A)
                if( NW && PACKED ) {
                    const uint32_t tab = (uint32_t) __cvta_generic_to_shared( prof ) + NW_TAB * 4 + lane * 16;
#pragma unroll
                    for( int q = 0; q < R / 4; q++ ) {
                        const uint4 h = lds128( tab + q * 512 ), e = lds128( tab + R * 128 + q * 512 );
                        Hp[4 * q] = (T) h.x;
                        Hp[4 * q + 1] = (T) h.y;
                        Hp[4 * q + 2] = (T) h.z;
                        Hp[4 * q + 3] = (T) h.w;
                        E[4 * q] = (T) e.x;
                        E[4 * q + 1] = (T) e.y;
                        E[4 * q + 2] = (T) e.z;
                        E[4 * q + 3] = (T) e.w;
                    }
                }
                else {
                    T hv = A::bcast( 0 ), ev = A::bcast( 0 );
                    if( NW ) {
                        hv = A::bcast( a.gap_open + ( row0 + 1 ) * a.gap_extend - a.offset );
                        ev = A::bcast( 2 * a.gap_open + ( row0 + 2 ) * a.gap_extend - a.offset );
                        // opaque to the optimiser: otherwise the 2R start values are hoisted out of the step loop as loop
                        // invariants and occupy 2R registers for the whole kernel
                        asm volatile( "" : "+r"( hv ), "+r"( ev ) );
                    }
#pragma unroll
                    for( int r = 0; r < R; r++ ) {
                        Hp[r] = hv;
                        E[r] = ev;
                        if( NW ) {
                            hv = A::add( hv, RG );
                            ev = A::add( ev, RG );
                        }
                    }
                }
                if( NW && PACKED ) {   // first diagonal: H of the array position above, column -1 (virtual rows on top: nw_start)
                    const int d1 = nw_start( row0 - 1, a.nw_shift, a.gap_open, a.gap_extend ) - a.offset;
                    const int d2 = DUAL ? nw_start( row0 - 1, a.nw_shift2, a.gap_open, a.gap_extend ) - a.offset : d1;
                    diag = (T) ( ( (uint32_t) (uint16_t) (int16_t) d1 << 16 ) | (uint32_t) (uint16_t) (int16_t) d2 );
                }
                else
                    diag = NW ? A::bcast( ( row0 == 0 ? 0 : a.gap_open + row0 * a.gap_extend ) - a.offset ) : A::bcast( 0 );
                if( !NW ) S = A::bcast( 0 );
                jcol = 0;
                f8 = 0x7fffffff;
                f16 = 0x7fffffff;
            }

            const bool b_valid = ( cib & COL_VALID ) != 0;
            const bool slow = HAS_SLOW;
            if( slow ) {
                // column by column: the 32-bit tier (one sequence per group; NW reads its result after the column in which the
                // sequence ends)
                single_column( cia, ha_in, fa_in, ha_last, fa_last );
                if( NW ) nw_capture( cia );
                if( b_valid ) {
                    single_column( cib, hb_in, fb_in, hb_last, fb_last );
                    if( NW ) nw_capture( cib );
                }
            }
            else {
                // ---- two columns per step, packed tier ----
                // Row r of column B needs row r of column A (E, H) and row r-1 of column B (F); row r+1 of column A
                // needs neither, so the two F chains overlap.  The diagonal candidate of the NEXT row of column A
                // (old H + substitution score) is formed before that row's H register is overwritten, so every H
                // stays in one physical register across steps.
                const uint32_t hs = DUAL ? half * 5u : 0u;   // DUAL: the residue of this launch half's sequence
                const uint32_t ra = ( cia >> hs ) & 31u, rb = ( cia >> 5 ) & 31u, sa = ( cib >> hs ) & 31u, sb = ( cib >> 5 ) & 31u;
                T fa = fa_in, fb = fb_in;
                T ha_prev = ha_in;                       // H of column A, previous row (diagonal of column B)
                T hda = diag;                            // H of column A-1, previous row (diagonal of column A)
                diag = hb_in;
                // rows 2k, 2k+1 of both columns; va0/va1 and vb0/vb1 = (A:B) substitution scores in column A resp. B
                auto dp_pair = [&]( const int k, const uint32_t va0, const uint32_t va1, const uint32_t vb0, const uint32_t vb1 ) {
                    // diagonal candidates of column A are formed from the OLD H registers before they are overwritten
                    const T ta0 = A::add( hda, (T) va0 );
                    const T ta1 = A::add( Hp[2 * k], (T) va1 );
                    hda = Hp[2 * k + 1];
                    T hq;
                    // column A, row 2k
                    const T ha0 = NW ? A::max3( ta0, E[2 * k], fa ) : A::max3_relu( ta0, E[2 * k], fa );
                    hq = A::add( ha0, QR );
                    const T ea0 = A::addmax( E[2 * k], RG, hq );
                    fa = A::addmax( fa, RG, hq );
                    // column B, row 2k
                    const T tb0 = A::add( ha_prev, (T) vb0 );
                    Hp[2 * k] = NW ? A::max3( tb0, ea0, fb ) : A::max3_relu( tb0, ea0, fb );
                    hq = A::add( Hp[2 * k], QR );
                    E[2 * k] = A::addmax( ea0, RG, hq );
                    fb = A::addmax( fb, RG, hq );
                    // column A, row 2k+1
                    const T ha1 = NW ? A::max3( ta1, E[2 * k + 1], fa ) : A::max3_relu( ta1, E[2 * k + 1], fa );
                    hq = A::add( ha1, QR );
                    const T ea1 = A::addmax( E[2 * k + 1], RG, hq );
                    fa = A::addmax( fa, RG, hq );
                    // column B, row 2k+1
                    const T tb1 = A::add( ha0, (T) vb1 );
                    Hp[2 * k + 1] = NW ? A::max3( tb1, ea1, fb ) : A::max3_relu( tb1, ea1, fb );
                    hq = A::add( Hp[2 * k + 1], QR );
                    E[2 * k + 1] = A::addmax( ea1, RG, hq );
                    fb = A::addmax( fb, RG, hq );
                    ha_prev = ha1;
                    // The running maximum is taken over the diagonal candidates t = Hdiag + V instead of over H:
                    // with non-positive gap costs E and F never exceed an earlier H, so max(H) = max(0, t).  Giving
                    // t a second consumer also keeps ptxas from fusing the add into a VIADDMNMX, i.e. it stays a
                    // VIADD.16x2 on the FMA-side pipe and off the (binding) ALU pipe.  NW needs no maximum: the packed tier
                    // only runs pairs whose value range is statically inside int16 (packed_safe_len).
                    if( !NW ) {
                        S = A::max3( S, ta0, ta1 );
                        S = A::max3( S, tb0, tb1 );
                    }
                    else if( PACKED ) {
                        // (NW: a candidate with a single consumer would be fused into VIADDMNMX + VIMNMX, 5 instead of 4 ALU-pipe
                        // instructions per word; one multiply-add per two words on the FMA-side pipe keeps it a value - sw32_kernel.cuh)
                        S = (T) ( (uint32_t) ta0 * (uint32_t) S + (uint32_t) ta1 );
                        S = (T) ( (uint32_t) tb0 * (uint32_t) S + (uint32_t) tb1 );
                    }
                };
                // one shared load per sequence and column fetches the scores of four query rows
#pragma unroll
                for( int q = 0; q < MR / 4; q++ ) {   // one zero-extended score per word: merge by multiply-add
                    const uint4 laa = lds128( prof_base + ( ra << 9 ) + q * 16384 ), lab = lds128( prof_base + ( sa << 9 ) + q * 16384 );
                    if( DUAL ) {   // the entry holds the (q1 : q2) words of four rows
                        dp_pair( 2 * q, laa.x, laa.y, lab.x, lab.y );
                        dp_pair( 2 * q + 1, laa.z, laa.w, lab.z, lab.w );
                    }
                    else {
                        const uint4 lba = lds128( prof_base + ( rb << 9 ) + q * 16384 ), lbb = lds128( prof_base + ( sb << 9 ) + q * 16384 );
                        dp_pair( 2 * q, laa.x * K65536 + lba.x, laa.y * K65536 + lba.y, lab.x * K65536 + lbb.x, lab.y * K65536 + lbb.y );
                        dp_pair( 2 * q + 1, laa.z * K65536 + lba.z, laa.w * K65536 + lba.w, lab.z * K65536 + lbb.z, lab.w * K65536 + lbb.w );
                    }
                }
#pragma unroll
                for( int q = 0; q < NQ16; q++ ) {      // two scores per word, four rows per load
                    const uint2 laa = lds64( prof_base16 + ( ra << 8 ) + q * 8192 ), lba = lds64( prof_base16 + ( rb << 8 ) + q * 8192 );
                    const uint2 lab = lds64( prof_base16 + ( sa << 8 ) + q * 8192 ), lbb = lds64( prof_base16 + ( sb << 8 ) + q * 8192 );
                    dp_pair( MR / 2 + 2 * q, __byte_perm( lba.x, laa.x, 0x5410 ), __byte_perm( lba.x, laa.x, 0x7632 ),
                            __byte_perm( lbb.x, lab.x, 0x5410 ), __byte_perm( lbb.x, lab.x, 0x7632 ) );
                    dp_pair( MR / 2 + 2 * q + 1, __byte_perm( lba.y, laa.y, 0x5410 ), __byte_perm( lba.y, laa.y, 0x7632 ),
                            __byte_perm( lbb.y, lab.y, 0x5410 ), __byte_perm( lbb.y, lab.y, 0x7632 ) );
                }
                ha_last = ha_prev;
                fa_last = fa;
                hb_last = Hp[R - 1];
                fb_last = fb;
                if( NW ) {   // packed NW: the query's last row is the last row of its lane - H of that row is what the lane hands on
                    const uint32_t pos = bnd_base + (uint32_t) ( t - gl );
                    if( DUAL ) {
                        uint16_t * o = (uint16_t *) ( (uint32_t *) a.out + half * a.out_stride + pos );
#pragma unroll
                        for( int c = 0; c < 2; c++ ) {
                            if( ( c ? cib : cia ) & ( half ? COL_LAST_B : COL_LAST_A ) ) {
                                const uint32_t v = (uint32_t) ( c ? hb_last : ha_last );
                                if( gl == a.cap_lane ) o[1] = (uint16_t) ( v >> 16 );       // first query = high half
                                if( gl == a.cap_lane2 ) o[0] = (uint16_t) ( v & 0xffffu );
                            }
                        }
                    }
                    else if( gl == a.cap_lane ) {
#pragma unroll
                        for( int c = 0; c < 2; c++ ) {
                            const uint32_t ci = c ? cib : cia, v = (uint32_t) ( c ? hb_last : ha_last );
                            if( ci & COL_LAST_B ) cap_b = (T) v;
                            if( ci & COL_LAST_A ) ( (uint32_t *) a.out )[pos] = __byte_perm( (uint32_t) cap_b, v, 0x7610 );
                        }
                    }
                }
            }

            // ---- SW results at sequence ends (rare, divergent) ----
            if( !NW ) {
                const uint32_t last_mask = PACKED && !DUAL ? COL_LAST_A : ( half ? COL_LAST_B : COL_LAST_A );
                s_out = A::bcast( 0 );
                f8_out = 0x7fffffff;
                f16_out = 0x7fffffff;
                if( ( cia | cib ) & last_mask ) {
                    s_out = A::max2( s_in, S );
                    if( TRACK ) {
                        f8_out = min( f8_in, f8 );
                        f16_out = min( f16_in, f16 );
                    }
                    if( is_bottom ) {
                        if( PACKED ) {
                            uint32_t * o = (uint32_t *) a.out + ( DUAL ? half * a.out_stride : 0u ) + ( bnd_base + (uint32_t) ( t - gmask ) );
                            *o = a.first_pass ? (uint32_t) s_out : __vmaxs2( *o, (uint32_t) s_out );
                        }
                        else {
                            int32_t * o = (int32_t *) a.out + out_idx;
                            *o = a.first_pass ? (int32_t) s_out : max( *o, (int32_t) s_out );
                            if( TRACK ) {
                                atomicMin( a.first8 + out_idx, f8_out );
                                atomicMin( a.first16 + out_idx, f16_out );
                            }
                        }
                    }
                }
            }
            if( MULTI && is_bottom && a.bnd_out )   // this lane's last row becomes the next pass' top boundary
                a.bnd_out[bnd_base + t - gmask] = make_uint4( (uint32_t) ha_last, (uint32_t) hb_last, (uint32_t) fa_last, (uint32_t) fb_last );
        }
        ha_out = ha_last;
        fa_out = fa_last;
        hb_out = hb_last;
        fb_out = fb_last;

        t++;
    }
    if( NW && PACKED && (uint32_t) S == 0x9e3779b9u && a.nbatches == 0xffffffffu ) ( (uint32_t *) a.out )[0] = (uint32_t) S;   // never true
}

}  // namespace ssa
