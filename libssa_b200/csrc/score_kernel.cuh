// Scoring kernels for sm_100a: affine-gap Smith-Waterman / Needleman-Wunsch score of one query
// against a stream of database sequences.  Replaces the reference's SIMD kernels
// (src/algo/simd/search_simd_{sw,nw}.c), its profile transposition (src/algo/{8,16}/search_*_util.c),
// the scalar 64-bit tier (src/algo/64/*.c) and the chunk loop around them.
//
// Decomposition (one warp = one systolic array):
//   * lane l of a warp owns R consecutive query rows [rowbase + l*R, rowbase + (l+1)*R); their H (previous
//     column) and E state live in registers for the whole launch;
//   * database columns stream through the lanes: at step t lane l works on column t-l and hands the H and F
//     of its last row (plus the column's residues) to lane l+1 with __shfl_up_sync - the anti-diagonal
//     wavefront, one shuffle hop per step;
//   * the columns of MANY database sequences are concatenated into one stream per work batch; flag bits in the
//     stream mark first/last columns, where a lane resets its state / reports its result.  The wavefront
//     therefore fills and drains once per batch (31 steps per >= 4096 columns), not once per sequence;
//   * PACKED tier: every 32-bit register holds the same DP cell of TWO database sequences (A in the high, B in
//     the low 16 bits): inter-sequence SIMD-in-word.  All arithmetic is one SASS instruction per packed word:
//     VIADD.16x2 (FMA-side pipe), VIMNMX3.S16x2[.RELU] and VIADDMNMX.S16x2 (ALU pipe, DPX);
//   * substitution scores come from a query profile in shared memory laid out [row pair][residue][lane], so a
//     warp-wide lookup is bank-conflict free whatever residues the lanes hold: two LDS fetch the scores of two
//     query rows for sequence A resp. B, two PRMT interleave them into the (A:B) words of the two rows;
//   * WIDE tier (PACKED=false): the same systolic array on plain int32 for the (rare) sequences whose scores
//     cannot be guaranteed to fit 16 bits - the counterpart of the reference's 16->64 bit overflow cascade.
//     Optionally tracks the first column whose running maximum reaches 255 / 65535, which is what the
//     reference's 8/16-bit tiers would have overflowed at (needed to reproduce its hit order, replay.cpp);
//   * queries longer than 32*R rows are processed in passes; the H/F of the last row of a pass is written to a
//     boundary buffer in global memory and read back by lane 0 of the next pass.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace ssa {

constexpr uint32_t COL_FIRST = 1u << 10;   // first column of a sequence pair
constexpr uint32_t COL_LAST_B = 1u << 11;  // last real column of sequence B (low half)
constexpr uint32_t COL_LAST_A = 1u << 12;  // last column of sequence A = last column of the pair
constexpr uint32_t COL_VALID = 1u << 15;   // set by lane 0 for real columns

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
    const uint32_t * profile;  // this pass' query profile, laid out for shared memory
    uint32_t group;            // lanes per systolic array: 4, 8, 16 or 32
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
    void * out;                // PACKED: uint32 (A:B) per pair; WIDE: int32 per slot
    int32_t * first8;          // TRACK: first column reaching thr8 / thr16, per slot (atomicMin)
    int32_t * first16;
    int32_t thr8, thr16;
    const uint32_t * bnd_h_in;  // boundary row from the previous pass (nullptr in pass 0)
    const uint32_t * bnd_f_in;
    uint32_t * bnd_h_out;       // boundary row for the next pass (nullptr in the last pass)
    uint32_t * bnd_f_out;
    int32_t * guard;            // PACKED NW: largest diagonal candidate seen (both halves), for the host's range check
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

// words of shared memory the profile of one pass needs
template<bool PACKED, int R> __host__ __device__ constexpr int profile_words() { return PACKED ? ( R / 2 ) * 1024 : R * 1024; }

// resident CTAs per SM the register allocator has to make room for (64K registers / 256 threads / blocks)
template<bool PACKED, int R> __host__ __device__ constexpr int min_blocks() { return PACKED ? ( R <= 12 ? 4 : ( R <= 20 ? 3 : 2 ) ) : 2; }

template<bool PACKED, int R, bool NW, bool MULTI, bool TRACK>
__global__ void __launch_bounds__( SCORE_THREADS, min_blocks<PACKED, R>() ) score_kernel( const ScoreArgs a ) {
    typedef Arith<PACKED> A;
    typedef typename A::T T;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ uint32_t prof[];

    for( int i = threadIdx.x; i < profile_words<PACKED, R>(); i += blockDim.x ) prof[i] = a.profile[i];
    __syncthreads();

    // A warp is split into 32/G independent systolic arrays of G lanes (G = 4, 8, 16 or 32, chosen per query so
    // that G*R covers the query with few padded rows while R stays large enough to amortise the per-step work).
    const int lane = threadIdx.x & 31;
    const int G = (int) a.group;
    const int gmask = G - 1;
    const int gl = lane & gmask;                                   // lane within the group
    const unsigned group_mask = ( G == 32 ? FULL : ( ( 1u << G ) - 1u ) << ( lane & ~gmask ) );
    const bool is_top = gl == 0;
    const bool is_bottom = gl == gmask;
    const uint32_t prof_base = (uint32_t) __cvta_generic_to_shared( prof ) + lane * 4;
    const T QR = (T) a.qr_word;   // gapO + gapE, replicated per half for the packed tier
    const T RG = (T) a.rg_word;   // gapE
    const int row0 = a.rowbase + gl * R;
    const bool gen_top = is_top && ( a.rowbase == 0 );  // this lane synthesises the matrix' top boundary

    // Values of a sequence start (NW: reference needleman_wunsch_64.c:39-42,51-52) are recomputed from row0 where a
    // sequence starts (rare) instead of being held in registers.
    T Hp[R], E[R];
#pragma unroll
    for( int r = 0; r < R; r++ ) {
        Hp[r] = A::bcast( 0 );
        E[r] = A::bcast( 0 );
    }
    T S = A::bcast( NW ? -32768 : 0 );  // SW: running maximum of this lane's rows; NW: range guard
    T diag = A::bcast( 0 );             // H of (row0-1, previous column)
    T top_h = A::bcast( 0 ), top_f = A::bcast( 0 );
    T h_out = A::bcast( 0 ), f_out = A::bcast( 0 ), s_out = A::bcast( 0 );
    T cap_b = A::bcast( 0 );            // NW PACKED: score of sequence B, kept until the pair ends
    int jcol = 0;                       // TRACK: column index within the sequence
    int f8 = 0x7fffffff, f16 = 0x7fffffff, f8_out = 0x7fffffff, f16_out = 0x7fffffff;
    int guard_max = -0x7fffffff;

    // current work batch of this group
    const uint16_t * st = a.stream;
    const uint32_t * bnd_h_in = nullptr, * bnd_f_in = nullptr;
    uint32_t * bnd_h_out = nullptr, * bnd_f_out = nullptr;
    int ncols = 0, nsteps = 0, t = 0, out_idx = 0;
    uint32_t half = 0;
    bool done = false;
    uint32_t ci_n1 = 0, ci_n2 = 0;                  // column words of the next two steps (software prefetch)
    uint32_t bh_n1 = 0, bf_n1 = 0, bh_n2 = 0, bf_n2 = 0, wh = 0, wf = 0;

    while( true ) {
        if( t >= nsteps ) {  // group-uniform: this group's batch is finished, pull the next one from the queue
            uint32_t b = 0xffffffffu;
            if( is_top && !done ) b = atomicAdd( a.counter, 1u );
            b = __shfl_sync( group_mask, b, 0, G );
            if( b >= a.nbatches ) {
                done = true;
                ncols = 0;
                nsteps = 0x7fffffff;
                ci_n1 = 0;
                ci_n2 = 0;
            }
            else {
                const Batch bt = a.batches[b];
                st = a.stream + bt.col0;
                ncols = (int) bt.ncols;
                nsteps = ncols + gmask;
                out_idx = (int) bt.out0;
                half = bt.half;
                if( MULTI ) {
                    bnd_h_in = a.bnd_h_in ? a.bnd_h_in + bt.bnd0 : nullptr;
                    bnd_f_in = a.bnd_f_in ? a.bnd_f_in + bt.bnd0 : nullptr;
                    bnd_h_out = a.bnd_h_out ? a.bnd_h_out + bt.bnd0 : nullptr;
                    bnd_f_out = a.bnd_f_out ? a.bnd_f_out + bt.bnd0 : nullptr;
                }
                // lane gl works on column t - gl at step t: preload the columns of steps 0 and 1
                ci_n1 = ( gl == 0 ) ? ( (uint32_t) st[0] | COL_VALID ) : 0;
                ci_n2 = ( 1 - gl >= 0 && 1 - gl < ncols ) ? ( (uint32_t) st[1 - gl] | COL_VALID ) : 0;
                if( MULTI && bnd_h_in && is_top ) {
                    bh_n1 = bnd_h_in[0];
                    bf_n1 = bnd_f_in[0];
                    bh_n2 = ncols > 1 ? bnd_h_in[1] : 0;
                    bf_n2 = ncols > 1 ? bnd_f_in[1] : 0;
                }
            }
            t = 0;
        }
        if( __all_sync( FULL, done ) ) break;

        // Every lane reads the column word of ITS column (t - gl) straight from the stream, two steps ahead of use:
        // the G lanes of a group touch G consecutive 16-bit words, all groups run the same instruction stream.
        const uint32_t ci = ci_n1;
        ci_n1 = ci_n2;
        {
            const int c = t + 2 - gl;
            ci_n2 = ( c >= 0 && c < ncols ) ? ( (uint32_t) st[c] | COL_VALID ) : 0;
        }
        T h_in = (T) __shfl_up_sync( FULL, h_out, 1, G );
        T f_in = (T) __shfl_up_sync( FULL, f_out, 1, G );
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
                if( NW && ( ci & COL_FIRST ) ) {
                    top_h = A::bcast( a.gap_open + a.gap_extend - a.offset );
                    top_f = A::bcast( 2 * a.gap_open + 2 * a.gap_extend - a.offset );
                }
                h_in = top_h;
                f_in = top_f;
                if( NW ) {
                    top_h = A::add( top_h, RG );
                    top_f = A::add( top_f, RG );
                }
            }
            else if( MULTI ) {
                h_in = (T) bh_n1;
                f_in = (T) bf_n1;
                bh_n1 = bh_n2;
                bf_n1 = bf_n2;
                if( bnd_h_in && t + 2 < ncols ) {
                    bh_n2 = bnd_h_in[t + 2];
                    bf_n2 = bnd_f_in[t + 2];
                }
            }
        }

        T h_last = A::bcast( 0 ), f_last = A::bcast( 0 );
        if( ci & COL_VALID ) {
            if( ci & COL_FIRST ) {  // a new sequence (pair) starts in this lane
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
                diag = NW ? A::bcast( ( row0 == 0 ? 0 : a.gap_open + row0 * a.gap_extend ) - a.offset ) : A::bcast( 0 );
                if( !NW ) S = A::bcast( 0 );
                jcol = 0;
                f8 = 0x7fffffff;
                f16 = 0x7fffffff;
            }

            T f = f_in;
            T hd = diag;
            diag = h_in;
            T colmax = A::bcast( 0 );
            if( PACKED ) {
                const uint32_t pa = prof_base + ( ( ci & 31u ) << 7 );
                const uint32_t pb = prof_base + ( ( ( ci >> 5 ) & 31u ) << 7 );
                // The diagonal candidate of row r+1 (old H of row r + its substitution score) is formed BEFORE row r's
                // H register is overwritten, so every H stays in one physical register across steps (no rotation moves).
                uint32_t la = lds32( pa ), lb = lds32( pb );
                T tc = A::add( hd, (T) __byte_perm( lb, la, 0x5410 ) );
#pragma unroll
                for( int k = 0; k < R / 2; k++ ) {
                    const T t0 = tc;
                    const T t1 = A::add( Hp[2 * k], (T) __byte_perm( lb, la, 0x7632 ) );
                    if( k + 1 < R / 2 ) {
                        la = lds32( pa + ( k + 1 ) * 4096 );
                        lb = lds32( pb + ( k + 1 ) * 4096 );
                        tc = A::add( Hp[2 * k + 1], (T) __byte_perm( lb, la, 0x5410 ) );
                    }
                    T hq;
                    Hp[2 * k] = NW ? A::max3( t0, E[2 * k], f ) : A::max3_relu( t0, E[2 * k], f );
                    hq = A::add( Hp[2 * k], QR );
                    E[2 * k] = A::addmax( E[2 * k], RG, hq );
                    f = A::addmax( f, RG, hq );
                    Hp[2 * k + 1] = NW ? A::max3( t1, E[2 * k + 1], f ) : A::max3_relu( t1, E[2 * k + 1], f );
                    hq = A::add( Hp[2 * k + 1], QR );
                    E[2 * k + 1] = A::addmax( E[2 * k + 1], RG, hq );
                    f = A::addmax( f, RG, hq );
                    // The running maximum is taken over the diagonal candidates t = Hdiag + V instead of over
                    // H: with non-positive gap costs E and F never exceed an earlier H, so max(H) = max(0, t).
                    // Giving t a second consumer also keeps ptxas from fusing the add into a VIADDMNMX, i.e. it
                    // stays a VIADD.16x2 on the FMA-side pipe and off the (binding) ALU pipe.  NW uses the same
                    // accumulator as an upper range guard.
                    S = A::max3( S, t0, t1 );
                }
            }
            else {
                const uint32_t d = half ? ( ( ci >> 5 ) & 31u ) : ( ci & 31u );
                const uint32_t pa = prof_base + ( d << 7 );
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

            // ---- results at sequence ends (rare, divergent) ----
            if( NW ) {
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
                    if( PACKED ) {
                        if( ci & COL_LAST_B ) cap_b = sel;
                        if( ci & COL_LAST_A ) {
                            ( (uint32_t *) a.out )[out_idx++] = __byte_perm( (uint32_t) cap_b, (uint32_t) sel, 0x7610 );
                        }
                    }
                    else {
                        if( ci & ( half ? COL_LAST_B : COL_LAST_A ) ) ( (int32_t *) a.out )[out_idx] = (int32_t) sel + a.offset;
                    }
                }
            }
            else {
                const uint32_t last_mask = PACKED ? COL_LAST_A : ( half ? COL_LAST_B : COL_LAST_A );
                s_out = A::bcast( 0 );
                f8_out = 0x7fffffff;
                f16_out = 0x7fffffff;
                if( ci & last_mask ) {
                    s_out = A::max2( s_in, S );
                    if( TRACK ) {
                        f8_out = min( f8_in, f8 );
                        f16_out = min( f16_in, f16 );
                    }
                    if( is_bottom ) {
                        if( PACKED ) {
                            uint32_t * o = (uint32_t *) a.out + out_idx++;
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
        }
        h_out = h_last;
        f_out = f_last;

        if( MULTI ) {  // the bottom lane's last row becomes the next pass' top boundary
            const int cb = t - gmask;
            const uint32_t vh = __shfl_sync( FULL, (uint32_t) h_out, gmask, G );
            const uint32_t vf = __shfl_sync( FULL, (uint32_t) f_out, gmask, G );
            if( bnd_h_out && cb >= 0 && cb < ncols ) {
                if( gl == ( cb & gmask ) ) {
                    wh = vh;
                    wf = vf;
                }
                if( ( cb & gmask ) == gmask || cb == ncols - 1 ) {
                    const int c = ( cb & ~gmask ) + gl;
                    if( c <= cb ) {
                        bnd_h_out[c] = wh;
                        bnd_f_out[c] = wf;
                    }
                }
            }
        }
        if( NW && PACKED && t == nsteps - 1 ) {
            const int hi = (int) (int16_t) ( (uint32_t) S >> 16 ), lo = (int) (int16_t) ( (uint32_t) S & 0xffffu );
            guard_max = max( guard_max, max( hi, lo ) );
        }
        t++;
    }
    if( NW && PACKED && a.guard ) atomicMax( a.guard, guard_max );
}

}  // namespace ssa
