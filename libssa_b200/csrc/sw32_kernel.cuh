// Smith-Waterman / Needleman-Wunsch scoring on full-warp systolic arrays, packed tier (two database sequences per int16 x 2
// word): the hot kernel of protein searches (reference src/algo/simd/search_simd_sw.c:172-441, search_simd_nw.c and their
// chunk loop).  Same recurrence and same data as score_kernel.cuh (which keeps the narrow groups, the 32-bit tier and the
// tracker); what is different is everything AROUND the recurrence - that was 143 of 591 instructions per step of 64 packed words:
//
//   * one warp = one array of 32 lanes x R rows (or, for queries of up to 16 x 32 rows, two independent arrays of 16 lanes with
//     their own chains and rings), known at compile time;
//   * the work batches of a warp are CHAINED: the warp runs on a virtual step axis on which its batches follow each other
//     without a gap, lane l works on virtual step T - l, so the 32-lane wavefront fills and drains once per launch instead
//     of once per batch, and batches can be small (fine-grained balance of the work queue at no cost);
//   * everything a step needs from memory is staged in a per-warp ring in shared memory, 32 steps at a time and 64 steps
//     ahead, with cp.async (no registers, no latency on the step's path): the column words of the step, its position in the
//     column stream and - in passes of a long query - the boundary row of the previous pass.  A step reads its column
//     word with one LDS; the top lane reads the boundary with one predicated LDS.128 (was: a per-lane predicated global
//     load with 64-bit address arithmetic, and four indexed shuffles + four selects out of eight live registers);
//   * results are stored by stream position (the column pair in which a pair of sequences ends), boundary rows as one
//     16-byte entry {H col A, H col B, F col A, F col B} per column pair, so no lane has to know which batch it is in;
//   * there is no one-column path (pairs are padded with a PAD_RESIDUE column) and no per-step exit vote.  Needleman-Wunsch
//     needs none either: the query is shifted down by virtual rows so that its last row is the last row of a lane, and the
//     result - H of that row in the column in which a sequence ends - is a value the lane hands on anyway (Sw32Args::nw_shift);
//   * ALL full passes of a long query run in ONE launch: a CTA works through the passes in order (its warps meet at a CTA
//     barrier to swap the query profile in shared memory), every pass has its own work queue over the same batches, and a
//     batch of pass p+1 starts as soon as the SAME batch of pass p is finished (a per-batch counter in global memory,
//     release / acquire) - there is no grid-wide barrier between passes, so the idle tail of a pass (the longest pairs, the
//     last small batches; 10-15 % of a pass on a 70 000-record shard) is filled with the next pass' work.  Waits only ever
//     point to a lower pass of a batch some running warp has already claimed, so they cannot deadlock.
#pragma once

#include "score_kernel.cuh"

namespace ssa {

// An array of G lanes (G = 32: the warp; G = 16: two independent arrays per warp, one pass only) stages its steps in blocks of
// G, two blocks ahead; its ring holds 4 G steps (2 G ahead of the top lane, G - 1 behind it, one block being written); a batch
// occupies at least 2 G virtual steps, so a block crosses at most one batch boundary.
__host__ __device__ constexpr int sw32_ring( int G ) { return 4 * G; }
__host__ __device__ constexpr int sw32_min_steps( int G ) { return 2 * G; }
constexpr uint32_t SW32_IDLE = 0xffffffffu;
constexpr int SW32_MAX_PASSES = 32;   // passes one launch can chain (counters live in one small array)

struct Sw32Args {
    const uint32_t * stream32;   // the column stream, two 16-bit column words per element (one column pair = one step)
    const Batch * batches;       // whole pairs; col0 even
    uint32_t nbatches;
    uint32_t * counters;         // work queue heads, one per pass of this launch (zero at launch)
    const uint32_t * profile;    // the query profiles of this launch's passes (fill_profile_packed for G = 32), one after the other
    uint32_t k65536;             // see score_kernel.cuh
    uint32_t qr_word, rg_word;   // gapO + gapE and gapE, replicated in both halves
    uint32_t npass;              // passes this launch runs (1 unless MULTIPASS)
    uint32_t pass0;              // position of the first of them among the query's passes (pass 0 stores results, later ones max-combine)
    uint32_t more_passes;        // 1 = the query has passes after this launch (the last pass here writes a boundary row too)
    uint32_t * out_by_col;       // (A:B) running maxima, indexed by the column pair in which the pair of sequences ends
    uint4 * bnd0, * bnd1;        // MULTIPASS: boundary rows per column pair; an even pass reads bnd1 and writes bnd0, an odd pass the reverse
    uint32_t * done;             // MULTIPASS, npass > 1: passes of this launch completed per batch (zero at launch)
    // Needleman-Wunsch
    int32_t gap_open, gap_extend;   // gapO, gapE (non-positive)
    int32_t offset;                 // values are stored minus this offset (centres the value range on 0)
    int32_t rowbase;                // array position of the first row of this launch's first pass (its later passes follow at 32 R rows each)
    int32_t nw_shift, nw_shift2;    // virtual rows on top of the query (DUAL: of the first / second query): query row = position - shift,
                                    // chosen so that the query's last row is the LAST row of a lane (search.cu:fill_profile_packed)
    int32_t cap_pass;               // pass of the QUERY (not of the launch) that holds the query's last row ...
    int32_t cap_lane, cap_row;      // ... and the lane / row in the lane
    // DUAL (two QUERIES per packed word, score_kernel.cuh): one launch per sequence of the pairs
    uint32_t dual_half;             // 0 = sequence A of every pair (residue bits 0-4), 1 = sequence B (bits 5-9)
    uint32_t out_stride;            // words between the result arrays of the two launch halves
    int32_t cap_lane2, cap_row2;    // NW: lane / row of the second query's last row (cap_lane / cap_row: the first query's)
};

// words of shared memory per warp for the ring: column words, stream positions, boundary entries
template<bool MULTIPASS, int G = 32> __host__ __device__ constexpr int sw32_ring_words() {
    return 2 * sw32_ring( G ) + ( MULTIPASS ? 5 * sw32_ring( G ) : 0 ) + 16;
}
// Needleman-Wunsch: start values of a lane's rows (H(i, -1) and the E entering column 0), [H | E][row quad][lane] as 128-bit entries
template<int R> __host__ __device__ constexpr int sw32_nw_table_words() { return 2 * R * 32; }
template<int R, bool DUAL> __host__ __device__ constexpr int sw32_profile_words() { return DUAL ? R * 1024 : profile_words<true, R>(); }
template<int R, bool MULTIPASS, bool NW = false, int G = 32, bool DUAL = false> __host__ __device__ constexpr int sw32_smem_bytes() {
    return ( sw32_profile_words<R, DUAL>() + ( SCORE_THREADS / G ) * sw32_ring_words<MULTIPASS, G>() + ( NW ? sw32_nw_table_words<R>() : 0 ) ) * 4;
}

__device__ __forceinline__ void cp_async4( uint32_t smem_addr, const void * gmem ) {
    asm volatile( "cp.async.ca.shared.global [%0], [%1], 4;" ::"r"( smem_addr ), "l"( gmem ) : "memory" );
}
__device__ __forceinline__ void cp_async16( uint32_t smem_addr, const void * gmem ) {
    asm volatile( "cp.async.cg.shared.global [%0], [%1], 16;" ::"r"( smem_addr ), "l"( gmem ) : "memory" );
}
__device__ __forceinline__ void cp_async_commit() { asm volatile( "cp.async.commit_group;" ::: "memory" ); }
template<int N> __device__ __forceinline__ void cp_async_wait() { asm volatile( "cp.async.wait_group %0;" ::"n"( N ) : "memory" ); }
__device__ __forceinline__ void sts32( uint32_t addr, uint32_t v ) { asm volatile( "st.shared.u32 [%0], %1;" ::"r"( addr ), "r"( v ) : "memory" ); }
__device__ __forceinline__ uint32_t ld_acquire( const uint32_t * p ) {
    uint32_t v;
    asm volatile( "ld.acquire.gpu.global.u32 %0, [%1];" : "=r"( v ) : "l"( p ) : "memory" );
    return v;
}
__device__ __forceinline__ void st_release( uint32_t * p, uint32_t v ) {
    asm volatile( "st.release.gpu.global.u32 [%0], %1;" ::"l"( p ), "r"( v ) : "memory" );
}
__device__ __forceinline__ void sts128( uint32_t addr, uint4 v ) {
    asm volatile( "st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"( addr ), "r"( v.x ), "r"( v.y ), "r"( v.z ), "r"( v.w ) : "memory" );
}

template<int R, bool MULTIPASS, bool NW, int G = 32, bool DUAL = false>
__global__ void __launch_bounds__( SCORE_THREADS, min_blocks<true, R>() ) sw32_kernel( const Sw32Args a ) {
    typedef Arith<true> A;
    typedef uint32_t T;
    static_assert( G == 32 || G == 16, "arrays of 32 or 16 lanes" );
    static_assert( !MULTIPASS || ( G == 32 && !DUAL ), "several passes: full-warp arrays, one query per word" );
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int PW = sw32_profile_words<R, DUAL>();
    constexpr int MR = DUAL ? R : merge_rows( R );   // (DUAL: every row has a 128-bit entry of ready-made (q1 : q2) words)
    constexpr int NQ16 = ( R - MR ) / 4;
    constexpr int RING = sw32_ring( G ), BLK = G, MIN_STEPS = sw32_min_steps( G );
    constexpr int RW = sw32_ring_words<MULTIPASS, G>();   // shared-memory words of one array's ring
    extern __shared__ uint32_t smem[];

    const int lane = threadIdx.x & 31;
    const int gl = lane & ( G - 1 );                       // lane within the array
    const int grp = threadIdx.x / G;                       // array within the CTA
    const unsigned gmask = G == 32 ? FULL : 0xffffu << ( lane & 16 );   // the lanes of this array
    {   // every ring slot idle: lanes below the top one start on steps "before the launch"
        uint32_t * ring = smem + PW + grp * RW;
        for( int i = gl; i < RW; i += G ) ring[i] = 0;
    }

    const uint32_t smem_base = (uint32_t) __cvta_generic_to_shared( smem );
    const uint32_t ring_w = smem_base + ( PW + grp * RW ) * 4;                               // [RING] column words of the step
    const uint32_t ring_g = ring_w + RING * 4;                                               // [RING] column pair index in the stream
    const uint32_t ring_b = ring_g + RING * 4;                                               // [RING] boundary entry (MULTIPASS)
    const uint32_t ring_e = ring_b + RING * 16;                                              // [RING] batch + 1 on its last step (MULTIPASS)
    const uint32_t ring_ws = ring_w + ( RW - 16 ) * 4;                                       // [12] the array's batch bookkeeping (WS_*)
    const uint32_t ring_cap = ring_w + ( RW - 4 ) * 4;                                       // NW: score of sequence B until the pair ends
    constexpr int NW_TAB = PW + ( SCORE_THREADS / G ) * RW;                                  // first word of the NW start-value table
    const uint32_t nw_tab = smem_base + NW_TAB * 4 + lane * 16;
    const uint32_t prof_base = smem_base + lane * 16;            // 128-bit entries of the first MR rows
    const uint32_t prof_base16 = smem_base + MR * 4096 + lane * 8;   // 64-bit entries of the other rows
    const uint32_t K65536 = a.k65536;
    const T QR = a.qr_word, RG = a.rg_word;
    const bool is_top = gl == 0, is_bottom = gl == G - 1;
    const uint32_t hs = DUAL ? a.dual_half * 5u : 0u;                                        // DUAL: which residue of a column word
    const uint32_t end_mask = DUAL ? ( a.dual_half ? COL_LAST_B : COL_LAST_A ) : COL_LAST_A; // the column that ends what this launch scores
    uint32_t * const out = a.out_by_col + ( DUAL ? a.dual_half * a.out_stride : 0u );

    T Hp[R], E[R];
#pragma unroll
    for( int r = 0; r < R; r++ ) {
        Hp[r] = 0;
        E[r] = 0;
    }
    T S = 0, diag = 0;
    T ha_out = 0, fa_out = 0, hb_out = 0, fb_out = 0, s_out = 0;
    // Needleman-Wunsch: the matrix' top boundary H(-1, j) = gapO + (j + 1) gapE is generated by the top lane of the query's first
    // pass (F(-1, j) = H(-1, j) + gapO + gapE); a lane's rows start a sequence at H(i, -1) = gapO + (i + 1) gapE
    T top_h = 0;
    bool is_cap = false;           // this lane holds the query's last row in the current pass

    // ---- the warp's batches on its virtual step axis (warp-uniform) ----
    // cur / nxt: the two batches the staging front can touch.  A batch of n column pairs occupies max( n, SW32_MIN_STEPS )
    // virtual steps; steps past its columns, and everything after the queue ran dry, are idle.  This bookkeeping is touched
    // once per 32 steps, so it lives in shared memory (every lane writes and reads the same words) and not in registers the
    // recurrence needs: WS_CUR_S virtual begin of the current batch, _E its end (= begin of the next), _N column pairs, _G first
    // column pair in the stream, _B batch index; WS_LAST virtual end of the last real batch; WS_DRY the pass' queue is exhausted.
    enum { WS_CUR_S = 0, WS_CUR_E, WS_CUR_N, WS_CUR_G, WS_CUR_B, WS_NXT_E, WS_NXT_N, WS_NXT_G, WS_NXT_B, WS_LAST, WS_DRY };
    // (Smith-Waterman with R >= 24 has the registers to spare and is 1.4 % faster with the bookkeeping in them [B200];
    // every access uses a compile-time index, so the array below never leaves the register file.)
    constexpr bool WS_SMEM = NW || R < 24;
    uint32_t ws_reg[WS_DRY + 1] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
    auto ws_get = [&]( int i ) { return WS_SMEM ? lds32( ring_ws + i * 4 ) : ws_reg[i]; };
    auto ws_set = [&]( int i, uint32_t v ) {
        if( WS_SMEM )
            sts32( ring_ws + i * 4, v );
        else
            ws_reg[i] = v;
    };
    uint32_t pp = 0;                                          // pass of this launch the CTA works on
    uint32_t * counter = a.counters;
    const uint4 * bnd_in = nullptr;
    uint4 * bnd_out = nullptr;
    bool first_pass = true, publish = false;

    // pulls the next batch from the pass' queue; it begins at virtual step `begin`; fills slot WS_CUR_* (which = 0) or WS_NXT_* (1)
    auto fetch = [&]( const uint32_t begin, const int which ) {
        uint32_t b = 0xffffffffu;
        if( !ws_get( WS_DRY ) ) {
            if( is_top ) b = atomicAdd( counter, 1u );
            b = __shfl_sync( gmask, b, lane & ~( G - 1 ) );
        }
        uint32_t e, n = 0, g = 0, bidx = 0;
        if( b >= a.nbatches ) {
            ws_set( WS_DRY, 1 );
            e = begin + MIN_STEPS;
        }
        else {
            const Batch bt = a.batches[b];
            n = bt.ncols >> 1;
            g = (uint32_t) ( bt.col0 >> 1 );
            bidx = b;
            e = begin + max( n, (uint32_t) MIN_STEPS );
            ws_set( WS_LAST, e );
            if( MULTIPASS && pp > 0 ) {   // the previous pass of this launch must have finished this batch (its boundary row, its result)
                if( is_top )
                    while( ld_acquire( a.done + b ) < pp ) __nanosleep( 200 );
                __syncwarp( gmask );
            }
        }
        ws_set( which ? WS_NXT_E : WS_CUR_E, e );
        ws_set( which ? WS_NXT_N : WS_CUR_N, n );
        ws_set( which ? WS_NXT_G : WS_CUR_G, g );
        ws_set( which ? WS_NXT_B : WS_CUR_B, bidx );
        return e;
    };

    // stages the G virtual steps [v0, v0 + G) of this array: its lane l prepares step v0 + l
    auto stage = [&]( const uint32_t v0 ) {
        const uint32_t v = v0 + gl;
        const uint32_t slot = v & ( RING - 1 );
        const uint32_t cur_e = ws_get( WS_CUR_E );
        uint32_t c, n, g, bidx;
        if( v < cur_e ) {
            c = v - ws_get( WS_CUR_S );
            n = ws_get( WS_CUR_N );
            g = ws_get( WS_CUR_G );
            bidx = ws_get( WS_CUR_B );
        }
        else {
            c = v - cur_e;
            n = ws_get( WS_NXT_N );
            g = ws_get( WS_NXT_G );
            bidx = ws_get( WS_NXT_B );
        }
        if( c < n ) {
            const uint32_t pos = g + c;
            cp_async4( ring_w + slot * 4, a.stream32 + pos );
            sts32( ring_g + slot * 4, pos );
            if( MULTIPASS ) {
                if( bnd_in )
                    cp_async16( ring_b + slot * 16, bnd_in + pos );
                else
                    sts128( ring_b + slot * 16, make_uint4( 0, 0, 0, 0 ) );
                sts32( ring_e + slot * 4, c + 1 == n ? bidx + 1 : 0 );
            }
        }
        else {
            sts32( ring_w + slot * 4, 0 );
            sts32( ring_g + slot * 4, SW32_IDLE );
        }
        cp_async_commit();
        if( v0 + BLK >= cur_e ) {   // the front leaves the current batch: the next one becomes current, another is pulled
            __syncwarp( gmask );    // (every lane has read the old values)
            const uint32_t nxt_e = ws_get( WS_NXT_E ), nxt_n = ws_get( WS_NXT_N ), nxt_g = ws_get( WS_NXT_G ), nxt_b = ws_get( WS_NXT_B );
            __syncwarp( gmask );
            ws_set( WS_CUR_S, cur_e );
            ws_set( WS_CUR_E, nxt_e );
            ws_set( WS_CUR_N, nxt_n );
            ws_set( WS_CUR_G, nxt_g );
            ws_set( WS_CUR_B, nxt_b );
            fetch( nxt_e, 1 );
            __syncwarp( gmask );
        }
    };

    uint32_t t0 = 0;
  for( pp = 0; pp < a.npass; pp++ ) {
    // ---- the CTA moves to the next pass: every warp has drained its chain, the profile in shared memory is swapped ----
    if( pp > 0 ) __syncthreads();
    {
        const uint32_t * src = a.profile + (size_t) pp * PW;
        for( int i = threadIdx.x; i < PW; i += blockDim.x ) smem[i] = src[i];
    }
    if( NW ) {   // start values of this pass' rows (column -1; both minus the offset): real rows H(i, -1) = gapO + (i + 1) gapE, virtual
                 // rows gapO, the last virtual row 0 (= the corner H(-1, -1) row 0 takes as its first diagonal); E = H + gapO + gapE
        for( int i = threadIdx.x; i < R * 32; i += blockDim.x ) {
            const int r = ( i >> 7 ) * 4 + ( i & 3 ), l = ( i >> 2 ) & 31;   // word i of the H half: [row quad][lane][row in quad]
            const int pos = a.rowbase + (int) pp * 32 * R + ( l & ( G - 1 ) ) * R + r;
            const int h1 = nw_start( pos, a.nw_shift, a.gap_open, a.gap_extend ) - a.offset;
            const int h2 = DUAL ? nw_start( pos, a.nw_shift2, a.gap_open, a.gap_extend ) - a.offset : h1;
            const int qr = a.gap_open + a.gap_extend;
            smem[NW_TAB + i] = ( (uint32_t) (uint16_t) (int16_t) h1 << 16 ) | (uint32_t) (uint16_t) (int16_t) h2;
            smem[NW_TAB + R * 32 + i] = ( (uint32_t) (uint16_t) (int16_t) ( h1 + qr ) << 16 ) | (uint32_t) (uint16_t) (int16_t) ( h2 + qr );
        }
    }
    __syncthreads();
    {
        const uint32_t abs_pass = a.pass0 + pp;
        first_pass = abs_pass == 0;
        bnd_in = MULTIPASS && abs_pass > 0 ? ( ( abs_pass & 1 ) ? a.bnd0 : a.bnd1 ) : nullptr;
        bnd_out = MULTIPASS && ( pp + 1 < a.npass || a.more_passes ) ? ( ( abs_pass & 1 ) ? a.bnd1 : a.bnd0 ) : nullptr;
        publish = MULTIPASS && pp + 1 < a.npass;
        counter = a.counters + pp;
        if( NW ) is_cap = (int) abs_pass == a.cap_pass && ( gl == a.cap_lane || ( DUAL && gl == a.cap_lane2 ) );
    }
    // the virtual axis simply continues: the steps already staged beyond t0 are idle ones and are staged again
    __syncwarp();
    ws_set( WS_DRY, 0 );
    ws_set( WS_LAST, 0 );
    ws_set( WS_CUR_S, t0 );
    __syncwarp();
    fetch( fetch( t0, 0 ), 1 );
    __syncwarp();
    stage( t0 );
    stage( t0 + BLK );

    // Needleman-Wunsch: the query's last row is the last row of a lane (nw_shift), so the result is what that lane hands to the next
    // one: H of its last row in the column in which sequence B / A ends (the score of B waits in shared memory until the pair
    // ends).  DUAL: the sequence of this launch half ends; each query's last row may live in another lane (16-bit stores).
    auto nw_capture = [&]( const uint32_t ci, const T v, const uint32_t slot4 ) {
        if( DUAL ) {
            if( ci & end_mask ) {
                uint16_t * o = (uint16_t *) ( out + lds32( ring_g + slot4 ) );
                if( gl == a.cap_lane ) o[1] = (uint16_t) ( v >> 16 );        // first query = high half
                if( gl == a.cap_lane2 ) o[0] = (uint16_t) ( v & 0xffffu );
            }
            return;
        }
        if( ci & COL_LAST_B ) sts32( ring_cap, v );
        if( ci & COL_LAST_A ) out[lds32( ring_g + slot4 )] = __byte_perm( lds32( ring_cap ), v, 0x7610 );
    };

    for( ;; t0 += BLK ) {
        stage( t0 + 2 * BLK );
        cp_async_wait<2>();     // everything but the two youngest blocks has landed: block t0 is complete
        __syncwarp();
        {   // the bottom lane has passed the last real step (of every array of the warp)
            const bool fin = ws_get( WS_DRY ) && t0 >= ws_get( WS_LAST ) + BLK;
            if( G == 32 ? fin : __all_sync( FULL, fin ) ) break;
        }

#pragma unroll 1
        for( uint32_t t = t0; t < t0 + BLK; t++ ) {
            const uint32_t slot4 = ( ( t - gl ) & ( RING - 1 ) ) * 4;   // lane l of an array works on virtual step t - l
            const uint32_t w = lds32( ring_w + slot4 );
            const uint32_t cia = w & 0xffffu, cib = w >> 16;

            T ha_in = __shfl_up_sync( FULL, ha_out, 1, G );
            T fa_in = __shfl_up_sync( FULL, fa_out, 1, G );
            T hb_in = __shfl_up_sync( FULL, hb_out, 1, G );
            T fb_in = __shfl_up_sync( FULL, fb_out, 1, G );
            T s_in = 0;
            if( !NW ) s_in = __shfl_up_sync( FULL, s_out, 1, G );
            if( is_top ) {
                if( MULTIPASS && !( NW && first_pass ) ) {
                    const uint4 b = lds128( ring_b + slot4 * 4 );
                    ha_in = b.x;
                    hb_in = b.y;
                    fa_in = b.z;
                    fb_in = b.w;
                }
                else if( NW ) {
                    if( cia & COL_FIRST ) top_h = A::bcast( a.gap_open + a.gap_extend - a.offset );
                    ha_in = top_h;
                    fa_in = A::add( top_h, QR );
                    hb_in = A::add( top_h, RG );
                    fb_in = A::add( hb_in, QR );
                    top_h = A::add( hb_in, RG );
                }
                else {
                    ha_in = 0;
                    hb_in = 0;
                    fa_in = 0;
                    fb_in = 0;
                }
                s_in = 0;
            }

            s_out = 0;
            if( cia & COL_VALID ) {
                T ha_last, fa_last, hb_last, fb_last;   // (an idle step leaves the hand-over registers alone: the next lane's step is idle too)
                if( cia & COL_FIRST ) {   // a new pair of sequences starts in this lane (always in column A)
                    if( NW ) {   // H(i, -1) and E of this lane's rows from the pass' table, H(row0 - 1, -1) as the first diagonal
#pragma unroll
                        for( int q = 0; q < R / 4; q++ ) {
                            const uint4 h = lds128( nw_tab + q * 512 ), e = lds128( nw_tab + R * 128 + q * 512 );
                            Hp[4 * q] = h.x;
                            Hp[4 * q + 1] = h.y;
                            Hp[4 * q + 2] = h.z;
                            Hp[4 * q + 3] = h.w;
                            E[4 * q] = e.x;
                            E[4 * q + 1] = e.y;
                            E[4 * q + 2] = e.z;
                            E[4 * q + 3] = e.w;
                        }
                        const int pos0 = a.rowbase + (int) pp * 32 * R + gl * R;   // first diagonal: H of the position above, column -1
                        const int d1 = nw_start( pos0 - 1, a.nw_shift, a.gap_open, a.gap_extend ) - a.offset;
                        const int d2 = DUAL ? nw_start( pos0 - 1, a.nw_shift2, a.gap_open, a.gap_extend ) - a.offset : d1;
                        diag = ( (uint32_t) (uint16_t) (int16_t) d1 << 16 ) | (uint32_t) (uint16_t) (int16_t) d2;
                    }
                    else {
#pragma unroll
                        for( int r = 0; r < R; r++ ) {
                            Hp[r] = 0;
                            E[r] = 0;
                        }
                        diag = 0;
                        S = 0;
                    }
                }
                const uint32_t ra = ( cia >> hs ) & 31u, rb = ( cia >> 5 ) & 31u, sa = ( cib >> hs ) & 31u, sb = ( cib >> 5 ) & 31u;
                {
                // ---- two columns per step (see score_kernel.cuh for the recurrence and its instruction selection) ----
                // Full-warp arrays with 64-bit profile entries (R >= 20, the shapes of long queries): every diagonal candidate of
                // column A (old H of the row above + substitution score) is formed BEFORE the pair of rows that overwrites that H
                // register - the first candidate of the NEXT quad of rows from the next quad's profile entry, fetched one quad
                // ahead - so no H value has to outlive its register and none needs a copy: 8 register moves fewer per step
                // (+1.3 % SW, +2.9 % NW on the 5147-aa query [B200]).  The other shapes keep the order below: on half-warp arrays
                // the same change costs 2.7 % (16 x 24 rows, 375 aa) and 5 % with two queries per word.
                constexpr bool EARLY_CANDIDATES = G == 32 && !DUAL && MR == 0;
                T fa = fa_in, fb = fb_in;
                T ha_prev = ha_in;
                T hda = diag;
                diag = hb_in;
                if( EARLY_CANDIDATES ) {
                auto rows2 = [&]( const int r, const T ta0, const T ta1, const uint32_t vb0, const uint32_t vb1 ) {
                    T hq;
                    const T ha0 = NW ? A::max3( ta0, E[r], fa ) : A::max3_relu( ta0, E[r], fa );
                    hq = A::add( ha0, QR );
                    const T ea0 = A::addmax( E[r], RG, hq );
                    fa = A::addmax( fa, RG, hq );
                    const T tb0 = A::add( ha_prev, vb0 );
                    Hp[r] = NW ? A::max3( tb0, ea0, fb ) : A::max3_relu( tb0, ea0, fb );
                    hq = A::add( Hp[r], QR );
                    E[r] = A::addmax( ea0, RG, hq );
                    fb = A::addmax( fb, RG, hq );
                    const T ha1 = NW ? A::max3( ta1, E[r + 1], fa ) : A::max3_relu( ta1, E[r + 1], fa );
                    hq = A::add( ha1, QR );
                    const T ea1 = A::addmax( E[r + 1], RG, hq );
                    fa = A::addmax( fa, RG, hq );
                    const T tb1 = A::add( ha0, vb1 );
                    Hp[r + 1] = NW ? A::max3( tb1, ea1, fb ) : A::max3_relu( tb1, ea1, fb );
                    hq = A::add( Hp[r + 1], QR );
                    E[r + 1] = A::addmax( ea1, RG, hq );
                    fb = A::addmax( fb, RG, hq );
                    ha_prev = ha1;
                    if( !NW ) {
                        S = A::max3( S, ta0, ta1 );
                        S = A::max3( S, tb0, tb1 );
                    }
                    else {
                        // Needleman-Wunsch has no running maximum, and a diagonal candidate with a single consumer is fused by
                        // ptxas into VIADDMNMX + VIMNMX - two instructions of the binding ALU pipe instead of VIADD.16x2 (FMA-side
                        // pipe) + VIMNMX3 (5 instead of 4 ALU-pipe instructions per word).  One multiply-add per two words on the
                        // FMA-side pipe keeps the candidates alive as values (its result is never used for anything real).
                        S = ta0 * S + ta1;
                        S = tb0 * S + tb1;
                    }
                };
                {   // 64-bit entries: four rows, two scores per word; PRMT interleaves A and B
                    uint2 laa = lds64( prof_base16 + ( ra << 8 ) ), lba = lds64( prof_base16 + ( rb << 8 ) );
                    T ta_first = A::add( hda, __byte_perm( lba.x, laa.x, 0x5410 ) );
#pragma unroll
                    for( int q = 0; q < NQ16; q++ ) {
                        const uint2 lab = lds64( prof_base16 + ( sa << 8 ) + q * 8192 ), lbb = lds64( prof_base16 + ( sb << 8 ) + q * 8192 );
                        const int r = 4 * q;
                        const T ta0 = ta_first;
                        const T ta1 = A::add( Hp[r], __byte_perm( lba.x, laa.x, 0x7632 ) );
                        const T ta2 = A::add( Hp[r + 1], __byte_perm( lba.y, laa.y, 0x5410 ) );
                        const T ta3 = A::add( Hp[r + 2], __byte_perm( lba.y, laa.y, 0x7632 ) );
                        if( q + 1 < NQ16 ) {
                            laa = lds64( prof_base16 + ( ra << 8 ) + ( q + 1 ) * 8192 );
                            lba = lds64( prof_base16 + ( rb << 8 ) + ( q + 1 ) * 8192 );
                            ta_first = A::add( Hp[r + 3], __byte_perm( lba.x, laa.x, 0x5410 ) );
                        }
                        rows2( r, ta0, ta1, __byte_perm( lbb.x, lab.x, 0x5410 ), __byte_perm( lbb.x, lab.x, 0x7632 ) );
                        rows2( r + 2, ta2, ta3, __byte_perm( lbb.y, lab.y, 0x5410 ), __byte_perm( lbb.y, lab.y, 0x7632 ) );
                    }
                }
                }
                else {
                auto dp_pair = [&]( const int k, const uint32_t va0, const uint32_t va1, const uint32_t vb0, const uint32_t vb1 ) {
                    const T ta0 = A::add( hda, va0 );
                    const T ta1 = A::add( Hp[2 * k], va1 );
                    hda = Hp[2 * k + 1];
                    T hq;
                    const T ha0 = NW ? A::max3( ta0, E[2 * k], fa ) : A::max3_relu( ta0, E[2 * k], fa );
                    hq = A::add( ha0, QR );
                    const T ea0 = A::addmax( E[2 * k], RG, hq );
                    fa = A::addmax( fa, RG, hq );
                    const T tb0 = A::add( ha_prev, vb0 );
                    Hp[2 * k] = NW ? A::max3( tb0, ea0, fb ) : A::max3_relu( tb0, ea0, fb );
                    hq = A::add( Hp[2 * k], QR );
                    E[2 * k] = A::addmax( ea0, RG, hq );
                    fb = A::addmax( fb, RG, hq );
                    const T ha1 = NW ? A::max3( ta1, E[2 * k + 1], fa ) : A::max3_relu( ta1, E[2 * k + 1], fa );
                    hq = A::add( ha1, QR );
                    const T ea1 = A::addmax( E[2 * k + 1], RG, hq );
                    fa = A::addmax( fa, RG, hq );
                    const T tb1 = A::add( ha0, vb1 );
                    Hp[2 * k + 1] = NW ? A::max3( tb1, ea1, fb ) : A::max3_relu( tb1, ea1, fb );
                    hq = A::add( Hp[2 * k + 1], QR );
                    E[2 * k + 1] = A::addmax( ea1, RG, hq );
                    fb = A::addmax( fb, RG, hq );
                    ha_prev = ha1;
                    if( !NW ) {
                        S = A::max3( S, ta0, ta1 );
                        S = A::max3( S, tb0, tb1 );
                    }
                    else {
                        // Needleman-Wunsch has no running maximum, and a diagonal candidate with a single consumer is fused by
                        // ptxas into VIADDMNMX + VIMNMX - two instructions of the binding ALU pipe instead of VIADD.16x2 (FMA-side
                        // pipe) + VIMNMX3 (5 instead of 4 ALU-pipe instructions per word).  One multiply-add per two words on the
                        // FMA-side pipe keeps the candidates alive as values (its result is never used for anything real).
                        S = ta0 * S + ta1;
                        S = tb0 * S + tb1;
                    }
                };
#pragma unroll
                for( int q = 0; q < MR / 4; q++ ) {
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
                for( int q = 0; q < NQ16; q++ ) {
                    const uint2 laa = lds64( prof_base16 + ( ra << 8 ) + q * 8192 ), lba = lds64( prof_base16 + ( rb << 8 ) + q * 8192 );
                    const uint2 lab = lds64( prof_base16 + ( sa << 8 ) + q * 8192 ), lbb = lds64( prof_base16 + ( sb << 8 ) + q * 8192 );
                    dp_pair( MR / 2 + 2 * q, __byte_perm( lba.x, laa.x, 0x5410 ), __byte_perm( lba.x, laa.x, 0x7632 ),
                            __byte_perm( lbb.x, lab.x, 0x5410 ), __byte_perm( lbb.x, lab.x, 0x7632 ) );
                    dp_pair( MR / 2 + 2 * q + 1, __byte_perm( lba.y, laa.y, 0x5410 ), __byte_perm( lba.y, laa.y, 0x7632 ),
                            __byte_perm( lbb.y, lab.y, 0x5410 ), __byte_perm( lbb.y, lab.y, 0x7632 ) );
                }
                }
                ha_last = ha_prev;
                fa_last = fa;
                hb_last = Hp[R - 1];
                fb_last = fb;
                if( NW && is_cap ) {
                    nw_capture( cia, ha_last, slot4 );
                    nw_capture( cib, hb_last, slot4 );
                }
                }

                // ---- the pair ends in this step: the running maximum travels down the lanes with the wavefront ----
                if( !NW && ( ( cia | cib ) & end_mask ) ) {
                    s_out = A::max2( s_in, S );
                    if( is_bottom ) {   // (the earlier pass may have run on another SM: read around L1)
                        uint32_t * o = out + lds32( ring_g + slot4 );
                        *o = first_pass ? s_out : __vmaxs2( __ldcg( o ), s_out );
                    }
                }
                if( MULTIPASS && is_bottom && bnd_out ) {   // this lane's last row is the next pass' top boundary
                    bnd_out[lds32( ring_g + slot4 )] = make_uint4( ha_last, hb_last, fa_last, fb_last );
                    if( publish ) {   // last step of a batch: the same batch of the next pass may start
                        const uint32_t e = lds32( ring_e + slot4 );
                        if( e ) {
                            __threadfence();
                            st_release( a.done + ( e - 1 ), pp + 1 );
                        }
                    }
                }
                ha_out = ha_last;
                fa_out = fa_last;
                hb_out = hb_last;
                fb_out = fb_last;
            }
        }
    }
  }
    cp_async_wait<0>();
    if( NW && S == 0x9e3779b9u && a.nbatches == 0xffffffffu ) a.out_by_col[0] = S;   // never true: keeps the candidate chain alive
}

}  // namespace ssa
