// Smith-Waterman scoring on full-warp systolic arrays, packed tier (two database sequences per int16 x 2 word): the hot
// kernel of protein searches (reference src/algo/simd/search_simd_sw.c:172-441 and its chunk loop).  Same recurrence and
// same data as score_kernel.cuh (which keeps the narrow groups, Needleman-Wunsch, the 32-bit tier and the tracker); what is
// different is everything AROUND the recurrence - that was 143 of 591 instructions per step of 64 packed words:
//
//   * one warp = one array of 32 lanes x R rows, known at compile time;
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
//   * there is no one-column path (pairs are padded with a PAD_RESIDUE column) and no per-step exit vote;
//   * ALL full passes of a long query run in ONE launch: a CTA works through the passes in order (its warps meet at a CTA
//     barrier to swap the query profile in shared memory), every pass has its own work queue over the same batches, and a
//     batch of pass p+1 starts as soon as the SAME batch of pass p is finished (a per-batch counter in global memory,
//     release / acquire) - there is no grid-wide barrier between passes, so the idle tail of a pass (the longest pairs, the
//     last small batches; 10-15 % of a pass on a 70 000-record shard) is filled with the next pass' work.  Waits only ever
//     point to a lower pass of a batch some running warp has already claimed, so they cannot deadlock.
#pragma once

#include "score_kernel.cuh"

namespace ssa {

constexpr int SW32_RING = 128;        // steps a warp's ring holds (power of two): 64 ahead of the top lane, 31 behind it, one block of slack
constexpr int SW32_MIN_STEPS = 64;    // virtual length of a batch: a block of 32 steps crosses at most one batch boundary
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
};

// words of shared memory per warp for the ring: column words, stream positions, boundary entries
template<bool MULTIPASS> __host__ __device__ constexpr int sw32_ring_words() { return 2 * SW32_RING + ( MULTIPASS ? 5 * SW32_RING : 0 ); }
template<int R, bool MULTIPASS> __host__ __device__ constexpr int sw32_smem_bytes() {
    return ( profile_words<true, R>() + ( SCORE_THREADS / 32 ) * sw32_ring_words<MULTIPASS>() ) * 4;
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

template<int R, bool MULTIPASS>
__global__ void __launch_bounds__( SCORE_THREADS, min_blocks<true, R>() ) sw32_kernel( const Sw32Args a ) {
    typedef Arith<true> A;
    typedef uint32_t T;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int PW = profile_words<true, R>();
    constexpr int MR = merge_rows( R );
    constexpr int NQ16 = ( R - MR ) / 4;
    extern __shared__ uint32_t smem[];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    {   // every ring slot idle: lanes below the top one start on steps "before the launch"
        uint32_t * ring = smem + PW + warp * sw32_ring_words<MULTIPASS>();
        for( int i = lane; i < sw32_ring_words<MULTIPASS>(); i += 32 ) ring[i] = 0;
    }

    const uint32_t smem_base = (uint32_t) __cvta_generic_to_shared( smem );
    const uint32_t ring_w = smem_base + ( PW + warp * sw32_ring_words<MULTIPASS>() ) * 4;   // [RING] column words of the step
    const uint32_t ring_g = ring_w + SW32_RING * 4;                                          // [RING] column pair index in the stream
    const uint32_t ring_b = ring_g + SW32_RING * 4;                                          // [RING] boundary entry (MULTIPASS)
    const uint32_t ring_e = ring_b + SW32_RING * 16;                                         // [RING] batch + 1 on its last step (MULTIPASS)
    const uint32_t prof_base = smem_base + lane * 16;            // 128-bit entries of the first MR rows
    const uint32_t prof_base16 = smem_base + MR * 4096 + lane * 8;   // 64-bit entries of the other rows
    const uint32_t K65536 = a.k65536;
    const T QR = a.qr_word, RG = a.rg_word;
    const bool is_top = lane == 0, is_bottom = lane == 31;

    T Hp[R], E[R];
#pragma unroll
    for( int r = 0; r < R; r++ ) {
        Hp[r] = 0;
        E[r] = 0;
    }
    T S = 0, diag = 0;
    T ha_out = 0, fa_out = 0, hb_out = 0, fb_out = 0, s_out = 0;

    // ---- the warp's batches on its virtual step axis (warp-uniform) ----
    // cur / nxt: the two batches the staging front can touch.  A batch of n column pairs occupies max( n, SW32_MIN_STEPS )
    // virtual steps; steps past its columns, and everything after the queue ran dry, are idle.
    uint32_t cur_s = 0, cur_e = 0, cur_n = 0, cur_g = 0, cur_b = 0;   // virtual begin / end, column pairs, first column pair in the stream, batch
    uint32_t nxt_e = 0, nxt_n = 0, nxt_g = 0, nxt_b = 0;              // (begins at cur_e)
    uint32_t last_real_e = 0;                                 // virtual end of the last real batch
    bool dry = false;                                         // the work queue of the current pass is exhausted
    uint32_t pp = 0;                                          // pass of this launch the CTA works on
    uint32_t * counter = a.counters;
    const uint4 * bnd_in = nullptr;
    uint4 * bnd_out = nullptr;
    bool first_pass = true, publish = false;

    auto fetch = [&]( uint32_t begin, uint32_t & e, uint32_t & n, uint32_t & g, uint32_t & bidx ) {
        uint32_t b = 0xffffffffu;
        if( !dry ) {
            if( is_top ) b = atomicAdd( counter, 1u );
            b = __shfl_sync( FULL, b, 0 );
        }
        if( b >= a.nbatches ) {
            dry = true;
            n = 0;
            g = 0;
            bidx = 0;
            e = begin + SW32_MIN_STEPS;
        }
        else {
            const Batch bt = a.batches[b];
            n = bt.ncols >> 1;
            g = (uint32_t) ( bt.col0 >> 1 );
            bidx = b;
            e = begin + max( n, (uint32_t) SW32_MIN_STEPS );
            last_real_e = e;
            if( MULTIPASS && pp > 0 ) {   // the previous pass of this launch must have finished this batch (its boundary row, its result)
                if( is_top )
                    while( ld_acquire( a.done + b ) < pp ) __nanosleep( 200 );
                __syncwarp();
            }
        }
    };

    // stages the 32 virtual steps [v0, v0 + 32): lane l prepares step v0 + l
    auto stage = [&]( const uint32_t v0 ) {
        const uint32_t v = v0 + lane;
        const uint32_t slot = v & ( SW32_RING - 1 );
        uint32_t c, n, g, bidx;
        if( v < cur_e ) {
            c = v - cur_s;
            n = cur_n;
            g = cur_g;
            bidx = cur_b;
        }
        else {
            c = v - cur_e;
            n = nxt_n;
            g = nxt_g;
            bidx = nxt_b;
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
        if( v0 + 32 >= cur_e ) {   // the front leaves the current batch: the next one becomes current, another is pulled
            cur_s = cur_e;
            cur_e = nxt_e;
            cur_n = nxt_n;
            cur_g = nxt_g;
            cur_b = nxt_b;
            fetch( cur_e, nxt_e, nxt_n, nxt_g, nxt_b );
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
    __syncthreads();
    {
        const uint32_t abs_pass = a.pass0 + pp;
        first_pass = abs_pass == 0;
        bnd_in = MULTIPASS && abs_pass > 0 ? ( ( abs_pass & 1 ) ? a.bnd0 : a.bnd1 ) : nullptr;
        bnd_out = MULTIPASS && ( pp + 1 < a.npass || a.more_passes ) ? ( ( abs_pass & 1 ) ? a.bnd1 : a.bnd0 ) : nullptr;
        publish = MULTIPASS && pp + 1 < a.npass;
        counter = a.counters + pp;
    }
    // the virtual axis simply continues: the steps already staged beyond t0 are idle ones and are staged again
    dry = false;
    cur_s = t0;
    fetch( t0, cur_e, cur_n, cur_g, cur_b );
    fetch( cur_e, nxt_e, nxt_n, nxt_g, nxt_b );
    stage( t0 );
    stage( t0 + 32 );

    for( ;; t0 += 32 ) {
        stage( t0 + 64 );
        cp_async_wait<2>();     // everything but the two youngest blocks has landed: block t0 is complete
        __syncwarp();
        if( dry && t0 >= last_real_e + 32 ) break;   // the bottom lane has passed the last real step

#pragma unroll 1
        for( uint32_t t = t0; t < t0 + 32; t++ ) {
            const uint32_t slot4 = ( ( t - lane ) & ( SW32_RING - 1 ) ) * 4;   // lane l works on virtual step t - l
            const uint32_t w = lds32( ring_w + slot4 );
            const uint32_t cia = w & 0xffffu, cib = w >> 16;

            T ha_in = __shfl_up_sync( FULL, ha_out, 1 );
            T fa_in = __shfl_up_sync( FULL, fa_out, 1 );
            T hb_in = __shfl_up_sync( FULL, hb_out, 1 );
            T fb_in = __shfl_up_sync( FULL, fb_out, 1 );
            T s_in = __shfl_up_sync( FULL, s_out, 1 );
            if( is_top ) {
                if( MULTIPASS ) {
                    const uint4 b = lds128( ring_b + slot4 * 4 );
                    ha_in = b.x;
                    hb_in = b.y;
                    fa_in = b.z;
                    fb_in = b.w;
                }
                else {
                    ha_in = 0;
                    hb_in = 0;
                    fa_in = 0;
                    fb_in = 0;
                }
                s_in = 0;
            }

            T ha_last = 0, fa_last = 0, hb_last = 0, fb_last = 0;
            s_out = 0;
            if( cia & COL_VALID ) {
                if( cia & COL_FIRST ) {   // a new pair of sequences starts in this lane (always in column A)
#pragma unroll
                    for( int r = 0; r < R; r++ ) {
                        Hp[r] = 0;
                        E[r] = 0;
                    }
                    diag = 0;
                    S = 0;
                }
                // ---- two columns per step (see score_kernel.cuh for the recurrence and its instruction selection) ----
                const uint32_t ra = cia & 31u, rb = ( cia >> 5 ) & 31u, sa = cib & 31u, sb = ( cib >> 5 ) & 31u;
                T fa = fa_in, fb = fb_in;
                T ha_prev = ha_in;
                T hda = diag;
                diag = hb_in;
                auto dp_pair = [&]( const int k, const uint32_t va0, const uint32_t va1, const uint32_t vb0, const uint32_t vb1 ) {
                    const T ta0 = A::add( hda, va0 );
                    const T ta1 = A::add( Hp[2 * k], va1 );
                    hda = Hp[2 * k + 1];
                    T hq;
                    const T ha0 = A::max3_relu( ta0, E[2 * k], fa );
                    hq = A::add( ha0, QR );
                    const T ea0 = A::addmax( E[2 * k], RG, hq );
                    fa = A::addmax( fa, RG, hq );
                    const T tb0 = A::add( ha_prev, vb0 );
                    Hp[2 * k] = A::max3_relu( tb0, ea0, fb );
                    hq = A::add( Hp[2 * k], QR );
                    E[2 * k] = A::addmax( ea0, RG, hq );
                    fb = A::addmax( fb, RG, hq );
                    const T ha1 = A::max3_relu( ta1, E[2 * k + 1], fa );
                    hq = A::add( ha1, QR );
                    const T ea1 = A::addmax( E[2 * k + 1], RG, hq );
                    fa = A::addmax( fa, RG, hq );
                    const T tb1 = A::add( ha0, vb1 );
                    Hp[2 * k + 1] = A::max3_relu( tb1, ea1, fb );
                    hq = A::add( Hp[2 * k + 1], QR );
                    E[2 * k + 1] = A::addmax( ea1, RG, hq );
                    fb = A::addmax( fb, RG, hq );
                    ha_prev = ha1;
                    S = A::max3( S, ta0, ta1 );
                    S = A::max3( S, tb0, tb1 );
                };
#pragma unroll
                for( int q = 0; q < MR / 4; q++ ) {
                    const uint4 laa = lds128( prof_base + ( ra << 9 ) + q * 16384 ), lba = lds128( prof_base + ( rb << 9 ) + q * 16384 );
                    const uint4 lab = lds128( prof_base + ( sa << 9 ) + q * 16384 ), lbb = lds128( prof_base + ( sb << 9 ) + q * 16384 );
                    dp_pair( 2 * q, laa.x * K65536 + lba.x, laa.y * K65536 + lba.y, lab.x * K65536 + lbb.x, lab.y * K65536 + lbb.y );
                    dp_pair( 2 * q + 1, laa.z * K65536 + lba.z, laa.w * K65536 + lba.w, lab.z * K65536 + lbb.z, lab.w * K65536 + lbb.w );
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
                ha_last = ha_prev;
                fa_last = fa;
                hb_last = Hp[R - 1];
                fb_last = fb;

                // ---- the pair ends in this step: the running maximum travels down the lanes with the wavefront ----
                if( ( cia | cib ) & COL_LAST_A ) {
                    s_out = A::max2( s_in, S );
                    if( is_bottom ) {   // (the earlier pass may have run on another SM: read around L1)
                        uint32_t * o = a.out_by_col + lds32( ring_g + slot4 );
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
            }
            ha_out = ha_last;
            fa_out = fa_last;
            hb_out = hb_last;
            fb_out = fb_last;
        }
    }
  }
    cp_async_wait<0>();
}

}  // namespace ssa
