// The exact 32-bit tier for LONG queries: all passes of a sequence run at the same time, as a pipeline.
//
// The 32-bit tier is the counterpart of the reference's 16 -> 64 bit recomputation (src/algo/64/*.c, the overflow lists of
// src/algo/searcher.c): the few sequences whose scores leave the int16 range.  A query longer than one systolic array is
// processed in passes of 32 lanes x R rows; score_kernel<PACKED = false> does one launch per pass, and within a pass one
// warp owns one sequence - a handful of 10 kbp sequences then keeps a handful of warps busy for `passes` launches in a row
// (ten 10 kbp x 10 kbp pairs: 54 ms [B200]).  Here a task is (group of 8 sequences, pass): a CTA takes tasks from one queue
// ordered (group, pass 0), (group, pass 1), ..., loads the pass' query profile and gives each of its 8 warps one sequence;
// pass p of a sequence starts as soon as pass p-1 has written the first boundary entries it needs and then follows it
// block by block (a per-(sequence, pass) progress counter, release / acquire) - the passes of one sequence are a pipeline
// of warps on different multiprocessors, 64-95 steps apart.  A wait only ever points to a task that was taken from the
// queue earlier, by a CTA that is therefore resident, so the chain of waits ends at a pass 0 and cannot deadlock.
//
// The recurrence is the reference's (Gotoh, affine gaps; SW clamps at 0 and reports the maximum, NW reports H of the last
// cell); two columns per step, one sequence per warp; same stream, batches and profile layout as score_kernel<false>.
#include "score_launch.h"
#include "sw32_kernel.cuh"

#include <cstdio>
#include <cstdlib>

namespace ssa {

constexpr int WIDE_PIPE_SEQS = SCORE_THREADS / 32;   // sequences per task = warps per CTA

template<int R, bool NW>
__global__ void __launch_bounds__( SCORE_THREADS, 2 ) wide_pipe_kernel( const WidePipeArgs a ) {
    typedef int32_t T;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int PW = R * 1024;   // profile words of one pass: [row][residue][lane]
    extern __shared__ uint32_t prof[];
    __shared__ uint32_t s_task;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t nb = a.nbatches_dev ? *a.nbatches_dev : a.nbatches;
    const uint32_t ntasks = ( ( nb + WIDE_PIPE_SEQS - 1 ) / WIDE_PIPE_SEQS ) * a.npass;
    const uint32_t prof_base = (uint32_t) __cvta_generic_to_shared( prof ) + lane * 4;
    const T QR = a.gap_open + a.gap_extend, RG = a.gap_extend;

    for( ;; ) {
        __syncthreads();   // every warp is done with the previous task's profile (and has read s_task)
        if( threadIdx.x == 0 ) s_task = atomicAdd( a.counter, 1u );
        __syncthreads();
        const uint32_t task = s_task;
        if( task >= ntasks ) break;
        const uint32_t p = task % a.npass, b = ( task / a.npass ) * WIDE_PIPE_SEQS + warp;
        {
            const uint32_t * src = a.profile + (size_t) p * PW;
            for( int i = threadIdx.x; i < PW; i += blockDim.x ) prof[i] = src[i];
        }
        __syncthreads();
        if( b >= nb ) continue;   // (warp-uniform; the CTA barriers at the top of the loop are reached by every warp)

        const Batch bt = a.batches[b];
        const uint32_t * st = (const uint32_t *) ( a.stream + bt.col0 );   // two 16-bit column words per element
        const int ncp = (int) ( bt.ncols >> 1 ), nsteps = ncp + 31;
        const uint32_t half_shift = bt.half ? 5u : 0u, last_mask = bt.half ? COL_LAST_B : COL_LAST_A;
        const uint32_t bnd_base = (uint32_t) ( bt.bnd0 >> 1 );
        // pass p writes buffer p & 1 and reads what pass p-1 wrote; pass p+2 reuses the buffer of pass p, always behind the reads of p+1
        const uint4 * bnd_in = p > 0 ? ( ( ( p - 1 ) & 1 ) ? a.bnd1 : a.bnd0 ) + bnd_base : nullptr;
        uint4 * bnd_out = p + 1 < a.npass ? ( ( p & 1 ) ? a.bnd1 : a.bnd0 ) + bnd_base : nullptr;
        const uint32_t * prog_in = p > 0 ? a.progress + (size_t) b * a.npass + ( p - 1 ) : nullptr;
        uint32_t * prog_out = p + 1 < a.npass ? a.progress + (size_t) b * a.npass + p : nullptr;
        const int row0 = (int) p * 32 * R + lane * R;
        const bool is_cap = NW && (int) p == a.cap_pass && lane == a.cap_lane;

        T Hp[R], E[R];
#pragma unroll
        for( int r = 0; r < R; r++ ) {   // column -1: SW zeros; NW H(i, -1) = gapO + (i + 1) gapE, E entering column 0 = H + gapO + gapE
            Hp[r] = NW ? a.gap_open + ( row0 + r + 1 ) * a.gap_extend : 0;
            E[r] = NW ? Hp[r] + QR : 0;
        }
        T diag = NW ? ( row0 == 0 ? 0 : a.gap_open + row0 * a.gap_extend ) : 0;   // H(row0 - 1, -1)
        T S = 0;
        T ha_out = 0, fa_out = 0, hb_out = 0, fb_out = 0, s_out = 0;

        // blocks of 32 boundary entries: lane l holds the entry of column pair 32 * block + l; the top lane gets its entry by shuffle
        auto wait_for = [&]( int upto ) {   // the previous pass has written the boundary of column pairs [0, min( upto, ncp ))
            const uint32_t need = (uint32_t) min( upto, ncp );
            if( lane == 0 )
                while( ld_acquire( prog_in ) < need ) __nanosleep( 100 );
            __syncwarp();
        };
        uint4 b_cur = make_uint4( 0, 0, 0, 0 ), b_nxt = make_uint4( 0, 0, 0, 0 );
        if( bnd_in ) {
            wait_for( 64 );
            if( lane < ncp ) b_cur = __ldcg( bnd_in + lane );
            if( 32 + lane < ncp ) b_nxt = __ldcg( bnd_in + 32 + lane );
        }
        uint32_t w_n1 = lane == 0 ? st[0] : 0, w_n2 = ( 1 - lane >= 0 && 1 - lane < ncp ) ? st[1 - lane] : 0;

        // one database column through this lane's R rows
        auto column = [&]( const uint32_t d, const T h_in, const T f_in, T & h_last, T & f_last ) {
            T f = f_in, hd = diag;
            diag = h_in;
            const uint32_t pa = prof_base + ( d << 7 );
            T colmax = 0;
#pragma unroll
            for( int r = 0; r < R; r++ ) {
                const T v = (T) lds32( pa + r * 4096 );
                const T tt = hd + v;
                hd = Hp[r];
                const T h = NW ? __vimax3_s32( tt, E[r], f ) : __vimax3_s32_relu( tt, E[r], f );
                const T hq = h + QR;
                E[r] = __viaddmax_s32( E[r], RG, hq );
                f = __viaddmax_s32( f, RG, hq );
                Hp[r] = h;
                if( !NW ) colmax = max( colmax, h );
            }
            if( !NW ) S = max( S, colmax );
            h_last = Hp[R - 1];
            f_last = f;
        };
        auto nw_capture = [&]( const uint32_t ci ) {   // H of the query's last row at the sequence's last column
            if( is_cap && ( ci & last_mask ) ) {
                T sel = Hp[0];
#pragma unroll
                for( int r = 1; r < R; r++ ) {
                    T cand = Hp[r];
                    asm volatile( "" : "+r"( cand ) );
                    if( r == a.cap_row ) sel = cand;
                    asm volatile( "" : "+r"( sel ) );
                }
                a.out[bt.out0] = sel;
            }
        };

#pragma unroll 1
        for( int t = 0; t < nsteps; t++ ) {
            const uint32_t w = w_n1;
            w_n1 = w_n2;
            {
                const int c2 = t + 2 - lane;
                w_n2 = ( c2 >= 0 && c2 < ncp ) ? st[c2] : 0;
            }
            const uint32_t cia = w & 0xffffu, cib = w >> 16;
            T ha_in = __shfl_up_sync( FULL, ha_out, 1 ), fa_in = __shfl_up_sync( FULL, fa_out, 1 );
            T hb_in = __shfl_up_sync( FULL, hb_out, 1 ), fb_in = __shfl_up_sync( FULL, fb_out, 1 );
            T s_in = NW ? 0 : __shfl_up_sync( FULL, s_out, 1 );
            if( bnd_in ) {   // warp-uniform: t is the column pair of the top lane
                if( ( t & 31 ) == 0 && t > 0 ) {
                    b_cur = b_nxt;
                    if( t + 32 < ncp ) {
                        wait_for( t + 64 );
                        if( t + 32 + lane < ncp ) b_nxt = __ldcg( bnd_in + t + 32 + lane );
                    }
                }
                const uint32_t bha = __shfl_sync( FULL, b_cur.x, t & 31 ), bhb = __shfl_sync( FULL, b_cur.y, t & 31 );
                const uint32_t bfa = __shfl_sync( FULL, b_cur.z, t & 31 ), bfb = __shfl_sync( FULL, b_cur.w, t & 31 );
                if( lane == 0 ) {
                    ha_in = (T) bha;
                    hb_in = (T) bhb;
                    fa_in = (T) bfa;
                    fb_in = (T) bfb;
                }
            }
            else if( lane == 0 ) {   // the matrix' top boundary: SW zeros, NW H(-1, j) = gapO + (j + 1) gapE, F = H + gapO + gapE
                ha_in = NW ? a.gap_open + ( 2 * t + 1 ) * a.gap_extend : 0;
                hb_in = NW ? ha_in + RG : 0;
                fa_in = NW ? ha_in + QR : 0;
                fb_in = NW ? hb_in + QR : 0;
            }
            if( lane == 0 ) s_in = 0;

            T ha_last = 0, fa_last = 0, hb_last = 0, fb_last = 0;
            s_out = 0;
            if( cia & COL_VALID ) {
                column( ( cia >> half_shift ) & 31u, ha_in, fa_in, ha_last, fa_last );
                if( NW ) nw_capture( cia );
                column( ( cib >> half_shift ) & 31u, hb_in, fb_in, hb_last, fb_last );
                if( NW ) nw_capture( cib );
                if( !NW && ( ( cia | cib ) & last_mask ) ) {   // the sequence ends in this step: the maximum travels down the lanes
                    s_out = max( s_in, S );
                    S = 0;   // (columns behind the end only hold padding)
                    if( lane == 31 ) atomicMax( a.out + bt.out0, s_out );
                }
                if( lane == 31 && bnd_out ) {   // this lane's last row is the next pass' top boundary
                    const int c = t - 31;
                    bnd_out[c] = make_uint4( (uint32_t) ha_last, (uint32_t) hb_last, (uint32_t) fa_last, (uint32_t) fb_last );
                    if( ( ( c + 1 ) & 31 ) == 0 || c + 1 == ncp ) {
                        __threadfence();
                        st_release( prog_out, (uint32_t) ( c + 1 ) );
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

template<int R, bool NW> static int launch( const WidePipeArgs & a, cudaStream_t stream ) {
    const int smem = R * 1024 * 4;
    const std::pair<int, int> lim = kernel_limits( wide_pipe_kernel<R, NW>, smem );
    const long tasks = ( ( (long) a.nbatches + WIDE_PIPE_SEQS - 1 ) / WIDE_PIPE_SEQS ) * a.npass;
    int grid = lim.first * lim.second;
    if( tasks < grid ) grid = (int) ( tasks < 1 ? 1 : tasks );
    wide_pipe_kernel<R, NW><<<grid, SCORE_THREADS, smem, stream>>>( a );
    return grid;
}

int launch_wide_pipe( bool nw, const WidePipeArgs & a, cudaStream_t s ) {
    return nw ? launch<WIDE_PIPE_R, true>( a, s ) : launch<WIDE_PIPE_R, false>( a, s );
}

}  // namespace ssa
