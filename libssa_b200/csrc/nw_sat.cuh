// Emulation of ONE channel of the reference's saturating 8-/16-bit NW kernel per database sequence
// (reference src/algo/simd/search_simd_nw.c:180-316 ALIGNCORE / aligncolumns_first / aligncolumns_rest,
// :390-500 boundary injection, per-block overflow test, result read-out).  It reproduces two things the exact
// tiers cannot know.  (1) The ORDER in which the reference's narrow NW tiers hand records to the heap: a record
// that overflows a tier leaves its SIMD channel after the 4-column block in which the overflow is detected and
// is re-run in the next tier after the chunk.  (2) The SCORE a narrow tier reports for a record it does NOT flag:
// the injected top boundary H(-1, j) saturates silently (it is not part of the overflow test), so a short query
// against a sequence of more than ~120 columns can finish the 8-bit tier with a score that is too high - e.g.
// query "AA", match 5, gaps -1/-1, a 150-nt record: -118 instead of -139 - and the reference reports that value
// (found by the differential fuzz campaign of round 2).  sat_score receives it; the search substitutes it for the
// exact score of every record the entry tier (or the 16-bit tier behind it) accepts.
// One thread per sequence, 4 columns per block exactly like the reference, H/E of the previous block in
// global scratch laid out [row][thread] (coalesced); a thread stops at its first overflowing block.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace ssa {

struct NwSatArgs {
    const uint16_t * stream;      // packed column stream
    const uint64_t * pair_col;    // first column of every pair
    const uint32_t * sorted_len;  // length by local sorted position
    const uint32_t * list;        // local sorted positions to emulate
    uint32_t count;
    const uint8_t * query;        // query codes
    uint32_t qlen;
    const int32_t * matrix;       // 32 x 32, (db << 5) + query
    int32_t gap_open, gap_extend;
    int32_t bits;                 // 8 or 16
    int16_t * scratch_h;          // [qlen][count]
    int16_t * scratch_e;
    int32_t * first_col;          // out, by local sorted position: 4 * overflow block, untouched if no overflow
    int32_t * sat_score;          // out (may be null), by local sorted position: the tier's own result if it does not overflow
};

__device__ __forceinline__ int sat_clamp( int v, int lo, int hi ) { return v < lo ? lo : ( v > hi ? hi : v ); }
__device__ __forceinline__ int sat_wrap( int v, int bits ) { return bits == 8 ? (int) (int8_t) v : (int) (int16_t) v; }

__global__ void __launch_bounds__( 128 ) nw_sat_kernel( const NwSatArgs a ) {
    __shared__ int32_t M[1024];
    for( int i = threadIdx.x; i < 1024; i += blockDim.x ) M[i] = sat_wrap( a.matrix[i], a.bits );
    __syncthreads();
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if( k >= a.count ) return;
    const uint32_t s = a.list[k];
    const uint32_t dlen = a.sorted_len[s];
    const uint16_t * st = a.stream + a.pair_col[s >> 1];
    const int shift = ( s & 1 ) ? 5 : 0;
    const int bits = a.bits;
    const int LO = bits == 8 ? -128 : -32768, HI = bits == 8 ? 127 : 32767;
    const int QR = sat_wrap( a.gap_open + a.gap_extend, bits ), RG = sat_wrap( a.gap_extend, bits );
    const int score_min = sat_wrap( LO - a.gap_open - a.gap_extend - 1, bits );  // search_simd_nw.c:372
    int16_t * H = a.scratch_h + k;
    int16_t * E = a.scratch_e + k;
    const size_t stride = a.count;
    int Hb[4], Fb[4], Sm[4] = { 0, 0, 0, 0 };
    Hb[0] = 0;
    for( int c = 1; c < 4; c++ ) Hb[c] = sat_wrap( a.gap_open + c * a.gap_extend, bits );           // :445-448
    for( int c = 0; c < 4; c++ ) Fb[c] = sat_wrap( a.gap_open + ( c + 1 ) * a.gap_extend, bits );   // :450-453
    const uint32_t nblocks = ( dlen + 3 ) / 4;
    int ovf = -1;
    for( uint32_t blk = 0; blk < nblocks && ovf < 0; blk++ ) {
        int res[4], h[4], f[4], hn[4];
        for( int c = 0; c < 4; c++ ) {
            const uint32_t col = 4 * blk + c;
            res[c] = col < dlen ? ( ( st[col] >> shift ) & 31 ) : 0;   // zero padding past the end, search_8_util.h:60-74
            h[c] = Hb[c];
            f[c] = sat_clamp( Fb[c] + QR, LO, HI );
        }
        int hmin = 0, hmax = 0, mge = QR;
        for( uint32_t i = 0; i < a.qlen; i++ ) {
            int h4, e;
            if( blk == 0 ) {  // masked initialisation of a restarting channel, :205-240
                h4 = sat_clamp( mge, LO, HI );
                e = sat_clamp( sat_clamp( mge, LO, HI ) + QR, LO, HI );
                mge = sat_clamp( mge + RG, LO, HI );
            }
            else {
                h4 = H[i * stride];
                e = E[i * stride];
            }
            const int qi = a.query[i];
#pragma unroll
            for( int c = 0; c < 4; c++ ) {
                int x = sat_clamp( h[c] + M[( res[c] << 5 ) + qi], LO, HI );
                x = max( x, f[c] );
                x = max( x, e );
                hmin = min( hmin, x );
                hmax = max( hmax, x );
                hn[c] = x;
                x = sat_clamp( x + QR, LO, HI );
                f[c] = max( sat_clamp( f[c] + RG, LO, HI ), x );
                e = max( sat_clamp( e + RG, LO, HI ), x );
            }
            H[i * stride] = (int16_t) hn[3];
            E[i * stride] = (int16_t) e;
            h[0] = h4;
            h[1] = hn[0];
            h[2] = hn[1];
            h[3] = hn[2];
        }
        Sm[0] = h[1];
        Sm[1] = h[2];
        Sm[2] = h[3];
        Sm[3] = hn[3];
        if( score_min > hmin || hmax == HI ) {  // :483-485
            ovf = (int) blk;
            break;
        }
        Fb[0] = sat_clamp( Fb[3] + RG, LO, HI );  // :494-502
        for( int c = 1; c < 4; c++ ) Fb[c] = sat_clamp( Fb[c - 1] + RG, LO, HI );
        Hb[0] = sat_clamp( Hb[3] + RG, LO, HI );
        for( int c = 1; c < 4; c++ ) Hb[c] = sat_clamp( Hb[c - 1] + RG, LO, HI );
    }
    if( ovf < 0 ) {
        const int score = Sm[( dlen + 3 ) % 4];  // :423-426
        if( !( score > LO && score < HI ) ) ovf = nblocks ? (int) nblocks - 1 : 0;
    }
    if( dlen == 0 ) ovf = 0;   // (no such records: empty ones are dropped when the database is packed)
    if( ovf >= 0 )
        a.first_col[s] = 4 * ovf;
    else if( a.sat_score )
        a.sat_score[s] = Sm[( dlen + 3 ) % 4];
}

}  // namespace ssa
