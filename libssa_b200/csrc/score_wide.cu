// Instantiations + launcher of the wide (int32) scoring kernels: the exact tier for sequences whose
// scores may not fit 16 bits, and the overflow-position tracker used to emulate the reference's
// 8/16-bit tier order.
#include "score_launch.h"

#include <cstdio>
#include <cstdlib>

namespace ssa {

template<typename K> static int launch( K kernel, int smem_bytes, const ScoreArgs & args, cudaStream_t stream ) {
    const std::pair<int, int> lim = kernel_limits( kernel, smem_bytes );
    const int sms = lim.first, per_sm = lim.second;
    // persistent grid: a whole number of CTAs per SM; every warp pulls batches from the work queue
    const long groups_per_cta = ( SCORE_THREADS / 32 ) * ( 32 / args.group );
    long want = ( (long) args.nbatches + groups_per_cta - 1 ) / groups_per_cta;
    int grid = sms * per_sm;
    if( want < grid ) grid = (int) ( want < 1 ? 1 : want );
    kernel<<<grid, SCORE_THREADS, smem_bytes, stream>>>( args );
    return grid;
}

template<int R> static int pick( bool nw, bool multi, bool track, const ScoreArgs & a, cudaStream_t s ) {
    const int smem = profile_words<false, R>() * 4;
    if( nw ) return multi ? launch( score_kernel<false, R, true, 1, false>, smem, a, s )
                          : launch( score_kernel<false, R, true, 0, false>, smem, a, s );
    if( track ) return multi ? launch( score_kernel<false, R, false, 1, true>, smem, a, s )
                             : launch( score_kernel<false, R, false, 0, true>, smem, a, s );
    return multi ? launch( score_kernel<false, R, false, 1, false>, smem, a, s )
                 : launch( score_kernel<false, R, false, 0, false>, smem, a, s );
}

int launch_score_wide( int R, bool nw, bool multi, bool track, const ScoreArgs & a, cudaStream_t s ) {
    switch( R ) {
    case 8: return pick<8>( nw, multi, track, a, s );
    case 16: return pick<16>( nw, multi, track, a, s );
    case 32: return pick<32>( nw, multi, track, a, s );
    }
    fprintf( stderr, "libssa_b200: no wide kernel for R=%d\n", R );
    exit( 1 );
}

}  // namespace ssa
