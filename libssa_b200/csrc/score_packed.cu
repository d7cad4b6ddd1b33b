// Instantiations + launcher of the packed (2 x int16 per word) scoring kernels.
#include "score_launch.h"

#include <cstdio>
#include <cstdlib>

namespace ssa {

template<typename K> static int launch( K kernel, int smem_bytes, const ScoreArgs & args, cudaStream_t stream ) {
    static thread_local int cached_dev = -1;
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice( &dev );
    (void) cached_dev;
    cudaFuncSetAttribute( kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes );
    cudaDeviceGetAttribute( &sms, cudaDevAttrMultiProcessorCount, dev );
    cudaOccupancyMaxActiveBlocksPerMultiprocessor( &per_sm, kernel, SCORE_THREADS, smem_bytes );
    if( per_sm < 1 ) per_sm = 1;
    // persistent grid: a whole number of CTAs per SM; every warp pulls batches from the work queue
    long want = ( (long) args.nbatches + ( SCORE_THREADS / 32 ) - 1 ) / ( SCORE_THREADS / 32 );
    int grid = sms * per_sm;
    if( want < grid ) grid = (int) ( want < 1 ? 1 : want );
    kernel<<<grid, SCORE_THREADS, smem_bytes, stream>>>( args );
    return grid;
}

template<int R> static int pick( bool nw, bool multi, const ScoreArgs & a, cudaStream_t s ) {
    const int smem = profile_words<true, R>() * 4;
    if( nw ) return multi ? launch( score_kernel<true, R, true, true, false>, smem, a, s )
                          : launch( score_kernel<true, R, true, false, false>, smem, a, s );
    return multi ? launch( score_kernel<true, R, false, true, false>, smem, a, s )
                 : launch( score_kernel<true, R, false, false, false>, smem, a, s );
}

int launch_score_packed( int R, bool nw, bool multi, const ScoreArgs & a, cudaStream_t s ) {
    switch( R ) {
    case 4: return pick<4>( nw, multi, a, s );
    case 8: return pick<8>( nw, multi, a, s );
    case 12: return pick<12>( nw, multi, a, s );
    case 16: return pick<16>( nw, multi, a, s );
    case 24: return pick<24>( nw, multi, a, s );
    case 32: return pick<32>( nw, multi, a, s );
    }
    fprintf( stderr, "libssa_b200: no packed kernel for R=%d\n", R );
    exit( 1 );
}

}  // namespace ssa
