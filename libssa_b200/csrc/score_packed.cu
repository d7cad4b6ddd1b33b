// Instantiations + launcher of the packed (2 x int16 per word) scoring kernels.
#include "score_launch.h"

#include <cstdio>
#include <cstdlib>

namespace ssa {

template<typename K> static int launch( K kernel, int smem_bytes, const ScoreArgs & args, cudaStream_t stream ) {
    const std::pair<int, int> lim = kernel_limits( kernel, smem_bytes );
    const int sms = lim.first, per_sm = lim.second;
    // persistent grid: a whole number of CTAs per SM; every warp pulls batches from the work queue
    const long groups_per_cta = ( SCORE_THREADS / 32 ) * ( 32 / args.group );
    long want = ( (long) args.nbatches + groups_per_cta - 1 ) / groups_per_cta;  // never more groups than batches
    int grid = sms * per_sm;
    if( want < grid ) grid = (int) ( want < 1 ? 1 : want );
    kernel<<<grid, SCORE_THREADS, smem_bytes, stream>>>( args );
    return grid;
}

template<int R> static int pick( bool nw, bool multi, const ScoreArgs & a, cudaStream_t s ) {
    const int smem = score_smem_bytes<true, R>( nw );
    if( multi && a.group != 32 )  // remainder pass of a long query on narrow groups
        return nw ? launch( score_kernel<true, R, true, 2, false>, smem, a, s ) : launch( score_kernel<true, R, false, 2, false>, smem, a, s );
    if( nw ) return multi ? launch( score_kernel<true, R, true, 1, false>, smem, a, s )
                          : launch( score_kernel<true, R, true, 0, false>, smem, a, s );
    return multi ? launch( score_kernel<true, R, false, 1, false>, smem, a, s )
                 : launch( score_kernel<true, R, false, 0, false>, smem, a, s );
}

template<typename K> static int occupancy( K kernel, int smem_bytes ) { return kernel_limits( kernel, smem_bytes ).second; }

template<int R> static int occ_pick( bool nw, bool multi ) {
    const int smem = score_smem_bytes<true, R>( nw );
    if( nw ) return multi ? occupancy( score_kernel<true, R, true, 1, false>, smem ) : occupancy( score_kernel<true, R, true, 0, false>, smem );
    return multi ? occupancy( score_kernel<true, R, false, 1, false>, smem ) : occupancy( score_kernel<true, R, false, 0, false>, smem );
}

int packed_blocks_per_sm( int R, bool nw, bool multi ) {
    switch( R ) {
    case 4: return occ_pick<4>( nw, multi );
    case 8: return occ_pick<8>( nw, multi );
    case 12: return occ_pick<12>( nw, multi );
    case 16: return occ_pick<16>( nw, multi );
    case 20: return occ_pick<20>( nw, multi );
    case 24: return occ_pick<24>( nw, multi );
    case 28: return occ_pick<28>( nw, multi );
    case 32: return occ_pick<32>( nw, multi );
    }
    return 1;
}

int launch_score_packed( int R, bool nw, bool multi, const ScoreArgs & a, cudaStream_t s ) {
    switch( R ) {
    case 4: return pick<4>( nw, multi, a, s );
    case 8: return pick<8>( nw, multi, a, s );
    case 12: return pick<12>( nw, multi, a, s );
    case 16: return pick<16>( nw, multi, a, s );
    case 20: return pick<20>( nw, multi, a, s );
    case 24: return pick<24>( nw, multi, a, s );
    case 28: return pick<28>( nw, multi, a, s );
    case 32: return pick<32>( nw, multi, a, s );
    }
    fprintf( stderr, "libssa_b200: no packed kernel for R=%d\n", R );
    exit( 1 );
}

}  // namespace ssa
