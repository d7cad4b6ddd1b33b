// Instantiations + launcher of the two-queries-per-word kernels (score_kernel<PACKED, R, NW, 0, false, DUAL = true>).
#include "score_launch.h"

#include <cstdio>
#include <cstdlib>

namespace ssa {

template<int R> static constexpr int dual_smem( bool nw ) { return ( R * 1024 + ( nw ? 2 * R * 32 : 0 ) ) * 4; }

template<typename K> static int launch( K kernel, int smem_bytes, const ScoreArgs & args, cudaStream_t stream ) {
    const std::pair<int, int> lim = kernel_limits( kernel, smem_bytes );
    const long groups_per_cta = ( SCORE_THREADS / 32 ) * ( 32 / args.group );
    long want = ( (long) args.nbatches + groups_per_cta - 1 ) / groups_per_cta;  // never more groups than batches
    int grid = lim.first * lim.second;
    if( want < grid ) grid = (int) ( want < 1 ? 1 : want );
    kernel<<<grid, SCORE_THREADS, smem_bytes, stream>>>( args );
    return grid;
}

template<int R> static int pick( bool nw, const ScoreArgs & a, cudaStream_t s ) {
    return nw ? launch( score_kernel<true, R, true, 0, false, true>, dual_smem<R>( true ), a, s )
              : launch( score_kernel<true, R, false, 0, false, true>, dual_smem<R>( false ), a, s );
}

template<int R> static int occ( bool nw ) {
    return nw ? kernel_limits( score_kernel<true, R, true, 0, false, true>, dual_smem<R>( true ) ).second
              : kernel_limits( score_kernel<true, R, false, 0, false, true>, dual_smem<R>( false ) ).second;
}

int launch_score_dual( int R, bool nw, const ScoreArgs & a, cudaStream_t s ) {
    switch( R ) {
    case 4: return pick<4>( nw, a, s );
    case 8: return pick<8>( nw, a, s );
    case 12: return pick<12>( nw, a, s );
    case 16: return pick<16>( nw, a, s );
    case 20: return pick<20>( nw, a, s );
    case 24: return pick<24>( nw, a, s );
    case 28: return pick<28>( nw, a, s );
    case 32: return pick<32>( nw, a, s );
    }
    fprintf( stderr, "libssa_b200: no two-query kernel for R=%d\n", R );
    exit( 1 );
}

int dual_blocks_per_sm( int R, bool nw ) {
    switch( R ) {
    case 4: return occ<4>( nw );
    case 8: return occ<8>( nw );
    case 12: return occ<12>( nw );
    case 16: return occ<16>( nw );
    case 20: return occ<20>( nw );
    case 24: return occ<24>( nw );
    case 28: return occ<28>( nw );
    case 32: return occ<32>( nw );
    }
    return 1;
}

}  // namespace ssa
