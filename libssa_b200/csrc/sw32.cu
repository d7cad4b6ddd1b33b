// Instantiations + launcher of the full-warp Smith-Waterman kernel (sw32_kernel.cuh).
#include "score_launch.h"

#include <cstdio>
#include <cstdlib>

namespace ssa {

template<typename K> static int launch( K kernel, int smem_bytes, const Sw32Args & args, cudaStream_t stream ) {
    const std::pair<int, int> lim = kernel_limits( kernel, smem_bytes );
    // persistent grid: a whole number of CTAs per SM, never more warps than batches
    const long warps_per_cta = SCORE_THREADS / 32;
    long want = ( (long) args.nbatches + warps_per_cta - 1 ) / warps_per_cta;
    int grid = lim.first * lim.second;
    if( want < grid ) grid = (int) ( want < 1 ? 1 : want );
    kernel<<<grid, SCORE_THREADS, smem_bytes, stream>>>( args );
    return grid;
}

template<int R> static int pick( bool multipass, const Sw32Args & a, cudaStream_t s ) {
    return multipass ? launch( sw32_kernel<R, true>, sw32_smem_bytes<R, true>(), a, s )
                     : launch( sw32_kernel<R, false>, sw32_smem_bytes<R, false>(), a, s );
}

template<int R> static int occ( bool multipass ) {
    return multipass ? kernel_limits( sw32_kernel<R, true>, sw32_smem_bytes<R, true>() ).second
                     : kernel_limits( sw32_kernel<R, false>, sw32_smem_bytes<R, false>() ).second;
}

int launch_sw32( int R, bool multipass, const Sw32Args & a, cudaStream_t s ) {
    switch( R ) {
    case 4: return pick<4>( multipass, a, s );
    case 8: return pick<8>( multipass, a, s );
    case 12: return pick<12>( multipass, a, s );
    case 16: return pick<16>( multipass, a, s );
    case 20: return pick<20>( multipass, a, s );
    case 24: return pick<24>( multipass, a, s );
    case 28: return pick<28>( multipass, a, s );
    case 32: return pick<32>( multipass, a, s );
    }
    fprintf( stderr, "libssa_b200: no full-warp kernel for R=%d\n", R );
    exit( 1 );
}

int sw32_blocks_per_sm( int R, bool multipass ) {
    switch( R ) {
    case 4: return occ<4>( multipass );
    case 8: return occ<8>( multipass );
    case 12: return occ<12>( multipass );
    case 16: return occ<16>( multipass );
    case 20: return occ<20>( multipass );
    case 24: return occ<24>( multipass );
    case 28: return occ<28>( multipass );
    case 32: return occ<32>( multipass );
    }
    return 1;
}

}  // namespace ssa
