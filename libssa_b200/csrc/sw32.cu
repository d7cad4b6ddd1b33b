// Full-warp kernel, Smith-Waterman instantiations (sw32_kernel.cuh, sw32_launch.inc) and the SW / NW dispatch.
#define SW32_NW false
#define SW32_LAUNCH_NAME launch_sw32_sw
#define SW32_OCC_NAME sw32_blocks_per_sm_sw
#include "sw32_launch.inc"

namespace ssa {
int launch_sw32_nw( int R, bool multipass, const Sw32Args & a, cudaStream_t s );
int sw32_blocks_per_sm_nw( int R, bool multipass );

int launch_sw32( int R, bool nw, bool multipass, const Sw32Args & a, cudaStream_t s ) {
    return nw ? launch_sw32_nw( R, multipass, a, s ) : launch_sw32_sw( R, multipass, a, s );
}
int sw32_blocks_per_sm( int R, bool nw, bool multipass ) { return nw ? sw32_blocks_per_sm_nw( R, multipass ) : sw32_blocks_per_sm_sw( R, multipass ); }
}  // namespace ssa
