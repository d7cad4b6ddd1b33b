// Chained kernel, Smith-Waterman / full-warp instantiations (sw32_kernel.cuh, sw32_launch.inc) and the dispatch over all sets.
#define SW32_NW false
#define SW32_LAUNCH_NAME launch_sw32_sw
#define SW32_OCC_NAME sw32_blocks_per_sm_sw
#include "sw32_launch.inc"

namespace ssa {
#define SW32_DECL( L, O ) int L( int R, bool multipass, const Sw32Args & a, cudaStream_t s ); int O( int R, bool multipass );
SW32_DECL( launch_sw32_nw, sw32_blocks_per_sm_nw )
SW32_DECL( launch_sw32_g16_sw, sw32_occ_g16_sw )
SW32_DECL( launch_sw32_g16_nw, sw32_occ_g16_nw )
SW32_DECL( launch_sw32_dual32_sw, sw32_occ_dual32_sw )
SW32_DECL( launch_sw32_dual32_nw, sw32_occ_dual32_nw )
SW32_DECL( launch_sw32_dual16_sw, sw32_occ_dual16_sw )
SW32_DECL( launch_sw32_dual16_nw, sw32_occ_dual16_nw )

int launch_sw32( int R, bool nw, bool multipass, const Sw32Args & a, cudaStream_t s, int G, bool dual ) {
    if( dual ) {
        if( G == 16 ) return nw ? launch_sw32_dual16_nw( R, false, a, s ) : launch_sw32_dual16_sw( R, false, a, s );
        return nw ? launch_sw32_dual32_nw( R, false, a, s ) : launch_sw32_dual32_sw( R, false, a, s );
    }
    if( G == 16 ) return nw ? launch_sw32_g16_nw( R, false, a, s ) : launch_sw32_g16_sw( R, false, a, s );
    return nw ? launch_sw32_nw( R, multipass, a, s ) : launch_sw32_sw( R, multipass, a, s );
}

int sw32_blocks_per_sm( int R, bool nw, bool multipass, int G, bool dual ) {
    if( dual ) {
        if( G == 16 ) return nw ? sw32_occ_dual16_nw( R, false ) : sw32_occ_dual16_sw( R, false );
        return nw ? sw32_occ_dual32_nw( R, false ) : sw32_occ_dual32_sw( R, false );
    }
    if( G == 16 ) return nw ? sw32_occ_g16_nw( R, false ) : sw32_occ_g16_sw( R, false );
    return nw ? sw32_blocks_per_sm_nw( R, multipass ) : sw32_blocks_per_sm_sw( R, multipass );
}
}  // namespace ssa
