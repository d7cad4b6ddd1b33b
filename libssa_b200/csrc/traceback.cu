#include "traceback.h"
void traceback_hits( cudaStream_t, const uint8_t *, size_t, bool, std::vector<HitAlignment> &, ssa_b200_stats_t * ) {
    fatal( "libssa_b200: COMPUTE_ALIGNMENT is not implemented yet." );
}
void traceback_release() {}
