// Host-side launchers of the scoring kernels (score_packed.cu / score_wide.cu).
#pragma once
#include "score_kernel.cuh"

namespace ssa {
// rows per lane the kernels are compiled for
constexpr int PACKED_R[] = { 4, 8, 12, 16, 20, 24, 28, 32 };
constexpr int WIDE_R[] = { 8, 16, 32 };
// launches score_kernel<PACKED=true,R,nw,multi,false>; returns the grid size used
int launch_score_packed( int R, bool nw, bool multi, const ScoreArgs & args, cudaStream_t stream );
// resident CTAs per SM of the packed kernel variant (for sizing the work batches)
int packed_blocks_per_sm( int R, bool nw, bool multi );
int launch_score_wide( int R, bool nw, bool multi, bool track, const ScoreArgs & args, cudaStream_t stream );
}  // namespace ssa
