// Full-warp kernel, Needleman-Wunsch instantiations (sw32_kernel.cuh, sw32_launch.inc).
#define SW32_NW true
#define SW32_LAUNCH_NAME launch_sw32_nw
#define SW32_OCC_NAME sw32_blocks_per_sm_nw
#include "sw32_launch.inc"
