// Chained kernel instantiations (sw32_kernel.cuh, sw32_launch.inc): Needleman-Wunsch, arrays of 32 lanes, two queries per word.
#define SW32_NW true
#define SW32_G 32
#define SW32_DUAL true
#define SW32_LAUNCH_NAME launch_sw32_dual32_nw
#define SW32_OCC_NAME sw32_occ_dual32_nw
#include "sw32_launch.inc"
