// Chained kernel instantiations (sw32_kernel.cuh, sw32_launch.inc): Needleman-Wunsch, arrays of 16 lanes, one query per word.
#define SW32_NW true
#define SW32_G 16
#define SW32_DUAL false
#define SW32_LAUNCH_NAME launch_sw32_g16_nw
#define SW32_OCC_NAME sw32_occ_g16_nw
#include "sw32_launch.inc"
