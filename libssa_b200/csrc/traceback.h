// Alignment (region + CIGAR) of the reported hits on the GPU; see traceback.cu.
#pragma once
#include "common.h"
#include <cuda_runtime.h>

void traceback_hits( cudaStream_t stream, const uint8_t * query, size_t qlen, bool nw, std::vector<HitAlignment> & hits,
        ssa_b200_stats_t * stats, bool scored_sequence );
void traceback_release();
