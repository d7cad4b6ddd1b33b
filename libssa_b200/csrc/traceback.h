// Alignment (region + CIGAR) of the reported hits on the GPU; see traceback.cu.
#pragma once
#include "common.h"
#include <cuda_runtime.h>

void traceback_hits( cudaStream_t stream, const uint8_t * query, size_t qlen, bool nw, std::vector<HitAlignment> & hits,
        ssa_b200_stats_t * stats, bool scored_sequence );
struct TbQuery {
    const uint8_t * seq;
    size_t len;
};
// hits of several queries in one go: hit k belongs to queries[query_of_hit[k]]
void traceback_hits_multi( cudaStream_t stream, const std::vector<TbQuery> & queries, const std::vector<uint32_t> & query_of_hit, bool nw,
        std::vector<HitAlignment> & hits, ssa_b200_stats_t * stats, const std::vector<char> & scored_sequence );
void traceback_release();
