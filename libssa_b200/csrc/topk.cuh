// Top-k hit selection on the GPU: radix select of the k-th largest score over the (all-gathered) score vector,
// tie count at that score, and compaction of the hits.
//
// Why this is exact: the reference keeps hits in a bounded min-heap that drops a newcomer equal to its minimum
// (reference src/util/minheap.c:75-91), so WHICH of several records tied at the cut-off score T survive depends on
// the arrival order - but only then.  If at most k records score >= T, no record with score >= T can ever be
// rejected or evicted (that would need k+1 records >= T), so the reported set is exactly {score >= T} whatever
// the order, and the final sort (score desc, ID desc: src/algo/manager.c:68-77) fixes the list.  Only when the
// tie at T straddles the cut-off does the search fall back to the full arrival-order replay (replay.cpp).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace ssa {

struct SelectState {
    uint32_t prefix;       // bits of the k-th largest key decided so far
    uint32_t mask;         // which bits those are
    uint32_t k_remaining;  // rank of the k-th element among the keys matching the prefix
    uint32_t count_gt;     // keys strictly greater than every key matching the prefix
    uint32_t count_eq;     // after the last pass: keys equal to the k-th largest key
    uint32_t n_out;        // compaction counter
    uint32_t pad[2];
    uint32_t hist[2048];
};

struct SelectArgs {
    const int32_t * scores;   // `world` blocks of `stride` scores (rank-major gathered layout)
    const int32_t * pos;      // global sorted position -> arrival position, < 0 = record the reference never reaches
    uint32_t stride, world, nglobal;
    SelectState * state;
};

__device__ __forceinline__ uint32_t score_key( int32_t s ) { return (uint32_t) s ^ 0x80000000u; }  // unsigned order = signed order

__global__ void select_init_kernel( SelectState * st, uint32_t k ) {
    for( int i = threadIdx.x; i < 2048; i += blockDim.x ) st->hist[i] = 0;
    if( threadIdx.x == 0 ) {
        st->prefix = 0;
        st->mask = 0;
        st->k_remaining = k;
        st->count_gt = 0;
        st->count_eq = 0;
        st->n_out = 0;
    }
}

__global__ void __launch_bounds__( 256 ) select_hist_kernel( const SelectArgs a, int shift, uint32_t binmask ) {
    __shared__ uint32_t h[2048];
    for( int i = threadIdx.x; i < 2048; i += blockDim.x ) h[i] = 0;
    __syncthreads();
    const uint32_t prefix = a.state->prefix, mask = a.state->mask;
    const uint32_t total = a.stride * a.world;
    for( uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x ) {
        const uint32_t r = i / a.stride, sl = i % a.stride;
        const uint64_t g = (uint64_t) sl * a.world + r;
        if( g >= a.nglobal || a.pos[g] < 0 ) continue;
        const uint32_t key = score_key( a.scores[i] );
        if( ( key & mask ) == prefix ) atomicAdd( &h[( key >> shift ) & binmask], 1u );
    }
    __syncthreads();
    for( int i = threadIdx.x; i < 2048; i += blockDim.x )
        if( h[i] ) atomicAdd( &a.state->hist[i], h[i] );
}

// one block: walks the bins from the top until the k-th element's bin is found
__global__ void __launch_bounds__( 1024 ) select_pick_kernel( SelectState * st, int shift, uint32_t binmask, int last ) {
    __shared__ uint32_t h[2048];
    __shared__ uint32_t chunk_sum[32];
    const uint32_t nbins = binmask + 1;
    for( uint32_t i = threadIdx.x; i < 2048; i += blockDim.x ) h[i] = i < nbins ? st->hist[i] : 0;
    __syncthreads();
    // 32 chunks of 64 bins, summed by one warp each
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t s = h[warp * 64 + lane] + h[warp * 64 + 32 + lane];
    for( int o = 16; o > 0; o >>= 1 ) s += __shfl_down_sync( 0xffffffffu, s, o );
    if( lane == 0 ) chunk_sum[warp] = s;
    __syncthreads();
    if( threadIdx.x == 0 ) {
        uint32_t k = st->k_remaining, above = 0;
        int c = 31;
        while( c > 0 && above + chunk_sum[c] < k ) above += chunk_sum[c--];
        int b = c * 64 + 63;
        while( b > c * 64 && above + h[b] < k ) above += h[b--];
        st->count_gt += above;
        st->k_remaining = k - above;
        st->prefix |= (uint32_t) b << shift;
        st->mask |= binmask << shift;
        if( last ) st->count_eq = h[b];
    }
    __syncthreads();
    for( uint32_t i = threadIdx.x; i < 2048; i += blockDim.x ) st->hist[i] = 0;
}

// (score, provider index) of every record whose score is >= the selected k-th largest score
__global__ void __launch_bounds__( 256 ) select_compact_kernel( const SelectArgs a, const uint32_t * sorted_all, int32_t * out_scores,
        uint32_t * out_ids, uint32_t cap ) {
    const uint32_t tkey = a.state->prefix;
    const uint32_t total = a.stride * a.world;
    for( uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x ) {
        const uint32_t r = i / a.stride, sl = i % a.stride;
        const uint64_t g = (uint64_t) sl * a.world + r;
        if( g >= a.nglobal || a.pos[g] < 0 ) continue;
        const int32_t s = a.scores[i];
        if( score_key( s ) >= tkey ) {
            const uint32_t o = atomicAdd( &a.state->n_out, 1u );
            if( o < cap ) {
                out_scores[o] = s;
                out_ids[o] = sorted_all[g];
            }
        }
    }
}

}  // namespace ssa
