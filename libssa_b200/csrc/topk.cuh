// Top-k hit selection on the GPU: radix select of the k-th largest score over the (all-gathered) score vector (three
// launches, 11 + 11 + 10 bits), tie count at that score, and compaction of the hits (one launch).
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
    uint32_t ge8, ge16;    // records scoring >= 255 / >= 65535 (the reference's overflow counters), final after pass 0
    uint32_t ticket;       // blocks of the running pass that have delivered their histogram
    uint32_t acc8, acc16;  // accumulators of ge8 / ge16
    uint32_t pad;
    uint32_t hist[2048];   // zero between passes
};
constexpr size_t SELECT_HEADER_BYTES = 48;   // what the host reads back

struct SelectArgs {
    const int32_t * scores;   // `world` blocks of `stride` scores (rank-major gathered layout)
    const int32_t * pos;      // global sorted position -> arrival position, < 0 = record the reference never reaches
    uint32_t stride, world, nglobal;
    SelectState * state;
};

__device__ __forceinline__ uint32_t score_key( int32_t s ) { return (uint32_t) s ^ 0x80000000u; }  // unsigned order = signed order

// One pass of the radix select in ONE launch: every block histograms its share of the keys that match the prefix decided
// so far; the block that delivers last (ticket) walks the bins from the top until it finds the bin of the k-th element,
// updates the state and leaves the histogram and the ticket zeroed for the next pass.  Pass 0 also initialises the state
// for `k` (no separate init launch; *state must be zero before the very first launch) and counts the records at or above
// the reference's 8- and 16-bit overflow thresholds.
__global__ void __launch_bounds__( 256 ) select_pass_kernel( const SelectArgs a, int pass, int shift, uint32_t binmask, uint32_t k, int last ) {
    __shared__ uint32_t h[2048];
    __shared__ uint32_t chunk_sum[32];
    __shared__ uint32_t ge[2];
    __shared__ bool is_last;
    SelectState * st = a.state;
    for( int i = threadIdx.x; i < 2048; i += blockDim.x ) h[i] = 0;
    if( threadIdx.x < 2 ) ge[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t prefix = pass ? st->prefix : 0u, mask = pass ? st->mask : 0u;
    const uint32_t total = a.stride * a.world;
    for( uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x ) {
        const uint32_t r = i / a.stride, sl = i % a.stride;
        const uint64_t g = (uint64_t) sl * a.world + r;
        if( g >= a.nglobal ) continue;
        const int32_t s = a.scores[i];
        if( pass == 0 && s >= 255 ) {
            atomicAdd( &ge[0], 1u );
            if( s >= 65535 ) atomicAdd( &ge[1], 1u );
        }
        if( a.pos[g] < 0 ) continue;
        const uint32_t key = score_key( s );
        if( ( key & mask ) == prefix ) atomicAdd( &h[( key >> shift ) & binmask], 1u );
    }
    __syncthreads();
    for( int i = threadIdx.x; i < 2048; i += blockDim.x )
        if( h[i] ) atomicAdd( &st->hist[i], h[i] );
    if( pass == 0 && threadIdx.x < 2 && ge[threadIdx.x] ) atomicAdd( threadIdx.x ? &st->acc16 : &st->acc8, ge[threadIdx.x] );
    __threadfence();
    __syncthreads();
    if( threadIdx.x == 0 ) is_last = atomicAdd( &st->ticket, 1u ) == gridDim.x - 1;
    __syncthreads();
    if( !is_last ) return;
    __threadfence();
    // ---- the last block: all histograms are in ----
    const uint32_t nbins = binmask + 1;
    for( uint32_t i = threadIdx.x; i < 2048; i += blockDim.x ) h[i] = i < nbins ? __ldcg( &st->hist[i] ) : 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;   // 32 chunks of 64 bins, four per warp
    for( int c = warp; c < 32; c += 8 ) {
        uint32_t s = h[c * 64 + lane] + h[c * 64 + 32 + lane];
        for( int o = 16; o > 0; o >>= 1 ) s += __shfl_down_sync( 0xffffffffu, s, o );
        if( lane == 0 ) chunk_sum[c] = s;
    }
    __syncthreads();
    if( threadIdx.x == 0 ) {
        uint32_t kk = pass ? st->k_remaining : k, above = 0;
        int c = 31;
        while( c > 0 && above + chunk_sum[c] < kk ) above += chunk_sum[c--];
        int b = c * 64 + 63;
        while( b > c * 64 && above + h[b] < kk ) above += h[b--];
        st->count_gt = ( pass ? st->count_gt : 0u ) + above;
        st->k_remaining = kk - above;
        st->prefix = prefix | ( (uint32_t) b << shift );
        st->mask = mask | ( binmask << shift );
        st->count_eq = last ? h[b] : 0u;
        if( pass == 0 ) {
            st->n_out = 0;
            st->ge8 = __ldcg( &st->acc8 );
            st->ge16 = __ldcg( &st->acc16 );
            st->acc8 = 0;
            st->acc16 = 0;
        }
        st->ticket = 0;
    }
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
