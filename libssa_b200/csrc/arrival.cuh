// GPU emulation of the reference's one-thread heap-arrival order (see replay.cpp for the model and the
// reference citations).  One warp per chunk of the provider order; the lanes play the SIMD channels of the
// reference scheduler (32 for the 8-bit tier, 16 for the 16-bit tier): every iteration the channel with the
// smallest (finish block, channel index) key emits its record and is refilled with the chunk's next record -
// found with warp-wide REDUX reductions over a 64-bit key.  Records that overflow a tier are collected in emission order and run through
// the next tier, exactly like search_8_chunk / search_16_chunk (reference src/algo/8/search_8.c:94-124,
// src/algo/16/search_16.c:92-114).  The overflow positions come straight from the tracking kernel's output, so
// nothing but the final order has to cross PCIe.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace ssa {

constexpr int ARRIVAL_MAX_CHUNK = 2048;  // records per chunk the kernel handles (shared-memory lists)

struct ArrivalArgs {
    const uint32_t * lens;            // record lengths in provider order (0 = skipped)
    uint32_t n_all;                   // provider records
    uint32_t chunk;                   // records per chunk
    uint32_t nchunks;                 // chunks the reference processes (it stops at the first empty one)
    const uint32_t * chunk_base;      // arrival position of each chunk's first emitted record
    int32_t bit_width;                // 8, 16 or 64
    const int32_t * sorted_pos_of_id; // provider index -> global sorted position
    const int32_t * first_all;        // nullptr, or per rank r: [stride] first column >= 255 | [stride] first column >= 65535
    uint32_t stride, world;
    uint32_t * order;                 // out: arrival position -> provider index
    int32_t * pos_of_sorted;          // out: global sorted position -> arrival position
};

__global__ void __launch_bounds__( 32 ) arrival_kernel( const ArrivalArgs a ) {
    __shared__ uint32_t list_a[ARRIVAL_MAX_CHUNK];  // records entering the current tier (provider indices)
    __shared__ uint32_t blk_a[ARRIVAL_MAX_CHUNK];   // blocks each occupies its channel
    __shared__ uint32_t list_b[ARRIVAL_MAX_CHUNK];  // records that overflowed the current tier, in emission order
    constexpr unsigned FULL = 0xffffffffu;
    const uint32_t c = blockIdx.x;
    if( c >= a.nchunks ) return;
    const int lane = threadIdx.x;
    const uint32_t id0 = c * a.chunk;
    const uint32_t id1 = min( a.n_all, id0 + a.chunk );
    uint32_t out = a.chunk_base[c];

    auto overflow_block = [&]( uint32_t id, int tier ) -> int32_t {  // -1 = no overflow in this tier (8 or 16)
        if( !a.first_all ) return -1;
        const uint32_t g = (uint32_t) a.sorted_pos_of_id[id];
        const uint32_t r = g % a.world, sl = g / a.world;
        const int32_t col = a.first_all[(size_t) r * 2 * a.stride + ( tier == 8 ? 0 : a.stride ) + sl];
        return col == 0x7fffffff ? -1 : col / 4;
    };
    auto emit_final = [&]( uint32_t id, uint32_t where ) {
        a.order[where] = id;
        a.pos_of_sorted[a.sorted_pos_of_id[id]] = (int32_t) where;
    };

    // non-empty records of the chunk, provider order
    uint32_t m = 0;
    for( uint32_t base = id0; base < id1; base += 32 ) {
        const uint32_t id = base + lane;
        const bool keep = id < id1 && a.lens[id] > 0;
        const unsigned bal = __ballot_sync( FULL, keep );
        if( keep ) list_a[m + __popc( bal & ( ( 1u << lane ) - 1u ) )] = id;
        m += __popc( bal );
    }
    __syncwarp();
    if( a.bit_width == 64 ) {  // search_64_chunk: provider order
        for( uint32_t i = lane; i < m; i += 32 ) emit_final( list_a[i], out + i );
        return;
    }

    for( int tier = ( a.bit_width == 8 ? 8 : 16 ); tier <= 16; tier += 8 ) {
        const int channels = tier == 8 ? 32 : 16;
        for( uint32_t i = lane; i < m; i += 32 ) {
            const uint32_t id = list_a[i];
            uint32_t b = ( a.lens[id] + 3 ) / 4;
            const int32_t ob = overflow_block( id, tier );
            if( ob >= 0 ) b = min( b, (uint32_t) ob + 1 );
            blk_a[i] = b;
        }
        __syncwarp();
        // channel state: the block after which the channel's record finishes (64 bit: a chunk of 2048 records of up to 2^22
        // blocks each can accumulate more than 2^32 blocks in one channel); the reference's scan finds the channel with
        // the smallest (finish block, channel index) first (replay.cpp: simd_emission uses the same 64-bit key)
        uint32_t rec = lane;
        uint64_t fin = 0;
        bool active = lane < channels && (uint32_t) lane < m;
        if( active ) fin = blk_a[lane];
        uint32_t next = min( (uint32_t) channels, m ), nb = 0;
        for( uint32_t e = 0; e < m; e++ ) {
            const uint32_t hi = (uint32_t) ( fin >> 32 ), lo = (uint32_t) fin;
            const uint32_t hmin = __reduce_min_sync( FULL, active ? hi : 0xffffffffu );
            const uint32_t lmin = __reduce_min_sync( FULL, ( active && hi == hmin ) ? lo : 0xffffffffu );
            const unsigned tied = __ballot_sync( FULL, active && hi == hmin && lo == lmin );
            bool to_b = false;
            if( lane == __ffs( tied ) - 1 ) {  // exactly one lane: the channel the reference's scan finds finished first
                const uint32_t id = list_a[rec];
                to_b = overflow_block( id, tier ) >= 0;
                if( to_b )
                    list_b[nb] = id;
                else
                    emit_final( id, out );
                if( next < m ) {
                    rec = next;
                    fin += blk_a[next];
                }
                else {
                    active = false;
                }
            }
            if( __ballot_sync( FULL, to_b ) )
                nb++;
            else
                out++;
            if( next < m ) next++;
        }
        __syncwarp();
        // the overflowed records enter the next tier in the order they were collected
        for( uint32_t i = lane; i < nb; i += 32 ) list_a[i] = list_b[i];
        m = nb;
        __syncwarp();
    }
    for( uint32_t i = lane; i < m; i += 32 ) emit_final( list_a[i], out + i );  // 64-bit tier: list order
}

}  // namespace ssa
