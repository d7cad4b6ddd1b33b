// Host-side launchers of the scoring kernels (score_packed.cu / score_wide.cu).
#pragma once
#include "score_kernel.cuh"
#include "sw32_kernel.cuh"

namespace ssa {
// rows per lane the kernels are compiled for
constexpr int PACKED_R[] = { 4, 8, 12, 16, 20, 24, 28, 32 };
constexpr int WIDE_R[] = { 8, 16, 32 };
// launches score_kernel<PACKED=true,R,nw,multi,false>; returns the grid size used
int launch_score_packed( int R, bool nw, bool multi, const ScoreArgs & args, cudaStream_t stream );
// resident CTAs per SM of the packed kernel variant (for sizing the work batches)
int packed_blocks_per_sm( int R, bool nw, bool multi );
int launch_score_wide( int R, bool nw, bool multi, bool track, const ScoreArgs & args, cudaStream_t stream );
// two queries per packed word (score_kernel<..., DUAL = true>, one pass): args.dual_half selects sequence A or B of every pair
int launch_score_dual( int R, bool nw, const ScoreArgs & args, cudaStream_t stream );
int dual_blocks_per_sm( int R, bool nw );
// The 32-bit tier for queries of more than one pass (wide_pipe.cu): all passes of a sequence run as a pipeline in ONE launch.
constexpr int WIDE_PIPE_R = 16;   // rows per lane: 64 KB of profile per CTA, two CTAs per multiprocessor
struct WidePipeArgs {
    const uint16_t * stream;
    const Batch * batches;             // one sequence each (col0, ncols, half, out0 = output slot, bnd0 = first boundary column)
    uint32_t nbatches;                 // sizes the grid ...
    const uint32_t * nbatches_dev;     // ... if set, the real number is read from device memory
    uint32_t * counter;                // task queue head (zero at launch)
    const uint32_t * profile;          // npass profiles of WIDE_PIPE_R x 1024 words, [row][residue][lane]
    uint32_t npass;
    int32_t gap_open, gap_extend;
    int32_t cap_pass, cap_lane, cap_row;   // NW: where the query's last row lives
    int32_t * out;                     // score per output slot (SW: zero at launch, passes combine by atomicMax)
    uint4 * bnd0, * bnd1;              // boundary rows { H col A, H col B, F col A, F col B } per column pair, two buffers
    uint32_t * progress;               // [batch][pass] column pairs whose boundary row is written (zero at launch)
};
int launch_wide_pipe( bool nw, const WidePipeArgs & args, cudaStream_t stream );
// launches sw32_kernel<R, multipass, nw> (packed tier, full-warp arrays, chained batches); returns the grid size used
// G: lanes per array (32, or 16 - one pass only); dual: two queries per packed word (one pass only)
int launch_sw32( int R, bool nw, bool multipass, const Sw32Args & args, cudaStream_t stream, int G = 32, bool dual = false );
int sw32_blocks_per_sm( int R, bool nw, bool multipass, int G = 32, bool dual = false );
}  // namespace ssa

#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <utility>

namespace ssa {
// Multiprocessors of the current device and resident CTAs per multiprocessor of `kernel` with `smem_bytes` of dynamic shared
// memory; sets the kernel's shared-memory limit.  Cached per (device, kernel): the runtime queries cost more host time per
// launch than everything else a search enqueues.  Called from one host thread per GPU, hence the lock.
template<typename K> inline std::pair<int, int> kernel_limits( K kernel, int smem_bytes ) {
    static std::map<std::pair<int, const void *>, std::pair<int, int>> cache;
    static std::mutex lock;
    int dev = 0;
    cudaGetDevice( &dev );
    const std::pair<int, const void *> key( dev, (const void *) kernel );
    std::lock_guard<std::mutex> guard( lock );
    auto it = cache.find( key );
    if( it != cache.end() ) return it->second;
    int sms = 0, per_sm = 0;
    cudaError_t e = cudaFuncSetAttribute( kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes );
    if( e == cudaSuccess ) e = cudaDeviceGetAttribute( &sms, cudaDevAttrMultiProcessorCount, dev );
    if( e == cudaSuccess ) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor( &per_sm, kernel, SCORE_THREADS, smem_bytes );
    if( e != cudaSuccess ) {
        fprintf( stderr, "libssa_b200: cannot configure a scoring kernel (%d bytes of shared memory): %s\n", smem_bytes, cudaGetErrorString( e ) );
        exit( 1 );
    }
    if( per_sm < 1 ) per_sm = 1;
    return cache[key] = std::make_pair( sms, per_sm );
}
}  // namespace ssa
