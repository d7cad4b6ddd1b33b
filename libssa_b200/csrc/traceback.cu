// Alignment of the reported hits on the GPU: local region, direction matrix, backward walk, CIGAR.
// Replaces reference src/algo/aligner.c:150-181 (thread fan-out over hits), src/algo/align.c:39-182 and
// src/algo/cigar.c:48-346.  The reference's exact semantics - including the ones that are not textbook
// Gotoh - are restated:
//   * SW region end = first strict maximum in (database outer, query inner) order of a forward pass whose
//     E/F are advanced BEFORE H and whose matrix is read as M[query][db] (align.c:56-83);
//   * SW region begin = first cell, in reverse scan order, of a reverse pass with -1 sentinels whose score
//     reaches the forward score; its row state is only re-initialised for indices 0..b_end although it is
//     indexed by the query position, so entries b_end+1..a_end still hold forward-pass values (align.c:86-127);
//   * direction bits per cell of a column-major Gotoh pass with strict comparisons (cigar.c:48-198);
//   * walk: left/ext-left before up/ext-up before diagonal, looking only at the CURRENT cell (cigar.c:299-335).
//
// Engine: the DP matrix of a hit is cut into tiles of 256 rows x 256 columns.  ONE WARP sweeps a tile as a systolic
// array - lane l owns 8 consecutive rows whose left-neighbour state (H, E of the previous column) lives in
// registers, columns stream through the lanes one per step, the H/F of a lane's last row reach the next lane by
// __shfl_up_sync - exactly the decomposition of the scoring kernels (score_kernel.cuh), in exact int32 and with the
// reference's cell functions.  Tiles on one anti-diagonal of the tile grid are independent: one launch per tile
// anti-diagonal processes them for all hits at once (a 10 kbp x 10 kbp pair has up to 40 tiles in flight per hit);
// tile boundaries (bottom row H/F, right column H/E) pass through global memory.  All three passes share that
// engine and differ only in the cell function and the boundary values.  Substitution scores come from a per-tile
// profile in shared memory laid out [row][residue][lane] (bank-conflict free whatever residues the lanes hold).
// The direction bytes are stored in 128-byte blocks of 8 rows x 16 columns, so that a lane writes its 8 rows of a
// column with one 64-bit store and the backward walk - a serial chain of dependent loads - stays inside one cache
// line for many steps.  Arithmetic is int32, exact for the supported lengths (< 2^24 symbols: |value| < 2^31).
#include "traceback.h"

#include <algorithm>
#include <climits>
#include <cstring>

#define CUDA_OK( call )                                                                                     \
    do {                                                                                                    \
        cudaError_t e_ = ( call );                                                                          \
        if( e_ != cudaSuccess ) fatal( "CUDA error %s at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString( e_ ) ); \
    } while( 0 )

namespace {

constexpr int TB_R = 8;                 // rows per lane
constexpr int TILE_R = 32 * TB_R;       // rows per tile
constexpr int TILE_C = 128;             // columns per tile

enum Pass { PASS_FORWARD = 0, PASS_REVERSE = 1, PASS_DIRECTIONS = 2 };

struct HitTask {              // one reported hit
    const uint8_t * a;        // query codes (shared by all hits)
    const uint8_t * b;        // database sequence codes
    int32_t alen, blen;
    // tile boundaries, set 0 for the forward and reverse pass, set 1 for the direction pass (which runs beside the reverse pass)
    int32_t * rowH[2], * rowF[2];   // [tile row][ldc]: H / F of the last row of every tile row, per column
    int32_t * colH[2], * colE[2];   // [tile column][ldr]: H / E of the last column of every tile column, per row
    int32_t * HHf, * EEf;     // forward pass: H / E of the last column, per row (inherited by the reverse pass)
    int64_t * tile_key;       // per tile: partial results of the forward (best cell) / reverse (first cell) pass
    int32_t * tile_val;
    int32_t ldr, ldc, ntc_cap;
    uint8_t * dir;            // direction bytes of rows 0..a_end x columns 0..b_end, blocked 8 rows x 16 columns
    int32_t * result;         // [0] score [1] a_end [2] b_end [3] a_begin [4] b_begin [5] found [6] nruns
    uint32_t * runs;          // walk output: (count << 8) | op, in walk order
    int32_t runs_cap;
};

struct TbParams {
    int32_t gapO, gapE;
    int32_t nw;
};

__device__ __forceinline__ int32_t max2( int32_t a, int32_t b ) { return a > b ? a : b; }

__device__ __forceinline__ size_t dir_offset( int nrb, int r, int c ) {
    return ( (size_t) ( c >> 4 ) * nrb + ( r >> 3 ) ) * 128 + ( ( c & 15 ) << 3 ) + ( r & 7 );
}

// boundary values of row -1 / column -1 of the three passes
template<int PASS> struct Bound {
    int32_t gapO, gapE, a_end, b_end;
    bool nw;
    const int32_t * HHf, * EEf;
    __device__ __forceinline__ int32_t top_h( int c ) const {   // also the diagonal neighbour of (0, c+1)
        if( PASS == PASS_FORWARD ) return 0;
        if( PASS == PASS_REVERSE ) return -1;
        return nw ? gapO + ( c + 1 ) * gapE : 0;
    }
    __device__ __forceinline__ int32_t top_f( int c ) const {
        if( PASS == PASS_FORWARD ) return gapO;
        if( PASS == PASS_REVERSE ) return -1;
        return nw ? 2 * gapO + ( c + 2 ) * gapE : 0;
    }
    __device__ __forceinline__ int32_t corner() const { return 0; }
    __device__ __forceinline__ int32_t left_h( int r ) const {
        if( PASS == PASS_FORWARD ) return 0;
        if( PASS == PASS_REVERSE ) {
            // query positions above b_end were not reset to -1 (align.c:86-89): they keep the forward pass' last column
            const int j = a_end - r;
            return j > b_end ? HHf[j] : -1;
        }
        return nw ? gapO + ( r + 1 ) * gapE : 0;
    }
    __device__ __forceinline__ int32_t left_e( int r ) const {
        if( PASS == PASS_FORWARD ) return gapO;
        if( PASS == PASS_REVERSE ) {
            const int j = a_end - r;
            return j > b_end ? EEf[j] : -1;
        }
        return nw ? 2 * gapO + ( r + 2 ) * gapE : 0;
    }
};

// One tile: see the header comment.
template<int PASS>
__device__ __forceinline__ void tile_sweep( int32_t * prof, int32_t * s_toph, int32_t * s_topf, uint8_t * s_code, const HitTask * tasks, const int32_t * matrix, const TbParams prm, int diag, int maxd ) {
    const int lane = threadIdx.x;
    const HitTask t = tasks[blockIdx.x / maxd];
    constexpr int SET = PASS == PASS_DIRECTIONS ? 1 : 0;
    int32_t * const rowH = t.rowH[SET], * const rowF = t.rowF[SET], * const colH = t.colH[SET], * const colE = t.colE[SET];
    const int slot = blockIdx.x % maxd;
    const int32_t gapO = prm.gapO, gapE = prm.gapE;
    const int32_t score = t.result[0], a_end = t.result[1], b_end = t.result[2];
    const int nr = PASS == PASS_FORWARD ? t.alen : a_end + 1;
    const int nc = PASS == PASS_FORWARD ? t.blen : b_end + 1;
    if( PASS != PASS_FORWARD && score <= 0 && !prm.nw ) return;   // no local alignment: reported as an error by the host
    const int ntr = ( nr + TILE_R - 1 ) / TILE_R, ntc = ( nc + TILE_C - 1 ) / TILE_C;
    const int ti = ( diag - ( ntc - 1 ) > 0 ? diag - ( ntc - 1 ) : 0 ) + slot;
    const int tj = diag - ti;
    if( ti >= ntr || tj < 0 ) return;
    const int r0 = ti * TILE_R + lane * TB_R;   // first row of this lane
    const int c0 = tj * TILE_C;
    const int ncols = nc - c0 < TILE_C ? nc - c0 : TILE_C;
    const int nvalid = nr - r0 < 0 ? 0 : ( nr - r0 > TB_R ? TB_R : nr - r0 );   // real rows of this lane
    const uint8_t * a = t.a;
    const uint8_t * b = t.b;
    const Bound<PASS> bound{ gapO, gapE, a_end, b_end, prm.nw != 0, t.HHf, t.EEf };

    // profile of this tile's rows: forward / reverse read M[query][db] (align.c:64,106), the direction pass M[db][query] (cigar.c:79,152)
#pragma unroll
    for( int r = 0; r < TB_R; r++ ) {
        const int gr = r0 + r;
        int code = 0;
        if( gr < nr ) code = PASS == PASS_REVERSE ? a[a_end - gr] : a[gr];
        for( int res = 0; res < 32; res++ )
            prof[( r * 32 + res ) * 32 + lane] = __ldg( matrix + ( PASS == PASS_DIRECTIONS ? ( res << 5 ) + code : ( code << 5 ) + res ) );
    }
    __syncwarp();

    int32_t Hl[TB_R], El[TB_R];
#pragma unroll
    for( int r = 0; r < TB_R; r++ ) Hl[r] = El[r] = 0;
    int32_t prev_up = 0, h_out = 0, f_out = 0;
    int32_t best_val = 0, best_pos = 0x7fffffff;     // forward: best cell of this lane; reverse: first cell reaching the score
    const int32_t * rowH_in = rowH + (size_t) ( ti > 0 ? ti - 1 : 0 ) * t.ldc;   // only read when ti > 0
    const int32_t * rowF_in = rowF + (size_t) ( ti > 0 ? ti - 1 : 0 ) * t.ldc;
    const int nrb = ( nr + 7 ) >> 3;

    // the tile's column codes and top boundary go through shared memory (global latency would otherwise sit on every
    // step's critical path); every lane's left boundary is fetched before the sweep starts
    for( int i = lane; i < ncols; i += 32 ) {
        s_code[i] = PASS == PASS_REVERSE ? b[b_end - ( c0 + i )] : b[c0 + i];
        s_toph[i] = ti == 0 ? bound.top_h( c0 + i ) : rowH_in[c0 + i];
        s_topf[i] = ti == 0 ? bound.top_f( c0 + i ) : rowF_in[c0 + i];
    }
    if( nvalid > 0 ) {
        if( tj == 0 ) {
#pragma unroll
            for( int r = 0; r < TB_R; r++ ) {
                Hl[r] = bound.left_h( r0 + r );
                El[r] = bound.left_e( r0 + r );
            }
        }
        else {
            const int32_t * ch = colH + (size_t) ( tj - 1 ) * t.ldr + r0;
            const int32_t * ce = colE + (size_t) ( tj - 1 ) * t.ldr + r0;
#pragma unroll
            for( int r = 0; r < TB_R; r++ ) {
                Hl[r] = ch[r];
                El[r] = ce[r];
            }
        }
        // the cell above-left of this lane's first row
        if( r0 == 0 )
            prev_up = c0 == 0 ? bound.corner() : bound.top_h( c0 - 1 );
        else if( c0 == 0 )
            prev_up = bound.left_h( r0 - 1 );
        else
            prev_up = colH[(size_t) ( tj - 1 ) * t.ldr + r0 - 1];
    }
    __syncwarp();
    uint32_t code_n = lane == 0 ? s_code[0] : 0;   // column code of the next step

    const int nsteps = ncols + 31;
    for( int s = 0; s < nsteps; s++ ) {
        const int cc = s - lane;
        const uint32_t code = code_n;
        code_n = ( cc + 1 >= 0 && cc + 1 < ncols ) ? s_code[cc + 1] : 0;
        int32_t h_up = __shfl_up_sync( 0xffffffffu, h_out, 1 );
        int32_t f_up = __shfl_up_sync( 0xffffffffu, f_out, 1 );
        if( cc < 0 || cc >= ncols || nvalid == 0 ) continue;   // the shuffles above are executed by all lanes
        if( lane == 0 ) {
            h_up = s_toph[cc];
            f_up = s_topf[cc];
        }
        const int gc = c0 + cc;
        int32_t hd = prev_up;
        prev_up = h_up;
        int32_t hu = h_up, fu = f_up;
        const int32_t * pp = prof + ( code << 5 ) + lane;
        uint32_t bits_lo = 0, bits_hi = 0;
#pragma unroll
        for( int r = 0; r < TB_R; r++ ) {
            const int32_t v = pp[r * 1024];
            const int32_t hl = Hl[r], el = El[r];
            int32_t h, e, f;
            if( PASS == PASS_DIRECTIONS ) {   // cigar.c:48-198, strict comparisons
                uint32_t bits = 0;
                h = hd + v;
                f = fu;
                e = el;
                if( f > h ) {
                    bits |= 1;
                    h = f;
                }
                if( e > h ) {
                    bits |= 2;
                    h = e;
                }
                if( !prm.nw && h < 0 ) h = 0;
                const int32_t open = h + gapO + gapE;
                f += gapE;
                e += gapE;
                if( f > open )
                    bits |= 4;
                else
                    f = open;
                if( e > open )
                    bits |= 8;
                else
                    e = open;
                if( r < 4 )
                    bits_lo |= bits << ( 8 * r );
                else
                    bits_hi |= bits << ( 8 * ( r - 4 ) );
            }
            else {   // align.c:56-83 / 99-127: E and F are advanced before H
                f = max2( fu, hu + gapO ) + gapE;
                e = max2( el, hl + gapO ) + gapE;
                h = hd + v;
                if( PASS == PASS_FORWARD && h < 0 ) h = 0;
                if( f > h ) h = f;
                if( e > h ) h = e;
                if( PASS == PASS_FORWARD ) {
                    // first strict maximum in (column outer, row inner) order; a lane meets its cells in that order
                    if( h > best_val && r < nvalid ) {
                        best_val = h;
                        best_pos = ( cc << 3 ) | r;
                    }
                }
                else {
                    if( h >= score && best_pos == 0x7fffffff && r < nvalid ) best_pos = ( cc << 3 ) | r;
                }
            }
            hd = hl;
            Hl[r] = h;
            El[r] = e;
            hu = h;
            fu = f;
        }
        h_out = hu;
        f_out = fu;
        if( PASS == PASS_DIRECTIONS )
            *(uint2 *) ( t.dir + dir_offset( nrb, r0, gc ) ) = make_uint2( bits_lo, bits_hi );
        if( lane == 31 ) {  // bottom boundary of this tile row (only read if a tile row follows)
            rowH[(size_t) ti * t.ldc + gc] = h_out;
            rowF[(size_t) ti * t.ldc + gc] = f_out;
        }
        if( cc == ncols - 1 ) {  // right boundary of this tile column
            int32_t * ch = colH + (size_t) tj * t.ldr + r0;
            int32_t * ce = colE + (size_t) tj * t.ldr + r0;
#pragma unroll
            for( int r = 0; r < TB_R; r++ ) {
                ch[r] = Hl[r];
                ce[r] = El[r];
                if( PASS == PASS_FORWARD && gc == nc - 1 && r < nvalid ) {  // state the reverse pass inherits (align.c:86-89)
                    t.HHf[r0 + r] = Hl[r];
                    t.EEf[r0 + r] = El[r];
                }
            }
        }
    }

    if( PASS != PASS_DIRECTIONS ) {
        // tile result: forward = largest value, then smallest (column, row); reverse = smallest (column, row)
        int64_t key = INT64_MAX;
        if( best_pos != 0x7fffffff ) key = ( (int64_t) ( c0 + ( best_pos >> 3 ) ) << 32 ) | (uint32_t) ( r0 + ( best_pos & 7 ) );
        int32_t val = PASS == PASS_FORWARD ? best_val : 0;
        for( int o = 16; o > 0; o >>= 1 ) {
            const int64_t k2 = __shfl_xor_sync( 0xffffffffu, key, o );
            const int32_t v2 = __shfl_xor_sync( 0xffffffffu, val, o );
            if( v2 > val || ( v2 == val && k2 < key ) ) {
                val = v2;
                key = k2;
            }
        }
        if( lane == 0 ) {
            t.tile_key[ti * t.ntc_cap + tj] = key;
            t.tile_val[ti * t.ntc_cap + tj] = val;
        }
    }
}

// One launch = one anti-diagonal `diag` of the tile grid, for all hits: block = one warp = one tile.  SECOND >= 0 runs a
// second pass beside the first (blockIdx.y selects): the reverse and the direction pass are independent of each other.
template<int FIRST, int SECOND>
__global__ void __launch_bounds__( 32 ) traceback_tile_kernel( const HitTask * tasks, const int32_t * matrix, TbParams prm, int diag, int maxd ) {
    __shared__ int32_t prof[TB_R * 32 * 32];   // [row][residue][lane]
    __shared__ int32_t s_toph[TILE_C], s_topf[TILE_C];
    __shared__ uint8_t s_code[TILE_C];
    if( SECOND >= 0 && blockIdx.y == 1 )
        tile_sweep < SECOND < 0 ? FIRST : SECOND > ( prof, s_toph, s_topf, s_code, tasks, matrix, prm, diag, maxd );
    else
        tile_sweep<FIRST>( prof, s_toph, s_topf, s_code, tasks, matrix, prm, diag, maxd );
}

// combines the tile results of a pass into the hit's result record
template<int PASS>
__global__ void traceback_reduce_kernel( const HitTask * tasks, TbParams prm ) {
    const HitTask t = tasks[blockIdx.x];
    const int32_t score = t.result[0], a_end = t.result[1], b_end = t.result[2];
    const int nr = PASS == PASS_FORWARD ? t.alen : a_end + 1;
    const int nc = PASS == PASS_FORWARD ? t.blen : b_end + 1;
    const int ntr = ( nr + TILE_R - 1 ) / TILE_R, ntc = ( nc + TILE_C - 1 ) / TILE_C;
    int64_t key = INT64_MAX;
    int32_t val = 0;
    if( PASS == PASS_FORWARD || score > 0 )
        for( int i = threadIdx.x; i < ntr * ntc; i += blockDim.x ) {
            const int idx = ( i / ntc ) * t.ntc_cap + ( i % ntc );
            const int64_t k2 = t.tile_key[idx];
            const int32_t v2 = t.tile_val[idx];
            if( v2 > val || ( v2 == val && k2 < key ) ) {
                val = v2;
                key = k2;
            }
        }
    __shared__ int64_t skey[128];
    __shared__ int32_t sval[128];
    skey[threadIdx.x] = key;
    sval[threadIdx.x] = val;
    __syncthreads();
    for( int s = blockDim.x / 2; s > 0; s >>= 1 ) {
        if( (int) threadIdx.x < s ) {
            const int32_t v2 = sval[threadIdx.x + s];
            const int64_t k2 = skey[threadIdx.x + s];
            if( v2 > sval[threadIdx.x] || ( v2 == sval[threadIdx.x] && k2 < skey[threadIdx.x] ) ) {
                sval[threadIdx.x] = v2;
                skey[threadIdx.x] = k2;
            }
        }
        __syncthreads();
    }
    if( threadIdx.x == 0 ) {
        key = skey[0];
        val = sval[0];
        if( PASS == PASS_FORWARD ) {
            t.result[0] = val;
            t.result[1] = val > 0 ? (int32_t) ( key & 0xffffffff ) : 0;  // a_end
            t.result[2] = val > 0 ? (int32_t) ( key >> 32 ) : 0;         // b_end
        }
        else {
            const bool found = key != INT64_MAX && score > 0;
            t.result[5] = found ? 1 : 0;
            t.result[3] = found ? a_end - (int32_t) ( key & 0xffffffff ) : 0;  // a_begin
            t.result[4] = found ? b_end - (int32_t) ( key >> 32 ) : 0;         // b_begin
        }
    }
}

// cigar.c:299-346: backward walk over the direction matrix, run-length encoded in walk order
// One warp per hit, lane 0 walks: the walk is a serial chain of dependent loads with data-dependent branches, and
// several walks in one warp would serialise each other's divergent paths.
__global__ void traceback_walk_kernel( const HitTask * tasks, int ntasks ) {
    const int k = blockIdx.x;
    if( k >= ntasks || threadIdx.x != 0 ) return;
    const HitTask t = tasks[k];
    const int32_t a_end = t.result[1], b_end = t.result[2], a_begin = t.result[3], b_begin = t.result[4];
    const int nrb = ( a_end + 1 + 7 ) >> 3;
    int i = a_end, j = b_end;
    int nruns = 0;
    uint32_t prev = 0, count = 0;
    while( i >= 0 && j >= 0 && i >= a_begin && j >= b_begin ) {
        const uint8_t d = t.dir[dir_offset( nrb, i, j )];
        uint32_t op;
        if( d & ( 2 | 8 ) ) {
            j--;
            op = 'I';
        }
        else if( d & ( 1 | 4 ) ) {
            i--;
            op = 'D';
        }
        else {
            i--;
            j--;
            op = 'M';
        }
        if( op == prev ) {
            count++;
        }
        else {
            if( count > 0 && nruns < t.runs_cap ) t.runs[nruns++] = ( count << 8 ) | prev;
            prev = op;
            count = 1;
        }
    }
    if( count > 0 && nruns < t.runs_cap ) t.runs[nruns++] = ( count << 8 ) | prev;
    t.result[6] = nruns;
}

struct TbState {
    void * d_pool = nullptr;
    size_t pool_cap = 0;
    void * h_stage = nullptr;   // pinned staging buffer for the one upload / one download of a call
    size_t stage_cap = 0;
    int32_t * d_matrix = nullptr;
    uint64_t matrix_version = ~0ull;
} tb;

void * pool_reserve( size_t bytes ) {
    if( bytes > tb.pool_cap ) {
        if( tb.d_pool ) cudaFree( tb.d_pool );
        tb.pool_cap = bytes + bytes / 4 + 4096;
        CUDA_OK( cudaMalloc( &tb.d_pool, tb.pool_cap ) );
    }
    return tb.d_pool;
}

void * stage_reserve( size_t bytes ) {
    if( bytes > tb.stage_cap ) {
        if( tb.h_stage ) cudaFreeHost( tb.h_stage );
        tb.stage_cap = bytes + bytes / 4 + 4096;
        CUDA_OK( cudaHostAlloc( &tb.h_stage, tb.stage_cap, cudaHostAllocDefault ) );
    }
    return tb.h_stage;
}

size_t align_up( size_t v, size_t a ) { return ( v + a - 1 ) / a * a; }

}  // namespace

void traceback_release() {
    if( tb.d_pool ) cudaFree( tb.d_pool );
    if( tb.d_matrix ) cudaFree( tb.d_matrix );
    if( tb.h_stage ) cudaFreeHost( tb.h_stage );
    tb.h_stage = nullptr;
    tb.stage_cap = 0;
    tb.d_pool = nullptr;
    tb.pool_cap = 0;
    tb.d_matrix = nullptr;
    tb.matrix_version = ~0ull;
}

void traceback_hits( cudaStream_t stream, const uint8_t * query, size_t qlen, bool nw, std::vector<HitAlignment> & hits,
        ssa_b200_stats_t * stats, bool scored_sequence ) {
    const std::vector<TbQuery> queries( 1, TbQuery{ query, qlen } );
    const std::vector<uint32_t> query_of_hit( hits.size(), 0 );
    const std::vector<char> scored( hits.size(), scored_sequence ? 1 : 0 );
    traceback_hits_multi( stream, queries, query_of_hit, nw, hits, stats, scored );
}

// The hits of SEVERAL queries in one go (ssa_b200_align_batch aligns the hits of many queries together: the tile launches are
// latency-bound for a single query's ten hits - 0.3 ms for 350 x 350 pairs - and throughput-bound for thousands of tiles).
void traceback_hits_multi( cudaStream_t stream, const std::vector<TbQuery> & queries, const std::vector<uint32_t> & query_of_hit, bool nw,
        std::vector<HitAlignment> & hits, ssa_b200_stats_t * stats, const std::vector<char> & scored_sequence ) {
    const int n = (int) hits.size();
    if( n == 0 ) return;
    auto qlen_of = [&]( int k ) { return queries[query_of_hit[k]].len; };
    if( !tb.d_matrix ) CUDA_OK( cudaMalloc( &tb.d_matrix, 1024 * sizeof( int32_t ) ) );
    if( tb.matrix_version != g_matrix.version ) {
        int32_t m32[1024];
        for( int i = 0; i < 1024; i++ ) m32[i] = (int32_t) g_matrix.m[i];
        CUDA_OK( cudaMemcpyAsync( tb.d_matrix, m32, sizeof( m32 ), cudaMemcpyHostToDevice, stream ) );
        CUDA_OK( cudaStreamSynchronize( stream ) );
        tb.matrix_version = g_matrix.version;
        stats->h2d_bytes += sizeof( m32 );
    }
    TbParams prm{ g_gapO, g_gapE, nw ? 1 : 0 };

    // Everything is enqueued without a host round trip in between: the passes after the forward pass take their
    // extents (a_end, b_end) from the device-side result record, launch counts and buffer sizes use the upper bounds
    // (query length x database sequence length).  One upload (tasks, result records, sequences), one download
    // (result records, run lists).
    auto tiles_r = []( size_t rows ) { return (int) ( ( rows + TILE_R - 1 ) / TILE_R ); };
    auto tiles_c = []( size_t cols ) { return (int) ( ( cols + TILE_C - 1 ) / TILE_C ); };
    auto dir_bytes = []( size_t rows, size_t cols ) { return ( ( cols + 15 ) / 16 ) * ( ( rows + 7 ) / 8 ) * 128; };
    size_t off = 0;
    auto carve = [&]( size_t bytes ) {
        const size_t at = off;
        off = align_up( off + bytes, 256 );
        return at;
    };
    // upload blob
    const size_t off_tasks = carve( n * sizeof( HitTask ) );
    std::vector<size_t> off_query( queries.size() );
    for( size_t q = 0; q < queries.size(); q++ ) off_query[q] = carve( queries[q].len + 1 );
    std::vector<size_t> off_b( n );
    for( int k = 0; k < n; k++ ) off_b[k] = carve( hits[k].db_codes.size() + 1 );
    // download blob: result records (also uploaded: NW presets them), run lists
    const size_t off_results = carve( (size_t) n * 8 * sizeof( int32_t ) );
    const size_t upload_bytes = off;
    std::vector<size_t> off_runs( n );
    for( int k = 0; k < n; k++ ) off_runs[k] = carve( ( qlen_of( k ) + hits[k].db_codes.size() + 2 ) * sizeof( uint32_t ) );
    const size_t download_bytes = off - off_results;
    // device-only: tile boundaries, partial results, direction bytes
    std::vector<std::vector<size_t>> offs( n );
    for( int k = 0; k < n; k++ ) {
        const size_t dlen = hits[k].db_codes.size(), qlen = qlen_of( k );
        const size_t ntr = (size_t) tiles_r( qlen ), ntc = (size_t) tiles_c( dlen );
        const size_t ldr = ntr * TILE_R, ldc = ntc * TILE_C;
        std::vector<size_t> & o = offs[k];
        for( int set = 0; set < 2; set++ ) {
            o.push_back( carve( ntr * ldc * sizeof( int32_t ) ) );   // rowH
            o.push_back( carve( ntr * ldc * sizeof( int32_t ) ) );   // rowF
            o.push_back( carve( ntc * ldr * sizeof( int32_t ) ) );   // colH
            o.push_back( carve( ntc * ldr * sizeof( int32_t ) ) );   // colE
        }
        o.push_back( carve( ldr * sizeof( int32_t ) ) );             // HHf
        o.push_back( carve( ldr * sizeof( int32_t ) ) );             // EEf
        o.push_back( carve( ntr * ntc * sizeof( int64_t ) ) );       // tile_key
        o.push_back( carve( ntr * ntc * sizeof( int32_t ) ) );       // tile_val
        o.push_back( carve( dir_bytes( qlen, dlen ) ) );             // dir
    }
    uint8_t * pool = (uint8_t *) pool_reserve( off );
    uint8_t * stage = (uint8_t *) stage_reserve( std::max( upload_bytes, download_bytes ) );

    HitTask * tasks = (HitTask *) ( stage + off_tasks );
    for( size_t q = 0; q < queries.size(); q++ ) memcpy( stage + off_query[q], queries[q].seq, queries[q].len );
    int ntr_max = 0, ntc_max = 0;
    for( int k = 0; k < n; k++ ) {
        HitTask & t = tasks[k];
        const size_t dlen = hits[k].db_codes.size(), qlen = qlen_of( k );
        const std::vector<size_t> & o = offs[k];
        t.a = pool + off_query[query_of_hit[k]];
        t.b = pool + off_b[k];
        t.alen = (int32_t) qlen;
        t.blen = (int32_t) dlen;
        for( int set = 0; set < 2; set++ ) {
            t.rowH[set] = (int32_t *) ( pool + o[4 * set] );
            t.rowF[set] = (int32_t *) ( pool + o[4 * set + 1] );
            t.colH[set] = (int32_t *) ( pool + o[4 * set + 2] );
            t.colE[set] = (int32_t *) ( pool + o[4 * set + 3] );
        }
        t.HHf = (int32_t *) ( pool + o[8] );
        t.EEf = (int32_t *) ( pool + o[9] );
        t.tile_key = (int64_t *) ( pool + o[10] );
        t.tile_val = (int32_t *) ( pool + o[11] );
        t.ldr = tiles_r( qlen ) * TILE_R;
        t.ldc = tiles_c( dlen ) * TILE_C;
        t.ntc_cap = tiles_c( dlen );
        t.dir = pool + o[12];
        t.result = (int32_t *) ( pool + off_results ) + 8 * k;
        t.runs = (uint32_t *) ( pool + off_runs[k] );
        t.runs_cap = (int32_t) ( qlen + dlen + 2 );
        ntr_max = std::max( ntr_max, tiles_r( qlen ) );
        ntc_max = std::max( ntc_max, tiles_c( dlen ) );
        memcpy( stage + off_b[k], hits[k].db_codes.data(), dlen );
        int32_t * r = (int32_t *) ( stage + off_results ) + 8 * k;
        memset( r, 0, 8 * sizeof( int32_t ) );
        if( nw ) {  // align.c:141-150: the whole rectangle
            r[0] = (int32_t) hits[k].hit.score;
            r[1] = t.alen - 1;
            r[2] = t.blen - 1;
            r[5] = 1;
        }
    }
    CUDA_OK( cudaMemcpyAsync( pool, stage, upload_bytes, cudaMemcpyHostToDevice, stream ) );
    stats->h2d_bytes += upload_bytes;
    const HitTask * d_tasks = (const HitTask *) ( pool + off_tasks );
    const int maxd = std::min( ntr_max, ntc_max ), ndiag = ntr_max + ntc_max - 1;
    if( !nw ) {  // SW: the forward pass finds the region end
        for( int d = 0; d < ndiag; d++ )
            traceback_tile_kernel<PASS_FORWARD, -1><<<n * maxd, 32, 0, stream>>>( d_tasks, tb.d_matrix, prm, d, maxd );
        traceback_reduce_kernel<PASS_FORWARD><<<n, 128, 0, stream>>>( d_tasks, prm );
        stats->launches += ndiag + 1;
    }
    // direction bytes over (a_end+1) x (b_end+1) beside the reverse pass (SW: region begin), then the walk
    for( int d = 0; d < ndiag; d++ ) {
        if( nw )
            traceback_tile_kernel<PASS_DIRECTIONS, -1><<<n * maxd, 32, 0, stream>>>( d_tasks, tb.d_matrix, prm, d, maxd );
        else
            traceback_tile_kernel<PASS_REVERSE, PASS_DIRECTIONS><<<dim3( n * maxd, 2 ), 32, 0, stream>>>( d_tasks, tb.d_matrix, prm, d, maxd );
    }
    if( !nw ) traceback_reduce_kernel<PASS_REVERSE><<<n, 128, 0, stream>>>( d_tasks, prm );
    traceback_walk_kernel<<<n, 32, 0, stream>>>( d_tasks, n );
    stats->launches += ndiag + 2;
    CUDA_OK( cudaMemcpyAsync( stage, pool + off_results, download_bytes, cudaMemcpyDeviceToHost, stream ) );
    CUDA_OK( cudaStreamSynchronize( stream ) );
    stats->d2h_bytes += download_bytes;

    bool matrix_symmetric = true;
    for( int x = 0; x < SCORE_MATRIX_DIM && matrix_symmetric; x++ )
        for( int y = 0; y < x; y++ )
            if( g_matrix.m[( x << 5 ) + y] != g_matrix.m[( y << 5 ) + x] ) {
                matrix_symmetric = false;
                break;
            }
    for( int k = 0; k < n; k++ ) {
        const int32_t * r = (const int32_t *) stage + 8 * k;
        // The search scores with M[(db << 5) + query], the region passes read M[query][db] like the reference's align.c:
        // with an asymmetric user matrix the two scores legitimately differ, and the reference returns the alignment of
        // its own forward score (align.c:39-131).  The cross-check is therefore an internal assertion for symmetric
        // matrices only.
        if( scored_sequence[k] && !nw && matrix_symmetric && r[0] != (int32_t) hits[k].hit.score )
            fatal( "libssa_b200: internal error: alignment score %d differs from search score %ld for sequence %ld", r[0],
                    (long) hits[k].hit.score, (long) hits[k].hit.db_id );
        if( !r[5] ) fatal( "Internal error in align function." );  // align.c:131
        const uint32_t * runs = (const uint32_t *) ( stage + ( off_runs[k] - off_results ) );
        const int nruns = r[6];
        std::string cigar;  // cigar.c:221-263: runs front to back, count omitted when 1
        for( int i = nruns - 1; i >= 0; i-- ) {
            const uint32_t count = runs[i] >> 8;
            if( count > 1 ) cigar += std::to_string( count );
            cigar.push_back( (char) ( runs[i] & 0xff ) );
        }
        hits[k].cigar = cigar;
        hits[k].q_start = (size_t) r[3];
        hits[k].q_end = (size_t) r[1];
        hits[k].d_start = (size_t) r[4];
        hits[k].d_end = (size_t) r[2];
    }
}
