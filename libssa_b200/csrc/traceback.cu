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
// One CTA per hit sweeps the DP matrix by anti-diagonals; every cell needs its diagonal, upper and left
// neighbour, which live in rotating per-diagonal arrays (3 x H, 2 x E, 2 x F).  All three passes share that
// engine and differ only in the cell function and the boundary values.  Arithmetic is int32, exact for the
// supported lengths (< 2^24 symbols: |value| < 2^31).
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

constexpr int TB_THREADS = 512;

enum Pass { PASS_FORWARD = 0, PASS_REVERSE = 1, PASS_DIRECTIONS = 2 };

struct HitTask {              // one reported hit
    const uint8_t * a;        // query codes (shared by all hits)
    const uint8_t * b;        // database sequence codes
    int32_t alen, blen;
    int32_t * scratch;        // 9 * alen int32: H[3], E[2], F[2], final HH, final EE
    uint8_t * dir;            // (a_end+1) x (b_end+1) direction bytes, column-major
    int32_t * result;         // [0] score [1] a_end [2] b_end [3] a_begin [4] b_begin [5] found [6] nruns
    uint32_t * runs;          // walk output: (count << 8) | op, in walk order
    int32_t runs_cap;
};

struct TbParams {
    int32_t gapO, gapE;
    int32_t nw;
};

__device__ __forceinline__ int32_t max2( int32_t a, int32_t b ) { return a > b ? a : b; }

// Sweeps rows 0..nr-1 x columns 0..nc-1 by anti-diagonals.  F(r, c, diag, h_up, f_up, h_left, e_left, &h, &e, &f)
// computes one cell; B supplies the boundary values of row -1 / column -1.
template<typename Cell, typename Bound>
__device__ void sweep( int nr, int nc, int32_t * H0, int32_t * H1, int32_t * H2, int32_t * E0, int32_t * E1, int32_t * F0, int32_t * F1,
        Cell cell, Bound bound ) {
    int32_t * H[3] = { H0, H1, H2 };
    int32_t * E[2] = { E0, E1 };
    int32_t * F[2] = { F0, F1 };
    const int ndiag = nr + nc - 1;
    for( int d = 0; d < ndiag; d++ ) {
        const int lo = d - nc + 1 > 0 ? d - nc + 1 : 0;
        const int hi = d < nr - 1 ? d : nr - 1;
        int32_t * Hc = H[d % 3];
        const int32_t * Hp = H[( d + 2 ) % 3];   // diagonal d-1
        const int32_t * Hpp = H[( d + 1 ) % 3];  // diagonal d-2
        int32_t * Ec = E[d & 1];
        const int32_t * Ep = E[( d + 1 ) & 1];
        int32_t * Fc = F[d & 1];
        const int32_t * Fp = F[( d + 1 ) & 1];
        for( int r = lo + (int) threadIdx.x; r <= hi; r += blockDim.x ) {
            const int c = d - r;
            int32_t diag, h_up, f_up, h_left, e_left;
            if( r == 0 ) {
                h_up = bound.top_h( c );
                f_up = bound.top_f( c );
                diag = c == 0 ? bound.corner() : bound.top_diag( c );
            }
            else {
                h_up = Hp[r - 1];
                f_up = Fp[r - 1];
                diag = c == 0 ? bound.left_h( r - 1 ) : Hpp[r - 1];
            }
            if( c == 0 ) {
                h_left = bound.left_h( r );
                e_left = bound.left_e( r );
            }
            else {
                h_left = Hp[r];
                e_left = Ep[r];
            }
            int32_t h, e, f;
            cell( r, c, diag, h_up, f_up, h_left, e_left, h, e, f );
            Hc[r] = h;
            Ec[r] = e;
            Fc[r] = f;
        }
        __syncthreads();
    }
}

// ---- block-wide lexicographic reductions ----
__device__ void block_best_max( int64_t & key, int32_t & val ) {  // max val, ties: min key
    __shared__ int64_t skey[TB_THREADS];
    __shared__ int32_t sval[TB_THREADS];
    skey[threadIdx.x] = key;
    sval[threadIdx.x] = val;
    __syncthreads();
    for( int s = blockDim.x / 2; s > 0; s >>= 1 ) {
        if( (int) threadIdx.x < s ) {
            int32_t v2 = sval[threadIdx.x + s];
            int64_t k2 = skey[threadIdx.x + s];
            if( v2 > sval[threadIdx.x] || ( v2 == sval[threadIdx.x] && k2 < skey[threadIdx.x] ) ) {
                sval[threadIdx.x] = v2;
                skey[threadIdx.x] = k2;
            }
        }
        __syncthreads();
    }
    key = skey[0];
    val = sval[0];
    __syncthreads();
}

__device__ void block_min_key( int64_t & key ) {
    __shared__ int64_t skey[TB_THREADS];
    skey[threadIdx.x] = key;
    __syncthreads();
    for( int s = blockDim.x / 2; s > 0; s >>= 1 ) {
        if( (int) threadIdx.x < s && skey[threadIdx.x + s] < skey[threadIdx.x] ) skey[threadIdx.x] = skey[threadIdx.x + s];
        __syncthreads();
    }
    key = skey[0];
    __syncthreads();
}

__global__ void __launch_bounds__( TB_THREADS ) traceback_pass_kernel( const HitTask * tasks, const int32_t * matrix, TbParams prm, int pass ) {
    __shared__ int32_t M[1024];
    for( int i = threadIdx.x; i < 1024; i += blockDim.x ) M[i] = matrix[i];
    __syncthreads();
    const HitTask t = tasks[blockIdx.x];
    const int32_t gapO = prm.gapO, gapE = prm.gapE;
    const uint8_t * a = t.a;
    const uint8_t * b = t.b;
    int32_t * S = t.scratch;
    const int alen = t.alen, blen = t.blen;
    int32_t * H0 = S, * H1 = S + alen, * H2 = S + 2 * alen, * E0 = S + 3 * alen, * E1 = S + 4 * alen, * F0 = S + 5 * alen,
            * F1 = S + 6 * alen, * HHf = S + 7 * alen, * EEf = S + 8 * alen;

    if( pass == PASS_FORWARD ) {
        // align.c:49-83.  rows = query positions, columns = database positions.
        int64_t best_key = INT64_MAX;
        int32_t best_val = 0;
        struct B {
            int32_t gapO;
            __device__ int32_t top_h( int ) const { return 0; }
            __device__ int32_t top_f( int ) const { return gapO; }
            __device__ int32_t top_diag( int ) const { return 0; }
            __device__ int32_t corner() const { return 0; }
            __device__ int32_t left_h( int ) const { return 0; }
            __device__ int32_t left_e( int ) const { return gapO; }
        } bound{ gapO };
        auto cell = [&]( int r, int c, int32_t diag, int32_t h_up, int32_t f_up, int32_t h_left, int32_t e_left, int32_t & h,
                            int32_t & e, int32_t & f ) {
            f = max2( f_up, h_up + gapO ) + gapE;
            e = max2( e_left, h_left + gapO ) + gapE;
            h = diag + M[( (int) a[r] << 5 ) + b[c]];
            if( h < 0 ) h = 0;
            if( f > h ) h = f;
            if( e > h ) h = e;
            // first strict maximum in (column outer, row inner) order = largest value, then smallest (c, r)
            const int64_t key = ( (int64_t) c << 32 ) | (uint32_t) r;
            if( h > best_val || ( h == best_val && h > 0 && key < best_key ) ) {
                best_val = h;
                best_key = key;
            }
            if( c == blen - 1 ) {  // state the reverse pass inherits for rows it fails to re-initialise
                HHf[r] = h;
                EEf[r] = e;
            }
        };
        sweep( alen, blen, H0, H1, H2, E0, E1, F0, F1, cell, bound );
        block_best_max( best_key, best_val );
        if( threadIdx.x == 0 ) {
            t.result[0] = best_val;
            t.result[1] = best_val > 0 ? (int32_t) ( best_key & 0xffffffff ) : 0;  // a_end
            t.result[2] = best_val > 0 ? (int32_t) ( best_key >> 32 ) : 0;         // b_end
        }
    }
    else if( pass == PASS_REVERSE ) {
        // align.c:86-127 on reversed coordinates: row r <-> query position a_end - r, column c <-> db position b_end - c
        const int32_t score = t.result[0], a_end = t.result[1], b_end = t.result[2];
        const int nr = a_end + 1, nc = b_end + 1;
        int64_t first_key = INT64_MAX;
        struct B {
            int32_t a_end, b_end;
            const int32_t * HHf, * EEf;
            __device__ int32_t top_h( int ) const { return -1; }
            __device__ int32_t top_f( int ) const { return -1; }
            __device__ int32_t top_diag( int ) const { return -1; }
            __device__ int32_t corner() const { return 0; }
            // query positions above b_end were not reset to -1 (align.c:86-89): they keep the forward pass' last column
            __device__ int32_t left_h( int r ) const { int j = a_end - r; return j > b_end ? HHf[j] : -1; }
            __device__ int32_t left_e( int r ) const { int j = a_end - r; return j > b_end ? EEf[j] : -1; }
        } bound{ a_end, b_end, HHf, EEf };
        auto cell = [&]( int r, int c, int32_t diag, int32_t h_up, int32_t f_up, int32_t h_left, int32_t e_left, int32_t & h,
                            int32_t & e, int32_t & f ) {
            f = max2( f_up, h_up + gapO ) + gapE;
            e = max2( e_left, h_left + gapO ) + gapE;
            h = diag + M[( (int) a[a_end - r] << 5 ) + b[b_end - c]];
            if( f > h ) h = f;
            if( e > h ) h = e;
            if( h >= score ) {
                const int64_t key = ( (int64_t) c << 32 ) | (uint32_t) r;
                if( key < first_key ) first_key = key;
            }
        };
        sweep( nr, nc, H0, H1, H2, E0, E1, F0, F1, cell, bound );
        block_min_key( first_key );
        if( threadIdx.x == 0 ) {
            const bool found = first_key != INT64_MAX && score > 0;
            t.result[5] = found ? 1 : 0;
            t.result[3] = found ? a_end - (int32_t) ( first_key & 0xffffffff ) : 0;  // a_begin
            t.result[4] = found ? b_end - (int32_t) ( first_key >> 32 ) : 0;         // b_begin
        }
    }
    else {
        // cigar.c:48-198 restricted to the rectangle the walk can reach: rows 0..a_end, columns 0..b_end
        const int32_t a_end = t.result[1], b_end = t.result[2];
        const int nr = a_end + 1, nc = b_end + 1;
        const bool nw = prm.nw != 0;
        uint8_t * dir = t.dir;
        struct B {
            int32_t gapO, gapE;
            bool nw;
            __device__ int32_t top_h( int ) const { return 0; }  // unused by this cell function
            __device__ int32_t top_f( int c ) const { return nw ? 2 * gapO + ( c + 2 ) * gapE : 0; }
            __device__ int32_t top_diag( int c ) const { return nw ? gapO + c * gapE : 0; }
            __device__ int32_t corner() const { return 0; }
            __device__ int32_t left_h( int r ) const { return nw ? gapO + ( r + 1 ) * gapE : 0; }
            __device__ int32_t left_e( int r ) const { return nw ? 2 * gapO + ( r + 2 ) * gapE : 0; }
        } bound{ gapO, gapE, nw };
        auto cell = [&]( int r, int c, int32_t diag, int32_t h_up, int32_t f_up, int32_t h_left, int32_t e_left, int32_t & h,
                            int32_t & e, int32_t & f ) {
            uint8_t bits = 0;
            h = diag + M[( (int) b[c] << 5 ) + a[r]];
            f = f_up;
            e = e_left;
            if( f > h ) {
                bits |= 1;
                h = f;
            }
            if( e > h ) {
                bits |= 2;
                h = e;
            }
            if( !nw && h < 0 ) h = 0;
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
            dir[(size_t) nr * c + r] = bits;
        };
        sweep( nr, nc, H0, H1, H2, E0, E1, F0, F1, cell, bound );
    }
}

// cigar.c:299-346: backward walk over the direction matrix, run-length encoded in walk order
__global__ void traceback_walk_kernel( const HitTask * tasks, int ntasks ) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if( k >= ntasks ) return;
    const HitTask t = tasks[k];
    const int32_t a_end = t.result[1], b_end = t.result[2], a_begin = t.result[3], b_begin = t.result[4];
    const int nr = a_end + 1;
    int i = a_end, j = b_end;
    int nruns = 0;
    uint32_t prev = 0, count = 0;
    while( i >= 0 && j >= 0 && i >= a_begin && j >= b_begin ) {
        const uint8_t d = t.dir[(size_t) nr * j + i];
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

size_t align_up( size_t v, size_t a ) { return ( v + a - 1 ) / a * a; }

}  // namespace

void traceback_release() {
    if( tb.d_pool ) cudaFree( tb.d_pool );
    if( tb.d_matrix ) cudaFree( tb.d_matrix );
    tb.d_pool = nullptr;
    tb.pool_cap = 0;
    tb.d_matrix = nullptr;
    tb.matrix_version = ~0ull;
}

void traceback_hits( cudaStream_t stream, const uint8_t * query, size_t qlen, bool nw, std::vector<HitAlignment> & hits,
        ssa_b200_stats_t * stats, bool scored_sequence ) {
    const int n = (int) hits.size();
    if( n == 0 ) return;
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

    // ---- stage A: sequences, scratch, results; SW forward pass finds the region end ----
    size_t off = 0;
    const size_t off_tasks = off;
    off = align_up( off + n * sizeof( HitTask ), 256 );
    const size_t off_query = off;
    off = align_up( off + qlen + 1, 256 );
    std::vector<size_t> off_b( n ), off_scratch( n ), off_result( n ), off_runs( n ), off_dir( n );
    for( int k = 0; k < n; k++ ) {
        off_b[k] = off;
        off = align_up( off + hits[k].db_codes.size() + 1, 256 );
        off_scratch[k] = off;
        off = align_up( off + 9 * qlen * sizeof( int32_t ), 256 );
        off_result[k] = off;
        off = align_up( off + 8 * sizeof( int32_t ), 256 );
        off_runs[k] = off;
        off = align_up( off + ( qlen + hits[k].db_codes.size() + 2 ) * sizeof( uint32_t ), 256 );
    }
    const size_t stage_a_bytes = off;
    // the direction matrices are sized after the forward pass; reserve the worst case up front so that the pool
    // does not move between the stages
    size_t worst = stage_a_bytes;
    for( int k = 0; k < n; k++ ) worst = align_up( worst + qlen * hits[k].db_codes.size(), 256 );
    uint8_t * pool = (uint8_t *) pool_reserve( worst );

    std::vector<HitTask> tasks( n );
    CUDA_OK( cudaMemcpyAsync( pool + off_query, query, qlen, cudaMemcpyHostToDevice, stream ) );
    stats->h2d_bytes += qlen;
    for( int k = 0; k < n; k++ ) {
        HitTask & t = tasks[k];
        t.a = pool + off_query;
        t.b = pool + off_b[k];
        t.alen = (int32_t) qlen;
        t.blen = (int32_t) hits[k].db_codes.size();
        t.scratch = (int32_t *) ( pool + off_scratch[k] );
        t.dir = nullptr;
        t.result = (int32_t *) ( pool + off_result[k] );
        t.runs = (uint32_t *) ( pool + off_runs[k] );
        t.runs_cap = (int32_t) ( qlen + hits[k].db_codes.size() + 2 );
        CUDA_OK( cudaMemcpyAsync( pool + off_b[k], hits[k].db_codes.data(), hits[k].db_codes.size(), cudaMemcpyHostToDevice, stream ) );
        stats->h2d_bytes += hits[k].db_codes.size();
    }
    std::vector<int32_t> res( 8 * n, 0 );
    if( nw ) {  // align.c:141-150: the whole rectangle
        for( int k = 0; k < n; k++ ) {
            int32_t r[8] = { (int32_t) hits[k].hit.score, tasks[k].alen - 1, tasks[k].blen - 1, 0, 0, 1, 0, 0 };
            CUDA_OK( cudaMemcpyAsync( tasks[k].result, r, sizeof( r ), cudaMemcpyHostToDevice, stream ) );
            memcpy( &res[8 * k], r, sizeof( r ) );
        }
    }
    CUDA_OK( cudaMemcpyAsync( pool + off_tasks, tasks.data(), n * sizeof( HitTask ), cudaMemcpyHostToDevice, stream ) );
    if( !nw ) {
        traceback_pass_kernel<<<n, TB_THREADS, 0, stream>>>( (const HitTask *) ( pool + off_tasks ), tb.d_matrix, prm, PASS_FORWARD );
        stats->launches++;
        for( int k = 0; k < n; k++ )
            CUDA_OK( cudaMemcpyAsync( &res[8 * k], tasks[k].result, 8 * sizeof( int32_t ), cudaMemcpyDeviceToHost, stream ) );
        CUDA_OK( cudaStreamSynchronize( stream ) );
        stats->d2h_bytes += n * 8 * sizeof( int32_t );
    }

    // ---- stage B: direction matrices over (a_end+1) x (b_end+1), reverse pass, walk ----
    off = stage_a_bytes;
    for( int k = 0; k < n; k++ ) {
        const size_t cells = (size_t) ( res[8 * k + 1] + 1 ) * (size_t) ( res[8 * k + 2] + 1 );
        off_dir[k] = off;
        off = align_up( off + cells, 256 );
        tasks[k].dir = pool + off_dir[k];
    }
    CUDA_OK( cudaMemcpyAsync( pool + off_tasks, tasks.data(), n * sizeof( HitTask ), cudaMemcpyHostToDevice, stream ) );
    if( !nw ) {
        traceback_pass_kernel<<<n, TB_THREADS, 0, stream>>>( (const HitTask *) ( pool + off_tasks ), tb.d_matrix, prm, PASS_REVERSE );
        stats->launches++;
    }
    traceback_pass_kernel<<<n, TB_THREADS, 0, stream>>>( (const HitTask *) ( pool + off_tasks ), tb.d_matrix, prm, PASS_DIRECTIONS );
    traceback_walk_kernel<<<( n + 31 ) / 32, 32, 0, stream>>>( (const HitTask *) ( pool + off_tasks ), n );
    stats->launches += 2;
    for( int k = 0; k < n; k++ )
        CUDA_OK( cudaMemcpyAsync( &res[8 * k], tasks[k].result, 8 * sizeof( int32_t ), cudaMemcpyDeviceToHost, stream ) );
    CUDA_OK( cudaStreamSynchronize( stream ) );
    stats->d2h_bytes += n * 8 * sizeof( int32_t );

    std::vector<uint32_t> runs;
    for( int k = 0; k < n; k++ ) {
        const int32_t * r = &res[8 * k];
        if( scored_sequence && !nw && r[0] != (int32_t) hits[k].hit.score )
            fatal( "libssa_b200: internal error: alignment score %d differs from search score %ld for sequence %ld", r[0],
                    (long) hits[k].hit.score, (long) hits[k].hit.db_id );
        if( !r[5] ) fatal( "Internal error in align function." );  // align.c:131
        const int nruns = r[6];
        runs.resize( nruns );
        if( nruns > 0 ) {
            CUDA_OK( cudaMemcpyAsync( runs.data(), tasks[k].runs, nruns * sizeof( uint32_t ), cudaMemcpyDeviceToHost, stream ) );
            CUDA_OK( cudaStreamSynchronize( stream ) );
            stats->d2h_bytes += nruns * sizeof( uint32_t );
        }
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
