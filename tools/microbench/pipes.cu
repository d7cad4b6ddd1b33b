// Integer-pipe throughput microbenchmark for sm_100a (B200).
// Measures lane-ops / clk / SM for the SASS instructions the SW/NW scoring
// kernels are built from, alone and co-issued, to fix the "L" of the
// integer-pipe roofline.  Build: see Makefile target `pipes`.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){fprintf(stderr,"CUDA %s @%d: %s\n",#x,__LINE__,cudaGetErrorString(e)); exit(1);} }while(0)

constexpr int NCH = 8;      // independent dependency chains per thread
constexpr int ITER = 4096;  // loop trips (each trip = UNR ops per chain)
constexpr int UNR = 8;

enum Op { OP_IADD, OP_IMAD, OP_VIMNMX3, OP_VIADDMNMX, OP_VIMNMX2, OP_VIADD2, OP_HFMA2, OP_HMNMX2, OP_PRMT, OP_LOP3,
          OP_MIX_IMAD_VIMNMX3, OP_MIX_IMAD_VIADDMNMX, OP_MIX_HFMA2_VIMNMX3, OP_MIX_IADD_VIMNMX3, OP_SHFL, OP_LDS,
          OP_MIX_LDS_VIMNMX3, OP_MIX3, OP_COUNT };
static const char* names[] = {"IADD3","IMAD","VIMNMX3.U16x2","VIADDMNMX.U16x2","VIMNMX.S16x2","VIADD.16x2","HFMA2","HMNMX2","PRMT","LOP3",
          "IMAD+VIMNMX3 (1:1)","IMAD+VIADDMNMX (1:1)","HFMA2+VIMNMX3 (1:1)","IADD3+VIMNMX3 (1:1)","SHFL.UP","LDS.32",
          "LDS+VIMNMX3 (1:1)","IMAD+VIMNMX3+LDS (4:4:1)"};
// ops counted per "unit" of the unrolled body, per chain
static const int ops_per_unit[] = {1,1,1,1,1,1,1,1,1,1, 2,2,2,2, 1,1, 2,9};

template<int OP>
__global__ void __launch_bounds__(256) bench(unsigned* out, unsigned seed, unsigned one, long long* clk) {
    __shared__ unsigned sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * seed;
    __syncthreads();
    unsigned a[NCH], b[NCH];
    #pragma unroll
    for (int c = 0; c < NCH; c++) { a[c] = seed * (threadIdx.x + 1 + c); b[c] = seed ^ (c * 77 + threadIdx.x); }
    unsigned k1 = seed | 0x00010001u, k2 = seed * 3u;
    unsigned lidx = threadIdx.x;  // conflict-free LDS index
    long long t0 = clock64();
    for (int it = 0; it < ITER; it++) {
        #pragma unroll
        for (int u = 0; u < UNR; u++) {
            #pragma unroll
            for (int c = 0; c < NCH; c++) {
                if (OP == OP_IADD)      { asm volatile("add.u32 %0,%0,%1;" : "+r"(a[c]) : "r"(k1)); }
                if (OP == OP_IMAD)      { asm volatile("mad.lo.u32 %0,%0,%1,%2;" : "+r"(a[c]) : "r"(one), "r"(k1)); }
                if (OP == OP_VIMNMX3)   { a[c] = __vimax3_u16x2(a[c], b[c], k1); }
                if (OP == OP_VIADDMNMX) { a[c] = __viaddmax_u16x2(a[c], k1, b[c]); }
                if (OP == OP_VIMNMX2)   { a[c] = __vmaxs2(a[c], b[c]) ^ 0; asm volatile("" : "+r"(a[c])); }
                if (OP == OP_VIADD2)    { a[c] = __vadd2(a[c], k1); }
                if (OP == OP_HFMA2)     { asm volatile("fma.rn.f16x2 %0,%0,%1,%2;" : "+r"(a[c]) : "r"(k1), "r"(k2)); }
                if (OP == OP_HMNMX2)    { asm volatile("max.f16x2 %0,%0,%1;" : "+r"(a[c]) : "r"(b[c])); }
                if (OP == OP_PRMT)      { asm volatile("prmt.b32 %0,%0,%1,%2;" : "+r"(a[c]) : "r"(b[c]), "r"(k2)); }
                if (OP == OP_LOP3)      { asm volatile("lop3.b32 %0,%0,%1,%2,0x96;" : "+r"(a[c]) : "r"(b[c]), "r"(k1)); }
                if (OP == OP_MIX_IMAD_VIMNMX3)  { asm volatile("mad.lo.u32 %0,%0,%1,%2;" : "+r"(b[c]) : "r"(one), "r"(k1)); a[c] = __vimax3_u16x2(a[c], k2, k1); asm volatile("" : "+r"(a[c])); }
                if (OP == OP_MIX_IMAD_VIADDMNMX){ asm volatile("mad.lo.u32 %0,%0,%1,%2;" : "+r"(b[c]) : "r"(one), "r"(k1)); a[c] = __viaddmax_u16x2(a[c], k1, k2); asm volatile("" : "+r"(a[c])); }
                if (OP == OP_MIX_HFMA2_VIMNMX3) { asm volatile("fma.rn.f16x2 %0,%0,%1,%2;" : "+r"(b[c]) : "r"(k1), "r"(k2)); a[c] = __vimax3_u16x2(a[c], k2, k1); asm volatile("" : "+r"(a[c])); }
                if (OP == OP_MIX_IADD_VIMNMX3)  { asm volatile("add.u32 %0,%0,%1;" : "+r"(b[c]) : "r"(k1)); a[c] = __vimax3_u16x2(a[c], k2, k1); asm volatile("" : "+r"(a[c])); }
                if (OP == OP_SHFL)      { a[c] = __shfl_up_sync(0xffffffffu, a[c], 1); }
                if (OP == OP_LDS)       { unsigned v; asm volatile("ld.shared.u32 %0,[%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(&sm[lidx + (c * UNR + u) * 32 + (it & 1) * 1024]))); a[c] ^= v; }
                if (OP == OP_MIX_LDS_VIMNMX3)   { unsigned v; asm volatile("ld.shared.u32 %0,[%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(&sm[lidx + (c * UNR + u) * 32 + (it & 1) * 1024]))); a[c] = __vimax3_u16x2(a[c], v, k1); }
                if (OP == OP_MIX3) {
                    // 4 IMAD + 4 VIMNMX3 + 1 LDS per unit: the planned SW cell-word instruction mix
                    unsigned v; asm volatile("ld.shared.u32 %0,[%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(&sm[lidx + (c * UNR + u) * 32 + (it & 1) * 1024])));
                    unsigned t, hq, er, fr;
                    asm volatile("mad.lo.u32 %0,%1,%2,%3;" : "=r"(t)  : "r"(a[c]), "r"(one), "r"(v));
                    unsigned h = __vimax3_u16x2(t, b[c], k2);
                    asm volatile("mad.lo.u32 %0,%1,%2,%3;" : "=r"(hq) : "r"(h), "r"(one), "r"(k1));
                    asm volatile("mad.lo.u32 %0,%1,%2,%3;" : "=r"(er) : "r"(b[c]), "r"(one), "r"(k1));
                    asm volatile("mad.lo.u32 %0,%1,%2,%3;" : "=r"(fr) : "r"(k2), "r"(one), "r"(k1));
                    b[c] = __vimax3_u16x2(er, hq, k1);
                    k2 = __vimax3_u16x2(fr, hq, k1);
                    a[c] = __vimax3_u16x2(a[c], h, k1);
                }
            }
        }
    }
    long long t1 = clock64();
    unsigned r = k2;
    #pragma unroll
    for (int c = 0; c < NCH; c++) r ^= a[c] ^ b[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) { clk[2 * blockIdx.x] = t0; clk[2 * blockIdx.x + 1] = t1; }
}

template<int OP> void run(int sms, int blocks_per_sm, unsigned* d_out, long long* d_clk, long long* h_clk) {
    int grid = sms * blocks_per_sm;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    bench<OP><<<grid, 256>>>(d_out, 12345u, 1u, d_clk);  // warm-up
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    bench<OP><<<grid, 256>>>(d_out, 12345u, 1u, d_clk);
    CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    CK(cudaMemcpy(h_clk, d_clk, sizeof(long long) * 2 * grid, cudaMemcpyDeviceToHost));
    double avg = 0; for (int b = 0; b < grid; b++) avg += double(h_clk[2*b+1] - h_clk[2*b]); avg /= grid;
    double lane_ops_per_block = 256.0 * NCH * UNR * ITER * ops_per_unit[OP];
    double per_clk_sm = lane_ops_per_block * blocks_per_sm / avg;
    double total = lane_ops_per_block * grid;
    double mhz = avg / (ms * 1e3);
    printf("{\"op\": \"%s\", \"blocks_per_sm\": %d, \"lane_ops_per_clk_per_sm\": %.2f, \"ms\": %.4f, \"tera_lane_ops_per_s\": %.3f, \"avg_block_clk\": %.0f, \"implied_sm_mhz\": %.0f}\n",
           names[OP], blocks_per_sm, per_clk_sm, ms, total / (ms * 1e-3) / 1e12, avg, mhz);
    fflush(stdout);
}

template<int OP> void run_all(int sms, unsigned* d_out, long long* d_clk, long long* h_clk) {
    run<OP>(sms, 2, d_out, d_clk, h_clk);
    run<OP>(sms, 4, d_out, d_clk, h_clk);
    run<OP>(sms, 8, d_out, d_clk, h_clk);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, sms, p.clockRate);
    unsigned* d_out; long long *d_clk, *h_clk;
    CK(cudaMalloc(&d_out, sizeof(unsigned) * sms * 8 * 256));
    CK(cudaMalloc(&d_clk, sizeof(long long) * 2 * sms * 8));
    h_clk = (long long*)malloc(sizeof(long long) * 2 * sms * 8);
    run_all<OP_IADD>(sms, d_out, d_clk, h_clk);
    run_all<OP_IMAD>(sms, d_out, d_clk, h_clk);
    run_all<OP_VIMNMX3>(sms, d_out, d_clk, h_clk);
    run_all<OP_VIADDMNMX>(sms, d_out, d_clk, h_clk);
    run_all<OP_VIMNMX2>(sms, d_out, d_clk, h_clk);
    run_all<OP_VIADD2>(sms, d_out, d_clk, h_clk);
    run_all<OP_HFMA2>(sms, d_out, d_clk, h_clk);
    run_all<OP_HMNMX2>(sms, d_out, d_clk, h_clk);
    run_all<OP_PRMT>(sms, d_out, d_clk, h_clk);
    run_all<OP_LOP3>(sms, d_out, d_clk, h_clk);
    run_all<OP_MIX_IMAD_VIMNMX3>(sms, d_out, d_clk, h_clk);
    run_all<OP_MIX_IMAD_VIADDMNMX>(sms, d_out, d_clk, h_clk);
    run_all<OP_MIX_HFMA2_VIMNMX3>(sms, d_out, d_clk, h_clk);
    run_all<OP_MIX_IADD_VIMNMX3>(sms, d_out, d_clk, h_clk);
    run_all<OP_SHFL>(sms, d_out, d_clk, h_clk);
    run_all<OP_LDS>(sms, d_out, d_clk, h_clk);
    run_all<OP_MIX_LDS_VIMNMX3>(sms, d_out, d_clk, h_clk);
    run_all<OP_MIX3>(sms, d_out, d_clk, h_clk);
    return 0;
}
