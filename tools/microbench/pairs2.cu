// Follow-up to pairs.cu: which pipe does VHMNMX (three-input packed fp16 max, what ptxas makes of two chained max.f16x2)
// issue on?  1:1 mixes with the DPX ops and the FMA-side adds; combined tera-lane-ops/s, 18.6 = same pipe, 37 = different pipes.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){fprintf(stderr,"CUDA %s @%d: %s\n",#x,__LINE__,cudaGetErrorString(e)); exit(1);} }while(0)
constexpr int NCH = 6, ITER = 8192, UNR = 4;
enum { VHMNMX, VIMNMX3, VIADDMNMX, VIADD2, IMAD, NOPS };
static const char* nm[] = {"VHMNMX","VIMNMX3","VIADDMNMX","VIADD.16x2","IMAD"};
template<int OP> __device__ __forceinline__ void op(unsigned& a, unsigned b, unsigned k1, unsigned k2, unsigned one) {
    if (OP == VHMNMX)    { asm volatile("max.f16x2 %0,%0,%1;" : "+r"(a) : "r"(b)); asm volatile("max.f16x2 %0,%0,%1;" : "+r"(a) : "r"(k1)); }
    if (OP == VIMNMX3)   { a = __vimax3_u16x2(a, b, k1); asm volatile("" : "+r"(a)); }
    if (OP == VIADDMNMX) { a = __viaddmax_u16x2(a, k1, b); asm volatile("" : "+r"(a)); }
    if (OP == VIADD2)    { a = __vadd2(a, b); asm volatile("" : "+r"(a)); }
    if (OP == IMAD)      asm volatile("mad.lo.u32 %0,%0,%1,%2;" : "+r"(a) : "r"(one), "r"(b));
}
template<int A, int B>
__global__ void __launch_bounds__(256) bench(unsigned* out, unsigned seed, unsigned one) {
    unsigned a[NCH], b[NCH], x[NCH];
    #pragma unroll
    for (int c = 0; c < NCH; c++) { a[c] = seed * (threadIdx.x + 1 + c); b[c] = seed ^ (c * 77 + threadIdx.x); x[c] = seed + c * 13 + threadIdx.x * 7; }
    unsigned k1 = seed | 0x00010001u, k2 = (seed * 3u) & 0x7777;
    for (int it = 0; it < ITER; it++) {
        #pragma unroll
        for (int u = 0; u < UNR; u++) {
            #pragma unroll
            for (int c = 0; c < NCH; c++) { op<A>(a[c], x[c], k1, k2, one); op<B>(b[c], x[c], k1, k2, one); }
        }
    }
    unsigned r = 0;
    #pragma unroll
    for (int c = 0; c < NCH; c++) r ^= a[c] ^ b[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
static unsigned* d_out; static int sms;
template<int A, int B> void run() {
    int grid = sms * 4;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 3; w++) bench<A,B><<<grid, 256>>>(d_out, 12345u, 1u);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int w = 0; w < 5; w++) bench<A,B><<<grid, 256>>>(d_out, 12345u, 1u);
    CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 5;
    double total = 2.0 * 256.0 * NCH * UNR * ITER * grid;   // SASS instructions x lanes (a fused VHMNMX counts once)
    printf("{\"a\": \"%s\", \"b\": \"%s\", \"tera_lane_ops_per_s\": %.2f, \"ms\": %.3f}\n", nm[A], nm[B], total / (ms * 1e-3) / 1e12, ms);
    fflush(stdout);
}
template<int B> struct Row { static void go() { run<VHMNMX,B>(); if constexpr (B + 1 < NOPS) Row<B+1>::go(); } };
int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0)); sms = p.multiProcessorCount;
    CK(cudaMalloc(&d_out, sizeof(unsigned) * sms * 4 * 256));
    for (int w = 0; w < 200; w++) bench<IMAD,VIMNMX3><<<sms*4,256>>>(d_out, 1u, 1u);
    CK(cudaDeviceSynchronize());
    Row<0>::go();
    return 0;
}
