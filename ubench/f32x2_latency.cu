// ubench/f32x2_latency.cu -- how fast can ONE warp (and W warps per SM sub-partition) issue packed fp32x2 FMAs?
// Times K independent dependent-chains of fma.rn.f32x2 (and scalar fma) with clock64 on one SM:
//   cycles per instruction per warp  ->  dependent-issue latency (K = 1) and the per-warp issue ceiling (K large),
//   for W = 1, 2, 4 warps per SMSP   ->  whether the FMA pipe needs several warps to reach one packed op per 2 cycles.
// This decides how the compositor's blend loop must be scheduled for a tile that runs alone on its SM (kernel tail,
// sparse multi-GPU shards).  build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench/f32x2_latency ubench/f32x2_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
typedef unsigned long long u64;

template <int K, bool PACKED>
__global__ void chain(long long *cycles, float *sink, float a, float b) {
    u64 x[K], aa, bb;
    float y[K];
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
#pragma unroll
    for (int i = 0; i < K; ++i) { float v = threadIdx.x * 1e-3f + i; asm("mov.b64 %0, {%1, %1};" : "=l"(x[i]) : "f"(v)); y[i] = v; }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < K; ++i) {
                if (PACKED) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(aa), "l"(bb));
                else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(y[i]) : "f"(a), "f"(b));
            }
        }
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < K; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i])); s += lo + hi + y[i]; }
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
}

template <int K, bool PACKED>
void run(int warps_per_smsp, long long *d_cyc, float *d_sink) {
    chain<K, PACKED><<<1, 128 * warps_per_smsp>>>(d_cyc, d_sink, 0.999f, 0.001f);
    chain<K, PACKED><<<1, 128 * warps_per_smsp>>>(d_cyc, d_sink, 0.999f, 0.001f);
    long long c;
    cudaMemcpy(&c, d_cyc, sizeof c, cudaMemcpyDeviceToHost);
    const double per_warp_inst = (double)ITERS * 4 * K;
    printf("%s K=%d chains, %d warp(s)/SMSP: %.2f cycles per instruction per warp, %.3f warp-instr/cycle/SMSP\n", PACKED ? "FFMA2" : "FFMA ", K,
           warps_per_smsp, c / per_warp_inst, per_warp_inst * warps_per_smsp / c);
}

int main() {
    long long *d_cyc; float *d_sink;
    cudaMalloc(&d_cyc, 8); cudaMalloc(&d_sink, 4 * 4096);
    for (int w = 1; w <= 4; w *= 2) {
        run<1, true>(w, d_cyc, d_sink); run<2, true>(w, d_cyc, d_sink); run<3, true>(w, d_cyc, d_sink); run<4, true>(w, d_cyc, d_sink);
        run<6, true>(w, d_cyc, d_sink); run<8, true>(w, d_cyc, d_sink);
        run<1, false>(w, d_cyc, d_sink); run<2, false>(w, d_cyc, d_sink); run<4, false>(w, d_cyc, d_sink); run<8, false>(w, d_cyc, d_sink);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
