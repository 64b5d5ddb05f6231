// ubench/f32x2.cu -- does fma.rn.f32x2 (SASS FFMA2) double FP32 throughput per issue slot on sm_100a?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench/f32x2 ubench/f32x2.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
__global__ void k_scalar(float *out, float a, float b) {
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = __fmaf_rn(x[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_packed(float *out, float a, float b) {
    unsigned long long x[8], aa, bb;
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
#pragma unroll
    for (int i = 0; i < 8; ++i) { float v = threadIdx.x * 1e-3f + i; asm("mov.b64 %0, {%1, %1};" : "=l"(x[i]) : "f"(v)); }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(aa), "l"(bb));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i])); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_mixed(float *out, float a, float b) {  // packed FMA + scalar FMNMX/integer mix like the blend loop
    unsigned long long x[4], aa, bb;
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
    float y[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { float v = threadIdx.x * 1e-3f + i; asm("mov.b64 %0, {%1, %1};" : "=l"(x[i]) : "f"(v)); y[i] = v; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(aa), "l"(bb));
            asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(x[i]) : "l"(aa));
            asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x[i]) : "l"(bb));
            y[i] = fmaxf(y[i] * a, -127.0f);
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i])); s += lo + hi + y[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 8, threads = 256;
    float *out; cudaMalloc(&out, sizeof(float) * blocks * threads);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
        float ms;
        cudaEventRecord(e0); k_scalar<<<blocks, threads>>>(out, 0.999f, 0.001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double fma = (double)blocks * threads * ITERS * 8;
        printf("scalar FFMA : %.3f ms  %.1f Gfma/s  (%.2f warp-inst/clk/SM @1.965GHz)\n", ms, fma / ms / 1e6, fma / 32 / (ms * 1e-3) / sms / 1.965e9);
        cudaEventRecord(e0); k_packed<<<blocks, threads>>>(out, 0.999f, 0.001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("packed FFMA2: %.3f ms  %.1f Gfma/s  (%.2f warp-inst/clk/SM)\n", ms, 2 * fma / ms / 1e6, fma / 32 / (ms * 1e-3) / sms / 1.965e9);
        cudaEventRecord(e0); k_mixed<<<blocks, threads>>>(out, 0.999f, 0.001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double inst = (double)blocks * threads * ITERS * 4 * 5;  // 3 packed + FMUL + FMNMX per i
        printf("mixed (3 x2-ops + FMUL + FMNMX): %.3f ms  %.2f warp-inst/clk/SM\n", ms, inst / 32 / (ms * 1e-3) / sms / 1.965e9);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
