// ingest.cu -- AoS -> SoA transposition of uploaded splats.
//
// The reference keeps the 60-float std430 `Splat` struct in one AoS storage buffer
// (gsplat_projection.glsl:33-40, written by util/ply_file.gd:71).  libgsr accepts exactly that struct at the
// boundary (gsr_upload_splats_aos) and stores it as 15 float4 planes so that the projection kernel issues
// coalesced 128-bit loads and culled splats touch one plane only.
#include "common.cuh"

namespace gsr {

namespace {

constexpr int SPLATS_PER_BLOCK = 128;

__global__ void __launch_bounds__(256) aos_to_soa_kernel(const float4 *__restrict__ aos, uint64_t count, float4 *__restrict__ soa,
                                                         uint64_t plane_stride, uint64_t first) {
    __shared__ float4 s[SPLATS_PER_BLOCK * NUM_PLANES];
    const uint64_t s0 = (uint64_t)blockIdx.x * SPLATS_PER_BLOCK;
    const uint32_t here = (uint32_t)((count - s0) < (uint64_t)SPLATS_PER_BLOCK ? (count - s0) : SPLATS_PER_BLOCK);
    const float4 *src = aos + s0 * NUM_PLANES;
    for (uint32_t i = threadIdx.x; i < here * NUM_PLANES; i += blockDim.x) s[i] = src[i];  // coalesced AoS read
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < here * NUM_PLANES; i += blockDim.x) {
        const uint32_t plane = i / here, k = i - plane * here;  // consecutive threads -> consecutive splats of one plane
        soa[(uint64_t)plane * plane_stride + first + s0 + k] = s[k * NUM_PLANES + plane];
    }
}

// ---- scope row f1: PLY vertex -> Splat on the device (util/ply_file.gd:44-69) -------------------------------------
// One thread per vertex; a CTA stages its 128 vertices (nprops floats each, standard 3DGS property order) through
// shared memory so that the AoS read is coalesced, then every thread evaluates exp(scale) and the sigmoid in float64
// (GDScript floats are doubles; the results are narrowed when they enter Vector3 / PackedFloat32Array), builds
// Basis(Quaternion).transposed(), Sigma = (S R)^T (S R) with Godot's Basis*Basis operation order (including the
// structurally-zero products, so signed zeros match), re-interleaves the SH coefficients and writes the 15 SoA planes.
constexpr int INGEST_SPLATS = 128;

__device__ __forceinline__ void godot_basis_mul(const float a[3][3], const float b[3][3], float o[3][3]) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) o[i][j] = (b[0][j] * a[i][0] + b[1][j] * a[i][1]) + b[2][j] * a[i][2];
}

__global__ void __launch_bounds__(INGEST_SPLATS) ply_to_soa_kernel(const float *__restrict__ ply, uint32_t nprops, uint64_t count, float creation_time,
                                                                  float4 *__restrict__ soa, uint64_t plane_stride, uint64_t first) {
#ifndef GSR_CPU_EMU
    extern __shared__ float s_v[];  // [INGEST_SPLATS][nprops]
#else  // tests/kernel_emu (CPU logic pre-flight): nprops <= 256 (gsr_upload_ply_raw rejects more)
    __shared__ float s_v[INGEST_SPLATS * 256];
#endif
    const uint64_t s0 = (uint64_t)blockIdx.x * INGEST_SPLATS;
    const uint32_t here = (uint32_t)((count - s0) < (uint64_t)INGEST_SPLATS ? (count - s0) : INGEST_SPLATS);
    const float *src = ply + s0 * nprops;
    for (uint32_t i = threadIdx.x; i < here * nprops; i += blockDim.x) s_v[i] = src[i];
    __syncthreads();
    if (threadIdx.x >= here) return;
    const float *p = s_v + (size_t)threadIdx.x * nprops;
    const uint64_t id = first + s0 + threadIdx.x;

    const float sc0 = (float)exp((double)p[55]), sc1 = (float)exp((double)p[56]), sc2 = (float)exp((double)p[57]);
    const float qx = p[59], qy = p[60], qz = p[61], qw = p[58];  // Quaternion(rot_1, rot_2, rot_3, rot_0)
    const float d = ((qx * qx + qy * qy) + qz * qz) + qw * qw;
    const float s = 2.0f / d;
    const float xs = qx * s, ys = qy * s, zs = qz * s;
    const float wx = qw * xs, wy = qw * ys, wz = qw * zs;
    const float xx = qx * xs, xy = qx * ys, xz = qx * zs;
    const float yy = qy * ys, yz = qy * zs, zz = qz * zs;
    const float Bq[3][3] = {{1.0f - (yy + zz), xy - wz, xz + wy}, {xy + wz, 1.0f - (xx + zz), yz - wx}, {xz - wy, yz + wx, 1.0f - (xx + yy)}};
    float R[3][3], M[3][3], Mt[3][3], Cv[3][3];
    const float S[3][3] = {{sc0, 0.0f, 0.0f}, {0.0f, sc1, 0.0f}, {0.0f, 0.0f, sc2}};
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) R[r][c] = Bq[c][r];  // .transposed()
    godot_basis_mul(S, R, M);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) Mt[r][c] = M[c][r];
    godot_basis_mul(Mt, M, Cv);
    const float opacity = (float)(1.0 / (1.0 + exp(-(double)p[54])));

    soa[0 * plane_stride + id] = make_float4(p[0], p[1], p[2], creation_time);
    soa[1 * plane_stride + id] = make_float4(Cv[0][0], Cv[0][1], Cv[0][2], Cv[1][1]);
    soa[2 * plane_stride + id] = make_float4(Cv[1][2], Cv[2][2], opacity, 0.0f);
    float sh[48];  // coefficient-major RGB: DC, then f_rest R 0..14 | G 15..29 | B 30..44 re-interleaved (:65-69)
    sh[0] = p[6]; sh[1] = p[7]; sh[2] = p[8];
#pragma unroll
    for (int k = 0; k < 15; ++k) { sh[3 + 3 * k + 0] = p[9 + k]; sh[3 + 3 * k + 1] = p[9 + 15 + k]; sh[3 + 3 * k + 2] = p[9 + 30 + k]; }
#pragma unroll
    for (int k = 0; k < 12; ++k) soa[(uint64_t)(3 + k) * plane_stride + id] = make_float4(sh[4 * k], sh[4 * k + 1], sh[4 * k + 2], sh[4 * k + 3]);
}

}  // namespace

#ifndef GSR_CPU_EMU  // host side: CUDA only
// Force-load this file's kernels (CUDA loads modules lazily; a first launch that has to load code while another context's
// kernel spins on a flag this launch would satisfy can stall the host: see gsr_group_attach).
int preload_ingest_kernels() {
    cudaFuncAttributes fa;
    GSR_CUDA_TRY(cudaFuncGetAttributes(&fa, ply_to_soa_kernel));
    GSR_CUDA_TRY(cudaFuncGetAttributes(&fa, aos_to_soa_kernel));
    return GSR_OK;
}
int launch_ply_to_soa(const float *ply, uint32_t nprops, uint64_t count, float creation_time, float4 *soa, uint64_t plane_stride, uint64_t first,
                      cudaStream_t stream) {
    if (count == 0) return GSR_OK;
    const size_t smem = sizeof(float) * (size_t)INGEST_SPLATS * nprops;
    if (smem > 48 * 1024) GSR_CUDA_TRY(cudaFuncSetAttribute(ply_to_soa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint32_t blocks = (uint32_t)((count + INGEST_SPLATS - 1) / INGEST_SPLATS);
    ply_to_soa_kernel<<<blocks, INGEST_SPLATS, smem, stream>>>(ply, nprops, count, creation_time, soa, plane_stride, first);
    GSR_CUDA_TRY(cudaGetLastError());
    return GSR_OK;
}

int launch_aos_to_soa(const float4 *aos, uint64_t count, float4 *soa, uint64_t plane_stride, uint64_t first, cudaStream_t stream) {
    if (count == 0) return GSR_OK;
    const uint32_t blocks = (uint32_t)((count + SPLATS_PER_BLOCK - 1) / SPLATS_PER_BLOCK);
    aos_to_soa_kernel<<<blocks, 256, 0, stream>>>(aos, count, soa, plane_stride, first);
    GSR_CUDA_TRY(cudaGetLastError());
    return GSR_OK;
}
#endif  // GSR_CPU_EMU

}  // namespace gsr
