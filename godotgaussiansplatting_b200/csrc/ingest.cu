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

// RGBA32F -> RGB32F packing for the host read-back: alpha is the constant 1.0 (gsplat_render.glsl:101), so it does not
// have to cross PCIe.  Thread i converts pixels 4i..4i+3: four float4 loads, three float4 stores (both contiguous).
__global__ void __launch_bounds__(256) pack_rgb_kernel(const float4 *__restrict__ rgba, float4 *__restrict__ rgb, uint64_t quads, uint64_t pixels) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= quads) return;
    const uint64_t p0 = 4 * i;
    if (p0 + 4 <= pixels) {
        const float4 a = rgba[p0], b = rgba[p0 + 1], c = rgba[p0 + 2], d = rgba[p0 + 3];
        rgb[3 * i + 0] = make_float4(a.x, a.y, a.z, b.x);
        rgb[3 * i + 1] = make_float4(b.y, b.z, c.x, c.y);
        rgb[3 * i + 2] = make_float4(c.z, d.x, d.y, d.z);
    } else {  // ragged tail (pixel count not a multiple of 4)
        float *o = reinterpret_cast<float *>(rgb) + 3 * p0;
        for (uint64_t p = p0; p < pixels; ++p) { const float4 a = rgba[p]; *o++ = a.x; *o++ = a.y; *o++ = a.z; }
    }
}

}  // namespace

int launch_pack_rgb(const float4 *rgba, float4 *rgb, uint64_t pixels, cudaStream_t stream) {
    const uint64_t quads = (pixels + 3) / 4;
    if (!quads) return GSR_OK;
    pack_rgb_kernel<<<(uint32_t)((quads + 255) / 256), 256, 0, stream>>>(rgba, rgb, quads, pixels);
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

}  // namespace gsr
