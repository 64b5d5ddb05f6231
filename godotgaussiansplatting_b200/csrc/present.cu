// present.cu -- presentation hand-off (scope row f3): converts the RGBA32F frame (render_texture, rasterizer.gd:92) into what
// the consumer wants before it crosses PCIe / lands in an imported external image:
//   * GSR_OUT_RGB32F   -- alpha of the reference's output is the constant 1.0 (gsplat_render.glsl:101): packed away, -25 % bytes;
//   * GSR_OUT_RGBA16F  -- IEEE binary16, round-to-nearest-even per channel, 8 B/pixel;
//   * GSR_OUT_RGBA8    -- UNORM8: rint(clamp(x, 0, 1) * 255), 4 B/pixel;
//   * | GSR_OUT_SRGB_TO_LINEAR -- the conversion the reference's presentation shader applies when it samples the texture
//     (resources/shaders/spatial/main.gdshader:7-11,18), fused here so that the host / the blit does not have to:
//         higher = pow((x + 0.055) / 1.055, 2.4); lower = x / 12.92; x < 0.04045 ? lower : higher        (rgb; alpha untouched)
//     pow() is det_pow() of the arithmetic contract (common.cuh), so the oracle reproduces the result bit for bit.
#include <cuda_fp16.h>

#include "common.cuh"

namespace gsr {

namespace {

__device__ __forceinline__ float srgb_to_linear(float x) {   // main.gdshader:7-11
    const float higher = det_pow(__fdiv_rn(__fadd_rn(x, 0.055f), 1.055f), 2.4f);
    const float lower = __fdiv_rn(x, 12.92f);
    return (x < 0.04045f) ? lower : higher;
}

__device__ __forceinline__ uint32_t unorm8(float x) {
    const float c = g_clamp(x, 0.0f, 1.0f);          // NaN -> comparisons false -> x itself; (uint) of NaN saturates to 0
    return __float2uint_rn(__fmul_rn(c, 255.0f));
}

template <int FMT, bool LINEARIZE>
__global__ void __launch_bounds__(256) present_kernel(const float4 *__restrict__ rgba, void *__restrict__ out, uint64_t pixels) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (FMT == GSR_OUT_RGB32F) {   // thread = 4 pixels: four float4 loads, three float4 stores
        const uint64_t p0 = 4 * i;
        if (p0 >= pixels) return;
        float v[12];
        const uint64_t np = pixels - p0 < 4 ? pixels - p0 : 4;
        for (uint64_t k = 0; k < np; ++k) {
            const float4 a = rgba[p0 + k];
            v[3 * k + 0] = LINEARIZE ? srgb_to_linear(a.x) : a.x;
            v[3 * k + 1] = LINEARIZE ? srgb_to_linear(a.y) : a.y;
            v[3 * k + 2] = LINEARIZE ? srgb_to_linear(a.z) : a.z;
        }
        if (np == 4) {
            float4 *o = reinterpret_cast<float4 *>(out) + 3 * i;
            o[0] = make_float4(v[0], v[1], v[2], v[3]);
            o[1] = make_float4(v[4], v[5], v[6], v[7]);
            o[2] = make_float4(v[8], v[9], v[10], v[11]);
        } else {
            float *o = reinterpret_cast<float *>(out) + 3 * p0;
            for (uint64_t k = 0; k < 3 * np; ++k) o[k] = v[k];
        }
        return;
    }
    if (i >= pixels) return;
    float4 a = rgba[i];
    if (LINEARIZE) { a.x = srgb_to_linear(a.x); a.y = srgb_to_linear(a.y); a.z = srgb_to_linear(a.z); }
    if (FMT == GSR_OUT_RGBA32F) {
        reinterpret_cast<float4 *>(out)[i] = a;
    } else if (FMT == GSR_OUT_RGBA16F) {
        const __half2 lo = __floats2half2_rn(a.x, a.y), hi = __floats2half2_rn(a.z, a.w);
        uint2 w;
        w.x = *reinterpret_cast<const uint32_t *>(&lo);
        w.y = *reinterpret_cast<const uint32_t *>(&hi);
        reinterpret_cast<uint2 *>(out)[i] = w;
    } else {  // GSR_OUT_RGBA8
        reinterpret_cast<uint32_t *>(out)[i] = unorm8(a.x) | (unorm8(a.y) << 8) | (unorm8(a.z) << 16) | (unorm8(a.w) << 24);
    }
}

}  // namespace

#ifndef GSR_CPU_EMU  // host side: CUDA only (tests/kernel_emu drives the kernel above itself)
// Force-load this file's kernels (CUDA loads modules lazily; a first launch that has to load code while another context's
// kernel spins on a flag this launch would satisfy can stall the host: see gsr_group_attach).
int preload_present_kernels() {
    cudaFuncAttributes fa;
    GSR_CUDA_TRY(cudaFuncGetAttributes(&fa, present_kernel<GSR_OUT_RGBA32F, true>));
    GSR_CUDA_TRY(cudaFuncGetAttributes(&fa, present_kernel<GSR_OUT_RGB32F, false>));
    GSR_CUDA_TRY(cudaFuncGetAttributes(&fa, present_kernel<GSR_OUT_RGB32F, true>));
    GSR_CUDA_TRY(cudaFuncGetAttributes(&fa, present_kernel<GSR_OUT_RGBA16F, false>));
    GSR_CUDA_TRY(cudaFuncGetAttributes(&fa, present_kernel<GSR_OUT_RGBA16F, true>));
    GSR_CUDA_TRY(cudaFuncGetAttributes(&fa, present_kernel<GSR_OUT_RGBA8, false>));
    GSR_CUDA_TRY(cudaFuncGetAttributes(&fa, present_kernel<GSR_OUT_RGBA8, true>));
    return GSR_OK;
}
size_t present_bytes_per_pixel(int format) {
    switch (format & 0xFF) {
        case GSR_OUT_RGBA32F: return 16;
        case GSR_OUT_RGB32F: return 12;
        case GSR_OUT_RGBA16F: return 8;
        case GSR_OUT_RGBA8: return 4;
        default: return 0;
    }
}

int launch_present(const float4 *rgba, void *out, uint64_t pixels, int format, cudaStream_t stream) {
    if (!pixels) return GSR_OK;
    const bool lin = (format & GSR_OUT_SRGB_TO_LINEAR) != 0;
    const int fmt = format & 0xFF;
    const uint64_t items = fmt == GSR_OUT_RGB32F ? (pixels + 3) / 4 : pixels;
    const uint32_t grid = (uint32_t)((items + 255) / 256);
#define GSR_PRESENT(F)                                                                             \
    do {                                                                                           \
        if (lin) present_kernel<F, true><<<grid, 256, 0, stream>>>(rgba, out, pixels);             \
        else present_kernel<F, false><<<grid, 256, 0, stream>>>(rgba, out, pixels);                \
    } while (0)
    switch (fmt) {
        case GSR_OUT_RGBA32F: GSR_PRESENT(GSR_OUT_RGBA32F); break;
        case GSR_OUT_RGB32F: GSR_PRESENT(GSR_OUT_RGB32F); break;
        case GSR_OUT_RGBA16F: GSR_PRESENT(GSR_OUT_RGBA16F); break;
        case GSR_OUT_RGBA8: GSR_PRESENT(GSR_OUT_RGBA8); break;
        default: set_last_error("unknown output format %d", format); return GSR_ERR_INVALID;
    }
#undef GSR_PRESENT
    GSR_CUDA_TRY(cudaGetLastError());
    return GSR_OK;
}
#endif  // GSR_CPU_EMU

}  // namespace gsr
