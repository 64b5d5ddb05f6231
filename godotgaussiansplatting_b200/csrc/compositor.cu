// compositor.cu -- stage 4: per-tile front-to-back alpha blend.  Replaces gsplat_render.glsl:50-111.
//
// One CTA per 16x16 tile, one thread per pixel (the reference's workgroup shape), 256-splat chunks staged
// in shared memory.  What changes is the data movement, not the arithmetic:
//   * the gather `culled_buffer[sort_buffer[...]]` (:72) for chunk i+1 is issued into registers BEFORE the
//     blend loop of chunk i runs (software double buffering), so the random 48-B record gathers overlap
//     the ~256*35 FP32 instructions of the blend instead of sitting between two barriers;
//   * only the 9 floats the blend needs (image_pos, conic, colour+opacity) are staged (36 B, not 48 B);
//   * the tile-stop vote `atomicAdd(shared_t, uint(t*255))` (:97) becomes a warp reduction + 8 shared
//     words (same sum, one barrier less per chunk);
//   * the per-pixel early-out `t > 1/255` (:79) additionally breaks the warp's chunk loop when no lane is
//     live, which is a pure skip of no-op iterations.
// Arithmetic: "gsr deterministic math" (common.cuh; compiled -fmad=false): the two GLSL-legal contractions
// of :84 and the three of :89 are explicit __fmaf_rn, exp() is det_exp().  Bit-identical to the oracle.
#include "common.cuh"

namespace gsr {

namespace {

constexpr int CHUNK = 256;  // gsplat_render.glsl:9 WORKGROUP_SIZE
constexpr float MIN_ALPHA = 1.0f / 255.0f;

struct Staged {  // registers holding one gathered record
    float4 a;    // image_pos.xy, conic.x, conic.y
    float4 b;    // conic.z, color.rgb
    float o;     // opacity
};

__device__ __forceinline__ Staged gather(const float4 *__restrict__ records, const uint32_t *__restrict__ values, uint32_t idx) {
    const uint32_t v = __ldg(values + idx);
    const float4 *r = records + (uint64_t)v * 3u;
    const float4 r0 = __ldg(r + 0), r1 = __ldg(r + 1), r2 = __ldg(r + 2);
    Staged s;
    s.a = make_float4(r0.x, r0.y, r1.x, r1.y);
    s.b = make_float4(r1.z, r2.x, r2.y, r2.z);
    s.o = r2.w;
    return s;
}

__global__ void __launch_bounds__(CHUNK) composite_kernel(const __grid_constant__ CompositeArgs p) {
    __shared__ float4 s_a[2][CHUNK];
    __shared__ float4 s_b[2][CHUNK];
    __shared__ float s_o[2][CHUNK];
    __shared__ uint32_t s_vote[CHUNK / 32];

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t tile_id = (uint32_t)p.tile_begin + blockIdx.x;
    const uint32_t tx = tile_id % (uint32_t)p.tiles_x, ty = tile_id / (uint32_t)p.tiles_x;
    const int px = (int)(tx * TILE + (tid & 15u)), py = (int)(ty * TILE + (tid >> 4));
    const float fpx = (float)px, fpy = (float)py;

    const uint2 bounds = p.bounds[tile_id];
    const int32_t diff = (int32_t)(bounds.y - bounds.x);
    const int num_splats = diff > 0 ? diff : 0;                             // :61
    const int num_iterations = (int)ceilf((float)num_splats / (float)CHUNK);  // :62

    float cr = 0.0f, cg = 0.0f, cb = 0.0f, t = 1.0f;
    uint32_t staged = 0;  // SURVEY 8 symbol C: sum of consumed chunk sizes (uniform across the CTA)

    Staged nxt;
    nxt.a = make_float4(0.f, 0.f, 0.f, 0.f); nxt.b = nxt.a; nxt.o = 0.f;
    if (num_iterations > 0 && (int)tid < num_splats) nxt = gather(p.records, p.values, bounds.x + tid);

    for (int i = 0; i < num_iterations; ++i) {
        const int buf = i & 1;
        const int sort_offset = CHUNK * i;
        const int chunk = (num_splats - sort_offset) < CHUNK ? (num_splats - sort_offset) : CHUNK;
        staged += (uint32_t)chunk;
        s_a[buf][tid] = nxt.a;
        s_b[buf][tid] = nxt.b;
        s_o[buf][tid] = nxt.o;
        __syncthreads();
        // prefetch the next chunk's record while this one is blended
        if (i + 1 < num_iterations && sort_offset + CHUNK + (int)tid < num_splats)
            nxt = gather(p.records, p.values, bounds.x + (uint32_t)(sort_offset + CHUNK) + tid);

        // :79-91
        for (int j = 0; j < chunk; ++j) {
            const bool live = t > MIN_ALPHA;
            if (!__any_sync(0xffffffffu, live)) break;
            if (live) {
                const float4 a = s_a[buf][j];
                const float4 b = s_b[buf][j];
                const float op = s_o[buf][j];
                const float ox = a.x - fpx, oy = a.y - fpy;
                // power = -0.5*(cx*ox*ox + cz*oy*oy) - cy*ox*oy
                const float q = __fmaf_rn(b.x * oy, oy, a.z * ox * ox);
                const float power = __fmaf_rn(-(a.w * ox), oy, -0.5f * q);
                const float alpha = op * det_exp(power);
                cr = __fmaf_rn(b.y * alpha, t, cr);
                cg = __fmaf_rn(b.z * alpha, t, cg);
                cb = __fmaf_rn(b.w * alpha, t, cb);
                t = t * (1.0f - alpha);
            }
        }

        // :97 tile-stop vote: continue only if sum over the 256 threads of uint(t*255) > 255
        const uint32_t wsum = __reduce_add_sync(0xffffffffu, (uint32_t)(t * 255.0f));
        if (lane == 0) s_vote[warp] = wsum;
        __syncthreads();
        uint32_t shared_t = 0;
#pragma unroll
        for (int w = 0; w < CHUNK / 32; ++w) shared_t += s_vote[w];
        if (!(shared_t > 255u)) break;
    }

    if (tid == 0 && staged && p.frame) atomicAdd(&p.frame->staged, (unsigned long long)staged);

    // :100-101
    const float hx = (float)num_splats * 5e-4f;
    const float h0 = 0.0f * (1.0f - hx) + 1.0f * hx, h1 = 0.0f * (1.0f - hx) + 0.2f * hx, h2 = 1.0f * (1.0f - hx) + 0.2f * hx;
    const float k = 1.0f - t;
    if (px < p.width && py < p.height) {
        float4 o;
        o.x = cr + h0 * k * p.heatmap_factor;
        o.y = cg + h1 * k * p.heatmap_factor;
        o.z = cb + h2 * k * p.heatmap_factor;
        o.w = 1.0f;
        p.out[(uint64_t)py * (uint64_t)p.width + (uint64_t)px] = o;
    }

    // :105-110 pick: the elected (first) lane of each 32-wide subgroup of the target tile
    if (lane == 0 && tile_id == p.target_tile_id && t != 1.0f) {
        const uint32_t v = p.values[bounds.x + (bounds.y - bounds.x) / 10u];
        const float4 r0 = p.records[(uint64_t)v * 3u + 0], r1 = p.records[(uint64_t)v * 3u + 1];
        *p.pick = make_float4(r0.z, r0.w, r1.w, (float)num_splats);
    }
}

}  // namespace

int launch_composite(const CompositeArgs &a, cudaStream_t stream) {
    if (a.num_tiles <= 0) return GSR_OK;
    composite_kernel<<<a.num_tiles, CHUNK, 0, stream>>>(a);
    GSR_CUDA_TRY(cudaGetLastError());
    return GSR_OK;
}

}  // namespace gsr
