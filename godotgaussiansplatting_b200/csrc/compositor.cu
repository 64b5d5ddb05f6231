// compositor.cu -- stage 4: per-tile front-to-back alpha blend.  Replaces gsplat_render.glsl:50-111.
//
// One CTA per 16x16 tile like the reference's workgroup, 256-splat chunks staged in shared memory, the
// same per-pixel arithmetic and the same tile-stop vote.  What is Blackwell-specific is how the blend is
// issued: the kernel is instruction-issue bound (ncu: 82 % issue-active, 2 % DRAM), so
//   * every thread owns TWO horizontally adjacent pixels and the blend runs on packed fp32x2
//     instructions (PTX add/sub/mul/fma.rn.f32x2 -> SASS FADD2/FMUL2/FFMA2, sm_100+).  Each lane of a
//     packed op is an ordinary IEEE binary32 operation, so results stay bit-identical to the oracle while
//     the FP32 work of two pixels costs one issue slot (measured on B200: FFMA2 sustains the full
//     128 lane-FMA/clk/SM at 2 warp-instructions/clk/SM, ubench/f32x2.cu);
//   * per-splat control flow is gone: dead pixels (t <= 1/255, gsplat_render.glsl:79) are masked by
//     selecting alpha = 0 (an exact no-op on colour and transmittance), the warp-level "all dead" test
//     runs once per 4 splats, and the last chunk is padded with null splats (opacity 0);
//   * the conic is pre-scaled at staging time (-0.5*cx, -0.5*cz, -cy: exact power-of-two/sign changes) so
//     the `-0.5 * (...)` multiply of :84 disappears from the inner loop without changing any rounding;
//   * the gather `culled_buffer[sort_buffer[...]]` (:72) for chunk i+1 is issued into registers before the
//     blend loop of chunk i (software prefetch); one shared buffer suffices because the vote barrier
//     already separates blend(i) from store(i+1);
//   * the tile-stop vote `atomicAdd(shared_t, uint(t*255))` (:97) is a warp reduction + 4 shared words;
//   * tiles are RE-QUEUEABLE: the grid is persistent (SMs x resident CTAs); a CTA blends at most GSR_COMP_QUANTUM
//     chunks of a tile, then spills the tile's 4 KB of per-pixel state and pushes the tile back to a device queue,
//     so a 19-chunk tile no longer pins one SM while others idle (ncu before: SMs active 60 % of the kernel).
//     State is saved/restored verbatim, so the result is unchanged.
// Arithmetic contract: "gsr deterministic math" (common.cuh): the GLSL-legal contractions of :84 and :89 are
// explicit fma, exp() is the det_exp() polynomial (evaluated here two lanes at a time).
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace gsr {

namespace {

constexpr int CHUNK = 256;    // gsplat_render.glsl:9 WORKGROUP_SIZE: splats per staged chunk / pixels per tile
constexpr int THREADS = 128;  // 2 pixels per thread
constexpr float MIN_ALPHA = 1.0f / 255.0f;

typedef unsigned long long u64;

#ifndef GSR_CPU_EMU
__device__ __forceinline__ u64 pk(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
#else  // tests/kernel_emu: the kernels of this file compiled for the CPU (test infrastructure; libgsr never defines GSR_CPU_EMU).
       // A packed op is two independent IEEE binary32 operations -- exactly what the PTX f32x2 instructions are.
inline u64 pk(float lo, float hi) { uint32_t a, b; memcpy(&a, &lo, 4); memcpy(&b, &hi, 4); return (u64)a | ((u64)b << 32); }
inline void upk(u64 v, float &lo, float &hi) { const uint32_t a = (uint32_t)v, b = (uint32_t)(v >> 32); memcpy(&lo, &a, 4); memcpy(&hi, &b, 4); }
inline u64 fma2(u64 a, u64 b, u64 c) { float al, ah, bl, bh, cl, ch; upk(a, al, ah); upk(b, bl, bh); upk(c, cl, ch); return pk(fmaf(al, bl, cl), fmaf(ah, bh, ch)); }
inline u64 mul2(u64 a, u64 b) { float al, ah, bl, bh; upk(a, al, ah); upk(b, bl, bh); return pk(al * bl, ah * bh); }
inline u64 add2(u64 a, u64 b) { float al, ah, bl, bh; upk(a, al, ah); upk(b, bl, bh); return pk(al + bl, ah + bh); }
inline u64 sub2(u64 a, u64 b) { float al, ah, bl, bh; upk(a, al, ah); upk(b, bl, bh); return pk(al - bl, ah - bh); }
#endif
__device__ __forceinline__ u64 bc(float x) { return pk(x, x); }

struct Staged {  // one gathered record, pre-scaled for the inner loop
    float4 a;    // image_pos.x, image_pos.y, -0.5*conic.x, -0.5*conic.z
    float4 b;    // -conic.y, opacity, color.r, color.g
    float c;     // color.b
};

__device__ __forceinline__ Staged null_splat() {
    Staged s;
    s.a = make_float4(0.f, 0.f, 0.f, 0.f);
    s.b = make_float4(0.f, 0.f, 0.f, 0.f);
    s.c = 0.f;
    return s;
}

__device__ __forceinline__ Staged gather(const float4 *__restrict__ records, const uint32_t *__restrict__ values, uint32_t idx) {
    const uint32_t v = __ldg(values + idx);
    const float4 *r = records + (uint64_t)v * 3u;
    const float4 r0 = __ldg(r + 0), r1 = __ldg(r + 1), r2 = __ldg(r + 2);
    Staged s;
    s.a = make_float4(r0.x, r0.y, -0.5f * r1.x, -0.5f * r1.z);
    s.b = make_float4(-r1.y, r2.w, r2.x, r2.y);
    s.c = r2.z;
    return s;
}

constexpr int COMP_MAX_PUSHES = GSR_COMP_MAX_PUSHES;  // common.cuh: queue capacity = COMP_MAX_PUSHES * tiles
#ifndef GSR_COMP_QUANTUM
#define GSR_COMP_QUANTUM 2  // chunks blended before an unfinished tile is handed back to the queue
#endif
constexpr uint32_t EXIT_TILE = 0xFFFFFFFFu;
#ifndef GSR_COMP_GROUP
#define GSR_COMP_GROUP 4  // splats per software-pipelined group of the blend loop
#endif
constexpr int GU = GSR_COMP_GROUP;
#ifndef GSR_COMP_WS_DEFAULT
#define GSR_COMP_WS_DEFAULT 0
#endif

#ifndef GSR_CPU_EMU
__device__ __forceinline__ unsigned long long globaltimer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ uint32_t smid() { uint32_t r; asm volatile("mov.u32 %0, %smid;" : "=r"(r)); return r; }
#else
inline unsigned long long globaltimer_ns() { return 0ull; }
inline uint32_t smid() { return 0u; }
#endif

// Persistent CTAs + re-queueable tiles.  Work item = (tile, first chunk).  Tickets [0, num_tiles) are the tiles
// themselves in natural order; an unfinished tile spills its per-pixel state (t, rgb of 256 pixels = 4 KB) and is
// pushed to `queue`, where ticket num_tiles + k finds it.  Every item ends in exactly one of {tile done, tile
// pushed}, so a CTA waiting for a queue slot either gets one or sees comp_done == num_tiles and leaves.
struct BlendK {  // broadcast constants of det_exp() for the packed lanes
    u64 L2E2, MAGIC2, ONE2, C6, C5, C4, C3, C2, C1;
};
__device__ __forceinline__ BlendK make_blend_k() {
    BlendK k;
    k.L2E2 = bc(0x1.715476p+0f); k.MAGIC2 = bc(12582912.0f); k.ONE2 = bc(1.0f);
    k.C6 = bc(0x1.446c7ep-13f); k.C5 = bc(0x1.5f48c8p-10f); k.C4 = bc(0x1.3b29d8p-7f); k.C3 = bc(0x1.c6aeccp-5f);
    k.C2 = bc(0x1.ebfbe0p-3f); k.C1 = bc(0x1.62e430p-1f);
    return k;
}

// ---- phase A: alpha = opacity * exp(power) of GU splats (slots j .. j+GU-1 of the staged chunk) for this thread's
//      two pixels, written stage by stage so that the GU ~25-instruction dependency chains can be interleaved (a lone
//      warp otherwise runs this at IPC 0.23: measured 25 us per chunk for a tile that owns its SM).  No dependence on
//      the transmittance: this part of gsplat_render.glsl:84-88 can run ahead of the sequential blend.
//      HWEXP (experiment, GSR_COMP_HWEXP=1): exp() through the SFU (MUFU.EX2 on the XU pipe) instead of the det_exp()
//      polynomial: 10 of the 28 FMA-pipe operations per splat and pixel pair disappear, but the result is no longer
//      bit-reproducible on a CPU (pixels within 1e-4 except where a `t > 1/255` exit or a tile-stop vote flips).
template <bool HWEXP>
__device__ __forceinline__ void phase_a(const float4 *s_a, const float4 *s_b, int j, u64 npx2, float fpy, const BlendK &K, u64 al2[GU]) {
    float4 A[GU];
    float bx[GU], by[GU], oy[GU];
    u64 ox2[GU], pw2[GU], tm2[GU], e2[GU];
#pragma unroll
    for (int u = 0; u < GU; ++u) { A[u] = s_a[j + u]; const float4 b = s_b[j + u]; bx[u] = b.x; by[u] = b.y; }
#pragma unroll
    for (int u = 0; u < GU; ++u) { ox2[u] = add2(bc(A[u].x), npx2); oy[u] = A[u].y - fpy; }
    // power = -0.5*(cx*ox*ox + cz*oy*oy) - cy*ox*oy  with q = fma(cz*oy, oy, cx*ox*ox), power = fma(-(cy*ox), oy, -0.5*q)
#pragma unroll
    for (int u = 0; u < GU; ++u) pw2[u] = mul2(bc(A[u].z), ox2[u]);
#pragma unroll
    for (int u = 0; u < GU; ++u) pw2[u] = mul2(pw2[u], ox2[u]);
#pragma unroll
    for (int u = 0; u < GU; ++u) pw2[u] = fma2(bc(A[u].w * oy[u]), bc(oy[u]), pw2[u]);
#pragma unroll
    for (int u = 0; u < GU; ++u) e2[u] = mul2(bc(bx[u]), ox2[u]);
#pragma unroll
    for (int u = 0; u < GU; ++u) pw2[u] = fma2(e2[u], bc(oy[u]), pw2[u]);
    // exp(power): det_exp(), two lanes at a time
#pragma unroll
    for (int u = 0; u < GU; ++u) pw2[u] = mul2(pw2[u], K.L2E2);
    if (HWEXP) {
#pragma unroll
        for (int u = 0; u < GU; ++u) {
            float tl, th;
            upk(pw2[u], tl, th);
#ifndef GSR_CPU_EMU
            asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(tl));  // one MUFU.EX2; results below 2^-126 flush to 0
            asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(th));
#else
            tl = exp2f(tl); th = exp2f(th);
#endif
            al2[u] = mul2(bc(by[u]), pk(tl, th));
        }
        return;
    }
#pragma unroll
    for (int u = 0; u < GU; ++u) {
        float tl, th;
        upk(pw2[u], tl, th);
        tl = g_min(g_max(tl, -127.0f), 128.0f);
        th = g_min(g_max(th, -127.0f), 128.0f);
        pw2[u] = pk(tl, th);
    }
#pragma unroll
    for (int u = 0; u < GU; ++u) tm2[u] = add2(pw2[u], K.MAGIC2);
#pragma unroll
    for (int u = 0; u < GU; ++u) al2[u] = sub2(tm2[u], K.MAGIC2);
#pragma unroll
    for (int u = 0; u < GU; ++u) pw2[u] = sub2(pw2[u], al2[u]);  // f
#pragma unroll
    for (int u = 0; u < GU; ++u) e2[u] = fma2(K.C6, pw2[u], K.C5);
#pragma unroll
    for (int u = 0; u < GU; ++u) e2[u] = fma2(e2[u], pw2[u], K.C4);
#pragma unroll
    for (int u = 0; u < GU; ++u) e2[u] = fma2(e2[u], pw2[u], K.C3);
#pragma unroll
    for (int u = 0; u < GU; ++u) e2[u] = fma2(e2[u], pw2[u], K.C2);
#pragma unroll
    for (int u = 0; u < GU; ++u) e2[u] = fma2(e2[u], pw2[u], K.C1);
#pragma unroll
    for (int u = 0; u < GU; ++u) e2[u] = fma2(e2[u], pw2[u], K.ONE2);
#pragma unroll
    for (int u = 0; u < GU; ++u) {
        float ml, mh;
        upk(tm2[u], ml, mh);
        tm2[u] = pk(__uint_as_float((__float_as_uint(ml) << 23) + 0x3F800000u),
                    __uint_as_float((__float_as_uint(mh) << 23) + 0x3F800000u));
    }
#pragma unroll
    for (int u = 0; u < GU; ++u) e2[u] = mul2(e2[u], tm2[u]);
#pragma unroll
    for (int u = 0; u < GU; ++u) al2[u] = mul2(bc(by[u]), e2[u]);
}

// ---- phase B: the sequential part (gsplat_render.glsl:89-90).  Dead pixels take alpha = 0: the reference's loop exit ----
__device__ __forceinline__ void phase_b(const float4 *s_b, const float *s_c, int j, const u64 al2[GU], const BlendK &K, u64 &cr2, u64 &cg2,
                                        u64 &cb2, float &t0, float &t1) {
#pragma unroll
    for (int u = 0; u < GU; ++u) {
        const float4 b = s_b[j + u];
        const float cbl = s_c[j + u];
        float al, ah;
        upk(al2[u], al, ah);
        al = (t0 > MIN_ALPHA) ? al : 0.0f;
        ah = (t1 > MIN_ALPHA) ? ah : 0.0f;
        const u64 m2 = pk(al, ah);
        const u64 t2 = pk(t0, t1);
        cr2 = fma2(mul2(bc(b.z), m2), t2, cr2);
        cg2 = fma2(mul2(bc(b.w), m2), t2, cg2);
        cb2 = fma2(mul2(bc(cbl), m2), t2, cb2);
        upk(mul2(t2, sub2(K.ONE2, m2)), t0, t1);
    }
}

#ifndef GSR_COMP_MIN_BLOCKS
#define GSR_COMP_MIN_BLOCKS 4  // lets ptxas spend registers on interleaving the per-splat dependency chains
#endif
template <bool HWEXP>
__global__ void __launch_bounds__(THREADS, GSR_COMP_MIN_BLOCKS) composite_kernel(const __grid_constant__ CompositeArgs p) {
    __shared__ float4 s_a[CHUNK];
    __shared__ float4 s_b[CHUNK];
    __shared__ float s_c[CHUNK];
    __shared__ uint32_t s_vote[THREADS / 32];
    __shared__ uint32_t s_tile, s_resume;

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const BlendK K = make_blend_k();
    uint32_t staged = 0;  // SURVEY 8 symbol C, summed over the items this CTA processed (uniform across the CTA)
    unsigned long long t_start = 0;  // trace only

    for (;;) {
        if (tid == 0) {
            // FIFO: fresh tiles are tickets [0, num_tiles); ticket num_tiles + k waits for the k-th hand-back.  (Serving
            // continuing tiles first was measured to be worse: the in-flight set monopolises the CTAs and the
            // remaining busy tiles start as a second wave.)
            const uint32_t ticket = atomicAdd(&p.frame->comp_head, 1u);
            if (ticket < (uint32_t)p.num_tiles) {
                // owned tiles: rows tile_begin/tiles_x + k*row_step, all columns (row_step == 1: one contiguous band)
                const uint32_t k = p.order ? p.order[ticket] : ticket;   // owned-tile index: longest lists first when an order is given
                s_tile = (uint32_t)p.tile_begin + (k / (uint32_t)p.tiles_x) * (uint32_t)(p.row_step * p.tiles_x) + k % (uint32_t)p.tiles_x;
                s_resume = 0u;
            } else {
                // a tile is pushed at most COMP_MAX_PUSHES times: later tickets can never be served and must not touch the queue
                const uint32_t qi = ticket - (uint32_t)p.num_tiles;
                const bool in_q = qi < (uint32_t)COMP_MAX_PUSHES * (uint32_t)p.num_tiles;
                volatile uint32_t *slot = p.queue + (in_q ? qi : 0u);
                volatile uint32_t *done = &p.frame->comp_done;
                uint32_t v = 0u;
                while (in_q && (v = *slot) == 0u && *done < (uint32_t)p.num_tiles) __nanosleep(200);
                if (in_q && v == 0u) v = *slot;  // a push may have landed between the two reads
                s_tile = v ? v - 1u : EXIT_TILE;
                s_resume = 1u;
                __threadfence();
            }
            if (p.trace) t_start = globaltimer_ns();
        }
        __syncthreads();
        const uint32_t tile_id = s_tile;
        const bool resume = s_resume != 0u;
        if (tile_id == EXIT_TILE) break;

        const uint32_t tx = tile_id % (uint32_t)p.tiles_x, ty = tile_id / (uint32_t)p.tiles_x;
        const int px0 = (int)(tx * TILE + 2u * (tid & 7u)), py = (int)(ty * TILE + (tid >> 3));
        const u64 npx2 = pk(-(float)px0, -(float)(px0 + 1));  // ox = image_pos.x - pixel.x  ==  image_pos.x + (-pixel.x)
        const float fpy = (float)py;

        const uint2 bounds = p.bounds[tile_id];
        const int32_t diff = (int32_t)(bounds.y - bounds.x);
        const int num_splats = diff > 0 ? diff : 0;                              // :61
        const int num_iterations = (int)ceilf((float)num_splats / (float)CHUNK);  // :62

        u64 cr2 = pk(0.f, 0.f), cg2 = cr2, cb2 = cr2;  // blended colour of the two pixels
        float t0 = 1.0f, t1 = 1.0f;                    // transmittance of the two pixels
        int i0 = 0;
        // index of this tile among the owned tiles (slot of its spilled state)
        const uint32_t rel = tile_id - (uint32_t)p.tile_begin;
        const uint32_t local_tile = (rel / (uint32_t)(p.row_step * p.tiles_x)) * (uint32_t)p.tiles_x + rel % (uint32_t)p.tiles_x;
        float4 *st = p.state + (uint64_t)local_tile * (2u * THREADS);
        if (resume) {  // written by another SM during this launch: read through L2
            const float4 sa = __ldcg(st + tid), sb = __ldcg(st + THREADS + tid);
            cr2 = pk(sa.x, sa.y); cg2 = pk(sa.z, sa.w); cb2 = pk(sb.x, sb.y);
            t0 = sb.z; t1 = sb.w;
            i0 = (int)__ldcg(p.state_chunk + local_tile);
        }

        Staged n0 = null_splat(), n1 = null_splat();
        if (i0 < num_iterations) {
            const int nb = CHUNK * i0;
            if (nb + (int)tid < num_splats) n0 = gather(p.records, p.values, bounds.x + (uint32_t)nb + tid);
            if (nb + (int)tid + THREADS < num_splats) n1 = gather(p.records, p.values, bounds.x + (uint32_t)nb + tid + THREADS);
        }

        // at most COMP_MAX_PUSHES hand-backs per tile bound the queue: long lists get a proportionally longer quantum
        // a resumed item means every fresh tile has been handed out (tickets are monotonic): yielding again would only cost a
        // spill + restore, so (requeue_only_if_fresh) it runs to completion; a fresh tile yields after `quantum` chunks
        const int q_min = p.quantum > (num_iterations + COMP_MAX_PUSHES) / (COMP_MAX_PUSHES + 1)
                              ? p.quantum : (num_iterations + COMP_MAX_PUSHES) / (COMP_MAX_PUSHES + 1);
        const int quantum = (p.requeue_only_if_fresh && resume) ? 0x3fffffff : q_min;
        const int i_begin = i0;
        bool finished = true;
        for (int i = i_begin; i < num_iterations; ++i) {
            const int sort_offset = CHUNK * i;
            const int chunk = (num_splats - sort_offset) < CHUNK ? (num_splats - sort_offset) : CHUNK;
            staged += (uint32_t)chunk;
            s_a[tid] = n0.a; s_b[tid] = n0.b; s_c[tid] = n0.c;
            s_a[tid + THREADS] = n1.a; s_b[tid + THREADS] = n1.b; s_c[tid + THREADS] = n1.c;
            __syncthreads();
            // prefetch the next chunk's records while this one is blended (slots past the list end become null splats)
            n0 = null_splat(); n1 = null_splat();
            if (i + 1 < num_iterations && i + 1 - i_begin < quantum) {
                const int nb = sort_offset + CHUNK;
                if (nb + (int)tid < num_splats) n0 = gather(p.records, p.values, bounds.x + (uint32_t)nb + tid);
                if (nb + (int)tid + THREADS < num_splats) n1 = gather(p.records, p.values, bounds.x + (uint32_t)nb + tid + THREADS);
            }

            // :79-91, four splats per liveness test; `chunk` rounded up to 4 reads null splats (opacity 0 => exact no-op)
            const int chunk4 = (chunk + GU - 1) & ~(GU - 1);
            for (int j = 0; j < chunk4; j += GU) {
                if (!__any_sync(0xffffffffu, (t0 > MIN_ALPHA) || (t1 > MIN_ALPHA))) break;
                u64 al2[GU];
                phase_a<HWEXP>(s_a, s_b, j, npx2, fpy, K, al2);
                phase_b(s_b, s_c, j, al2, K, cr2, cg2, cb2, t0, t1);
            }

            // :97 tile-stop vote: continue only if the sum over the tile's 256 pixels of uint(t*255) exceeds 255
            const uint32_t wsum = __reduce_add_sync(0xffffffffu, (uint32_t)(t0 * 255.0f) + (uint32_t)(t1 * 255.0f));
            if (lane == 0) s_vote[warp] = wsum;
            __syncthreads();
            uint32_t shared_t = 0;
#pragma unroll
            for (int w = 0; w < THREADS / 32; ++w) shared_t += s_vote[w];
            if (!(shared_t > 255u)) break;  // finished stays true
            if (i + 1 < num_iterations && i + 1 - i_begin >= quantum) {  // quantum used up, tile still live
                finished = false;
                i0 = i + 1;  // resume point, spilled below
                break;
            }
        }

        float r0, r1, g0, g1, b0, b1;
        upk(cr2, r0, r1); upk(cg2, g0, g1); upk(cb2, b0, b1);
        if (!finished) {
            // spill the tile and hand it back: any CTA on any SM resumes it
            __stcg(st + tid, make_float4(r0, r1, g0, g1));
            __stcg(st + THREADS + tid, make_float4(b0, b1, t0, t1));
            if (tid == 0) __stcg(p.state_chunk + local_tile, (uint32_t)i0);
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                const uint32_t slot = atomicAdd(&p.frame->comp_tail, 1u);
                __threadfence();
                *(volatile uint32_t *)(p.queue + slot) = tile_id + 1u;
            }
        } else {
            // :100-101
            const float hx = (float)num_splats * 5e-4f;
            const float h0 = 0.0f * (1.0f - hx) + 1.0f * hx, h1 = 0.0f * (1.0f - hx) + 0.2f * hx, h2c = 1.0f * (1.0f - hx) + 0.2f * hx;
            if (py < p.height) {
                float4 *row = p.out + (uint64_t)py * (uint64_t)p.width;
                const float k0 = 1.0f - t0, k1 = 1.0f - t1;
                if (px0 < p.width)
                    row[px0] = make_float4(r0 + h0 * k0 * p.heatmap_factor, g0 + h1 * k0 * p.heatmap_factor, b0 + h2c * k0 * p.heatmap_factor, 1.0f);
                if (px0 + 1 < p.width)
                    row[px0 + 1] = make_float4(r1 + h0 * k1 * p.heatmap_factor, g1 + h1 * k1 * p.heatmap_factor, b1 + h2c * k1 * p.heatmap_factor, 1.0f);
            }
            // :105-110 pick: the elected (first) lane of each 32-wide subgroup of the reference's 16x16 workgroup is local
            // index 32*s = pixel (0, 2*s) of the tile = first pixel of thread 16*s here
            if ((tid & 15u) == 0u && tile_id == p.target_tile_id && t0 != 1.0f) {
                const uint32_t v = p.values[bounds.x + (bounds.y - bounds.x) / 10u];
                const float4 q0 = p.records[(uint64_t)v * 3u + 0], q1 = p.records[(uint64_t)v * 3u + 1];
                *p.pick = make_float4(q0.z, q0.w, q1.w, (float)num_splats);
            }
            if (tid == 0) atomicAdd(&p.frame->comp_done, 1u);
        }
        if (p.trace && tid == 0) {
            const uint32_t k = atomicAdd(p.trace_count, 1u);
            if (k < p.trace_cap) p.trace[k] = make_ulonglong4(((unsigned long long)tile_id << 32) | smid(), t_start, globaltimer_ns(), ((unsigned long long)(uint32_t)i_begin << 32) | (uint32_t)(finished ? 1u : 0u) | ((uint32_t)num_iterations << 1));
        }
        __syncthreads();  // s_tile / staging buffers are reused by the next item
    }
    if (tid == 0 && staged && p.count_staged) atomicAdd(&p.frame->staged, (unsigned long long)staged);
}

// ==============================================================================================================
// EXPERIMENTAL "v2" staging (GSR_COMP_V2=1; off by default, NOT yet run on a GPU -- written at the end of round 1 from the
// ncu stall profile of composite_kernel: barrier 2.68 and math_pipe_throttle 1.57 warps per issue, FMA pipe 53 % busy,
// 106 registers => 4 CTAs/SM).  Same tiles, same queue, same arithmetic, same results; what changes is the staging:
//   * the records of chunk i+1 travel global -> shared with cp.async (LDGSTS) into per-thread private raw slots while
//     chunk i is blended, instead of living in 18 registers across the blend loop (the register budget decides how many
//     CTAs share an SM, and co-resident CTAs are what fills the barrier stalls);
//   * the copies land directly in the OTHER half of a double-buffered staging area (24 KB in all), where each thread
//     then pre-scales its own two records in place, so the
//     "staged data visible" barrier and the tile-stop-vote barrier of the reference (:70,:77,:98) become ONE
//     __syncthreads per chunk (vote words double-buffered by chunk parity for the same reason).
// phase A / phase B of composite_kernel over the one-array staging layout of the v2 / p4 kernels (slot k = float4[3k..3k+2])
__device__ __forceinline__ void phase_a_st(const float4 *s, int j, u64 npx2, float fpy, const BlendK &K, u64 al2[GU]) {
    float4 A[GU];
    float bx[GU], by[GU], oy[GU];
    u64 ox2[GU], pw2[GU], tm2[GU], e2[GU];
#pragma unroll
    for (int u = 0; u < GU; ++u) { A[u] = s[3 * (j + u)]; const float4 b = s[3 * (j + u) + 1]; bx[u] = b.x; by[u] = b.y; }
#pragma unroll
    for (int u = 0; u < GU; ++u) { ox2[u] = add2(bc(A[u].x), npx2); oy[u] = A[u].y - fpy; }
    // power = -0.5*(cx*ox*ox + cz*oy*oy) - cy*ox*oy  with q = fma(cz*oy, oy, cx*ox*ox), power = fma(-(cy*ox), oy, -0.5*q)
#pragma unroll
    for (int u = 0; u < GU; ++u) pw2[u] = mul2(bc(A[u].z), ox2[u]);
#pragma unroll
    for (int u = 0; u < GU; ++u) pw2[u] = mul2(pw2[u], ox2[u]);
#pragma unroll
    for (int u = 0; u < GU; ++u) pw2[u] = fma2(bc(A[u].w * oy[u]), bc(oy[u]), pw2[u]);
#pragma unroll
    for (int u = 0; u < GU; ++u) e2[u] = mul2(bc(bx[u]), ox2[u]);
#pragma unroll
    for (int u = 0; u < GU; ++u) pw2[u] = fma2(e2[u], bc(oy[u]), pw2[u]);
    // exp(power): det_exp(), two lanes at a time
#pragma unroll
    for (int u = 0; u < GU; ++u) pw2[u] = mul2(pw2[u], K.L2E2);
#pragma unroll
    for (int u = 0; u < GU; ++u) {
        float tl, th;
        upk(pw2[u], tl, th);
        tl = g_min(g_max(tl, -127.0f), 128.0f);
        th = g_min(g_max(th, -127.0f), 128.0f);
        pw2[u] = pk(tl, th);
    }
#pragma unroll
    for (int u = 0; u < GU; ++u) tm2[u] = add2(pw2[u], K.MAGIC2);
#pragma unroll
    for (int u = 0; u < GU; ++u) al2[u] = sub2(tm2[u], K.MAGIC2);
#pragma unroll
    for (int u = 0; u < GU; ++u) pw2[u] = sub2(pw2[u], al2[u]);  // f
#pragma unroll
    for (int u = 0; u < GU; ++u) e2[u] = fma2(K.C6, pw2[u], K.C5);
#pragma unroll
    for (int u = 0; u < GU; ++u) e2[u] = fma2(e2[u], pw2[u], K.C4);
#pragma unroll
    for (int u = 0; u < GU; ++u) e2[u] = fma2(e2[u], pw2[u], K.C3);
#pragma unroll
    for (int u = 0; u < GU; ++u) e2[u] = fma2(e2[u], pw2[u], K.C2);
#pragma unroll
    for (int u = 0; u < GU; ++u) e2[u] = fma2(e2[u], pw2[u], K.C1);
#pragma unroll
    for (int u = 0; u < GU; ++u) e2[u] = fma2(e2[u], pw2[u], K.ONE2);
#pragma unroll
    for (int u = 0; u < GU; ++u) {
        float ml, mh;
        upk(tm2[u], ml, mh);
        tm2[u] = pk(__uint_as_float((__float_as_uint(ml) << 23) + 0x3F800000u),
                    __uint_as_float((__float_as_uint(mh) << 23) + 0x3F800000u));
    }
#pragma unroll
    for (int u = 0; u < GU; ++u) e2[u] = mul2(e2[u], tm2[u]);
#pragma unroll
    for (int u = 0; u < GU; ++u) al2[u] = mul2(bc(by[u]), e2[u]);
}

__device__ __forceinline__ void phase_b_st(const float4 *s, int j, const u64 al2[GU], const BlendK &K, u64 &cr2, u64 &cg2,
                                        u64 &cb2, float &t0, float &t1) {
#pragma unroll
    for (int u = 0; u < GU; ++u) {
        const float4 b = s[3 * (j + u) + 1];
        const float cbl = s[3 * (j + u) + 2].x;
        float al, ah;
        upk(al2[u], al, ah);
        al = (t0 > MIN_ALPHA) ? al : 0.0f;
        ah = (t1 > MIN_ALPHA) ? ah : 0.0f;
        const u64 m2 = pk(al, ah);
        const u64 t2 = pk(t0, t1);
        cr2 = fma2(mul2(bc(b.z), m2), t2, cr2);
        cg2 = fma2(mul2(bc(b.w), m2), t2, cg2);
        cb2 = fma2(mul2(bc(cbl), m2), t2, cb2);
        upk(mul2(t2, sub2(K.ONE2, m2)), t0, t1);
    }
}

#ifndef GSR_CPU_EMU
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
#else
inline void cp_async16(void *smem_dst, const void *gmem_src) { memcpy(smem_dst, gmem_src, 16); }
inline void cp_async_commit_wait_all() {}
#endif

template <int MIN_BLOCKS>   // CTAs per SM the register allocation targets: 5 -> 82 registers, 6 -> 70, 8 -> 62 (ptxas, no spills)
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS) composite_v2_kernel(const __grid_constant__ CompositeArgs p) {
    __shared__ float4 s_st[2][CHUNK * 3];   // two staging halves; slot k = float4[3k] (a), [3k+1] (b), [3k+2].x (c): the 48-byte
                                            // record lands there raw (cp.async) and is pre-scaled in place by the thread that fetched it
    __shared__ uint32_t s_vote[2][THREADS / 32];
    __shared__ uint32_t s_tile, s_resume;

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const BlendK K = make_blend_k();
    uint32_t staged = 0;
    unsigned long long t_start = 0;

    for (;;) {
        if (tid == 0) {
            const uint32_t ticket = atomicAdd(&p.frame->comp_head, 1u);
            if (ticket < (uint32_t)p.num_tiles) {
                const uint32_t k = p.order ? p.order[ticket] : ticket;   // owned-tile index: longest lists first when an order is given
                s_tile = (uint32_t)p.tile_begin + (k / (uint32_t)p.tiles_x) * (uint32_t)(p.row_step * p.tiles_x) + k % (uint32_t)p.tiles_x;
                s_resume = 0u;
            } else {
                // a tile is pushed at most COMP_MAX_PUSHES times: later tickets can never be served and must not touch the queue
                const uint32_t qi = ticket - (uint32_t)p.num_tiles;
                const bool in_q = qi < (uint32_t)COMP_MAX_PUSHES * (uint32_t)p.num_tiles;
                volatile uint32_t *slot = p.queue + (in_q ? qi : 0u);
                volatile uint32_t *done = &p.frame->comp_done;
                uint32_t v = 0u;
                while (in_q && (v = *slot) == 0u && *done < (uint32_t)p.num_tiles) __nanosleep(200);
                if (in_q && v == 0u) v = *slot;
                s_tile = v ? v - 1u : EXIT_TILE;
                s_resume = 1u;
                __threadfence();
            }
            if (p.trace) t_start = globaltimer_ns();
        }
        __syncthreads();
        const uint32_t tile_id = s_tile;
        const bool resume = s_resume != 0u;
        if (tile_id == EXIT_TILE) break;

        const uint32_t tx = tile_id % (uint32_t)p.tiles_x, ty = tile_id / (uint32_t)p.tiles_x;
        const int px0 = (int)(tx * TILE + 2u * (tid & 7u)), py = (int)(ty * TILE + (tid >> 3));
        const u64 npx2 = pk(-(float)px0, -(float)(px0 + 1));
        const float fpy = (float)py;

        const uint2 bounds = p.bounds[tile_id];
        const int32_t diff = (int32_t)(bounds.y - bounds.x);
        const int num_splats = diff > 0 ? diff : 0;
        const int num_iterations = (int)ceilf((float)num_splats / (float)CHUNK);

        u64 cr2 = pk(0.f, 0.f), cg2 = cr2, cb2 = cr2;
        float t0 = 1.0f, t1 = 1.0f;
        int i0 = 0;
        const uint32_t rel = tile_id - (uint32_t)p.tile_begin;
        const uint32_t local_tile = (rel / (uint32_t)(p.row_step * p.tiles_x)) * (uint32_t)p.tiles_x + rel % (uint32_t)p.tiles_x;
        float4 *st = p.state + (uint64_t)local_tile * (2u * THREADS);
        if (resume) {
            const float4 sa = __ldcg(st + tid), sb = __ldcg(st + THREADS + tid);
            cr2 = pk(sa.x, sa.y); cg2 = pk(sa.z, sa.w); cb2 = pk(sb.x, sb.y);
            t0 = sb.z; t1 = sb.w;
            i0 = (int)__ldcg(p.state_chunk + local_tile);
        }

        // splat ids of this thread's two slots of a chunk (one chunk ahead of the records, two ahead of the blend)
        auto load_ids = [&](int ci, uint32_t &v0, uint32_t &v1) {
            const int k0 = CHUNK * ci + (int)tid, k1 = k0 + THREADS;
            v0 = (ci < num_iterations && k0 < num_splats) ? __ldg(p.values + bounds.x + (uint32_t)k0) : 0xFFFFFFFFu;
            v1 = (ci < num_iterations && k1 < num_splats) ? __ldg(p.values + bounds.x + (uint32_t)k1) : 0xFFFFFFFFu;
        };
        auto issue = [&](uint32_t v0, uint32_t v1, int b) {   // raw records of the two splats -> this thread's slots of half b
            if (v0 != 0xFFFFFFFFu) {
                const float4 *r = p.records + (uint64_t)v0 * 3u;
                float4 *d = &s_st[b][3 * tid];
                cp_async16(d + 0, r + 0); cp_async16(d + 1, r + 1); cp_async16(d + 2, r + 2);
            }
            if (v1 != 0xFFFFFFFFu) {
                const float4 *r = p.records + (uint64_t)v1 * 3u;
                float4 *d = &s_st[b][3 * (tid + THREADS)];
                cp_async16(d + 0, r + 0); cp_async16(d + 1, r + 1); cp_async16(d + 2, r + 2);
            }
        };
        auto finalize = [&](uint32_t v0, uint32_t v1, int b) {   // wait for the own copies, pre-scale in place like gather()
            cp_async_commit_wait_all();
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                float4 *d = &s_st[b][3 * (tid + (uint32_t)hh * THREADS)];
                Staged sgd = null_splat();
                if ((hh ? v1 : v0) != 0xFFFFFFFFu) {
                    const float4 r0 = d[0], r1 = d[1], r2 = d[2];
                    sgd.a = make_float4(r0.x, r0.y, -0.5f * r1.x, -0.5f * r1.z);
                    sgd.b = make_float4(-r1.y, r2.w, r2.x, r2.y);
                    sgd.c = r2.z;
                }
                d[0] = sgd.a; d[1] = sgd.b; d[2] = make_float4(sgd.c, 0.f, 0.f, 0.f);
            }
        };

        // a resumed item means every fresh tile has been handed out (tickets are monotonic): yielding again would only cost a
        // spill + restore, so (requeue_only_if_fresh) it runs to completion; a fresh tile yields after `quantum` chunks
        const int q_min = p.quantum > (num_iterations + COMP_MAX_PUSHES) / (COMP_MAX_PUSHES + 1)
                              ? p.quantum : (num_iterations + COMP_MAX_PUSHES) / (COMP_MAX_PUSHES + 1);
        const int quantum = (p.requeue_only_if_fresh && resume) ? 0x3fffffff : q_min;
        const int i_begin = i0;
        bool finished = true;
        int b = 0;
        uint32_t va0, va1, vb0 = 0xFFFFFFFFu, vb1 = 0xFFFFFFFFu;   // ids of the chunk being fetched / of the one after it
        if (i0 < num_iterations) {
            load_ids(i0, va0, va1);
            issue(va0, va1, 0);
            load_ids(i0 + 1, vb0, vb1);
            finalize(va0, va1, 0);
        }
        __syncthreads();   // staging half 0 visible (also orders the previous tile's last reads before this tile's writes)
        for (int i = i_begin; i < num_iterations; ++i) {
            const int sort_offset = CHUNK * i;
            const int chunk = (num_splats - sort_offset) < CHUNK ? (num_splats - sort_offset) : CHUNK;
            staged += (uint32_t)chunk;
            const bool fetch_next = (i + 1 < num_iterations) && (i + 1 - i_begin < quantum);
            if (fetch_next) {   // chunk i+1: records in flight during this blend, ids of chunk i+2 behind them
                va0 = vb0; va1 = vb1;
                issue(va0, va1, b ^ 1);   // half b^1 was last read in blend(i-1): everybody is past that barrier
                load_ids(i + 2, vb0, vb1);
            }

            const int chunk4 = (chunk + GU - 1) & ~(GU - 1);
            for (int j = 0; j < chunk4; j += GU) {
                if (!__any_sync(0xffffffffu, (t0 > MIN_ALPHA) || (t1 > MIN_ALPHA))) break;
                u64 al2[GU];
                phase_a_st(s_st[b], j, npx2, fpy, K, al2);
                phase_b_st(s_st[b], j, al2, K, cr2, cg2, cb2, t0, t1);
            }
            if (fetch_next) finalize(va0, va1, b ^ 1);   // nobody reads half b^1 before the barrier below

            const uint32_t wsum = __reduce_add_sync(0xffffffffu, (uint32_t)(t0 * 255.0f) + (uint32_t)(t1 * 255.0f));
            if (lane == 0) s_vote[i & 1][warp] = wsum;
            __syncthreads();   // the ONE barrier of the chunk: votes of chunk i and staging half b^1 (chunk i+1) visible
            uint32_t shared_t = 0;
#pragma unroll
            for (int w = 0; w < THREADS / 32; ++w) shared_t += s_vote[i & 1][w];
            if (!(shared_t > 255u)) break;
            if (i + 1 < num_iterations && i + 1 - i_begin >= quantum) {
                finished = false;
                i0 = i + 1;
                break;
            }
            b ^= 1;
        }

        float r0, r1, g0, g1, b0, b1;
        upk(cr2, r0, r1); upk(cg2, g0, g1); upk(cb2, b0, b1);
        if (!finished) {
            __stcg(st + tid, make_float4(r0, r1, g0, g1));
            __stcg(st + THREADS + tid, make_float4(b0, b1, t0, t1));
            if (tid == 0) __stcg(p.state_chunk + local_tile, (uint32_t)i0);
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                const uint32_t slot = atomicAdd(&p.frame->comp_tail, 1u);
                __threadfence();
                *(volatile uint32_t *)(p.queue + slot) = tile_id + 1u;
            }
        } else {
            const float hx = (float)num_splats * 5e-4f;  // :100-101
            const float h0 = 0.0f * (1.0f - hx) + 1.0f * hx, h1 = 0.0f * (1.0f - hx) + 0.2f * hx, h2c = 1.0f * (1.0f - hx) + 0.2f * hx;
            if (py < p.height) {
                float4 *row = p.out + (uint64_t)py * (uint64_t)p.width;
                const float k0 = 1.0f - t0, k1 = 1.0f - t1;
                if (px0 < p.width)
                    row[px0] = make_float4(r0 + h0 * k0 * p.heatmap_factor, g0 + h1 * k0 * p.heatmap_factor, b0 + h2c * k0 * p.heatmap_factor, 1.0f);
                if (px0 + 1 < p.width)
                    row[px0 + 1] = make_float4(r1 + h0 * k1 * p.heatmap_factor, g1 + h1 * k1 * p.heatmap_factor, b1 + h2c * k1 * p.heatmap_factor, 1.0f);
            }
            if ((tid & 15u) == 0u && tile_id == p.target_tile_id && t0 != 1.0f) {  // :105-110 pick
                const uint32_t v = p.values[bounds.x + (bounds.y - bounds.x) / 10u];
                const float4 q0 = p.records[(uint64_t)v * 3u + 0], q1 = p.records[(uint64_t)v * 3u + 1];
                *p.pick = make_float4(q0.z, q0.w, q1.w, (float)num_splats);
            }
            if (tid == 0) atomicAdd(&p.frame->comp_done, 1u);
        }
        if (p.trace && tid == 0) {
            const uint32_t k = atomicAdd(p.trace_count, 1u);
            if (k < p.trace_cap) p.trace[k] = make_ulonglong4(((unsigned long long)tile_id << 32) | smid(), t_start, globaltimer_ns(), ((unsigned long long)(uint32_t)i_begin << 32) | (uint32_t)(finished ? 1u : 0u) | ((uint32_t)num_iterations << 1));
        }
        __syncthreads();  // s_tile / staging buffers are reused by the next item
    }
    if (tid == 0 && staged && p.count_staged) atomicAdd(&p.frame->staged, (unsigned long long)staged);
}

// ==============================================================================================================
// "v3": the v2 structure (24 KB double-buffered staging, ONE __syncthreads per chunk, >= 6 CTAs/SM) with
//   * TMA staging: every thread fetches the two raw 48-byte records of its slots with `cp.async.bulk` (SASS UBLKCP) onto the
//     half's mbarrier -- no address registers held across the blend, no LDGSTS triplets; thread 0 arms the barrier with the
//     chunk's byte count; everybody waits on it once, then pre-scales its own two records in place;
//   * a two-instruction transmittance chain: alpha and (1 - alpha) are formed off the loop-carried path, the per-splat update is
//     FMUL2 + FSEL (`t = alive ? t * (1 - alpha) : t`, bit-identical to multiplying by 1 - 0), instead of
//     FSETP -> FSEL -> FADD2 -> FMUL2: a lone tile (tail of the kernel, sparse multi-GPU shards) advances faster;
//   * CVT (experiment): round-to-nearest of the exp2 argument with F2I.RN / I2F (conversion pipe) instead of the two magic-constant
//     FADD2s (FMA pipe) -- same integer, same fraction, same bits.
#ifndef GSR_CPU_EMU
__device__ __forceinline__ uint32_t comp_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void comp_mbar_init(uint64_t *bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(comp_smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void comp_fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void comp_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void comp_mbar_expect_tx(uint64_t *bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(comp_smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void comp_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(comp_smem_u32(dst)), "l"(src), "r"(bytes), "r"(comp_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void comp_mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(comp_smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
#else
inline void comp_mbar_init(uint64_t *, uint32_t) {}
inline void comp_fence_mbar_init() {}
inline void comp_fence_proxy_async() {}
inline void comp_mbar_expect_tx(uint64_t *, uint32_t) {}
inline void comp_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *) { memcpy(dst, src, bytes); }
inline void comp_mbar_wait(uint64_t *, uint32_t) {}
#endif

// phase A over the one-array staging layout; also returns om2 = 1 - alpha (off the transmittance chain)
template <bool CVT>
__device__ __forceinline__ void phase_a_v3(const float4 A[GU], const float4 B[GU], u64 npx2, float fpy, const BlendK &K, u64 al2[GU], u64 om2[GU]) {
    float bx[GU], by[GU], oy[GU];
    u64 ox2[GU], pw2[GU], tm2[GU], e2[GU];
#pragma unroll
    for (int u = 0; u < GU; ++u) { bx[u] = B[u].x; by[u] = B[u].y; }
#pragma unroll
    for (int u = 0; u < GU; ++u) { ox2[u] = add2(bc(A[u].x), npx2); oy[u] = A[u].y - fpy; }
#pragma unroll
    for (int u = 0; u < GU; ++u) pw2[u] = mul2(bc(A[u].z), ox2[u]);
#pragma unroll
    for (int u = 0; u < GU; ++u) pw2[u] = mul2(pw2[u], ox2[u]);
#pragma unroll
    for (int u = 0; u < GU; ++u) pw2[u] = fma2(bc(A[u].w * oy[u]), bc(oy[u]), pw2[u]);
#pragma unroll
    for (int u = 0; u < GU; ++u) e2[u] = mul2(bc(bx[u]), ox2[u]);
#pragma unroll
    for (int u = 0; u < GU; ++u) pw2[u] = fma2(e2[u], bc(oy[u]), pw2[u]);
#pragma unroll
    for (int u = 0; u < GU; ++u) pw2[u] = mul2(pw2[u], K.L2E2);
#pragma unroll
    for (int u = 0; u < GU; ++u) {
        float tl, th;
        upk(pw2[u], tl, th);
        tl = g_min(g_max(tl, -127.0f), 128.0f);
        th = g_min(g_max(th, -127.0f), 128.0f);
        pw2[u] = pk(tl, th);
    }
    if (CVT) {
        // n = rint(t) (ties to even, |t| <= 128): F2I.RN gives the integer the 1.5*2^23 trick leaves in the mantissa; the scale
        // 2^n is assembled from it directly, and f = t - float(n) is the same subtraction
#pragma unroll
        for (int u = 0; u < GU; ++u) {
            float tl, th;
            upk(pw2[u], tl, th);
#ifndef GSR_CPU_EMU
            const int nl = __float2int_rn(tl), nh = __float2int_rn(th);
#else
            const int nl = (int)nearbyintf(tl), nh = (int)nearbyintf(th);
#endif
            al2[u] = pk((float)nl, (float)nh);
            tm2[u] = pk(__uint_as_float(((uint32_t)nl << 23) + 0x3F800000u), __uint_as_float(((uint32_t)nh << 23) + 0x3F800000u));
        }
    } else {
#pragma unroll
        for (int u = 0; u < GU; ++u) tm2[u] = add2(pw2[u], K.MAGIC2);
#pragma unroll
        for (int u = 0; u < GU; ++u) al2[u] = sub2(tm2[u], K.MAGIC2);
    }
#pragma unroll
    for (int u = 0; u < GU; ++u) pw2[u] = sub2(pw2[u], al2[u]);  // f
#pragma unroll
    for (int u = 0; u < GU; ++u) e2[u] = fma2(K.C6, pw2[u], K.C5);
#pragma unroll
    for (int u = 0; u < GU; ++u) e2[u] = fma2(e2[u], pw2[u], K.C4);
#pragma unroll
    for (int u = 0; u < GU; ++u) e2[u] = fma2(e2[u], pw2[u], K.C3);
#pragma unroll
    for (int u = 0; u < GU; ++u) e2[u] = fma2(e2[u], pw2[u], K.C2);
#pragma unroll
    for (int u = 0; u < GU; ++u) e2[u] = fma2(e2[u], pw2[u], K.C1);
#pragma unroll
    for (int u = 0; u < GU; ++u) e2[u] = fma2(e2[u], pw2[u], K.ONE2);
    if (!CVT) {
#pragma unroll
        for (int u = 0; u < GU; ++u) {
            float ml, mh;
            upk(tm2[u], ml, mh);
            tm2[u] = pk(__uint_as_float((__float_as_uint(ml) << 23) + 0x3F800000u), __uint_as_float((__float_as_uint(mh) << 23) + 0x3F800000u));
        }
    }
#pragma unroll
    for (int u = 0; u < GU; ++u) e2[u] = mul2(e2[u], tm2[u]);
#pragma unroll
    for (int u = 0; u < GU; ++u) al2[u] = mul2(bc(by[u]), e2[u]);
#pragma unroll
    for (int u = 0; u < GU; ++u) om2[u] = sub2(K.ONE2, al2[u]);
}

// phase B with the short transmittance chain: per splat FMUL2 (t * (1 - alpha)) and one select per pixel
__device__ __forceinline__ void phase_b_v3(const float4 *s, int j, const float4 B[GU], const u64 al2[GU], const u64 om2[GU], u64 &cr2, u64 &cg2, u64 &cb2,
                                           float &t0, float &t1, bool &alive_mid) {
#pragma unroll
    for (int u = 0; u < GU; ++u) {
        const float4 b = B[u];
        const float cbl = s[3 * (j + u) + 2].x;
        if (u == GU / 2) alive_mid = (t0 > MIN_ALPHA) || (t1 > MIN_ALPHA);   // liveness half a group early: the vote leaves the critical path
        const bool a0 = t0 > MIN_ALPHA, a1 = t1 > MIN_ALPHA;   // gsplat_render.glsl:79: a dead pixel has left the loop
        float al, ah, pl, ph;
        upk(al2[u], al, ah);
        const u64 t2 = pk(t0, t1);
        upk(mul2(t2, om2[u]), pl, ph);
        const u64 m2 = pk(a0 ? al : 0.0f, a1 ? ah : 0.0f);
        cr2 = fma2(mul2(bc(b.z), m2), t2, cr2);
        cg2 = fma2(mul2(bc(b.w), m2), t2, cg2);
        cb2 = fma2(mul2(bc(cbl), m2), t2, cb2);
        t0 = a0 ? pl : t0;
        t1 = a1 ? ph : t1;
    }
}

template <int MIN_BLOCKS, bool CVT, bool PIPE>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS) composite_v3_kernel(const __grid_constant__ CompositeArgs p) {
    __shared__ __align__(128) float4 s_st[2][CHUNK * 3];   // two staging halves; slot k = float4[3k..3k+2]: the raw 48-byte record lands
                                                           // there (TMA) and is pre-scaled in place by the thread that fetched it
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ uint32_t s_vote[2][THREADS / 32];
    __shared__ uint32_t s_tile, s_resume;

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const BlendK K = make_blend_k();
    uint32_t staged = 0;
    unsigned long long t_start = 0;
    uint32_t phase0 = 0u, phase1 = 0u;   // mbarrier phase parity of the two halves (uniform across the CTA)
    if (tid == 0) { comp_mbar_init(&s_bar[0], 1); comp_mbar_init(&s_bar[1], 1); comp_fence_mbar_init(); }
    __syncthreads();

    for (;;) {
        if (tid == 0) {
            const uint32_t ticket = atomicAdd(&p.frame->comp_head, 1u);
            if (ticket < (uint32_t)p.num_tiles) {
                const uint32_t k = p.order ? p.order[ticket] : ticket;   // owned-tile index: longest lists first when an order is given
                s_tile = (uint32_t)p.tile_begin + (k / (uint32_t)p.tiles_x) * (uint32_t)(p.row_step * p.tiles_x) + k % (uint32_t)p.tiles_x;
                s_resume = 0u;
            } else {
                // a tile is pushed at most COMP_MAX_PUSHES times: later tickets can never be served and must not touch the queue
                const uint32_t qi = ticket - (uint32_t)p.num_tiles;
                const bool in_q = qi < (uint32_t)COMP_MAX_PUSHES * (uint32_t)p.num_tiles;
                volatile uint32_t *slot = p.queue + (in_q ? qi : 0u);
                volatile uint32_t *done = &p.frame->comp_done;
                uint32_t v = 0u;
                while (in_q && (v = *slot) == 0u && *done < (uint32_t)p.num_tiles) __nanosleep(200);
                if (in_q && v == 0u) v = *slot;
                s_tile = v ? v - 1u : EXIT_TILE;
                s_resume = 1u;
                __threadfence();
            }
            if (p.trace) t_start = globaltimer_ns();
        }
        __syncthreads();
        const uint32_t tile_id = s_tile;
        const bool resume = s_resume != 0u;
        if (tile_id == EXIT_TILE) break;

        const uint32_t tx = tile_id % (uint32_t)p.tiles_x, ty = tile_id / (uint32_t)p.tiles_x;
        const int px0 = (int)(tx * TILE + 2u * (tid & 7u)), py = (int)(ty * TILE + (tid >> 3));
        const u64 npx2 = pk(-(float)px0, -(float)(px0 + 1));
        const float fpy = (float)py;

        const uint2 bounds = p.bounds[tile_id];
        const int32_t diff = (int32_t)(bounds.y - bounds.x);
        const int num_splats = diff > 0 ? diff : 0;
        const int num_iterations = (int)ceilf((float)num_splats / (float)CHUNK);

        u64 cr2 = pk(0.f, 0.f), cg2 = cr2, cb2 = cr2;
        float t0 = 1.0f, t1 = 1.0f;
        int i0 = 0;
        const uint32_t rel = tile_id - (uint32_t)p.tile_begin;
        const uint32_t local_tile = (rel / (uint32_t)(p.row_step * p.tiles_x)) * (uint32_t)p.tiles_x + rel % (uint32_t)p.tiles_x;
        float4 *st = p.state + (uint64_t)local_tile * (2u * THREADS);
        if (resume) {
            const float4 sa = __ldcg(st + tid), sb = __ldcg(st + THREADS + tid);
            cr2 = pk(sa.x, sa.y); cg2 = pk(sa.z, sa.w); cb2 = pk(sb.x, sb.y);
            t0 = sb.z; t1 = sb.w;
            i0 = (int)__ldcg(p.state_chunk + local_tile);
        }

        auto load_ids = [&](int ci, uint32_t &v0, uint32_t &v1) {
            const int k0 = CHUNK * ci + (int)tid, k1 = k0 + THREADS;
            v0 = (ci < num_iterations && k0 < num_splats) ? __ldg(p.values + bounds.x + (uint32_t)k0) : 0xFFFFFFFFu;
            v1 = (ci < num_iterations && k1 < num_splats) ? __ldg(p.values + bounds.x + (uint32_t)k1) : 0xFFFFFFFFu;
        };
        auto issue = [&](int ci, uint32_t v0, uint32_t v1, int b) {   // TMA: raw records of the two splats -> this thread's slots of half b
            comp_fence_proxy_async();   // this thread's generic-proxy stores to its slots (pre-scale of an earlier chunk) precede the async writes
            if (tid == 0) {
                const int left = num_splats - CHUNK * ci;
                comp_mbar_expect_tx(&s_bar[b], 48u * (uint32_t)(left < CHUNK ? left : CHUNK));
            }
            if (v0 != 0xFFFFFFFFu) comp_bulk_g2s(&s_st[b][3 * tid], p.records + (uint64_t)v0 * 3u, 48u, &s_bar[b]);
            if (v1 != 0xFFFFFFFFu) comp_bulk_g2s(&s_st[b][3 * (tid + THREADS)], p.records + (uint64_t)v1 * 3u, 48u, &s_bar[b]);
        };
        auto finalize = [&](uint32_t v0, uint32_t v1, int b) {   // wait for the half, pre-scale the own two records in place like gather()
            if (b == 0) { comp_mbar_wait(&s_bar[0], phase0); phase0 ^= 1u; } else { comp_mbar_wait(&s_bar[1], phase1); phase1 ^= 1u; }
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                float4 *d = &s_st[b][3 * (tid + (uint32_t)hh * THREADS)];
                Staged sgd = null_splat();
                if ((hh ? v1 : v0) != 0xFFFFFFFFu) {
                    const float4 r0 = d[0], r1 = d[1], r2 = d[2];
                    sgd.a = make_float4(r0.x, r0.y, -0.5f * r1.x, -0.5f * r1.z);
                    sgd.b = make_float4(-r1.y, r2.w, r2.x, r2.y);
                    sgd.c = r2.z;
                }
                d[0] = sgd.a; d[1] = sgd.b; d[2] = make_float4(sgd.c, 0.f, 0.f, 0.f);
            }
        };

        // a resumed item means every fresh tile has been handed out (tickets are monotonic): yielding again would only cost a
        // spill + restore, so (requeue_only_if_fresh) it runs to completion; a fresh tile yields after `quantum` chunks
        const int q_min = p.quantum > (num_iterations + COMP_MAX_PUSHES) / (COMP_MAX_PUSHES + 1)
                              ? p.quantum : (num_iterations + COMP_MAX_PUSHES) / (COMP_MAX_PUSHES + 1);
        const int quantum = (p.requeue_only_if_fresh && resume) ? 0x3fffffff : q_min;
        const int i_begin = i0;
        bool finished = true;
        int b = 0;
        uint32_t va0, va1, vb0 = 0xFFFFFFFFu, vb1 = 0xFFFFFFFFu;   // ids of the chunk being fetched / of the one after it
        if (i0 < num_iterations) {
            load_ids(i0, va0, va1);
            issue(i0, va0, va1, 0);
            load_ids(i0 + 1, vb0, vb1);
            finalize(va0, va1, 0);
        }
        __syncthreads();   // staging half 0 visible (also orders the previous tile's last reads before this tile's writes)
        for (int i = i_begin; i < num_iterations; ++i) {
            const int sort_offset = CHUNK * i;
            const int chunk = (num_splats - sort_offset) < CHUNK ? (num_splats - sort_offset) : CHUNK;
            staged += (uint32_t)chunk;
            const bool fetch_next = (i + 1 < num_iterations) && (i + 1 - i_begin < quantum);
            if (fetch_next) {   // chunk i+1: records in flight during this blend, ids of chunk i+2 behind them
                va0 = vb0; va1 = vb1;
                issue(i + 1, va0, va1, b ^ 1);   // half b^1 was last read in blend(i-1): everybody is past that barrier
                load_ids(i + 2, vb0, vb1);
            }

            const int chunk4 = (chunk + GU - 1) & ~(GU - 1);
            if (PIPE) {
                // software-pipelined: the (a, b) words of group g+1 are loaded while group g is blended, and the warp's "anybody alive?"
                // test uses the transmittance after the first half of the group -- a dead warp may blend one more (fully masked,
                // exact no-op) group before it leaves, in exchange the vote and the shared-memory latency leave the loop-carried path
                float4 A[GU], B[GU];
#pragma unroll
                for (int u = 0; u < GU; ++u) { A[u] = s_st[b][3 * u]; B[u] = s_st[b][3 * u + 1]; }
                bool go = __any_sync(0xffffffffu, (t0 > MIN_ALPHA) || (t1 > MIN_ALPHA));
                for (int j = 0; j < chunk4 && go; j += GU) {
                    u64 al2[GU], om2[GU];
                    phase_a_v3<CVT>(A, B, npx2, fpy, K, al2, om2);
                    float4 Bc[GU];
#pragma unroll
                    for (int u = 0; u < GU; ++u) Bc[u] = B[u];
                    const int jn = (j + GU < CHUNK) ? j + GU : j;   // the last group re-reads itself instead of running off the half
#pragma unroll
                    for (int u = 0; u < GU; ++u) { A[u] = s_st[b][3 * (jn + u)]; B[u] = s_st[b][3 * (jn + u) + 1]; }
                    bool alive_mid = true;
                    phase_b_v3(s_st[b], j, Bc, al2, om2, cr2, cg2, cb2, t0, t1, alive_mid);
                    go = __any_sync(0xffffffffu, alive_mid);
                }
            } else {
                for (int j = 0; j < chunk4; j += GU) {
                    if (!__any_sync(0xffffffffu, (t0 > MIN_ALPHA) || (t1 > MIN_ALPHA))) break;
                    float4 A[GU], B[GU];
#pragma unroll
                    for (int u = 0; u < GU; ++u) { A[u] = s_st[b][3 * (j + u)]; B[u] = s_st[b][3 * (j + u) + 1]; }
                    u64 al2[GU], om2[GU];
                    bool alive_mid;
                    phase_a_v3<CVT>(A, B, npx2, fpy, K, al2, om2);
                    phase_b_v3(s_st[b], j, B, al2, om2, cr2, cg2, cb2, t0, t1, alive_mid);
                }
            }
            if (fetch_next) finalize(va0, va1, b ^ 1);   // nobody reads half b^1 before the barrier below

            const uint32_t wsum = __reduce_add_sync(0xffffffffu, (uint32_t)(t0 * 255.0f) + (uint32_t)(t1 * 255.0f));
            if (lane == 0) s_vote[i & 1][warp] = wsum;
            __syncthreads();   // the ONE barrier of the chunk: votes of chunk i and staging half b^1 (chunk i+1) visible
            uint32_t shared_t = 0;
#pragma unroll
            for (int w = 0; w < THREADS / 32; ++w) shared_t += s_vote[i & 1][w];
            if (!(shared_t > 255u)) break;
            if (i + 1 < num_iterations && i + 1 - i_begin >= quantum) {
                finished = false;
                i0 = i + 1;
                break;
            }
            b ^= 1;
        }

        float r0, r1, g0, g1, b0, b1;
        upk(cr2, r0, r1); upk(cg2, g0, g1); upk(cb2, b0, b1);
        if (!finished) {
            __stcg(st + tid, make_float4(r0, r1, g0, g1));
            __stcg(st + THREADS + tid, make_float4(b0, b1, t0, t1));
            if (tid == 0) __stcg(p.state_chunk + local_tile, (uint32_t)i0);
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                const uint32_t slot = atomicAdd(&p.frame->comp_tail, 1u);
                __threadfence();
                *(volatile uint32_t *)(p.queue + slot) = tile_id + 1u;
            }
        } else {
            const float hx = (float)num_splats * 5e-4f;  // :100-101
            const float h0 = 0.0f * (1.0f - hx) + 1.0f * hx, h1 = 0.0f * (1.0f - hx) + 0.2f * hx, h2c = 1.0f * (1.0f - hx) + 0.2f * hx;
            if (py < p.height) {
                float4 *row = p.out + (uint64_t)py * (uint64_t)p.width;
                const float k0 = 1.0f - t0, k1 = 1.0f - t1;
                if (px0 < p.width)
                    row[px0] = make_float4(r0 + h0 * k0 * p.heatmap_factor, g0 + h1 * k0 * p.heatmap_factor, b0 + h2c * k0 * p.heatmap_factor, 1.0f);
                if (px0 + 1 < p.width)
                    row[px0 + 1] = make_float4(r1 + h0 * k1 * p.heatmap_factor, g1 + h1 * k1 * p.heatmap_factor, b1 + h2c * k1 * p.heatmap_factor, 1.0f);
            }
            if ((tid & 15u) == 0u && tile_id == p.target_tile_id && t0 != 1.0f) {  // :105-110 pick
                const uint32_t v = p.values[bounds.x + (bounds.y - bounds.x) / 10u];
                const float4 q0 = p.records[(uint64_t)v * 3u + 0], q1 = p.records[(uint64_t)v * 3u + 1];
                *p.pick = make_float4(q0.z, q0.w, q1.w, (float)num_splats);
            }
            if (tid == 0) atomicAdd(&p.frame->comp_done, 1u);
        }
        if (p.trace && tid == 0) {
            const uint32_t k = atomicAdd(p.trace_count, 1u);
            if (k < p.trace_cap) p.trace[k] = make_ulonglong4(((unsigned long long)tile_id << 32) | smid(), t_start, globaltimer_ns(), ((unsigned long long)(uint32_t)i_begin << 32) | (uint32_t)(finished ? 1u : 0u) | ((uint32_t)num_iterations << 1));
        }
        __syncthreads();  // s_tile / staging buffers are reused by the next item
    }
    if (tid == 0 && staged && p.count_staged) atomicAdd(&p.frame->staged, (unsigned long long)staged);
}

// ==============================================================================================================
// EXPERIMENTAL "p4" variant (GSR_COMP_P4=1; off by default, NOT yet run on a GPU; logic checked by tests/test_kernel_emu.py).
// v2 staging, but 64 threads per tile and FOUR horizontally adjacent pixels (two packed pairs) per thread: the shared-memory
// loads and the row terms (oy, cz*oy) of a splat are shared by both pairs, and every splat offers two independent
// dependency chains, so a group of two splats gives ptxas the four chains the 2-pixel kernel needs four splats for.
// Same arithmetic per pixel, same vote, same queue => same results.
constexpr int P4_THREADS = 64;
constexpr int P4_GU = 2;
#ifndef GSR_COMP_P4_MIN_BLOCKS
#define GSR_COMP_P4_MIN_BLOCKS 8
#endif

// phase A for two pixel pairs (pixels px0,px0+1 | px0+2,px0+3 of one row) and P4_GU splats, stage by stage
// staging layout of the p4 kernel: ONE array per half, slot k = float4[3k] (a), float4[3k+1] (b), float4[3k+2].x (c) -- the
// 48-byte record lands there raw (cp.async) and is pre-scaled in place by the thread that fetched it
__device__ __forceinline__ void phase_a_p4(const float4 *s, int j, u64 npxA, u64 npxB, float fpy, const BlendK &K, u64 al2[P4_GU][2]) {
    float4 A[P4_GU];
    float bx[P4_GU], by[P4_GU], oy[P4_GU], czoy[P4_GU];
    u64 ox2[P4_GU][2], pw2[P4_GU][2], tm2[P4_GU][2], e2[P4_GU][2];
#pragma unroll
    for (int u = 0; u < P4_GU; ++u) { A[u] = s[3 * (j + u)]; const float4 b = s[3 * (j + u) + 1]; bx[u] = b.x; by[u] = b.y; }
#pragma unroll
    for (int u = 0; u < P4_GU; ++u) {
        ox2[u][0] = add2(bc(A[u].x), npxA); ox2[u][1] = add2(bc(A[u].x), npxB);
        oy[u] = A[u].y - fpy; czoy[u] = A[u].w * oy[u];
    }
#define P4_EACH for (int u = 0; u < P4_GU; ++u) for (int q = 0; q < 2; ++q)
#pragma unroll
    P4_EACH pw2[u][q] = mul2(bc(A[u].z), ox2[u][q]);
#pragma unroll
    P4_EACH pw2[u][q] = mul2(pw2[u][q], ox2[u][q]);
#pragma unroll
    P4_EACH pw2[u][q] = fma2(bc(czoy[u]), bc(oy[u]), pw2[u][q]);
#pragma unroll
    P4_EACH e2[u][q] = mul2(bc(bx[u]), ox2[u][q]);
#pragma unroll
    P4_EACH pw2[u][q] = fma2(e2[u][q], bc(oy[u]), pw2[u][q]);
#pragma unroll
    P4_EACH pw2[u][q] = mul2(pw2[u][q], K.L2E2);
#pragma unroll
    P4_EACH {
        float tl, th;
        upk(pw2[u][q], tl, th);
        tl = g_min(g_max(tl, -127.0f), 128.0f);
        th = g_min(g_max(th, -127.0f), 128.0f);
        pw2[u][q] = pk(tl, th);
    }
#pragma unroll
    P4_EACH tm2[u][q] = add2(pw2[u][q], K.MAGIC2);
#pragma unroll
    P4_EACH al2[u][q] = sub2(tm2[u][q], K.MAGIC2);
#pragma unroll
    P4_EACH pw2[u][q] = sub2(pw2[u][q], al2[u][q]);  // f
#pragma unroll
    P4_EACH e2[u][q] = fma2(K.C6, pw2[u][q], K.C5);
#pragma unroll
    P4_EACH e2[u][q] = fma2(e2[u][q], pw2[u][q], K.C4);
#pragma unroll
    P4_EACH e2[u][q] = fma2(e2[u][q], pw2[u][q], K.C3);
#pragma unroll
    P4_EACH e2[u][q] = fma2(e2[u][q], pw2[u][q], K.C2);
#pragma unroll
    P4_EACH e2[u][q] = fma2(e2[u][q], pw2[u][q], K.C1);
#pragma unroll
    P4_EACH e2[u][q] = fma2(e2[u][q], pw2[u][q], K.ONE2);
#pragma unroll
    P4_EACH {
        float ml, mh;
        upk(tm2[u][q], ml, mh);
        tm2[u][q] = pk(__uint_as_float((__float_as_uint(ml) << 23) + 0x3F800000u), __uint_as_float((__float_as_uint(mh) << 23) + 0x3F800000u));
    }
#pragma unroll
    P4_EACH e2[u][q] = mul2(e2[u][q], tm2[u][q]);
#pragma unroll
    P4_EACH al2[u][q] = mul2(bc(by[u]), e2[u][q]);
#undef P4_EACH
}

__device__ __forceinline__ void phase_b_p4(const float4 *s, int j, const u64 al2[P4_GU][2], const BlendK &K, u64 cr2[2], u64 cg2[2], u64 cb2[2],
                                           float t[4]) {
#pragma unroll
    for (int u = 0; u < P4_GU; ++u) {
        const float4 b = s[3 * (j + u) + 1];
        const float cbl = s[3 * (j + u) + 2].x;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            float al, ah;
            upk(al2[u][q], al, ah);
            al = (t[2 * q] > MIN_ALPHA) ? al : 0.0f;
            ah = (t[2 * q + 1] > MIN_ALPHA) ? ah : 0.0f;
            const u64 m2 = pk(al, ah);
            const u64 t2 = pk(t[2 * q], t[2 * q + 1]);
            cr2[q] = fma2(mul2(bc(b.z), m2), t2, cr2[q]);
            cg2[q] = fma2(mul2(bc(b.w), m2), t2, cg2[q]);
            cb2[q] = fma2(mul2(bc(cbl), m2), t2, cb2[q]);
            upk(mul2(t2, sub2(K.ONE2, m2)), t[2 * q], t[2 * q + 1]);
        }
    }
}

__global__ void __launch_bounds__(P4_THREADS, GSR_COMP_P4_MIN_BLOCKS) composite_p4_kernel(const __grid_constant__ CompositeArgs p) {
    __shared__ float4 s_st[2][CHUNK * 3];   // two staging halves, 12 KB each (see phase_a_p4)
    __shared__ uint32_t s_vote[2][P4_THREADS / 32];
    __shared__ uint32_t s_tile, s_resume;
    constexpr int SL = CHUNK / P4_THREADS;   // staging slots per thread (4): slot tid + k*64

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const BlendK K = make_blend_k();
    uint32_t staged = 0;
    unsigned long long t_start = 0;

    for (;;) {
        if (tid == 0) {
            const uint32_t ticket = atomicAdd(&p.frame->comp_head, 1u);
            if (ticket < (uint32_t)p.num_tiles) {
                const uint32_t k = p.order ? p.order[ticket] : ticket;   // owned-tile index: longest lists first when an order is given
                s_tile = (uint32_t)p.tile_begin + (k / (uint32_t)p.tiles_x) * (uint32_t)(p.row_step * p.tiles_x) + k % (uint32_t)p.tiles_x;
                s_resume = 0u;
            } else {
                // a tile is pushed at most COMP_MAX_PUSHES times: later tickets can never be served and must not touch the queue
                const uint32_t qi = ticket - (uint32_t)p.num_tiles;
                const bool in_q = qi < (uint32_t)COMP_MAX_PUSHES * (uint32_t)p.num_tiles;
                volatile uint32_t *slot = p.queue + (in_q ? qi : 0u);
                volatile uint32_t *done = &p.frame->comp_done;
                uint32_t v = 0u;
                while (in_q && (v = *slot) == 0u && *done < (uint32_t)p.num_tiles) __nanosleep(200);
                if (in_q && v == 0u) v = *slot;
                s_tile = v ? v - 1u : EXIT_TILE;
                s_resume = 1u;
                __threadfence();
            }
            if (p.trace) t_start = globaltimer_ns();
        }
        __syncthreads();
        const uint32_t tile_id = s_tile;
        const bool resume = s_resume != 0u;
        if (tile_id == EXIT_TILE) break;

        const uint32_t tx = tile_id % (uint32_t)p.tiles_x, ty = tile_id / (uint32_t)p.tiles_x;
        const int px0 = (int)(tx * TILE + 4u * (tid & 3u)), py = (int)(ty * TILE + (tid >> 2));
        const u64 npxA = pk(-(float)px0, -(float)(px0 + 1)), npxB = pk(-(float)(px0 + 2), -(float)(px0 + 3));
        const float fpy = (float)py;

        const uint2 bounds = p.bounds[tile_id];
        const int32_t diff = (int32_t)(bounds.y - bounds.x);
        const int num_splats = diff > 0 ? diff : 0;
        const int num_iterations = (int)ceilf((float)num_splats / (float)CHUNK);

        u64 cr2[2] = {pk(0.f, 0.f), pk(0.f, 0.f)}, cg2[2] = {cr2[0], cr2[0]}, cb2[2] = {cr2[0], cr2[0]};
        float t[4] = {1.0f, 1.0f, 1.0f, 1.0f};
        int i0 = 0;
        const uint32_t rel = tile_id - (uint32_t)p.tile_begin;
        const uint32_t local_tile = (rel / (uint32_t)(p.row_step * p.tiles_x)) * (uint32_t)p.tiles_x + rel % (uint32_t)p.tiles_x;
        float4 *st = p.state + (uint64_t)local_tile * 256u;   // 4 float4 per thread: r[4], g[4], b[4], t[4]
        if (resume) {
            const float4 sr = __ldcg(st + tid), sg = __ldcg(st + 64 + tid), sb = __ldcg(st + 128 + tid), stt = __ldcg(st + 192 + tid);
            cr2[0] = pk(sr.x, sr.y); cr2[1] = pk(sr.z, sr.w);
            cg2[0] = pk(sg.x, sg.y); cg2[1] = pk(sg.z, sg.w);
            cb2[0] = pk(sb.x, sb.y); cb2[1] = pk(sb.z, sb.w);
            t[0] = stt.x; t[1] = stt.y; t[2] = stt.z; t[3] = stt.w;
            i0 = (int)__ldcg(p.state_chunk + local_tile);
        }

        auto load_ids = [&](int ci, uint32_t v[SL]) {
#pragma unroll
            for (int k = 0; k < SL; ++k) {
                const int idx = CHUNK * ci + (int)tid + k * P4_THREADS;
                v[k] = (ci < num_iterations && idx < num_splats) ? __ldg(p.values + bounds.x + (uint32_t)idx) : 0xFFFFFFFFu;
            }
        };
        auto issue = [&](const uint32_t v[SL], int b) {   // raw records straight into the half that is not being blended
#pragma unroll
            for (int k = 0; k < SL; ++k)
                if (v[k] != 0xFFFFFFFFu) {
                    const float4 *r = p.records + (uint64_t)v[k] * 3u;
                    float4 *d = &s_st[b][3 * (tid + (uint32_t)k * P4_THREADS)];
                    cp_async16(d + 0, r + 0); cp_async16(d + 1, r + 1); cp_async16(d + 2, r + 2);
                }
        };
        auto finalize = [&](const uint32_t v[SL], int b) {   // own slots: wait, pre-scale in place (gather()'s layout)
            cp_async_commit_wait_all();
#pragma unroll
            for (int k = 0; k < SL; ++k) {
                float4 *d = &s_st[b][3 * (tid + (uint32_t)k * P4_THREADS)];
                Staged sgd = null_splat();
                if (v[k] != 0xFFFFFFFFu) {
                    const float4 r0 = d[0], r1 = d[1], r2 = d[2];
                    sgd.a = make_float4(r0.x, r0.y, -0.5f * r1.x, -0.5f * r1.z);
                    sgd.b = make_float4(-r1.y, r2.w, r2.x, r2.y);
                    sgd.c = r2.z;
                }
                d[0] = sgd.a; d[1] = sgd.b; d[2] = make_float4(sgd.c, 0.f, 0.f, 0.f);
            }
        };

        // a resumed item means every fresh tile has been handed out (tickets are monotonic): yielding again would only cost a
        // spill + restore, so (requeue_only_if_fresh) it runs to completion; a fresh tile yields after `quantum` chunks
        const int q_min = p.quantum > (num_iterations + COMP_MAX_PUSHES) / (COMP_MAX_PUSHES + 1)
                              ? p.quantum : (num_iterations + COMP_MAX_PUSHES) / (COMP_MAX_PUSHES + 1);
        const int quantum = (p.requeue_only_if_fresh && resume) ? 0x3fffffff : q_min;
        const int i_begin = i0;
        bool finished = true;
        int b = 0;
        uint32_t va[SL], vb[SL];
#pragma unroll
        for (int k = 0; k < SL; ++k) va[k] = vb[k] = 0xFFFFFFFFu;
        if (i0 < num_iterations) {
            load_ids(i0, va);
            issue(va, 0);
            load_ids(i0 + 1, vb);
            finalize(va, 0);
        }
        __syncthreads();
        for (int i = i_begin; i < num_iterations; ++i) {
            const int sort_offset = CHUNK * i;
            const int chunk = (num_splats - sort_offset) < CHUNK ? (num_splats - sort_offset) : CHUNK;
            staged += (uint32_t)chunk;
            const bool fetch_next = (i + 1 < num_iterations) && (i + 1 - i_begin < quantum);
            if (fetch_next) {
#pragma unroll
                for (int k = 0; k < SL; ++k) va[k] = vb[k];
                issue(va, b ^ 1);   // half b^1 was last read in blend(i-1): everybody is past that barrier
                load_ids(i + 2, vb);
            }

            const int chunkg = (chunk + P4_GU - 1) & ~(P4_GU - 1);   // null splats (opacity 0) pad the group: exact no-ops
            for (int j = 0; j < chunkg; j += P4_GU) {
                if (!__any_sync(0xffffffffu, (t[0] > MIN_ALPHA) || (t[1] > MIN_ALPHA) || (t[2] > MIN_ALPHA) || (t[3] > MIN_ALPHA))) break;
                u64 al2[P4_GU][2];
                phase_a_p4(s_st[b], j, npxA, npxB, fpy, K, al2);
                phase_b_p4(s_st[b], j, al2, K, cr2, cg2, cb2, t);
            }
            if (fetch_next) finalize(va, b ^ 1);

            const uint32_t wsum = __reduce_add_sync(0xffffffffu, (uint32_t)(t[0] * 255.0f) + (uint32_t)(t[1] * 255.0f) +
                                                                    (uint32_t)(t[2] * 255.0f) + (uint32_t)(t[3] * 255.0f));
            if (lane == 0) s_vote[i & 1][warp] = wsum;
            __syncthreads();
            uint32_t shared_t = 0;
#pragma unroll
            for (int w = 0; w < P4_THREADS / 32; ++w) shared_t += s_vote[i & 1][w];
            if (!(shared_t > 255u)) break;
            if (i + 1 < num_iterations && i + 1 - i_begin >= quantum) {
                finished = false;
                i0 = i + 1;
                break;
            }
            b ^= 1;
        }

        float r[4], g[4], bl[4];
        upk(cr2[0], r[0], r[1]); upk(cr2[1], r[2], r[3]);
        upk(cg2[0], g[0], g[1]); upk(cg2[1], g[2], g[3]);
        upk(cb2[0], bl[0], bl[1]); upk(cb2[1], bl[2], bl[3]);
        if (!finished) {
            __stcg(st + tid, make_float4(r[0], r[1], r[2], r[3]));
            __stcg(st + 64 + tid, make_float4(g[0], g[1], g[2], g[3]));
            __stcg(st + 128 + tid, make_float4(bl[0], bl[1], bl[2], bl[3]));
            __stcg(st + 192 + tid, make_float4(t[0], t[1], t[2], t[3]));
            if (tid == 0) __stcg(p.state_chunk + local_tile, (uint32_t)i0);
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                const uint32_t slot = atomicAdd(&p.frame->comp_tail, 1u);
                __threadfence();
                *(volatile uint32_t *)(p.queue + slot) = tile_id + 1u;
            }
        } else {
            const float hx = (float)num_splats * 5e-4f;  // :100-101
            const float h0 = 0.0f * (1.0f - hx) + 1.0f * hx, h1 = 0.0f * (1.0f - hx) + 0.2f * hx, h2c = 1.0f * (1.0f - hx) + 0.2f * hx;
            if (py < p.height) {
                float4 *row = p.out + (uint64_t)py * (uint64_t)p.width;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float kk = 1.0f - t[k];
                    if (px0 + k < p.width)
                        row[px0 + k] = make_float4(r[k] + h0 * kk * p.heatmap_factor, g[k] + h1 * kk * p.heatmap_factor, bl[k] + h2c * kk * p.heatmap_factor, 1.0f);
                }
            }
            // :105-110 pick: the elected lanes of the reference's 8 subgroups are the pixels (0, 2s): even rows, first column
            if ((tid & 7u) == 0u && tile_id == p.target_tile_id && t[0] != 1.0f) {
                const uint32_t v = p.values[bounds.x + (bounds.y - bounds.x) / 10u];
                const float4 q0 = p.records[(uint64_t)v * 3u + 0], q1 = p.records[(uint64_t)v * 3u + 1];
                *p.pick = make_float4(q0.z, q0.w, q1.w, (float)num_splats);
            }
            if (tid == 0) atomicAdd(&p.frame->comp_done, 1u);
        }
        if (p.trace && tid == 0) {
            const uint32_t k = atomicAdd(p.trace_count, 1u);
            if (k < p.trace_cap) p.trace[k] = make_ulonglong4(((unsigned long long)tile_id << 32) | smid(), t_start, globaltimer_ns(), ((unsigned long long)(uint32_t)i_begin << 32) | (uint32_t)(finished ? 1u : 0u) | ((uint32_t)num_iterations << 1));
        }
        __syncthreads();
    }
    if (tid == 0 && staged && p.count_staged) atomicAdd(&p.frame->staged, (unsigned long long)staged);
}

// ==============================================================================================================
// EXPERIMENTAL warp-specialised variant (GSR_COMP_WS=1; off by default).  Measured on B200 (c3): 0.77 ms vs 0.54 ms for
// composite_kernel -- the per-4-splat hand-off through shared memory (poll, 4 x 64-bit loads, two warp syncs) costs more
// than the instruction-level parallelism it buys, both in the saturated phase (-30 % throughput) and for a lone tile
// (47 us vs 34 us per 512 splats).  Kept for the next round: a coarser hand-off (16+ splats) is the obvious follow-up.
// Same tiles, same queue, same arithmetic -- but each 64-pixel group of a tile is served by
// TWO warps: a blend warp that owns the pixels' state and an alpha warp that runs phase A ahead of it and hands the
// packed alphas over through a shared-memory ring.  Phase A does not depend on the transmittance, so it parallelises
// over splats; only phase B is sequential.  Of every three 4-splat groups the alpha warp computes two and the blend warp
// one (plus all three blends): 152 vs 172 FMA-pipe operations, i.e. a lone tile advances ~1.9x faster, which is what
// bounds the kernel's tail (and all of it when a GPU owns only a slice of the frame).
constexpr int WS_THREADS = 256;  // warps 0-3: blend, warps 4-7: alpha; pixel group g = warp & 3
constexpr int RING_D = 4;        // ring depth in 4-splat groups per pixel group

__global__ void __launch_bounds__(WS_THREADS, 2) composite_ws_kernel(const __grid_constant__ CompositeArgs p) {
    __shared__ float4 s_a[CHUNK];
    __shared__ float4 s_b[CHUNK];
    __shared__ float s_c[CHUNK];
    __shared__ u64 s_ring[4][RING_D][GU][32];
    __shared__ uint32_t s_prod[4], s_cons[4], s_stop[4];
    __shared__ uint32_t s_vote[4];
    __shared__ uint32_t s_tile, s_resume;

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const bool is_blend = warp < 4u;
    const uint32_t g = warp & 3u;
    const uint32_t ptid = g * 32u + lane;  // pixel-pair index 0..127 inside the tile (same mapping as composite_kernel)
    const BlendK K = make_blend_k();
    uint32_t staged = 0;
    unsigned long long t_start = 0;

    for (;;) {
        if (tid == 0) {
            const uint32_t ticket = atomicAdd(&p.frame->comp_head, 1u);
            if (ticket < (uint32_t)p.num_tiles) {
                const uint32_t k = p.order ? p.order[ticket] : ticket;   // owned-tile index: longest lists first when an order is given
                s_tile = (uint32_t)p.tile_begin + (k / (uint32_t)p.tiles_x) * (uint32_t)(p.row_step * p.tiles_x) + k % (uint32_t)p.tiles_x;
                s_resume = 0u;
            } else {
                // a tile is pushed at most COMP_MAX_PUSHES times: later tickets can never be served and must not touch the queue
                const uint32_t qi = ticket - (uint32_t)p.num_tiles;
                const bool in_q = qi < (uint32_t)COMP_MAX_PUSHES * (uint32_t)p.num_tiles;
                volatile uint32_t *slot = p.queue + (in_q ? qi : 0u);
                volatile uint32_t *done = &p.frame->comp_done;
                uint32_t v = 0u;
                while (in_q && (v = *slot) == 0u && *done < (uint32_t)p.num_tiles) __nanosleep(200);
                if (in_q && v == 0u) v = *slot;
                s_tile = v ? v - 1u : EXIT_TILE;
                s_resume = 1u;
                __threadfence();
            }
            if (p.trace) t_start = globaltimer_ns();
        }
        __syncthreads();
        const uint32_t tile_id = s_tile;
        const bool resume = s_resume != 0u;
        if (tile_id == EXIT_TILE) break;

        const uint32_t tx = tile_id % (uint32_t)p.tiles_x, ty = tile_id / (uint32_t)p.tiles_x;
        const int px0 = (int)(tx * TILE + 2u * (ptid & 7u)), py = (int)(ty * TILE + (ptid >> 3));
        const u64 npx2 = pk(-(float)px0, -(float)(px0 + 1));
        const float fpy = (float)py;

        const uint2 bounds = p.bounds[tile_id];
        const int32_t diff = (int32_t)(bounds.y - bounds.x);
        const int num_splats = diff > 0 ? diff : 0;                              // :61
        const int num_iterations = (int)ceilf((float)num_splats / (float)CHUNK);  // :62

        u64 cr2 = pk(0.f, 0.f), cg2 = cr2, cb2 = cr2;
        float t0 = 1.0f, t1 = 1.0f;
        int i0 = 0;
        const uint32_t rel = tile_id - (uint32_t)p.tile_begin;
        const uint32_t local_tile = (rel / (uint32_t)(p.row_step * p.tiles_x)) * (uint32_t)p.tiles_x + rel % (uint32_t)p.tiles_x;
        float4 *st = p.state + (uint64_t)local_tile * (2u * THREADS);
        if (resume) {
            if (is_blend) {
                const float4 sa = __ldcg(st + ptid), sb = __ldcg(st + THREADS + ptid);
                cr2 = pk(sa.x, sa.y); cg2 = pk(sa.z, sa.w); cb2 = pk(sb.x, sb.y);
                t0 = sb.z; t1 = sb.w;
            }
            i0 = (int)__ldcg(p.state_chunk + local_tile);
        }

        Staged n0 = null_splat();  // 256 threads stage one record each
        if (i0 < num_iterations && CHUNK * i0 + (int)tid < num_splats) n0 = gather(p.records, p.values, bounds.x + (uint32_t)(CHUNK * i0) + tid);

        // a resumed item means every fresh tile has been handed out (tickets are monotonic): yielding again would only cost a
        // spill + restore, so (requeue_only_if_fresh) it runs to completion; a fresh tile yields after `quantum` chunks
        const int q_min = p.quantum > (num_iterations + COMP_MAX_PUSHES) / (COMP_MAX_PUSHES + 1)
                              ? p.quantum : (num_iterations + COMP_MAX_PUSHES) / (COMP_MAX_PUSHES + 1);
        const int quantum = (p.requeue_only_if_fresh && resume) ? 0x3fffffff : q_min;
        const int i_begin = i0;
        bool finished = true;
        for (int i = i_begin; i < num_iterations; ++i) {
            const int sort_offset = CHUNK * i;
            const int chunk = (num_splats - sort_offset) < CHUNK ? (num_splats - sort_offset) : CHUNK;
            staged += (uint32_t)chunk;
            s_a[tid] = n0.a; s_b[tid] = n0.b; s_c[tid] = n0.c;
            if (lane == 0) {
                if (is_blend) { s_cons[g] = 0u; s_stop[g] = 0u; }
                else s_prod[g] = 0u;
            }
            __syncthreads();
            n0 = null_splat();
            if (i + 1 < num_iterations && i + 1 - i_begin < quantum && sort_offset + CHUNK + (int)tid < num_splats)
                n0 = gather(p.records, p.values, bounds.x + (uint32_t)(sort_offset + CHUNK) + tid);

            const int ngroups = ((chunk + GU - 1) & ~(GU - 1)) / GU;
            if (is_blend) {
                uint32_t jq = 0;  // alpha-warp groups consumed so far
                for (int k = 0; k < ngroups; ++k) {
                    if (!__any_sync(0xffffffffu, (t0 > MIN_ALPHA) || (t1 > MIN_ALPHA))) {
                        if (lane == 0) *(volatile uint32_t *)&s_stop[g] = 1u;  // tell the alpha warp to stop producing
                        break;
                    }
                    u64 al2[GU];
                    if (k % 3 == 0) {
                        phase_a<false>(s_a, s_b, k * GU, npx2, fpy, K, al2);
                    } else {
                        if (lane == 0) { while (*(volatile uint32_t *)&s_prod[g] <= jq) {} }
                        __syncwarp();
#pragma unroll
                        for (int u = 0; u < GU; ++u) al2[u] = *(volatile u64 *)&s_ring[g][jq % RING_D][u][lane];
                        __syncwarp();
                        ++jq;
                        if (lane == 0) *(volatile uint32_t *)&s_cons[g] = jq;
                    }
                    phase_b(s_b, s_c, k * GU, al2, K, cr2, cg2, cb2, t0, t1);
                }
            } else {
                uint32_t jq = 0;  // groups produced so far
                for (int k = 0; k < ngroups; ++k) {
                    if (k % 3 == 0) continue;
                    uint32_t stop = 0u;
                    if (lane == 0) {
                        while (jq - *(volatile uint32_t *)&s_cons[g] >= (uint32_t)RING_D && *(volatile uint32_t *)&s_stop[g] == 0u) {}
                        stop = *(volatile uint32_t *)&s_stop[g];
                    }
                    stop = __shfl_sync(0xffffffffu, stop, 0);
                    if (stop) break;
                    u64 al2[GU];
                    phase_a<false>(s_a, s_b, k * GU, npx2, fpy, K, al2);
#pragma unroll
                    for (int u = 0; u < GU; ++u) s_ring[g][jq % RING_D][u][lane] = al2[u];
                    __syncwarp();
                    ++jq;
                    if (lane == 0) { __threadfence_block(); *(volatile uint32_t *)&s_prod[g] = jq; }
                }
            }

            // :97 tile-stop vote over the tile's 256 pixels (owned by the four blend warps)
            if (is_blend) {
                const uint32_t wsum = __reduce_add_sync(0xffffffffu, (uint32_t)(t0 * 255.0f) + (uint32_t)(t1 * 255.0f));
                if (lane == 0) s_vote[g] = wsum;
            }
            __syncthreads();
            const uint32_t shared_t = s_vote[0] + s_vote[1] + s_vote[2] + s_vote[3];
            if (!(shared_t > 255u)) break;
            if (i + 1 < num_iterations && i + 1 - i_begin >= quantum) {
                finished = false;
                i0 = i + 1;
                break;
            }
        }

        float r0, r1, g0, g1, b0, b1;
        upk(cr2, r0, r1); upk(cg2, g0, g1); upk(cb2, b0, b1);
        if (!finished) {
            if (is_blend) {
                __stcg(st + ptid, make_float4(r0, r1, g0, g1));
                __stcg(st + THREADS + ptid, make_float4(b0, b1, t0, t1));
            }
            if (tid == 0) __stcg(p.state_chunk + local_tile, (uint32_t)i0);
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                const uint32_t slot = atomicAdd(&p.frame->comp_tail, 1u);
                __threadfence();
                *(volatile uint32_t *)(p.queue + slot) = tile_id + 1u;
            }
        } else {
            if (is_blend) {
                const float hx = (float)num_splats * 5e-4f;  // :100-101
                const float h0 = 0.0f * (1.0f - hx) + 1.0f * hx, h1 = 0.0f * (1.0f - hx) + 0.2f * hx, h2c = 1.0f * (1.0f - hx) + 0.2f * hx;
                if (py < p.height) {
                    float4 *row = p.out + (uint64_t)py * (uint64_t)p.width;
                    const float k0 = 1.0f - t0, k1 = 1.0f - t1;
                    if (px0 < p.width)
                        row[px0] = make_float4(r0 + h0 * k0 * p.heatmap_factor, g0 + h1 * k0 * p.heatmap_factor, b0 + h2c * k0 * p.heatmap_factor, 1.0f);
                    if (px0 + 1 < p.width)
                        row[px0 + 1] = make_float4(r1 + h0 * k1 * p.heatmap_factor, g1 + h1 * k1 * p.heatmap_factor, b1 + h2c * k1 * p.heatmap_factor, 1.0f);
                }
                if ((ptid & 15u) == 0u && tile_id == p.target_tile_id && t0 != 1.0f) {  // :105-110 pick
                    const uint32_t v = p.values[bounds.x + (bounds.y - bounds.x) / 10u];
                    const float4 q0 = p.records[(uint64_t)v * 3u + 0], q1 = p.records[(uint64_t)v * 3u + 1];
                    *p.pick = make_float4(q0.z, q0.w, q1.w, (float)num_splats);
                }
            }
            if (tid == 0) atomicAdd(&p.frame->comp_done, 1u);
        }
        if (p.trace && tid == 0) {
            const uint32_t k = atomicAdd(p.trace_count, 1u);
            if (k < p.trace_cap) p.trace[k] = make_ulonglong4(((unsigned long long)tile_id << 32) | smid(), t_start, globaltimer_ns(), ((unsigned long long)(uint32_t)i_begin << 32) | (uint32_t)(finished ? 1u : 0u) | ((uint32_t)num_iterations << 1));
        }
        __syncthreads();
    }
    if (tid == 0 && staged && p.count_staged) atomicAdd(&p.frame->staged, (unsigned long long)staged);
}

}  // namespace

#ifndef GSR_CPU_EMU
int preload_composite_kernels() {
    cudaFuncAttributes fa;
    GSR_CUDA_TRY(cudaFuncGetAttributes(&fa, composite_kernel<false>));
    GSR_CUDA_TRY(cudaFuncGetAttributes(&fa, composite_v3_kernel<6, false, false>));
    return GSR_OK;
}
int launch_composite(const CompositeArgs &a, cudaStream_t stream) {
    if (a.num_tiles <= 0) return GSR_OK;
    static int ctas_per_sm = 0, sms = 0, use_ws = GSR_COMP_WS_DEFAULT, use_hwexp = 0, use_v2 = 0, use_p4 = 0, use_v3 = 0, use_cvt = 0, use_pipe = 0, cfg_dev = -1;
    int dev = 0;
    GSR_CUDA_TRY(cudaGetDevice(&dev));
    if (cfg_dev != dev) {
        cfg_dev = dev;
        GSR_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const char *w = getenv("GSR_COMP_WS");  // experiment knob: 1 = warp-specialised kernel, 0 = plain
        if (w) use_ws = atoi(w) != 0;
        const char *h = getenv("GSR_COMP_HWEXP");  // experiment knob: 1 = exp() on the SFU (not bit-reproducible; see phase_a)
        use_hwexp = (h && atoi(h) != 0) ? 1 : 0;
        const char *v2e = getenv("GSR_COMP_V2");  // experiment knob: cp.async staging, one barrier per chunk (bit-identical results);
        use_v2 = (v2e && atoi(v2e) != 0 && !use_ws && !use_hwexp) ? atoi(v2e) : 0;   // value = CTAs/SM target: 1|5 -> 5, 6, 8
        const char *p4e = getenv("GSR_COMP_P4");  // experiment knob: 1 = four pixels per thread on top of the v2 staging (bit-identical results)
        use_p4 = (p4e && atoi(p4e) != 0 && !use_ws && !use_hwexp) ? 1 : 0;
        if (use_p4) use_v2 = 0;
        const char *v3e = getenv("GSR_COMP_V3");   // experiment knob: TMA staging + short transmittance chain; value = CTAs/SM target (6 or 8)
        use_v3 = (v3e && atoi(v3e) != 0 && !use_ws && !use_hwexp && !use_p4) ? atoi(v3e) : 0;
        const char *cvte = getenv("GSR_COMP_CVT");
        use_cvt = (cvte && atoi(cvte) != 0) ? 1 : 0;
        if (use_v3) use_v2 = 0;
        const char *pipee = getenv("GSR_COMP_PIPE");
        use_pipe = (pipee && atoi(pipee) != 0) ? 1 : 0;
        if (use_v3 == 4 && use_pipe) GSR_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, composite_v3_kernel<4, false, true>, THREADS, 0));
        else if (use_v3 == 4) GSR_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, composite_v3_kernel<4, false, false>, THREADS, 0));
        else if (use_v3 >= 8 && use_cvt) GSR_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, composite_v3_kernel<8, true, false>, THREADS, 0));
        else if (use_v3 >= 8) GSR_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, composite_v3_kernel<8, false, false>, THREADS, 0));
        else if (use_v3 && use_cvt) GSR_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, composite_v3_kernel<6, true, false>, THREADS, 0));
        else if (use_v3) GSR_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, composite_v3_kernel<6, false, false>, THREADS, 0));
        else if (use_p4) GSR_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, composite_p4_kernel, P4_THREADS, 0));
        else if (use_v2 >= 8) GSR_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, composite_v2_kernel<8>, THREADS, 0));
        else if (use_v2 >= 6) GSR_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, composite_v2_kernel<6>, THREADS, 0));
        else if (use_v2) GSR_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, composite_v2_kernel<5>, THREADS, 0));
        else if (use_ws) GSR_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, composite_ws_kernel, WS_THREADS, 0));
        else if (use_hwexp) GSR_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, composite_kernel<true>, THREADS, 0));
        else GSR_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, composite_kernel<false>, THREADS, 0));
        if (ctas_per_sm < 1) ctas_per_sm = 1;
        const char *e = getenv("GSR_COMP_CTAS_PER_SM");  // experiment knob
        if (e && atoi(e) > 0 && atoi(e) < ctas_per_sm) ctas_per_sm = atoi(e);
    }
    const int per_sm = (a.ctas_per_sm > 0 && a.ctas_per_sm < ctas_per_sm) ? a.ctas_per_sm : ctas_per_sm;
    const int grid = a.num_tiles < sms * per_sm ? a.num_tiles : sms * per_sm;
    if (use_v3 == 4 && use_pipe) composite_v3_kernel<4, false, true><<<grid, THREADS, 0, stream>>>(a);
    else if (use_v3 == 4) composite_v3_kernel<4, false, false><<<grid, THREADS, 0, stream>>>(a);
    else if (use_v3 >= 8 && use_cvt) composite_v3_kernel<8, true, false><<<grid, THREADS, 0, stream>>>(a);
    else if (use_v3 >= 8) composite_v3_kernel<8, false, false><<<grid, THREADS, 0, stream>>>(a);
    else if (use_v3 && use_cvt) composite_v3_kernel<6, true, false><<<grid, THREADS, 0, stream>>>(a);
    else if (use_v3) composite_v3_kernel<6, false, false><<<grid, THREADS, 0, stream>>>(a);
    else if (use_p4) composite_p4_kernel<<<grid, P4_THREADS, 0, stream>>>(a);
    else if (use_v2 >= 8) composite_v2_kernel<8><<<grid, THREADS, 0, stream>>>(a);
    else if (use_v2 >= 6) composite_v2_kernel<6><<<grid, THREADS, 0, stream>>>(a);
    else if (use_v2) composite_v2_kernel<5><<<grid, THREADS, 0, stream>>>(a);
    else if (use_ws) composite_ws_kernel<<<grid, WS_THREADS, 0, stream>>>(a);
    else if (use_hwexp) composite_kernel<true><<<grid, THREADS, 0, stream>>>(a);
    else composite_kernel<false><<<grid, THREADS, 0, stream>>>(a);
    GSR_CUDA_TRY(cudaGetLastError());
    return GSR_OK;
}
#endif  // GSR_CPU_EMU

}  // namespace gsr
