// kernel_emu.cpp -- libgsr's projection, sort, tile-range and compositor kernels (csrc/projection.cu, radix_sort.cu, ranges.cu,
// compositor.cu) compiled for the CPU.
// TEST INFRASTRUCTURE: see cuda_shim.h.  Built by tests/kernel_emu/build.py into tests/kernel_emu/libkernel_emu.so.
#define GSR_CPU_EMU 1
#include <vector>

#include "cuda_shim.h"

namespace cuda_emu { dim g_block_dim{128, 1, 1}, g_grid_dim{1, 1, 1}; }
namespace gsr { void set_last_error(const char *, ...) {} }

#include "../../godotgaussiansplatting_b200/csrc/compositor.cu"
#include "../../godotgaussiansplatting_b200/csrc/ranges.cu"
#include "../../godotgaussiansplatting_b200/csrc/radix_sort.cu"
#include "../../godotgaussiansplatting_b200/csrc/projection.cu"
#include "../../godotgaussiansplatting_b200/csrc/ingest.cu"
#include "../../godotgaussiansplatting_b200/csrc/present.cu"
#include "../../godotgaussiansplatting_b200/csrc/group.cu"

namespace {
struct Launch { const gsr::CompositeArgs *args; int variant; };
void body(void *p) {
    const Launch *l = static_cast<const Launch *>(p);
    if (l->variant == 1) gsr::composite_kernel<false>(*l->args);   // GSR_FLAG_UNCONTRACTED_BLEND
    else gsr::composite_kernel<true>(*l->args);
}
}  // namespace

// One persistent block works through every tile of the launch in ticket order (natural, or longest-chain-first through
// tile_order_kernel with or without the previous frame's consumed-chunk hints).
extern "C" int emu_composite(int variant, const void *records, const uint32_t *values, const uint32_t *bounds, float *out_rgba, int width,
                             int height, int tile_begin, int row_step, int num_tiles, float heatmap_factor, uint32_t target_tile_id,
                             float *pick4, unsigned long long *staged_out, uint32_t *hint /* nullable: [num_tiles] in/out */, int sched_flags) {
    const int tiles_x = (width + 15) / 16;
    gsr::FrameState frame;
    memset(&frame, 0, sizeof frame);
    const size_t nt = (size_t)(num_tiles > 0 ? num_tiles : 1);
    gsr::CompositeArgs a;
    memset(&a, 0, sizeof a);
    a.records = static_cast<const float4 *>(records);
    a.values = values;
    a.bounds = reinterpret_cast<const uint2 *>(bounds);
    a.out = reinterpret_cast<float4 *>(out_rgba);
    a.width = width; a.height = height; a.tiles_x = tiles_x;
    a.tile_begin = tile_begin; a.row_step = row_step; a.num_tiles = num_tiles;
    a.heatmap_factor = heatmap_factor; a.target_tile_id = target_tile_id;
    a.pick = reinterpret_cast<float4 *>(pick4);
    a.frame = &frame; a.count_staged = 1;
    a.consumed = hint; a.ctas_per_sm = 1; a.sm_count = 1; a.contract = variant == 1 ? 0 : 1;
    uint32_t *order = static_cast<uint32_t *>(calloc(nt, sizeof(uint32_t)));
    if ((sched_flags & 1) && num_tiles > 0) {   // longest-chain-first ticket order (csrc/ranges.cu tile_order_kernel, one block of 1024)
        struct OL { const uint2 *b; int tb, rs, tx, n; uint32_t *h, *o; gsr::FrameState *f; } ol{a.bounds, tile_begin, row_step, tiles_x, num_tiles, hint, order, &frame};
        cuda_emu::g_block_dim = cuda_emu::dim{1024, 1, 1};
        cuda_emu::g_grid_dim = cuda_emu::dim{1, 1, 1};
        glsl::run_workgroup(glsl::uvec3(0, 0, 0), glsl::uvec3(1024, 1, 1),
                            [](void *p) { OL *l = static_cast<OL *>(p); gsr::tile_order_kernel(l->b, l->tb, l->rs, l->tx, l->n, l->h, l->o, l->f, 3u, 1u); }, &ol);
        a.order = order;
        // a permutation of the owned tiles?
        std::vector<char> seen(nt, 0);
        for (int k = 0; k < num_tiles; ++k) { if (order[k] >= (uint32_t)num_tiles || seen[order[k]]) { free(order); return 2; } seen[order[k]] = 1; }
    }
    Launch l{&a, variant};
    cuda_emu::g_block_dim = cuda_emu::dim{128, 1, 1};
    if (num_tiles > 0) glsl::run_workgroup(glsl::uvec3(0, 0, 0), glsl::uvec3(128, 1, 1), &body, &l);
    if (staged_out) *staged_out = frame.staged;
    const int ok = frame.comp_head >= (uint32_t)(num_tiles > 0 ? num_tiles : 0);
    free(order);
    return ok ? 0 : 1;
}

// ---- csrc/ranges.cu: `grid` blocks of 256 threads, run one after another (the kernel has no inter-block dependence) ----
namespace {
struct RangesLaunch { const uint32_t *keys; const gsr::FrameState *frame; uint2 *bounds; uint32_t num_tiles; int quirks, sharded; int32_t *sync_word; };
void ranges_body(void *p) {
    const RangesLaunch *l = static_cast<const RangesLaunch *>(p);
    gsr::tile_ranges_kernel(l->keys, l->frame, l->bounds, l->num_tiles, l->quirks, l->sharded, l->sync_word);
}
struct FixupLaunch { const int32_t *last; float4 *out; int32_t w, h, tiles_x, T, y0, y1, mod, rem; };
void fixup_body(void *p) {
    const FixupLaunch *l = static_cast<const FixupLaunch *>(p);
    gsr::band_fixup_kernel(l->last, l->out, l->w, l->h, l->tiles_x, l->T, l->y0, l->y1, l->mod, l->rem);
}
}  // namespace

extern "C" int emu_tile_ranges(const uint32_t *sorted_keys, uint32_t m, uint32_t *bounds, uint32_t num_tiles, int quirks, int sharded,
                               int32_t global_last_tile, int32_t *sync_word, int grid) {
    gsr::FrameState frame;
    memset(&frame, 0, sizeof frame);
    frame.dup_total = m; frame.dup_sorted = m; frame.last_tile_plus1 = global_last_tile + 1;
    memset(bounds, 0, sizeof(uint32_t) * 2 * (size_t)num_tiles);   // rasterizer.gd:128
    RangesLaunch l{sorted_keys, &frame, reinterpret_cast<uint2 *>(bounds), num_tiles, quirks, sharded, sync_word};
    cuda_emu::g_block_dim = cuda_emu::dim{256, 1, 1};
    cuda_emu::g_grid_dim = cuda_emu::dim{(unsigned)grid, 1, 1};
    for (int b = 0; b < grid; ++b) glsl::run_workgroup(glsl::uvec3((unsigned)b, 0, 0), glsl::uvec3(256, 1, 1), &ranges_body, &l);
    cuda_emu::g_block_dim = cuda_emu::dim{128, 1, 1};
    cuda_emu::g_grid_dim = cuda_emu::dim{1, 1, 1};
    return 0;
}

extern "C" int emu_band_fixup(int32_t global_last_plus1, float *out_rgba, int width, int height, int band_y0, int band_y1, int row_mod, int row_rem) {
    const int tiles_x = (width + 15) / 16, tiles_y = (height + 15) / 16;
    FixupLaunch l{&global_last_plus1, reinterpret_cast<float4 *>(out_rgba), width, height, tiles_x, tiles_x * tiles_y, band_y0, band_y1, row_mod, row_rem};
    cuda_emu::g_block_dim = cuda_emu::dim{256, 1, 1};
    glsl::run_workgroup(glsl::uvec3(0, 0, 0), glsl::uvec3(256, 1, 1), &fixup_body, &l);
    cuda_emu::g_block_dim = cuda_emu::dim{128, 1, 1};
    return 0;
}

// ---- csrc/radix_sort.cu: the histogram kernel on `hist_grid` blocks, then four onesweep passes, each by ONE persistent block that
//      pulls every tile in ticket order (so a look-back always finds its predecessors published) ----
namespace {
struct HistLaunch { const uint32_t *keys, *n_ptr; uint32_t n_max; uint32_t *hist, *status; uint32_t tile_keys, max_tiles; };
void hist_body(void *p) {
    const HistLaunch *l = static_cast<const HistLaunch *>(p);
    gsr::sort_hist_kernel(l->keys, l->n_ptr, l->n_max, l->hist, l->status, l->tile_keys, l->max_tiles);
}
struct SweepLaunch { const uint32_t *kin; uint32_t *kout; const uint32_t *vin; uint32_t *vout; const uint32_t *n_ptr; uint32_t n_max; const uint32_t *hist; uint32_t *status, *ticket; int shift; };
void sweep_pairs_body(void *p) {
    const SweepLaunch *l = static_cast<const SweepLaunch *>(p);
    gsr::onesweep_kernel<gsr::SWEEP_THREADS, gsr::SWEEP_ITEMS, true>(l->kin, l->kout, l->vin, l->vout, l->n_ptr, l->n_max, l->hist, l->status, l->ticket, l->shift);
}
void sweep_keys_body(void *p) {
    const SweepLaunch *l = static_cast<const SweepLaunch *>(p);
    gsr::onesweep_kernel<gsr::SWEEP_THREADS, gsr::SWEEP_ITEMS, false>(l->kin, l->kout, nullptr, nullptr, l->n_ptr, l->n_max, l->hist, l->status, l->ticket, l->shift);
}
}  // namespace

// keys/values: n_max entries each, sorted in place (values may be null: keys only).  n <= n_max is read "from the device".
extern "C" int emu_sort_pairs(uint32_t *keys, uint32_t *values, uint32_t n, uint32_t n_max, int hist_grid) {
    const uint32_t tile = gsr::SWEEP_TILE, max_tiles = (n_max + tile - 1) / tile;
    uint32_t *hist = static_cast<uint32_t *>(calloc(4 * 256 + 8, sizeof(uint32_t)));
    uint32_t *status = static_cast<uint32_t *>(malloc(sizeof(uint32_t) * 4ull * max_tiles * 256));
    memset(status, 0xCD, sizeof(uint32_t) * 4ull * max_tiles * 256);   // the histogram kernel must clear what the passes use
    uint32_t *alt_k = static_cast<uint32_t *>(malloc(sizeof(uint32_t) * (size_t)n_max));
    uint32_t *alt_v = values ? static_cast<uint32_t *>(malloc(sizeof(uint32_t) * (size_t)n_max)) : nullptr;
    uint32_t *tickets = hist + 4 * 256;
    const uint32_t n_dev = n;
    HistLaunch h{keys, &n_dev, n_max, hist, status, tile, max_tiles};
    cuda_emu::g_block_dim = cuda_emu::dim{512, 1, 1};
    cuda_emu::g_grid_dim = cuda_emu::dim{(unsigned)hist_grid, 1, 1};
    for (int b = 0; b < hist_grid; ++b) glsl::run_workgroup(glsl::uvec3((unsigned)b, 0, 0), glsl::uvec3(512, 1, 1), &hist_body, &h);
    cuda_emu::g_block_dim = cuda_emu::dim{(unsigned)gsr::SWEEP_THREADS, 1, 1};
    cuda_emu::g_grid_dim = cuda_emu::dim{1, 1, 1};
    uint32_t *kin = keys, *kout = alt_k, *vin = values, *vout = alt_v;
    for (int pass = 0; pass < 4; ++pass) {
        SweepLaunch l{kin, kout, vin, vout, &n_dev, n_max, hist + pass * 256, status + (size_t)pass * max_tiles * 256, tickets + pass, 8 * pass};
        glsl::run_workgroup(glsl::uvec3(0, 0, 0), glsl::uvec3((unsigned)gsr::SWEEP_THREADS, 1, 1), values ? &sweep_pairs_body : &sweep_keys_body, &l);
        uint32_t *t = kin; kin = kout; kout = t;
        t = vin; vin = vout; vout = t;
    }
    cuda_emu::g_block_dim = cuda_emu::dim{128, 1, 1};
    free(hist); free(status); free(alt_k); free(alt_v);
    return 0;
}

// ---- csrc/projection.cu: blocks run one after another; each takes its scan position from the ticket, so a look-back always finds
//      its predecessors published.  Inside a block the warps meet through shared-memory flags (spin-waits yield to the scheduler). ----
namespace { void run_blocks(unsigned blocks, unsigned threads, void (*body)(void *), void *arg); }
namespace {
void projection_body(void *p) { gsr::projection_kernel(*static_cast<const gsr::ProjectionArgs *>(p)); }
void projection_sharded_body(void *p) { gsr::projection_sharded_kernel(*static_cast<const gsr::ProjectionArgs *>(p)); }
}  // namespace

// soa: 15 planes x plane_stride float4 (the library's SoA layout); vp: 32 floats; uniforms32: the 32-byte block.
// Per-frame constants are derived exactly like render_enqueue() in gsr_api.cu.  Returns M; outputs like the library's buffers.
extern "C" long long emu_projection(const void *soa, unsigned long long plane_stride, unsigned num_splats, const float *vp, const void *uniforms32,
                                    int band_y0, int band_y1, int row_mod, int row_rem, int fast_reject, int sh_bulk_min, void *records,
                                    uint32_t *keys, uint32_t *values, unsigned capacity, unsigned *visible_out, int *last_tile_out, unsigned *overflow_out,
                                    const uint32_t *extents /* nullable: gsr_shard_use_extents */, uint32_t *extents_out /* nullable: ONLY compute
                                    the extents of [ext_first, ext_first + ext_count), like gsr_shard_extents_compute */, unsigned ext_first, unsigned ext_count) {
    gsr::ProjectionArgs pa;
    memset(&pa, 0, sizeof pa);
    pa.soa = static_cast<const float4 *>(soa); pa.plane_stride = plane_stride; pa.num_splats = num_splats;
    memcpy(pa.vp, vp, sizeof pa.vp);
    memcpy(&pa.u, uniforms32, sizeof pa.u);
    {
        const float tfi0 = vp[16 + 0], tfi1 = vp[16 + 5];
        const volatile float hw = (float)pa.u.dims[0] * 0.5f, hh = (float)pa.u.dims[1] * 0.5f;
        const volatile float f0 = hw * tfi0, f1 = hh * tfi1;
        const volatile float t0 = 1.0f / tfi0, t1 = 1.0f / tfi1;
        const volatile float n0 = -t0, n1 = -t1;
        pa.focal_base[0] = f0; pa.focal_base[1] = f1;
        pa.lim_lo[0] = n0 * 1.3f; pa.lim_lo[1] = n1 * 1.3f;
        pa.lim_hi[0] = t0 * 1.3f; pa.lim_hi[1] = t1 * 1.3f;
    }
    const bool fast = row_mod > 1;
    pa.band_y0 = band_y0; pa.band_y1 = band_y1; pa.row_mod = row_mod; pa.row_rem = row_rem;
    pa.fast_reject = (fast_reject && fast) ? 1 : 0;
    pa.fast_mode = fast ? 1 : 0;
    pa.sh_bulk_min = sh_bulk_min > 0 ? sh_bulk_min : (fast ? 1 : 12);
    {
        float g[3][3];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) { g[i][j] = 0.0f; for (int r = 0; r < 3; ++r) g[i][j] += vp[4 * i + r] * vp[4 * j + r]; }
        float nrm = 0.0f;
        for (int i = 0; i < 3; ++i) { float row = 0.0f; for (int j = 0; j < 3; ++j) row += g[i][j] < 0.0f ? -g[i][j] : g[i][j]; nrm = row > nrm ? row : nrm; }
        pa.w_frob2 = nrm * 1.0001f;
    }
    gsr::FrameState frame;
    memset(&frame, 0, sizeof frame);
    const bool sharded_kernel = pa.fast_reject != 0;
    const unsigned per_block = sharded_kernel ? (unsigned)gsr::SH_SPLATS : (unsigned)gsr::PROJ_THREADS;
    const unsigned blocks = (num_splats + per_block - 1) / per_block;
    unsigned long long *lookback = static_cast<unsigned long long *>(calloc(blocks ? blocks : 1, sizeof(unsigned long long)));
    pa.records = static_cast<float4 *>(records); pa.keys = keys; pa.values = values; pa.capacity = capacity;
    pa.lookback = lookback; pa.frame = &frame;
    cuda_emu::g_block_dim = cuda_emu::dim{(unsigned)gsr::PROJ_THREADS, 1, 1};
    cuda_emu::g_grid_dim = cuda_emu::dim{blocks, 1, 1};
    for (unsigned b = 0; b < blocks; ++b)
        glsl::run_workgroup(glsl::uvec3(b, 0, 0), glsl::uvec3((unsigned)gsr::PROJ_THREADS, 1, 1),
                            pa.fast_reject ? &projection_sharded_body : &projection_body, &pa);
    cuda_emu::g_block_dim = cuda_emu::dim{128, 1, 1};
    cuda_emu::g_grid_dim = cuda_emu::dim{1, 1, 1};
    free(lookback);
    if (visible_out) *visible_out = frame.visible;
    if (last_tile_out) *last_tile_out = frame.last_tile_plus1 - 1;
    if (overflow_out) *overflow_out = frame.overflow;
    return (long long)frame.dup_total;
}

// ---- projection_scatter_kernel (group mode): rank `rank` of `world` projects its slice [first, first + count) and scatters pairs and
//      records into `world` destination buffers (here: plain host arrays standing in for the peers' arenas).  Returns 0 and fills
//      counts[world] (pairs sent to each destination, as published in the flag words) and *last_plus1.
extern "C" int emu_projection_scatter(const void *soa, unsigned long long plane_stride, unsigned num_splats, const float *vp, const void *uniforms32, int world,
                                      int rank, unsigned first, unsigned count, unsigned seg_cap, void **dst_records, uint32_t **dst_keys, uint32_t **dst_values,
                                      unsigned long long *counts, unsigned *last_plus1) {
    if (world < 1 || world > gsr::GROUP_MAX) return 3;
    gsr::ProjectionArgs pa;
    memset(&pa, 0, sizeof pa);
    pa.soa = static_cast<const float4 *>(soa); pa.plane_stride = plane_stride; pa.num_splats = num_splats;
    memcpy(pa.vp, vp, sizeof pa.vp);
    memcpy(&pa.u, uniforms32, sizeof pa.u);
    {
        const float tfi0 = vp[16 + 0], tfi1 = vp[16 + 5];
        const volatile float hw = (float)pa.u.dims[0] * 0.5f, hh = (float)pa.u.dims[1] * 0.5f;
        const volatile float f0 = hw * tfi0, f1 = hh * tfi1;
        const volatile float t0 = 1.0f / tfi0, t1 = 1.0f / tfi1;
        const volatile float n0 = -t0, n1 = -t1;
        pa.focal_base[0] = f0; pa.focal_base[1] = f1;
        pa.lim_lo[0] = n0 * 1.3f; pa.lim_lo[1] = n1 * 1.3f;
        pa.lim_hi[0] = t0 * 1.3f; pa.lim_hi[1] = t1 * 1.3f;
    }
    pa.band_y0 = 0; pa.band_y1 = (pa.u.dims[1] + gsr::TILE - 1) / gsr::TILE; pa.row_mod = 1; pa.row_rem = 0; pa.sh_bulk_min = 12;
    gsr::FrameState frame;
    memset(&frame, 0, sizeof frame);
    pa.frame = &frame;
    static gsr::GroupFlags flags[gsr::GROUP_MAX];
    memset(flags, 0, sizeof flags);
    gsr::ScatterPeers sp;
    memset(&sp, 0, sizeof sp);
    sp.world = world; sp.rank = rank; sp.parity = 1; sp.seq = 91u; sp.first = first; sp.count = count; sp.seg_cap = seg_cap;
    for (int d = 0; d < world; ++d) { sp.records[d] = static_cast<float4 *>(dst_records[d]); sp.keys[d] = dst_keys[d]; sp.values[d] = dst_values[d]; sp.flags[d] = &flags[d]; }
    const unsigned blocks = count ? (count + gsr::PROJ_THREADS - 1) / gsr::PROJ_THREADS : 1u;
    sp.lookback = static_cast<unsigned long long *>(calloc((size_t)blocks * world, sizeof(unsigned long long)));
    struct SL { gsr::ProjectionArgs a; gsr::ScatterPeers sp; } sl{pa, sp};
    run_blocks(blocks, (unsigned)gsr::PROJ_THREADS, [](void *p) { SL *l = static_cast<SL *>(p); gsr::projection_scatter_kernel(l->a, l->sp); }, &sl);
    free(sp.lookback);
    int rc = 0;
    for (int d = 0; d < world; ++d) {   // what the kernel's last block published to destination d
        const unsigned long long w0 = flags[d].seg_meta[1][rank][0], w1 = flags[d].seg_meta[1][rank][1];
        if ((uint32_t)(w0 >> 32) != 91u || (uint32_t)(w1 >> 32) != 91u) rc = 1;
        counts[d] = (uint32_t)w0;
        if (last_plus1) *last_plus1 = (uint32_t)w1;
    }
    if (flags[rank].scat_ticket != 0u || flags[rank].scat_last != 0) rc = 2;   // protocol words reset for the next frame
    return rc;
}

// ---- csrc/group.cu, destination side of the scatter projection: wait for the sources' flag words, publish prefix / M / overflow / last
//      tile, pack the receive segments.  counts_in[world] / last_in[world] are what the sources "published" (seq already matching).
extern "C" int emu_group_receive(int world, unsigned seg_cap, unsigned capacity, const unsigned *counts_in, const unsigned *last_in, const uint32_t *rx_keys,
                                 const uint32_t *rx_vals, uint32_t *keys, uint32_t *vals, unsigned long long *dup_total, unsigned *dup_sorted, unsigned *overflow,
                                 int *last_plus1, unsigned *prefix_out) {
    static gsr::GroupFlags flags;
    memset(&flags, 0, sizeof flags);
    for (int r = 0; r < world; ++r) { flags.seg_meta[1][r][0] = (77ull << 32) | counts_in[r]; flags.seg_meta[1][r][1] = (77ull << 32) | last_in[r]; }
    gsr::FrameState frame;
    memset(&frame, 0, sizeof frame);
    struct WL { gsr::GroupFlags *f; int world; unsigned seg_cap, cap; gsr::FrameState *fr; } wl{&flags, world, seg_cap, capacity, &frame};
    run_blocks(1, 32, [](void *p) { WL *l = static_cast<WL *>(p); gsr::group_wait_segments_kernel(l->f, 1, l->world, 77u, l->seg_cap, l->cap, l->fr, 1000ull); }, &wl);
    struct GL { const gsr::GroupFlags *f; int world; unsigned seg_cap; const uint32_t *rk, *rv; uint32_t *k, *v; } gl{&flags, world, seg_cap, rx_keys, rx_vals, keys, vals};
    run_blocks(3, 256, [](void *p) { GL *l = static_cast<GL *>(p); gsr::gather_segments_kernel(l->f, l->world, l->seg_cap, l->rk, l->rv, l->k, l->v); }, &gl);
    *dup_total = frame.dup_total; *dup_sorted = frame.dup_sorted; *overflow = frame.overflow; *last_plus1 = frame.last_tile_plus1;
    for (int r = 0; r <= world; ++r) prefix_out[r] = flags.seg_prefix[r];
    return flags.error ? 1 : 0;
}

// ---- csrc/ingest.cu ----
namespace {
struct AosLaunch { const float4 *aos; uint64_t count; float4 *soa; uint64_t stride, first; };
void aos_body(void *p) { const AosLaunch *l = static_cast<const AosLaunch *>(p); gsr::aos_to_soa_kernel(l->aos, l->count, l->soa, l->stride, l->first); }
struct PlyLaunch { const float *ply; uint32_t nprops; uint64_t count; float creation; float4 *soa; uint64_t stride, first; };
void ply_body(void *p) { const PlyLaunch *l = static_cast<const PlyLaunch *>(p); gsr::ply_to_soa_kernel(l->ply, l->nprops, l->count, l->creation, l->soa, l->stride, l->first); }
struct PackLaunch { const float4 *rgba; float4 *rgb; uint64_t quads, pixels; };
void pack_body(void *p) { const PackLaunch *l = static_cast<const PackLaunch *>(p); gsr::present_kernel<GSR_OUT_RGB32F, false>(l->rgba, l->rgb, l->pixels); }
struct PresentLaunch { const float4 *rgba; void *out; uint64_t pixels; int format; };
void present_body(void *p) {
    const PresentLaunch *l = static_cast<const PresentLaunch *>(p);
    const bool lin = (l->format & GSR_OUT_SRGB_TO_LINEAR) != 0;
    switch (l->format & 0xFF) {
        case GSR_OUT_RGBA32F: lin ? gsr::present_kernel<GSR_OUT_RGBA32F, true>(l->rgba, l->out, l->pixels) : gsr::present_kernel<GSR_OUT_RGBA32F, false>(l->rgba, l->out, l->pixels); break;
        case GSR_OUT_RGB32F: lin ? gsr::present_kernel<GSR_OUT_RGB32F, true>(l->rgba, l->out, l->pixels) : gsr::present_kernel<GSR_OUT_RGB32F, false>(l->rgba, l->out, l->pixels); break;
        case GSR_OUT_RGBA16F: lin ? gsr::present_kernel<GSR_OUT_RGBA16F, true>(l->rgba, l->out, l->pixels) : gsr::present_kernel<GSR_OUT_RGBA16F, false>(l->rgba, l->out, l->pixels); break;
        default: lin ? gsr::present_kernel<GSR_OUT_RGBA8, true>(l->rgba, l->out, l->pixels) : gsr::present_kernel<GSR_OUT_RGBA8, false>(l->rgba, l->out, l->pixels); break;
    }
}
void run_blocks(unsigned blocks, unsigned threads, void (*body)(void *), void *arg) {
    cuda_emu::g_block_dim = cuda_emu::dim{threads, 1, 1};
    cuda_emu::g_grid_dim = cuda_emu::dim{blocks, 1, 1};
    for (unsigned b = 0; b < blocks; ++b) glsl::run_workgroup(glsl::uvec3(b, 0, 0), glsl::uvec3(threads, 1, 1), body, arg);
    cuda_emu::g_block_dim = cuda_emu::dim{128, 1, 1};
    cuda_emu::g_grid_dim = cuda_emu::dim{1, 1, 1};
}
}  // namespace

extern "C" int emu_aos_to_soa(const void *aos60, unsigned long long count, void *soa, unsigned long long plane_stride, unsigned long long first) {
    AosLaunch l{static_cast<const float4 *>(aos60), count, static_cast<float4 *>(soa), plane_stride, first};
    run_blocks((unsigned)((count + gsr::SPLATS_PER_BLOCK - 1) / gsr::SPLATS_PER_BLOCK), 256, &aos_body, &l);
    return 0;
}
extern "C" int emu_ply_to_soa(const float *ply, unsigned nprops, unsigned long long count, float creation_time, void *soa,
                              unsigned long long plane_stride, unsigned long long first) {
    if (nprops > 256) return 1;
    PlyLaunch l{ply, nprops, count, creation_time, static_cast<float4 *>(soa), plane_stride, first};
    run_blocks((unsigned)((count + gsr::INGEST_SPLATS - 1) / gsr::INGEST_SPLATS), (unsigned)gsr::INGEST_SPLATS, &ply_body, &l);
    return 0;
}
extern "C" int emu_present(const void *rgba, void *out, unsigned long long pixels, int format) {
    PresentLaunch l{static_cast<const float4 *>(rgba), out, pixels, format};
    const unsigned long long items = (format & 0xFF) == GSR_OUT_RGB32F ? (pixels + 3) / 4 : pixels;
    run_blocks((unsigned)((items + 255) / 256), 256, &present_body, &l);
    return 0;
}
extern "C" int emu_pack_rgb(const void *rgba, void *rgb, unsigned long long pixels) {
    const unsigned long long quads = (pixels + 3) / 4;
    PackLaunch l{static_cast<const float4 *>(rgba), static_cast<float4 *>(rgb), quads, pixels};
    run_blocks((unsigned)((quads + 255) / 256), 256, &pack_body, &l);
    return 0;
}
