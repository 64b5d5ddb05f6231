// ranges.cu -- stage 3: tile-range scan.  Replaces gsplat_boundaries.glsl:23-50.
//
// Adjacent difference over the high 16 bits of the sorted keys -> uint2 bounds[tile] = (start, end).
// bounds is cleared to 0 before the launch (rasterizer.gd:128).  The element count M is read from
// FrameState on the device (the reference dispatches indirectly from grid_dims[3..5], rasterizer.gd:153).
//
// Reference quirks (Q10), reproduced when `quirks` != 0:
//   * the first occupied tile keeps start = 0 from the clear (correct by construction);
//   * every thread whose tile is T-1 stores bounds[T-1].y = M-1 (:47-49), dropping the final splat;
//   * the last OCCUPIED tile, when it is not tile T-1, never gets its end written and renders nothing
//     (gsplat_render.glsl:61: max(0, int(0 - start))).
// Sharded runs (`sharded` != 0): this context holds one tile-row band of the frame; its band-local last
// occupied tile is only "the frame's last occupied tile" if it equals FrameState.last_tile (computed by the
// projection over ALL splats); otherwise its end is M_local, exactly what the un-sharded scan would write.
#include "common.cuh"

namespace gsr {

namespace {

__global__ void __launch_bounds__(256) tile_ranges_kernel(const uint32_t *__restrict__ keys, const FrameState *__restrict__ frame,
                                                          uint2 *__restrict__ bounds, uint32_t num_tiles, int quirks, int sharded) {
    const uint32_t m = frame->dup_sorted;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t id = blockIdx.x * blockDim.x + threadIdx.x; id < m; id += stride) {
        const uint32_t b = keys[id] >> 16;
        if (id > 0) {
            const uint32_t a = keys[id - 1] >> 16;
            if (a != b) {
                bounds[a].y = id;
                bounds[b].x = id;
            }
        }
        if (id == m - 1) {  // tail rules for the last key's tile
            if (quirks) {
                if (b == num_tiles - 1u) {
                    if (m - 1u >= 1u) bounds[b].y = m - 1u;
                } else if (sharded && (int32_t)b != frame->last_tile_plus1 - 1) {
                    bounds[b].y = m;
                }
            } else {
                bounds[b].y = m;
            }
        }
    }
}

}  // namespace

int launch_tile_ranges(const uint32_t *sorted_keys, const FrameState *frame, uint2 *bounds, uint32_t num_tiles, int quirks,
                       int sharded, int grid, cudaStream_t stream) {
    tile_ranges_kernel<<<grid, 256, 0, stream>>>(sorted_keys, frame, bounds, num_tiles, quirks, sharded);
    GSR_CUDA_TRY(cudaGetLastError());
    return GSR_OK;
}

}  // namespace gsr
