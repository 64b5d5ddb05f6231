// group.cu -- device-side synchronisation of a multi-GPU shard group (one gsr_ctx per GPU, gsr_group_attach).
//
// The reference is single-device; SURVEY 8e shards the frame by tile rows.  In group mode NOTHING on the frame path goes
// through the host or through NCCL: every rank projects ITS slice of the splats and stores the resulting pairs and records
// straight into the owning rank's memory over NVLink/NVSwitch (projection_scatter_kernel, projection.cu), composites straight
// into the presenting rank's frame (compositor.cu), and all of it is ordered by the sequence-numbered flag words below, which
// live in every rank's "arena" and are written by the peers with system-scope stores.  Every wait is bounded (%globaltimer deadline): a lost peer raises GroupFlags::error
// instead of hanging the GPU.
#include "common.cuh"

namespace gsr {

namespace {

#ifndef GSR_CPU_EMU
__device__ __forceinline__ unsigned long long now_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#else
inline unsigned long long now_ns() { return 0ull; }
#endif

// Destination side of the scatter projection.  Lane r < world waits until source r's two flag words of this parity carry `seq`
// (the source stored its pairs and records first, then a system fence, then the flags).  Then: exclusive prefix of the received
// segment lengths (clamped to the segment capacity) -> seg_prefix; M, overflow and the frame-global last occupied tile (max over the
// sources' slices; gsplat_boundaries.glsl:47-49 needs it) -> this rank's FrameState, before its sort runs.
__global__ void __launch_bounds__(32) group_wait_segments_kernel(GroupFlags *flags, int parity, int world, uint32_t seq, uint32_t seg_cap, uint32_t capacity,
                                                                 FrameState *frame, unsigned long long timeout_ns) {
    const int lane = (int)threadIdx.x;
    uint32_t count = 0u, last = 0u;
    if (lane < world) {
        volatile unsigned long long *w = &flags->seg_meta[parity][lane][0];
        const unsigned long long deadline = now_ns() + timeout_ns;
        unsigned long long v0, v1;
        while ((uint32_t)((v0 = w[0]) >> 32) != seq || (uint32_t)((v1 = w[1]) >> 32) != seq) {
            if (now_ns() > deadline) { atomicExch(&flags->error, 1u); v0 = v1 = 0ull; break; }
            __nanosleep(100);
        }
        count = (uint32_t)v0; last = (uint32_t)v1;
    }
    __threadfence_system();  // acquire side: the pairs and records were written before the flags
    const uint32_t kept = count < seg_cap ? count : seg_cap;
    uint32_t incl = kept;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const uint32_t m = __shfl_sync(0xffffffffu, incl, world - 1);   // (warp collectives stay outside divergent code)
    if (lane <= world) flags->seg_prefix[lane] = lane == world ? m : incl - kept;
    const unsigned long long total = (unsigned long long)__reduce_add_sync(0xffffffffu, count > 0x7FFFFFFFu ? 0x7FFFFFFFu : count);  // 16 x 2^31 fits
    const uint32_t any_over = __any_sync(0xffffffffu, count > seg_cap) ? 1u : 0u;
    const uint32_t g = __reduce_max_sync(0xffffffffu, last);
    if (lane == 0) {
        frame->dup_total = total;
        frame->dup_sorted = m < capacity ? m : capacity;
        frame->overflow = (any_over || m > capacity) ? 1u : 0u;
        if (g) atomicMax(&frame->last_tile_plus1, (int32_t)g);
    }
}

// Pack the `world` receive segments (source r's pairs at [r * seg_cap, r * seg_cap + len_r)) into the contiguous sort input, in source
// order = splat-id order.  A pure streaming copy of 8 * M bytes; the segment of an element is found in the <= 17-entry prefix table.
__global__ void __launch_bounds__(256) gather_segments_kernel(const GroupFlags *__restrict__ flags, int world, uint32_t seg_cap, const uint32_t *__restrict__ rx_keys,
                                                              const uint32_t *__restrict__ rx_vals, uint32_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    __shared__ uint32_t s_prefix[GROUP_MAX + 1];
    if (threadIdx.x <= (uint32_t)world) s_prefix[threadIdx.x] = flags->seg_prefix[threadIdx.x];
    __syncthreads();
    const uint32_t m = s_prefix[world];
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
        int sgm = 0;
        while (sgm + 1 < world && i >= s_prefix[sgm + 1]) ++sgm;
        const uint64_t src = (uint64_t)sgm * seg_cap + (i - s_prefix[sgm]);
        keys[i] = __ldg(rx_keys + src);
        vals[i] = __ldg(rx_vals + src);
    }
}

__global__ void group_wait_u32_kernel(GroupFlags *flags, const volatile uint32_t *word, int count, uint32_t need, unsigned long long timeout_ns) {
    const int lane = (int)threadIdx.x;
    if (lane < count) {
        const unsigned long long deadline = now_ns() + timeout_ns;
        while ((int32_t)(word[lane] - need) < 0) {
            if (now_ns() > deadline) { atomicExch(&flags->error, 2u); break; }
            __nanosleep(100);
        }
    }
    __threadfence_system();
}

// one system-scope 32-bit store per destination: dst[i] points at the word to set on peer i
__global__ void group_store_u32_kernel(GroupPeers peers, int which, int index, int count, uint32_t value) {
    const int lane = (int)threadIdx.x;
    __threadfence_system();  // everything this stream did before (kernel boundary) is ordered before the flag
    if (lane < count) {
        GroupFlags *f = peers.flags[which < 0 ? lane : which];
        volatile uint32_t *w = index < 0 ? &f->released : &f->done[index];
        *w = value;
    }
}

__global__ void __launch_bounds__(32) publish_frame_state_kernel(const FrameState *__restrict__ frame, FrameState *host_mapped) {
    if (threadIdx.x < 4u) {
        const uint4 v = reinterpret_cast<const uint4 *>(frame)[threadIdx.x];
        reinterpret_cast<volatile uint4 *>(host_mapped)[threadIdx.x].x = v.x;
        reinterpret_cast<volatile uint4 *>(host_mapped)[threadIdx.x].y = v.y;
        reinterpret_cast<volatile uint4 *>(host_mapped)[threadIdx.x].z = v.z;
        reinterpret_cast<volatile uint4 *>(host_mapped)[threadIdx.x].w = v.w;
    }
    __threadfence_system();
}

}  // namespace

#ifndef GSR_CPU_EMU
// Force-load this file's kernels (CUDA loads modules lazily; a first launch that has to load code while another context's
// kernel spins on a flag this launch would satisfy can stall the host: see gsr_group_attach).
int preload_group_kernels() {
    cudaFuncAttributes fa;
    GSR_CUDA_TRY(cudaFuncGetAttributes(&fa, group_wait_segments_kernel));
    GSR_CUDA_TRY(cudaFuncGetAttributes(&fa, gather_segments_kernel));
    GSR_CUDA_TRY(cudaFuncGetAttributes(&fa, group_wait_u32_kernel));
    GSR_CUDA_TRY(cudaFuncGetAttributes(&fa, group_store_u32_kernel));
    GSR_CUDA_TRY(cudaFuncGetAttributes(&fa, publish_frame_state_kernel));
    return GSR_OK;
}
int launch_publish_frame_state(const FrameState *frame, FrameState *host_mapped, cudaStream_t stream) {
    publish_frame_state_kernel<<<1, 32, 0, stream>>>(frame, host_mapped);
    GSR_CUDA_TRY(cudaGetLastError());
    return GSR_OK;
}
int launch_group_wait_segments(GroupFlags *flags, int parity, int world, uint32_t seq, uint32_t seg_cap, uint32_t capacity, FrameState *frame, cudaStream_t stream) {
    group_wait_segments_kernel<<<1, 32, 0, stream>>>(flags, parity, world, seq, seg_cap, capacity, frame, GSR_GROUP_TIMEOUT_NS);
    GSR_CUDA_TRY(cudaGetLastError());
    return GSR_OK;
}
int launch_gather_segments(const GroupFlags *flags, int world, uint32_t seg_cap, const uint32_t *rx_keys, const uint32_t *rx_vals, uint32_t *keys, uint32_t *vals,
                           int grid, cudaStream_t stream) {
    gather_segments_kernel<<<grid, 256, 0, stream>>>(flags, world, seg_cap, rx_keys, rx_vals, keys, vals);
    GSR_CUDA_TRY(cudaGetLastError());
    return GSR_OK;
}
int launch_group_wait_released(GroupFlags *flags, uint32_t need, cudaStream_t stream) {
    group_wait_u32_kernel<<<1, 32, 0, stream>>>(flags, &flags->released, 1, need, GSR_GROUP_TIMEOUT_NS);
    GSR_CUDA_TRY(cudaGetLastError());
    return GSR_OK;
}
int launch_group_wait_done(GroupFlags *flags, int world, uint32_t seq, cudaStream_t stream) {
    group_wait_u32_kernel<<<1, 32, 0, stream>>>(flags, flags->done, world, seq, GSR_GROUP_TIMEOUT_NS);
    GSR_CUDA_TRY(cudaGetLastError());
    return GSR_OK;
}
int launch_group_signal_done(const GroupPeers &peers, int root, int rank, uint32_t seq, cudaStream_t stream) {
    group_store_u32_kernel<<<1, 32, 0, stream>>>(peers, root, rank, 1, seq);
    GSR_CUDA_TRY(cudaGetLastError());
    return GSR_OK;
}
int launch_group_release(const GroupPeers &peers, int world, uint32_t value, cudaStream_t stream) {
    group_store_u32_kernel<<<1, 32, 0, stream>>>(peers, -1, -1, world, value);
    GSR_CUDA_TRY(cudaGetLastError());
    return GSR_OK;
}
#endif

}  // namespace gsr
