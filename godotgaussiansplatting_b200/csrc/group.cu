// group.cu -- device-side synchronisation of a multi-GPU shard group (one gsr_ctx per GPU, gsr_group_attach).
//
// The reference is single-device; SURVEY 8e shards the frame by tile rows.  In group mode NOTHING on the frame path goes
// through the host or through NCCL: the ranks exchange the per-splat tile-row extents with peer stores over
// NVLink/NVSwitch (extent kernel, projection.cu), composite straight into the presenting rank's frame (compositor.cu) and
// order all of it with the sequence-numbered flag words below, which live in every rank's "arena" and are written by the
// peers with system-scope stores.  Every wait is bounded (%globaltimer deadline): a lost peer raises GroupFlags::error
// instead of hanging the GPU.
#include "common.cuh"

namespace gsr {

namespace {

__device__ __forceinline__ unsigned long long now_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

// lane r < world waits until meta[parity][r] carries `seq`; the max of the ranks' slice-local last tiles is the frame-global
// last occupied tile (gsplat_boundaries.glsl:47-49 needs it), published to this rank's FrameState before its projection runs.
__global__ void __launch_bounds__(32) group_wait_extents_kernel(GroupFlags *flags, int parity, int world, uint32_t seq, FrameState *frame,
                                                                unsigned long long timeout_ns) {
    const int lane = (int)threadIdx.x;
    uint32_t last = 0u;
    if (lane < world) {
        volatile unsigned long long *w = &flags->meta[parity][lane];
        const unsigned long long deadline = now_ns() + timeout_ns;
        unsigned long long v;
        while ((uint32_t)((v = *w) >> 32) != seq) {
            if (now_ns() > deadline) { atomicExch(&flags->error, 1u); break; }
            __nanosleep(100);
        }
        last = (uint32_t)v;
    }
    __threadfence_system();  // acquire side: the table slices were written before the flag
    const uint32_t g = __reduce_max_sync(0xffffffffu, last);
    if (lane == 0 && g) atomicMax(&frame->last_tile_plus1, (int32_t)g);
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
    GSR_CUDA_TRY(cudaFuncGetAttributes(&fa, group_wait_extents_kernel));
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
int launch_group_wait_extents(GroupFlags *flags, int parity, int world, uint32_t seq, FrameState *frame, cudaStream_t stream) {
    group_wait_extents_kernel<<<1, 32, 0, stream>>>(flags, parity, world, seq, frame, GSR_GROUP_TIMEOUT_NS);
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
