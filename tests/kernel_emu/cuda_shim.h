/*
 * cuda_shim.h -- just enough of the CUDA execution model to run ONE thread block of a libgsr kernel on the CPU.
 *
 * TEST INFRASTRUCTURE (tests/kernel_emu).  The kernels' source files are compiled by g++ with -DGSR_CPU_EMU, unchanged
 * except for the few PTX helpers that have a C twin behind that macro.  Threads of the block are fibers of
 * oracle/glsl_cpu/glsl_emu.hpp's scheduler: __syncthreads() is a real barrier, the warp collectives
 * (__any_sync / __ballot_sync / __reduce_add_sync) are real 32-wide collectives, __shared__ variables are shared by the
 * block, atomics are sequential.  What this checks is the kernels' LOGIC (indexing, double buffering, work queue, spill
 * and resume, vote parity) bit for bit against the oracle before a GPU minute is spent; what it cannot check is the
 * memory model (fences, races) and anything timing-related.  The product never links this: libgsr has no CPU path.
 */
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../oracle/glsl_cpu/glsl_emu.hpp"

#undef __shared__
#define __shared__ static /* one block at a time: a function-level static IS block-shared memory */
#undef __global__
#define __global__ static
#undef __device__
#define __device__
#undef __grid_constant__
#define __grid_constant__
#undef __launch_bounds__
#define __launch_bounds__(...)

namespace cuda_emu {
struct dim { unsigned x, y, z; };
inline dim thread_idx() { auto *c = ::glsl::sched().cur; return dim{c->local_id.x, c->local_id.y, c->local_id.z}; }
inline dim block_idx() { auto *c = ::glsl::sched().cur; return dim{c->group_id.x, c->group_id.y, c->group_id.z}; }
extern dim g_block_dim, g_grid_dim;
}  // namespace cuda_emu
#define threadIdx (::cuda_emu::thread_idx())
#define blockIdx (::cuda_emu::block_idx())
#define blockDim (::cuda_emu::g_block_dim)
#define gridDim (::cuda_emu::g_grid_dim)

inline void __syncthreads() { ::glsl::barrier(); }
inline void __syncwarp(unsigned = 0xffffffffu) {}
inline int __any_sync(unsigned, int pred) { ::glsl::subgroup_collective(pred ? 1u : 0u); return ::glsl::sched().cur->sg_ballot != 0u; }
inline int __all_sync(unsigned, int pred) { ::glsl::subgroup_collective(pred ? 0u : 1u); return ::glsl::sched().cur->sg_ballot == 0u; }
inline unsigned __ballot_sync(unsigned, int pred) { ::glsl::subgroup_collective(pred ? 1u : 0u); return ::glsl::sched().cur->sg_ballot; }
inline unsigned __reduce_add_sync(unsigned, unsigned v) { ::glsl::subgroup_collective(v); return ::glsl::sched().cur->sg_sum; }
inline unsigned __shfl_sync(unsigned, unsigned v, int src) { ::glsl::subgroup_collective(v); return ::glsl::sched().cur->sg_vals[src & 31]; }
inline int __shfl_sync(unsigned m, int v, int src) { return (int)__shfl_sync(m, (unsigned)v, src); }
inline float __shfl_sync(unsigned m, float v, int src) { unsigned u; memcpy(&u, &v, 4); u = __shfl_sync(m, u, src); memcpy(&v, &u, 4); return v; }
inline unsigned long long __shfl_sync(unsigned m, unsigned long long v, int src) {
    const unsigned lo = __shfl_sync(m, (unsigned)v, src), hi = __shfl_sync(m, (unsigned)(v >> 32), src);
    return (unsigned long long)lo | ((unsigned long long)hi << 32);
}
inline unsigned __shfl_up_sync(unsigned, unsigned v, unsigned delta) {
    ::glsl::subgroup_collective(v);
    auto *c = ::glsl::sched().cur;
    return c->sg_invocation >= delta ? c->sg_vals[c->sg_invocation - delta] : v;
}
inline unsigned __shfl_xor_sync(unsigned, unsigned v, int m) { ::glsl::subgroup_collective(v); auto *c = ::glsl::sched().cur; return c->sg_vals[(c->sg_invocation ^ (unsigned)m) & 31]; }
inline unsigned long long __shfl_xor_sync(unsigned mk, unsigned long long v, int m) {
    const unsigned lo = __shfl_xor_sync(mk, (unsigned)v, m), hi = __shfl_xor_sync(mk, (unsigned)(v >> 32), m);
    return (unsigned long long)lo | ((unsigned long long)hi << 32);
}
inline int __reduce_max_sync(unsigned, int v) {
    ::glsl::subgroup_collective((unsigned)v);
    auto *c = ::glsl::sched().cur;
    int m = v;
    for (int i = 0; i < 32; ++i) if (c->sg_part & (1u << i)) { const int x = (int)c->sg_vals[i]; m = x > m ? x : m; }
    return m;
}
inline unsigned __float2uint_rn(float x) { return x != x ? 0u : (x <= 0.0f ? 0u : (x >= 4294967296.0f ? 0xFFFFFFFFu : (unsigned)nearbyintf(x))); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
inline void __threadfence() {}
inline void __threadfence_system() {}
inline void __threadfence_block() {}
inline void __nanosleep(unsigned) { ::glsl::spin_yield(); }   /* a spin-wait iteration: let everybody else run first (the scheduler
                                                                  aborts if nobody ever satisfies the wait) */

template <class T> inline T __ldg(const T *p) { return *p; }
template <class T> inline T __ldcg(const T *p) { return *p; }
template <class T> inline void __stcg(T *p, T v) { *p = v; }

inline unsigned atomicAdd(unsigned *p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
inline unsigned atomicExch(unsigned *p, unsigned v) { unsigned o = *p; *p = v; return o; }
inline int atomicMax(int *p, int v) { int o = *p; *p = o < v ? v : o; return o; }

inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
