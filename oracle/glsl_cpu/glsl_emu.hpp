/*
 * glsl_emu.hpp -- a small CPU execution environment for Vulkan-GLSL compute shaders.
 *
 * TEST INFRASTRUCTURE (part of oracle/).  It exists so that the reference's OWN shader sources
 * (the .glsl files under /root/reference/resources/shaders/compute) can be executed in the build container:
 * oracle/glsl_cpu/translate.py turns each .glsl file -- read where it lies, never copied into this
 * repository -- into one C++ translation unit whose body is the shader text (declarations rewrapped,
 * expressions untouched), compiled against this header into oracle/_ref/libgsr_refshaders.so.
 * tests/test_refshaders.py then pins the C restatement (gsr_oracle.c) against those outputs.
 *
 * What this header provides
 *   - GLSL value types (vec2/3/4, ivec2/4, uvec2/3/4, bvec2/3, mat3, mat4) with the swizzles, the
 *     constructors and the implicit int->uint->float conversions the six shaders use;
 *   - built-ins: min/max/clamp/mix written as their GLSL definitions, ceil, sqrt, normalize,
 *     transpose, lessThan/greaterThan/any, bitCount, bitfieldExtract, atomics, image store;
 *   - exp()/pow(): GLSL leaves them implementation-defined; GLSL_EMU_LIBM selects glibc's
 *     expf/powf, otherwise they are the oracle's orc_test_exp/orc_test_pow (so that a bit-exact
 *     comparison with gsr_oracle.c is meaningful);
 *   - an invocation scheduler: every invocation of a workgroup is a fiber (ucontext); barrier()
 *     and the subgroup operations (32-wide: subgroupElect/Ballot/Add/ExclusiveAdd) are blocking
 *     collectives over the invocations that reach them ("maximal reconvergence").  Fibers run in
 *     ascending invocation order and workgroups in ascending order, which makes every atomic
 *     deterministic (the projection's atomicAdd hands out offsets in splat-id order, the
 *     convention gsr_oracle.c fixes for the reference's nondeterminism Q13).
 *
 * Floating-point model: every GLSL operator is one IEEE binary32 operation, vector and matrix
 * operators componentwise / left-to-right dot products, no contraction (build with
 * -ffp-contract=off).  That is one conforming evaluation of the GLSL text.
 */
#pragma once
#include <ucontext.h>

#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

extern "C" {
float orc_test_exp(float);
float orc_test_pow(float, float);
float sqrtf(float);
float ceilf(float);
float expf(float);
float powf(float, float);
}

namespace glsl {

typedef unsigned int uint;

/* ---------------------------------------------------------------------------------------------- */
/* vectors                                                                                        */
/* ---------------------------------------------------------------------------------------------- */
/* swizzle proxy: lives in a union with the components; converts to / assigns from the vector V */
template <class V, class T, int N, int... I>
struct swz {
    T d[N];
    operator V() const { return V(d[I]...); }
    swz& operator=(const V& v) {
        const int idx[] = {I...};
        for (unsigned k = 0; k < sizeof...(I); ++k) d[idx[k]] = v[k];
        return *this;
    }
};

struct ivec2;
struct uvec2;
struct ivec4;

struct vec2 {
    union {
        struct { float x, y; };
        swz<vec2, float, 2, 0, 1> xy;
    };
    vec2() = default;
    explicit vec2(float s) { x = s; y = s; }
    vec2(float a, float b) { x = a; y = b; }
    vec2(const ivec2& v);
    vec2(const uvec2& v);
    float& operator[](int i) { return (&x)[i]; }
    const float& operator[](int i) const { return (&x)[i]; }
};

struct vec3 {
    union {
        struct { float x, y, z; };
        struct { float r, g, b; };
        swz<vec2, float, 3, 0, 1> xy;
        swz<vec3, float, 3, 0, 1, 2> xyz, rgb;
    };
    vec3() = default;
    explicit vec3(float s) { x = s; y = s; z = s; }
    vec3(float a, float b, float c) { x = a; y = b; z = c; }
    vec3(const vec2& v, float c) { x = v.x; y = v.y; z = c; }
    float& operator[](int i) { return (&x)[i]; }
    const float& operator[](int i) const { return (&x)[i]; }
};

struct vec4 {
    union {
        struct { float x, y, z, w; };
        struct { float r, g, b, a; };
        swz<vec2, float, 4, 0, 1> xy;
        swz<vec2, float, 4, 3, 3> ww;
        swz<vec3, float, 4, 0, 1, 2> xyz, rgb;
    };
    vec4() = default;
    explicit vec4(float s) { x = s; y = s; z = s; w = s; }
    vec4(float a, float b, float c, float d) { x = a; y = b; z = c; w = d; }
    vec4(const vec3& v, float d) { x = v.x; y = v.y; z = v.z; w = d; }
    float& operator[](int i) { return (&x)[i]; }
    const float& operator[](int i) const { return (&x)[i]; }
};

struct ivec2 {
    union {
        struct { int x, y; };
        swz<ivec2, int, 2, 0, 1> xy;
    };
    ivec2() = default;
    explicit ivec2(int s) { x = s; y = s; }
    ivec2(int a, int b) { x = a; y = b; }
    explicit ivec2(const vec2& v) { x = int(v.x); y = int(v.y); } /* float -> int: truncation */
    int& operator[](int i) { return (&x)[i]; }
    const int& operator[](int i) const { return (&x)[i]; }
};

struct uvec2 {
    union {
        struct { uint x, y; };
        swz<uvec2, uint, 2, 0, 1> xy;
    };
    uvec2() = default;
    explicit uvec2(uint s) { x = s; y = s; }
    uvec2(uint a, uint b) { x = a; y = b; }
    uvec2(const ivec2& v) { x = uint(v.x); y = uint(v.y); } /* GLSL implicit int -> uint */
    uint& operator[](int i) { return (&x)[i]; }
    const uint& operator[](int i) const { return (&x)[i]; }
};

struct uvec3 {
    union {
        struct { uint x, y, z; };
        swz<uvec2, uint, 3, 0, 1> xy;
    };
    uvec3() = default;
    uvec3(uint a, uint b, uint c) { x = a; y = b; z = c; }
    uint& operator[](int i) { return (&x)[i]; }
    const uint& operator[](int i) const { return (&x)[i]; }
};

struct ivec4 {
    int x, y, z, w;
    ivec4() = default;
    ivec4(int a, int b, int c, int d) : x(a), y(b), z(c), w(d) {}
    ivec4(const vec2& a, const vec2& b) : x(int(a.x)), y(int(a.y)), z(int(b.x)), w(int(b.y)) {}
};

struct uvec4 {
    uint x, y, z, w;
    uvec4() = default;
    explicit uvec4(uint s) : x(s), y(s), z(s), w(s) {}
    uvec4(uint a, uint b, uint c, uint d) : x(a), y(b), z(c), w(d) {}
    uvec4(const ivec4& v) : x(uint(v.x)), y(uint(v.y)), z(uint(v.z)), w(uint(v.w)) {}
    uint& operator[](int i) { return (&x)[i]; }
    const uint& operator[](int i) const { return (&x)[i]; }
};

struct bvec2 { bool x, y; };
struct bvec3 { bool x, y, z; };

inline vec2::vec2(const ivec2& v) { x = float(v.x); y = float(v.y); }
inline vec2::vec2(const uvec2& v) { x = float(v.x); y = float(v.y); }

/* componentwise float operators */
#define GLSL_VEC_OPS(V, N)                                                                                  \
    inline V operator+(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] + b[i]; return r; } \
    inline V operator-(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] - b[i]; return r; } \
    inline V operator*(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] * b[i]; return r; } \
    inline V operator/(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] / b[i]; return r; } \
    inline V operator+(const V& a, float s) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] + s; return r; }      \
    inline V operator-(const V& a, float s) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] - s; return r; }      \
    inline V operator*(const V& a, float s) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] * s; return r; }      \
    inline V operator/(const V& a, float s) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] / s; return r; }      \
    inline V operator+(float s, const V& a) { V r; for (int i = 0; i < N; ++i) r[i] = s + a[i]; return r; }      \
    inline V operator-(float s, const V& a) { V r; for (int i = 0; i < N; ++i) r[i] = s - a[i]; return r; }      \
    inline V operator*(float s, const V& a) { V r; for (int i = 0; i < N; ++i) r[i] = s * a[i]; return r; }      \
    inline V operator/(float s, const V& a) { V r; for (int i = 0; i < N; ++i) r[i] = s / a[i]; return r; }      \
    inline V operator-(const V& a) { V r; for (int i = 0; i < N; ++i) r[i] = -a[i]; return r; }                  \
    inline V& operator+=(V& a, const V& b) { a = a + b; return a; }                                              \
    inline V& operator-=(V& a, const V& b) { a = a - b; return a; }                                              \
    inline V& operator*=(V& a, const V& b) { a = a * b; return a; }                                              \
    inline V& operator*=(V& a, float s) { a = a * s; return a; }
GLSL_VEC_OPS(vec2, 2)
GLSL_VEC_OPS(vec3, 3)
GLSL_VEC_OPS(vec4, 4)
#undef GLSL_VEC_OPS

/* integer vectors: only what the shaders use */
inline ivec2 operator+(const ivec2& a, int s) { return ivec2(a.x + s, a.y + s); }
inline ivec2 operator-(const ivec2& a, int s) { return ivec2(a.x - s, a.y - s); }
inline ivec2 operator/(const ivec2& a, int s) { return ivec2(a.x / s, a.y / s); }
inline vec2 operator*(const ivec2& a, float s) { return vec2(float(a.x) * s, float(a.y) * s); } /* int -> float */
inline uvec2 operator*(const uvec2& a, uint s) { return uvec2(a.x * s, a.y * s); }
inline uvec2 operator+(const uvec2& a, const uvec2& b) { return uvec2(a.x + b.x, a.y + b.y); }
inline uvec4 operator&(const uvec4& a, const uvec4& b) { return uvec4(a.x & b.x, a.y & b.y, a.z & b.z, a.w & b.w); }
inline uvec4 operator^(const uvec4& a, const uvec4& b) { return uvec4(a.x ^ b.x, a.y ^ b.y, a.z ^ b.z, a.w ^ b.w); }
inline uvec4& operator&=(uvec4& a, const uvec4& b) { a = a & b; return a; }

/* ---------------------------------------------------------------------------------------------- */
/* matrices: column-major, m[c][r]; products are left-to-right dot products                        */
/* ---------------------------------------------------------------------------------------------- */
struct mat4 {
    vec4 c[4];
    vec4& operator[](int i) { return c[i]; }
    const vec4& operator[](int i) const { return c[i]; }
};
struct mat3 {
    vec3 c[3];
    mat3() = default;
    mat3(float a0, float a1, float a2, float b0, float b1, float b2, float c0, float c1, float c2) {
        c[0] = vec3(a0, a1, a2); c[1] = vec3(b0, b1, b2); c[2] = vec3(c0, c1, c2);
    }
    explicit mat3(const mat4& m) { for (int i = 0; i < 3; ++i) c[i] = vec3(m[i].x, m[i].y, m[i].z); }
    vec3& operator[](int i) { return c[i]; }
    const vec3& operator[](int i) const { return c[i]; }
};
inline mat3 operator*(const mat3& a, float s) { mat3 r; for (int i = 0; i < 3; ++i) r[i] = a[i] * s; return r; }
inline mat3 operator*(const mat3& a, const mat3& b) {
    mat3 o;
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) o[c][r] = (a[0][r] * b[c][0] + a[1][r] * b[c][1]) + a[2][r] * b[c][2];
    return o;
}
inline vec4 operator*(const mat4& m, const vec4& v) {
    vec4 o;
    for (int r = 0; r < 4; ++r) o[r] = ((m[0][r] * v[0] + m[1][r] * v[1]) + m[2][r] * v[2]) + m[3][r] * v[3];
    return o;
}
inline mat3 transpose(const mat3& a) {
    mat3 o;
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) o[c][r] = a[r][c];
    return o;
}

/* ---------------------------------------------------------------------------------------------- */
/* built-in functions (GLSL 4.60 spec section 8, written as their definitions)                     */
/* ---------------------------------------------------------------------------------------------- */
inline float max(float x, float y) { return (x < y) ? y : x; }
inline float min(float x, float y) { return (y < x) ? y : x; }
inline int max(int x, int y) { return (x < y) ? y : x; }
inline int min(int x, int y) { return (y < x) ? y : x; }
inline uint max(uint x, uint y) { return (x < y) ? y : x; }
inline uint min(uint x, uint y) { return (y < x) ? y : x; }
inline uint min(int x, uint y) { return min(uint(x), y); } /* implicit int -> uint */
inline uint min(uint x, int y) { return min(x, uint(y)); }
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
inline vec2 max(const vec2& a, const vec2& b) { return vec2(max(a.x, b.x), max(a.y, b.y)); }
inline vec3 max(const vec3& a, const vec3& b) { return vec3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
inline vec2 min(const vec2& a, const vec2& b) { return vec2(min(a.x, b.x), min(a.y, b.y)); }
inline vec2 clamp(const vec2& x, const vec2& lo, const vec2& hi) { return min(max(x, lo), hi); }
inline float mix(float x, float y, float a) { return x * (1.0f - a) + y * a; }
inline vec3 mix(const vec3& x, const vec3& y, float a) { return vec3(mix(x.x, y.x, a), mix(x.y, y.y, a), mix(x.z, y.z, a)); }
inline float sqrt(float x) { return ::sqrtf(x); }
inline float ceil(float x) { return ::ceilf(x); }
inline vec2 ceil(const vec2& v) { return vec2(ceil(v.x), ceil(v.y)); }
#ifdef GLSL_EMU_LIBM
inline float exp(float x) { return ::expf(x); }
inline float pow(float x, float y) { return ::powf(x, y); }
#else
inline float exp(float x) { return ::orc_test_exp(x); }
inline float pow(float x, float y) { return ::orc_test_pow(x, y); }
#endif
/* normalize(v) = v * (1/length(v)); length = sqrt of the left-to-right dot product */
inline vec3 normalize(const vec3& v) {
    float inv_len = 1.0f / sqrt((v.x * v.x + v.y * v.y) + v.z * v.z);
    return v * inv_len;
}
inline bvec3 lessThan(const vec3& a, const vec3& b) { return bvec3{a.x < b.x, a.y < b.y, a.z < b.z}; }
inline bvec3 greaterThan(const vec3& a, const vec3& b) { return bvec3{a.x > b.x, a.y > b.y, a.z > b.z}; }
inline bvec2 lessThan(const vec2& a, const vec2& b) { return bvec2{a.x < b.x, a.y < b.y}; }
inline bool any(const bvec3& v) { return v.x || v.y || v.z; }
inline bool any(const bvec2& v) { return v.x || v.y; }
inline uvec4 bitCount(const uvec4& v) {
    return uvec4(uint(__builtin_popcount(v.x)), uint(__builtin_popcount(v.y)), uint(__builtin_popcount(v.z)),
                 uint(__builtin_popcount(v.w)));
}
inline uint bitfieldExtract(uint value, int offset, int bits) { return (value >> offset) & ((1u << bits) - 1u); }
/* one scheduler thread runs every invocation, so plain read-modify-write is atomic */
inline uint atomicAdd(uint& mem, uint v) { uint old = mem; mem = old + v; return old; }
inline uint atomicMax(uint& mem, uint v) { uint old = mem; mem = (old < v) ? v : old; return old; }

/* ---------------------------------------------------------------------------------------------- */
/* resources                                                                                      */
/* ---------------------------------------------------------------------------------------------- */
/* storage-buffer array member with robustBufferAccess semantics: out-of-range loads return zero,
 * out-of-range stores are dropped */
template <class T>
struct buffer_array {
    T* p = nullptr;
    size_t n = 0;
    T dummy;
    void bind(void* base, size_t offset, size_t buffer_bytes) {
        p = reinterpret_cast<T*>(static_cast<char*>(base) + offset);
        n = buffer_bytes > offset ? (buffer_bytes - offset) / sizeof(T) : 0;
    }
    int length() const { return int(n); }
    T& operator[](size_t i) {
        if (i < n) return p[i];
        std::memset(&dummy, 0, sizeof(T));
        return dummy;
    }
    T& operator[](int i) { return (*this)[size_t(i < 0 ? ~size_t(0) : size_t(i))]; }
    T& operator[](uint i) { return (*this)[size_t(i)]; }
};

struct image2D {
    float* texels = nullptr; /* rgba32f, row-major */
    int width = 0, height = 0;
};
inline ivec2 imageSize(const image2D& im) { return ivec2(im.width, im.height); }
inline void imageStore(image2D& im, const ivec2& p, const vec4& v) {
    if (p.x < 0 || p.y < 0 || p.x >= im.width || p.y >= im.height) return;
    float* o = im.texels + (size_t(p.y) * im.width + p.x) * 4;
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}

/* contents of `shared` variables at workgroup start (GLSL: undefined); see refshader_*_set_shared_fill */
inline uint shared_fill_word = 0;
inline void fill_words(void* mem, size_t bytes) {
    uint* w = static_cast<uint*>(mem);
    for (size_t i = 0; i < bytes / sizeof(uint); ++i) w[i] = shared_fill_word;
}

constexpr size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
/* std430 / std140 base alignment and size of the member types the shaders' blocks use */
template <class T> struct layout_of { static constexpr size_t align = alignof(T), size = sizeof(T); };
template <> struct layout_of<vec2> { static constexpr size_t align = 8, size = 8; };
template <> struct layout_of<ivec2> { static constexpr size_t align = 8, size = 8; };
template <> struct layout_of<uvec2> { static constexpr size_t align = 8, size = 8; };
template <> struct layout_of<vec3> { static constexpr size_t align = 16, size = 12; };
template <> struct layout_of<vec4> { static constexpr size_t align = 16, size = 16; };
template <> struct layout_of<mat4> { static constexpr size_t align = 16, size = 64; };

/* ---------------------------------------------------------------------------------------------- */
/* invocation scheduler                                                                           */
/* ---------------------------------------------------------------------------------------------- */
enum { EMU_SUBGROUP_WIDTH = 32 };
enum inv_state { INV_READY, INV_AT_BARRIER, INV_AT_SUBGROUP, INV_DONE, INV_YIELD };

#if defined(__x86_64__) && !defined(GLSL_EMU_UCONTEXT)
#define GLSL_EMU_FAST_SWITCH 1
/* minimal System-V x86-64 context switch (callee-saved registers + stack pointer); swapcontext() costs a
 * sigprocmask system call per switch, and a sort pass needs millions of switches */
__attribute__((naked, noinline, used)) static void ctx_switch(void** /*save_sp: rdi*/, void* /*load_sp: rsi*/) {
    __asm__ volatile(
        "pushq %rbp\n\tpushq %rbx\n\tpushq %r12\n\tpushq %r13\n\tpushq %r14\n\tpushq %r15\n\t"
        "movq %rsp, (%rdi)\n\tmovq %rsi, %rsp\n\t"
        "popq %r15\n\tpopq %r14\n\tpopq %r13\n\tpopq %r12\n\tpopq %rbx\n\tpopq %rbp\n\tret\n\t");
}
#endif

struct invocation {
#ifdef GLSL_EMU_FAST_SWITCH
    void* sp;
#else
    ucontext_t ctx;
#endif
    inv_state state;
    uvec3 global_id, local_id, group_id;
    uint local_index, sg_invocation, sg_id, num_subgroups;
    /* subgroup collective exchange */
    uint sg_in, sg_sum, sg_excl, sg_ballot;
    uint sg_vals[32]; /* every participant's sg_in (0 for non-participants): lane shuffles of the CUDA shim */
    uint sg_part;     /* participation mask of the last collective */
    bool sg_first;
};

struct scheduler {
    invocation* inv = nullptr;
    char* stacks = nullptr;
    size_t capacity = 0;
#ifdef GLSL_EMU_FAST_SWITCH
    void* main_sp = nullptr;
#else
    ucontext_t main_ctx;
#endif
    invocation* cur = nullptr;
    void (*body)(void*) = nullptr;
    void* self = nullptr;
    static constexpr size_t STACK = 256 * 1024;
    ~scheduler() { std::free(inv); std::free(stacks); }
};
inline scheduler& sched() { static thread_local scheduler s; return s; }

#ifdef GLSL_EMU_FAST_SWITCH
inline void yield_to_scheduler() { scheduler& s = sched(); ctx_switch(&s.cur->sp, s.main_sp); }
inline void resume(scheduler& s, invocation& v) { s.cur = &v; ctx_switch(&s.main_sp, v.sp); }
#else
inline void yield_to_scheduler() { scheduler& s = sched(); swapcontext(&s.cur->ctx, &s.main_ctx); }
inline void resume(scheduler& s, invocation& v) { s.cur = &v; swapcontext(&s.main_ctx, &v.ctx); }
#endif
inline void barrier() { sched().cur->state = INV_AT_BARRIER; yield_to_scheduler(); }
/* a spin-wait iteration (CUDA shim: __nanosleep): the invocation stays runnable, everybody else gets a turn first */
inline void spin_yield() { sched().cur->state = INV_YIELD; yield_to_scheduler(); }
inline void subgroup_collective(uint v) {
    invocation* me = sched().cur;
    me->sg_in = v;
    me->state = INV_AT_SUBGROUP;
    yield_to_scheduler();
}
inline bool subgroupElect() { subgroup_collective(0); return sched().cur->sg_first; }
inline uvec4 subgroupBallot(bool b) { subgroup_collective(b ? 1u : 0u); return uvec4(sched().cur->sg_ballot, 0, 0, 0); }
inline uint subgroupAdd(uint v) { subgroup_collective(v); return sched().cur->sg_sum; }
inline uint subgroupExclusiveAdd(uint v) { subgroup_collective(v); return sched().cur->sg_excl; }

inline void fiber_entry() {
    scheduler& s = sched();
    s.body(s.self);
    s.cur->state = INV_DONE;
#ifdef GLSL_EMU_FAST_SWITCH
    yield_to_scheduler(); /* never resumed */
    std::abort();
#endif
    /* ucontext: uc_link returns to the scheduler */
}

/* run one workgroup: `body(self)` is the shader's main() for the current invocation */
inline void run_workgroup(uvec3 group, uvec3 local_size, void (*body)(void*), void* self) {
    scheduler& s = sched();
    const size_t n = size_t(local_size.x) * local_size.y * local_size.z;
    if (n > s.capacity) {
        std::free(s.inv); std::free(s.stacks);
        s.inv = static_cast<invocation*>(std::calloc(n, sizeof(invocation)));
        s.stacks = static_cast<char*>(std::malloc(n * scheduler::STACK));
        s.capacity = n;
        if (!s.inv || !s.stacks) { std::fprintf(stderr, "glsl_emu: out of memory\n"); std::abort(); }
    }
    s.body = body; s.self = self;
    const uint nsub = uint((n + EMU_SUBGROUP_WIDTH - 1) / EMU_SUBGROUP_WIDTH);
    for (size_t i = 0; i < n; ++i) {
        invocation& v = s.inv[i];
        v.local_index = uint(i);
        v.local_id = uvec3(uint(i % local_size.x), uint((i / local_size.x) % local_size.y), uint(i / (size_t(local_size.x) * local_size.y)));
        v.group_id = group;
        v.global_id = uvec3(group.x * local_size.x + v.local_id.x, group.y * local_size.y + v.local_id.y, group.z * local_size.z + v.local_id.z);
        v.sg_invocation = uint(i % EMU_SUBGROUP_WIDTH);
        v.sg_id = uint(i / EMU_SUBGROUP_WIDTH);
        v.num_subgroups = nsub;
        v.state = INV_READY;
#ifdef GLSL_EMU_FAST_SWITCH
        /* initial frame: six zeroed callee-saved registers, the entry point as return address, a null
         * return address above it (so that fiber_entry starts with the ABI's rsp = 16k + 8) */
        void** top = reinterpret_cast<void**>((reinterpret_cast<uintptr_t>(s.stacks + (i + 1) * scheduler::STACK)) & ~uintptr_t(15));
        top[-1] = nullptr;
        top[-2] = reinterpret_cast<void*>(&fiber_entry);
        for (int k = 3; k <= 8; ++k) top[-k] = nullptr;
        v.sp = top - 8;
#else
        getcontext(&v.ctx);
        v.ctx.uc_stack.ss_sp = s.stacks + i * scheduler::STACK;
        v.ctx.uc_stack.ss_size = scheduler::STACK;
        v.ctx.uc_link = &s.main_ctx;
        makecontext(&v.ctx, fiber_entry, 0);
#endif
    }
    unsigned long stuck = 0;
    for (;;) {
        bool progressed = false;
        for (uint sg = 0; sg < nsub; ++sg) {
            const size_t lo = size_t(sg) * EMU_SUBGROUP_WIDTH, hi = (lo + EMU_SUBGROUP_WIDTH < n) ? lo + EMU_SUBGROUP_WIDTH : n;
            for (;;) {
                bool spinning = false;
                for (size_t i = lo; i < hi; ++i)
                    if (s.inv[i].state == INV_READY) {
                        resume(s, s.inv[i]);
                        if (s.inv[i].state == INV_YIELD) spinning = true; else progressed = true;
                    }
                /* a spinning invocation may still join the collective its subgroup is waiting in: come back later */
                if (spinning) break;
                /* everyone in the subgroup is now blocked or done: resolve a pending collective */
                uint sum = 0, ballot = 0, part = 0; bool any_waiting = false, first_seen = false;
                uint vals[32] = {0};
                for (size_t i = lo; i < hi; ++i) {
                    invocation& v = s.inv[i];
                    if (v.state != INV_AT_SUBGROUP) continue;
                    any_waiting = true;
                    v.sg_excl = sum;
                    vals[i - lo] = v.sg_in;
                    part |= 1u << (i - lo);
                    sum += v.sg_in;
                    if (v.sg_in) ballot |= 1u << (i - lo);
                    v.sg_first = !first_seen;
                    first_seen = true;
                }
                if (!any_waiting) break;
                progressed = true;
                for (size_t i = lo; i < hi; ++i) {
                    invocation& v = s.inv[i];
                    if (v.state != INV_AT_SUBGROUP) continue;
                    v.sg_sum = sum; v.sg_ballot = ballot; v.sg_part = part; v.state = INV_READY;
                    std::memcpy(v.sg_vals, vals, sizeof vals);
                }
            }
        }
        bool runnable = false, at_barrier = false;
        for (size_t i = 0; i < n; ++i) {
            if (s.inv[i].state == INV_YIELD) { s.inv[i].state = INV_READY; runnable = true; }
            else if (s.inv[i].state == INV_READY || s.inv[i].state == INV_AT_SUBGROUP) runnable = true;
            else if (s.inv[i].state == INV_AT_BARRIER) at_barrier = true;
        }
        if (!runnable) {
            if (!at_barrier) break; /* everybody is done */
            for (size_t i = 0; i < n; ++i)
                if (s.inv[i].state == INV_AT_BARRIER) s.inv[i].state = INV_READY;
            progressed = true;
        }
        if (progressed) stuck = 0;
        else if (++stuck > 1000000ul) { std::fprintf(stderr, "glsl_emu: no invocation makes progress (a spin-wait that nobody satisfies)\n"); std::abort(); }
    }
    s.cur = nullptr;
}

}  // namespace glsl

/* built-in variables of the current invocation */
#define gl_GlobalInvocationID (::glsl::sched().cur->global_id)
#define gl_LocalInvocationID (::glsl::sched().cur->local_id)
#define gl_WorkGroupID (::glsl::sched().cur->group_id)
#define gl_LocalInvocationIndex (::glsl::sched().cur->local_index)
#define gl_SubgroupInvocationID (::glsl::sched().cur->sg_invocation)
#define gl_SubgroupID (::glsl::sched().cur->sg_id)
#define gl_NumSubgroups (::glsl::sched().cur->num_subgroups)
#define gl_SubgroupSize (uint(::glsl::EMU_SUBGROUP_WIDTH))
