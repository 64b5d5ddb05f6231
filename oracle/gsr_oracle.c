/*
 * gsr_oracle.c -- CPU restatement of the 2Retr0/GodotGaussianSplatting forward-rasterizer hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference leg may load it.  The product path
 * (godotgaussiansplatting_b200/csrc -> libgsr.so) never links, imports or calls anything here.
 *
 * PARITY STATUS: the shader path is PINNED against the reference's own shader sources, executed on the CPU
 * (oracle/glsl_cpu -> oracle/_ref/libgsr_refshaders.so, tests/test_refshaders.py): projection records, keys,
 * values, M, the sort, the tile ranges and -- in the uncontracted mode, orc_set_blend_contraction(0) -- the
 * pixels are bit-identical.  The reference ships no tests, golden vectors or known-answer files and cannot be
 * executed as a whole here (Godot 4.3 + Vulkan), so the two GDScript host functions restated below
 * (orc_preprocess_ply, orc_pack_camera) remain "parity unpinned"; secondary evidence: a literal emulation of
 * the radix-sort shaders, an independent float64 transliteration (oracle/refmath_numpy.py) and the SURVEY.md
 * Appendix-B statistics on resources/demo.ply.
 *
 * Each function cites the reference file:line it follows (paths relative to /root/reference).
 *
 * Floating-point contract ("gsr deterministic math"):
 *   - every + - * / sqrt is one IEEE-754 binary32 operation, evaluated in the GLSL parse order
 *     (left-associative), NO implicit contraction (build with -ffp-contract=off);
 *   - fmaf() appears only where this file writes it explicitly (GLSL permits contracting a*b+c);
 *   - GLSL exp()/pow() are implementation-defined; this restatement fixes them to
 *     orc_exp()/orc_pow() below (range reduction + fixed polynomials, <= 2 ulp class, inside the
 *     Vulkan precision envelope), so that the CUDA kernels can reproduce them bit for bit;
 *   - min/max/clamp follow the GLSL definitions literally (orc_min/orc_max).
 *
 * Build: see oracle/Makefile  (gcc -O3 -march=x86-64-v3 -ffp-contract=off -fopenmp -shared -fPIC)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_TILE 16
#define ORC_WG   256 /* gsplat_render.glsl:9 WORKGROUP_SIZE */

/* ------------------------------------------------------------------------------------------ */
/* scalar helpers                                                                             */
/* ------------------------------------------------------------------------------------------ */
static inline float orc_max(float x, float y) { return (x < y) ? y : x; } /* GLSL max: y if x<y else x */
static inline float orc_min(float x, float y) { return (y < x) ? y : x; } /* GLSL min: y if y<x else x */
static inline float orc_clamp(float x, float lo, float hi) { return orc_min(orc_max(x, lo), hi); }

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* 2^t for any float t.  t is clamped to [-127,128]; n = round-half-even(t) via the 1.5*2^23 magic
 * constant; 2^f on f in [-0.5,0.5] by a degree-6 polynomial (Horner, fmaf); the scale 2^n is built
 * in the exponent field (n=-127 -> 0.0, n=128 -> +inf).  Pure bit/IEEE ops => reproducible. */
static inline float orc_exp2(float t) {
    const float MAGIC = 12582912.0f; /* 1.5 * 2^23 */
    float tc = orc_min(orc_max(t, -127.0f), 128.0f);
    float tm = tc + MAGIC;
    float nf = tm - MAGIC;
    float f = tc - nf;
    float p = 0x1.446c7ep-13f;
    p = fmaf(p, f, 0x1.5f48c8p-10f);
    p = fmaf(p, f, 0x1.3b29d8p-7f);
    p = fmaf(p, f, 0x1.c6aeccp-5f);
    p = fmaf(p, f, 0x1.ebfbe0p-3f);
    p = fmaf(p, f, 0x1.62e430p-1f);
    p = fmaf(p, f, 1.0f);
    uint32_t sbits = (f2u(tm) << 23) + 0x3F800000u;
    return p * u2f(sbits);
}

/* GLSL exp(x) := 2^(x * log2(e)) with log2(e) rounded to binary32 (what GPU drivers do). */
static inline float orc_exp(float x) { return orc_exp2(x * 0x1.715476p+0f); }

/* log2(x) for x > 0 (normal or subnormal).  x = m * 2^e with m in [sqrt(1/2), sqrt(2));
 * s = (m-1)/(m+1); log2(m) = s * (l0 + z(l1 + z(l2 + z l3))), z = s*s. */
static inline float orc_log2(float x) {
    int32_t eadj = 0;
    if (x < 0x1p-126f) { x = x * 0x1p+32f; eadj = -32; }
    uint32_t u = f2u(x);
    int32_t e = (int32_t)(u >> 23) - 127;
    uint32_t mb = (u & 0x007FFFFFu) | 0x3F800000u; /* m in [1,2) */
    float m = u2f(mb);
    if (m >= 0x1.6a09e6p+0f) { m = m * 0.5f; e += 1; } /* m >= sqrt(2) -> [sqrt(.5), 1) */
    float s = (m - 1.0f) / (m + 1.0f);
    float z = s * s;
    float g = 0x1.ba18b8p-2f;
    g = fmaf(g, z, 0x1.27471ep-1f);
    g = fmaf(g, z, 0x1.ec70e6p-1f);
    g = fmaf(g, z, 0x1.715476p+1f);
    return (float)(e + eadj) + s * g;
}

/* GLSL pow(x, y) := exp2(y * log2(x)) for x > 0; pow(x<=0, y>0) := 0 (only x == 0 can occur). */
static inline float orc_pow(float x, float y) {
    if (!(x > 0.0f)) return 0.0f;
    return orc_exp2(y * orc_log2(x));
}

/* mat3 stored like GLSL: m[c][r] (column-major). */
typedef struct { float m[3][3]; } mat3;

/* GLSL a*b for mat3: result[c][r] = sum_k a[k][r]*b[c][k], summed left to right. */
static inline mat3 mat3_mul(const mat3 *a, const mat3 *b) {
    mat3 o;
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r)
            o.m[c][r] = (a->m[0][r] * b->m[c][0] + a->m[1][r] * b->m[c][1]) + a->m[2][r] * b->m[c][2];
    return o;
}
static inline mat3 mat3_transpose(const mat3 *a) {
    mat3 o;
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) o.m[c][r] = a->m[r][c];
    return o;
}

/* ------------------------------------------------------------------------------------------ */
/* records                                                                                    */
/* ------------------------------------------------------------------------------------------ */
/* RasterizeData, std430, 48 B (gsplat_projection.glsl:42-48, gsplat_render.glsl:13-19) */
typedef struct {
    float image_pos[2];
    float pos_xy[2];
    float conic[3];
    float pos_z;
    float color[4];
} orc_record;

/* Uniforms block, std140, 32 B (gsplat_projection.glsl:75-80; written by rasterizer.gd:126) */
typedef struct {
    float camera_pos[3];
    float model_scale;
    int32_t dims[2];
    float time;
    float _pad;
} orc_uniforms;

/* ------------------------------------------------------------------------------------------ */
/* ingest: util/ply_file.gd:44-69  (62-float PLY vertex -> 60-float std430 Splat)              */
/* ------------------------------------------------------------------------------------------ */
/* Godot Basis(Quaternion) (engine core/math/basis.cpp set_quaternion, float32 real_t; stated from
 * knowledge of Godot 4.x -- the engine source is not vendored in the reference). rows[r][c]. */
static void godot_basis_from_quat(float qx, float qy, float qz, float qw, float rows[3][3]) {
    float d = ((qx * qx + qy * qy) + qz * qz) + qw * qw;
    float s = 2.0f / d;
    float xs = qx * s, ys = qy * s, zs = qz * s;
    float wx = qw * xs, wy = qw * ys, wz = qw * zs;
    float xx = qx * xs, xy = qx * ys, xz = qx * zs;
    float yy = qy * ys, yz = qy * zs, zz = qz * zs;
    rows[0][0] = 1.0f - (yy + zz); rows[0][1] = xy - wz;          rows[0][2] = xz + wy;
    rows[1][0] = xy + wz;          rows[1][1] = 1.0f - (xx + zz); rows[1][2] = yz - wx;
    rows[2][0] = xz - wy;          rows[2][1] = yz + wx;          rows[2][2] = 1.0f - (xx + yy);
}
/* Godot Basis*Basis: (A*B)[i][j] = B[0][j]*A[i][0] + B[1][j]*A[i][1] + B[2][j]*A[i][2] (tdot order) */
static void godot_basis_mul(const float a[3][3], const float b[3][3], float o[3][3]) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            o[i][j] = (b[0][j] * a[i][0] + b[1][j] * a[i][1]) + b[2][j] * a[i][2];
}

/* ply62: vertices with `nprops` float properties in the standard 3DGS order
 * (x,y,z,nx,ny,nz,f_dc_0..2,f_rest_0..44,opacity,scale_0..2,rot_0..3); splat60: 60 floats each. */
void orc_preprocess_ply(const float *ply, int64_t n, int nprops, float creation_time, float *splat60) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const float *p = ply + (size_t)i * nprops;
        float *o = splat60 + (size_t)i * 60;
        /* position + creation time (ply_file.gd:46-47) */
        o[0] = p[0]; o[1] = p[1]; o[2] = p[2]; o[3] = creation_time;
        /* covariance (ply_file.gd:49-59): exp() is GDScript float64, narrowed into Vector3 */
        float sc[3] = {(float)exp((double)p[55]), (float)exp((double)p[56]), (float)exp((double)p[57])};
        float S[3][3] = {{sc[0], 0, 0}, {0, sc[1], 0}, {0, 0, sc[2]}};
        float B[3][3], R[3][3], M[3][3], Mt[3][3], C[3][3];
        godot_basis_from_quat(p[59], p[60], p[61], p[58], B); /* Quaternion(rot_1,rot_2,rot_3,rot_0) */
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R[r][c] = B[c][r]; /* .transposed() */
        godot_basis_mul(S, R, M);
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Mt[r][c] = M[c][r];
        godot_basis_mul(Mt, M, C);
        /* cov.x[0], cov.y[0], cov.z[0], cov.y[1], cov.z[1], cov.z[2]; Basis.x is a COLUMN */
        o[4] = C[0][0]; o[5] = C[0][1]; o[6] = C[0][2]; o[7] = C[1][1]; o[8] = C[1][2]; o[9] = C[2][2];
        /* opacity (ply_file.gd:62), float64 then narrowed */
        o[10] = (float)(1.0 / (1.0 + exp(-(double)p[54])));
        o[11] = 0.0f;
        /* SH (ply_file.gd:65-69): DC then f_rest re-interleaved to coefficient-major RGB */
        o[12] = p[6]; o[13] = p[7]; o[14] = p[8];
        for (int k = 0; k < 15; ++k) {
            o[15 + 3 * k + 0] = p[9 + k];
            o[15 + 3 * k + 1] = p[9 + 15 + k];
            o[15 + 3 * k + 2] = p[9 + 30 + k];
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* camera packing: util/gaussian_splatting_rasterizer.gd:175-195                               */
/* ------------------------------------------------------------------------------------------ */
/* cam: camera global transform as Godot Projection columns x,y,z,w (each 4 floats, column-major
 * 16 floats, w = origin); proj: Godot Projection columns x,y,z,w (16 floats). out: 32 floats. */
void orc_pack_camera(const float *cam, const float *proj, float *out) {
    const float *x = cam, *y = cam + 4, *z = cam + 8, *w = cam + 12;
    /* Vector4.dot: x*x + y*y + z*z + w*w left to right (w components are 0 for axes) */
    float wdx = ((w[0] * x[0] + w[1] * x[1]) + w[2] * x[2]) + w[3] * x[3];
    float wdny = ((w[0] * -y[0] + w[1] * -y[1]) + w[2] * -y[2]) + w[3] * -y[3];
    float wdz = ((w[0] * z[0] + w[1] * z[1]) + w[2] * z[2]) + w[3] * z[3];
    float v[16] = {-x[0], y[0], -z[0], 0.0f, -x[1], y[1], -z[1], 0.0f,
                   x[2],  -y[2], z[2], 0.0f, -wdx,  -wdny, -wdz, 1.0f};
    memcpy(out, v, sizeof v);
    const float *px = proj, *py = proj + 4, *pz = proj + 8, *pw = proj + 12;
    float q[16] = {px[0], px[1], px[2], 0.0f, py[0], py[1], py[2], 0.0f,
                   pz[0], pz[1], pz[2], -1.0f, pw[0], pw[1], pw[2], 0.0f};
    memcpy(out + 16, q, sizeof q);
}

/* ------------------------------------------------------------------------------------------ */
/* a1: gsplat_projection.glsl:150-227                                                          */
/* ------------------------------------------------------------------------------------------ */
#define SH_C0 0.28209479177387814f
#define SH_C1 0.4886025119029199f
#define SH_C2_0 1.0925484305920792f
#define SH_C2_1 1.0925484305920792f
#define SH_C2_2 0.31539156525252005f
#define SH_C2_3 1.0925484305920792f
#define SH_C2_4 0.5462742152960396f
#define SH_C3_0 0.5900435899266435f
#define SH_C3_1 2.890611442640554f
#define SH_C3_2 0.4570457994644658f
#define SH_C3_3 0.3731763325901154f
#define SH_C3_4 0.4570457994644658f
#define SH_C3_5 1.445305721320277f
#define SH_C3_6 0.5900435899266435f

/* gsplat_projection.glsl:94-121 -- one colour channel `ch` of get_color */
static inline float sh_channel(const float *sh, int ch, float x, float y, float z) {
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
#define SHC(k) (sh[3 * (k) + ch])
    float r = 0.5f + SHC(0) * SH_C0;
    r = r - SHC(1) * SH_C1 * y;
    r = r + SHC(2) * SH_C1 * z;
    r = r - SHC(3) * SH_C1 * x;
    r = r + SHC(4) * SH_C2_0 * xy;
    r = r - SHC(5) * SH_C2_1 * yz;
    r = r + SHC(6) * SH_C2_2 * (2.0f * zz - xx - yy);
    r = r - SHC(7) * SH_C2_3 * xz;
    r = r + SHC(8) * SH_C2_4 * (xx - yy);
    r = r - SHC(9) * SH_C3_0 * y * (3.0f * xx - yy);
    r = r + SHC(10) * SH_C3_1 * x * yz;
    r = r - SHC(11) * SH_C3_2 * y * (4.0f * zz - xx - yy);
    r = r + SHC(12) * SH_C3_3 * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
    r = r - SHC(13) * SH_C3_4 * x * (4.0f * zz - xx - yy);
    r = r + SHC(14) * SH_C3_5 * z * (xx - yy);
    r = r - SHC(15) * SH_C3_6 * x * (xx - 3.0f * yy);
#undef SHC
    return orc_max(0.0f, r);
}

static inline float ease_out_cubic(float x) { /* gsplat_projection.glsl:87-90 */
    float a = 1.0f - x;
    return 1.0f - a * a * a;
}

/* per-splat result of the projection stage */
typedef struct {
    uint32_t rect[4]; /* x0,y0,x1,y1 in tiles (after optional band clamp on y) */
    uint32_t depth;
    uint32_t ntiles;  /* 0 => culled */
} orc_proj;

/* One invocation of gsplat_projection.glsl:main up to (not including) the atomicAdd.
 * vp: 32 floats (view_matrix, projection_matrix; GLSL column-major).  band_y0/band_y1: tile-row
 * band [y0,y1) owned by this rank (multi-GPU tile-row sharding; full frame = [0, grid_y)). */
static void project_one(const float *s, const float *vp, const orc_uniforms *u, int band_y0, int band_y1,
                        orc_record *rec, orc_proj *out) {
    const float *V = vp, *P = vp + 16; /* X[c][r] = X[4*c + r] */
    const int W = u->dims[0], H = u->dims[1];
    const uint32_t gx = (uint32_t)((W + ORC_TILE - 1) / ORC_TILE), gy = (uint32_t)((H + ORC_TILE - 1) / ORC_TILE);
    const float ms = u->model_scale;
    out->ntiles = 0;

    /* :158-166 frustum cull */
    float sp[3] = {s[0] * ms, s[1] * ms, s[2] * ms};
    float view[4], clip[4];
    for (int r = 0; r < 4; ++r) view[r] = ((V[0 + r] * sp[0] + V[4 + r] * sp[1]) + V[8 + r] * sp[2]) + V[12 + r] * 1.0f;
    for (int r = 0; r < 4; ++r) clip[r] = ((P[0 + r] * view[0] + P[4 + r] * view[1]) + P[8 + r] * view[2]) + P[12 + r] * view[3];
    float vb = clip[3] * 1.2f;
    if (clip[0] < -vb || clip[1] < -vb || clip[2] < 0.0f || clip[0] > vb || clip[1] > vb || clip[2] > clip[3]) return;

    /* :169-174 load-in animation */
    float splat_time = u->time - s[3];
    float tf = ease_out_cubic(orc_clamp(splat_time, 0.0f, 1.0f));
    float tfl = ease_out_cubic(orc_clamp(splat_time - 0.35f, 0.0f, 1.0f));
    float splat_opacity = s[10] * tfl * tfl;
    float splat_scale = ms * (2.0f * (1.0f - tfl) + 1.0f * tfl); /* mix(2.0, 1.0, tfl) */

    /* :124-142 project_covariance */
    const float *c = s + 4;
    mat3 cov3 = {{{c[0], c[1], c[2]}, {c[1], c[3], c[4]}, {c[2], c[4], c[5]}}};
    for (int cc = 0; cc < 3; ++cc) for (int r = 0; r < 3; ++r) cov3.m[cc][r] = cov3.m[cc][r] * splat_scale * splat_scale;
    float tfi[2] = {P[0], P[5]};
    float focal[2] = {((float)W * 0.5f) * tfi[0], ((float)H * 0.5f) * tfi[1]};
    float tanfov[2] = {1.0f / tfi[0], 1.0f / tfi[1]};
    float z_inv = 1.0f / view[2];
    focal[0] *= z_inv; focal[1] *= z_inv;
    float mx = orc_clamp(view[0] * z_inv, -tanfov[0] * 1.3f, tanfov[0] * 1.3f);
    float my = orc_clamp(view[1] * z_inv, -tanfov[1] * 1.3f, tanfov[1] * 1.3f);
    /* jacobian = mat3(focal.x, 0, -focal.y*mean.x,  0, focal.y, -focal.y*mean.y,  0, 0, 0)   (:134-137)
     * b = transpose(mat3(view_matrix)) * jacobian;  cov_2d = transpose(b) * cov_3d * b              (:138-140)
     * gsr spec: the structurally-zero terms of b are skipped (b[0][r] = V[r][0]*J00 + V[r][2]*J02,
     * b[1][r] = V[r][1]*J11 + V[r][2]*J12, b[2] = 0) and only the three entries :141 reads are formed;
     * every remaining sum is left-to-right as mat3_mul would do it. */
    float j02 = -focal[1] * mx, j12 = -focal[1] * my;
    float B0[3], B1[3], T0[3], T1[3];
    for (int r = 0; r < 3; ++r) {
        B0[r] = V[4 * r + 0] * focal[0] + V[4 * r + 2] * j02;
        B1[r] = V[4 * r + 1] * focal[1] + V[4 * r + 2] * j12;
    }
    for (int cc = 0; cc < 3; ++cc) { /* t1 = transpose(b) * cov_3d: T0[c] = t1[c][0], T1[c] = t1[c][1] */
        T0[cc] = (B0[0] * cov3.m[cc][0] + B0[1] * cov3.m[cc][1]) + B0[2] * cov3.m[cc][2];
        T1[cc] = (B1[0] * cov3.m[cc][0] + B1[1] * cov3.m[cc][1]) + B1[2] * cov3.m[cc][2];
    }
    float c2_00 = (T0[0] * B0[0] + T0[1] * B0[1]) + T0[2] * B0[2];
    float c2_01 = (T1[0] * B0[0] + T1[1] * B0[1]) + T1[2] * B0[2];
    float c2_11 = (T1[0] * B1[0] + T1[1] * B1[1]) + T1[2] * B1[2];
    float cx = c2_00 + 0.3f, cy = c2_01, cz = c2_11 + 0.3f;

    /* :177-182 */
    float det = cx * cz - cy * cy;
    if (det == 0.0f) return;
    float mid = 0.5f * (cx + cz);
    float sq = sqrtf(orc_max(0.1f, mid * mid - det));
    float e1 = mid + 1.0f * sq, e2 = mid + -1.0f * sq;
    if (e1 < 0.0f || e2 < 0.0f) return;

    /* :184-185 */
    float ndc[3] = {clip[0] / clip[3], clip[1] / clip[3], clip[2] / clip[3]};
    float ipx = ((ndc[0] + 1.0f) * 0.5f - 1.0f * (1.0f - tf)) * (float)(W - 1);
    float ipy = ((ndc[1] + 1.0f) * 0.5f - 0.75f * (1.0f - tf)) * (float)(H - 1);

    /* :190-194 */
    float radius = orc_pow(splat_opacity, 0.2f) * 2.5f * sqrtf(orc_max(e1, e2));
    /* gsr spec: non-finite image_pos/radius (reference behaviour undefined: int(NaN)) => culled */
    if (!(fabsf(ipx) <= 3.0e38f) || !(fabsf(ipy) <= 3.0e38f) || !(radius <= 3.0e38f)) return;
    float fgx = (float)gx, fgy = (float)gy;
    int32_t x0 = (int32_t)orc_clamp((ipx - radius) / 16.0f, 0.0f, fgx);
    int32_t y0 = (int32_t)orc_clamp((ipy - radius) / 16.0f, 0.0f, fgy);
    int32_t x1 = (int32_t)orc_clamp(ceilf((ipx + radius) / 16.0f), 0.0f, fgx);
    int32_t y1 = (int32_t)orc_clamp(ceilf((ipy + radius) / 16.0f), 0.0f, fgy);
    /* multi-GPU band clamp (no-op for the full frame) */
    if (y0 < band_y0) y0 = band_y0;
    if (y1 > band_y1) y1 = band_y1;
    if (y1 < y0) y1 = y0;
    uint32_t n = (uint32_t)(x1 - x0) * (uint32_t)(y1 - y0);
    if (n == 0) return;

    /* :198-206 */
    float d[3] = {sp[0] - u->camera_pos[0], sp[1] - u->camera_pos[1], sp[2] - u->camera_pos[2]};
    float inv_len = 1.0f / sqrtf((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]); /* normalize := v * (1/|v|) */
    float dx = d[0] * inv_len, dy = d[1] * inv_len, dz = d[2] * inv_len;
    rec->image_pos[0] = ipx; rec->image_pos[1] = ipy;
    rec->conic[0] = cz / det; rec->conic[1] = -cy / det; rec->conic[2] = cx / det;
    const float *sh = s + 12;
    rec->color[0] = sh_channel(sh, 0, dx, dy, dz);
    rec->color[1] = sh_channel(sh, 1, dx, dy, dz);
    rec->color[2] = sh_channel(sh, 2, dx, dy, dz);
    rec->color[3] = splat_opacity;
    rec->pos_xy[0] = sp[0]; rec->pos_xy[1] = sp[1]; rec->pos_z = sp[2];

    /* :218 */
    out->depth = ((uint32_t)(ndc[2] * ndc[2] * ndc[2] * 65535.0f)) & 0xFFFFu;
    out->rect[0] = (uint32_t)x0; out->rect[1] = (uint32_t)y0; out->rect[2] = (uint32_t)x1; out->rect[3] = (uint32_t)y1;
    out->ntiles = n;
}

/* Projection + key duplication.  Emission order (Q13: nondeterministic in the reference because of
 * the atomicAdd at :196) is fixed to ascending splat id, row-major within a splat's rect (:219-226).
 * records: 48*n bytes indexed by splat id (culled entries untouched).  keys/values: capacity `cap`.
 * Returns M (total duplicates, may exceed cap: entries >= cap are not written).
 * visible_out (nullable): number of splats with ntiles > 0.  max_tile_out (nullable): largest tile id
 * touched by the UNBANDED rect of any visible-in-band-or-not splat (for sharded Q10 handling). */
int64_t orc_project(const float *splat60, int64_t n, const float *vp, const orc_uniforms *u, int band_y0, int band_y1,
                    orc_record *records, uint32_t *keys, uint32_t *values, int64_t cap, int64_t *visible_out,
                    int64_t *last_tile_out) {
    const uint32_t gx = (uint32_t)((u->dims[0] + ORC_TILE - 1) / ORC_TILE);
    const int gy = (u->dims[1] + ORC_TILE - 1) / ORC_TILE;
    orc_proj *pr = (orc_proj *)malloc(sizeof(orc_proj) * (size_t)(n > 0 ? n : 1));
    int64_t *off = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + 1));
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) project_one(splat60 + (size_t)i * 60, vp, u, band_y0, band_y1, records + i, pr + i);
    int64_t m = 0, vis = 0;
    for (int64_t i = 0; i < n; ++i) { off[i] = m; m += pr[i].ntiles; vis += pr[i].ntiles != 0; }
    off[n] = m;
#pragma omp parallel for schedule(dynamic, 4096)
    for (int64_t i = 0; i < n; ++i) {
        if (!pr[i].ntiles) continue;
        int64_t o = off[i];
        for (uint32_t y = pr[i].rect[1]; y < pr[i].rect[3]; ++y)
            for (uint32_t x = pr[i].rect[0]; x < pr[i].rect[2]; ++x) {
                uint32_t tile_id = y * gx + x;
                if (o < cap) { keys[o] = (tile_id << 16) | pr[i].depth; values[o] = (uint32_t)i; }
                ++o;
            }
    }
    if (visible_out) *visible_out = vis;
    if (last_tile_out) {
        /* global last occupied tile: needs the un-banded rects, so re-project with the full band */
        int64_t last = -1;
        if (band_y0 == 0 && band_y1 >= gy) {
            for (int64_t i = 0; i < n; ++i)
                if (pr[i].ntiles) { int64_t t = (int64_t)(pr[i].rect[3] - 1) * gx + (pr[i].rect[2] - 1); if (t > last) last = t; }
        } else {
#pragma omp parallel
            {
                int64_t l = -1; orc_record tmp; orc_proj q;
#pragma omp for schedule(static) nowait
                for (int64_t i = 0; i < n; ++i) {
                    project_one(splat60 + (size_t)i * 60, vp, u, 0, gy, &tmp, &q);
                    if (q.ntiles) { int64_t t = (int64_t)(q.rect[3] - 1) * gx + (q.rect[2] - 1); if (t > l) l = t; }
                }
#pragma omp critical
                { if (l > last) last = l; }
            }
        }
        *last_tile_out = last;
    }
    free(pr); free(off);
    return m;
}

/* ------------------------------------------------------------------------------------------ */
/* a2-a4: radix sort.  Semantics of radix_sort_{upsweep,spine,downsweep}.glsl x 4 passes        */
/* (rasterizer.gd:143-149) = stable LSD sort of (key,value) pairs on all 32 key bits.           */
/* ------------------------------------------------------------------------------------------ */
void orc_sort_pairs(uint32_t *keys, uint32_t *values, int64_t n) {
    if (n <= 1) return;
    uint32_t *k2 = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)n);
    uint32_t *v2 = values ? (uint32_t *)malloc(sizeof(uint32_t) * (size_t)n) : NULL;
    uint32_t *ki = keys, *ko = k2, *vi = values, *vo = v2;
    int nt = 1;
#ifdef _OPENMP
    nt = omp_get_max_threads();
#endif
    if (n < 1 << 16) nt = 1;
    int64_t *hist = (int64_t *)malloc(sizeof(int64_t) * 256 * (size_t)nt);
    for (int pass = 0; pass < 4; ++pass) {
        const int sh = 8 * pass;
        memset(hist, 0, sizeof(int64_t) * 256 * (size_t)nt);
#pragma omp parallel num_threads(nt)
        {
            int t = 0, nth = 1;
#ifdef _OPENMP
            t = omp_get_thread_num();
            nth = omp_get_num_threads(); /* may be fewer than requested */
#endif
            int64_t lo = n * t / nth, hi = n * (t + 1) / nth;
            int64_t *h = hist + 256 * t;
            for (int64_t i = lo; i < hi; ++i) h[(ki[i] >> sh) & 255]++;
#pragma omp barrier
#pragma omp single
            {
                int64_t run = 0;
                for (int d = 0; d < 256; ++d)
                    for (int tt = 0; tt < nth; ++tt) { int64_t c = hist[256 * tt + d]; hist[256 * tt + d] = run; run += c; }
            }
            for (int64_t i = lo; i < hi; ++i) {
                int64_t dst = h[(ki[i] >> sh) & 255]++;
                ko[dst] = ki[i];
                if (vi) vo[dst] = vi[i];
            }
        }
        uint32_t *tk = ki; ki = ko; ko = tk;
        uint32_t *tv = vi; vi = vo; vo = tv;
    }
    /* 4 passes: result is back in the caller's arrays */
    free(hist); free(k2); free(v2);
}

/* Literal emulation of the three vendored shaders, workgroup by workgroup, subgroup (32 lanes) by
 * subgroup, including the 0xffffffff pad keys, the `dst < element_count` guard on keys only
 * (radix_sort_downsweep.glsl:195) and the ping-pong offsets of rasterizer.gd:145.
 * keys/values: 2*cap uints each (both halves); n <= cap.  Result in half 0.  Small n only. */
void orc_sort_pairs_shader_emulation(uint32_t *keys, uint32_t *values, int64_t n, int64_t cap) {
    enum { RADIX = 256, WG = 512, PD = 8, PS = PD * WG, NSG = 16, SG = 32 };
    const int64_t parts_cap = (cap + PS - 1) / PS;
    uint32_t *global_hist = (uint32_t *)calloc(4 * RADIX, 4);           /* cleared by rasterizer.gd:127 */
    uint32_t *part_hist = (uint32_t *)calloc((size_t)(parts_cap + 1) * RADIX, 4);
    const uint32_t element_count = (uint32_t)n;
    const int64_t nparts = (n + PS - 1) / PS; /* grid_dims[0] (gsplat_projection.glsl:212) */
    uint32_t *local_histogram = (uint32_t *)malloc(PS * 4);
    uint32_t *local_histogram_sum = (uint32_t *)malloc(RADIX * 4 * 2);
    for (int pass = 0; pass < 4; ++pass) {
        const uint32_t in_offset = (uint32_t)(cap * (pass % 2)), out_offset = (uint32_t)(cap * (1 - (pass % 2)));
        /* --- upsweep (radix_sort_upsweep.glsl:35-65) --- */
        for (int64_t part = 0; part < nparts; ++part) {
            uint32_t pstart = (uint32_t)(part * PS);
            if (pstart >= element_count) continue;
            uint32_t lh[RADIX]; memset(lh, 0, sizeof lh);
            for (int i = 0; i < PD; ++i)
                for (uint32_t index = 0; index < WG; ++index) {
                    uint32_t key_index = pstart + WG * i + index;
                    uint32_t key = key_index < element_count ? keys[key_index + in_offset] : 0xffffffffu;
                    lh[(key >> (8 * pass)) & 255]++;
                }
            for (int d = 0; d < RADIX; ++d) { part_hist[RADIX * part + d] = lh[d]; global_hist[RADIX * pass + d] += lh[d]; }
        }
        /* --- spine (radix_sort_spine.glsl:35-92) --- */
        {
            uint32_t partition_count = (element_count + PS - 1) / PS;
            for (int radix = 0; radix < RADIX; ++radix) {
                uint32_t reduction = 0;
                for (uint32_t p = 0; p < partition_count; ++p) { uint32_t v = part_hist[RADIX * p + radix]; part_hist[RADIX * p + radix] = reduction; reduction += v; }
            }
            uint32_t run = 0;
            for (int d = 0; d < RADIX; ++d) { uint32_t v = global_hist[RADIX * pass + d]; global_hist[RADIX * pass + d] = run; run += v; }
        }
        /* --- downsweep (radix_sort_downsweep.glsl:59-214) --- */
        for (int64_t part = 0; part < nparts; ++part) {
            uint32_t pstart = (uint32_t)(part * PS);
            if (pstart >= element_count) continue;
            static uint32_t local_keys[WG][PD], local_values[WG][PD], local_radix[WG][PD], local_offsets[WG][PD], sg_hist[WG][PD];
            memset(local_histogram, 0, PS * 4);
            for (int sgi = 0; sgi < NSG; ++sgi)
                for (int i = 0; i < PD; ++i) {
                    uint32_t rk[SG];
                    for (int lane = 0; lane < SG; ++lane) {
                        int index = sgi * SG + lane;
                        uint32_t key_index = pstart + (PD * SG) * sgi + i * SG + lane;
                        uint32_t key = key_index < element_count ? keys[key_index + in_offset] : 0xffffffffu;
                        local_keys[index][i] = key;
                        local_values[index][i] = key_index < element_count ? values[key_index + in_offset] : 0;
                        rk[lane] = (key >> (8 * pass)) & 255;
                        local_radix[index][i] = rk[lane];
                    }
                    for (int lane = 0; lane < SG; ++lane) {
                        /* 8 ballots => mask of lanes with the same digit (:95-102) */
                        uint32_t mask = 0;
                        for (int l2 = 0; l2 < SG; ++l2) if (rk[l2] == rk[lane]) mask |= 1u << l2;
                        uint32_t subgroup_offset = (uint32_t)__builtin_popcount(mask & ((1u << lane) - 1));
                        uint32_t radix_count = (uint32_t)__builtin_popcount(mask);
                        int index = sgi * SG + lane;
                        if (subgroup_offset == 0) { local_histogram[NSG * rk[lane] + sgi] += radix_count; sg_hist[index][i] = radix_count; }
                        else sg_hist[index][i] = 0;
                        local_offsets[index][i] = subgroup_offset;
                    }
                }
            /* :121-163 exclusive scan of the 4096 (radix, subgroup) counters */
            { uint32_t run = 0; for (int k = 0; k < RADIX * NSG; ++k) { uint32_t v = local_histogram[k]; local_histogram[k] = run; run += v; } }
            /* :166-175 post-scan, row by row with barriers => rows accumulate in order */
            for (int i = 0; i < PD; ++i) {
                for (int index = 0; index < WG; ++index) local_offsets[index][i] += local_histogram[NSG * local_radix[index][i] + index / SG];
                for (int index = 0; index < WG; ++index) if (sg_hist[index][i] > 0) local_histogram[NSG * local_radix[index][i] + index / SG] += sg_hist[index][i];
            }
            /* :178-181 */
            for (int index = 0; index < RADIX; ++index) {
                uint32_t v = index == 0 ? 0 : local_histogram[NSG * index - 1];
                local_histogram_sum[index] = global_hist[RADIX * pass + index] + part_hist[RADIX * part + index] - v;
            }
            /* :186-201 keys */
            static uint32_t dsts[PS];
            for (int index = 0; index < WG; ++index) for (int i = 0; i < PD; ++i) local_histogram[local_offsets[index][i]] = local_keys[index][i];
            for (uint32_t i = 0; i < PS; ++i) {
                uint32_t key = local_histogram[i];
                uint32_t dst = local_histogram_sum[(key >> (8 * pass)) & 255] + i;
                if (dst < element_count) keys[dst + out_offset] = key;
                dsts[i] = dst;
            }
            /* :205-213 values (no guard in the reference; guard here only against leaving the buffer) */
            for (int index = 0; index < WG; ++index) for (int i = 0; i < PD; ++i) local_histogram[local_offsets[index][i]] = local_values[index][i];
            for (uint32_t i = 0; i < PS; ++i) if ((int64_t)dsts[i] < cap) values[dsts[i] + out_offset] = local_histogram[i];
        }
    }
    free(global_hist); free(part_hist); free(local_histogram); free(local_histogram_sum);
}

/* ------------------------------------------------------------------------------------------ */
/* a5: gsplat_boundaries.glsl:23-50                                                            */
/* ------------------------------------------------------------------------------------------ */
/* bounds: 2*T uints, cleared to 0 first (rasterizer.gd:128).
 * quirks != 0: reference behaviour (Q10).  The "every thread with tile_id == T-1 writes
 * bounds[T-1].y = M-1" store races with the boundary thread's `.y` store only for tile_id_prev,
 * never for T-1 itself, so the outcome is deterministic.
 * quirks == 0: corrected ranges (every occupied tile gets [start,end)).
 * global_last_tile >= 0 (sharded runs): the tile that is last-occupied over ALL bands; the
 * band-local last occupied tile gets its end written unless it is that global tile. */
void orc_boundaries(const uint32_t *keys, int64_t m, int64_t T, uint32_t *bounds, int quirks, int64_t global_last_tile) {
    memset(bounds, 0, sizeof(uint32_t) * 2 * (size_t)T);
    if (m <= 0) return;
    for (int64_t id = 1; id < m; ++id) {
        uint32_t a = keys[id - 1] >> 16, b = keys[id] >> 16;
        if (a != b) { bounds[2 * a + 1] = (uint32_t)id; bounds[2 * b + 0] = (uint32_t)id; }
    }
    uint32_t last = keys[m - 1] >> 16;
    if (quirks) {
        if ((int64_t)last == T - 1) {
            /* :47-49 -- any id>=1 whose tile is T-1 stores M-1 (needs at least one such thread) */
            int64_t first_of_last = bounds[2 * last + 0];
            int has_thread = (m - 1 >= 1) && (m - 1 >= first_of_last);
            if (has_thread) bounds[2 * last + 1] = (uint32_t)(m - 1);
        } else if (global_last_tile >= 0 && (int64_t)last != global_last_tile) {
            bounds[2 * last + 1] = (uint32_t)m; /* sharded: this band's last tile is not the frame's last */
        }
    } else {
        bounds[2 * last + 1] = (uint32_t)m;
    }
}

/* Q20 (found by running the shader under oracle/glsl_cpu): gsplat_boundaries.glsl:27 makes invocation 0 of
 * workgroup 0 return before its load at :33, so the word `local[1]` that invocation id = 1 reads as its left
 * neighbour (:36) is never written -- an uninitialised `shared` read.  With that word = G the shader does
 * `bounds[G].y = 1` (dropped when G >= T) and `bounds[keys[1]>>16].x = 1` whenever G != keys[1]>>16, i.e. the
 * front-most instance of the first occupied tile is lost.  orc_boundaries() above defines the word as the
 * author evidently meant it (G = keys[0]>>16); this variant reproduces the shader for any G, stores applied
 * in ascending invocation order (the order oracle/glsl_cpu runs them).  Reference quirks (Q10) included. */
void orc_boundaries_uninit(const uint32_t *keys, int64_t m, int64_t T, uint32_t *bounds, uint32_t garbage) {
    memset(bounds, 0, sizeof(uint32_t) * 2 * (size_t)T);
    for (int64_t id = 1; id < m; ++id) {
        uint32_t a = (id == 1) ? garbage : (keys[id - 1] >> 16), b = keys[id] >> 16;
        if (a != b) {
            if ((int64_t)a < T) bounds[2 * a + 1] = (uint32_t)id;
            if ((int64_t)b < T) bounds[2 * b + 0] = (uint32_t)id;
        }
        if ((int64_t)b == T - 1) bounds[2 * b + 1] = (uint32_t)(m - 1); /* :47-49 */
    }
}

/* ------------------------------------------------------------------------------------------ */
/* a6: gsplat_render.glsl:50-111                                                               */
/* ------------------------------------------------------------------------------------------ */
/* out: W*H*4 floats (row-major, RGBA).  pick (nullable, 4 floats: splat_pos.xyz, num_tile_splats):
 * written only when the reference would write it (:105-110), otherwise left untouched.
 * tile_y0/tile_y1: tile-row band to render (pixels outside the band are not touched).
 * staged_out (nullable): C = sum over tiles and consumed chunks of chunk_size (SURVEY 8 symbol C). */
/* 1 (default): the gsr spec's five explicit contractions in the blend; 0: none (see orc_render) */
static int g_blend_contraction = 1;
void orc_set_blend_contraction(int on) { g_blend_contraction = on != 0; }

void orc_render(const orc_record *records, const uint32_t *values, const uint32_t *bounds, int W, int H, float heatmap_factor,
                uint32_t target_tile_id, int tile_y0, int tile_y1, float *out, float *pick, int64_t *staged_out,
                uint32_t *tile_staged /* nullable: per-tile consumed instance count */) {
    const int gx = (W + ORC_TILE - 1) / ORC_TILE;
    const float MIN_ALPHA = 1.0f / 255.0f;
    int64_t staged_total = 0;
#pragma omp parallel for schedule(dynamic, 1) collapse(2) reduction(+ : staged_total)
    for (int ty = tile_y0; ty < tile_y1; ++ty)
        for (int tx = 0; tx < gx; ++tx) {
            const uint32_t tile_id = (uint32_t)(ty * gx + tx);
            const uint32_t bx = bounds[2 * tile_id], by = bounds[2 * tile_id + 1];
            int32_t diff = (int32_t)(by - bx);
            const int num_splats = diff > 0 ? diff : 0;
            const int num_iterations = (int)ceilf((float)num_splats / 256.0f);
            float col[ORC_WG][3], t[ORC_WG];
            for (int l = 0; l < ORC_WG; ++l) { col[l][0] = col[l][1] = col[l][2] = 0.0f; t[l] = 1.0f; }
            uint32_t shared_t = 0xFFFFFFFFu, tile_consumed = 0;
            for (int i = 0; i < num_iterations && shared_t > 255u; ++i) {
                const int sort_offset = ORC_WG * i;
                const int chunk = (num_splats - sort_offset) < ORC_WG ? (num_splats - sort_offset) : ORC_WG;
                staged_total += chunk;
                tile_consumed += (uint32_t)chunk;
                shared_t = 0;
                for (int l = 0; l < ORC_WG; ++l) {
                    const float px = (float)(tx * ORC_TILE + (l & 15)), py = (float)(ty * ORC_TILE + (l >> 4));
                    float tt = t[l], r = col[l][0], g = col[l][1], b = col[l][2];
                    if (!g_blend_contraction) {
                        /* strict evaluation of :84-90, no contraction anywhere: what the reference's own shader
                         * text gives under oracle/glsl_cpu (tests/test_refshaders.py compares bit for bit) */
                        for (int j = 0; j < chunk && tt > MIN_ALPHA; ++j) {
                            const orc_record *s = &records[values[bx + (uint32_t)sort_offset + (uint32_t)j]];
                            float ox = s->image_pos[0] - px, oy = s->image_pos[1] - py;
                            float power = -0.5f * (s->conic[0] * ox * ox + s->conic[2] * oy * oy) - s->conic[1] * ox * oy;
                            float alpha = s->color[3] * orc_exp(power);
                            r = r + s->color[0] * alpha * tt;
                            g = g + s->color[1] * alpha * tt;
                            b = b + s->color[2] * alpha * tt;
                            tt = tt * (1.0f - alpha);
                        }
                    } else
                    for (int j = 0; j < chunk && tt > MIN_ALPHA; ++j) {
                        const orc_record *s = &records[values[bx + (uint32_t)sort_offset + (uint32_t)j]];
                        float ox = s->image_pos[0] - px, oy = s->image_pos[1] - py;
                        /* power = -0.5*(cx*ox*ox + cz*oy*oy) - cy*ox*oy  (:84), with the two GLSL-legal
                         * contractions fixed by the gsr spec: q = fma(cz*oy, oy, cx*ox*ox);
                         * power = fma(-(cy*ox), oy, -0.5*q). */
                        float q = fmaf(s->conic[2] * oy, oy, s->conic[0] * ox * ox);
                        float power = fmaf(-(s->conic[1] * ox), oy, -0.5f * q);
                        float alpha = s->color[3] * orc_exp(power);
                        /* blended += color.rgb*alpha*t  => fma((c*alpha), t, blended) */
                        r = fmaf(s->color[0] * alpha, tt, r);
                        g = fmaf(s->color[1] * alpha, tt, g);
                        b = fmaf(s->color[2] * alpha, tt, b);
                        tt = tt * (1.0f - alpha);
                    }
                    t[l] = tt; col[l][0] = r; col[l][1] = g; col[l][2] = b;
                    shared_t += (uint32_t)(tt * 255.0f); /* :97 atomicAdd(shared_t, uint(t*MIN_FACTOR)) */
                }
            }
            if (tile_staged) tile_staged[tile_id] = tile_consumed;
            /* :100-101 */
            const float hx = (float)num_splats * 5e-4f;
            const float h0 = 0.0f * (1.0f - hx) + 1.0f * hx, h1 = 0.0f * (1.0f - hx) + 0.2f * hx, h2 = 1.0f * (1.0f - hx) + 0.2f * hx;
            for (int l = 0; l < ORC_WG; ++l) {
                const int px = tx * ORC_TILE + (l & 15), py = ty * ORC_TILE + (l >> 4);
                if (px >= W || py >= H) continue;
                float *o = out + ((size_t)py * W + px) * 4;
                float k = 1.0f - t[l];
                o[0] = col[l][0] + h0 * k * heatmap_factor;
                o[1] = col[l][1] + h1 * k * heatmap_factor;
                o[2] = col[l][2] + h2 * k * heatmap_factor;
                o[3] = 1.0f;
            }
            /* :105-110 pick: elected lane of each 32-wide subgroup = local index 32*s */
            if (pick && tile_id == target_tile_id) {
                int any = 0;
                for (int sgi = 0; sgi < 8; ++sgi) if (t[32 * sgi] != 1.0f) any = 1;
                if (any) {
                    const orc_record *s = &records[values[bx + (by - bx) / 10u]];
#pragma omp critical
                    { pick[0] = s->pos_xy[0]; pick[1] = s->pos_xy[1]; pick[2] = s->pos_z; pick[3] = (float)num_splats; }
                }
            }
        }
    if (staged_out) *staged_out = staged_total;
}

/* ------------------------------------------------------------------------------------------ */
/* a0: rasterizer.gd:122-160 -- one frame                                                      */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int64_t visible, duplicates, staged, last_tile;
    double ms_projection, ms_sort, ms_boundaries, ms_render;
} orc_frame_stats;

static double now_ms(void) {
#ifdef _OPENMP
    return omp_get_wtime() * 1e3;
#else
    return 0.0;
#endif
}

/* Scratch (records/keys/values/bounds) is caller-provided so tests can inspect every stage.
 * keys/values have capacity cap.  Returns 0, or 1 if M > cap (Q12 overflow; frame not rendered). */
int orc_frame(const float *splat60, int64_t n, const float *vp, const orc_uniforms *u, float heatmap_factor, int quirks,
              int band_y0, int band_y1, orc_record *records, uint32_t *keys, uint32_t *values, int64_t cap,
              uint32_t *bounds, float *out, orc_frame_stats *st) {
    const int gx = (u->dims[0] + ORC_TILE - 1) / ORC_TILE, gy = (u->dims[1] + ORC_TILE - 1) / ORC_TILE;
    if (band_y1 > gy) band_y1 = gy;
    const int sharded = !(band_y0 == 0 && band_y1 == gy);
    double t0 = now_ms();
    int64_t vis = 0, last = -1;
    int64_t m = orc_project(splat60, n, vp, u, band_y0, band_y1, records, keys, values, cap, &vis, &last);
    double t1 = now_ms();
    if (st) { st->visible = vis; st->duplicates = m; st->last_tile = last; st->ms_projection = t1 - t0; }
    if (m > cap) return 1;
    orc_sort_pairs(keys, values, m);
    double t2 = now_ms();
    orc_boundaries(keys, m, (int64_t)gx * gy, bounds, quirks, sharded ? last : -1);
    double t3 = now_ms();
    int64_t staged = 0;
    orc_render(records, values, bounds, u->dims[0], u->dims[1], heatmap_factor, 0xFFFFFFFFu, band_y0, band_y1, out, NULL, &staged, NULL);
    double t4 = now_ms();
    if (st) { st->staged = staged; st->ms_sort = t2 - t1; st->ms_boundaries = t3 - t2; st->ms_render = t4 - t3; }
    return 0;
}

/* exported scalar taps so the tests can pin the deterministic math against libm */
/* ---- presentation (scope row f3): resources/shaders/spatial/main.gdshader:7-11 srgb_to_linear() and the output packings the
 *      library offers (include/gsr.h GSR_OUT_*).  format: 0 RGBA32F, 1 RGB32F, 2 RGBA16F, 3 RGBA8; | 0x100 = sRGB -> linear on rgb. ---- */
static inline float orc_srgb_to_linear(float x) {
    const float higher = orc_pow((x + 0.055f) / 1.055f, 2.4f); /* pow(): the gsr deterministic pow, like every pow of the path */
    const float lower = x / 12.92f;
    return (x < 0.04045f) ? lower : higher;                    /* mix(higher, lower, lessThan(x, 0.04045)) */
}
static inline uint16_t orc_f32_to_f16(float f) { /* IEEE binary32 -> binary16, round to nearest even */
    const uint32_t u = f2u(f), sign = (u >> 16) & 0x8000u;
    const uint32_t a = u & 0x7FFFFFFFu;
    if (a >= 0x7F800000u) return (uint16_t)(sign | 0x7C00u | ((a > 0x7F800000u) ? (0x0200u | ((a >> 13) & 0x3FFu)) : 0u));
    if (a >= 0x477FF000u) return (uint16_t)(sign | 0x7C00u);                /* >= 65520 rounds to inf */
    if (a < 0x33000001u) return (uint16_t)sign;                             /* <= 2^-25 rounds to zero */
    int e = (int)(a >> 23) - 127;
    uint32_t m = (a & 0x7FFFFFu) | 0x800000u;
    int shift = (e < -14) ? (13 + (-14 - e)) : 13;                          /* subnormal halves lose extra bits */
    uint32_t half_m = m >> shift, rem = m & ((1u << shift) - 1u), halfway = 1u << (shift - 1);
    if (rem > halfway || (rem == halfway && (half_m & 1u))) half_m += 1u;
    uint32_t h = (e < -14) ? half_m : (((uint32_t)(e + 15) << 10) + (half_m - 0x400u));
    return (uint16_t)(sign | h);
}
static inline uint32_t orc_unorm8(float x) {
    const float c = orc_clamp(x, 0.0f, 1.0f);
    if (!(c == c)) return 0u;
    return (uint32_t)nearbyintf(c * 255.0f);
}
void orc_present(const float *rgba, int64_t pixels, int format, void *out) {
    const int lin = (format & 0x100) != 0, fmt = format & 0xFF;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < pixels; ++i) {
        float v[4] = {rgba[4 * i], rgba[4 * i + 1], rgba[4 * i + 2], rgba[4 * i + 3]};
        if (lin) for (int k = 0; k < 3; ++k) v[k] = orc_srgb_to_linear(v[k]);
        if (fmt == 0) memcpy((float *)out + 4 * i, v, 16);
        else if (fmt == 1) memcpy((float *)out + 3 * i, v, 12);
        else if (fmt == 2) for (int k = 0; k < 4; ++k) ((uint16_t *)out)[4 * i + k] = orc_f32_to_f16(v[k]);
        else ((uint32_t *)out)[i] = orc_unorm8(v[0]) | (orc_unorm8(v[1]) << 8) | (orc_unorm8(v[2]) << 16) | (orc_unorm8(v[3]) << 24);
    }
}

float orc_test_exp(float x) { return orc_exp(x); }
float orc_test_pow(float x, float y) { return orc_pow(x, y); }
float orc_test_log2(float x) { return orc_log2(x); }
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
