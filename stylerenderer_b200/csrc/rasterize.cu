// Z-buffer mesh rasteriser (forward, reference-shaped dcoeff, fused backward) for sm_100a.
//
// Replaces rasterize_gpu / rasterize_gpu_backward and their kernels (reference op/rasterize.cu:40-138,
// math op/rasterize.h:9-228) and the sparse-matrix scatter of op/rasterize.py:39-80.
//
// Determinism and bit-exactness: the reference CUDA kernel races (z test and payload write are not
// one atomic) and its CPU loop is the only well-defined behaviour: strict `zbuf < z`, so among equal
// depths the FIRST triangle in list order wins.  Here every (pixel, triangle) candidate is folded into
// one 64-bit key  [ order-preserving(z) : 32 | ~triangle_id : 32 ]  with a single atomicMax, which is
// order independent and reproduces exactly that rule; a resolve pass then recomputes the winner's
// coefficients with the same device function.  All geometry arithmetic uses round-to-nearest
// intrinsics (__fmul_rn / __fadd_rn / __fdiv_rn ...) which the compiler never contracts into FMAs,
// mirroring the FMA-free host build of the reference, so ids AND coefficients are bit-exact with
// `rasterize_cpu` (oracle/raster_body.inc).  float64 uses a three-pass variant (max z, min id, resolve).
//
// Work decomposition (BFM-size meshes have < 1 covered pixel per triangle): one lane per triangle for
// set-up and small boxes; triangles whose clamped box exceeds 32 pixels are re-distributed over the
// warp (ballot + shuffle broadcast, 32 pixels per step) so a few big triangles cannot serialise a lane.
#include "common.cuh"
#include <float.h>

namespace sr {
namespace {

constexpr int kThreads = 256;

// ---- strictly rounded arithmetic ----------------------------------------------------------------
template <typename T> struct RN;
template <> struct RN<float> {
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ float lowest() { return -FLT_MAX; }
};
template <> struct RN<double> {
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ double lowest() { return -DBL_MAX; }
};

__device__ __forceinline__ uint32_t order_bits(float z) {
    uint32_t u = __float_as_uint(z);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ uint64_t order_bits(double z) {
    uint64_t u = (uint64_t)__double_as_longlong(z);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

template <typename T>
struct Tri {
    T p[9];        // projected corners: x, y in pixel units, z untouched
    T E[9];        // edge-function matrix, rows: constant, d/dx, d/dy
    T det;
    int x_lo, x_hi, y_lo, y_hi;
};

// ceil / floor + clamp to the raster.  The float overloads stay in fp32 (ceilf of a float is exact and equals
// ceil of its double value), which keeps the slow fp64 pipe out of the per-triangle path.
__device__ __forceinline__ int clamp_lo(double v) { return v <= 0.0 ? 0 : (v > 1.0e9 ? 1000000000 : (int)ceil(v)); }
__device__ __forceinline__ int clamp_hi(double v, int span) {
    return v >= (double)(span - 1) ? span - 1 : (v < -1.0 ? -1 : (int)floor(v));
}
__device__ __forceinline__ int clamp_lo(float v) { return v <= 0.0f ? 0 : (v > 1.0e9f ? 1000000000 : (int)ceilf(v)); }
__device__ __forceinline__ int clamp_hi(float v, int span) {
    return v >= (float)(span - 1) ? span - 1 : (v < -1.0f ? -1 : (int)floorf(v));
}

// NDC -> pixel coordinates of ONE vertex (reference op/rasterize.h:15-22), in place: q = (x, y, z) -> (X, Y, z).
// The perspective rejection (z >= -eps) is NOT decided here: it needs the untouched z, which stays in q[2].
template <typename T>
__device__ __forceinline__ void project_vertex(T *q, int span_x, int span_y, bool perspective, T eps)
{
    using R = RN<T>;
    if (perspective && !(q[2] >= -eps)) {
        q[0] = R::div(q[0], -q[2]);
        q[1] = R::div(q[1], -q[2]);
    }
    // ((1 + x) * W / 2) - .5 : the reference subtracts a double .5 and rounds back, which for
    // +,-,*,/ equals the single-precision operation (innocuous double rounding)
    // `/ 2` as `* 0.5`: bit-identical for every finite or infinite input (a power-of-two scaling, checked on the CPU
    // over all 2^32 float patterns) and one FMUL instead of an IEEE division (FCHK + slow path)
    q[0] = R::sub(R::mul(R::mul(R::add((T)1, q[0]), (T)span_x), (T)0.5), (T)0.5);
    q[1] = R::sub(R::mul(R::mul(R::sub((T)1, q[1]), (T)span_y), (T)0.5), (T)0.5);
}

// reference op/rasterize.h:9-75 (`barycentric` with det_ != NULL); mirrors oracle sr_tri_setup.
// PROJECTED: t.p already holds pixel coordinates (project_vertex, e.g. from the per-vertex pre-pass).
// CLIP: compute the clamped pixel box and reject empty ones (the resolve pass knows its pixel and skips it).
template <typename T, bool PROJECTED = false, bool CLIP = true>
__device__ __forceinline__ bool tri_setup(Tri<T> &t, int span_x, int span_y, bool perspective, T eps)
{
    using R = RN<T>;
    T xmin = 0, xmax = 0, ymin = 0, ymax = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        T *q = t.p + 3 * c;
        if (perspective && q[2] >= -eps) return false;
        if (!PROJECTED) project_vertex<T>(q, span_x, span_y, perspective, eps);
        if (CLIP) {
            if (c == 0) { xmin = xmax = q[0]; ymin = ymax = q[1]; }
            else {
                if (xmin > q[0]) xmin = q[0]; else if (xmax < q[0]) xmax = q[0];
                if (ymin > q[1]) ymin = q[1]; else if (ymax < q[1]) ymax = q[1];
            }
        }
    }
    if (CLIP) {
        t.x_lo = clamp_lo(xmin); t.x_hi = clamp_hi(xmax, span_x);
        t.y_lo = clamp_lo(ymin); t.y_hi = clamp_hi(ymax, span_y);
        if (t.x_hi < t.x_lo || t.y_hi < t.y_lo) return false;
    }
    const T *p = t.p;
    T *E = t.E;
    E[0] = R::sub(R::mul(p[3], p[7]), R::mul(p[4], p[6]));
    E[1] = R::sub(R::mul(p[1], p[6]), R::mul(p[0], p[7]));
    E[2] = R::sub(R::mul(p[0], p[4]), R::mul(p[1], p[3]));
    T det = R::add(R::add(E[0], E[1]), E[2]);
    if (det > eps) return false;
    E[3] = R::sub(p[4], p[7]); E[4] = R::sub(p[7], p[1]); E[5] = R::sub(p[1], p[4]);
    E[6] = R::sub(p[6], p[3]); E[7] = R::sub(p[0], p[6]); E[8] = R::sub(p[3], p[0]);
    if (det < 0) {
#pragma unroll
        for (int c = 0; c < 9; ++c) E[c] = -E[c];
        det = -det;
    }
    t.det = det;
    return true;
}

// longest edge I of a zero-area triangle: project onto it (segment) or test the point (rasterize.h:105-120)
template <typename T, int I>
__device__ __forceinline__ bool degenerate_case(const Tri<T> &t, T px, T py, T len_i, T eps, T w[3])
{
    using R = RN<T>;
    constexpr int J = (I + 1) % 3, K = (J + 1) % 3;
    const T *E = t.E, *p = t.p;
    if (len_i > eps) {
        const T lj = R::add(R::mul(-R::sub(px, p[3 * K]), E[6 + I]), R::mul(R::sub(py, p[3 * K + 1]), E[3 + I]));
        const T lk = R::sub(R::mul(R::sub(px, p[3 * J]), E[6 + I]), R::mul(R::sub(py, p[3 * J + 1]), E[3 + I]));
        const T li = R::add(lj, lk);
        w[I] = 0; w[J] = R::div(lj, li); w[K] = R::div(lk, li);
        return w[J] >= -eps && w[K] >= -eps;
    }
    w[J] = 0; w[K] = 0; w[I] = 1;
    const T dx = R::sub(px, p[3 * I]), dy = R::sub(py, p[3 * I + 1]);
    const T d2 = R::add(R::mul(dx, dx), R::mul(dy, dy));
    return d2 < eps;
}

// reference op/rasterize.h:76-142 (`normalize_coeff` + depth of `assign_buffer`); mirrors sr_tri_sample.
template <typename T>
__device__ __forceinline__ bool tri_sample(const Tri<T> &t, T px, T py, bool perspective, T eps, T w[3], T &z)
{
    using R = RN<T>;
    const T *E = t.E, *p = t.p;
    w[0] = R::add(R::add(E[0], R::mul(E[3], px)), R::mul(E[6], py));
    w[1] = R::add(R::add(E[1], R::mul(E[4], px)), R::mul(E[7], py));
    w[2] = R::add(R::add(E[2], R::mul(E[5], px)), R::mul(E[8], py));
    if (w[0] < -eps || w[1] < -eps || w[2] < -eps) return false;
    if (t.det > eps) {
        T s = R::add(R::add(w[0], w[1]), w[2]);
        w[0] = R::div(w[0], s); w[1] = R::div(w[1], s); w[2] = R::div(w[2], s);
    } else {                                           // degenerate triangle (rasterize.h:87-120)
        const T l0 = R::add(R::mul(E[3], E[3]), R::mul(E[6], E[6]));
        const T l1 = R::add(R::mul(E[4], E[4]), R::mul(E[7], E[7]));
        const T l2 = R::add(R::mul(E[5], E[5]), R::mul(E[8], E[8]));
        int i = (l0 > l1) ? 0 : 1;
        i = (((i == 0) ? l0 : l1) > l2) ? i : 2;
        // compile-time indices only (dynamic indexing would push the whole Tri into local memory)
        bool ok;
        if (i == 0) ok = degenerate_case<T, 0>(t, px, py, l0, eps, w);
        else if (i == 1) ok = degenerate_case<T, 1>(t, px, py, l1, eps, w);
        else ok = degenerate_case<T, 2>(t, px, py, l2, eps, w);
        if (!ok) return false;
    }
    if (perspective) {
        w[0] = R::div(w[0], p[2]); w[1] = R::div(w[1], p[5]); w[2] = R::div(w[2], p[8]);
        T s = R::add(R::add(w[0], w[1]), w[2]);
        if (s >= -eps) return false;
        w[0] = R::mul(w[0], s); w[1] = R::mul(w[1], s); w[2] = R::mul(w[2], s);
        z = s;
    } else {
        z = R::add(R::add(R::mul(w[0], p[2]), R::mul(w[1], p[5])), R::mul(w[2], p[8]));
    }
    return true;
}

struct RasterGeom {
    int64_t b, nv, nf;
    int h, w;
    int shared_v, shared_f, perspective;
    // exact division by h*w and by w through host-prepared magic numbers: the pixel index arithmetic of the resolve /
    // backward passes was ~85 % of their instructions as 64-bit divisions and address chains (ncu, round 2); every pixel
    // count is < 2^31 (checked by the entry points)
    FastDiv div_hw, div_w;
};

template <typename T>
__device__ __forceinline__ bool load_tri(Tri<T> &t, const T *__restrict__ V, const int64_t *__restrict__ F,
                                         int64_t f, int64_t nv, int64_t ids[3])
{
    ids[0] = __ldg(F + 3 * f); ids[1] = __ldg(F + 3 * f + 1); ids[2] = __ldg(F + 3 * f + 2);
    if (ids[0] < 0 || ids[1] < 0 || ids[2] < 0 || ids[0] >= nv || ids[1] >= nv || ids[2] >= nv) return false;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const T *s = V + 3 * ids[k];
        t.p[3 * k] = __ldg(s); t.p[3 * k + 1] = __ldg(s + 1); t.p[3 * k + 2] = __ldg(s + 2);
    }
    return true;
}

// ---- per-vertex / per-triangle pre-pass (float path) -------------------------------------------------------------
// A BFM-size mesh has 6 triangle corners per vertex and the resolve pass sets the winning triangle up once more per
// pixel: the NDC -> pixel transform of a vertex used to run ~9 times per image.  The pre-pass runs it ONCE per (image,
// vertex, raster size) with the same single-rounded operations (bit-identical coordinates) and narrows the triangle list
// to int32 (-1 in the first slot = a corner outside [0, nv): skipped like reference op/rasterize.cpp:30-33), so the
// triangle and resolve passes gather 12-byte ids and ready-made pixel coordinates.
//   P [level][image or 1][nv][3] = (X, Y, z)      F32 [image or 1][nf][3]
struct PreLevels { int n; int size[SR_RASTER_MAX_LEVELS]; };

template <typename T>
__global__ void __launch_bounds__(kThreads)
raster_prepass_kernel(const RasterGeom g, const PreLevels L, const T *__restrict__ verts, const int64_t *__restrict__ tris,
                      T *__restrict__ P, int32_t *__restrict__ F32, T eps)
{
    const int64_t nvert = (g.shared_v ? 1 : g.b) * g.nv, ntri = (g.shared_f ? 1 : g.b) * g.nf;
    const int64_t stride = (int64_t)gridDim.x * kThreads, first = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    for (int64_t i = first; i < nvert; i += stride) {
        const T x = __ldg(verts + 3 * i), y = __ldg(verts + 3 * i + 1), z = __ldg(verts + 3 * i + 2);
        for (int l = 0; l < L.n; ++l) {
            T q[3] = {x, y, z};
            // reference passes (h, w) for (w, h): x spans h, y spans w (op/rasterize.cpp:38); square targets only
            project_vertex<T>(q, L.size[l], L.size[l], g.perspective != 0, eps);
            T *o = P + ((int64_t)l * nvert + i) * 3;
            o[0] = q[0]; o[1] = q[1]; o[2] = q[2];
        }
    }
    for (int64_t i = first; i < ntri; i += stride) {
        const int64_t a = __ldg(tris + 3 * i), b = __ldg(tris + 3 * i + 1), c = __ldg(tris + 3 * i + 2);
        const bool ok = a >= 0 && b >= 0 && c >= 0 && a < g.nv && b < g.nv && c < g.nv;
        F32[3 * i] = ok ? (int32_t)a : -1; F32[3 * i + 1] = ok ? (int32_t)b : 0; F32[3 * i + 2] = ok ? (int32_t)c : 0;
    }
}

template <typename T>
__device__ __forceinline__ bool load_tri_pre(Tri<T> &t, const T *__restrict__ P, const int32_t *__restrict__ F, int64_t f,
                                             int32_t ids[3])
{
    ids[0] = __ldg(F + 3 * f); ids[1] = __ldg(F + 3 * f + 1); ids[2] = __ldg(F + 3 * f + 2);
    if (ids[0] < 0) return false;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const T *s = P + 3 * (int64_t)ids[k];
        t.p[3 * k] = __ldg(s); t.p[3 * k + 1] = __ldg(s + 1); t.p[3 * k + 2] = __ldg(s + 2);
    }
    return true;
}

// PASS 0: float, packed (z, ~id) key.  PASS 1: double, max ordered z.  PASS 2: double, min id among z == zmax.
template <typename T, int PASS>
__device__ __forceinline__ void emit(const Tri<T> &t, int x, int y, uint32_t f, const RasterGeom &g, T eps,
                                     uint64_t *__restrict__ zkeys, uint32_t *__restrict__ idkeys, int64_t img)
{
    // zkeys / idkeys already point at this image's planes (cover_warp's callers add img * h * w once)
    (void)img;
    T w[3], z;
    if (!tri_sample<T>(t, (T)x, (T)y, g.perspective != 0, eps, w, z)) return;
    if (!(z > RN<T>::lowest())) return;            // the buffer starts at -MAX and the test is strict; drops NaN
    z = z + (T)0;                                  // -0 -> +0: they compare equal on the host
    // reference pixel index is x + y*w (op/rasterize.cpp:42) with x spanning h and y spanning w (quirk 6)
    const uint32_t pix = (uint32_t)y * (uint32_t)g.w + (uint32_t)x;
    if (PASS == 0) {
        const uint64_t key = ((uint64_t)order_bits((float)z) << 32) | (uint64_t)(0xffffffffu - f);
        atomicMax(reinterpret_cast<unsigned long long *>(zkeys + pix), (unsigned long long)key);
    } else if (PASS == 1) {
        atomicMax(reinterpret_cast<unsigned long long *>(zkeys + pix), (unsigned long long)order_bits((double)z));
    } else {
        if (zkeys[pix] == order_bits((double)z)) atomicMin(idkeys + pix, f);
    }
}

// Coverage of one set-up triangle per lane (t valid where `live`): small boxes are walked by their own lane, boxes of
// more than 32 pixels are re-distributed over the warp.  Must be called by all 32 lanes (ballot / shuffles).
template <typename T, int PASS>
__device__ __forceinline__ void cover_warp(const Tri<T> &t, bool live, uint32_t f, uint32_t lane, const RasterGeom &g, T eps,
                                           uint64_t *__restrict__ zkeys, uint32_t *__restrict__ idkeys, int64_t img)
{
    int bw = 0, area = 0;
    if (live) { bw = t.x_hi - t.x_lo + 1; area = bw * (t.y_hi - t.y_lo + 1); }
    const bool big = live && area > 32;
    if (live && !big) {
        for (int y = t.y_lo; y <= t.y_hi; ++y)
            for (int x = t.x_lo; x <= t.x_hi; ++x) emit<T, PASS>(t, x, y, f, g, eps, zkeys, idkeys, img);
    }
    // big boxes: the whole warp walks the box of one lane at a time, 32 pixels per step
    uint32_t pending = __ballot_sync(0xffffffffu, big);
    while (pending) {
        const int src = __ffs(pending) - 1;
        pending &= pending - 1;
        Tri<T> s;
#pragma unroll
        for (int c = 0; c < 9; ++c) { s.p[c] = __shfl_sync(0xffffffffu, t.p[c], src); s.E[c] = __shfl_sync(0xffffffffu, t.E[c], src); }
        s.det = __shfl_sync(0xffffffffu, t.det, src);
        s.x_lo = __shfl_sync(0xffffffffu, t.x_lo, src); s.y_lo = __shfl_sync(0xffffffffu, t.y_lo, src);
        const int sbw = __shfl_sync(0xffffffffu, bw, src), sarea = __shfl_sync(0xffffffffu, area, src);
        const uint32_t sf = __shfl_sync(0xffffffffu, f, src);
        for (int k = lane; k < sarea; k += 32) {
            const int yy = k / sbw, xx = k - yy * sbw;
            emit<T, PASS>(s, s.x_lo + xx, s.y_lo + yy, sf, g, eps, zkeys, idkeys, img);
        }
    }
}

template <typename T, int PASS>
__global__ void __launch_bounds__(kThreads)
raster_tri_kernel(const RasterGeom g, const T *__restrict__ verts, const int64_t *__restrict__ tris,
                  uint64_t *__restrict__ zkeys, uint32_t *__restrict__ idkeys, T eps,
                  const T *__restrict__ P = nullptr, const int32_t *__restrict__ F32 = nullptr)
{
    const int64_t img = blockIdx.y;                      // one grid row per image
    const int64_t total = g.nf;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    const uint32_t lane = threadIdx.x & 31u;
    const T *V = verts + (g.shared_v ? 0 : img * g.nv * 3);
    const int64_t *F = tris + (g.shared_f ? 0 : img * g.nf * 3);
    const T *Pi = P ? P + (g.shared_v ? 0 : img * g.nv * 3) : nullptr;         // pre-pass products of this image
    const int32_t *Fi = F32 ? F32 + (g.shared_f ? 0 : img * g.nf * 3) : nullptr;
    // warp-uniform trip count: every lane of a warp runs the same number of iterations
    for (int64_t base = (int64_t)blockIdx.x * kThreads + (threadIdx.x & ~31u); base < total; base += stride) {
        const int64_t item = base + lane;
        Tri<T> t;
        bool live = item < total;
        const uint32_t f = (uint32_t)item;
        if (live) {
            if (Pi) {
                int32_t ids[3];
                live = load_tri_pre<T>(t, Pi, Fi, f, ids) && tri_setup<T, true>(t, g.h, g.w, g.perspective != 0, eps);
            } else {
                int64_t ids[3];
                live = load_tri<T>(t, V, F, f, g.nv, ids);
                // reference passes (h, w) for (w, h): x spans h, y spans w (op/rasterize.cpp:38)
                if (live) live = tri_setup<T>(t, g.h, g.w, g.perspective != 0, eps);
            }
        }
        cover_warp<T, PASS>(t, live, f, lane, g, eps, zkeys + img * g.h * (int64_t)g.w,
                            idkeys ? idkeys + img * g.h * (int64_t)g.w : nullptr, img);
    }
}

// ---- resolution pyramid: the same mesh rasterised at several square sizes in one launch per pass -------------------
// GeneratorWithMap renders the normal map at 4, 8, ..., 256 pixels every forward (reference model.py:260-270: seven
// independent rasterize calls).  One triangle pass loads each triangle (3 int64 ids + 9 gathered floats) ONCE and sets
// it up per level -- the pixel transform depends on the size, so the per-level arithmetic stays exactly the reference's
// and the result is bit-identical to seven single-size calls; at the coarse levels nearly every triangle leaves after
// the empty-box test.  One resolve launch and one backward launch walk the concatenated 256-pixel blocks of all levels.
constexpr int kMaxLevels = SR_RASTER_MAX_LEVELS;

template <typename T>
struct Pyramid {
    int n;
    int size[kMaxLevels];
    int64_t key_off[kMaxLevels];          // first key of the level inside the workspace
    int64_t blk_off[kMaxLevels + 1];      // prefix sum of 256-pixel blocks
    int64_t *ids[kMaxLevels];
    T *bary[kMaxLevels];
    T *out[kMaxLevels];
    const T *gout[kMaxLevels];
    int planar;                           // forward: interpolated maps are written as [b, c, size, size] planes (NCHW)
    FastDiv div_hw[kMaxLevels], div_w[kMaxLevels];
};

template <typename T>
__global__ void __launch_bounds__(kThreads)
raster_tri_pyramid_kernel(const RasterGeom g0, const Pyramid<T> L, const T *__restrict__ verts,
                          const int64_t *__restrict__ tris, uint64_t *__restrict__ zkeys, T eps,
                          const T *__restrict__ P, const int32_t *__restrict__ F32)
{
    const int64_t img = blockIdx.y;
    const int64_t total = g0.nf;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    const uint32_t lane = threadIdx.x & 31u;
    const int64_t level_stride = (g0.shared_v ? 1 : g0.b) * g0.nv * 3;         // P: [level][image or 1][nv][3]
    const T *Pi = P + (g0.shared_v ? 0 : img * g0.nv * 3);
    const int32_t *Fi = F32 + (g0.shared_f ? 0 : img * g0.nf * 3);
    for (int64_t base = (int64_t)blockIdx.x * kThreads + (threadIdx.x & ~31u); base < total; base += stride) {
        const int64_t item = base + lane;
        bool loaded = item < total;
        const uint32_t f = (uint32_t)item;
        int32_t ids[3] = {-1, 0, 0};
        if (loaded) {
            ids[0] = __ldg(Fi + 3 * item); ids[1] = __ldg(Fi + 3 * item + 1); ids[2] = __ldg(Fi + 3 * item + 2);
            loaded = ids[0] >= 0;
        }
        if (!__any_sync(0xffffffffu, loaded)) continue;
        for (int l = 0; l < L.n; ++l) {                      // warp-uniform
            RasterGeom g = g0;
            g.h = g.w = L.size[l];
            Tri<T> t;
            bool live = loaded;
            if (live) {
                const T *Pl = Pi + l * level_stride;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const T *s = Pl + 3 * (int64_t)ids[k];
                    t.p[3 * k] = __ldg(s); t.p[3 * k + 1] = __ldg(s + 1); t.p[3 * k + 2] = __ldg(s + 2);
                }
                live = tri_setup<T, true>(t, g.h, g.w, g.perspective != 0, eps);
            }
            cover_warp<T, 0>(t, live, f, lane, g, eps, zkeys + L.key_off[l] + img * g.h * (int64_t)g.w, nullptr, img);
        }
    }
}

// One thread per pixel: decode the winner, recompute its coefficients, write ids / bary (/ interpolated tex).
// The three outputs are 24 + 12 (+ 4c) bytes per pixel: written per thread they would be strided 8/4-byte stores, so a
// CTA stages its 256 pixels in shared memory and streams them out as contiguous 16-byte stores.
constexpr int kMaxStageC = 4;

template <typename V4>
__device__ __forceinline__ void stream_out(void *gdst, const void *ssrc, int bytes, bool vec_ok)
{
    if (vec_ok) {
        const int n16 = bytes >> 4;
        for (int i = threadIdx.x; i < n16; i += kThreads)
            reinterpret_cast<int4 *>(gdst)[i] = reinterpret_cast<const int4 *>(ssrc)[i];
        const int rem = bytes - (n16 << 4);                  // 0 or a multiple of 4
        if ((int)threadIdx.x < (rem >> 2))
            reinterpret_cast<int *>(gdst)[(n16 << 2) + threadIdx.x] = reinterpret_cast<const int *>(ssrc)[(n16 << 2) + threadIdx.x];
    } else {
        for (int i = threadIdx.x; i < (bytes >> 2); i += kThreads)
            reinterpret_cast<int *>(gdst)[i] = reinterpret_cast<const int *>(ssrc)[i];
    }
}

// full 256-pixel block, 16-byte aligned destination: N16 int4 per CTA as straight-line code (no loop, no remainder)
template <int N16>
__device__ __forceinline__ void stream_out_full(void *gdst, const void *ssrc)
{
#pragma unroll
    for (int r = 0; r < (N16 + kThreads - 1) / kThreads; ++r) {
        const int i = r * kThreads + (int)threadIdx.x;
        if (N16 % kThreads == 0 || i < N16) reinterpret_cast<int4 *>(gdst)[i] = reinterpret_cast<const int4 *>(ssrc)[i];
    }
}

template <typename T>
struct ResolveStage {
    int64_t ids[kThreads * 3];
    T w[kThreads * 3];
    T out[kThreads * kMaxStageC];
};

// 256 consecutive pixels (block `blk`) of one raster geometry; called by the whole CTA.  All pixel indices are 32-bit
// (npix < 2^31) and the (image, row, column) split uses the host-prepared exact divisions of RasterGeom.
template <typename T, bool PACKED>
__device__ __forceinline__ void resolve_block(const RasterGeom &g, const T *__restrict__ verts, const int64_t *__restrict__ tris,
                                              const uint64_t *__restrict__ zkeys, const uint32_t *__restrict__ idkeys, T eps,
                                              int64_t *__restrict__ ids_out, T *__restrict__ bary_out,
                                              const T *__restrict__ tex, int c, T *__restrict__ out, int vec_ok,
                                              uint32_t blk, uint32_t npix, ResolveStage<T> &st, bool planar = false,
                                              const T *__restrict__ P = nullptr, const int32_t *__restrict__ F32 = nullptr)
{
    // planar: the map leaves as [b, c, h, w] planes -- consecutive threads are consecutive pixels of one plane, so the
    // per-thread 4-byte stores coalesce by themselves; ids_out / bary_out may then be NULL (forward-only callers: the mesh
    // is sampled under no_grad in the training loop, reference train.py:249-251, so nothing reads those buffers back)
    const bool stage_out = tex != nullptr && c <= kMaxStageC && !planar;
    const uint32_t pix0 = blk * kThreads, pix = pix0 + threadIdx.x;
    const int count = (int)((npix - pix0 < (uint32_t)kThreads) ? (npix - pix0) : (uint32_t)kThreads);
    int64_t ids[3] = {0, 0, 0};
    T w[3] = {0, 0, 0};
    bool hit = false;
    uint32_t img = 0, rem = 0;
    if (pix < npix) {
        const uint64_t key = zkeys[pix];
        hit = key != 0;
        g.div_hw.divmod(pix, img, rem);
        if (hit) {
            const uint32_t f = PACKED ? (0xffffffffu - (uint32_t)key) : idkeys[pix];
            uint32_t y, x;
            g.div_w.divmod(rem, y, x);
            const int64_t voff = g.shared_v ? 0 : (int64_t)img * g.nv * 3, foff = g.shared_f ? 0 : (int64_t)img * g.nf * 3;
            Tri<T> t;
            T z;
            if (P) {             // pre-projected corners, no pixel box needed: the edge matrix and the sample only
                int32_t i32[3];
                hit = load_tri_pre<T>(t, P + voff, F32 + foff, f, i32) &&
                      tri_setup<T, true, false>(t, g.h, g.w, g.perspective != 0, eps) &&
                      tri_sample<T>(t, (T)(int)x, (T)(int)y, g.perspective != 0, eps, w, z);
                ids[0] = i32[0]; ids[1] = i32[1]; ids[2] = i32[2];
            } else {
                hit = load_tri<T>(t, verts + voff, tris + foff, f, g.nv, ids) && tri_setup<T>(t, g.h, g.w, g.perspective != 0, eps) &&
                      tri_sample<T>(t, (T)(int)x, (T)(int)y, g.perspective != 0, eps, w, z);
            }
            if (hit && !g.shared_v) { const int64_t o = g.nv * (int64_t)img; ids[0] += o; ids[1] += o; ids[2] += o; }
            if (!hit) { ids[0] = ids[1] = ids[2] = 0; w[0] = w[1] = w[2] = 0; }
        }
    }
    if (ids_out) {
#pragma unroll
        for (int k = 0; k < 3; ++k) st.ids[3 * threadIdx.x + k] = ids[k];
    }
    if (bary_out) {
#pragma unroll
        for (int k = 0; k < 3; ++k) st.w[3 * threadIdx.x + k] = w[k];
    }
    if (tex && pix < npix) {
        // op/rasterize.py:29-37: sum_k tex[ids_k] * bary_k  (background: row 0 times 0 = 0)
        if (c == 3) {                                    // normals / colours: the case every caller of the model has
            T v[3] = {0, 0, 0};
            if (hit) {
                const T *t0 = tex + ids[0] * 3, *t1 = tex + ids[1] * 3, *t2 = tex + ids[2] * 3;
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) v[ch] = __ldg(t0 + ch) * w[0] + __ldg(t1 + ch) * w[1] + __ldg(t2 + ch) * w[2];
            }
            if (stage_out) { st.out[threadIdx.x * 3] = v[0]; st.out[threadIdx.x * 3 + 1] = v[1]; st.out[threadIdx.x * 3 + 2] = v[2]; }
            else if (planar) {
                T *o = out + ((size_t)img * 3 * g.h) * g.w + rem;
                const size_t plane = (size_t)g.h * g.w;
                o[0] = v[0]; o[plane] = v[1]; o[2 * plane] = v[2];
            } else { T *o = out + (size_t)pix * 3; o[0] = v[0]; o[1] = v[1]; o[2] = v[2]; }
        } else {
            for (int ch = 0; ch < c; ++ch) {
                T v = 0;
                if (hit) v = tex[ids[0] * c + ch] * w[0] + tex[ids[1] * c + ch] * w[1] + tex[ids[2] * c + ch] * w[2];
                if (stage_out) st.out[threadIdx.x * c + ch] = v;
                else if (planar) out[(((size_t)img * c + ch) * g.h) * g.w + rem] = v;
                else out[(size_t)pix * c + ch] = v;
            }
        }
    }
    if (!ids_out && !bary_out && !stage_out) return;         // maps-only planar output: nothing staged
    __syncthreads();
    if (count == kThreads && vec_ok && sizeof(T) == 4) {     // common case: straight-line 16-byte copies
        if (ids_out) stream_out_full<kThreads * 24 / 16>(ids_out + (size_t)pix0 * 3, st.ids);
        if (bary_out) stream_out_full<kThreads * 12 / 16>(bary_out + (size_t)pix0 * 3, st.w);
        if (stage_out && c == 3) stream_out_full<kThreads * 12 / 16>(out + (size_t)pix0 * 3, st.out);
        else if (stage_out) stream_out<int4>(out + (size_t)pix0 * c, st.out, count * c * (int)sizeof(T), ((size_t)pix0 * c * sizeof(T)) % 16 == 0);
    } else {
        if (ids_out) stream_out<int4>(ids_out + (size_t)pix0 * 3, st.ids, count * 3 * (int)sizeof(int64_t), vec_ok);
        if (bary_out) stream_out<int4>(bary_out + (size_t)pix0 * 3, st.w, count * 3 * (int)sizeof(T), vec_ok);
        if (stage_out) stream_out<int4>(out + (size_t)pix0 * c, st.out, count * c * (int)sizeof(T),
                                        vec_ok && (((size_t)pix0 * c * sizeof(T)) % 16 == 0));
    }
    __syncthreads();
}

template <typename T, bool PACKED>
__global__ void __launch_bounds__(kThreads)
raster_resolve_kernel(const RasterGeom g, const T *__restrict__ verts, const int64_t *__restrict__ tris,
                      const uint64_t *__restrict__ zkeys, const uint32_t *__restrict__ idkeys, T eps,
                      int64_t *__restrict__ ids_out, T *__restrict__ bary_out,
                      const T *__restrict__ tex, int c, T *__restrict__ out, int vec_ok,
                      const T *__restrict__ P = nullptr, const int32_t *__restrict__ F32 = nullptr)
{
    __shared__ __align__(16) ResolveStage<T> st;
    const uint32_t npix = (uint32_t)(g.b * g.h * (int64_t)g.w);
    const uint32_t nblk = (npix + kThreads - 1) / kThreads;
    for (uint32_t blk = blockIdx.x; blk < nblk; blk += gridDim.x)
        resolve_block<T, PACKED>(g, verts, tris, zkeys, idkeys, eps, ids_out, bary_out, tex, c, out, vec_ok, blk, npix, st,
                                 false, P, F32);
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
raster_resolve_pyramid_kernel(const RasterGeom g0, const Pyramid<T> L, const T *__restrict__ verts,
                              const int64_t *__restrict__ tris, const uint64_t *__restrict__ zkeys, T eps,
                              const T *__restrict__ tex, int c, int vec_ok, const T *__restrict__ P,
                              const int32_t *__restrict__ F32)
{
    const int64_t level_stride = (g0.shared_v ? 1 : g0.b) * g0.nv * 3;
    __shared__ __align__(16) ResolveStage<T> st;
    const int64_t nblk = L.blk_off[L.n];
    for (int64_t blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        int l = 0;
        while (blk >= L.blk_off[l + 1]) ++l;                 // CTA-uniform
        RasterGeom g = g0;
        g.h = g.w = L.size[l];
        g.div_hw = L.div_hw[l]; g.div_w = L.div_w[l];
        const uint32_t npix = (uint32_t)(g.b * g.h * (int64_t)g.w);
        resolve_block<T, true>(g, verts, tris, zkeys + L.key_off[l], nullptr, eps, L.ids[l], L.bary[l], tex, c, L.out[l],
                               vec_ok, (uint32_t)(blk - L.blk_off[l]), npix, st, L.planar != 0, P + l * level_stride, F32);
    }
}

// reference op/rasterize.h:168-228 (`barycentric_grad`); mirrors oracle sr_bary_grad.
template <typename T>
__device__ __forceinline__ bool bary_grad(const T q[9], T px, T py, T axis_x, T axis_y, bool perspective, T eps,
                                          T out[27])
{
    using R = RN<T>;
    const T u = R::div(R::add(R::sub(R::mul(px, (T)2), axis_x), (T)1), axis_x);
    const T v = R::div(R::sub(R::add(R::mul(py, (T)-2), axis_y), (T)1), axis_y);
    T E[9], det, cf[3];
    E[0] = R::sub(R::mul(q[3], q[7]), R::mul(q[4], q[6]));
    E[1] = R::sub(R::mul(q[1], q[6]), R::mul(q[0], q[7]));
    E[2] = R::sub(R::mul(q[0], q[4]), R::mul(q[1], q[3]));
    if (perspective) {
        if (q[2] >= -eps || q[5] >= -eps || q[8] >= -eps) return false;
        det = R::add(R::add(R::mul(E[0], q[2]), R::mul(E[1], q[5])), R::mul(E[2], q[8]));
        if ((det >= -eps) & (det <= eps)) return false;
        E[0] = -E[0]; E[1] = -E[1]; E[2] = -E[2];
        E[3] = R::sub(R::mul(q[4], q[8]), R::mul(q[5], q[7]));
        E[4] = R::sub(R::mul(q[2], q[7]), R::mul(q[1], q[8]));
        E[5] = R::sub(R::mul(q[1], q[5]), R::mul(q[2], q[4]));
        E[6] = R::sub(R::mul(q[5], q[6]), R::mul(q[3], q[8]));
        E[7] = R::sub(R::mul(q[0], q[8]), R::mul(q[2], q[6]));
        E[8] = R::sub(R::mul(q[2], q[3]), R::mul(q[0], q[5]));
    } else {
        det = R::add(R::add(E[0], E[1]), E[2]);
        E[3] = R::sub(q[4], q[7]); E[4] = R::sub(q[7], q[1]); E[5] = R::sub(q[1], q[4]);
        E[6] = R::sub(q[6], q[3]); E[7] = R::sub(q[0], q[6]); E[8] = R::sub(q[3], q[0]);
    }
    if (!(det < -eps || det > eps)) return false;
#pragma unroll
    for (int l = 0; l < 9; ++l) E[l] = R::div(E[l], det);
    cf[0] = R::add(R::add(E[0], R::mul(E[3], u)), R::mul(E[6], v));
    cf[1] = R::add(R::add(E[1], R::mul(E[4], u)), R::mul(E[7], v));
    cf[2] = R::add(R::add(E[2], R::mul(E[5], u)), R::mul(E[8], v));
#pragma unroll
    for (int l = 0; l < 27; ++l) {
        const int i = l / 9, j = (l + 1) % 3, k = (l / 3) % 3;
        out[l] = R::mul(-cf[k], E[i + j * 3]);
    }
    if (perspective) {
        const T tot = R::add(R::add(cf[0], cf[1]), cf[2]);
#pragma unroll
        for (int l = 0; l < 9; ++l) {
            const T dsum = R::add(R::add(out[l], out[l + 9]), out[l + 18]);
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const T corr = R::div(R::mul(cf[i], dsum), tot);
                out[l + i * 9] = (l % 3 == 2) ? R::div(R::sub(-out[l + i * 9], corr), tot)
                                              : R::div(R::sub(out[l + i * 9], corr), tot);
            }
        }
    } else {
#pragma unroll
        for (int l = 0; l < 9; ++l) out[2 + l * 3] = 0;
    }
    return true;
}

template <typename T>
__device__ __forceinline__ bool gather_pixel_tri(const T *__restrict__ verts, const int64_t *__restrict__ I,
                                                 int64_t limit, T q[9])
{
    const int64_t a = I[0], b = I[1], c = I[2];
    if (a == b || a == c || b == c) return false;                       // op/rasterize.cpp:74
    if (a < 0 || b < 0 || c < 0 || a >= limit || b >= limit || c >= limit) return false;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        q[k] = __ldg(verts + 3 * a + k); q[3 + k] = __ldg(verts + 3 * b + k); q[6 + k] = __ldg(verts + 3 * c + k);
    }
    return true;
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
raster_dcoeff_kernel(int64_t b, int64_t n, int h, int w, int perspective, const T *__restrict__ verts,
                     const int64_t *__restrict__ ids, T *__restrict__ dcoeff, T eps)
{
    const int64_t npix = b * h * (int64_t)w;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t pix = (int64_t)blockIdx.x * kThreads + threadIdx.x; pix < npix; pix += stride) {
        T q[9], d[27];
        if (!gather_pixel_tri<T>(verts, ids + 3 * pix, n * b, q)) continue;
        const int rem = (int)(pix % (h * (int64_t)w));
        // reference passes ((scalar)h, (scalar)w) for (w, h) (op/rasterize.cpp:87)
        if (!bary_grad<T>(q, (T)(rem % w), (T)(rem / w), (T)h, (T)w, perspective != 0, eps, d)) continue;
#pragma unroll
        for (int l = 0; l < 27; ++l) dcoeff[pix * 27 + l] = d[l];
    }
}

// Fused backward: no dcoeff tensor, no sparse matrix -- contract in registers, scatter with atomics.
template <typename T>
__device__ __forceinline__ void backward_pixel(int64_t b, int64_t n, int h, int w, int c, int perspective,
                                               const T *__restrict__ verts, const T *__restrict__ tex,
                                               const int64_t *__restrict__ ids, const T *__restrict__ bary,
                                               const T *__restrict__ gout, T *__restrict__ grad_v, T *__restrict__ grad_tex,
                                               T eps, uint32_t pix, const FastDiv &div_hw, const FastDiv &div_w)
{
    const T w0 = bary[3 * pix], w1 = bary[3 * pix + 1], w2 = bary[3 * pix + 2];
    if (w0 == (T)0 && w1 == (T)0 && w2 == (T)0) return;                 // background
    const int64_t I[3] = {ids[3 * pix], ids[3 * pix + 1], ids[3 * pix + 2]};
    const T wk[3] = {w0, w1, w2};
    T diff[3] = {0, 0, 0};
    for (int ch = 0; ch < c; ++ch) {
        const T gc = gout[pix * c + ch];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            diff[k] += gc * __ldg(tex + I[k] * c + ch);
            if (grad_tex) atomicAdd(grad_tex + I[k] * c + ch, gc * wk[k]);
        }
    }
    if (!grad_v) return;
    T q[9], d[27];
    if (!gather_pixel_tri<T>(verts, I, n * b, q)) return;
    uint32_t img_, rem, py, px;
    div_hw.divmod(pix, img_, rem);
    div_w.divmod(rem, py, px);
    if (!bary_grad<T>(q, (T)(int)px, (T)(int)py, (T)h, (T)w, perspective != 0, eps, d)) return;
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if (a == 2 && !perspective) continue;                       // d/dz is exactly 0 in orthographic mode
            const T s = diff[0] * d[k * 3 + a] + diff[1] * d[9 + k * 3 + a] + diff[2] * d[18 + k * 3 + a];
            atomicAdd(grad_v + I[k] * 3 + a, s);
        }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
raster_backward_kernel(int64_t b, int64_t n, int h, int w, int c, int perspective, const T *__restrict__ verts,
                       const T *__restrict__ tex, const int64_t *__restrict__ ids, const T *__restrict__ bary,
                       const T *__restrict__ gout, T *__restrict__ grad_v, T *__restrict__ grad_tex, T eps,
                       const FastDiv div_hw, const FastDiv div_w)
{
    const uint32_t npix = (uint32_t)(b * h * (int64_t)w);
    const uint32_t stride = gridDim.x * kThreads;
    for (uint32_t pix = blockIdx.x * kThreads + threadIdx.x; pix < npix; pix += stride)
        backward_pixel<T>(b, n, h, w, c, perspective, verts, tex, ids, bary, gout, grad_v, grad_tex, eps, pix, div_hw, div_w);
}

// every level of a pyramid scatters into the SAME grad_v / grad_tex (the sum autograd would form from per-level calls)
template <typename T>
__global__ void __launch_bounds__(kThreads)
raster_backward_pyramid_kernel(int64_t b, int64_t n, const Pyramid<T> L, int c, int perspective,
                               const T *__restrict__ verts, const T *__restrict__ tex, T *__restrict__ grad_v,
                               T *__restrict__ grad_tex, T eps)
{
    const int64_t nblk = L.blk_off[L.n];
    for (int64_t blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        int l = 0;
        while (blk >= L.blk_off[l + 1]) ++l;
        const int size = L.size[l];
        const uint32_t pix = (uint32_t)(blk - L.blk_off[l]) * kThreads + threadIdx.x;
        if (pix < (uint32_t)(b * size * (int64_t)size))
            backward_pixel<T>(b, n, size, size, c, perspective, verts, tex, L.ids[l], L.bary[l], L.gout[l], grad_v, grad_tex,
                              eps, pix, L.div_hw[l], L.div_w[l]);
    }
}

// Workspace of the float path: [keys: npix u64][P: levels x b x nv x 3 floats, 16-byte aligned][F32: b x nf x 3 int32]
// (sized for per-image vertex / triangle lists; shared ones use the first image's slice).
inline int64_t align16(int64_t v) { return (v + 15) & ~(int64_t)15; }
inline int64_t pre_bytes(int64_t b, int64_t nv, int64_t nf, int n_levels) {
    return align16(b * nv * 3 * 4 * n_levels) + align16(b * nf * 3 * 4);
}

int grid_for(int64_t items, int per_sm) {
    int64_t blocks = (items + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)kNumSMs * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

template <typename T>
int rasterize_forward(int64_t b, int64_t nv, int64_t nf, int64_t h, int64_t w, int shared_v, int shared_f,
                      int perspective, const T *verts, const int64_t *tris, int64_t *ids, T *bary,
                      uint64_t *keys, T eps, const T *tex, int64_t c, T *out, void *stream)
{
    constexpr bool kF32 = sizeof(T) == 4;
    SR_REQUIRE(b >= 0 && nv >= 0 && nf >= 0 && h >= 1 && w >= 1, "rasterize: bad sizes");
    SR_REQUIRE(h == w, "rasterize: only square targets (the reference swaps h and w; non-square renders wrongly)");
    SR_REQUIRE(h <= 32768, "rasterize: target too large");
    SR_REQUIRE(nf < 0xffffffffll, "rasterize: too many triangles");
    if (b == 0) return SR_OK;
    SR_REQUIRE(ids && bary && keys, "rasterize: null output/workspace");
    SR_REQUIRE(!tex || (out && c >= 1), "rasterize: tex given without out / channels");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t npix = b * h * w;
    RasterGeom g;
    g.b = b; g.nv = nv; g.nf = nf; g.h = (int)h; g.w = (int)w;
    g.shared_v = shared_v; g.shared_f = shared_f; g.perspective = perspective;
    SR_REQUIRE(npix < 0x7fffffffll, "rasterize: more than 2^31 pixels");
    g.div_hw = FastDiv((uint32_t)(h * w)); g.div_w = FastDiv((uint32_t)w);
    if (eps < 0) eps = -eps;
    uint32_t *idkeys = reinterpret_cast<uint32_t *>(keys + npix);
    // float path: per-vertex / per-triangle pre-pass products behind the keys (see raster_prepass_kernel)
    T *P = nullptr;
    int32_t *F32 = nullptr;
    if (kF32 && nf > 0 && verts && tris) {
        P = reinterpret_cast<T *>(reinterpret_cast<char *>(keys) + align16(npix * 8));
        F32 = reinterpret_cast<int32_t *>(reinterpret_cast<char *>(P) + align16(b * nv * 3 * 4));
    }
    cudaError_t e = cudaMemsetAsync(keys, 0, sizeof(uint64_t) * (size_t)npix, st);
    if (e == cudaSuccess && !kF32) e = cudaMemsetAsync(idkeys, 0xff, sizeof(uint32_t) * (size_t)npix, st);
    if (e != cudaSuccess) { set_error("rasterize: memset: %s", cudaGetErrorString(e)); return (int)e; }
    if (nf > 0 && verts && tris) {
        SR_REQUIRE(b <= 65535, "rasterize: batch too large for one launch");
        int gx = grid_for(nf, 16);
        const int per_img = (int)((int64_t)kNumSMs * 16 / b);              // about 16 CTAs per SM over the whole batch
        if (gx > per_img) gx = per_img < 1 ? 1 : per_img;
        const dim3 grid((unsigned)gx, (unsigned)b);
        if constexpr (kF32) {
            PreLevels pl;
            pl.n = 1; pl.size[0] = (int)h;
            const int64_t items = ((shared_v ? 1 : b) * nv > (shared_f ? 1 : b) * nf) ? (shared_v ? 1 : b) * nv : (shared_f ? 1 : b) * nf;
            raster_prepass_kernel<T><<<grid_for(items, 8), kThreads, 0, st>>>(g, pl, verts, tris, P, F32, eps);
            raster_tri_kernel<T, 0><<<grid, kThreads, 0, st>>>(g, verts, tris, keys, idkeys, eps, P, F32);
            count_launch(2);
        } else {
            raster_tri_kernel<T, 1><<<grid, kThreads, 0, st>>>(g, verts, tris, keys, idkeys, eps);
            raster_tri_kernel<T, 2><<<grid, kThreads, 0, st>>>(g, verts, tris, keys, idkeys, eps);
            count_launch(2);
        }
    }
    const int vec_ok = ((reinterpret_cast<uintptr_t>(ids) | reinterpret_cast<uintptr_t>(bary) |
                         reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
    raster_resolve_kernel<T, kF32><<<grid_for(npix, 8), kThreads, 0, st>>>(g, verts, tris, keys, idkeys, eps, ids, bary,
                                                                          tex, (int)c, out, vec_ok, P, F32);
    count_launch();
    return check_launch("sr_rasterize_forward");
}

template <typename T>
int rasterize_dcoeff(int64_t b, int64_t n, int64_t h, int64_t w, int perspective, const T *verts,
                     const int64_t *ids, T *dcoeff, T eps, void *stream)
{
    SR_REQUIRE(b >= 0 && n >= 0 && h >= 1 && w >= 1, "rasterize_dcoeff: bad sizes");
    if (b == 0) return SR_OK;
    SR_REQUIRE(verts && ids && dcoeff, "rasterize_dcoeff: null pointer");
    if (eps < 0) eps = -eps;
    raster_dcoeff_kernel<T><<<grid_for(b * h * w, 16), kThreads, 0, (cudaStream_t)stream>>>(
        b, n, (int)h, (int)w, perspective, verts, ids, dcoeff, eps);
    count_launch();
    return check_launch("sr_rasterize_dcoeff");
}

template <typename T>
int rasterize_backward(int64_t b, int64_t n, int64_t h, int64_t w, int64_t c, int perspective, const T *verts,
                       const T *tex, const int64_t *ids, const T *bary, const T *gout, T *grad_v, T *grad_tex,
                       T eps, void *stream)
{
    SR_REQUIRE(b >= 0 && n >= 0 && h >= 1 && w >= 1 && c >= 1, "rasterize_backward: bad sizes");
    if (b == 0 || (!grad_v && !grad_tex)) return SR_OK;
    SR_REQUIRE(verts && tex && ids && bary && gout, "rasterize_backward: null pointer");
    if (eps < 0) eps = -eps;
    SR_REQUIRE(b * h * w < 0x7fffffffll, "rasterize_backward: more than 2^31 pixels");
    raster_backward_kernel<T><<<grid_for(b * h * w, 16), kThreads, 0, (cudaStream_t)stream>>>(
        b, n, (int)h, (int)w, (int)c, perspective, verts, tex, ids, bary, gout, grad_v, grad_tex, eps,
        FastDiv((uint32_t)(h * w)), FastDiv((uint32_t)w));
    count_launch();
    return check_launch("sr_rasterize_backward");
}

// levels -> device table; `backward` keeps only the levels that carry a gradient
template <typename T>
int build_pyramid(Pyramid<T> &L, int64_t b, int n_levels, const sr_raster_level *levels, bool backward, bool &vec_ok,
                  bool maps_only = false)
{
    SR_REQUIRE(n_levels >= 1 && n_levels <= kMaxLevels && levels, "rasterize_pyramid: 1..SR_RASTER_MAX_LEVELS levels");
    L.n = 0;
    L.planar = 0;
    L.blk_off[0] = 0;
    int64_t keys = 0;
    uintptr_t bits = 0;
    for (int i = 0; i < n_levels; ++i) {
        const sr_raster_level &lv = levels[i];
        SR_REQUIRE(lv.size >= 1 && lv.size <= 32768, "rasterize_pyramid: bad level size");
        const int64_t npix = b * lv.size * lv.size;
        if (backward && !lv.gout) continue;
        SR_REQUIRE(maps_only || (lv.ids && lv.bary), "rasterize_pyramid: null ids / bary");
        const int l = L.n++;
        SR_REQUIRE(npix < 0x7fffffffll, "rasterize_pyramid: more than 2^31 pixels in a level");
        L.size[l] = (int)lv.size;
        L.div_hw[l] = FastDiv((uint32_t)(lv.size * lv.size)); L.div_w[l] = FastDiv((uint32_t)lv.size);
        L.key_off[l] = keys;
        L.blk_off[l + 1] = L.blk_off[l] + (npix + kThreads - 1) / kThreads;
        L.ids[l] = lv.ids;
        L.bary[l] = reinterpret_cast<T *>(lv.bary);
        L.out[l] = reinterpret_cast<T *>(lv.out);
        L.gout[l] = reinterpret_cast<const T *>(lv.gout);
        keys += npix;
        bits |= reinterpret_cast<uintptr_t>(lv.ids) | reinterpret_cast<uintptr_t>(lv.bary) | reinterpret_cast<uintptr_t>(lv.out);
    }
    vec_ok = (bits & 15u) == 0;
    return SR_OK;
}

int rasterize_pyramid_forward(int64_t b, int64_t nv, int64_t nf, int n_levels, const sr_raster_level *levels, int shared_v,
                              int shared_f, int perspective, const float *verts, const int64_t *tris, uint64_t *keys,
                              float eps, const float *tex, int64_t c, void *stream, bool maps_only = false, bool planar = false)
{
    SR_REQUIRE(b >= 0 && nv >= 0 && nf >= 0, "rasterize_pyramid: bad sizes");
    SR_REQUIRE(nf < 0xffffffffll, "rasterize_pyramid: too many triangles");
    SR_REQUIRE(!maps_only || tex, "rasterize_pyramid_maps: needs tex");
    if (b == 0) return SR_OK;
    Pyramid<float> L;
    bool vec_ok = false;
    if (int rc = build_pyramid<float>(L, b, n_levels, levels, false, vec_ok, maps_only)) return rc;
    L.planar = planar ? 1 : 0;
    SR_REQUIRE(keys, "rasterize_pyramid: null workspace");
    for (int l = 0; l < L.n; ++l) SR_REQUIRE(!tex || L.out[l], "rasterize_pyramid: tex given without out");
    SR_REQUIRE(!tex || c >= 1, "rasterize_pyramid: tex given without channels");
    cudaStream_t st = (cudaStream_t)stream;
    RasterGeom g;
    g.b = b; g.nv = nv; g.nf = nf; g.h = g.w = 0;
    g.shared_v = shared_v; g.shared_f = shared_f; g.perspective = perspective;
    if (eps < 0) eps = -eps;
    const int64_t nkeys = L.key_off[L.n - 1] + b * (int64_t)L.size[L.n - 1] * L.size[L.n - 1];
    cudaError_t e = cudaMemsetAsync(keys, 0, sizeof(uint64_t) * (size_t)nkeys, st);
    if (e != cudaSuccess) { set_error("rasterize_pyramid: memset: %s", cudaGetErrorString(e)); return (int)e; }
    float *P = reinterpret_cast<float *>(reinterpret_cast<char *>(keys) + align16(nkeys * 8));
    int32_t *F32 = reinterpret_cast<int32_t *>(reinterpret_cast<char *>(P) + align16(b * nv * 3 * 4 * (int64_t)L.n));
    if (nf > 0 && verts && tris) {
        SR_REQUIRE(b <= 65535, "rasterize_pyramid: batch too large for one launch");
        PreLevels pl;
        pl.n = L.n;
        for (int l = 0; l < L.n; ++l) pl.size[l] = L.size[l];
        const int64_t nvert = (shared_v ? 1 : b) * nv, ntri = (shared_f ? 1 : b) * nf;
        raster_prepass_kernel<float><<<grid_for(nvert > ntri ? nvert : ntri, 8), kThreads, 0, st>>>(g, pl, verts, tris, P, F32, eps);
        int gx = grid_for(nf, 16);
        const int per_img = (int)((int64_t)kNumSMs * 16 / b);
        if (gx > per_img) gx = per_img < 1 ? 1 : per_img;
        raster_tri_pyramid_kernel<float><<<dim3((unsigned)gx, (unsigned)b), kThreads, 0, st>>>(g, L, verts, tris, keys, eps, P, F32);
        count_launch(2);
    }
    const int64_t nblk = L.blk_off[L.n];
    const int64_t cap = (int64_t)kNumSMs * 8;
    raster_resolve_pyramid_kernel<float><<<(unsigned)(nblk < cap ? nblk : cap), kThreads, 0, st>>>(
        g, L, verts, tris, keys, eps, tex, (int)c, vec_ok ? 1 : 0, P, F32);
    count_launch();
    return check_launch("sr_rasterize_pyramid_forward_f32");
}

int rasterize_pyramid_backward(int64_t b, int64_t n, int n_levels, const sr_raster_level *levels, int64_t c, int perspective,
                               const float *verts, const float *tex, float *grad_v, float *grad_tex, float eps, void *stream)
{
    SR_REQUIRE(b >= 0 && n >= 0 && c >= 1, "rasterize_pyramid_backward: bad sizes");
    if (b == 0 || (!grad_v && !grad_tex)) return SR_OK;
    Pyramid<float> L;
    bool vec_ok = false;
    if (int rc = build_pyramid<float>(L, b, n_levels, levels, true, vec_ok)) return rc;
    if (L.n == 0) return SR_OK;                               // no level carries a gradient
    SR_REQUIRE(verts && tex, "rasterize_pyramid_backward: null pointer");
    if (eps < 0) eps = -eps;
    const int64_t nblk = L.blk_off[L.n];
    const int64_t cap = (int64_t)kNumSMs * 16;
    raster_backward_pyramid_kernel<float><<<(unsigned)(nblk < cap ? nblk : cap), kThreads, 0, (cudaStream_t)stream>>>(
        b, n, L, (int)c, perspective, verts, tex, grad_v, grad_tex, eps);
    count_launch();
    return check_launch("sr_rasterize_pyramid_backward_f32");
}

}  // namespace
}  // namespace sr

using namespace sr;

extern "C" int64_t sr_rasterize_pyramid_workspace_bytes(int64_t b, int64_t nv, int64_t nf, int n_levels, const int64_t *sizes) {
    int64_t npix = 0;
    for (int i = 0; i < n_levels; ++i) npix += b * sizes[i] * sizes[i];
    return align16(npix * 8) + pre_bytes(b, nv, nf, n_levels) + 16;
}
extern "C" int sr_rasterize_pyramid_forward_f32(int64_t b, int64_t nv, int64_t nf, int n_levels,
                                                const sr_raster_level *levels, int shared_v, int shared_f, int perspective,
                                                const float *verts, const int64_t *tris, uint64_t *keys, float eps,
                                                const float *tex, int64_t c, void *stream) {
    return rasterize_pyramid_forward(b, nv, nf, n_levels, levels, shared_v, shared_f, perspective, verts, tris, keys, eps, tex,
                                     c, stream);
}
extern "C" int sr_rasterize_pyramid_maps_f32(int64_t b, int64_t nv, int64_t nf, int n_levels, const sr_raster_level *levels,
                                             int shared_v, int shared_f, int perspective, const float *verts,
                                             const int64_t *tris, uint64_t *keys, float eps, const float *tex, int64_t c,
                                             int planar, void *stream) {
    return rasterize_pyramid_forward(b, nv, nf, n_levels, levels, shared_v, shared_f, perspective, verts, tris, keys, eps, tex,
                                     c, stream, true, planar != 0);
}
extern "C" int sr_rasterize_pyramid_backward_f32(int64_t b, int64_t n, int n_levels, const sr_raster_level *levels, int64_t c,
                                                 int perspective, const float *verts, const float *tex, float *grad_verts,
                                                 float *grad_tex, float eps, void *stream) {
    return rasterize_pyramid_backward(b, n, n_levels, levels, c, perspective, verts, tex, grad_verts, grad_tex, eps, stream);
}

extern "C" int64_t sr_rasterize_workspace_bytes(int64_t b, int64_t nv, int64_t nf, int64_t h, int64_t w, int is_f64) {
    const int64_t npix = b * h * w;
    return align16(npix * 8) + (is_f64 ? npix * 4 : pre_bytes(b, nv, nf, 1)) + 16;
}

extern "C" int sr_rasterize_forward_f32(int64_t b, int64_t nv, int64_t nf, int64_t h, int64_t w, int shared_v,
                                        int shared_f, int perspective, const float *verts, const int64_t *tris,
                                        int64_t *ids, float *bary, uint64_t *keys, float eps, const float *tex,
                                        int64_t c, float *out, void *stream) {
    return rasterize_forward<float>(b, nv, nf, h, w, shared_v, shared_f, perspective, verts, tris, ids, bary, keys,
                                    eps, tex, c, out, stream);
}
extern "C" int sr_rasterize_forward_f64(int64_t b, int64_t nv, int64_t nf, int64_t h, int64_t w, int shared_v,
                                        int shared_f, int perspective, const double *verts, const int64_t *tris,
                                        int64_t *ids, double *bary, uint64_t *keys, double eps, const double *tex,
                                        int64_t c, double *out, void *stream) {
    return rasterize_forward<double>(b, nv, nf, h, w, shared_v, shared_f, perspective, verts, tris, ids, bary, keys,
                                     eps, tex, c, out, stream);
}
extern "C" int sr_rasterize_dcoeff_f32(int64_t b, int64_t n, int64_t h, int64_t w, int perspective,
                                       const float *verts, const int64_t *ids, float *dcoeff, float eps, void *stream) {
    return rasterize_dcoeff<float>(b, n, h, w, perspective, verts, ids, dcoeff, eps, stream);
}
extern "C" int sr_rasterize_dcoeff_f64(int64_t b, int64_t n, int64_t h, int64_t w, int perspective,
                                       const double *verts, const int64_t *ids, double *dcoeff, double eps, void *stream) {
    return rasterize_dcoeff<double>(b, n, h, w, perspective, verts, ids, dcoeff, eps, stream);
}
extern "C" int sr_rasterize_backward_f32(int64_t b, int64_t n, int64_t h, int64_t w, int64_t c, int perspective,
                                         const float *verts, const float *tex, const int64_t *ids, const float *bary,
                                         const float *gout, float *grad_verts, float *grad_tex, float eps, void *stream) {
    return rasterize_backward<float>(b, n, h, w, c, perspective, verts, tex, ids, bary, gout, grad_verts, grad_tex, eps, stream);
}
extern "C" int sr_rasterize_backward_f64(int64_t b, int64_t n, int64_t h, int64_t w, int64_t c, int perspective,
                                         const double *verts, const double *tex, const int64_t *ids, const double *bary,
                                         const double *gout, double *grad_verts, double *grad_tex, double eps, void *stream) {
    return rasterize_backward<double>(b, n, h, w, c, perspective, verts, tex, ids, bary, gout, grad_verts, grad_tex, eps, stream);
}
