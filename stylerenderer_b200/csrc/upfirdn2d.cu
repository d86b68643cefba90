// upfirdn2d (zero-stuff -> pad -> FIR -> decimate) for sm_100a.
//
// Replaces the reference's upfirdn2d_op / upfirdn2d_kernel / upfirdn2d_kernel_large
// (reference op/upfirdn2d.cpp:2-83, op/upfirdn2d_kernel.cu:79-257).  Semantics are those of
// `upfirdn2d_native` (reference op/upfirdn2d.py:159-200), restated in oracle/sr_oracle.c:
//   out[o] = sum_k taps[K-1-k] * U[o*down + k - pad0],  U = input zero-stuffed by `up`.
//
// HBM-bound (algorithmic bytes 4*major*(in_h*in_w + out_h*out_w)); the kernels are
// organised so that the FIR costs ~1.3 shared-memory loads and K*K/up^2 FMAs per output:
//   * poly-phase form: for output o only taps k = k0 + j*up contribute, k0 = (pad0 - o*down) mod up,
//     reading input ib + j with ib = (o*down + k0 - pad0) / up -- no multiplies by stuffed zeros;
//   * a CTA stages a zero-padded input tile (several whole planes at low resolution, so 4x4..32x32
//     planes still fill the CTA) in shared memory, each thread then produces a 4 x VY register patch
//     from a register window filled with 128/64-bit shared loads; all phase/tap indices are
//     compile-time (the pad phase is a template parameter);
//   * 128-bit output stores when rows are 16-byte aligned.
// Everything that is not one of the specialised (up, down, taps, phase) combinations, or has
// minor > 1, takes the generic one-thread-per-output kernel.
#include "common.cuh"
#include "tma_host.cuh"
#include <stdlib.h>

namespace sr {
namespace {

constexpr int kThreads = 256;
constexpr int VX = 4;                 // outputs per thread along x

// ---- compile-time poly-phase index helpers (one axis) -----------------------------------------
__host__ __device__ constexpr int cmod(int a, int b) { return ((a % b) + b) % b; }
// first contributing tap of output (base + v), base*down = 0 (mod up), pad phase PH = pad0 mod up
__host__ __device__ constexpr int tap0(int v, int UP, int DOWN, int PH) { return cmod(PH - v * DOWN, UP); }
// input offset of output (base + v) relative to output `base`'s first input sample
__host__ __device__ constexpr int in_off(int v, int UP, int DOWN, int PH) {
    return (v * DOWN + tap0(v, UP, DOWN, PH) - PH) / UP;
}

struct TileGeom {
    int in_h, in_w, out_h, out_w;
    int pqx, pqy;                     // floor(pad0 / up) per axis
    int txs_log2, tys_log2;           // threads per tile along x / y (powers of two), planes/tile = 256 >> (sum)
    int tin_h, tin_w, tin_stride;     // staged input tile (per plane) and its padded row stride
    int tiles_x, tiles_y;             // tiles per plane
    int64_t major;
    FastDiv div_tiles_x, div_tiles_y;
    int vec_store;                    // out rows 16-byte aligned
};

template <int UP, int DOWN, int KH, int KW, int PHX, int PHY, int VY>
__global__ void __launch_bounds__(kThreads)
upfirdn2d_tile_kernel(float *__restrict__ out, const float *__restrict__ x, const float *__restrict__ taps,
                      const TileGeom g)
{
    static_assert(KH % UP == 0 && KW % UP == 0, "tap count must be a multiple of up");
    static_assert((VX * DOWN) % UP == 0 && (VY * DOWN) % UP == 0, "patch must cover whole phases");
    constexpr int JX = KW / UP, JY = KH / UP;                           // taps per output per axis
    constexpr int NWX = in_off(VX - 1, UP, DOWN, PHX) + JX;             // register window
    constexpr int NWY = in_off(VY - 1, UP, DOWN, PHY) + JY;
    constexpr int SX = VX * DOWN / UP, SY = VY * DOWN / UP;             // window step between patches
    constexpr int LDW = (SX % 4 == 0) ? 4 : ((SX % 2 == 0) ? 2 : 1);    // shared load width
    constexpr int NWXP = (NWX + LDW - 1) / LDW * LDW;

    extern __shared__ __align__(16) float s_in[];

    const int txs = 1 << g.txs_log2, tys = 1 << g.tys_log2;
    const int planes_per_tile = kThreads >> (g.txs_log2 + g.tys_log2);
    const int tow = txs * VX, toh = tys * VY;

    // tile decode: blockIdx.x = ((plane_group * tiles_y) + ty) * tiles_x + tx
    uint32_t t = blockIdx.x, tile_x, tile_y, pg;
    g.div_tiles_x.divmod(t, t, tile_x);
    g.div_tiles_y.divmod(t, pg, tile_y);
    const int64_t plane0 = (int64_t)pg * planes_per_tile;
    const int ox0 = tile_x * tow, oy0 = tile_y * toh;
    const int ix0 = ox0 * DOWN / UP - g.pqx, iy0 = oy0 * DOWN / UP - g.pqy;   // tile origin in the input

    // ---- taps -> registers, already flipped: tk[a][b] multiplies window sample with tap index (a, b)
    float tk[KH][KW];
#pragma unroll
    for (int a = 0; a < KH; ++a)
#pragma unroll
        for (int b = 0; b < KW; ++b) tk[a][b] = __ldg(taps + (KH - 1 - a) * KW + (KW - 1 - b));

    // ---- stage the zero-padded input tile with 4-byte cp.async (zero-fill for the padding): every lane
    //      issues all of its copies back to back, so the whole tile is in flight at once (the rows of a
    //      (2H+1)-wide plane are not 16-byte aligned, which rules out TMA / 16-byte copies in this layout).
    //      Rows are distributed over sub-warps of `lpr` lanes.
    {
        // only the part of the tile that produces in-range outputs is staged (edge tiles of odd-sized planes are
        // mostly empty), and tiles that lie fully inside the plane skip the per-element bounds tests
        const int vw = min(tow, g.out_w - ox0), vh = min(toh, g.out_h - oy0);
        const int need_cols = min(g.tin_stride, ((vw + VX - 1) / VX - 1) * SX + NWXP);
        const int need_rows = min(g.tin_h, ((vh + VY - 1) / VY - 1) * SY + NWY);
        int lpr = 32;
        while (lpr > 1 && (lpr >> 1) >= need_cols) lpr >>= 1;
        const int rows_per_pass = kThreads / lpr;
        const int sub = threadIdx.x / lpr, l = threadIdx.x % lpr;
        const int total_rows = planes_per_tile * need_rows;
        int p = sub / need_rows, ry = sub - p * need_rows;           // one division, then incremental
        const int dp = rows_per_pass / need_rows, dr = rows_per_pass - dp * need_rows;
        const bool interior = ix0 >= 0 && ix0 + need_cols <= g.in_w && iy0 >= 0 && iy0 + need_rows <= g.in_h &&
                              plane0 + planes_per_tile <= g.major;
        for (int r = sub; r < total_rows; r += rows_per_pass) {
            const int iy = iy0 + ry;
            const int64_t plane = plane0 + p;
            const float *src = x + (plane * g.in_h + iy) * (int64_t)g.in_w + ix0;
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_in + ((size_t)p * g.tin_h + ry) * g.tin_stride);
            if (interior) {
                for (int cx = l; cx < need_cols; cx += lpr)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" :: "r"(dst + 4u * cx), "l"(src + cx) : "memory");
            } else {
                const bool row_ok = (iy >= 0) && (iy < g.in_h) && (plane < g.major);
                for (int cx = l; cx < need_cols; cx += lpr) {
                    const int ix = ix0 + cx;
                    const bool ok = row_ok && ix >= 0 && ix < g.in_w;
                    const float *gp = ok ? src + cx : x;
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n"
                                 :: "r"(dst + 4u * cx), "l"(gp), "r"(ok ? 4 : 0) : "memory");
                }
            }
            p += dp; ry += dr;
            if (ry >= need_rows) { ry -= need_rows; ++p; }
        }
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
    }
    __syncthreads();

    // ---- thread -> (plane, patch) and its register window
    const int sx = threadIdx.x & (txs - 1);
    const int sy = (threadIdx.x >> g.txs_log2) & (tys - 1);
    const int p = threadIdx.x >> (g.txs_log2 + g.tys_log2);
    const int64_t plane = plane0 + p;
    const float *wbase = s_in + ((size_t)p * g.tin_h + sy * SY) * g.tin_stride + sx * SX;
    if (plane >= g.major || ox0 + sx * VX >= g.out_w || oy0 + sy * VY >= g.out_h) return;

    const int ox = ox0 + sx * VX, oyb = oy0 + sy * VY;
    if (plane >= g.major || ox >= g.out_w) return;

    if (UP == 1 && DOWN == 1 && VY > 2) {
        // Streaming strip (the blur layers): input rows flow through a 4-deep ring of open output rows, so a thread
        // produces VX x VY outputs from (VY + KH - 1) x 2 vector loads.  Rank-1 taps (every FIR the model builds is
        // an outer product, reference layers.py:7-12) take the separable form: KW FMAs for the row filter plus KH to
        // scatter it over the open rows = 8 instead of 16 FMAs per output; the kernel was issue bound, not HBM bound.
        int pa = 0, pb = 0;
        float best = 0.0f;
#pragma unroll
        for (int a = 0; a < KH; ++a)
#pragma unroll
            for (int b = 0; b < KW; ++b)
                if (fabsf(tk[a][b]) > best) { best = fabsf(tk[a][b]); pa = a; pb = b; }
        float kv[KH], kh[KW];
        bool sep = best > 0.0f;
        float pivot = 1.0f;
#pragma unroll
        for (int a = 0; a < KH; ++a)
#pragma unroll
            for (int b = 0; b < KW; ++b)
                if (a == pa && b == pb) pivot = tk[a][b];
#pragma unroll
        for (int a = 0; a < KH; ++a) {
            kv[a] = 0.0f;
#pragma unroll
            for (int b = 0; b < KW; ++b) if (b == pb) kv[a] = tk[a][b];
        }
#pragma unroll
        for (int b = 0; b < KW; ++b) {
            kh[b] = 0.0f;
#pragma unroll
            for (int a = 0; a < KH; ++a) if (a == pa) kh[b] = tk[a][b] / pivot;
        }
#pragma unroll
        for (int a = 0; a < KH; ++a)
#pragma unroll
            for (int b = 0; b < KW; ++b)
                sep = sep && fabsf(kv[a] * kh[b] - tk[a][b]) <= 1e-6f * best;

        float acc[KH][VX];
#pragma unroll
        for (int a = 0; a < KH; ++a)
#pragma unroll
            for (int vx = 0; vx < VX; ++vx) acc[a][vx] = 0.0f;
#pragma unroll
        for (int ir = 0; ir < VY + KH - 1; ++ir) {
            const float *row = wbase + ir * g.tin_stride;
            float v[NWXP];
#pragma unroll
            for (int b = 0; b < NWXP; b += 4) {
                const float4 q = *reinterpret_cast<const float4 *>(row + b);
                v[b] = q.x; v[b + 1] = q.y; v[b + 2] = q.z; v[b + 3] = q.w;
            }
            if (sep) {
                float hrow[VX];
#pragma unroll
                for (int vx = 0; vx < VX; ++vx) {
                    float t = 0.0f;
#pragma unroll
                    for (int b = 0; b < KW; ++b) t = fmaf(v[vx + b], kh[b], t);
                    hrow[vx] = t;
                }
#pragma unroll
                for (int a = 0; a < KH; ++a) {
                    const int orow = ir - a;                      // output row this input row feeds through tap row a
                    if (orow < 0 || orow >= VY) continue;
#pragma unroll
                    for (int vx = 0; vx < VX; ++vx) acc[orow % KH][vx] = fmaf(hrow[vx], kv[a], acc[orow % KH][vx]);
                }
            } else {
#pragma unroll
                for (int a = 0; a < KH; ++a) {
                    const int orow = ir - a;
                    if (orow < 0 || orow >= VY) continue;
#pragma unroll
                    for (int vx = 0; vx < VX; ++vx)
#pragma unroll
                        for (int b = 0; b < KW; ++b) acc[orow % KH][vx] = fmaf(v[vx + b], tk[a][b], acc[orow % KH][vx]);
                }
            }
            const int done = ir - (KH - 1);                       // output row completed by this input row
            if (done >= 0) {
                const int oy = oyb + done;
                if (oy < g.out_h) {
                    float *dst = out + (plane * g.out_h + oy) * (int64_t)g.out_w + ox;
                    if (g.vec_store && ox + VX <= g.out_w) {
                        *reinterpret_cast<float4 *>(dst) = make_float4(acc[done % KH][0], acc[done % KH][1], acc[done % KH][2], acc[done % KH][3]);
                    } else {
#pragma unroll
                        for (int vx = 0; vx < VX; ++vx)
                            if (ox + vx < g.out_w) dst[vx] = acc[done % KH][vx];
                    }
                }
#pragma unroll
                for (int vx = 0; vx < VX; ++vx) acc[done % KH][vx] = 0.0f;
            }
        }
        return;
    }

    float win[NWY][NWXP];
#pragma unroll
    for (int a = 0; a < NWY; ++a) {
        const float *row = wbase + a * g.tin_stride;
#pragma unroll
        for (int b = 0; b < NWXP; b += LDW) {
            if (LDW == 4) {
                float4 v = *reinterpret_cast<const float4 *>(row + b);
                win[a][b] = v.x; win[a][b + 1] = v.y; win[a][b + 2] = v.z; win[a][b + 3] = v.w;
            } else if (LDW == 2) {
                float2 v = *reinterpret_cast<const float2 *>(row + b);
                win[a][b] = v.x; win[a][b + 1] = v.y;
            } else {
                win[a][b] = row[b];
            }
        }
    }

#pragma unroll
    for (int vy = 0; vy < VY; ++vy) {
        const int oy = oyb + vy;
        if (oy >= g.out_h) break;
        float acc[VX];
#pragma unroll
        for (int vx = 0; vx < VX; ++vx) {
            float s = 0.0f;
#pragma unroll
            for (int jy = 0; jy < JY; ++jy)
#pragma unroll
                for (int jx = 0; jx < JX; ++jx)
                    s = fmaf(win[in_off(vy, UP, DOWN, PHY) + jy][in_off(vx, UP, DOWN, PHX) + jx],
                             tk[tap0(vy, UP, DOWN, PHY) + jy * UP][tap0(vx, UP, DOWN, PHX) + jx * UP], s);
            acc[vx] = s;
        }
        float *dst = out + (plane * g.out_h + oy) * (int64_t)g.out_w + ox;
        if (g.vec_store && ox + VX <= g.out_w) {
            *reinterpret_cast<float4 *>(dst) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        } else {
#pragma unroll
            for (int vx = 0; vx < VX; ++vx)
                if (ox + vx < g.out_w) dst[vx] = acc[vx];
        }
    }
}

// ---- NCHW blur (up = down = 1, 4x4 taps) straight from global memory, no shared-memory staging ------------------------
// The tile kernel above stages its input with 4-byte cp.async because the rows of a (2H+1)-wide plane are not 16-byte
// aligned -- one copy instruction plus address arithmetic per input element, which is what bounds it (ncu: sm 75 %,
// DRAM 32 %).  Here a thread owns 4 output columns x RY rows and reads, per input row, the THREE aligned 16-byte quads
// that cover the 7 floats it needs from the flat [major*H*W] array (consecutive lanes -> consecutive quads: every load
// instruction of a warp is one contiguous 512-byte run, neighbouring lanes' quads overlap in L1).  The position of the
// 7 floats inside the 12 loaded ones is the row's alignment phase (address mod 4), the same for every lane of a warp, so
// the FIR body is instantiated for the four phases and selected by a warp-uniform switch.  Rows stream through a 4-deep
// ring of open output rows; rank-1 taps use the separable form (8 FMAs per output).
constexpr int BR_RY = 16;                 // output rows per thread

struct RowsGeom {
    int in_h, in_w, out_h, out_w, pad_x0, pad_y0, txs, tys, bands;
    int64_t major, total;                 // total floats of the input
    FastDiv div_bands;
};

__device__ __forceinline__ float4 br_load_quad(const float *__restrict__ x, int64_t q, int64_t total) {
    if (q >= 0 && q + 4 <= total) return __ldg(reinterpret_cast<const float4 *>(x + q));
    float v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = (q + i >= 0 && q + i < total) ? __ldg(x + q + i) : 0.0f;
    return make_float4(v[0], v[1], v[2], v[3]);
}

template <int P>
__device__ __forceinline__ void br_row(const float (&f)[12], int c0, int in_w, bool edge, bool sep, const float (&kh)[4],
                                       const float (&kv)[4], const float (&tk)[4][4], float (&acc)[4][4], int ir)
{
    float u[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) u[i] = f[P + i];
    if (edge) {
#pragma unroll
        for (int i = 0; i < 7; ++i)
            if (c0 + i < 0 || c0 + i >= in_w) u[i] = 0.0f;
    }
    if (sep) {
        float h[4];
#pragma unroll
        for (int vx = 0; vx < 4; ++vx) {
            float t = 0.0f;
#pragma unroll
            for (int b = 0; b < 4; ++b) t = fmaf(u[vx + b], kh[b], t);
            h[vx] = t;
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int slot = (ir - a) & 3;                      // ring slot of output row ir - a (rows outside the strip are
#pragma unroll                                                  // never stored, their slots are reset before reuse)
            for (int vx = 0; vx < 4; ++vx) acc[slot][vx] = fmaf(h[vx], kv[a], acc[slot][vx]);
        }
    } else {
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int slot = (ir - a) & 3;
#pragma unroll
            for (int vx = 0; vx < 4; ++vx)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[slot][vx] = fmaf(u[vx + b], tk[a][b], acc[slot][vx]);
        }
    }
}

__global__ void __launch_bounds__(kThreads)
upfirdn2d_blur_rows_kernel(float *__restrict__ out, const float *__restrict__ x, const float *__restrict__ taps,
                           const RowsGeom g)
{
    uint32_t band, plane;
    g.div_bands.divmod(blockIdx.x, plane, band);
    const int sx = threadIdx.x % g.txs, sy = threadIdx.x / g.txs;
    const int ox = sx * 4, oy0 = (band * g.tys + sy) * BR_RY;
    if (ox >= g.out_w || oy0 >= g.out_h) return;

    float tk[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) tk[a][b] = __ldg(taps + (3 - a) * 4 + (3 - b));
    // rank-1 factorisation tk[a][b] = kv[a] * kh[b] through the largest tap (see upfirdn2d_tile_kernel)
    int pa = 0, pb = 0;
    float best = 0.0f;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
            if (fabsf(tk[a][b]) > best) { best = fabsf(tk[a][b]); pa = a; pb = b; }
    float kv[4], kh[4], pivot = 1.0f;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
            if (a == pa && b == pb) pivot = tk[a][b];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        kv[a] = 0.0f; kh[a] = 0.0f;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            if (b == pb) kv[a] = tk[a][b];
            if (b == pa) kh[a] = tk[b][a] / pivot;
        }
    }
    bool sep = best > 0.0f;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) sep = sep && fabsf(kv[a] * kh[b] - tk[a][b]) <= 1e-6f * best;

    const int c0 = ox - g.pad_x0;                               // first input column of this thread's window
    const bool edge = c0 < 0 || c0 + 6 >= g.in_w;
    const int64_t plane_base = (int64_t)plane * g.in_h * g.in_w;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int vx = 0; vx < 4; ++vx) acc[a][vx] = 0.0f;
    const int rows = min(BR_RY, g.out_h - oy0);
#pragma unroll 1
    for (int ir = 0; ir < rows + 3; ++ir) {
        const int iy = oy0 - g.pad_y0 + ir;
        if (iy >= 0 && iy < g.in_h) {                           // rows of the vertical padding contribute nothing
            const int64_t A = plane_base + (int64_t)iy * g.in_w + c0;
            const int64_t Q = A & ~(int64_t)3;
            const int phase = (int)(A - Q);
            float f[12];
#pragma unroll
            for (int qd = 0; qd < 3; ++qd) {
                const float4 v = br_load_quad(x, Q + 4 * qd, g.total);
                f[4 * qd] = v.x; f[4 * qd + 1] = v.y; f[4 * qd + 2] = v.z; f[4 * qd + 3] = v.w;
            }
            switch (phase) {
                case 0: br_row<0>(f, c0, g.in_w, edge, sep, kh, kv, tk, acc, ir); break;
                case 1: br_row<1>(f, c0, g.in_w, edge, sep, kh, kv, tk, acc, ir); break;
                case 2: br_row<2>(f, c0, g.in_w, edge, sep, kh, kv, tk, acc, ir); break;
                default: br_row<3>(f, c0, g.in_w, edge, sep, kh, kv, tk, acc, ir); break;
            }
        }
        const int done = ir - 3;                                // output row completed by this input row
        if (done >= 0) {
            const int slot = done & 3;
            float *dst = out + ((int64_t)plane * g.out_h + oy0 + done) * (int64_t)g.out_w + ox;
            if (ox + 4 <= g.out_w && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
                *reinterpret_cast<float4 *>(dst) = make_float4(acc[slot][0], acc[slot][1], acc[slot][2], acc[slot][3]);
            } else {
#pragma unroll
                for (int vx = 0; vx < 4; ++vx)
                    if (ox + vx < g.out_w) dst[vx] = acc[slot][vx];
            }
        }
        // the slot of output row ir + 1 (first touched by the next input row) must start from zero
#pragma unroll
        for (int vx = 0; vx < 4; ++vx) acc[(ir + 1) & 3][vx] = 0.0f;
    }
}

int launch_blur_rows(float *out, const float *x, const float *taps, int64_t major, int in_h, int in_w, int oh, int ow,
                     int pad_x0, int pad_y0, cudaStream_t st)
{
    // Opt-in (SR_UPFIRDN_ROWS=1): measured SLOWER than the shared-memory strip kernel (blur 257^2 -> 256^2: 0.90 vs
    // 0.81 ms; its backward 256^2 -> 257^2: 1.50 vs 1.02 ms, profiles/r1_kernel_bench_strip.jsonl) -- the three
    // overlapping quads per thread and row cost more L1 bandwidth than the 4-byte staging costs issue slots.
    static const char *on = getenv("SR_UPFIRDN_ROWS");
    if (!(on && on[0] == '1')) return SR_ERR_UNSUPPORTED;
    if (ow < 128 || pad_x0 < 0 || pad_x0 > 3 || pad_y0 < 0 || pad_y0 > 3 || (reinterpret_cast<uintptr_t>(x) & 15u))
        return SR_ERR_UNSUPPORTED;
    RowsGeom g;
    g.in_h = in_h; g.in_w = in_w; g.out_h = oh; g.out_w = ow; g.pad_x0 = pad_x0; g.pad_y0 = pad_y0;
    g.major = major; g.total = major * in_h * in_w;
    int txs = 32;
    while (txs * 4 < ow && txs < kThreads) txs <<= 1;           // threads across a row: 32, 64, 128 or 256
    if (txs * 4 < ow) return SR_ERR_UNSUPPORTED;                // rows wider than 1024 outputs: tile kernel
    g.txs = txs; g.tys = kThreads / txs;
    g.bands = (oh + g.tys * BR_RY - 1) / (g.tys * BR_RY);
    g.div_bands = FastDiv((uint32_t)g.bands);
    const int64_t blocks = major * g.bands;
    if (blocks >= 0x7fffffffll) return SR_ERR_UNSUPPORTED;
    upfirdn2d_blur_rows_kernel<<<(unsigned)blocks, kThreads, 0, st>>>(out, x, taps, g);
    return SR_OK;
}

// ---- channels-last (NHWC, minor = C) FIR, up = down = 1: the layout of the tcgen05 conv pipeline ---------
// One thread = 4 channels (float4) x 2 adjacent output columns x a strip of RY output rows.  Consecutive lanes
// take consecutive channel quads, so every load/store instruction of a warp is one contiguous 512-byte run;
// a KH-row x (KW+1)-column register window slides down the strip, so each input element is fetched from
// L1/L2 ~2.5 times per output instead of KH*KW times.
__device__ __forceinline__ float round_tf32_(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}

struct NhwcGeom {
    int64_t major, total_threads;
    int in_h, in_w, out_h, out_w, c4, pad_x0, pad_y0, strips_y, pairs_x, rows_per_strip;
    FastDiv div_c4, div_pairs, div_strips;
    // STYLED epilogue: y = lrelu(acc + noise_weight * noise[n, oy, ox] + bias[c]) * gain
    const float *noise, *noise_weight, *bias;
    long long noise_bstride;
    float alpha, gain;
    float *out2;                  // optional: tf32(y * scale2[n,c])
    const float *scale2;
    // SCALEDOT epilogue (backward of the up-sampling block): out = tf32(acc * scale2[n,c]), dot[n,c] += sum acc * other
    const float *other;
    float *dot;
    // STYLED with a style map (StyledMapConv, reference model.py:50): y = lrelu(acc * map0 + map1 + noise + bias) * gain;
    // `out` then receives acc (what the backward needs; map0 may be exactly 0) and only out2 sees y
    const float *stylemap;        // [B, 2, out_h, out_w] planes (batch stride map_bstride) or nullptr
    long long map_bstride;
    int op16;                     // the GEMM operand this pass produces (STYLED: out2, SCALEDOT: out) is a bfloat16 tensor
};

__device__ __forceinline__ uint2 pack4_bf16(float a, float b, float c, float d) {
    uint2 r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r.x) : "f"(b), "f"(a));
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r.y) : "f"(d), "f"(c));
    return r;
}

template <int KH, int KW, int MODE>     // MODE 0: plain, 1: STYLED forward tail, 2: SCALEDOT backward tail
__global__ void __launch_bounds__(kThreads, 2)      // <= 128 registers: two CTAs per SM keep enough loads in flight
upfirdn2d_nhwc_kernel(float *__restrict__ out, const float *__restrict__ x, const float *__restrict__ taps,
                      const NhwcGeom g)
{
    const int64_t tid = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    constexpr bool STYLED = (MODE == 1);
    if (tid >= g.total_threads) return;
    uint32_t t = (uint32_t)tid, c, xp, ys, n;
    g.div_c4.divmod(t, t, c);
    g.div_pairs.divmod(t, t, xp);
    g.div_strips.divmod(t, n, ys);

    float tk[KH][KW];
#pragma unroll
    for (int a = 0; a < KH; ++a)
#pragma unroll
        for (int b = 0; b < KW; ++b) tk[a][b] = __ldg(taps + (KH - 1 - a) * KW + (KW - 1 - b));

    const int ox0 = xp * 2;
    const int ix0 = ox0 - g.pad_x0;
    const int oy0 = ys * g.rows_per_strip;
    const int oy1 = min(g.out_h, oy0 + g.rows_per_strip);
    const float4 *xin = reinterpret_cast<const float4 *>(x) + (int64_t)n * g.in_h * g.in_w * g.c4 + c;
    float4 *yout = reinterpret_cast<float4 *>(out) + (int64_t)n * g.out_h * g.out_w * g.c4 + c;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);

    float nw = 0.0f;
    float4 bias4 = zero;
    const float *nz = nullptr;
    if (STYLED) {
        if (g.noise) { nw = __ldg(g.noise_weight); nz = g.noise + (int64_t)n * g.noise_bstride; }
        if (g.bias) bias4 = __ldg(reinterpret_cast<const float4 *>(g.bias) + c);
    }
    float4 sc2 = zero;
    if ((STYLED && g.out2) || MODE == 2) sc2 = __ldg(reinterpret_cast<const float4 *>(g.scale2) + (int64_t)n * g.c4 + c);
    float4 dot = zero;
    const float4 *oth = (MODE == 2 && g.other) ? reinterpret_cast<const float4 *>(g.other) + (int64_t)n * g.out_h * g.out_w * g.c4 + c : nullptr;

    float4 win[KH][KW + 1];
    auto load_row = [&](float4 (&row)[KW + 1], int iy) {
        const bool row_ok = iy >= 0 && iy < g.in_h;
        const float4 *src = xin + (int64_t)iy * g.in_w * g.c4;
#pragma unroll
        for (int b = 0; b < KW + 1; ++b) {
            const int ix = ix0 + b;
            row[b] = (row_ok && ix >= 0 && ix < g.in_w) ? __ldg(src + (int64_t)ix * g.c4) : zero;
        }
    };
#pragma unroll
    for (int a = 0; a < KH - 1; ++a) load_row(win[a], oy0 - g.pad_y0 + a);
    for (int oy = oy0; oy < oy1; ++oy) {
        load_row(win[KH - 1], oy - g.pad_y0 + KH - 1);
        float4 tt[2] = {zero, zero};
        if (MODE == 2 && oth) {                        // issue the loads of the dot operand before the FMA block
#pragma unroll
            for (int j = 0; j < 2; ++j)
                if (ox0 + j < g.out_w) tt[j] = __ldg(oth + ((int64_t)oy * g.out_w + ox0 + j) * g.c4);
        }
        float4 acc[2] = {zero, zero};
#pragma unroll
        for (int a = 0; a < KH; ++a)
#pragma unroll
            for (int b = 0; b < KW; ++b)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const float4 v = win[a][j + b];
                    const float k = tk[a][b];
                    acc[j].x = fmaf(v.x, k, acc[j].x); acc[j].y = fmaf(v.y, k, acc[j].y);
                    acc[j].z = fmaf(v.z, k, acc[j].z); acc[j].w = fmaf(v.w, k, acc[j].w);
                }
        float4 pre[2] = {acc[0], acc[1]};                  // filtered value before the tail (stored instead of y with a stylemap)
        if (STYLED) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                float add = (nz && ox0 + j < g.out_w) ? nw * __ldg(nz + (int64_t)oy * g.out_w + ox0 + j) : 0.0f;
                float m0 = 1.0f;
                if (g.stylemap && ox0 + j < g.out_w) {
                    const float *mp = g.stylemap + (int64_t)n * g.map_bstride + (int64_t)oy * g.out_w + ox0 + j;
                    m0 = __ldg(mp);
                    add += __ldg(mp + (int64_t)g.out_h * g.out_w);
                }
                float t;
                t = fmaf(acc[j].x, m0, add + bias4.x); acc[j].x = ((t > 0.f) ? t : t * g.alpha) * g.gain;
                t = fmaf(acc[j].y, m0, add + bias4.y); acc[j].y = ((t > 0.f) ? t : t * g.alpha) * g.gain;
                t = fmaf(acc[j].z, m0, add + bias4.z); acc[j].z = ((t > 0.f) ? t : t * g.alpha) * g.gain;
                t = fmaf(acc[j].w, m0, add + bias4.w); acc[j].w = ((t > 0.f) ? t : t * g.alpha) * g.gain;
            }
        }
        if (MODE == 2) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (ox0 + j >= g.out_w) break;
                dot.x = fmaf(acc[j].x, tt[j].x, dot.x); dot.y = fmaf(acc[j].y, tt[j].y, dot.y);
                dot.z = fmaf(acc[j].z, tt[j].z, dot.z); dot.w = fmaf(acc[j].w, tt[j].w, dot.w);
                acc[j].x = acc[j].x * sc2.x; acc[j].y = acc[j].y * sc2.y; acc[j].z = acc[j].z * sc2.z; acc[j].w = acc[j].w * sc2.w;
                if (!g.op16) {
                    acc[j].x = round_tf32_(acc[j].x); acc[j].y = round_tf32_(acc[j].y);
                    acc[j].z = round_tf32_(acc[j].z); acc[j].w = round_tf32_(acc[j].w);
                }
            }
        }
        const int64_t opix = (int64_t)n * g.out_h * g.out_w * g.c4 + c + ((int64_t)oy * g.out_w + ox0) * g.c4;   // in channel quads
        if (MODE == 2 && g.op16) {                         // the scaled gradient leaves as a bfloat16 operand
            uint2 *d16 = reinterpret_cast<uint2 *>(out) + opix;
            d16[0] = pack4_bf16(acc[0].x, acc[0].y, acc[0].z, acc[0].w);
            if (ox0 + 1 < g.out_w) d16[g.c4] = pack4_bf16(acc[1].x, acc[1].y, acc[1].z, acc[1].w);
        } else {
            float4 *dst = yout + ((int64_t)oy * g.out_w + ox0) * g.c4;
            const bool keep_pre = STYLED && g.stylemap;
            dst[0] = keep_pre ? pre[0] : acc[0];
            if (ox0 + 1 < g.out_w) dst[g.c4] = keep_pre ? pre[1] : acc[1];
        }
        if (STYLED && g.out2) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (ox0 + j >= g.out_w) break;
                if (g.op16) {
                    reinterpret_cast<uint2 *>(g.out2)[opix + (int64_t)j * g.c4] =
                        pack4_bf16(acc[j].x * sc2.x, acc[j].y * sc2.y, acc[j].z * sc2.z, acc[j].w * sc2.w);
                } else {
                    float4 o;
                    o.x = round_tf32_(acc[j].x * sc2.x);
                    o.y = round_tf32_(acc[j].y * sc2.y);
                    o.z = round_tf32_(acc[j].z * sc2.z);
                    o.w = round_tf32_(acc[j].w * sc2.w);
                    reinterpret_cast<float4 *>(g.out2)[opix + (int64_t)j * g.c4] = o;
                }
            }
        }
#pragma unroll
        for (int a = 0; a < KH - 1; ++a)
#pragma unroll
            for (int b = 0; b < KW + 1; ++b) win[a][b] = win[a + 1][b];
    }
    if (MODE == 2 && g.dot) {   // one 128-bit reduction per thread into dot[n, 4c .. 4c+3]
        float *dp = g.dot + ((int64_t)n * g.c4 + c) * 4;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(dp), "f"(dot.x), "f"(dot.y), "f"(dot.z), "f"(dot.w) : "memory");
    }
}

// ---- channels-last FIR, low-instruction-count form (default since round 2) ---------------------------------------------
// ncu on the 256^2 layer of the generator step (profiles/r2_ncu_hbm_passes.md): upfirdn2d_nhwc_kernel issues 566 M warp
// instructions for 67 M float4 outputs (270 per output: 64 FFMA for the 16 taps, 30 MOV to slide its 4 x 5 register window,
// the rest addressing / predicates / tail) and is ISSUE bound (issue active 67 % at 24 % occupancy, DRAM 55 %).  Here:
//   * rank-1 taps (every FIR the model builds, reference layers.py:7-12) take the separable form: a 4-tap row filter of the
//     incoming input row, then one FMA per open output row -- 32 instead of 64 FFMA per float4 output (general taps keep
//     the 2-D form: same loop, 16 FMAs);
//   * the four open output rows live in a register ring whose slot index is a compile-time constant inside a 4x unrolled
//     row loop -- no register moves;
//   * the next input row is fetched while the current one is consumed; column predicates and row pointers are hoisted.
// Thread decomposition and tails as upfirdn2d_nhwc_kernel (4 channels x 2 columns x a strip of rows).
template <int MODE>
__global__ void __launch_bounds__(kThreads, 2)
fir_nhwc_sep_kernel(float *__restrict__ out, const float *__restrict__ x, const float *__restrict__ taps, const NhwcGeom g)
{
    constexpr int K = 4;
    constexpr bool STYLED = (MODE == 1);
    constexpr bool PREFETCH = (MODE != 2);                 // the scale(+dot) tail needs the registers of the second row buffer
    const int64_t tid = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (tid >= g.total_threads) return;
    uint32_t t = (uint32_t)tid, c, xp, ys, n;
    g.div_c4.divmod(t, t, c);
    g.div_pairs.divmod(t, t, xp);
    g.div_strips.divmod(t, n, ys);

    // flipped taps tk[a][b] = taps[K-1-a][K-1-b]; rank-1 factorisation tk[a][b] = kv[a] * kh[b] through the largest tap
    float kv[K], kh[K];
    bool sep;
    {
        float tk[K][K];
#pragma unroll
        for (int a = 0; a < K; ++a)
#pragma unroll
            for (int b = 0; b < K; ++b) tk[a][b] = __ldg(taps + (K - 1 - a) * K + (K - 1 - b));
        int pa = 0, pb = 0;
        float best = 0.0f;
#pragma unroll
        for (int a = 0; a < K; ++a)
#pragma unroll
            for (int b = 0; b < K; ++b)
                if (fabsf(tk[a][b]) > best) { best = fabsf(tk[a][b]); pa = a; pb = b; }
        float pivot = 1.0f;
#pragma unroll
        for (int a = 0; a < K; ++a)
#pragma unroll
            for (int b = 0; b < K; ++b)
                if (a == pa && b == pb) pivot = tk[a][b];
#pragma unroll
        for (int a = 0; a < K; ++a) {
            kv[a] = 0.0f; kh[a] = 0.0f;
#pragma unroll
            for (int b = 0; b < K; ++b) {
                if (b == pb) kv[a] = tk[a][b];
                if (b == pa) kh[a] = tk[b][a] / pivot;
            }
        }
        sep = best > 0.0f;
#pragma unroll
        for (int a = 0; a < K; ++a)
#pragma unroll
            for (int b = 0; b < K; ++b) sep = sep && fabsf(kv[a] * kh[b] - tk[a][b]) <= 1e-6f * best;
    }

    const int ox0 = xp * 2;
    const int ix0 = ox0 - g.pad_x0;
    const int oy0 = ys * g.rows_per_strip;
    const int oy1 = min(g.out_h, oy0 + g.rows_per_strip);
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool col1 = ox0 + 1 < g.out_w;
    bool cok[K + 1];
#pragma unroll
    for (int b = 0; b < K + 1; ++b) cok[b] = ix0 + b >= 0 && ix0 + b < g.in_w;
    const int64_t row_stride = (int64_t)g.in_w * g.c4;                  // float4 units
    // first input pixel of this thread's window: image n, row (oy0 - pad), column ix0, channel quad c
    const float4 *src = reinterpret_cast<const float4 *>(x) + (int64_t)n * g.in_h * row_stride + (int64_t)(oy0 - g.pad_y0) * row_stride +
                        (int64_t)ix0 * g.c4 + c;
    int iy = oy0 - g.pad_y0;

    float nw = 0.0f;
    float4 bias4 = zero;
    const float *nz = nullptr;
    if (STYLED) {
        if (g.noise) { nw = __ldg(g.noise_weight); nz = g.noise + (int64_t)n * g.noise_bstride; }
        if (g.bias) bias4 = __ldg(reinterpret_cast<const float4 *>(g.bias) + c);
    }
    float4 sc2 = zero;
    if ((STYLED && g.out2) || MODE == 2) sc2 = __ldg(reinterpret_cast<const float4 *>(g.scale2) + (int64_t)n * g.c4 + c);
    float4 dot = zero;
    const int64_t img_quads = (int64_t)n * g.out_h * g.out_w * g.c4 + c;    // this image / channel quad in an output-shaped tensor
    const float4 *oth = (MODE == 2 && g.other) ? reinterpret_cast<const float4 *>(g.other) + img_quads : nullptr;
    const float *smap = (STYLED && g.stylemap) ? g.stylemap + (int64_t)n * g.map_bstride : nullptr;
    const int64_t map_plane = (int64_t)g.out_h * g.out_w;

    auto load_row = [&](float4 (&row)[K + 1], const float4 *p, int y) {
        const bool row_ok = y >= 0 && y < g.in_h;
#pragma unroll
        for (int b = 0; b < K + 1; ++b) row[b] = (row_ok && cok[b]) ? __ldg(p + (int64_t)b * g.c4) : zero;
    };

    float4 acc[K][2];                                      // ring of open output rows: slot of output row o is (o & 3)
#pragma unroll
    for (int s_ = 0; s_ < K; ++s_) { acc[s_][0] = zero; acc[s_][1] = zero; }
    const int nsteps = oy1 - oy0 + K - 1;
    float4 cur[K + 1], nxt[K + 1];
    load_row(cur, src, iy);
    for (int r0 = 0; r0 < nsteps; r0 += K) {
#pragma unroll
        for (int u = 0; u < K; ++u) {
            const int r = r0 + u;
            if (r < nsteps) {
                if (PREFETCH && r + 1 < nsteps) load_row(nxt, src + row_stride, iy + 1);   // in flight while row r is consumed
                if (sep) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        float4 h = zero;
#pragma unroll
                        for (int b = 0; b < K; ++b) {
                            h.x = fmaf(cur[j + b].x, kh[b], h.x); h.y = fmaf(cur[j + b].y, kh[b], h.y);
                            h.z = fmaf(cur[j + b].z, kh[b], h.z); h.w = fmaf(cur[j + b].w, kh[b], h.w);
                        }
#pragma unroll
                        for (int a = 0; a < K; ++a) {          // input row r is tap row a of output row r - a
                            float4 &d = acc[(u - a) & 3][j];
                            d.x = fmaf(h.x, kv[a], d.x); d.y = fmaf(h.y, kv[a], d.y);
                            d.z = fmaf(h.z, kv[a], d.z); d.w = fmaf(h.w, kv[a], d.w);
                        }
                    }
                } else {
#pragma unroll
                    for (int a = 0; a < K; ++a)
#pragma unroll
                        for (int b = 0; b < K; ++b)
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                float4 &d = acc[(u - a) & 3][j];
                                const float k = __ldg(taps + (K - 1 - a) * K + (K - 1 - b));    // (kept out of the register file)
                                d.x = fmaf(cur[j + b].x, k, d.x); d.y = fmaf(cur[j + b].y, k, d.y);
                                d.z = fmaf(cur[j + b].z, k, d.z); d.w = fmaf(cur[j + b].w, k, d.w);
                            }
                }
                const int oy = oy0 + r - (K - 1);              // the output row this step completes: ring slot (u + 1) & 3
                if (r >= K - 1) {
                    float4 o[2] = {acc[(u + 1) & 3][0], acc[(u + 1) & 3][1]};
                    const float4 pre[2] = {o[0], o[1]};
                    const int64_t opix = (int64_t)oy * g.out_w + ox0;
                    if (STYLED) {
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const bool in = (j == 0) || col1;
                            float add = (nz && in) ? nw * __ldg(nz + opix + j) : 0.0f;
                            float m0 = 1.0f;
                            if (smap && in) { m0 = __ldg(smap + opix + j); add += __ldg(smap + map_plane + opix + j); }
                            float v;
                            v = fmaf(o[j].x, m0, add + bias4.x); o[j].x = ((v > 0.f) ? v : v * g.alpha) * g.gain;
                            v = fmaf(o[j].y, m0, add + bias4.y); o[j].y = ((v > 0.f) ? v : v * g.alpha) * g.gain;
                            v = fmaf(o[j].z, m0, add + bias4.z); o[j].z = ((v > 0.f) ? v : v * g.alpha) * g.gain;
                            v = fmaf(o[j].w, m0, add + bias4.w); o[j].w = ((v > 0.f) ? v : v * g.alpha) * g.gain;
                        }
                    }
                    const int64_t q = img_quads + opix * g.c4;      // channel-quad index of (n, oy, ox0, c)
                    if (MODE == 2) {
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            if (j == 1 && !col1) break;
                            if (oth) {
                                const float4 tt = __ldg(oth + (opix + j) * g.c4);
                                dot.x = fmaf(o[j].x, tt.x, dot.x); dot.y = fmaf(o[j].y, tt.y, dot.y);
                                dot.z = fmaf(o[j].z, tt.z, dot.z); dot.w = fmaf(o[j].w, tt.w, dot.w);
                            }
                            o[j].x *= sc2.x; o[j].y *= sc2.y; o[j].z *= sc2.z; o[j].w *= sc2.w;
                            if (g.op16) {
                                reinterpret_cast<uint2 *>(out)[q + (int64_t)j * g.c4] = pack4_bf16(o[j].x, o[j].y, o[j].z, o[j].w);
                            } else {
                                reinterpret_cast<float4 *>(out)[q + (int64_t)j * g.c4] =
                                    make_float4(round_tf32_(o[j].x), round_tf32_(o[j].y), round_tf32_(o[j].z), round_tf32_(o[j].w));
                            }
                        }
                    } else {
                        const bool keep_pre = STYLED && smap;
                        reinterpret_cast<float4 *>(out)[q] = keep_pre ? pre[0] : o[0];
                        if (col1) reinterpret_cast<float4 *>(out)[q + g.c4] = keep_pre ? pre[1] : o[1];
                        if (STYLED && g.out2) {
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                if (j == 1 && !col1) break;
                                const float a0 = o[j].x * sc2.x, a1 = o[j].y * sc2.y, a2 = o[j].z * sc2.z, a3 = o[j].w * sc2.w;
                                if (g.op16) reinterpret_cast<uint2 *>(g.out2)[q + (int64_t)j * g.c4] = pack4_bf16(a0, a1, a2, a3);
                                else reinterpret_cast<float4 *>(g.out2)[q + (int64_t)j * g.c4] =
                                         make_float4(round_tf32_(a0), round_tf32_(a1), round_tf32_(a2), round_tf32_(a3));
                            }
                        }
                    }
                }
                acc[(u + 1) & 3][0] = zero; acc[(u + 1) & 3][1] = zero;      // the slot now belongs to output row r + 1
                src += row_stride;
                ++iy;
                if (PREFETCH) {
#pragma unroll
                    for (int b = 0; b < K + 1; ++b) cur[b] = nxt[b];
                } else if (r + 1 < nsteps) {
                    load_row(cur, src, iy);
                }
            }
        }
    }
    if (MODE == 2 && g.dot) {   // one 128-bit reduction per thread into dot[n, 4c .. 4c+3]
        float *dp = g.dot + ((int64_t)n * g.c4 + c) * 4;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(dp), "f"(dot.x), "f"(dot.y), "f"(dot.z), "f"(dot.w) : "memory");
    }
}

// ---- helpers of the ring kernels below (mbarrier wait, shared-space loads) ---------------------------------------------
__device__ __forceinline__ void ft_mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nFTW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra FTD_%=;\nbra FTW_%=;\nFTD_%=:\n}\n"
        :: "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}

// Shared-memory loads by 32-bit shared address.  The ring kernels carve their stages out of the dynamic shared array with
// integer pointer arithmetic (128-byte alignment), after which the compiler no longer knows the address space and emits
// GENERIC loads (LD.E: long-scoreboard latency) for every stage read; these keep them LDS.  volatile: never moved across the
// mbarrier waits (also volatile asm).
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float lds_f1(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

// Flipped taps tk[a][b] = taps[3-a][3-b] of a 4 x 4 FIR and, when the matrix is rank 1 (every FIR the model builds,
// reference layers.py:7-12), its factorisation tk[a][b] = kv[a] * kh[b] through the largest tap.  Returns whether it is.
__device__ __forceinline__ bool fir_rank1_taps(const float *__restrict__ taps, float (&kv)[4], float (&kh)[4])
{
    constexpr int K = 4;
    float tk[K][K];
#pragma unroll
    for (int a = 0; a < K; ++a)
#pragma unroll
        for (int b = 0; b < K; ++b) tk[a][b] = __ldg(taps + (K - 1 - a) * K + (K - 1 - b));
    int pa = 0, pb = 0;
    float best = 0.0f;
#pragma unroll
    for (int a = 0; a < K; ++a)
#pragma unroll
        for (int b = 0; b < K; ++b)
            if (fabsf(tk[a][b]) > best) { best = fabsf(tk[a][b]); pa = a; pb = b; }
    float pivot = 1.0f;
#pragma unroll
    for (int a = 0; a < K; ++a)
#pragma unroll
        for (int b = 0; b < K; ++b)
            if (a == pa && b == pb) pivot = tk[a][b];
#pragma unroll
    for (int a = 0; a < K; ++a) {
        kv[a] = 0.0f; kh[a] = 0.0f;
#pragma unroll
        for (int b = 0; b < K; ++b) {
            if (b == pb) kv[a] = tk[a][b];
            if (b == pa) kh[a] = tk[b][a] / pivot;
        }
    }
    bool sep = best > 0.0f;
#pragma unroll
    for (int a = 0; a < K; ++a)
#pragma unroll
        for (int b = 0; b < K; ++b) sep = sep && fabsf(kv[a] * kh[b] - tk[a][b]) <= 1e-6f * best;
    return sep;
}

// ---- channels-last FIR, row-streaming through a deep TMA ring (default for planes >= 32 x 32, C % 32 == 0) -------------
// ncu on the generator step (profiles/r2_ncu_hbm_passes.md): the register-window kernels above issue ~270 thread
// instructions per float4 output and saturate the ISSUE slots at 0.5-0.6 of the copy bandwidth; the tile-at-a-time TMA
// kernel waits for a whole 85 KB tile behind a 2-deep ring.  This form streams DOWN a column strip instead:
//   * work item = (image, block of 32 channels, strip of 32 output columns, segment of ~64 output rows); persistent CTAs
//     (2 per SM) walk the item list; the producer lane runs ahead across item boundaries;
//   * the producer fetches 4 input rows x 35 pixels x 128 B per stage with ONE TMA box (zero fill outside the plane = the
//     FIR padding: no bounds test on the load side) into a 6-deep ring -> ~105 KB in flight per CTA, 210 KB per SM;
//   * consumer thread = (output column, channel quad): per input row 4 conflict-free LDS.128, a 4-tap row filter, then ONE
//     FMA per open output row -- the four open rows live in a register ring whose slot index is a compile-time constant of
//     the 4-row stage loop; every step retires one output row: ~75 instructions per float4 output (32 FFMA);
//   * each warp stores 4 pixels x 128 contiguous bytes (whole lines); HBM sees each input byte 1.09x (3 halo columns per
//     32) + 3 halo rows per segment.
constexpr int FS_W = 32, FS_C = 32, FS_ROWS = 4, FS_MAX_STAGES = 6;
constexpr int FS_IW = FS_W + 3;
constexpr int FS_X_BYTES = FS_ROWS * FS_IW * FS_C * 4;              // 17920: the input box of a stage
constexpr int FS_NOISE_BYTES = FS_ROWS * FS_W * 4;                  // 512: noise[n, 4 output rows, 32 columns] (styled tail)
constexpr int FS_MAP_BYTES = 2 * FS_NOISE_BYTES;                    // 1024: the two style-map planes of the same pixels
constexpr int FS_CONSUMERS = 256;

struct FirStreamGeom {
    int in_h, in_w, out_h, out_w, c, cblocks, xtiles, nseg, seg_rows, total_items, pad_x0, pad_y0;
    FastDiv div_cb, div_xt, div_seg;
    const float *noise, *noise_weight, *bias;
    long long noise_bstride;
    float alpha, gain;
    float *out2;
    const float *scale2, *other;
    float *dot;
    const float *stylemap;
    long long map_bstride;
    int op16;
    int stages, stage_bytes;      // ring depth (6, or 5 with a style map) and bytes per stage (input box + side boxes)
    int side_tma;                 // bit 0: noise rows arrive through the ring (tmap_nz), bit 1: style-map rows (tmap_map)
};

// The per-pixel side inputs of the styled tail (noise, style map) ride the same ring as the input rows: fetched with
// per-thread loads they sat between the filter and the stores of every row (ncu: half of all stall samples on the
// `noise_weight * noise` multiply, 0.55 of the copy bandwidth against 0.89 for the tail without side inputs).
template <int MODE>     // MODE 0 plain, 1 styled forward tail (+ optional out2 / style map), 2 scale(+dot) backward tail
__global__ void __launch_bounds__(FS_CONSUMERS + 32, 2)
fir_nhwc_stream_kernel(float *__restrict__ out, const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_nz,
                       const __grid_constant__ CUtensorMap tmap_map, const float *__restrict__ taps, const FirStreamGeom g)
{
    extern __shared__ uint8_t fs_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(fs_raw) + 127) & ~(uintptr_t)127);
    const int FS_STAGES = g.stages, FS_STAGE_BYTES = g.stage_bytes;
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + FS_STAGES * FS_STAGE_BYTES);
    uint64_t *empty_bar = full_bar + FS_MAX_STAGES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool nz_tma = MODE == 1 && (g.side_tma & 1), map_tma = MODE == 1 && (g.side_tma & 2);

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_x) : "memory");
        for (int s = 0; s < FS_STAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(&full_bar[s])), "r"(1u) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(&empty_bar[s])), "r"((uint32_t)(FS_CONSUMERS / 32)) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == FS_CONSUMERS / 32) {                   // ===== producer warp (one lane)
        if (lane == 0) {
            uint32_t s = 0, ph = 0;
            for (uint32_t T = blockIdx.x; T < (uint32_t)g.total_items; T += gridDim.x) {
                uint32_t t = T, cb, xt, seg, n;
                g.div_cb.divmod(t, t, cb);
                g.div_xt.divmod(t, t, xt);
                g.div_seg.divmod(t, n, seg);
                const int oy_start = seg * g.seg_rows;
                const int oy_end = min(g.out_h, oy_start + g.seg_rows);
                const int nstages = (oy_end - oy_start + 3 + FS_ROWS - 1) / FS_ROWS;
                const int cx = (int)(xt * FS_W) - g.pad_x0;
                int cy = oy_start - g.pad_y0;
                const uint32_t tx_bytes = FS_X_BYTES + (nz_tma ? FS_NOISE_BYTES : 0) + (map_tma ? FS_MAP_BYTES : 0);
                for (int k = 0; k < nstages; ++k, cy += FS_ROWS) {
                    ft_mbar_wait(&empty_bar[s], ph ^ 1);
                    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&full_bar[s]);
                    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem + s * FS_STAGE_BYTES);
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(tx_bytes) : "memory");
                    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                                 :: "r"(dst), "l"(&tmap_x), "r"(bar), "r"((int)(cb * FS_C)), "r"(cx), "r"(cy), "r"((int)n) : "memory");
                    const int oy0 = oy_start + k * FS_ROWS - 3;          // first output row this stage retires (rows < 0: zero fill, unused)
                    if (nz_tma)
                        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                                     :: "r"(dst + FS_X_BYTES), "l"(&tmap_nz), "r"(bar), "r"((int)(xt * FS_W)), "r"(oy0),
                                        "r"(g.noise_bstride ? (int)n : 0) : "memory");
                    if (map_tma)
                        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                                     :: "r"(dst + FS_X_BYTES + FS_NOISE_BYTES), "l"(&tmap_map), "r"(bar), "r"((int)(xt * FS_W)), "r"(oy0),
                                        "r"(0), "r"((int)n) : "memory");
                    if (++s == FS_STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
        return;
    }

    // ===== consumers: thread = (output column of the strip, channel quad)
    constexpr int K = 4;
    float kv[K], kh[K];
    const bool sep = fir_rank1_taps(taps, kv, kh);
    const int quad = threadIdx.x & 7, col = threadIdx.x >> 3;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    float nw = 0.0f;
    if (MODE == 1 && g.noise) nw = __ldg(g.noise_weight);
    const long long map_plane = (long long)g.out_h * g.out_w;

    uint32_t s = 0, ph = 0;
    for (uint32_t T = blockIdx.x; T < (uint32_t)g.total_items; T += gridDim.x) {
        uint32_t t = T, cb, xt, seg, n;
        g.div_cb.divmod(t, t, cb);
        g.div_xt.divmod(t, t, xt);
        g.div_seg.divmod(t, n, seg);
        const int ox = xt * FS_W + col;
        const bool col_ok = ox < g.out_w;
        const int oy_start = seg * g.seg_rows;
        const int oy_end = min(g.out_h, oy_start + g.seg_rows);
        const int nstages = (oy_end - oy_start + 3 + FS_ROWS - 1) / FS_ROWS;
        const int ch = cb * FS_C + 4 * quad;
        float4 bias4 = zero, sc2 = zero;
        if (MODE == 1 && g.bias) bias4 = __ldg(reinterpret_cast<const float4 *>(g.bias + ch));
        if (MODE == 2) sc2 = make_float4(1.f, 1.f, 1.f, 1.f);
        if ((MODE == 1 && g.out2) || (MODE == 2 && g.scale2)) sc2 = __ldg(reinterpret_cast<const float4 *>(g.scale2 + (long long)n * g.c + ch));
        const long long pix0 = (long long)n * g.out_h * g.out_w;          // first output pixel of the image
        const float *nz = (MODE == 1 && g.noise) ? g.noise + (long long)n * g.noise_bstride : nullptr;
        const float *smap = (MODE == 1 && g.stylemap) ? g.stylemap + (long long)n * g.map_bstride : nullptr;
        float4 dot = zero;
        float4 acc[K];
#pragma unroll
        for (int a = 0; a < K; ++a) acc[a] = zero;

        for (int k = 0; k < nstages; ++k) {
            // side inputs of the four output rows this stage retires, fetched BEFORE the wait for the stage: their latency
            // overlaps the barrier wait instead of sitting between the filter and the stores of every row
            float add_[FS_ROWS], m0_[FS_ROWS];
            float4 tt_[FS_ROWS];
#pragma unroll
            for (int u = 0; u < FS_ROWS; ++u) {
                const int oy = oy_start + k * FS_ROWS + u - (K - 1);
                const bool emit = (k > 0 || u == K - 1) && oy < oy_end && col_ok;
                const int opix = oy * g.out_w + ox;
                add_[u] = 0.0f; m0_[u] = 1.0f; tt_[u] = zero;
                if (MODE == 1 && emit) {                       // (per-thread loads only when the side boxes could not be built)
                    if (nz && !nz_tma) add_[u] = nw * __ldg(nz + opix);
                    if (smap && !map_tma) { m0_[u] = __ldg(smap + opix); add_[u] += __ldg(smap + map_plane + opix); }
                }
                if (MODE == 2 && emit && g.other) tt_[u] = __ldg(reinterpret_cast<const float4 *>(g.other + (pix0 + opix) * g.c + ch));
            }
            ft_mbar_wait(&full_bar[s], ph);
            const uint32_t stg = (uint32_t)__cvta_generic_to_shared(smem + s * FS_STAGE_BYTES) + (col * (FS_C / 4) + quad) * 16;
#pragma unroll
            for (int u = 0; u < FS_ROWS; ++u) {
                const int oy = oy_start + k * FS_ROWS + u - (K - 1);        // the output row this step completes
                const bool emit = (k > 0 || u == K - 1) && oy < oy_end && col_ok;
                const int opix = oy * g.out_w + ox;
                float add = add_[u], m0 = m0_[u];
                const float4 tt = tt_[u];
                if (MODE == 1) {
                    const uint32_t side = (uint32_t)__cvta_generic_to_shared(smem + s * FS_STAGE_BYTES + FS_X_BYTES) + (u * FS_W + col) * 4;
                    if (nz_tma) add = nw * lds_f1(side);
                    if (map_tma) { m0 = lds_f1(side + FS_NOISE_BYTES); add += lds_f1(side + FS_NOISE_BYTES + FS_ROWS * FS_W * 4); }
                }
                float4 cur[K];
#pragma unroll
                for (int b = 0; b < K; ++b) cur[b] = lds_f4(stg + (u * FS_IW + b) * (FS_C / 4) * 16);
                if (sep) {
                    float4 h = zero;
#pragma unroll
                    for (int b = 0; b < K; ++b) {
                        h.x = fmaf(cur[b].x, kh[b], h.x); h.y = fmaf(cur[b].y, kh[b], h.y);
                        h.z = fmaf(cur[b].z, kh[b], h.z); h.w = fmaf(cur[b].w, kh[b], h.w);
                    }
#pragma unroll
                    for (int a = 0; a < K; ++a) {              // input row r is tap row a of output row r - a
                        float4 &d = acc[(u - a) & 3];
                        d.x = fmaf(h.x, kv[a], d.x); d.y = fmaf(h.y, kv[a], d.y);
                        d.z = fmaf(h.z, kv[a], d.z); d.w = fmaf(h.w, kv[a], d.w);
                    }
                } else {
#pragma unroll
                    for (int a = 0; a < K; ++a)
#pragma unroll
                        for (int b = 0; b < K; ++b) {
                            float4 &d = acc[(u - a) & 3];
                            const float kk = __ldg(taps + (K - 1 - a) * K + (K - 1 - b));
                            d.x = fmaf(cur[b].x, kk, d.x); d.y = fmaf(cur[b].y, kk, d.y);
                            d.z = fmaf(cur[b].z, kk, d.z); d.w = fmaf(cur[b].w, kk, d.w);
                        }
                }
                if (emit) {
                    float4 o = acc[(u + 1) & 3];
                    const float4 pre = o;
                    float *op = out + (pix0 + opix) * g.c + ch;
                    if (MODE == 1) {
                        float v;
                        v = fmaf(o.x, m0, add + bias4.x); o.x = ((v > 0.f) ? v : v * g.alpha) * g.gain;
                        v = fmaf(o.y, m0, add + bias4.y); o.y = ((v > 0.f) ? v : v * g.alpha) * g.gain;
                        v = fmaf(o.z, m0, add + bias4.z); o.z = ((v > 0.f) ? v : v * g.alpha) * g.gain;
                        v = fmaf(o.w, m0, add + bias4.w); o.w = ((v > 0.f) ? v : v * g.alpha) * g.gain;
                        *reinterpret_cast<float4 *>(op) = smap ? pre : o;
                        if (g.out2) {
                            const float a0 = o.x * sc2.x, a1 = o.y * sc2.y, a2 = o.z * sc2.z, a3 = o.w * sc2.w;
                            if (g.op16) reinterpret_cast<uint2 *>(g.out2)[((pix0 + opix) * g.c + ch) >> 2] = pack4_bf16(a0, a1, a2, a3);
                            else *reinterpret_cast<float4 *>(g.out2 + (pix0 + opix) * g.c + ch) =
                                     make_float4(round_tf32_(a0), round_tf32_(a1), round_tf32_(a2), round_tf32_(a3));
                        }
                    } else if (MODE == 2) {
                        dot.x = fmaf(o.x, tt.x, dot.x); dot.y = fmaf(o.y, tt.y, dot.y);
                        dot.z = fmaf(o.z, tt.z, dot.z); dot.w = fmaf(o.w, tt.w, dot.w);
                        o.x *= sc2.x; o.y *= sc2.y; o.z *= sc2.z; o.w *= sc2.w;
                        if (g.op16) reinterpret_cast<uint2 *>(out)[((pix0 + opix) * g.c + ch) >> 2] = pack4_bf16(o.x, o.y, o.z, o.w);
                        else *reinterpret_cast<float4 *>(op) =
                                 make_float4(round_tf32_(o.x), round_tf32_(o.y), round_tf32_(o.z), round_tf32_(o.w));
                    } else {
                        *reinterpret_cast<float4 *>(op) = o;
                    }
                }
                acc[(u + 1) & 3] = zero;                       // the slot now belongs to the output row three steps ahead
            }
            __syncwarp();                                      // this warp no longer reads the stage
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"((uint32_t)__cvta_generic_to_shared(&empty_bar[s])) : "memory");
            if (++s == FS_STAGES) { s = 0; ph ^= 1; }
        }
        if (MODE == 2 && g.dot) {   // lanes with equal channel quad (l, l^8, l^16, l^24) -> one 128-bit reduction per quad and warp
#pragma unroll
            for (int o = 8; o < 32; o <<= 1) {
                dot.x += __shfl_xor_sync(0xffffffffu, dot.x, o); dot.y += __shfl_xor_sync(0xffffffffu, dot.y, o);
                dot.z += __shfl_xor_sync(0xffffffffu, dot.z, o); dot.w += __shfl_xor_sync(0xffffffffu, dot.w, o);
            }
            if (lane < 8) {
                float *dp = g.dot + (long long)n * g.c + ch;
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(dp), "f"(dot.x), "f"(dot.y), "f"(dot.z), "f"(dot.w) : "memory");
            }
        }
    }
}

// ---- planes (NCHW, minor = 1) 4x4 FIR, row-streaming through a bulk-copy ring (round 2) ---------------------------------
// The reference layout [major = N*C, H, W] of `op.upfirdn2d`.  Rows of a (2H+1)-wide plane are not 16-byte aligned, so
// TMA tiles (16-byte strides) and vector loads do not apply; but FULL rows of a plane are contiguous in memory.  A stage
// of the ring is therefore a 1-D bulk copy (cp.async.bulk) of 4 consecutive input rows: the source is aligned down to 16
// bytes and the size rounded up (the over-fetch stays inside the tensor: base 16-byte aligned, element count % 4 == 0),
// consumers address the stage with the 0..3-float lead of that chunk.  Work item = (plane, ~128-row segment), persistent
// CTAs; consumer lane l of warp w owns output columns 128 w + l + 32 j (j = 0..3): every LDS / STG of a warp touches 32
// consecutive floats (conflict free, whole 128-byte lines); 4-tap row filter, then one FMA per open output row (register
// ring, compile-time slots).  Rows outside the plane are skipped (zero), edge columns take predicated loads.
constexpr int FP_ROWS = 4;

struct FirPlaneGeom {
    int in_h, in_w, out_h, out_w, pad_x0, pad_y0, nseg, seg_rows, total_items, stages, stage_bytes;
    int debug;                    // profiling only (SR_FIR_PLANES_DEBUG): bit 0 = no stores, bit 1 = no filter arithmetic
    FastDiv div_seg;
};

// J = output columns per consumer lane (32 apart).  Measured: the consumers are LATENCY bound (a serial row loop with short
// dependent chains), so what counts is consumer warps per SM: J = 1 gives a warp per 32 columns (9 warps for a 257-wide plane).
template <int J>
__global__ void __launch_bounds__(32 * (16 / J + 1))
fir_planes_stream_kernel(float *__restrict__ out, const float *__restrict__ x, const float *__restrict__ taps, const FirPlaneGeom g)
{
    extern __shared__ uint8_t fp_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(fp_raw) + 127) & ~(uintptr_t)127);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + g.stages * g.stage_bytes);
    uint64_t *empty_bar = full_bar + 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int consumer_warps = (blockDim.x >> 5) - 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < g.stages; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(&full_bar[s])), "r"(1u) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(&empty_bar[s])), "r"((uint32_t)consumer_warps) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int64_t plane_in = (int64_t)g.in_h * g.in_w;

    if (warp == consumer_warps) {                        // ===== producer warp (one lane)
        if (lane == 0) {
            uint32_t s = 0, ph = 0;
            for (uint32_t T = blockIdx.x; T < (uint32_t)g.total_items; T += gridDim.x) {
                uint32_t m, seg;
                g.div_seg.divmod(T, m, seg);
                const int oy_start = seg * g.seg_rows;
                const int oy_end = min(g.out_h, oy_start + g.seg_rows);
                const int nstages = (oy_end - oy_start + 3 + FP_ROWS - 1) / FP_ROWS;
                int iy0 = oy_start - g.pad_y0;
                for (int k = 0; k < nstages; ++k, iy0 += FP_ROWS) {
                    ft_mbar_wait(&empty_bar[s], ph ^ 1);
                    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&full_bar[s]);
                    const int r0 = max(iy0, 0), r1 = min(iy0 + FP_ROWS, g.in_h);
                    if (r1 > r0) {
                        const int64_t idx = (int64_t)m * plane_in + (int64_t)r0 * g.in_w;
                        const int lead = (int)(idx & 3);
                        const uint32_t bytes = (uint32_t)(((lead + (r1 - r0) * g.in_w) * 4 + 15) & ~15);
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                     :: "r"((uint32_t)__cvta_generic_to_shared(smem + s * g.stage_bytes)), "l"(x + (idx - lead)), "r"(bytes), "r"(bar) : "memory");
                    } else {
                        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
                    }
                    if (++s == (uint32_t)g.stages) { s = 0; ph ^= 1; }
                }
            }
        }
        return;
    }

    // ===== consumers
    constexpr int K = 4;
    float kv[K], kh[K];
    const bool sep = fir_rank1_taps(taps, kv, kh);
    const int col0 = warp * (32 * J) + lane;              // output columns col0 + 32 j
    bool live[J], edge[J];
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const int ox = col0 + 32 * j, ix = ox - g.pad_x0;
        live[j] = ox < g.out_w;
        edge[j] = ix < 0 || ix + K - 1 >= g.in_w;
    }
    uint32_t s = 0, ph = 0;
    for (uint32_t T = blockIdx.x; T < (uint32_t)g.total_items; T += gridDim.x) {
        uint32_t m, seg;
        g.div_seg.divmod(T, m, seg);
        const int oy_start = seg * g.seg_rows;
        const int oy_end = min(g.out_h, oy_start + g.seg_rows);
        const int nstages = (oy_end - oy_start + 3 + FP_ROWS - 1) / FP_ROWS;
        const int lead_plane = (int)(((int64_t)m * plane_in) & 3);
        float *orow = out + ((int64_t)m * g.out_h + oy_start - (K - 1)) * g.out_w + col0;    // row of the output the next step retires
        float acc[K][J];
#pragma unroll
        for (int a = 0; a < K; ++a)
#pragma unroll
            for (int j = 0; j < J; ++j) acc[a][j] = 0.0f;
        int iy0 = oy_start - g.pad_y0;
        for (int k = 0; k < nstages; ++k, iy0 += FP_ROWS) {
            ft_mbar_wait(&full_bar[s], ph);
            const int r0 = max(iy0, 0);
            const int lead = (lead_plane + r0 * g.in_w) & 3;
            const float *stg = reinterpret_cast<const float *>(smem + s * g.stage_bytes) + lead + col0 - g.pad_x0;
#pragma unroll
            for (int u = 0; u < FP_ROWS; ++u) {
                const int iy = iy0 + u;
                if (iy >= 0 && iy < g.in_h) {             // rows outside the plane contribute nothing
                    const float *row = stg + (iy - r0) * g.in_w;
#pragma unroll
                    for (int j = 0; j < J; ++j) {
                        float c[K];
                        if (!edge[j]) {
#pragma unroll
                            for (int b = 0; b < K; ++b) c[b] = row[32 * j + b];
                        } else {
                            const int ix = col0 + 32 * j - g.pad_x0;
#pragma unroll
                            for (int b = 0; b < K; ++b) c[b] = (live[j] && ix + b >= 0 && ix + b < g.in_w) ? row[32 * j + b] : 0.0f;
                        }
                        if (sep) {
                            float h = 0.0f;
#pragma unroll
                            for (int b = 0; b < K; ++b) h = fmaf(c[b], kh[b], h);
#pragma unroll
                            for (int a = 0; a < K; ++a) acc[(u - a) & 3][j] = fmaf(h, kv[a], acc[(u - a) & 3][j]);
                        } else {
#pragma unroll
                            for (int a = 0; a < K; ++a)
#pragma unroll
                                for (int b = 0; b < K; ++b)
                                    acc[(u - a) & 3][j] = fmaf(c[b], __ldg(taps + (K - 1 - a) * K + (K - 1 - b)), acc[(u - a) & 3][j]);
                        }
                    }
                }
                const int oy = oy_start + k * FP_ROWS + u - (K - 1);
                if ((k > 0 || u == K - 1) && oy < oy_end) {
#pragma unroll
                    for (int j = 0; j < J; ++j)
                        if (live[j]) orow[32 * j] = acc[(u + 1) & 3][j];
                }
#pragma unroll
                for (int j = 0; j < J; ++j) acc[(u + 1) & 3][j] = 0.0f;
                orow += g.out_w;
            }
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"((uint32_t)__cvta_generic_to_shared(&empty_bar[s])) : "memory");
            if (++s == (uint32_t)g.stages) { s = 0; ph ^= 1; }
        }
    }
}

// ---- planes FIR, vectorised consumers (round 2, second form) -------------------------------------------------------------
// fir_planes_stream_kernel above is ISSUE bound: 4-byte LDS / STG per lane, ~25 instructions per output.  Same ring, but a
// consumer thread owns FOUR ADJACENT output columns: the 7 inputs of a row come from three ALIGNED LDS.128 (12 floats) and a
// compile-time selection by the row's misalignment a = (chunk lead + row offset - pad) mod 4, which is uniform over the CTA
// (a 4-way switch around the row body); aligned output rows leave as STG.128, unaligned ones (odd widths) as four scalar
// stores.  The <= 3 columns beyond the last full quad are a scalar side job of warp 0.  ~12 instructions per output.
// rank-1 taps (every FIR the model builds): 4-tap row filter, then one FMA per open output row
template <int A, int U>
__device__ __forceinline__ void fpv_row(uint32_t q4, const float (&kh)[4], const float (&kv)[4], float (&acc)[4][4])
{
    const float4 v0 = lds_f4(q4), v1 = lds_f4(q4 + 16), v2 = lds_f4(q4 + 32);
    const float f[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float h = 0.0f;
#pragma unroll
        for (int b = 0; b < 4; ++b) h = fmaf(f[A + j + b], kh[b], h);
#pragma unroll
        for (int a = 0; a < 4; ++a) acc[(U - a) & 3][j] = fmaf(h, kv[a], acc[(U - a) & 3][j]);
    }
}

// general (rank > 1) taps: rare; a compact rolled loop over the tap column with scalar shared-memory reads, so that the hot
// loop stays small (ncu: the first version of this kernel, both forms inlined 16 times, stalled on instruction fetch)
template <int U>
__device__ __forceinline__ void fpv_row_general(uint32_t srow, const float *__restrict__ taps, float (&acc)[4][4])
{
#pragma unroll 1
    for (int b = 0; b < 4; ++b) {
        const float c0 = lds_f1(srow + 4 * b), c1 = lds_f1(srow + 4 * b + 4), c2 = lds_f1(srow + 4 * b + 8), c3 = lds_f1(srow + 4 * b + 12);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const float kk = __ldg(taps + (3 - a) * 4 + (3 - b));
            acc[(U - a) & 3][0] = fmaf(c0, kk, acc[(U - a) & 3][0]); acc[(U - a) & 3][1] = fmaf(c1, kk, acc[(U - a) & 3][1]);
            acc[(U - a) & 3][2] = fmaf(c2, kk, acc[(U - a) & 3][2]); acc[(U - a) & 3][3] = fmaf(c3, kk, acc[(U - a) & 3][3]);
        }
    }
}

// Interior quads [e0, e1) -- all 7 inputs of every row inside the plane -- run the vector path with no masking at all; the
// columns of the edge quads and the <= 3 columns past the last full quad (at most 32 in total) are a scalar side job of the
// lanes of warp 0 with predicated loads.
__global__ void __launch_bounds__(32 * 5)
fir_planes_vec_kernel(float *__restrict__ out, const float *__restrict__ x, const float *__restrict__ taps, const FirPlaneGeom g,
                      const int e0, const int e1)
{
    extern __shared__ uint8_t fp_raw[];
    // 128 bytes of slack in front of stage 0: an aligned quad may start up to 12 bytes before a chunk
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(fp_raw) + 127) & ~(uintptr_t)127) + 128;
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + g.stages * g.stage_bytes);
    uint64_t *empty_bar = full_bar + 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int consumer_warps = (blockDim.x >> 5) - 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < g.stages; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(&full_bar[s])), "r"(1u) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(&empty_bar[s])), "r"((uint32_t)consumer_warps) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int64_t plane_in = (int64_t)g.in_h * g.in_w;

    if (warp == consumer_warps) {                        // ===== producer warp (one lane): as fir_planes_stream_kernel
        if (lane == 0) {
            uint32_t s = 0, ph = 0;
            for (uint32_t T = blockIdx.x; T < (uint32_t)g.total_items; T += gridDim.x) {
                uint32_t m, seg;
                g.div_seg.divmod(T, m, seg);
                const int oy_start = seg * g.seg_rows;
                const int oy_end = min(g.out_h, oy_start + g.seg_rows);
                const int nstages = (oy_end - oy_start + 3 + FP_ROWS - 1) / FP_ROWS;
                int iy0 = oy_start - g.pad_y0;
                for (int k = 0; k < nstages; ++k, iy0 += FP_ROWS) {
                    ft_mbar_wait(&empty_bar[s], ph ^ 1);
                    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&full_bar[s]);
                    const int r0 = max(iy0, 0), r1 = min(iy0 + FP_ROWS, g.in_h);
                    if (r1 > r0) {
                        const int64_t idx = (int64_t)m * plane_in + (int64_t)r0 * g.in_w;
                        const int lead = (int)(idx & 3);
                        const uint32_t bytes = (uint32_t)(((lead + (r1 - r0) * g.in_w) * 4 + 15) & ~15);
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                     :: "r"((uint32_t)__cvta_generic_to_shared(smem + s * g.stage_bytes)), "l"(x + (idx - lead)), "r"(bytes), "r"(bar) : "memory");
                    } else {
                        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
                    }
                    if (++s == (uint32_t)g.stages) { s = 0; ph ^= 1; }
                }
            }
        }
        return;
    }

    // ===== consumers: thread t owns the interior quad t (output columns 4t .. 4t+3, e0 <= t < e1); lane l of warp 0 also owns
    // side column l (l < 4 e0) or 4 e1 + l - 4 e0
    float kv[4], kh[4];
    const bool sep = fir_rank1_taps(taps, kv, kh);
    const int t = threadIdx.x;
    const bool quad_live = t >= e0 && t < e1;
    const int sx = lane < 4 * e0 ? lane : 4 * e1 + lane - 4 * e0;              // side column of this lane
    const bool side_live = warp == 0 && sx < g.out_w;
    const int six = sx - g.pad_x0;                                             // its first input column
    const bool out_vec = (g.out_w & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0;
    uint32_t s = 0, ph = 0;
    for (uint32_t T = blockIdx.x; T < (uint32_t)g.total_items; T += gridDim.x) {
        uint32_t m, seg;
        g.div_seg.divmod(T, m, seg);
        const int oy_start = seg * g.seg_rows;
        const int oy_end = min(g.out_h, oy_start + g.seg_rows);
        const int nstages = (oy_end - oy_start + 3 + FP_ROWS - 1) / FP_ROWS;
        const int lead_plane = (int)(((int64_t)m * plane_in) & 3);
        float *orow = out + ((int64_t)m * g.out_h + oy_start - 3) * g.out_w;       // output row the next step retires
        float acc[4][4], tacc[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            tacc[a] = 0.0f;
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[a][j] = 0.0f;
        }
        int iy0 = oy_start - g.pad_y0;
        for (int k = 0; k < nstages; ++k, iy0 += FP_ROWS) {
            ft_mbar_wait(&full_bar[s], ph);
            const int r0 = max(iy0, 0);
            const uint32_t stage = (uint32_t)__cvta_generic_to_shared(smem + s * g.stage_bytes);
            // stage float index of the first tap of output column 0 in input row iy0 (rows advance by in_w)
            int base = ((lead_plane + r0 * g.in_w) & 3) + (iy0 - r0) * g.in_w - g.pad_x0;
            // one step per input row; U = ring phase (compile time): the row is tap a of the output row in slot (U - a) & 3,
            // the output row in slot (U + 1) & 3 is complete after it
#define FPV_STEP(U)                                                                                                      \
            {                                                                                                            \
                const int iy = iy0 + U;                                                                                  \
                if (iy >= 0 && iy < g.in_h && !(g.debug & 2)) {                                                          \
                    const int a = base & 3;                                  /* two's complement: right for base < 0 too */ \
                    if (quad_live) {                                                                                     \
                        if (sep) {                                                                                       \
                            const uint32_t q4 = stage + (base - a) * 4 + t * 16;                                         \
                            switch (a) {                                                                                 \
                            case 0: fpv_row<0, U>(q4, kh, kv, acc); break;                                               \
                            case 1: fpv_row<1, U>(q4, kh, kv, acc); break;                                               \
                            case 2: fpv_row<2, U>(q4, kh, kv, acc); break;                                               \
                            default: fpv_row<3, U>(q4, kh, kv, acc); break;                                              \
                            }                                                                                            \
                        } else {                                                                                         \
                            fpv_row_general<U>(stage + (base + 4 * t) * 4, taps, acc);                                   \
                        }                                                                                                \
                    }                                                                                                    \
                    if (side_live) {                                                                                     \
                        const uint32_t row = stage + (base + sx) * 4;                                                    \
                        float c[4];                                                                                      \
                        _Pragma("unroll") for (int b = 0; b < 4; ++b) c[b] = (six + b >= 0 && six + b < g.in_w) ? lds_f1(row + 4 * b) : 0.0f; \
                        if (sep) {                                                                                       \
                            float h = 0.0f;                                                                              \
                            _Pragma("unroll") for (int b = 0; b < 4; ++b) h = fmaf(c[b], kh[b], h);                      \
                            _Pragma("unroll") for (int a2 = 0; a2 < 4; ++a2) tacc[(U - a2) & 3] = fmaf(h, kv[a2], tacc[(U - a2) & 3]); \
                        } else {                                                                                         \
                            _Pragma("unroll") for (int a2 = 0; a2 < 4; ++a2)                                             \
                                _Pragma("unroll") for (int b = 0; b < 4; ++b)                                            \
                                    tacc[(U - a2) & 3] = fmaf(c[b], __ldg(taps + (3 - a2) * 4 + (3 - b)), tacc[(U - a2) & 3]); \
                        }                                                                                                \
                    }                                                                                                    \
                }                                                                                                        \
                base += g.in_w;                                                                                          \
                const int oy = oy_start + k * FP_ROWS + U - 3;                                                           \
                if ((k > 0 || U == 3) && oy < oy_end && !(g.debug & 1)) {                                                \
                    constexpr int D = (U + 1) & 3;                                                                       \
                    if (quad_live) {                                                                                     \
                        float *o = orow + 4 * t;                                                                         \
                        if (out_vec) *reinterpret_cast<float4 *>(o) = make_float4(acc[D][0], acc[D][1], acc[D][2], acc[D][3]); \
                        else { o[0] = acc[D][0]; o[1] = acc[D][1]; o[2] = acc[D][2]; o[3] = acc[D][3]; }                 \
                    }                                                                                                    \
                    if (side_live) orow[sx] = tacc[D];                                                                   \
                }                                                                                                        \
                _Pragma("unroll") for (int j = 0; j < 4; ++j) acc[(U + 1) & 3][j] = 0.0f;                                \
                tacc[(U + 1) & 3] = 0.0f;                                                                                \
                orow += g.out_w;                                                                                         \
            }
            FPV_STEP(0) FPV_STEP(1) FPV_STEP(2) FPV_STEP(3)
#undef FPV_STEP
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"((uint32_t)__cvta_generic_to_shared(&empty_bar[s])) : "memory");
            if (++s == (uint32_t)g.stages) { s = 0; ph ^= 1; }
        }
    }
}

inline bool vec_env_early() { const char *v = getenv("SR_FIR_PLANES_VEC"); return v && v[0] == '1'; }   // 1 = force the vector kernel

// Streaming path of the planes blur: SR_ERR_UNSUPPORTED when the shape does not qualify (the tile kernels take over).
int launch_planes_stream(float *out, const float *x, const float *taps, int64_t major, int in_h, int in_w, int oh, int ow,
                         int pad_x0, int pad_y0, cudaStream_t st)
{
    const char *off = getenv("SR_FIR_PLANES_STREAM");           // A/B switch, read per call: 0 = the tile / strip kernels
    if (off && off[0] == '0') return SR_ERR_UNSUPPORTED;
    if (ow < 64 || ow > 512 || oh < 16 || in_w < 8 || in_w > 1024) return SR_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(x) & 15u) || (major * (int64_t)in_h * in_w) % 4 != 0) return SR_ERR_UNSUPPORTED;
    FirPlaneGeom g;
    g.in_h = in_h; g.in_w = in_w; g.out_h = oh; g.out_w = ow; g.pad_x0 = pad_x0; g.pad_y0 = pad_y0;
    g.nseg = (oh + 64) / 128 > 0 ? (oh + 64) / 128 : 1;
    g.seg_rows = (oh + g.nseg - 1) / g.nseg;
    const int64_t total = major * g.nseg;
    if (total >= 0x7fffffffll) return SR_ERR_UNSUPPORTED;
    g.total_items = (int)total;
    g.div_seg = FastDiv((uint32_t)g.nseg);
    g.stage_bytes = ((FP_ROWS * in_w * 4 + 16 + 127) / 128) * 128;
    { const char *d = getenv("SR_FIR_PLANES_DEBUG"); g.debug = d ? atoi(d) : 0; }
    const char *st_env = getenv("SR_FIR_PLANES_STAGES");          // experiment switches, read per call
    const char *j_env = getenv("SR_FIR_PLANES_J");
    g.stages = st_env ? atoi(st_env) : 8;
    if (g.stages > 16) g.stages = 16;
    if (g.stages < 2) g.stages = 2;
    // measured (scratch/planes_sweep.py, B200; fraction of the copy bandwidth, tile kernels -> this kernel):
    //   257^2 -> 256^2: 0.41 -> 0.54 (J = 2 or 4)   256^2 -> 257^2: 0.32 -> 0.45 (J = 4)   129^2 -> 128^2: 0.40 -> 0.46 (J = 2)
    //   128^2 -> 129^2: 0.26 -> 0.31 (J = 2)        64^2 -> 65^2: 0.14 -> 0.22 (J = 2)     65^2 -> 64^2: 0.36 (tile kernel) vs 0.23
    // The kernel is ISSUE bound (4-byte LDS / STG per lane: ~25 instructions per output); J = 1 (more warps) is slowest.
    if (!j_env && !vec_env_early() && ow < 100 && in_w % 4 != 0) return SR_ERR_UNSUPPORTED;     // the tile kernel is faster there
    const int J = j_env ? atoi(j_env) : (ow >= 200 ? 4 : 2);
    if (J != 1 && J != 2 && J != 4) return SR_ERR_UNSUPPORTED;
    const int consumer_warps = (ow + 32 * J - 1) / (32 * J);
    if (consumer_warps > 16 / J) return SR_ERR_UNSUPPORTED;
    const int threads = 32 * (consumer_warps + 1);
    const size_t smem = 128 + (size_t)g.stages * g.stage_bytes + 32 * sizeof(uint64_t);
    if (smem > 100 * 1024) return SR_ERR_UNSUPPORTED;
    int per_sm = (int)(220 * 1024 / (smem + 1024));
    if (per_sm > 2048 / threads) per_sm = 2048 / threads;
    if (per_sm > 16) per_sm = 16;
    if (per_sm < 1) return SR_ERR_UNSUPPORTED;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(fir_planes_stream_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess ||
            cudaFuncSetAttribute(fir_planes_stream_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess ||
            cudaFuncSetAttribute(fir_planes_stream_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess)
            return SR_ERR_UNSUPPORTED;
        configured = true;
    }
    const int64_t slots = (int64_t)per_sm * kNumSMs;
    const int grid = (int)(total < slots ? total : slots);
    const char *vec_env = getenv("SR_FIR_PLANES_VEC");            // A/B switch: 0 = the scalar-lane kernel above
    // measured (scratch/planes_sweep.py, fraction of the copy bandwidth; tile kernels -> scalar-lane ring -> this kernel):
    //   257^2 -> 256^2: 0.41 -> 0.52 -> 0.65    256^2 -> 257^2: 0.32 -> 0.45 -> 0.63    513^2 -> 512^2: 0.43 -> 0.56 -> 0.66
    //   129^2 -> 128^2: 0.40 -> 0.46 -> 0.39    128^2 -> 129^2: 0.26 -> 0.31 -> 0.38     64^2 -> 65^2: 0.14 -> 0.22 -> 0.19
    // 4 stages beat 8 / 12 everywhere (more CTAs per SM); the bare ring without consumer work streams reads at 5.1 TB/s.
    const bool vec_shape = ow >= 200 || (in_w % 4 == 0 && ow >= 100);
    if (!(vec_env && vec_env[0] == '0') && !j_env && (vec_shape || (vec_env && vec_env[0] == '1'))) {
        if (!st_env) g.stages = 4;
        const int nq = ow / 4, cw = nq > 0 ? (nq + 31) / 32 : 1;
        // interior quads [e0, e1): 4t - pad >= 0 and 4t + 6 - pad < in_w; everything else (<= 32 columns) is the side job
        int e0 = pad_x0 > 0 ? (pad_x0 + 3) / 4 : 0, e1 = (in_w + pad_x0 - 7 >= 0) ? (in_w + pad_x0 - 7) / 4 + 1 : 0;
        if (e1 > nq) e1 = nq;
        if (e0 > e1) e0 = e1;
        if (cw <= 4 && 4 * e0 + (ow - 4 * e1) <= 32) {
            const int vthreads = 32 * (cw + 1);
            const size_t vsmem = 256 + (size_t)g.stages * g.stage_bytes + 32 * sizeof(uint64_t);
            int vper_sm = (int)(220 * 1024 / (vsmem + 1024));
            if (vper_sm > 2048 / vthreads) vper_sm = 2048 / vthreads;
            if (vper_sm > 16) vper_sm = 16;
            static bool vconf = false;
            if (!vconf) {
                if (cudaFuncSetAttribute(fir_planes_vec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess) return SR_ERR_UNSUPPORTED;
                vconf = true;
            }
            const int64_t vslots = (int64_t)vper_sm * kNumSMs;
            fir_planes_vec_kernel<<<(int)(total < vslots ? total : vslots), vthreads, vsmem, st>>>(out, x, taps, g, e0, e1);
            return SR_OK;
        }
    }
    if (J == 1) fir_planes_stream_kernel<1><<<grid, threads, smem, st>>>(out, x, taps, g);
    else if (J == 2) fir_planes_stream_kernel<2><<<grid, threads, smem, st>>>(out, x, taps, g);
    else fir_planes_stream_kernel<4><<<grid, threads, smem, st>>>(out, x, taps, g);
    return SR_OK;
}

// ---- generic: one thread per output element, any geometry ---------------------------------------
struct GenericGeom {
    int64_t major, minor, total;
    int in_h, in_w, out_h, out_w, kh, kw, up_x, up_y, down_x, down_y, pad_x0, pad_y0;
};

__global__ void __launch_bounds__(kThreads)
upfirdn2d_generic_kernel(float *__restrict__ out, const float *__restrict__ x, const float *__restrict__ taps,
                         const GenericGeom g)
{
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t idx = (int64_t)blockIdx.x * kThreads + threadIdx.x; idx < g.total; idx += stride) {
        int64_t r = idx;
        const int c = (int)(r % g.minor); r /= g.minor;
        const int ox = (int)(r % g.out_w); r /= g.out_w;
        const int oy = (int)(r % g.out_h);
        const int64_t m = r / g.out_h;
        // poly-phase walk in the same (ky, kx) row-major order as the oracle
        const int ky0 = pos_mod_i(g.pad_y0 - oy * g.down_y, g.up_y);
        const int kx0 = pos_mod_i(g.pad_x0 - ox * g.down_x, g.up_x);
        float acc = 0.0f;
        for (int ky = ky0; ky < g.kh; ky += g.up_y) {
            const int iy = (oy * g.down_y + ky - g.pad_y0) / g.up_y;       // exact division
            if (iy < 0 || iy >= g.in_h) continue;
            for (int kx = kx0; kx < g.kw; kx += g.up_x) {
                const int ix = (ox * g.down_x + kx - g.pad_x0) / g.up_x;
                if (ix < 0 || ix >= g.in_w) continue;
                acc = fmaf(__ldg(x + ((m * g.in_h + iy) * g.in_w + ix) * g.minor + c),
                           __ldg(taps + (g.kh - 1 - ky) * g.kw + (g.kw - 1 - kx)), acc);
            }
        }
        out[idx] = acc;
    }
}

int ilog2_ceil(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

template <int UP, int DOWN, int KH, int KW, int PHX, int PHY, int VY>
int launch_tile(float *out, const float *x, const float *taps, int64_t major, int in_h, int in_w,
                int out_h, int out_w, int pad_x0, int pad_y0, cudaStream_t st)
{
    constexpr int JX = KW / UP, JY = KH / UP;
    constexpr int NWX = in_off(VX - 1, UP, DOWN, PHX) + JX, NWY = in_off(VY - 1, UP, DOWN, PHY) + JY;
    constexpr int SX = VX * DOWN / UP, SY = VY * DOWN / UP;
    constexpr int LDW = (SX % 4 == 0) ? 4 : ((SX % 2 == 0) ? 2 : 1);
    constexpr int NWXP = (NWX + LDW - 1) / LDW * LDW;
    TileGeom g;
    g.in_h = in_h; g.in_w = in_w; g.out_h = out_h; g.out_w = out_w; g.major = major;
    g.pqx = floor_div_i(pad_x0, UP); g.pqy = floor_div_i(pad_y0, UP);
    int tx = ilog2_ceil((out_w + VX - 1) / VX);                             // <= 16 strips = 64 outputs wide
    if (tx > 4) {            // wide planes: 64, 32 or 16 columns per tile, whichever wastes the fewest columns
        int best = 4, best_w = (out_w + 63) / 64 * 64;
        for (int c = 3; c >= 2; --c) {
            const int tw = VX << c, padded = (out_w + tw - 1) / tw * tw;
            if (padded * 100 < best_w * 70) { best = c; best_w = padded; }   // measured: narrow tiles only pay when > 30% is saved
        }
        tx = best;
    }
    int ty = ilog2_ceil((out_h + VY - 1) / VY); if (ty > 8 - tx) ty = 8 - tx;
    g.txs_log2 = tx; g.tys_log2 = ty;
    const int txs = 1 << tx, tys = 1 << ty, ppt = kThreads >> (tx + ty);
    g.tin_w = (txs - 1) * SX + NWX;
    g.tin_h = (tys - 1) * SY + NWY;
    g.tin_stride = ((txs - 1) * SX + NWXP + 3) / 4 * 4;
    g.tiles_x = (out_w + txs * VX - 1) / (txs * VX);
    g.tiles_y = (out_h + tys * VY - 1) / (tys * VY);
    g.div_tiles_x = FastDiv((uint32_t)g.tiles_x);
    g.div_tiles_y = FastDiv((uint32_t)g.tiles_y);
    g.vec_store = (out_w % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15u) == 0);
    const int64_t groups = (major + ppt - 1) / ppt;
    const int64_t blocks = groups * g.tiles_x * g.tiles_y;
    if (blocks >= 0x7fffffffll) return SR_ERR_UNSUPPORTED;
    // +4 floats of slack: the widest vector read of the last row may run past the staged columns
    const size_t smem = sizeof(float) * ((size_t)ppt * g.tin_h * g.tin_stride + 4);
    auto kern = upfirdn2d_tile_kernel<UP, DOWN, KH, KW, PHX, PHY, VY>;
    if (smem > 48 * 1024) {
        if (smem > 200 * 1024) return SR_ERR_UNSUPPORTED;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    kern<<<(unsigned)blocks, kThreads, smem, st>>>(out, x, taps, g);
    return SR_OK;
}

// Streaming path of launch_nhwc (fir_nhwc_stream_kernel): returns SR_ERR_UNSUPPORTED when the shape does not qualify.
int launch_nhwc_stream(float *out, const float *x, const float *taps, int64_t major, int in_h, int in_w, int oh, int ow,
                       int64_t minor, int pad_x0, int pad_y0, int mode, const float *noise, long long noise_bstride,
                       const float *noise_weight, const float *bias, float alpha, float gain, cudaStream_t st, float *out2,
                       const float *scale2, const float *other, float *dot, const float *stylemap, long long map_bstride, int op16)
{
    const char *off = getenv("SR_FIR_STREAM");                 // A/B switch, read per call: 0 = the register-window kernels
    if (off && off[0] == '0') return SR_ERR_UNSUPPORTED;
    if (minor % FS_C != 0 || oh < 32 || ow < 32 || (reinterpret_cast<uintptr_t>(x) & 15u)) return SR_ERR_UNSUPPORTED;
    if ((int64_t)major * oh * ow * minor >= (1ll << 40)) return SR_ERR_UNSUPPORTED;
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return SR_ERR_UNSUPPORTED;
    CUtensorMap tm;
    cuuint64_t dims[4] = {(cuuint64_t)minor, (cuuint64_t)in_w, (cuuint64_t)in_h, (cuuint64_t)major};
    cuuint64_t strides[3] = {(cuuint64_t)minor * 4, (cuuint64_t)in_w * minor * 4, (cuuint64_t)in_h * in_w * minor * 4};
    cuuint32_t box[4] = {FS_C, FS_IW, FS_ROWS, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return SR_ERR_UNSUPPORTED;
    FirStreamGeom g;
    g.in_h = in_h; g.in_w = in_w; g.out_h = oh; g.out_w = ow; g.c = (int)minor; g.cblocks = (int)(minor / FS_C);
    g.xtiles = (ow + FS_W - 1) / FS_W;
    g.nseg = (oh + 32) / 64 > 0 ? (oh + 32) / 64 : 1;           // segments of ~64 output rows (3 halo rows each)
    g.seg_rows = (oh + g.nseg - 1) / g.nseg;
    const int64_t total = (int64_t)major * g.nseg * g.xtiles * g.cblocks;
    if (total >= 0x7fffffffll) return SR_ERR_UNSUPPORTED;
    g.total_items = (int)total; g.pad_x0 = pad_x0; g.pad_y0 = pad_y0;
    g.div_cb = FastDiv((uint32_t)g.cblocks); g.div_xt = FastDiv((uint32_t)g.xtiles); g.div_seg = FastDiv((uint32_t)g.nseg);
    g.noise = noise; g.noise_weight = noise_weight; g.bias = bias; g.noise_bstride = noise_bstride;
    g.alpha = alpha; g.gain = gain; g.out2 = out2; g.scale2 = scale2; g.other = other; g.dot = dot;
    g.stylemap = stylemap; g.map_bstride = map_bstride; g.op16 = op16;
    // side inputs of the styled tail through the ring: [b, oh, ow] noise rows and [b, 2, oh, ow] style-map rows as TMA boxes
    // of 4 rows x 32 columns (rows must be 16-byte multiples; otherwise the consumers load them per thread)
    CUtensorMap tm_nz = tm, tm_map = tm;
    g.side_tma = 0;
    if (mode == 1 && noise && ow % 4 == 0 && (reinterpret_cast<uintptr_t>(noise) & 15u) == 0 && noise_bstride % 4 == 0) {
        const cuuint64_t nb = noise_bstride ? (cuuint64_t)major : 1;
        cuuint64_t d3[3] = {(cuuint64_t)ow, (cuuint64_t)oh, nb};
        cuuint64_t s3[2] = {(cuuint64_t)ow * 4, (cuuint64_t)(noise_bstride ? noise_bstride : (long long)oh * ow) * 4};
        cuuint32_t b3[3] = {FS_W, FS_ROWS, 1}, e3[3] = {1, 1, 1};
        if (enc(&tm_nz, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(noise), d3, s3, b3, e3, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
            g.side_tma |= 1;
    }
    if (mode == 1 && stylemap && ow % 4 == 0 && (reinterpret_cast<uintptr_t>(stylemap) & 15u) == 0 && map_bstride % 4 == 0 &&
        map_bstride >= 2ll * oh * ow) {
        cuuint64_t d4[4] = {(cuuint64_t)ow, (cuuint64_t)oh, 2, (cuuint64_t)major};
        cuuint64_t s4[3] = {(cuuint64_t)ow * 4, (cuuint64_t)oh * ow * 4, (cuuint64_t)map_bstride * 4};
        cuuint32_t b4[4] = {FS_W, FS_ROWS, 2, 1}, e4[4] = {1, 1, 1, 1};
        if (enc(&tm_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(stylemap), d4, s4, b4, e4, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
            g.side_tma |= 2;
    }
    g.stage_bytes = FS_X_BYTES + ((g.side_tma & 1) ? FS_NOISE_BYTES : 0) + ((g.side_tma & 2) ? FS_MAP_BYTES : 0);
    if ((g.side_tma & 2) && !(g.side_tma & 1)) g.stage_bytes += FS_NOISE_BYTES;      // fixed offsets inside a stage
    g.stages = (g.side_tma & 2) ? 5 : FS_MAX_STAGES;                                 // two CTAs per SM must fit 227 KB
    const size_t smem = 128 + (size_t)g.stages * g.stage_bytes + 2 * FS_MAX_STAGES * sizeof(uint64_t);
    const size_t smem_max = 128 + (size_t)FS_MAX_STAGES * (FS_X_BYTES + FS_NOISE_BYTES) + 2 * FS_MAX_STAGES * sizeof(uint64_t);
    static_assert(5 * (FS_X_BYTES + FS_NOISE_BYTES + FS_MAP_BYTES) <= FS_MAX_STAGES * (FS_X_BYTES + FS_NOISE_BYTES), "ring sizes");
    const int slots = 2 * kNumSMs;
    const int grid = total < slots ? (int)total : slots;
    static bool configured[3] = {};
    auto launch = [&](auto kern, int idx) -> int {
        if (!configured[idx]) {
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max) != cudaSuccess) return SR_ERR_UNSUPPORTED;
            configured[idx] = true;
        }
        kern<<<grid, FS_CONSUMERS + 32, smem, st>>>(out, tm, tm_nz, tm_map, taps, g);
        return SR_OK;
    };
    if (mode == 2) return launch(fir_nhwc_stream_kernel<2>, 2);
    if (mode == 1) return launch(fir_nhwc_stream_kernel<1>, 1);
    return launch(fir_nhwc_stream_kernel<0>, 0);
}

int launch_nhwc(float *out, const float *x, const float *taps, int64_t major, int in_h, int in_w, int oh, int ow,
                int64_t minor, int pad_x0, int pad_y0, bool styled, const float *noise, long long noise_bstride,
                const float *noise_weight, const float *bias, float alpha, float gain, cudaStream_t st,
                float *out2 = nullptr, const float *scale2 = nullptr, const float *other = nullptr, float *dot = nullptr,
                const float *stylemap = nullptr, long long map_bstride = 0, int op16 = 0)
{
    {
        const int mode = (dot || (!styled && scale2)) ? 2 : (styled ? 1 : 0);
        if (launch_nhwc_stream(out, x, taps, major, in_h, in_w, oh, ow, minor, pad_x0, pad_y0, mode, noise, noise_bstride, noise_weight,
                               bias, alpha, gain, st, out2, scale2, other, dot, stylemap, map_bstride, op16) == SR_OK)
            return SR_OK;
    }
    NhwcGeom g;
    g.out2 = out2; g.scale2 = scale2; g.other = other; g.dot = dot;
    g.stylemap = stylemap; g.map_bstride = map_bstride; g.op16 = op16;
    g.major = major; g.in_h = in_h; g.in_w = in_w; g.out_h = oh; g.out_w = ow;
    g.c4 = (int)(minor / 4); g.pad_x0 = pad_x0; g.pad_y0 = pad_y0;
    g.rows_per_strip = oh >= 64 ? 16 : (oh >= 16 ? 8 : 4);
    g.strips_y = (oh + g.rows_per_strip - 1) / g.rows_per_strip;
    g.pairs_x = (ow + 1) / 2;
    g.total_threads = major * g.strips_y * g.pairs_x * g.c4;
    g.div_c4 = FastDiv((uint32_t)g.c4); g.div_pairs = FastDiv((uint32_t)g.pairs_x); g.div_strips = FastDiv((uint32_t)g.strips_y);
    g.noise = noise; g.noise_weight = noise_weight; g.bias = bias; g.noise_bstride = noise_bstride;
    g.alpha = alpha; g.gain = gain;
    if (g.total_threads >= 0x7fffffffll) return SR_ERR_UNSUPPORTED;
    const int64_t blocks = (g.total_threads + kThreads - 1) / kThreads;
    const int mode = (dot || (!styled && scale2)) ? 2 : (styled ? 1 : 0);
    // default: the low-instruction-count separable kernel (fir_nhwc_sep_kernel); SR_FIR_SEP=0 = the 4x5 input-window kernel
    const char *sep_env = getenv("SR_FIR_SEP");
    if (!(sep_env && sep_env[0] == '0')) {
        if (mode == 2) fir_nhwc_sep_kernel<2><<<(unsigned)blocks, kThreads, 0, st>>>(out, x, taps, g);
        else if (mode == 1) fir_nhwc_sep_kernel<1><<<(unsigned)blocks, kThreads, 0, st>>>(out, x, taps, g);
        else fir_nhwc_sep_kernel<0><<<(unsigned)blocks, kThreads, 0, st>>>(out, x, taps, g);
    } else {
        if (mode == 2) upfirdn2d_nhwc_kernel<4, 4, 2><<<(unsigned)blocks, kThreads, 0, st>>>(out, x, taps, g);
        else if (mode == 1) upfirdn2d_nhwc_kernel<4, 4, 1><<<(unsigned)blocks, kThreads, 0, st>>>(out, x, taps, g);
        else upfirdn2d_nhwc_kernel<4, 4, 0><<<(unsigned)blocks, kThreads, 0, st>>>(out, x, taps, g);
    }
    return SR_OK;
}

}  // namespace
}  // namespace sr

using namespace sr;

static int blur_nhwc_styled_any(float *out, float *out2, const float *scale2, const float *x, const float *taps,
                                int64_t batch, int64_t in_h, int64_t in_w, int64_t channels, int pad0, int pad1,
                                const float *noise, int64_t noise_batch_stride, const float *noise_weight,
                                const float *bias, float alpha, float gain, const float *stylemap,
                                int64_t stylemap_batch_stride, void *stream, int op16)
{
    SR_REQUIRE(out && x && taps, "blur_nhwc_styled: null pointer");
    SR_REQUIRE(channels >= 4 && channels % 4 == 0, "blur_nhwc_styled: channels must be a multiple of 4");
    SR_REQUIRE(((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out2) |
                 reinterpret_cast<uintptr_t>(scale2)) & 15u) == 0 &&
               (!bias || (reinterpret_cast<uintptr_t>(bias) & 15u) == 0), "blur_nhwc_styled: 16-byte alignment");
    SR_REQUIRE(!noise || noise_weight, "blur_nhwc_styled: noise needs noise_weight");
    SR_REQUIRE(!out2 || scale2, "blur_nhwc_styled: out2 needs scale2");
    const int64_t oh = in_h + pad0 + pad1 - 4 + 1, ow = in_w + pad0 + pad1 - 4 + 1;
    SR_REQUIRE(oh >= 1 && ow >= 1, "blur_nhwc_styled: FIR larger than the padded input");
    if (batch == 0) return SR_OK;
    int rc = launch_nhwc(out, x, taps, batch, (int)in_h, (int)in_w, (int)oh, (int)ow, channels, pad0, pad0, true, noise,
                         noise_batch_stride, noise_weight, bias, alpha, gain, (cudaStream_t)stream, out2, scale2, nullptr, nullptr,
                         stylemap, stylemap_batch_stride, op16);
    if (rc != SR_OK) { set_error("blur_nhwc_styled: problem too large"); return rc; }
    count_launch();
    return check_launch("sr_blur_nhwc_styled_f32");
}

extern "C" int sr_blur_nhwc_styled3_f32(float *out, float *out2, const float *scale2, const float *x, const float *taps,
                                        int64_t batch, int64_t in_h, int64_t in_w, int64_t channels, int pad0, int pad1,
                                        const float *noise, int64_t noise_batch_stride, const float *noise_weight,
                                        const float *bias, float alpha, float gain, const float *stylemap,
                                        int64_t stylemap_batch_stride, void *stream)
{
    return blur_nhwc_styled_any(out, out2, scale2, x, taps, batch, in_h, in_w, channels, pad0, pad1, noise, noise_batch_stride,
                                noise_weight, bias, alpha, gain, stylemap, stylemap_batch_stride, stream, 0);
}
// out2 is a bfloat16 tensor (the next layer's 16-bit GEMM operand); everything else as sr_blur_nhwc_styled3_f32
extern "C" int sr_blur_nhwc_styled3_bf16(float *out, void *out2, const float *scale2, const float *x, const float *taps,
                                         int64_t batch, int64_t in_h, int64_t in_w, int64_t channels, int pad0, int pad1,
                                         const float *noise, int64_t noise_batch_stride, const float *noise_weight,
                                         const float *bias, float alpha, float gain, const float *stylemap,
                                         int64_t stylemap_batch_stride, void *stream)
{
    return blur_nhwc_styled_any(out, reinterpret_cast<float *>(out2), scale2, x, taps, batch, in_h, in_w, channels, pad0, pad1, noise,
                                noise_batch_stride, noise_weight, bias, alpha, gain, stylemap, stylemap_batch_stride, stream, 1);
}

extern "C" int sr_blur_nhwc_styled2_f32(float *out, float *out2, const float *scale2, const float *x, const float *taps,
                                        int64_t batch, int64_t in_h, int64_t in_w, int64_t channels, int pad0, int pad1,
                                        const float *noise, int64_t noise_batch_stride, const float *noise_weight,
                                        const float *bias, float alpha, float gain, void *stream)
{
    return sr_blur_nhwc_styled3_f32(out, out2, scale2, x, taps, batch, in_h, in_w, channels, pad0, pad1, noise,
                                    noise_batch_stride, noise_weight, bias, alpha, gain, nullptr, 0, stream);
}

static int blur_nhwc_scaledot_any(float *out, float *dot, const float *x, const float *taps, const float *scale,
                                  const float *other, int64_t batch, int64_t in_h, int64_t in_w, int64_t channels,
                                  int pad0, int pad1, void *stream, int op16)
{
    SR_REQUIRE(out && x && taps && scale, "blur_nhwc_scaledot: null pointer");
    SR_REQUIRE((dot == nullptr) == (other == nullptr), "blur_nhwc_scaledot: dot and other go together (both NULL = scale only)");
    SR_REQUIRE(channels >= 4 && channels % 4 == 0, "blur_nhwc_scaledot: channels must be a multiple of 4");
    SR_REQUIRE(((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dot) |
                 reinterpret_cast<uintptr_t>(scale) | reinterpret_cast<uintptr_t>(other)) & 15u) == 0,
               "blur_nhwc_scaledot: 16-byte alignment");
    const int64_t oh = in_h + pad0 + pad1 - 4 + 1, ow = in_w + pad0 + pad1 - 4 + 1;
    SR_REQUIRE(oh >= 1 && ow >= 1, "blur_nhwc_scaledot: FIR larger than the padded input");
    if (dot) {
        cudaError_t er = cudaMemsetAsync(dot, 0, sizeof(float) * (size_t)(batch * channels), (cudaStream_t)stream);
        if (er != cudaSuccess) { set_error("blur_nhwc_scaledot: memset: %s", cudaGetErrorString(er)); return (int)er; }
    }
    if (batch == 0) return SR_OK;
    int rc = launch_nhwc(out, x, taps, batch, (int)in_h, (int)in_w, (int)oh, (int)ow, channels, pad0, pad0, false, nullptr, 0,
                         nullptr, nullptr, 0.f, 1.f, (cudaStream_t)stream, nullptr, scale, other, dot, nullptr, 0, op16);
    if (rc != SR_OK) { set_error("blur_nhwc_scaledot: problem too large"); return rc; }
    count_launch();
    return check_launch("sr_blur_nhwc_scaledot_f32");
}

extern "C" int sr_blur_nhwc_scaledot_f32(float *out, float *dot, const float *x, const float *taps, const float *scale,
                                         const float *other, int64_t batch, int64_t in_h, int64_t in_w, int64_t channels,
                                         int pad0, int pad1, void *stream)
{
    return blur_nhwc_scaledot_any(out, dot, x, taps, scale, other, batch, in_h, in_w, channels, pad0, pad1, stream, 0);
}
// out is a bfloat16 tensor (the 16-bit operand of the dgrad / wgrad GEMMs)
extern "C" int sr_blur_nhwc_scaledot_bf16(void *out, float *dot, const float *x, const float *taps, const float *scale,
                                          const float *other, int64_t batch, int64_t in_h, int64_t in_w, int64_t channels,
                                          int pad0, int pad1, void *stream)
{
    return blur_nhwc_scaledot_any(reinterpret_cast<float *>(out), dot, x, taps, scale, other, batch, in_h, in_w, channels, pad0,
                                  pad1, stream, 1);
}

extern "C" int sr_blur_nhwc_styled_f32(float *out, const float *x, const float *taps, int64_t batch, int64_t in_h,
                                       int64_t in_w, int64_t channels, int pad0, int pad1, const float *noise,
                                       int64_t noise_batch_stride, const float *noise_weight, const float *bias,
                                       float alpha, float gain, void *stream)
{
    return sr_blur_nhwc_styled2_f32(out, nullptr, nullptr, x, taps, batch, in_h, in_w, channels, pad0, pad1, noise,
                                    noise_batch_stride, noise_weight, bias, alpha, gain, stream);
}

extern "C" int sr_upfirdn2d_f32(float *out, const float *x, const float *taps,
                                int64_t major, int64_t in_h, int64_t in_w, int64_t minor,
                                int kernel_h, int kernel_w, int up_x, int up_y, int down_x, int down_y,
                                int pad_x0, int pad_x1, int pad_y0, int pad_y1, void *stream)
{
    SR_REQUIRE(up_x >= 1 && up_y >= 1 && down_x >= 1 && down_y >= 1, "upfirdn2d: up/down must be >= 1");
    SR_REQUIRE(kernel_h >= 1 && kernel_w >= 1, "upfirdn2d: empty FIR");
    SR_REQUIRE(major >= 0 && in_h >= 0 && in_w >= 0 && minor >= 0, "upfirdn2d: negative size");
    SR_REQUIRE(in_h < (1 << 24) && in_w < (1 << 24), "upfirdn2d: plane too large");
    const int64_t oh = (in_h * up_y + pad_y0 + pad_y1 - kernel_h) / down_y + 1;
    const int64_t ow = (in_w * up_x + pad_x0 + pad_x1 - kernel_w) / down_x + 1;
    SR_REQUIRE(in_h * up_y + pad_y0 + pad_y1 >= kernel_h && in_w * up_x + pad_x0 + pad_x1 >= kernel_w,
               "upfirdn2d: FIR larger than the padded input");
    if (major == 0 || minor == 0 || oh <= 0 || ow <= 0) return SR_OK;
    SR_REQUIRE(out && x && taps, "upfirdn2d: null pointer");
    cudaStream_t st = (cudaStream_t)stream;

    int rc = SR_ERR_UNSUPPORTED;
    if (minor == 1 && up_x == up_y && down_x == down_y && kernel_h == 4 && kernel_w == 4) {
#define SR_TILE(UP, DOWN, PHX, PHY, VY) \
    rc = launch_tile<UP, DOWN, 4, 4, PHX, PHY, VY>(out, x, taps, major, (int)in_h, (int)in_w, (int)oh, (int)ow, \
                                                   pad_x0, pad_y0, st)
        if (up_x == 1 && down_x == 1 &&
            launch_planes_stream(out, x, taps, major, (int)in_h, (int)in_w, (int)oh, (int)ow, pad_x0, pad_y0, st) == SR_OK) {
            rc = SR_OK;
        } else if (up_x == 1 && down_x == 1 && pad_x0 == pad_y0 &&
            launch_blur_rows(out, x, taps, major, (int)in_h, (int)in_w, (int)oh, (int)ow, pad_x0, pad_y0, st) == SR_OK) {
            rc = SR_OK;
        } else if (up_x == 1 && down_x == 1) {
            static const char *strip_env = getenv("SR_UPFIRDN_STRIP");      // A/B: 0 = 4x2 register patch (round-1 kernel)
            if (strip_env && strip_env[0] == '0') SR_TILE(1, 1, 0, 0, 2);
            else SR_TILE(1, 1, 0, 0, 8);
        }
        else if (up_x == 1 && down_x == 2) SR_TILE(1, 2, 0, 0, 2);
        else if (up_x == 2 && down_x == 1) {
            const int phx = pos_mod_i(pad_x0, 2), phy = pos_mod_i(pad_y0, 2);
            if (phx == 0 && phy == 0) SR_TILE(2, 1, 0, 0, 2);
            else if (phx == 1 && phy == 1) SR_TILE(2, 1, 1, 1, 2);
            else if (phx == 0 && phy == 1) SR_TILE(2, 1, 0, 1, 2);
            else SR_TILE(2, 1, 1, 0, 2);
        }
#undef SR_TILE
    }
    if (rc == SR_ERR_UNSUPPORTED && minor >= 4 && minor % 4 == 0 && up_x == 1 && up_y == 1 && down_x == 1 && down_y == 1 &&
        kernel_h == 4 && kernel_w == 4 && ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(x)) & 15u) == 0) {
        rc = launch_nhwc(out, x, taps, major, (int)in_h, (int)in_w, (int)oh, (int)ow, minor, pad_x0, pad_y0, false,
                         nullptr, 0, nullptr, nullptr, 0.f, 1.f, st);
    }
    if (rc == SR_ERR_UNSUPPORTED) {
        GenericGeom g;
        g.major = major; g.minor = minor; g.total = major * oh * ow * minor;
        g.in_h = (int)in_h; g.in_w = (int)in_w; g.out_h = (int)oh; g.out_w = (int)ow;
        g.kh = kernel_h; g.kw = kernel_w; g.up_x = up_x; g.up_y = up_y; g.down_x = down_x; g.down_y = down_y;
        g.pad_x0 = pad_x0; g.pad_y0 = pad_y0;
        int64_t blocks = (g.total + kThreads - 1) / kThreads;
        const int64_t cap = (int64_t)kNumSMs * 32;
        if (blocks > cap) blocks = cap;
        upfirdn2d_generic_kernel<<<(unsigned)blocks, kThreads, 0, st>>>(out, x, taps, g);
        rc = SR_OK;
    }
    if (rc != SR_OK) return rc;
    count_launch();
    return check_launch("sr_upfirdn2d_f32");
}
