// The style-map networks of GeneratorWithMap on B200: ResBlock(3 -> 2 / 4 channels, downsample = False) over the rasterised
// normal map at every generator resolution (reference model.py:194-216, 262, 271-275; block = reference layers.py:379-391,
// its ConvLayers layers.py:341-378).  The reference runs three cuDNN convolutions with 2-4 channels plus five elementwise
// passes per block; cuDNN's "indexed, without shared memory" kernels take milliseconds for them at 256^2 although the
// tensors are a few MB.  Here the whole block is ONE pass over [b, 3, h, w] planes, forward and backward:
//
//   y1  = lrelu(conv3x3(x,  w1 * s1) + b1c + b1a) * gain
//   y2  = lrelu(conv3x3(y1, w2 * s2) + b2c + b2a) * gain
//   out = (y2 + conv1x1(x, ws * ss)) / sqrt(2)
//
// A CTA owns tiles of 16 x 32 output pixels: the input tile with its halo and the intermediate y1 tile live in shared
// memory, nothing but x is read and nothing but out is written (HBM-bound: 12 + 4 * cout bytes per pixel).  The backward
// recomputes y1 and the two activation masks from x (cheaper than saving them), forms the two pre-activation gradients in
// shared memory and reduces the 81 + 27 cout + 3 cout + 3 + cout parameter gradients with one thread per gradient element,
// accumulated in a register across all tiles of a persistent CTA and added atomically once per CTA.
// The gradient with respect to x is not produced (the training loop rasterises under no_grad, reference train.py:249-251);
// callers that need it take the composed path.
#include "common.cuh"

namespace sr {
namespace {

constexpr int TH = 16, TW = 32, NT = 256;
constexpr int CI = 3;

struct NetGeom {
    int batch, h, w, tiles_x, tiles_y, total_tiles;
    FastDiv div_tx, div_ty;
    float s1, s2, ss, alpha, gain;
};

// smem plane strides are padded so that the three channel planes start 11 banks apart (one thread per weight-gradient
// element reads plane i at offset (ky, kx): 27 distinct addresses per warp)
__host__ __device__ constexpr int plane_stride(int rows, int cols) {
    int n = rows * cols;
    while (n % 32 != 11) ++n;
    return n;
}

template <int HALO>     // input tile with HALO pixels around the TH x TW core, zero outside the image
__device__ __forceinline__ void load_planes(float *dst, const float *__restrict__ src, int nch, int oy0, int ox0, int h, int w)
{
    constexpr int R = TH + 2 * HALO, C = TW + 2 * HALO, PS = plane_stride(R, C);
    for (int e = threadIdx.x; e < nch * R * C; e += NT) {
        const int c = e / (R * C), r = (e / C) % R, q = e % C;
        const int y = oy0 - HALO + r, x = ox0 - HALO + q;
        dst[c * PS + r * C + q] = (y >= 0 && y < h && x >= 0 && x < w) ? __ldg(src + ((int64_t)c * h + y) * w + x) : 0.0f;
    }
}

// y1 on the core + HALO ring from the x tile with HALO + 1 ring; exactly 0 outside the image (= conv2's zero padding)
template <int HALO>
__device__ __forceinline__ void compute_y1(float *y1s, const float *xs, const float *w1s, const float *b1s, int oy0, int ox0,
                                           int h, int w, float alpha, float gain)
{
    constexpr int R = TH + 2 * HALO, C = TW + 2 * HALO, PS = plane_stride(R, C);
    constexpr int XR = R + 2, XC = C + 2, XPS = plane_stride(XR, XC);
    for (int e = threadIdx.x; e < R * C; e += NT) {
        const int r = e / C, q = e % C;
        const int y = oy0 - HALO + r, x = ox0 - HALO + q;
        float a[CI];
#pragma unroll
        for (int o = 0; o < CI; ++o) a[o] = b1s[o];
#pragma unroll
        for (int i = 0; i < CI; ++i)
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const float v = xs[i * XPS + (r + ky) * XC + q + kx];
#pragma unroll
                    for (int o = 0; o < CI; ++o) a[o] = fmaf(w1s[((o * CI + i) * 3 + ky) * 3 + kx], v, a[o]);
                }
        const bool in = y >= 0 && y < h && x >= 0 && x < w;
#pragma unroll
        for (int o = 0; o < CI; ++o) y1s[o * PS + r * C + q] = in ? ((a[o] > 0.f) ? a[o] : a[o] * alpha) * gain : 0.0f;
    }
}

struct NetParams {
    const float *w1, *b1c, *b1a, *w2, *b2c, *b2a, *ws;
};

template <int CO>
__device__ __forceinline__ void stage_weights(float *w1s, float *b1s, float *w2s, float *b2s, float *wss, const NetParams p,
                                              const NetGeom g)
{
    for (int e = threadIdx.x; e < CI * CI * 9; e += NT) w1s[e] = __ldg(p.w1 + e) * g.s1;
    for (int e = threadIdx.x; e < CO * CI * 9; e += NT) w2s[e] = __ldg(p.w2 + e) * g.s2;
    for (int e = threadIdx.x; e < CO * CI; e += NT) wss[e] = __ldg(p.ws + e) * g.ss;
    if (threadIdx.x < CI) b1s[threadIdx.x] = (p.b1c ? __ldg(p.b1c + threadIdx.x) : 0.f) + (p.b1a ? __ldg(p.b1a + threadIdx.x) : 0.f);
    if (threadIdx.x < CO) b2s[threadIdx.x] = (p.b2c ? __ldg(p.b2c + threadIdx.x) : 0.f) + (p.b2a ? __ldg(p.b2a + threadIdx.x) : 0.f);
}

template <int CO>
__global__ void __launch_bounds__(NT)
stylemap_resblock_fwd_kernel(float *__restrict__ out, const float *__restrict__ x, const NetParams p, const NetGeom g)
{
    constexpr int XPS = plane_stride(TH + 4, TW + 4), YPS = plane_stride(TH + 2, TW + 2), YC = TW + 2, XC = TW + 4;
    __shared__ float xs[CI * XPS], y1s[CI * YPS];
    __shared__ float w1s[CI * CI * 9], w2s[CO * CI * 9], wss[CO * CI], b1s[CI], b2s[CO];
    stage_weights<CO>(w1s, b1s, w2s, b2s, wss, p, g);
    const float inv_sqrt2 = 0.70710678118654752440f;
    for (uint32_t T = blockIdx.x; T < (uint32_t)g.total_tiles; T += gridDim.x) {
        uint32_t t = T, tx, ty, n;
        g.div_tx.divmod(t, t, tx);
        g.div_ty.divmod(t, n, ty);
        const int oy0 = ty * TH, ox0 = tx * TW;
        __syncthreads();                                              // weights staged / previous tile consumed
        load_planes<2>(xs, x + (int64_t)n * CI * g.h * g.w, CI, oy0, ox0, g.h, g.w);
        __syncthreads();
        compute_y1<1>(y1s, xs, w1s, b1s, oy0, ox0, g.h, g.w, g.alpha, g.gain);
        __syncthreads();
        for (int e = threadIdx.x; e < TH * TW; e += NT) {
            const int r = e / TW, q = e % TW;
            const int y = oy0 + r, xx = ox0 + q;
            if (y >= g.h || xx >= g.w) continue;
            float a[CO];
#pragma unroll
            for (int o = 0; o < CO; ++o) a[o] = b2s[o];
#pragma unroll
            for (int i = 0; i < CI; ++i)
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const float v = y1s[i * YPS + (r + ky) * YC + q + kx];
#pragma unroll
                        for (int o = 0; o < CO; ++o) a[o] = fmaf(w2s[((o * CI + i) * 3 + ky) * 3 + kx], v, a[o]);
                    }
#pragma unroll
            for (int o = 0; o < CO; ++o) {
                float s = 0.0f;
#pragma unroll
                for (int i = 0; i < CI; ++i) s = fmaf(wss[o * CI + i], xs[i * XPS + (r + 2) * XC + q + 2], s);
                const float y2 = ((a[o] > 0.f) ? a[o] : a[o] * g.alpha) * g.gain;
                out[(((int64_t)n * CO + o) * g.h + y) * g.w + xx] = (y2 + s) * inv_sqrt2;
            }
        }
    }
}

// grads layout: [dW1 (CI*CI*9) | db1 (CI) | dW2 (CO*CI*9) | db2 (CO) | dWs (CO*CI)]
template <int CO>
__global__ void __launch_bounds__(NT)
stylemap_resblock_bwd_kernel(float *__restrict__ grads, const float *__restrict__ gout, const float *__restrict__ x,
                             const NetParams p, const NetGeom g)
{
    constexpr int XR = TH + 6, XC = TW + 6, XPS = plane_stride(XR, XC);            // x: 3-pixel ring
    constexpr int YR = TH + 4, YC = TW + 4, YPS = plane_stride(YR, YC);            // y1: 2-pixel ring
    constexpr int GR = TH + 2, GC = TW + 2, GPS = plane_stride(GR, GC);            // g / gp2: 1-pixel ring
    constexpr int PPS = plane_stride(TH, TW);                                      // gp1: core
    constexpr int NW1 = CI * CI * 9, NW2 = CO * CI * 9, NACC = NW1 + CI + NW2 + CO + CO * CI;
    static_assert(NACC <= NT, "one thread per gradient element");
    extern __shared__ float sm[];
    float *xs = sm, *y1s = xs + CI * XPS, *gs = y1s + CI * YPS, *gp2s = gs + CO * GPS, *gp1s = gp2s + CO * GPS;
    float *w1s = gp1s + CI * PPS, *w2s = w1s + NW1, *wss = w2s + NW2, *b1s = wss + CO * CI, *b2s = b1s + CI;
    stage_weights<CO>(w1s, b1s, w2s, b2s, wss, p, g);
    const float inv_sqrt2 = 0.70710678118654752440f;

    // this thread's gradient element: acc += A[p] * B[p] over the core pixels, A / B = shared-memory planes with offsets
    const int t = threadIdx.x;
    const float *pa = nullptr, *pb = nullptr;
    int sa = 0, sb = 0;
    float scale = 0.0f;
    if (t < NW1) {                                   // dW1[o,i,ky,kx] = sum gp1[o,p] * x[i, p + (ky-1, kx-1)]
        const int o = t / (CI * 9), i = (t / 9) % CI, ky = (t / 3) % 3, kx = t % 3;
        pa = gp1s + o * PPS; sa = TW;
        pb = xs + i * XPS + (3 + ky - 1) * XC + 3 + kx - 1; sb = XC;
        scale = g.s1;
    } else if (t < NW1 + CI) {                       // db1[o] = sum gp1[o,p]
        pa = gp1s + (t - NW1) * PPS; sa = TW;
        scale = 1.0f;
    } else if (t < NW1 + CI + NW2) {                 // dW2[o,i,ky,kx] = sum gp2[o,p] * y1[i, p + (ky-1, kx-1)]
        const int u = t - NW1 - CI;
        const int o = u / (CI * 9), i = (u / 9) % CI, ky = (u / 3) % 3, kx = u % 3;
        pa = gp2s + o * GPS + GC + 1; sa = GC;
        pb = y1s + i * YPS + (2 + ky - 1) * YC + 2 + kx - 1; sb = YC;
        scale = g.s2;
    } else if (t < NW1 + CI + NW2 + CO) {            // db2[o] = sum gp2[o,p]
        pa = gp2s + (t - NW1 - CI - NW2) * GPS + GC + 1; sa = GC;
        scale = 1.0f;
    } else if (t < NACC) {                           // dWs[o,i] = sum g[o,p] / sqrt(2) * x[i,p]
        const int u = t - NW1 - CI - NW2 - CO;
        pa = gs + (u / CI) * GPS + GC + 1; sa = GC;
        pb = xs + (u % CI) * XPS + 3 * XC + 3; sb = XC;
        scale = g.ss * inv_sqrt2;
    }
    float acc = 0.0f;

    for (uint32_t T = blockIdx.x; T < (uint32_t)g.total_tiles; T += gridDim.x) {
        uint32_t tt = T, tx, ty, n;
        g.div_tx.divmod(tt, tt, tx);
        g.div_ty.divmod(tt, n, ty);
        const int oy0 = ty * TH, ox0 = tx * TW;
        __syncthreads();
        load_planes<3>(xs, x + (int64_t)n * CI * g.h * g.w, CI, oy0, ox0, g.h, g.w);
        load_planes<1>(gs, gout + (int64_t)n * CO * g.h * g.w, CO, oy0, ox0, g.h, g.w);
        __syncthreads();
        compute_y1<2>(y1s, xs, w1s, b1s, oy0, ox0, g.h, g.w, g.alpha, g.gain);
        __syncthreads();
        // gp2 = d loss / d pre2 on the core + 1 ring (0 outside the image: g is 0 there)
        for (int e = t; e < GR * GC; e += NT) {
            const int r = e / GC, q = e % GC;
            float a[CO];
#pragma unroll
            for (int o = 0; o < CO; ++o) a[o] = b2s[o];
#pragma unroll
            for (int i = 0; i < CI; ++i)
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const float v = y1s[i * YPS + (r + ky) * YC + q + kx];
#pragma unroll
                        for (int o = 0; o < CO; ++o) a[o] = fmaf(w2s[((o * CI + i) * 3 + ky) * 3 + kx], v, a[o]);
                    }
#pragma unroll
            for (int o = 0; o < CO; ++o)
                gp2s[o * GPS + r * GC + q] = gs[o * GPS + r * GC + q] * (inv_sqrt2 * g.gain) * ((a[o] > 0.f) ? 1.0f : g.alpha);
        }
        __syncthreads();
        // gp1 = d loss / d pre1 on the core: transposed conv2 of gp2, times the activation slope of y1
        for (int e = t; e < TH * TW; e += NT) {
            const int r = e / TW, q = e % TW;
            float a[CI];
#pragma unroll
            for (int i = 0; i < CI; ++i) a[i] = 0.0f;
#pragma unroll
            for (int o = 0; o < CO; ++o)
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const float v = gp2s[o * GPS + (r + 1 + 1 - ky) * GC + q + 1 + 1 - kx];   // pre2 at p - (ky-1, kx-1)
#pragma unroll
                        for (int i = 0; i < CI; ++i) a[i] = fmaf(w2s[((o * CI + i) * 3 + ky) * 3 + kx], v, a[i]);
                    }
            const bool in = oy0 + r < g.h && ox0 + q < g.w;            // ragged tiles: no pre-activation outside the image
#pragma unroll
            for (int i = 0; i < CI; ++i)
                gp1s[i * PPS + r * TW + q] = in ? a[i] * g.gain * ((y1s[i * YPS + (r + 2) * YC + q + 2] > 0.f) ? 1.0f : g.alpha) : 0.0f;
        }
        __syncthreads();
        if (pa) {
            if (pb) {
                for (int r = 0; r < TH; ++r)
#pragma unroll 8
                    for (int q = 0; q < TW; ++q) acc = fmaf(pa[r * sa + q], pb[r * sb + q], acc);
            } else {
                for (int r = 0; r < TH; ++r)
#pragma unroll 8
                    for (int q = 0; q < TW; ++q) acc += pa[r * sa + q];
            }
        }
    }
    if (pa) atomicAdd(grads + t, acc * scale);
}

NetGeom make_geom(int64_t batch, int64_t h, int64_t w, int cin, int cout, float alpha, float gain)
{
    NetGeom g;
    g.batch = (int)batch; g.h = (int)h; g.w = (int)w;
    g.tiles_x = (int)((w + TW - 1) / TW); g.tiles_y = (int)((h + TH - 1) / TH);
    g.total_tiles = (int)(batch * g.tiles_x * g.tiles_y);
    g.div_tx = FastDiv((uint32_t)g.tiles_x); g.div_ty = FastDiv((uint32_t)g.tiles_y);
    g.s1 = 1.0f / sqrtf((float)(cin * 9)); g.s2 = g.s1; g.ss = 1.0f / sqrtf((float)cin);       // EqualConv2d scales, reference layers.py:209
    g.alpha = alpha; g.gain = gain;
    return g;
}

int check_args(const void *a, const void *b, const float *w1, const float *w2, const float *ws, int64_t batch, int cin, int cout,
               int64_t h, int64_t w, const char *what)
{
    SR_REQUIRE(a && b && w1 && w2 && ws, "%s: null pointer", what);
    SR_REQUIRE(cin == CI && (cout == 2 || cout == 4), "%s: built for 3 -> 2 or 3 -> 4 channels (got %d -> %d)", what, cin, cout);
    SR_REQUIRE(batch >= 1 && h >= 1 && w >= 1 && batch * ((h + TH - 1) / TH) * ((w + TW - 1) / TW) < 0x7fffffffll, "%s: bad sizes", what);
    return SR_OK;
}

// ---- small-channel convolution pair (<= 8 channels): any-order differentiable building blocks ---------------------------
// The regulariser iterations (path length: reference train.py:335-354 differentiates the image with respect to the latents AND
// the normal maps with create_graph=True) need the style-map nets twice differentiable.  torch's double backward of a
// cuDNN conv computes weight gradients as convolutions with 256 x 256 "kernels" (12.5 ms per call at [8,3,256,256]).  A
// stride-1 convolution and its weight gradient are bilinear maps whose derivatives are each other:
//   y  = conv(x, w)            dx = conv(gy, flipT(w))        dw = wgrad(gy, x)
//   dw = wgrad(g, x)           dg = conv(x, gdw)              dx = conv(g, flipT(gdw))
// so two kernels cover every order (fused.SmallConvFn / SmallWgradFn).
constexpr int SC_MAX = 8;

struct SmallConvGeom {
    int batch, ci, co, k, h, w;                 // k = 1 or 3 (zero padding k / 2), stride 1
    int64_t total;                              // batch * h * w
};

__global__ void __launch_bounds__(NT)
small_conv_kernel(float *__restrict__ y, const float *__restrict__ x, const float *__restrict__ wgt, const SmallConvGeom g)
{
    __shared__ float ws[SC_MAX * SC_MAX * 9];
    for (int e = threadIdx.x; e < g.co * g.ci * g.k * g.k; e += NT) ws[e] = __ldg(wgt + e);
    __syncthreads();
    const int pad = g.k / 2, kk = g.k * g.k;
    const int64_t plane = (int64_t)g.h * g.w;
    for (int64_t idx = (int64_t)blockIdx.x * NT + threadIdx.x; idx < g.total; idx += (int64_t)gridDim.x * NT) {
        const int px = (int)(idx % g.w), py = (int)((idx / g.w) % g.h);
        const int64_t n = idx / plane;
        float acc[SC_MAX];
#pragma unroll
        for (int o = 0; o < SC_MAX; ++o) acc[o] = 0.0f;
        for (int i = 0; i < g.ci; ++i) {
            const float *xp = x + (n * g.ci + i) * plane;
            for (int ky = 0; ky < g.k; ++ky) {
                const int yy = py + ky - pad;
                if (yy < 0 || yy >= g.h) continue;
                for (int kx = 0; kx < g.k; ++kx) {
                    const int xx = px + kx - pad;
                    if (xx < 0 || xx >= g.w) continue;
                    const float v = __ldg(xp + (int64_t)yy * g.w + xx);
                    const float *wp = ws + i * kk + ky * g.k + kx;
#pragma unroll
                    for (int o = 0; o < SC_MAX; ++o)
                        if (o < g.co) acc[o] = fmaf(wp[o * g.ci * kk], v, acc[o]);
                }
            }
        }
#pragma unroll
        for (int o = 0; o < SC_MAX; ++o)
            if (o < g.co) y[(n * g.co + o) * plane + (int64_t)py * g.w + px] = acc[o];
    }
}

// dw[o,i,ky,kx] = sum_{n,p} gy[n,o,p] * x[n,i,p + (ky,kx) - pad]; one thread per element (co*ci*k*k <= 256), tiles of
// TH x TW pixels staged in shared memory, register accumulation across the tiles of a persistent CTA, one atomic per CTA
__global__ void __launch_bounds__(NT)
small_wgrad_kernel(float *__restrict__ dw, const float *__restrict__ gy, const float *__restrict__ x, const SmallConvGeom g,
                   const NetGeom tg)
{
    constexpr int XR = TH + 2, XC = TW + 2, XPS = plane_stride(XR, XC), GPS = plane_stride(TH, TW);
    __shared__ float xs[SC_MAX * XPS], gs[SC_MAX * GPS];
    const int pad = g.k / 2, kk = g.k * g.k, nacc = g.co * g.ci * kk;
    const int t = threadIdx.x;
    const float *pa = nullptr, *pb = nullptr;
    if (t < nacc) {
        const int o = t / (g.ci * kk), i = (t / kk) % g.ci, ky = (t % kk) / g.k, kx = t % g.k;
        pa = gs + o * GPS;
        pb = xs + i * XPS + (1 + ky - pad) * XC + 1 + kx - pad;
    }
    float acc = 0.0f;
    for (uint32_t T = blockIdx.x; T < (uint32_t)tg.total_tiles; T += gridDim.x) {
        uint32_t tt = T, tx, ty, n;
        tg.div_tx.divmod(tt, tt, tx);
        tg.div_ty.divmod(tt, n, ty);
        const int oy0 = ty * TH, ox0 = tx * TW;
        __syncthreads();
        load_planes<1>(xs, x + (int64_t)n * g.ci * g.h * g.w, g.ci, oy0, ox0, g.h, g.w);
        load_planes<0>(gs, gy + (int64_t)n * g.co * g.h * g.w, g.co, oy0, ox0, g.h, g.w);
        __syncthreads();
        if (pa)
            for (int r = 0; r < TH; ++r)
#pragma unroll 8
                for (int q = 0; q < TW; ++q) acc = fmaf(pa[r * TW + q], pb[r * XC + q], acc);
    }
    if (pa) atomicAdd(dw + t, acc);
}

}  // namespace
}  // namespace sr

using namespace sr;

extern "C" int sr_small_conv_f32(float *y, const float *x, const float *w, int64_t batch, int cin, int cout, int ksize,
                                 int64_t h, int64_t wd, void *stream)
{
    SR_REQUIRE(y && x && w, "small_conv: null pointer");
    SR_REQUIRE(cin >= 1 && cin <= SC_MAX && cout >= 1 && cout <= SC_MAX && (ksize == 1 || ksize == 3),
               "small_conv: 1..8 channels, 1x1 or 3x3 (got %d -> %d, k = %d)", cin, cout, ksize);
    SR_REQUIRE(batch >= 1 && h >= 1 && wd >= 1 && h < (1 << 20) && wd < (1 << 20), "small_conv: bad sizes");
    SmallConvGeom g = {(int)batch, cin, cout, ksize, (int)h, (int)wd, batch * h * wd};
    int64_t blocks = (g.total + NT - 1) / NT;
    if (blocks > (int64_t)kNumSMs * 16) blocks = (int64_t)kNumSMs * 16;
    small_conv_kernel<<<(unsigned)blocks, NT, 0, (cudaStream_t)stream>>>(y, x, w, g);
    count_launch();
    return check_launch("small_conv");
}

extern "C" int sr_small_conv_wgrad_f32(float *dw, const float *gy, const float *x, int64_t batch, int cin, int cout, int ksize,
                                       int64_t h, int64_t wd, void *stream)
{
    SR_REQUIRE(dw && gy && x, "small_conv_wgrad: null pointer");
    SR_REQUIRE(cin >= 1 && cin <= SC_MAX && cout >= 1 && cout <= SC_MAX && (ksize == 1 || ksize == 3) &&
               cin * cout * ksize * ksize <= NT, "small_conv_wgrad: cin * cout * k * k must be <= 256 (got %d -> %d, k = %d)", cin, cout, ksize);
    SR_REQUIRE(batch >= 1 && h >= 1 && wd >= 1 && batch * ((h + TH - 1) / TH) * ((wd + TW - 1) / TW) < 0x7fffffffll, "small_conv_wgrad: bad sizes");
    SmallConvGeom g = {(int)batch, cin, cout, ksize, (int)h, (int)wd, batch * h * wd};
    const NetGeom tg = make_geom(batch, h, wd, cin, cout, 0.f, 1.f);
    cudaStream_t st = (cudaStream_t)stream;
    if (cudaMemsetAsync(dw, 0, sizeof(float) * cin * cout * ksize * ksize, st) != cudaSuccess) return check_launch("small_conv_wgrad (memset)");
    const int grid = tg.total_tiles < 2 * kNumSMs ? tg.total_tiles : 2 * kNumSMs;
    small_wgrad_kernel<<<grid, NT, 0, st>>>(dw, gy, x, g, tg);
    count_launch();
    return check_launch("small_conv_wgrad");
}

extern "C" int sr_stylemap_resblock_forward_f32(float *out, const float *x, const float *w1, const float *b1_conv, const float *b1_act,
                                                const float *w2, const float *b2_conv, const float *b2_act, const float *w_skip,
                                                int64_t batch, int cin, int cout, int64_t h, int64_t w, float alpha, float gain,
                                                void *stream)
{
    int rc = check_args(out, x, w1, w2, w_skip, batch, cin, cout, h, w, "stylemap_resblock_forward");
    if (rc != SR_OK) return rc;
    const NetGeom g = make_geom(batch, h, w, cin, cout, alpha, gain);
    const NetParams p = {w1, b1_conv, b1_act, w2, b2_conv, b2_act, w_skip};
    const int grid = g.total_tiles < 4 * kNumSMs ? g.total_tiles : 4 * kNumSMs;
    cudaStream_t st = (cudaStream_t)stream;
    if (cout == 2) stylemap_resblock_fwd_kernel<2><<<grid, NT, 0, st>>>(out, x, p, g);
    else stylemap_resblock_fwd_kernel<4><<<grid, NT, 0, st>>>(out, x, p, g);
    count_launch();
    return check_launch("stylemap_resblock_forward");
}

extern "C" int sr_stylemap_resblock_backward_f32(float *grads, const float *grad_out, const float *x, const float *w1,
                                                 const float *b1_conv, const float *b1_act, const float *w2, const float *b2_conv,
                                                 const float *b2_act, const float *w_skip, int64_t batch, int cin, int cout,
                                                 int64_t h, int64_t w, float alpha, float gain, void *stream)
{
    int rc = check_args(grads, grad_out, w1, w2, w_skip, batch, cin, cout, h, w, "stylemap_resblock_backward");
    if (rc != SR_OK) return rc;
    SR_REQUIRE(x, "stylemap_resblock_backward: null pointer");
    const NetGeom g = make_geom(batch, h, w, cin, cout, alpha, gain);
    const NetParams p = {w1, b1_conv, b1_act, w2, b2_conv, b2_act, w_skip};
    cudaStream_t st = (cudaStream_t)stream;
    const int nacc = cin * cin * 9 + cin + cout * cin * 9 + cout + cout * cin;
    if (cudaMemsetAsync(grads, 0, sizeof(float) * nacc, st) != cudaSuccess) return check_launch("stylemap_resblock_backward (memset)");
    const int grid = g.total_tiles < 2 * kNumSMs ? g.total_tiles : 2 * kNumSMs;
    auto launch = [&](auto kern, int co) -> int {
        const size_t smem = sizeof(float) * (CI * plane_stride(TH + 6, TW + 6) + CI * plane_stride(TH + 4, TW + 4) +
                                             2 * co * plane_stride(TH + 2, TW + 2) + CI * plane_stride(TH, TW) +
                                             CI * CI * 9 + co * CI * 9 + co * CI + CI + co);
        static bool configured[5] = {};
        if (!configured[co]) {
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
                return check_launch("stylemap_resblock_backward (smem)");
            configured[co] = true;
        }
        kern<<<grid, NT, smem, st>>>(grads, grad_out, x, p, g);
        return SR_OK;
    };
    rc = cout == 2 ? launch(stylemap_resblock_bwd_kernel<2>, 2) : launch(stylemap_resblock_bwd_kernel<4>, 4);
    if (rc != SR_OK) return rc;
    count_launch();
    return check_launch("stylemap_resblock_backward");
}
