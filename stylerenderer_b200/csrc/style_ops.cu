// Style path of ModulatedConv2d for ALL layers of a generator in a handful of launches (sm_100a).
//
// Reference (layers.py:232-239, 295-299): per layer  s = EqualLinear(style)  (one F.linear + scale / bias multiplies),
// w = scale * W * s, d = rsqrt(sum w^2 + 1e-8).  In the activation-scaling form used here (layers.py of this package)
//   s[b,i] = c * sum_k latent[b, li, k] * Wm[i,k] + lr_mul * bm[i]
//   d[b,o] = rsqrt( sum_i s[b,i]^2 * Wsq[o,i] + eps ),   Wsq[o,i] = scale^2 * sum_taps W[o,i,:]^2
// Through torch this is ~10 tiny launches per layer forward and ~25 backward (SIMT sgemms of 32 x 512 x 512, scale
// multiplies, pow / sum over the full 9.4 MB weights ...): ~900 launches, 2.7 ms of a 21.5 ms generator step.
// Here every layer of the network is one blockIdx.y slice of the same launch:
//   forward : rows_dot (s)  -> rows_dot (d)
//   backward: colsum (du, gs_total) -> outer (dWsq) -> outer (dWm, dbm) -> colsum (dlatent)
// plus the two elementwise passes over a conv weight (Wsq and its gradient).  All of it is bandwidth / latency trivial
// (a few MB per layer out of L2); the point is the launch count.
#include "common.cuh"

namespace sr {
namespace {

constexpr int kMaxStyleLayers = 32;
constexpr int kSB = 4;                                   // samples per CTA in the matrix-vector kernels

struct StyleTable { sr_style_layer l[kMaxStyleLayers]; };

// ---- out[b, r] = g( sum_c W[r, c] * f(X[b, c]) ) for r in a 64-row chunk, kSB samples ---------------------------------
//   KIND 0 (s): X = latent[b, li, :], W = Wm [cin, K], g = scale * dot + lr_mul * bias[r]
//   KIND 1 (d): X = s[b, :] squared,  W = Wsq [cout, cin], g = rsqrt(dot + eps)
template <int KIND>
__global__ void __launch_bounds__(256)
style_rows_dot_kernel(const StyleTable tab, const float *__restrict__ latent, int batch, int n_latent, int style_dim,
                      float mod_scale, float lr_mul, float eps)
{
    const sr_style_layer &L = tab.l[blockIdx.y];
    const int rows = KIND == 0 ? L.cin : L.cout;
    const int cols = KIND == 0 ? style_dim : L.cin;
    const float *W = KIND == 0 ? L.mod_weight : L.wsq;
    if (KIND == 1 && W == nullptr) return;
    const int r0 = blockIdx.x * 64;
    if (r0 >= rows) return;
    const int b0 = blockIdx.z * kSB;
    extern __shared__ float xs[];                        // [kSB][cols]
    for (int idx = threadIdx.x; idx < kSB * cols; idx += 256) {
        const int sb = idx / cols, c = idx - sb * cols;
        const int b = b0 + sb;
        float v = 0.0f;
        if (b < batch) {
            if (KIND == 0) v = __ldg(latent + ((int64_t)b * n_latent + L.latent_index) * style_dim + c);
            else { v = L.s[(int64_t)b * L.cin + c]; v *= v; }
        }
        xs[idx] = v;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int rr = warp; rr < 64; rr += 8) {
        const int r = r0 + rr;
        if (r >= rows) break;
        const float *wrow = W + (int64_t)r * cols;
        float acc[kSB];
#pragma unroll
        for (int sb = 0; sb < kSB; ++sb) acc[sb] = 0.0f;
        for (int c = lane; c < cols; c += 32) {
            const float w = __ldg(wrow + c);
#pragma unroll
            for (int sb = 0; sb < kSB; ++sb) acc[sb] = fmaf(w, xs[sb * cols + c], acc[sb]);
        }
#pragma unroll
        for (int sb = 0; sb < kSB; ++sb) acc[sb] = warp_sum(acc[sb]);
        if (lane < kSB && b0 + lane < batch) {
            float v = acc[0];
#pragma unroll
            for (int sb = 1; sb < kSB; ++sb) if (lane == sb) v = acc[sb];
            if (KIND == 0) L.s[(int64_t)(b0 + lane) * L.cin + r] = v * mod_scale + lr_mul * __ldg(L.mod_bias + r);
            else L.d[(int64_t)(b0 + lane) * L.cout + r] = rsqrtf(v + eps);
        }
    }
}

// ---- backward, per (layer, sample): du = -1/2 g_d d^3;  gs_total = g_s + 2 s (du @ Wsq) ------------------------------
// A CTA owns 32 input channels (one per lane); its 8 warps split the sum over the output channels and combine through
// shared memory (a thread-per-channel loop over all 512 outputs was a 0.12 ms latency chain per launch).
__global__ void __launch_bounds__(256)
style_bwd_gs_kernel(const StyleTable tab, int batch)
{
    const sr_style_layer &L = tab.l[blockIdx.y];
    const int i0 = blockIdx.x * 32;
    if (i0 >= L.cin) return;
    const int b0 = blockIdx.z * kSB;
    extern __shared__ float du_s[];                      // [kSB][cout]
    __shared__ float part[8][kSB][32];
    const bool demod = L.wsq != nullptr && L.g_d != nullptr;
    if (demod) {
        for (int idx = threadIdx.x; idx < kSB * L.cout; idx += 256) {
            const int sb = idx / L.cout, o = idx - sb * L.cout;
            const int b = b0 + sb;
            float v = 0.0f;
            if (b < batch) {
                const float dd = L.d[(int64_t)b * L.cout + o];
                v = -0.5f * __ldg(L.g_d + (int64_t)b * L.cout + o) * dd * dd * dd;
                if (blockIdx.x == 0) L.du[(int64_t)b * L.cout + o] = v;
            }
            du_s[idx] = v;
        }
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = i0 + lane;
    float t[kSB];
#pragma unroll
    for (int sb = 0; sb < kSB; ++sb) t[sb] = 0.0f;
    if (demod && i < L.cin) {
        for (int o = warp; o < L.cout; o += 8) {
            const float w = __ldg(L.wsq + (int64_t)o * L.cin + i);
#pragma unroll
            for (int sb = 0; sb < kSB; ++sb) t[sb] = fmaf(du_s[sb * L.cout + o], w, t[sb]);
        }
    }
#pragma unroll
    for (int sb = 0; sb < kSB; ++sb) part[warp][sb][lane] = t[sb];
    __syncthreads();
    if (warp < kSB && i < L.cin && b0 + warp < batch) {      // warp sb finishes sample sb
        float sum = 0.0f;
#pragma unroll
        for (int w = 0; w < 8; ++w) sum += part[w][warp][lane];
        const int b = b0 + warp;
        const float gs = L.g_s ? __ldg(L.g_s + (int64_t)b * L.cin + i) : 0.0f;
        L.gs_total[(int64_t)b * L.cin + i] = gs + 2.0f * L.s[(int64_t)b * L.cin + i] * sum;
    }
}

// ---- rank-`batch` outer products: out[r, c] = alpha * sum_b A[b, r] * f(B[b, c]) --------------------------------------
//   KIND 0: g_wsq[o, i]        = sum_b du[b,o] * s[b,i]^2
//   KIND 1: g_mod_weight[i, k] = scale * sum_b gs_total[b,i] * latent[b, li, k];  g_mod_bias[i] = lr_mul * sum_b gs_total[b,i]
template <int KIND>
__global__ void __launch_bounds__(256)
style_bwd_outer_kernel(const StyleTable tab, const float *__restrict__ latent, int batch, int n_latent, int style_dim,
                       float mod_scale, float lr_mul)
{
    const sr_style_layer &L = tab.l[blockIdx.y];
    const int rows = KIND == 0 ? L.cout : L.cin;
    const int cols = KIND == 0 ? L.cin : style_dim;
    float *out = KIND == 0 ? L.g_wsq : L.g_mod_weight;
    if (out == nullptr) return;
    if (KIND == 0 && (L.wsq == nullptr || L.g_d == nullptr)) {        // no demodulation gradient: dWsq = 0
        const int64_t total = (int64_t)rows * cols;
        for (int64_t i = ((int64_t)blockIdx.z * gridDim.x + blockIdx.x) * 256 + threadIdx.x; i < total;
             i += (int64_t)gridDim.x * gridDim.z * 256) out[i] = 0.0f;
        return;
    }
    const int c = blockIdx.x * 256 + threadIdx.x;
    const int r0 = blockIdx.z * 16;
    if (blockIdx.x * 256 >= cols || r0 >= rows) return;
    const float *A = KIND == 0 ? L.du : L.gs_total;               // [batch, rows]
    __shared__ float a_s[32][16];
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.0f;
    for (int bb = 0; bb < batch; bb += 32) {
        const int nb = min(32, batch - bb);
        __syncthreads();
        for (int idx = threadIdx.x; idx < 32 * 16; idx += 256) {
            const int b = idx >> 4, j = idx & 15;
            a_s[b][j] = (b < nb && r0 + j < rows) ? A[(int64_t)(bb + b) * rows + r0 + j] : 0.0f;
        }
        __syncthreads();
        if (c < cols) {
            for (int b = 0; b < nb; ++b) {
                float x;
                if (KIND == 0) { x = L.s[(int64_t)(bb + b) * L.cin + c]; x *= x; }
                else x = __ldg(latent + ((int64_t)(bb + b) * n_latent + L.latent_index) * style_dim + c);
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[j] = fmaf(a_s[b][j], x, acc[j]);
            }
        }
    }
    if (c < cols) {
        const float alpha = KIND == 0 ? 1.0f : mod_scale;
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (r0 + j < rows) out[(int64_t)(r0 + j) * cols + c] = alpha * acc[j];
    }
    if (KIND == 1 && blockIdx.x == 0 && threadIdx.x < 16 && r0 + threadIdx.x < rows && L.g_mod_bias) {
        float sum = 0.0f;
        for (int b = 0; b < batch; ++b) sum += L.gs_total[(int64_t)b * rows + r0 + threadIdx.x];
        L.g_mod_bias[r0 + threadIdx.x] = lr_mul * sum;
    }
}

// ---- g_latent[b, li, k] += scale * sum_i gs_total[b,i] * Wm[i,k] -------------------------------------------------------
// Same shape of work: a CTA owns 32 latent components (one per lane), its 8 warps split the sum over the channels.
__global__ void __launch_bounds__(256)
style_bwd_latent_kernel(const StyleTable tab, float *__restrict__ g_latent, int batch, int n_latent, int style_dim,
                        float mod_scale)
{
    const sr_style_layer &L = tab.l[blockIdx.y];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k = blockIdx.x * 32 + lane;
    const int b0 = blockIdx.z * kSB;
    extern __shared__ float gs_s[];                      // [kSB][cin]
    __shared__ float part[8][kSB][32];
    for (int idx = threadIdx.x; idx < kSB * L.cin; idx += 256) {
        const int sb = idx / L.cin, i = idx - sb * L.cin;
        gs_s[idx] = (b0 + sb < batch) ? L.gs_total[(int64_t)(b0 + sb) * L.cin + i] : 0.0f;
    }
    __syncthreads();
    float acc[kSB];
#pragma unroll
    for (int sb = 0; sb < kSB; ++sb) acc[sb] = 0.0f;
    if (k < style_dim) {
        for (int i = warp; i < L.cin; i += 8) {
            const float w = __ldg(L.mod_weight + (int64_t)i * style_dim + k);
#pragma unroll
            for (int sb = 0; sb < kSB; ++sb) acc[sb] = fmaf(gs_s[sb * L.cin + i], w, acc[sb]);
        }
    }
#pragma unroll
    for (int sb = 0; sb < kSB; ++sb) part[warp][sb][lane] = acc[sb];
    __syncthreads();
    if (warp < kSB && k < style_dim && b0 + warp < batch) {
        float sum = 0.0f;
#pragma unroll
        for (int w = 0; w < 8; ++w) sum += part[w][warp][lane];
        atomicAdd(g_latent + ((int64_t)(b0 + warp) * n_latent + L.latent_index) * style_dim + k, mod_scale * sum);
    }
}

// ---- Wsq[o,i] = scale^2 * sum_t W[o,i,t]^2 and its gradient  gW[o,i,t] = 2 scale^2 W[o,i,t] * gWsq[o,i] -----------------
__global__ void __launch_bounds__(256)
weight_sq_kernel(float *__restrict__ wsq, const float *__restrict__ w, float scale2, int64_t pairs, int taps)
{
    for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < pairs; p += (int64_t)gridDim.x * 256) {
        const float *src = w + p * taps;
        float acc = 0.0f;
        for (int t = 0; t < taps; ++t) { const float v = __ldg(src + t); acc = fmaf(v, v, acc); }
        wsq[p] = scale2 * acc;
    }
}

__global__ void __launch_bounds__(256)
weight_sq_backward_kernel(float *__restrict__ gw, const float *__restrict__ w, const float *__restrict__ g_wsq, float scale2,
                          uint32_t total, FastDiv div_taps)
{
    // 32-bit indices and a multiply-shift division: the 64-bit `i / taps` of a naive version costs more than the memory traffic
    for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < total; i += gridDim.x * 256u)
        gw[i] = 2.0f * scale2 * __ldg(w + i) * __ldg(g_wsq + div_taps.div(i));
}

// ---- gW[o,i,t] = scale * dwk[o,t,i]  (GEMM layout of the wgrad kernels -> reference layout) ----------------------------
// One CTA per (o, block of 256 input channels): the [taps][256] slab is read with coalesced rows, transposed through shared
// memory and written as one contiguous run of 256 * taps floats.
constexpr int kGL = 256;
__global__ void __launch_bounds__(256)
weight_grad_layout_kernel(float *__restrict__ gw, const float *__restrict__ dwk, float scale, int cout, int cin, int taps)
{
    extern __shared__ float slab[];                       // [taps][kGL + 1]
    const int o = blockIdx.y, i0 = blockIdx.x * kGL;
    const int n = min(kGL, cin - i0);
    for (int t = 0; t < taps; ++t)
        if ((int)threadIdx.x < n) slab[t * (kGL + 1) + threadIdx.x] = __ldg(dwk + ((int64_t)o * taps + t) * cin + i0 + threadIdx.x);
    __syncthreads();
    float *dst = gw + ((int64_t)o * cin + i0) * taps;
    for (int idx = threadIdx.x; idx < n * taps; idx += 256) {
        const int i = idx / taps, t = idx - i * taps;
        dst[idx] = scale * slab[t * (kGL + 1) + i];
    }
}

int fill_table(StyleTable &tab, const sr_style_layer *layers, int n, int &max_cin, int &max_cout)
{
    max_cin = max_cout = 0;
    for (int i = 0; i < n; ++i) {
        tab.l[i] = layers[i];
        if (layers[i].cin > max_cin) max_cin = layers[i].cin;
        if (layers[i].cout > max_cout) max_cout = layers[i].cout;
    }
    return SR_OK;
}

}  // namespace
}  // namespace sr

using namespace sr;

extern "C" int sr_style_scales_forward_f32(const sr_style_layer *layers, int n_layers, const float *latent, int64_t batch,
                                           int64_t n_latent, int64_t style_dim, float mod_scale, float lr_mul, float eps,
                                           void *stream)
{
    SR_REQUIRE(layers && latent && n_layers >= 1 && n_layers <= kMaxStyleLayers, "style_scales: 1..32 layers");
    SR_REQUIRE(batch >= 1 && style_dim >= 1 && n_latent >= 1, "style_scales: empty problem");
    bool any_demod = false;
    for (int i = 0; i < n_layers; ++i) {
        const sr_style_layer &l = layers[i];
        SR_REQUIRE(l.mod_weight && l.mod_bias && l.s && l.cin >= 1, "style_scales: layer %d: null tensor", i);
        SR_REQUIRE(l.latent_index >= 0 && l.latent_index < n_latent, "style_scales: layer %d: latent index out of range", i);
        SR_REQUIRE(!l.wsq || (l.d && l.cout >= 1), "style_scales: layer %d: wsq needs d and cout", i);
        any_demod |= l.wsq != nullptr;
    }
    StyleTable tab;
    int max_cin, max_cout;
    fill_table(tab, layers, n_layers, max_cin, max_cout);
    SR_REQUIRE(kSB * style_dim * 4 <= 48 * 1024 && kSB * max_cin * 4 <= 48 * 1024, "style_scales: style_dim / cin too large");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned gz = (unsigned)((batch + kSB - 1) / kSB);
    style_rows_dot_kernel<0><<<dim3((max_cin + 63) / 64, n_layers, gz), 256, kSB * style_dim * sizeof(float), st>>>(
        tab, latent, (int)batch, (int)n_latent, (int)style_dim, mod_scale, lr_mul, eps);
    count_launch();
    if (any_demod) {
        style_rows_dot_kernel<1><<<dim3((max_cout + 63) / 64, n_layers, gz), 256, kSB * max_cin * sizeof(float), st>>>(
            tab, latent, (int)batch, (int)n_latent, (int)style_dim, mod_scale, lr_mul, eps);
        count_launch();
    }
    return check_launch("sr_style_scales_forward_f32");
}

extern "C" int sr_style_scales_backward_f32(const sr_style_layer *layers, int n_layers, const float *latent, float *g_latent,
                                            int64_t batch, int64_t n_latent, int64_t style_dim, float mod_scale, float lr_mul,
                                            void *stream)
{
    SR_REQUIRE(layers && latent && g_latent && n_layers >= 1 && n_layers <= kMaxStyleLayers, "style_scales_backward: 1..32 layers");
    for (int i = 0; i < n_layers; ++i) {
        const sr_style_layer &l = layers[i];
        SR_REQUIRE(l.mod_weight && l.s && l.gs_total && l.g_mod_weight && l.g_mod_bias, "style_scales_backward: layer %d: null tensor", i);
        SR_REQUIRE(!(l.wsq && l.g_d) || (l.d && l.du && l.g_wsq), "style_scales_backward: layer %d: demodulation needs d, du, g_wsq", i);
    }
    StyleTable tab;
    int max_cin, max_cout;
    fill_table(tab, layers, n_layers, max_cin, max_cout);
    SR_REQUIRE(kSB * max_cin * 4 <= 48 * 1024 && kSB * (max_cout > 0 ? max_cout : 1) * 4 <= 48 * 1024,
               "style_scales_backward: cin / cout too large");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(g_latent, 0, sizeof(float) * (size_t)(batch * n_latent * style_dim), st);
    if (e != cudaSuccess) { set_error("style_scales_backward: memset: %s", cudaGetErrorString(e)); return (int)e; }
    const unsigned gz = (unsigned)((batch + kSB - 1) / kSB);
    style_bwd_gs_kernel<<<dim3((max_cin + 31) / 32, n_layers, gz), 256, kSB * (max_cout > 0 ? max_cout : 1) * sizeof(float), st>>>(
        tab, (int)batch);
    if (max_cout > 0)
        style_bwd_outer_kernel<0><<<dim3((max_cin + 255) / 256, n_layers, (max_cout + 15) / 16), 256, 0, st>>>(
            tab, latent, (int)batch, (int)n_latent, (int)style_dim, mod_scale, lr_mul);
    style_bwd_outer_kernel<1><<<dim3((unsigned)((style_dim + 255) / 256), n_layers, (max_cin + 15) / 16), 256, 0, st>>>(
        tab, latent, (int)batch, (int)n_latent, (int)style_dim, mod_scale, lr_mul);
    style_bwd_latent_kernel<<<dim3((unsigned)((style_dim + 31) / 32), n_layers, gz), 256, kSB * max_cin * sizeof(float), st>>>(
        tab, g_latent, (int)batch, (int)n_latent, (int)style_dim, mod_scale);
    count_launch(max_cout > 0 ? 4 : 3);
    return check_launch("sr_style_scales_backward_f32");
}

extern "C" int sr_weight_sq_f32(float *wsq, const float *w, float scale, int64_t cout, int64_t cin, int taps, void *stream)
{
    SR_REQUIRE(wsq && w && cout >= 1 && cin >= 1 && taps >= 1, "weight_sq: bad arguments");
    const int64_t pairs = cout * cin;
    int64_t blocks = (pairs + 255) / 256;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    weight_sq_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(wsq, w, scale * scale, pairs, taps);
    count_launch();
    return check_launch("sr_weight_sq_f32");
}

extern "C" int sr_weight_sq_backward_f32(float *gw, const float *w, const float *g_wsq, float scale, int64_t cout, int64_t cin,
                                         int taps, void *stream)
{
    SR_REQUIRE(gw && w && g_wsq && cout >= 1 && cin >= 1 && taps >= 1, "weight_sq_backward: bad arguments");
    const int64_t total = cout * cin * taps;
    SR_REQUIRE(total < 0x7fffffffll, "weight_sq_backward: weight too large");
    int64_t blocks = (total + 255) / 256;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    weight_sq_backward_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(gw, w, g_wsq, scale * scale, (uint32_t)total,
                                                                             FastDiv((uint32_t)taps));
    count_launch();
    return check_launch("sr_weight_sq_backward_f32");
}

extern "C" int sr_weight_grad_layout_f32(float *gw, const float *dwk, float scale, int64_t cout, int64_t cin, int taps,
                                         void *stream)
{
    SR_REQUIRE(gw && dwk && cout >= 1 && cin >= 1 && taps >= 1, "weight_grad_layout: bad arguments");
    SR_REQUIRE(cout <= 65535 && taps <= 32, "weight_grad_layout: cout <= 65535, taps <= 32");
    const dim3 grid((unsigned)((cin + kGL - 1) / kGL), (unsigned)cout);
    weight_grad_layout_kernel<<<grid, 256, sizeof(float) * taps * (kGL + 1), (cudaStream_t)stream>>>(gw, dwk, scale, (int)cout,
                                                                                                (int)cin, taps);
    count_launch();
    return check_launch("sr_weight_grad_layout_f32");
}
