// The Discriminator's stem on B200: ConvLayer(3, C, 1) = EqualConv2d 1x1 + bias + FusedLeakyReLU (reference model.py:303,
// layers.py:341-378) on [N,3,H,W] images, output channels-last [N,H,W,C] for the tensor-core ResBlocks that follow.
// Three input channels make this a bandwidth pass, not a GEMM: the reference runs a cuDNN convolution, a bias add and the
// fused_bias_act kernel (three passes over the 537 MB output at batch 16, 256^2, C = 128) and, backward, the activation
// backward, a bias reduction, and two [C, N*H*W] x [N*H*W, 3] products that cuBLAS serves with a SIMT kernel (1.5 ms each).
//   forward :  y[p,o] = lrelu(s * sum_i w[o,i] x[p,i] + b_conv[o] + b_act[o]) * gain            one write of y
//   backward:  gp = gy * gain * (pre > 0 ? 1 : alpha) with pre RECOMPUTED from x (3 FMAs), dw = s * sum_p gp x, db = sum_p gp,
//              dx[p,i] = s * sum_o gp[p,o] w[o,i] (optional)                                       one read of gy
#include "common.cuh"

namespace sr {
namespace {

constexpr int NT = 256;
constexpr int CI = 3;

struct StemGeom {
    int64_t pixels, hw;          // N*H*W (< 2^31), H*W
    FastDiv div_hw;
    int c4;                      // C / 4
    int x_nhwc;                  // x is [N,H,W,3] (channels_last image) instead of [N,3,H,W] planes
    float scale, alpha, gain;
};

__device__ __forceinline__ void load_x(float (&v)[CI], const float *__restrict__ x, int64_t p, const StemGeom &g)
{
    if (g.x_nhwc) {
#pragma unroll
        for (int i = 0; i < CI; ++i) v[i] = __ldg(x + p * CI + i);
    } else {
        uint32_t n, q;
        g.div_hw.divmod((uint32_t)p, n, q);
#pragma unroll
        for (int i = 0; i < CI; ++i) v[i] = __ldg(x + ((int64_t)n * CI + i) * g.hw + q);
    }
}

// thread = (pixel, channel quad); a warp covers 32 consecutive quads: 512 contiguous output bytes
__global__ void __launch_bounds__(NT)
stem_fwd_kernel(float *__restrict__ y, const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ b_conv,
                const float *__restrict__ b_act, const StemGeom g)
{
    if (NT % g.c4 == 0) {
        // the usual case (C = 64, 128, 256, 512): a thread keeps its channel quad, so weights and biases sit in registers
        const int quad = threadIdx.x % g.c4, c = quad * 4, ppb = NT / g.c4;
        float wv[4][CI], bias[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            bias[j] = (b_conv ? __ldg(b_conv + c + j) : 0.f) + (b_act ? __ldg(b_act + c + j) : 0.f);
#pragma unroll
            for (int i = 0; i < CI; ++i) wv[j][i] = __ldg(w + (c + j) * CI + i) * g.scale;
        }
        for (int64_t p = (int64_t)blockIdx.x * ppb + threadIdx.x / g.c4; p < g.pixels; p += (int64_t)gridDim.x * ppb) {
            float v[CI], o[4];
            load_x(v, x, p, g);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float a = bias[j];
#pragma unroll
                for (int i = 0; i < CI; ++i) a = fmaf(wv[j][i], v[i], a);
                o[j] = ((a > 0.f) ? a : a * g.alpha) * g.gain;
            }
            st_stream4(y + (p * g.c4 + quad) * 4, make_float4(o[0], o[1], o[2], o[3]));
        }
        return;
    }
    const int64_t total = g.pixels * g.c4;
    for (int64_t idx = (int64_t)blockIdx.x * NT + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * NT) {
        const int64_t p = idx / g.c4;
        const int c = (int)(idx - p * g.c4) * 4;
        float v[CI];
        load_x(v, x, p, g);
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float a = (b_conv ? __ldg(b_conv + c + j) : 0.f) + (b_act ? __ldg(b_act + c + j) : 0.f);
            float d = 0.0f;
#pragma unroll
            for (int i = 0; i < CI; ++i) d = fmaf(__ldg(w + (c + j) * CI + i), v[i], d);
            a = fmaf(d, g.scale, a);
            o[j] = ((a > 0.f) ? a : a * g.alpha) * g.gain;
        }
        st_stream4(y + idx * 4, make_float4(o[0], o[1], o[2], o[3]));
    }
}

// grads = [dw (C*3) | db (C)], zero-filled by the launcher.  A warp owns one pixel per iteration when C = 128 (32 quads):
// the dx reduction over channels is a warp reduction; parameter gradients accumulate in registers over the thread's
// pixels and are combined through shared memory, one atomic per (CTA, element).
template <bool WANT_DX>
__global__ void __launch_bounds__(NT)
stem_bwd_kernel(float *__restrict__ grads, float *__restrict__ dx, const float *__restrict__ gy, const float *__restrict__ x,
                const float *__restrict__ w, const float *__restrict__ b_conv, const float *__restrict__ b_act, const StemGeom g)
{
    extern __shared__ float red[];                           // [NT / 32 warps][C * 4] partial dw / db
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps_total = gridDim.x * (NT / 32);
    const int quads_per_pass = 32;                            // a warp covers 32 quads (128 channels) of one pixel per pass
    const int passes = (g.c4 + quads_per_pass - 1) / quads_per_pass;
    for (int pass = 0; pass < passes; ++pass) {
        const int quad = pass * quads_per_pass + lane;
        const bool live = quad < g.c4;
        const int c = quad * 4;
        float wv[4][CI], bias[4], dwv[4][CI], dbv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            bias[j] = live ? (b_conv ? __ldg(b_conv + c + j) : 0.f) + (b_act ? __ldg(b_act + c + j) : 0.f) : 0.f;
            dbv[j] = 0.0f;
#pragma unroll
            for (int i = 0; i < CI; ++i) { wv[j][i] = live ? __ldg(w + (c + j) * CI + i) * g.scale : 0.f; dwv[j][i] = 0.0f; }
        }
        constexpr int U = 2;                                  // pixels in flight per warp (one 512-byte gradient row each)
        for (int64_t p0 = (int64_t)blockIdx.x * (NT / 32) + warp; p0 < g.pixels; p0 += (int64_t)U * nwarps_total) {
          float vu[U][CI];
          float4 gu[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
              const int64_t p = p0 + (int64_t)u * nwarps_total;
              gu[u] = make_float4(0.f, 0.f, 0.f, 0.f);
              vu[u][0] = vu[u][1] = vu[u][2] = 0.f;
              if (p < g.pixels) {
                  load_x(vu[u], x, p, g);
                  if (live) gu[u] = ld_stream4(gy + (p * g.c4 + quad) * 4);
              }
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int64_t p = p0 + (int64_t)u * nwarps_total;
            if (p >= g.pixels) break;
            const float (&v)[CI] = vu[u];
            const float gg[4] = {gu[u].x, gu[u].y, gu[u].z, gu[u].w};
            float dxv[CI] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float a = bias[j];
#pragma unroll
                for (int i = 0; i < CI; ++i) a = fmaf(wv[j][i], v[i], a);
                const float gp = gg[j] * g.gain * ((a > 0.f) ? 1.0f : g.alpha);
                dbv[j] += gp;
#pragma unroll
                for (int i = 0; i < CI; ++i) {
                    dwv[j][i] = fmaf(gp, v[i], dwv[j][i]);
                    if (WANT_DX) dxv[i] = fmaf(gp, wv[j][i], dxv[i]);
                }
            }
            if (WANT_DX) {
#pragma unroll
                for (int i = 0; i < CI; ++i) dxv[i] = warp_sum(dxv[i]);
                if (lane < CI) {
                    const float r = lane == 0 ? dxv[0] : (lane == 1 ? dxv[1] : dxv[2]);
                    if (g.x_nhwc) {
                        if (pass == 0) dx[p * CI + lane] = r; else dx[p * CI + lane] += r;
                    } else {
                        uint32_t n, q;
                        g.div_hw.divmod((uint32_t)p, n, q);
                        float *d = dx + ((int64_t)n * CI + lane) * g.hw + q;
                        if (pass == 0) *d = r; else *d += r;
                    }
                }
            }
          }
        }
        // combine the warps of the CTA, then one atomic per element
        float *mine = red + warp * (g.c4 * 16);
        if (live) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
                for (int i = 0; i < CI; ++i) mine[(c + j) * CI + i] = dwv[j][i] * g.scale;
                mine[g.c4 * 12 + c + j] = dbv[j];
            }
        }
        __syncthreads();
        const int lo = pass * quads_per_pass * 4, hi = min(g.c4 * 4, lo + quads_per_pass * 4);
        for (int e = threadIdx.x; e < (hi - lo) * 4; e += NT) {
            // elements of this pass: dw rows lo..hi (3 each) and db lo..hi
            const int ch = lo + e / 4, k = e % 4;
            const int off = k < CI ? ch * CI + k : g.c4 * 12 + ch;
            float s = 0.0f;
            for (int wi = 0; wi < NT / 32; ++wi) s += red[wi * (g.c4 * 16) + off];
            atomicAdd(grads + off, s);
        }
        __syncthreads();
    }
}

int check(const void *a, const void *b, const void *c, int64_t n, int cin, int64_t cout, int64_t h, int64_t w, const char *what)
{
    SR_REQUIRE(a && b && c, "%s: null pointer", what);
    SR_REQUIRE(cin == CI, "%s: built for 3 input channels (got %d)", what, cin);
    SR_REQUIRE(cout >= 4 && cout % 4 == 0 && cout <= 1024, "%s: output channels must be a multiple of 4, <= 1024", what);
    SR_REQUIRE(n >= 1 && h >= 1 && w >= 1 && n * h * w < (1ll << 31), "%s: bad sizes", what);
    return SR_OK;
}

}  // namespace
}  // namespace sr

using namespace sr;

extern "C" int sr_stem_conv_forward_f32(float *y, const float *x, const float *w, const float *b_conv, const float *b_act,
                                        int64_t batch, int cin, int64_t cout, int64_t h, int64_t wd, int x_channels_last,
                                        float alpha, float gain, void *stream)
{
    int rc = check(y, x, w, batch, cin, cout, h, wd, "stem_conv_forward");
    if (rc != SR_OK) return rc;
    SR_REQUIRE((reinterpret_cast<uintptr_t>(y) & 15u) == 0, "stem_conv_forward: 16-byte alignment");
    StemGeom g = {batch * h * wd, h * wd, FastDiv((uint32_t)(h * wd)), (int)(cout / 4), x_channels_last, 1.0f / sqrtf((float)cin), alpha, gain};
    int64_t blocks = (g.pixels * g.c4 + NT - 1) / NT;
    if (blocks > (int64_t)kNumSMs * 32) blocks = (int64_t)kNumSMs * 32;
    stem_fwd_kernel<<<(unsigned)blocks, NT, 0, (cudaStream_t)stream>>>(y, x, w, b_conv, b_act, g);
    count_launch();
    return check_launch("stem_conv_forward");
}

extern "C" int sr_stem_conv_backward_f32(float *grads, float *dx, const float *gy, const float *x, const float *w,
                                         const float *b_conv, const float *b_act, int64_t batch, int cin, int64_t cout,
                                         int64_t h, int64_t wd, int x_channels_last, float alpha, float gain, void *stream)
{
    int rc = check(grads, gy, x, batch, cin, cout, h, wd, "stem_conv_backward");
    if (rc != SR_OK) return rc;
    SR_REQUIRE(w && (reinterpret_cast<uintptr_t>(gy) & 15u) == 0, "stem_conv_backward: null weight or unaligned gradient");
    StemGeom g = {batch * h * wd, h * wd, FastDiv((uint32_t)(h * wd)), (int)(cout / 4), x_channels_last, 1.0f / sqrtf((float)cin), alpha, gain};
    cudaStream_t st = (cudaStream_t)stream;
    if (cudaMemsetAsync(grads, 0, sizeof(float) * cout * 4, st) != cudaSuccess) return check_launch("stem_conv_backward (memset)");
    const size_t smem = sizeof(float) * (NT / 32) * cout * 4;
    const int grid = (int)(g.pixels < 4ll * kNumSMs * 8 ? (g.pixels + 7) / 8 : 4ll * kNumSMs);
    // dynamic shared memory: 8 warps x [C x 4] partials = 16 KB at C = 128, 128 KB at the largest supported C = 1024
    constexpr int kMaxSmem = (NT / 32) * 1024 * 4 * (int)sizeof(float);
    static bool conf = false;
    if (!conf) {
        if (cudaFuncSetAttribute(stem_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem) != cudaSuccess ||
            cudaFuncSetAttribute(stem_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem) != cudaSuccess)
            return check_launch("stem_conv_backward (shared memory)");
        conf = true;
    }
    if (dx) stem_bwd_kernel<true><<<grid, NT, smem, st>>>(grads, dx, gy, x, w, b_conv, b_act, g);
    else stem_bwd_kernel<false><<<grid, NT, smem, st>>>(grads, dx, gy, x, w, b_conv, b_act, g);
    count_launch();
    return check_launch("stem_conv_backward");
}
