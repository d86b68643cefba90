// Fused elementwise + reduction passes around the tensor-core GEMMs of a StyledConv block (NHWC, float32).
//
// They replace the chains of separate full-tensor passes the reference runs in PyTorch around its grouped
// conv (reference model.py:26-32, layers.py:296-332, op/fused_act.py:27-38): every kernel here reads each
// activation-sized tensor exactly once and produces, in the same pass, both the tensor the next GEMM
// consumes (already scaled and rounded to tf32) and the per-channel / per-(sample,channel) reductions the
// parameter gradients need.  HBM-bound: 128-bit accesses, consecutive lanes = consecutive channel quads
// (a warp covers one contiguous 512-byte run), reductions in registers -> shared memory -> one global
// atomic per (CTA, channel).
#include "common.cuh"
#include <stdlib.h>

namespace sr {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float round_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}
__device__ __forceinline__ float4 f4_mul(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ void f4_fma(float4 &acc, float4 a, float4 b) {
    acc.x = fmaf(a.x, b.x, acc.x); acc.y = fmaf(a.y, b.y, acc.y); acc.z = fmaf(a.z, b.z, acc.z); acc.w = fmaf(a.w, b.w, acc.w);
}
__device__ __forceinline__ void f4_add(float4 &acc, float4 a) { acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w; }

// Block-level reduction over the `lanes_p` pixel lanes of a CTA (threads with equal channel quad), then one
// atomicAdd per channel.  smem: kThreads float4.
__device__ __forceinline__ void block_reduce_quads(float4 v, float4 *smem, int c4, int pl, int C4, int lanes_p, float *dst)
{
    __syncthreads();
    smem[pl * C4 + c4] = v;
    __syncthreads();
    if (pl == 0) {
        float4 s = smem[c4];
        for (int i = 1; i < lanes_p; ++i) f4_add(s, smem[i * C4 + c4]);
        atomicAdd(dst + 4 * c4 + 0, s.x); atomicAdd(dst + 4 * c4 + 1, s.y);
        atomicAdd(dst + 4 * c4 + 2, s.z); atomicAdd(dst + 4 * c4 + 3, s.w);
    }
}

struct PrologueParams {
    float *ga;                    // out: tf32(g_pre * d[b,c]) (or g_pre when d == nullptr)
    float *g_bias;                // [C]      += sum g_pre
    float *g_noise_w;             // [1]      += sum g_pre * noise
    float *e;                     // [B,C]    += sum g_pre * (pre_activation - noise_w*noise - bias)   (nullptr: skip)
    const float *gy, *y, *noise, *noise_w, *bias, *d;
    // chained gradient sources (all optional): g_total = gy + gxs * s_next[b,c] + sum_k g_rgb[b,p,k] * rgb_w[b,k,c]
    const float *gxs, *s_next, *g_rgb, *rgb_w;
    float *ds_next;               // [B,C]   += sum_p gxs * y
    float *d_rgb_w;               // [B,3,C] += sum_p g_rgb[b,p,k] * y
    long long noise_bstride;
    int pixels, C4, chunks_per_image, pix_per_chunk;
    float alpha, gain;
    // StyledMapConv (reference model.py:48-52): t = conv_d * map0[b,p] + map1[b,p] + noise_w * noise + bias.  With a
    // stylemap `y` holds conv_d -- the demodulated (and, in an up-sampling block, filtered) conv output BEFORE the map
    // affine, which is what the forward kernels store for these blocks -- and the activated output is rebuilt here.  The
    // gradient handed to the GEMMs is gp * map0 (* d), e = sum gp * (conv_d * map0), and per pixel
    //   g_map1 = sum_c gp,   g_map0 = sum_c gp * conv_d          (no division: map0 may be exactly 0).
    int ga_bf16;                  // ga is a bfloat16 tensor (the 16-bit operand of the dgrad / wgrad GEMMs; only with d)
    const float *stylemap;        // [B, 2, pixels] planes (batch stride map_bstride, plane stride = pixels) or nullptr
    long long map_bstride;
    float *g_map;                 // [B, 2, pixels] contiguous, += (zeroed by the caller)
};

// grid.x = B * chunks_per_image.  MAP = StyledMapConv variant (compile-time: the plain path keeps its register budget).
// SPEC < 0: the optional inputs are run-time branches (one kernel for every call, 117 registers = 2 CTAs per SM).
// SPEC >= 0: bit mask of the inputs that are present (1 gy, 2 gxs, 4 ToRGB, 8 e, 16 noise, 32 d), compiled in for the
// combinations the chained generator produces -- the up-sampling blocks' combination needs 64 registers (4 CTAs per SM,
// MINB), the ToRGB ones 80 (profiles/r1_prologue_specialisation_ptxas.md).  SR_PROLOGUE_SPEC=0 forces the run-time kernel.
constexpr int kSpecGy = 1, kSpecGxs = 2, kSpecRgb = 4, kSpecE = 8, kSpecNoise = 16, kSpecD = 32;

template <bool MAP, int SPEC, int MINB, int U = 2>      // U = pixels per loop iteration (the ToRGB variants take 1: fewer temporaries)
__global__ void __launch_bounds__(kThreads, MINB)
styled_bwd_prologue_kernel(const PrologueParams p)
{
    constexpr bool S = SPEC >= 0;
    const bool has_gy = S ? ((SPEC & kSpecGy) != 0) : (p.gy != nullptr);
    const bool has_gxs = S ? ((SPEC & kSpecGxs) != 0) : (p.gxs != nullptr);
    const bool has_rgb = S ? ((SPEC & kSpecRgb) != 0) : (p.g_rgb != nullptr);
    const bool has_e = S ? ((SPEC & kSpecE) != 0) : (p.e != nullptr);
    const bool has_noise = S ? ((SPEC & kSpecNoise) != 0) : (p.noise != nullptr);
    const bool has_d = S ? ((SPEC & kSpecD) != 0) : (p.d != nullptr);
    __shared__ float4 s_red[kThreads];
    __shared__ float s_nw[kThreads / 32];
    const int C4 = p.C4, lanes_p = kThreads / C4;
    const int c4 = threadIdx.x % C4, pl = threadIdx.x / C4;
    const int b = blockIdx.x / p.chunks_per_image, chunk = blockIdx.x % p.chunks_per_image;
    const int p0 = chunk * p.pix_per_chunk, p1 = min(p.pixels, p0 + p.pix_per_chunk);
    const float4 *gy = has_gy ? reinterpret_cast<const float4 *>(p.gy) + (long long)b * p.pixels * C4 + c4 : nullptr;
    const float4 *gxs = has_gxs ? reinterpret_cast<const float4 *>(p.gxs) + (long long)b * p.pixels * C4 + c4 : nullptr;
    const float *grgb = has_rgb ? p.g_rgb + (long long)b * p.pixels * 3 : nullptr;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 sn = has_gxs ? __ldg(reinterpret_cast<const float4 *>(p.s_next) + (long long)b * C4 + c4) : zero4;
    float4 wrgb[3] = {zero4, zero4, zero4};
    if (has_rgb) {
#pragma unroll
        for (int k = 0; k < 3; ++k) wrgb[k] = __ldg(reinterpret_cast<const float4 *>(p.rgb_w) + ((long long)b * 3 + k) * C4 + c4);
    }
    float4 a_ds = zero4, a_wb[3] = {zero4, zero4, zero4};
    const float4 *y = reinterpret_cast<const float4 *>(p.y) + (long long)b * p.pixels * C4 + c4;
    float4 *ga = reinterpret_cast<float4 *>(p.ga) + (long long)b * p.pixels * C4 + c4;
    const float *nz = has_noise ? p.noise + (long long)b * p.noise_bstride : nullptr;
    const float nw = has_noise ? __ldg(p.noise_w) : 0.0f;
    const float4 bias = p.bias ? __ldg(reinterpret_cast<const float4 *>(p.bias) + c4) : zero4;
    const float4 dd = has_d ? __ldg(reinterpret_cast<const float4 *>(p.d) + (long long)b * C4 + c4) : make_float4(1.f, 1.f, 1.f, 1.f);
    const float ipos = 1.0f / p.gain, ineg = 1.0f / (p.gain * p.alpha);
    float4 a_bias = zero4, a_e = zero4;
    float a_nw = 0.0f;
    // U pixels per iteration: U times the loads in flight per thread (the pass is pure streaming)
    for (int px0 = p0 + pl; px0 < p1; px0 += U * lanes_p) {
        const int pxs[2] = {px0, px0 + lanes_p};
        float4 yy2[U], g2[U], gx2[U];
        float gk2[U][3], n2[U], m02[U], m12[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            ok[u] = pxs[u] < p1;
            const long long off = (long long)(ok[u] ? pxs[u] : px0) * C4;
            yy2[u] = __ldg(y + off);
            g2[u] = has_gy ? ld_stream4(reinterpret_cast<const float *>(gy + off)) : zero4;
            gx2[u] = has_gxs ? ld_stream4(reinterpret_cast<const float *>(gxs + off)) : zero4;
            const long long pp = ok[u] ? pxs[u] : px0;
#pragma unroll
            for (int k = 0; k < 3; ++k) gk2[u][k] = has_rgb ? __ldg(grgb + pp * 3 + k) : 0.0f;
            n2[u] = has_noise ? __ldg(nz + pp) : 0.0f;
            m02[u] = MAP ? __ldg(p.stylemap + (long long)b * p.map_bstride + pp) : 1.0f;
            m12[u] = MAP ? __ldg(p.stylemap + (long long)b * p.map_bstride + p.pixels + pp) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!ok[u]) continue;
            const int px = pxs[u];
            const float n = n2[u];
            // MAP: the saved tensor is cd = the demodulated conv output BEFORE the map affine (see PrologueParams); the
            // activated output is rebuilt from it with the forward's own expression.  Otherwise the saved tensor is y.
            float4 yy = yy2[u], cd = yy2[u];
            if (MAP) {
                const float sh = nw * n + m12[u];
                float t;
                t = fmaf(cd.x, m02[u], sh + bias.x); yy.x = ((t > 0.f) ? t : t * p.alpha) * p.gain;
                t = fmaf(cd.y, m02[u], sh + bias.y); yy.y = ((t > 0.f) ? t : t * p.alpha) * p.gain;
                t = fmaf(cd.z, m02[u], sh + bias.z); yy.z = ((t > 0.f) ? t : t * p.alpha) * p.gain;
                t = fmaf(cd.w, m02[u], sh + bias.w); yy.w = ((t > 0.f) ? t : t * p.alpha) * p.gain;
            }
            float4 g = g2[u];
            if (has_gxs) { f4_fma(g, gx2[u], sn); f4_fma(a_ds, gx2[u], yy); }
            if (has_rgb) {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float4 g4 = make_float4(gk2[u][k], gk2[u][k], gk2[u][k], gk2[u][k]);
                    f4_fma(g, g4, wrgb[k]);
                    f4_fma(a_wb[k], g4, yy);
                }
            }
            float4 gp;
            // reference op/fused_bias_act_kernel.cu:31: (ref > 0 ? g : g*alpha) * scale
            gp.x = ((yy.x > 0.f) ? g.x : g.x * p.alpha) * p.gain; gp.y = ((yy.y > 0.f) ? g.y : g.y * p.alpha) * p.gain;
            gp.z = ((yy.z > 0.f) ? g.z : g.z * p.alpha) * p.gain; gp.w = ((yy.w > 0.f) ? g.w : g.w * p.alpha) * p.gain;
            f4_add(a_bias, gp);
            a_nw += ((gp.x + gp.y) + (gp.z + gp.w)) * n;
            if (MAP) {
                // e = sum gp * (cd * map0); per pixel g_map0 = sum_c gp * cd and g_map1 = sum_c gp -- no division by the
                // map (background pixels of the rasterised normal map have map0 == 0 exactly at default init).
                // The C4 threads of a pixel are consecutive (whole warps when C4 >= 32, aligned sub-warp groups when C4
                // is a smaller power of two).
                if (has_e) {
                    const float4 cm = make_float4(cd.x * m02[u], cd.y * m02[u], cd.z * m02[u], cd.w * m02[u]);
                    f4_fma(a_e, gp, cm);
                }
                float s1 = (gp.x + gp.y) + (gp.z + gp.w);
                float s0 = fmaf(gp.x, cd.x, fmaf(gp.y, cd.y, fmaf(gp.z, cd.z, gp.w * cd.w)));
                const int width = C4 < 32 ? C4 : 32;
                const unsigned amask = __activemask();          // sub-warp pixel groups may be masked off independently
                for (int o_ = width >> 1; o_ > 0; o_ >>= 1) {
                    s1 += __shfl_xor_sync(amask, s1, o_);
                    s0 += __shfl_xor_sync(amask, s0, o_);
                }
                if ((threadIdx.x & (width - 1)) == 0) {
                    float *gm = p.g_map + (long long)b * 2 * p.pixels + px;
                    atomicAdd(gm, s0);
                    atomicAdd(gm + p.pixels, s1);
                }
                gp.x *= m02[u]; gp.y *= m02[u]; gp.z *= m02[u]; gp.w *= m02[u];
            } else if (has_e) {
                // the pre-activation is recoverable from y (gain, alpha > 0): t = y / gain or y / (gain * alpha)
                float4 uu;
                uu.x = yy.x * ((yy.x > 0.f) ? ipos : ineg); uu.y = yy.y * ((yy.y > 0.f) ? ipos : ineg);
                uu.z = yy.z * ((yy.z > 0.f) ? ipos : ineg); uu.w = yy.w * ((yy.w > 0.f) ? ipos : ineg);
                const float sh = nw * n;
                uu.x -= sh + bias.x; uu.y -= sh + bias.y; uu.z -= sh + bias.z; uu.w -= sh + bias.w;
                f4_fma(a_e, gp, uu);
            }
            float4 o = f4_mul(gp, dd);
            if (has_d && p.ga_bf16) {
                uint32_t lo, hi;
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(o.y), "f"(o.x));
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(o.w), "f"(o.z));
                reinterpret_cast<uint2 *>(p.ga)[((long long)b * p.pixels + px) * C4 + c4] = make_uint2(lo, hi);
            } else {
                if (has_d) o = make_float4(round_tf32(o.x), round_tf32(o.y), round_tf32(o.z), round_tf32(o.w));
                ga[(long long)px * C4] = o;
            }
        }
    }
    block_reduce_quads(a_bias, s_red, c4, pl, C4, lanes_p, p.g_bias);
    if (has_e) block_reduce_quads(a_e, s_red, c4, pl, C4, lanes_p, p.e + (long long)b * C4 * 4);
    if (has_gxs) block_reduce_quads(a_ds, s_red, c4, pl, C4, lanes_p, p.ds_next + (long long)b * C4 * 4);
    if (has_rgb) {
#pragma unroll
        for (int k = 0; k < 3; ++k)
            block_reduce_quads(a_wb[k], s_red, c4, pl, C4, lanes_p, p.d_rgb_w + ((long long)b * 3 + k) * C4 * 4);
    }
    if (has_noise) {
        a_nw = warp_sum(a_nw);
        if ((threadIdx.x & 31) == 0) s_nw[threadIdx.x >> 5] = a_nw;
        __syncthreads();
        if (threadIdx.x == 0) {
            float s = 0.f;
            for (int i = 0; i < kThreads / 32; ++i) s += s_nw[i];
            atomicAdd(p.g_noise_w, s);
        }
    }
}

struct ScaleDotParams {
    float *out;                   // a * scale[b,c]  (rounded to tf32 when round_out)
    float *dot;                   // [B,C] += sum_p a * other
    const float *a, *other, *scale;
    int pixels, C4, chunks_per_image, pix_per_chunk, round_out;
};

__global__ void __launch_bounds__(kThreads)
scale_dot_kernel(const ScaleDotParams p)
{
    __shared__ float4 s_red[kThreads];
    const int C4 = p.C4, lanes_p = kThreads / C4;
    const int c4 = threadIdx.x % C4, pl = threadIdx.x / C4;
    const int b = blockIdx.x / p.chunks_per_image, chunk = blockIdx.x % p.chunks_per_image;
    const int p0 = chunk * p.pix_per_chunk, p1 = min(p.pixels, p0 + p.pix_per_chunk);
    const long long base = (long long)b * p.pixels * C4 + c4;
    const float4 *a = reinterpret_cast<const float4 *>(p.a) + base;
    const float4 *o = p.other ? reinterpret_cast<const float4 *>(p.other) + base : nullptr;
    float4 *out = p.out ? reinterpret_cast<float4 *>(p.out) + base : nullptr;
    const float4 sc = p.scale ? __ldg(reinterpret_cast<const float4 *>(p.scale) + (long long)b * C4 + c4)
                              : make_float4(1.f, 1.f, 1.f, 1.f);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int px = p0 + pl; px < p1; px += lanes_p) {
        const float4 av = ld_stream4(reinterpret_cast<const float *>(a + (long long)px * C4));
        if (o) f4_fma(acc, av, ld_stream4(reinterpret_cast<const float *>(o + (long long)px * C4)));
        if (out) {
            float4 r = f4_mul(av, sc);
            if (p.round_out) r = make_float4(round_tf32(r.x), round_tf32(r.y), round_tf32(r.z), round_tf32(r.w));
            out[(long long)px * C4] = r;
        }
    }
    if (p.dot) block_reduce_quads(acc, s_red, c4, pl, C4, lanes_p, p.dot + (long long)b * C4 * 4);
}

int pick_chunks(int64_t batch, int64_t pixels, int C4, int *pix_per_chunk) {
    // about 4 CTAs per SM in flight, at least 64 pixels per pixel lane to amortise the final reduction
    const int lanes_p = kThreads / C4;
    int64_t want_ctas = (int64_t)kNumSMs * 8;
    int64_t chunks = (want_ctas + batch - 1) / batch;
    int64_t min_pix = (int64_t)lanes_p * 32;
    int64_t max_chunks = (pixels + min_pix - 1) / min_pix;
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    *pix_per_chunk = (int)((pixels + chunks - 1) / chunks);
    return (int)((pixels + *pix_per_chunk - 1) / *pix_per_chunk);
}

bool ok_channels(int64_t c) { return c % 4 == 0 && c / 4 <= kThreads && kThreads % (c / 4) == 0; }
bool al16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace
}  // namespace sr

using namespace sr;

static int styled_bwd_prologue_any(float *ga, float *g_bias, float *g_noise_w, float *e, float *ds_next,
                                   float *d_rgb_weight, const float *gy, const float *gxs, const float *s_next,
                                   const float *g_rgb, const float *rgb_weight, const float *y, const float *noise,
                                   int64_t noise_batch_stride, const float *noise_weight, const float *bias,
                                   const float *d, int64_t batch, int64_t pixels, int64_t channels, float alpha,
                                   float gain, const float *stylemap, int64_t stylemap_batch_stride, float *g_stylemap,
                                   void *stream, int ga_bf16)
{
    SR_REQUIRE(!stylemap || g_stylemap, "styled_bwd_prologue: stylemap needs g_stylemap");
    SR_REQUIRE(!stylemap || (channels / 4 >= 32 || ((channels / 4) & (channels / 4 - 1)) == 0),
               "styled_bwd_prologue: with a stylemap channels/4 must be >= 32 or a power of two");
    SR_REQUIRE(ga && g_bias && y && (gy || gxs || g_rgb), "styled_bwd_prologue: null pointer / no gradient source");
    SR_REQUIRE(ok_channels(channels), "styled_bwd_prologue: channels must be 4*k with k dividing 256 (got %lld)", (long long)channels);
    SR_REQUIRE(al16(ga) && (!gy || al16(gy)) && al16(y) && (!bias || al16(bias)) && (!d || al16(d)) && (!gxs || al16(gxs)) &&
               (!s_next || al16(s_next)) && (!rgb_weight || al16(rgb_weight)), "styled_bwd_prologue: 16-byte alignment");
    SR_REQUIRE(!noise || (noise_weight && g_noise_w), "styled_bwd_prologue: noise needs its weight and gradient slot");
    SR_REQUIRE(!gxs || (s_next && ds_next), "styled_bwd_prologue: gxs needs s_next and ds_next");
    SR_REQUIRE(!g_rgb || (rgb_weight && d_rgb_weight), "styled_bwd_prologue: g_rgb needs rgb_weight and d_rgb_weight");
    // alpha == 0 (plain ReLU, e.g. the VGG-shaped perceptual stack of the inversion loop) is fine as long as nothing has to
    // recover the pre-activation from y (e) or rebuild y from the pre-map value (stylemap)
    SR_REQUIRE(gain > 0 && (alpha > 0 || (alpha == 0 && !e && !stylemap)),
               "styled_bwd_prologue: gain must be positive and alpha positive (alpha == 0 only without e / stylemap)");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t er = cudaMemsetAsync(g_bias, 0, sizeof(float) * (size_t)channels, st);
    if (er == cudaSuccess && g_noise_w) er = cudaMemsetAsync(g_noise_w, 0, sizeof(float), st);
    if (er == cudaSuccess && e) er = cudaMemsetAsync(e, 0, sizeof(float) * (size_t)(batch * channels), st);
    if (er == cudaSuccess && gxs) er = cudaMemsetAsync(ds_next, 0, sizeof(float) * (size_t)(batch * channels), st);
    if (er == cudaSuccess && g_rgb) er = cudaMemsetAsync(d_rgb_weight, 0, sizeof(float) * (size_t)(batch * 3 * channels), st);
    if (er == cudaSuccess && stylemap) er = cudaMemsetAsync(g_stylemap, 0, sizeof(float) * (size_t)(batch * 2 * pixels), st);
    if (er != cudaSuccess) { set_error("styled_bwd_prologue: memset: %s", cudaGetErrorString(er)); return (int)er; }
    if (batch == 0 || pixels == 0) return SR_OK;
    PrologueParams p;
    p.ga = ga; p.g_bias = g_bias; p.g_noise_w = g_noise_w; p.e = e; p.gy = gy; p.y = y;
    p.noise = noise; p.noise_w = noise_weight; p.bias = bias; p.d = d; p.noise_bstride = noise_batch_stride;
    p.gxs = gxs; p.s_next = s_next; p.g_rgb = g_rgb; p.rgb_w = rgb_weight; p.ds_next = ds_next; p.d_rgb_w = d_rgb_weight;
    p.pixels = (int)pixels; p.C4 = (int)(channels / 4);
    p.chunks_per_image = pick_chunks(batch, pixels, p.C4, &p.pix_per_chunk);
    p.alpha = alpha; p.gain = gain;
    p.stylemap = stylemap; p.map_bstride = stylemap_batch_stride; p.g_map = g_stylemap;
    p.ga_bf16 = (ga_bf16 && d) ? 1 : 0;
    const unsigned nb = (unsigned)(batch * p.chunks_per_image);
    // kernels specialised on the inputs present (see styled_bwd_prologue_kernel); SR_PROLOGUE_SPEC=0: run-time checks only
    const char *spec_env = getenv("SR_PROLOGUE_SPEC");          // read per call (A/B runs in one process)
    const int spec = (gy ? kSpecGy : 0) | (gxs ? kSpecGxs : 0) | (g_rgb ? kSpecRgb : 0) | (e ? kSpecE : 0) |
                     (noise ? kSpecNoise : 0) | (d ? kSpecD : 0);
    bool launched = false;
    // measured (B200, B = 32 generator step, profiles/r2_prologue_spec.md): specialising the combinations WITHOUT the ToRGB
    // gradient (64 registers, no spill) takes the pass from 2.59 to 2.46 ms; the ToRGB combinations spilled at 80 registers
    // with two pixels per iteration and lost the gain (2.55 ms) -- they now take one pixel per iteration (no spill, 3 CTAs
    // per SM; ncu: the run-time kernel they replace sits at 45 % of DRAM, the spill-free specialised ones at 79 %).
    // Measured again: 2.48 ms with them, 2.44 ms without (profiles/r2_prologue_spec.md) -- the specialised ToRGB kernels still
    // spill 16-36 B at 80 registers, so they stay opt-in (SR_PROLOGUE_SPEC=1).
    const bool spec_rgb = spec_env && spec_env[0] == '1';
    // ToRGB combinations at up to 128 registers (106 used, no spill), 2 pixels per iteration: 2.475 -> 2.438 ms per step (default;
    // SR_PROLOGUE_SPEC=3 = the run-time kernel for them, as before)
    const bool spec_rgb2 = !spec_env || spec_env[0] == '2';
    if (spec_rgb2 && !stylemap && g_rgb) {
        launched = true;
        switch (spec) {
        case kSpecGxs | kSpecRgb | kSpecE | kSpecNoise | kSpecD:
            styled_bwd_prologue_kernel<false, kSpecGxs | kSpecRgb | kSpecE | kSpecNoise | kSpecD, 2, 2><<<nb, kThreads, 0, st>>>(p); break;
        case kSpecGy | kSpecRgb | kSpecE | kSpecNoise | kSpecD:
            styled_bwd_prologue_kernel<false, kSpecGy | kSpecRgb | kSpecE | kSpecNoise | kSpecD, 2, 2><<<nb, kThreads, 0, st>>>(p); break;
        default: launched = false;
        }
    }
    if (!launched && !(spec_env && spec_env[0] == '0') && !stylemap && (spec_rgb || !g_rgb)) {
        launched = true;
        switch (spec) {                                   // the combinations the chained generator produces
        case kSpecGy | kSpecD:                                                   // plain ConvLayer (Discriminator, perceptual stack)
            styled_bwd_prologue_kernel<false, kSpecGy | kSpecD, 4><<<nb, kThreads, 0, st>>>(p); break;
        case kSpecGxs | kSpecE | kSpecNoise:                                     // up-sampling block
            styled_bwd_prologue_kernel<false, kSpecGxs | kSpecE | kSpecNoise, 4><<<nb, kThreads, 0, st>>>(p); break;
        case kSpecGxs | kSpecE | kSpecNoise | kSpecD:                            // plain block without ToRGB
            styled_bwd_prologue_kernel<false, kSpecGxs | kSpecE | kSpecNoise | kSpecD, 4><<<nb, kThreads, 0, st>>>(p); break;
        case kSpecGxs | kSpecRgb | kSpecE | kSpecNoise | kSpecD:                 // plain block with ToRGB
            styled_bwd_prologue_kernel<false, kSpecGxs | kSpecRgb | kSpecE | kSpecNoise | kSpecD, 3, 1><<<nb, kThreads, 0, st>>>(p); break;
        case kSpecGy | kSpecRgb | kSpecE | kSpecNoise | kSpecD:                  // last block (image gradient + ToRGB)
            styled_bwd_prologue_kernel<false, kSpecGy | kSpecRgb | kSpecE | kSpecNoise | kSpecD, 3, 1><<<nb, kThreads, 0, st>>>(p); break;
        default: launched = false;
        }
    }
    // StyledMapConv blocks (GeneratorWithMap): the same four combinations with the style-map arithmetic compiled in (they ran
    // the 117-register run-time kernel at 45 % of DRAM: ncu, profiles/r2_ncu_full_summary.md); SR_PROLOGUE_SPEC=0/3 keeps it
    if (!launched && stylemap && !(spec_env && (spec_env[0] == '0' || spec_env[0] == '3'))) {
        launched = true;
        switch (spec) {
        case kSpecGxs | kSpecE | kSpecNoise:
            styled_bwd_prologue_kernel<true, kSpecGxs | kSpecE | kSpecNoise, 2, 2><<<nb, kThreads, 0, st>>>(p); break;
        case kSpecGxs | kSpecE | kSpecNoise | kSpecD:
            styled_bwd_prologue_kernel<true, kSpecGxs | kSpecE | kSpecNoise | kSpecD, 2, 2><<<nb, kThreads, 0, st>>>(p); break;
        case kSpecGxs | kSpecRgb | kSpecE | kSpecNoise | kSpecD:
            styled_bwd_prologue_kernel<true, kSpecGxs | kSpecRgb | kSpecE | kSpecNoise | kSpecD, 2, 2><<<nb, kThreads, 0, st>>>(p); break;
        case kSpecGy | kSpecRgb | kSpecE | kSpecNoise | kSpecD:
            styled_bwd_prologue_kernel<true, kSpecGy | kSpecRgb | kSpecE | kSpecNoise | kSpecD, 2, 2><<<nb, kThreads, 0, st>>>(p); break;
        default: launched = false;
        }
    }
    if (!launched) {
        if (stylemap) styled_bwd_prologue_kernel<true, -1, 1><<<nb, kThreads, 0, st>>>(p);
        else styled_bwd_prologue_kernel<false, -1, 1><<<nb, kThreads, 0, st>>>(p);
    }
    count_launch();
    return check_launch("sr_styled_bwd_prologue_f32");
}

extern "C" int sr_styled_bwd_prologue3_f32(float *ga, float *g_bias, float *g_noise_w, float *e, float *ds_next,
                                           float *d_rgb_weight, const float *gy, const float *gxs, const float *s_next,
                                           const float *g_rgb, const float *rgb_weight, const float *y, const float *noise,
                                           int64_t noise_batch_stride, const float *noise_weight, const float *bias,
                                           const float *d, int64_t batch, int64_t pixels, int64_t channels, float alpha,
                                           float gain, const float *stylemap, int64_t stylemap_batch_stride, float *g_stylemap,
                                           void *stream)
{
    return styled_bwd_prologue_any(ga, g_bias, g_noise_w, e, ds_next, d_rgb_weight, gy, gxs, s_next, g_rgb, rgb_weight, y, noise,
                                   noise_batch_stride, noise_weight, bias, d, batch, pixels, channels, alpha, gain, stylemap,
                                   stylemap_batch_stride, g_stylemap, stream, 0);
}
// bf16 operand form: with d != NULL, ga is a bfloat16 tensor (the GEMM operand); with d == NULL it stays fp32 (it then feeds
// the backward FIR, sr_blur_nhwc_scaledot_bf16).
extern "C" int sr_styled_bwd_prologue3_bf16(void *ga, float *g_bias, float *g_noise_w, float *e, float *ds_next,
                                            float *d_rgb_weight, const float *gy, const float *gxs, const float *s_next,
                                            const float *g_rgb, const float *rgb_weight, const float *y, const float *noise,
                                            int64_t noise_batch_stride, const float *noise_weight, const float *bias,
                                            const float *d, int64_t batch, int64_t pixels, int64_t channels, float alpha,
                                            float gain, const float *stylemap, int64_t stylemap_batch_stride,
                                            float *g_stylemap, void *stream)
{
    return styled_bwd_prologue_any(reinterpret_cast<float *>(ga), g_bias, g_noise_w, e, ds_next, d_rgb_weight, gy, gxs, s_next,
                                   g_rgb, rgb_weight, y, noise, noise_batch_stride, noise_weight, bias, d, batch, pixels, channels,
                                   alpha, gain, stylemap, stylemap_batch_stride, g_stylemap, stream, 1);
}

extern "C" int sr_styled_bwd_prologue2_f32(float *ga, float *g_bias, float *g_noise_w, float *e, float *ds_next,
                                           float *d_rgb_weight, const float *gy, const float *gxs, const float *s_next,
                                           const float *g_rgb, const float *rgb_weight, const float *y, const float *noise,
                                           int64_t noise_batch_stride, const float *noise_weight, const float *bias,
                                           const float *d, int64_t batch, int64_t pixels, int64_t channels, float alpha,
                                           float gain, void *stream)
{
    return sr_styled_bwd_prologue3_f32(ga, g_bias, g_noise_w, e, ds_next, d_rgb_weight, gy, gxs, s_next, g_rgb, rgb_weight, y,
                                       noise, noise_batch_stride, noise_weight, bias, d, batch, pixels, channels, alpha, gain,
                                       nullptr, 0, nullptr, stream);
}

extern "C" int sr_styled_bwd_prologue_f32(float *ga, float *g_bias, float *g_noise_w, float *e, const float *gy,
                                          const float *y, const float *noise, int64_t noise_batch_stride,
                                          const float *noise_weight, const float *bias, const float *d, int64_t batch,
                                          int64_t pixels, int64_t channels, float alpha, float gain, void *stream)
{
    return sr_styled_bwd_prologue2_f32(ga, g_bias, g_noise_w, e, nullptr, nullptr, gy, nullptr, nullptr, nullptr, nullptr, y,
                                       noise, noise_batch_stride, noise_weight, bias, d, batch, pixels, channels, alpha, gain,
                                       stream);
}

extern "C" int sr_scale_dot_nhwc_f32(float *out, float *dot, const float *a, const float *other, const float *scale,
                                     int64_t batch, int64_t pixels, int64_t channels, int round_out_tf32, void *stream)
{
    SR_REQUIRE(a && (out || dot), "scale_dot: nothing to do");
    SR_REQUIRE(!dot || other, "scale_dot: dot needs the second tensor");
    SR_REQUIRE(ok_channels(channels), "scale_dot: channels must be 4*k with k dividing 256 (got %lld)", (long long)channels);
    SR_REQUIRE(al16(a) && (!out || al16(out)) && (!other || al16(other)) && (!scale || al16(scale)), "scale_dot: 16-byte alignment");
    cudaStream_t st = (cudaStream_t)stream;
    if (dot) {
        cudaError_t er = cudaMemsetAsync(dot, 0, sizeof(float) * (size_t)(batch * channels), st);
        if (er != cudaSuccess) { set_error("scale_dot: memset: %s", cudaGetErrorString(er)); return (int)er; }
    }
    if (batch == 0 || pixels == 0) return SR_OK;
    ScaleDotParams p;
    p.out = out; p.dot = dot; p.a = a; p.other = dot ? other : nullptr; p.scale = scale;
    p.pixels = (int)pixels; p.C4 = (int)(channels / 4); p.round_out = round_out_tf32;
    p.chunks_per_image = pick_chunks(batch, pixels, p.C4, &p.pix_per_chunk);
    scale_dot_kernel<<<(unsigned)(batch * p.chunks_per_image), kThreads, 0, st>>>(p);
    count_launch();
    return check_launch("sr_scale_dot_nhwc_f32");
}
