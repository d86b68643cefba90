// Fused bias + leaky-ReLU (forward, backward with fused bias-gradient) for sm_100a.
//
// Replaces the reference's fused_bias_act_op / fused_bias_act_kernel
// (reference op/fused_bias_act.cpp:3-30, op/fused_bias_act_kernel.cu:14-112) and the second
// reduction pass of op/fused_act.py:33-38.
//
// HBM-bound elementwise work: 8 B/element forward, 12 B/element backward.
//   * 128-bit streaming loads/stores (ld.global.nc.L1::no_allocate), one float4 per lane;
//   * the channel of a float4 comes from a mul-hi "magic" division (no div/mod per element);
//   * backward: the bias gradient is reduced in the same pass -- warp shuffle when the warp sits
//     in one channel, per-CTA shared-memory bins otherwise, one global atomic per (CTA, channel);
//   * grid = a multiple of 148 SMs, grid-stride.
// Arithmetic order equals the reference kernel's ((x+b) -> select/alpha -> *scale), so forward and
// dx are bit-exact with it.
#include "common.cuh"

namespace sr {
namespace {

constexpr int kThreads = 256;
constexpr int kCtasPerSM = 8;     // 2048 resident threads per SM

enum ActMode { kLinear = 0, kLreluSelf = 1, kLreluRef = 2, kZero = 3 };

__device__ __forceinline__ float act1(float x, float b, float r, int mode, float alpha, float scale) {
    float t = x + b;
    float y;
    if (mode == kLreluSelf) y = (t > 0.0f) ? t : t * alpha;
    else if (mode == kLreluRef) y = (r > 0.0f) ? t : t * alpha;
    else if (mode == kLinear) y = t;
    else y = 0.0f;
    return y * scale;
}

// CH: 0 = no bias, 1 = plane-major (bias constant over a float4, channel = (i4 / step4) % C),
//     2 = channel-fastest (bias float4 at (4*i4) % C)
template <int MODE, int CH, bool REF>
__global__ void __launch_bounds__(kThreads)
bias_act_vec4_kernel(float *__restrict__ y, const float *__restrict__ x, const float *__restrict__ bias,
                     const float *__restrict__ ref, float alpha, float scale, uint32_t n4,
                     FastDiv step4, FastDiv nchan)
{
    const uint32_t stride = gridDim.x * kThreads;
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n4; i += stride) {
        float4 v = ld_stream4(x + 4ull * i);
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (REF) r = ld_stream4(ref + 4ull * i);
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (CH == 1) {
            uint32_t plane = step4.div(i), q, c;
            nchan.divmod(plane, q, c);
            float bv = __ldg(bias + c);
            b = make_float4(bv, bv, bv, bv);
        } else if (CH == 2) {
            uint32_t q, c;
            nchan.divmod(i, q, c);          // nchan here divides by C/4
            b = __ldg(reinterpret_cast<const float4 *>(bias) + c);
        }
        float4 o;
        o.x = act1(v.x, b.x, r.x, MODE, alpha, scale);
        o.y = act1(v.y, b.y, r.y, MODE, alpha, scale);
        o.z = act1(v.z, b.z, r.z, MODE, alpha, scale);
        o.w = act1(v.w, b.w, r.w, MODE, alpha, scale);
        st_stream4(y + 4ull * i, o);
    }
}

// Generic fallback: any shape / alignment, 64-bit indices.
__global__ void __launch_bounds__(kThreads)
bias_act_scalar_kernel(float *__restrict__ y, const float *__restrict__ x, const float *__restrict__ bias,
                       const float *__restrict__ ref, int mode, float alpha, float scale,
                       int64_t n, int64_t step_b, int64_t size_b)
{
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
        float b = bias ? bias[(i / step_b) % size_b] : 0.0f;
        float r = ref ? ref[i] : 0.0f;
        y[i] = act1(x[i], b, r, mode, alpha, scale);
    }
}

// ---- backward: dx = scale * (y > 0 ? g : alpha*g), dbias[c] = sum dx ---------------------------
// Plane-major layout (step_b % 4 == 0).  Loop bounds are warp-uniform so the shuffle reduction is legal.
__global__ void __launch_bounds__(kThreads)
lrelu_bwd_planes_kernel(float *__restrict__ dx, float *__restrict__ dbias, const float *__restrict__ gy,
                        const float *__restrict__ yref, float alpha, float scale, uint32_t n4,
                        FastDiv step4, FastDiv nchan, int C)
{
    extern __shared__ float s_db[];
    const bool reduce = dbias != nullptr;
    if (reduce) {
        for (int i = threadIdx.x; i < C; i += kThreads) s_db[i] = 0.0f;
        __syncthreads();
    }
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = (blockIdx.x * kThreads + threadIdx.x) >> 5;
    const uint32_t stride = (gridDim.x * kThreads);          // elements (float4) per sweep
    uint32_t cur_c = 0xffffffffu;                            // lane 0: channel of the running sum
    float acc = 0.0f;
    for (uint64_t base = (uint64_t)warp * 32u; base < n4; base += stride) {
        const uint32_t i = (uint32_t)base + lane;
        const bool valid = i < n4;
        float s = 0.0f;
        uint32_t c = 0xfffffffeu;
        if (valid) {
            float4 g = ld_stream4(gy + 4ull * i);
            float4 r = ld_stream4(yref + 4ull * i);
            float4 o;
            o.x = ((r.x > 0.0f) ? g.x : g.x * alpha) * scale;
            o.y = ((r.y > 0.0f) ? g.y : g.y * alpha) * scale;
            o.z = ((r.z > 0.0f) ? g.z : g.z * alpha) * scale;
            o.w = ((r.w > 0.0f) ? g.w : g.w * alpha) * scale;
            st_stream4(dx + 4ull * i, o);
            if (reduce) {
                s = (o.x + o.y) + (o.z + o.w);
                uint32_t q;
                nchan.divmod(step4.div(i), q, c);
            }
        }
        if (reduce) {
            const uint32_t c0 = __shfl_sync(0xffffffffu, c, 0);
            if (__all_sync(0xffffffffu, c == c0)) {          // whole warp inside one channel plane
                s = warp_sum(s);
                if (lane == 0) {
                    if (c0 != cur_c) {
                        if (cur_c != 0xffffffffu) atomicAdd(&s_db[cur_c], acc);
                        cur_c = c0;
                        acc = 0.0f;
                    }
                    acc += s;
                }
            } else if (valid) {
                atomicAdd(&s_db[c], s);
            }
        }
    }
    if (reduce) {
        if (lane == 0 && cur_c != 0xffffffffu) atomicAdd(&s_db[cur_c], acc);
        __syncthreads();
        for (int i = threadIdx.x; i < C; i += kThreads) {
            float v = s_db[i];
            if (v != 0.0f) atomicAdd(dbias + i, v);
        }
    }
}

// Channel-fastest layout ([B,C] or channels-last), C % 4 == 0: a thread keeps one channel quad as
// long as (grid * block * 4) % C == 0, which the launcher guarantees whenever possible.
__global__ void __launch_bounds__(kThreads)
lrelu_bwd_chlast_kernel(float *__restrict__ dx, float *__restrict__ dbias, const float *__restrict__ gy,
                        const float *__restrict__ yref, float alpha, float scale, uint32_t n4,
                        FastDiv cquads, int C)
{
    extern __shared__ float s_db[];
    const bool reduce = dbias != nullptr;
    if (reduce) {
        for (int i = threadIdx.x; i < C; i += kThreads) s_db[i] = 0.0f;
        __syncthreads();
    }
    const uint32_t stride = gridDim.x * kThreads;
    uint32_t cur = 0xffffffffu;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n4; i += stride) {
        float4 g = ld_stream4(gy + 4ull * i);
        float4 r = ld_stream4(yref + 4ull * i);
        float4 o;
        o.x = ((r.x > 0.0f) ? g.x : g.x * alpha) * scale;
        o.y = ((r.y > 0.0f) ? g.y : g.y * alpha) * scale;
        o.z = ((r.z > 0.0f) ? g.z : g.z * alpha) * scale;
        o.w = ((r.w > 0.0f) ? g.w : g.w * alpha) * scale;
        st_stream4(dx + 4ull * i, o);
        if (reduce) {
            uint32_t q, c;
            cquads.divmod(i, q, c);
            if (c != cur) {
                if (cur != 0xffffffffu) {
                    atomicAdd(&s_db[4 * cur + 0], acc.x); atomicAdd(&s_db[4 * cur + 1], acc.y);
                    atomicAdd(&s_db[4 * cur + 2], acc.z); atomicAdd(&s_db[4 * cur + 3], acc.w);
                }
                cur = c;
                acc = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
        }
    }
    if (reduce) {
        if (cur != 0xffffffffu) {
            atomicAdd(&s_db[4 * cur + 0], acc.x); atomicAdd(&s_db[4 * cur + 1], acc.y);
            atomicAdd(&s_db[4 * cur + 2], acc.z); atomicAdd(&s_db[4 * cur + 3], acc.w);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < C; i += kThreads) {
            float v = s_db[i];
            if (v != 0.0f) atomicAdd(dbias + i, v);
        }
    }
}

__global__ void __launch_bounds__(kThreads)
lrelu_bwd_scalar_kernel(float *__restrict__ dx, float *__restrict__ dbias, const float *__restrict__ gy,
                        const float *__restrict__ yref, float alpha, float scale, int64_t n,
                        int64_t step_b, int64_t size_b)
{
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
        float g = gy[i];
        float o = ((yref[i] > 0.0f) ? g : g * alpha) * scale;
        dx[i] = o;
        if (dbias) atomicAdd(dbias + (i / step_b) % size_b, o);
    }
}

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline int grid_for(uint64_t work_items) {
    uint64_t blocks = (work_items + kThreads - 1) / kThreads;
    uint64_t cap = (uint64_t)kNumSMs * kCtasPerSM;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

template <int MODE>
void launch_vec(float *y, const float *x, const float *bias, const float *ref, float alpha, float scale,
                uint32_t n4, int ch, FastDiv step4, FastDiv nchan, cudaStream_t st)
{
    const int grid = grid_for(n4);
#define SR_LAUNCH(CH, REF) bias_act_vec4_kernel<MODE, CH, REF><<<grid, kThreads, 0, st>>>( \
        y, x, bias, ref, alpha, scale, n4, step4, nchan)
    const bool use_ref = (MODE == kLreluRef) && ref != nullptr;
    if (ch == 0) { if (use_ref) SR_LAUNCH(0, true); else SR_LAUNCH(0, false); }
    else if (ch == 1) { if (use_ref) SR_LAUNCH(1, true); else SR_LAUNCH(1, false); }
    else { if (use_ref) SR_LAUNCH(2, true); else SR_LAUNCH(2, false); }
#undef SR_LAUNCH
}

}  // namespace
}  // namespace sr

using namespace sr;

extern "C" int sr_fused_bias_act_f32(float *y, const float *x, const float *bias, const float *ref,
                                     int act, int grad, float alpha, float scale,
                                     int64_t size_x, int64_t step_b, int64_t size_b, void *stream)
{
    SR_REQUIRE(size_x >= 0, "fused_bias_act: negative size");
    if (size_x == 0) return SR_OK;
    SR_REQUIRE(y && x, "fused_bias_act: null x/y");
    SR_REQUIRE(!bias || (step_b >= 1 && size_b >= 1), "fused_bias_act: bad bias geometry");
    cudaStream_t st = (cudaStream_t)stream;
    int mode;
    switch (act * 10 + grad) {                   // reference op/fused_bias_act_kernel.cu:28-39
    case 30: mode = kLreluSelf; break;
    case 31: mode = kLreluRef; break;
    case 12: case 32: mode = kZero; break;
    default: mode = kLinear; break;
    }
    // ref == NULL means ref = 0 (reference `use_ref` = 0): mode 31 then always takes the alpha branch
    const bool vec_ok = aligned16(y) && aligned16(x) && (!ref || aligned16(ref)) && (size_x % 4 == 0) &&
                        size_x / 4 < 0xffffffffll;
    int ch = -1;
    FastDiv step4, nchan;
    if (vec_ok) {
        if (!bias) ch = 0;
        else if (step_b % 4 == 0 && step_b / 4 < 0x7fffffffll && size_b < 0x7fffffffll) {
            ch = 1; step4 = FastDiv((uint32_t)(step_b / 4)); nchan = FastDiv((uint32_t)size_b);
        } else if (step_b == 1 && size_b % 4 == 0 && aligned16(bias)) {
            ch = 2; nchan = FastDiv((uint32_t)(size_b / 4));
        }
    }
    if (ch >= 0) {
        const uint32_t n4 = (uint32_t)(size_x / 4);
        switch (mode) {
        case kLreluSelf: launch_vec<kLreluSelf>(y, x, bias, ref, alpha, scale, n4, ch, step4, nchan, st); break;
        case kLreluRef: launch_vec<kLreluRef>(y, x, bias, ref, alpha, scale, n4, ch, step4, nchan, st); break;
        case kLinear: launch_vec<kLinear>(y, x, bias, ref, alpha, scale, n4, ch, step4, nchan, st); break;
        default: launch_vec<kZero>(y, x, bias, ref, alpha, scale, n4, ch, step4, nchan, st); break;
        }
    } else {
        bias_act_scalar_kernel<<<grid_for((uint64_t)size_x), kThreads, 0, st>>>(
            y, x, bias, ref, mode, alpha, scale, size_x, step_b > 0 ? step_b : 1, size_b > 0 ? size_b : 1);
    }
    count_launch();
    return check_launch("sr_fused_bias_act_f32");
}

extern "C" int sr_fused_lrelu_backward_f32(float *dx, float *dbias, const float *gy, const float *yref,
                                           float alpha, float scale, int64_t size_x, int64_t step_b,
                                           int64_t size_b, void *stream)
{
    SR_REQUIRE(size_x >= 0, "fused_lrelu_backward: negative size");
    cudaStream_t st = (cudaStream_t)stream;
    if (dbias) {
        SR_REQUIRE(step_b >= 1 && size_b >= 1, "fused_lrelu_backward: bad bias geometry");
        cudaError_t e = cudaMemsetAsync(dbias, 0, sizeof(float) * (size_t)size_b, st);
        if (e != cudaSuccess) { set_error("fused_lrelu_backward: memset: %s", cudaGetErrorString(e)); return (int)e; }
    }
    if (size_x == 0) return SR_OK;
    SR_REQUIRE(dx && gy && yref, "fused_lrelu_backward: null pointer");
    const bool vec_ok = aligned16(dx) && aligned16(gy) && aligned16(yref) && (size_x % 4 == 0) &&
                        size_x / 4 < 0xffffffffll;
    const size_t smem = dbias ? sizeof(float) * (size_t)size_b : 0;
    const uint32_t n4 = (uint32_t)(size_x / 4);
    if (vec_ok && smem <= 48 * 1024 && (!dbias || (step_b % 4 == 0 && step_b / 4 < 0x7fffffffll))) {
        FastDiv step4(dbias ? (uint32_t)(step_b / 4) : 1u), nchan(dbias ? (uint32_t)size_b : 1u);
        lrelu_bwd_planes_kernel<<<grid_for(n4), kThreads, smem, st>>>(dx, dbias, gy, yref, alpha, scale, n4,
                                                                      step4, nchan, (int)size_b);
    } else if (vec_ok && smem <= 48 * 1024 && step_b == 1 && size_b % 4 == 0) {
        // make (grid * 256 * 4) a multiple of C when C/4 divides a CTA multiple, so each thread keeps its quad
        int grid = grid_for(n4);
        const int64_t quads = size_b / 4;
        if (quads % kThreads == 0) { int64_t m = quads / kThreads; grid = (int)((grid + m - 1) / m * m); }
        FastDiv cquads((uint32_t)quads);
        lrelu_bwd_chlast_kernel<<<grid, kThreads, smem, st>>>(dx, dbias, gy, yref, alpha, scale, n4, cquads,
                                                              (int)size_b);
    } else {
        lrelu_bwd_scalar_kernel<<<grid_for((uint64_t)size_x), kThreads, 0, st>>>(
            dx, dbias, gy, yref, alpha, scale, size_x, step_b > 0 ? step_b : 1, size_b > 0 ? size_b : 1);
    }
    count_launch();
    return check_launch("sr_fused_lrelu_backward_f32");
}
