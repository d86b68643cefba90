// Shared host/device helpers for the sm_100a kernels of stylerenderer_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/stylerenderer_b200.h"

namespace sr {

constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs; grids are sized in multiples of this

// ---- error plumbing (thread-local message, see sr_last_error) ---------------------------------
void set_error(const char *fmt, ...);
void count_launch(int n = 1);

inline int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return SR_OK;
}

#define SR_REQUIRE(cond, ...)                      \
    do {                                           \
        if (!(cond)) {                             \
            ::sr::set_error(__VA_ARGS__);          \
            return SR_ERR_INVALID_ARGUMENT;        \
        }                                          \
    } while (0)

// ---- exact unsigned division by a runtime constant (host-prepared magic number) ----------------
// q = n / d for 0 <= n < 2^32, d >= 1: one mul.hi + shift instead of a ~20 instruction div.
struct FastDiv {
    uint32_t d, mul, shr;
    FastDiv() : d(1), mul(0), shr(0) {}
    explicit FastDiv(uint32_t div) : d(div) {
        if (div == 1) { mul = 0; shr = 0; return; }
        uint32_t l = 0;
        while ((1ull << l) < div) ++l;              // ceil(log2(d))
        uint64_t m = ((1ull << 32) * ((1ull << l) - div)) / div + 1;
        mul = (uint32_t)m;
        shr = l;
    }
    __host__ __device__ __forceinline__ uint32_t div(uint32_t n) const {
        if (d == 1) return n;
        uint32_t t = (uint32_t)(((uint64_t)n * mul) >> 32);
        return (t + ((n - t) >> 1)) >> (shr - 1);
    }
    __host__ __device__ __forceinline__ void divmod(uint32_t n, uint32_t &q, uint32_t &r) const {
        q = div(n);
        r = n - q * d;
    }
};

__host__ __device__ __forceinline__ int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// floor division / modulo for possibly negative numerators (positive denominators)
__host__ __device__ __forceinline__ int floor_div_i(int a, int b) {
    int q = a / b;
    return (q * b > a) ? q - 1 : q;
}
__host__ __device__ __forceinline__ int pos_mod_i(int a, int b) {
    int r = a % b;
    return r < 0 ? r + b : r;
}

// streaming (touch-once) global accesses: keep them out of L1
__device__ __forceinline__ float4 ld_stream4(const float *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream4(float *p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace sr
