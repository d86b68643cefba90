// Micro-benchmark: cycles per tcgen05.mma.kind::tf32 (M128 x N x K8, cta_group::1) as a function of the A descriptor's
// start row (multiple of 8 or not) and stride byte offset.  Shared memory holds zeros; only timing matters.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3ffffu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
template <int N>
__global__ void __launch_bounds__(128) k(long long *out, int shift_a, int sbo_a, int shift_b, int iters)
{
    extern __shared__ uint8_t raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;                  // 64 KB of zeros
    uint8_t *sB = smem + 65536;          // 64 KB of zeros
    uint64_t *mbar = (uint64_t *)(smem + 131072);
    uint32_t *slot = (uint32_t *)(mbar + 1);
    for (int i = threadIdx.x; i < 131072 / 4; i += 128) ((uint32_t *)smem)[i] = 0;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot)), "r"((uint32_t)N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t da0 = desc(smem_u32(sA) + shift_a * 128, sbo_a), db0 = desc(smem_u32(sB) + shift_b * 128, 1024);
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                uint64_t da = da0 + (uint64_t)(kk * 2), db = db0 + (uint64_t)(kk * 2);
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                             :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(mbar)) : "memory");
        mbar_wait(mbar, 0);
        long long t1 = clock64();
        out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"((uint32_t)N) : "memory");
}
template <int N> void run(long long *d, int sa, int sbo, int sb) {
    const int iters = 4000;
    cudaFuncSetAttribute(k<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 140000);
    k<N><<<1, 128, 140000>>>(d, sa, sbo, sb, iters);   // warm-up
    k<N><<<1, 128, 140000>>>(d, sa, sbo, sb, iters);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("N=%3d A start row %2d SBO %4d, B start row %d: %.1f cycles per MMA (M128xN%dxK8)%s\n", N, sa, sbo, sb,
           (double)h / (iters * 4.0), N, e == cudaSuccess ? "" : cudaGetErrorString(e));
}
int main() {
    long long *d; cudaMalloc(&d, 8);
    const int cfg[][3] = {{0, 1024, 0}, {1, 1024, 0}, {4, 1024, 0}, {8, 1024, 0}, {0, 1280, 0}, {1, 1280, 0}, {11, 1280, 0}, {0, 2048, 0},
                          {0, 1024, 1}, {0, 1024, 4}, {1, 1024, 1}};
    for (auto &c : cfg) run<128>(d, c[0], c[1], c[2]);
    for (auto &c : cfg) run<256>(d, c[0], c[1], c[2]);
    for (auto &c : cfg) run<64>(d, c[0], c[1], c[2]);
    return 0;
}
