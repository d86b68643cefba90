// Experiment: can a K-major SWIZZLE_128B UMMA operand start at a row that is not a multiple of 8 (1024 B)?
// A [136 rows x 32 tf32] is TMA-loaded (128B swizzle) at a 1024-aligned base; for shift r = 0..8 we run
// D = A[r:r+128, :] * B^T with the descriptor start advanced by r*128 bytes and base_offset = 0 (mode 0) or
// base_offset = r & 7 (mode 1), and compare with the exact result.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t base_off, uint32_t sbo = 1024) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3ffffu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_off & 7) << 49;
    d |= (uint64_t)2 << 61;
    return d;
}
__global__ void __launch_bounds__(128) k(const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap tb, float *out, int shift, int mode, int sbo)
{
    extern __shared__ uint8_t raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;                 // RA rows * 128 B
    uint8_t *sB = smem + 32768;         // 128 rows * 128 B
    uint64_t *bar = (uint64_t *)(smem + 32768 + 16384);
    uint64_t *mbar = bar + 1;
    uint32_t *slot = (uint32_t *)(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(240 * 128 + 128 * 128) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     :: "r"(smem_u32(sA)), "l"(&ta), "r"(smem_u32(bar)), "r"(0), "r"(0) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     :: "r"(smem_u32(sB)), "l"(&tb), "r"(smem_u32(bar)), "r"(0), "r"(0) : "memory");
        mbar_wait(bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t a_addr = smem_u32(sA) + shift * 128;
        for (int kk = 0; kk < 4; ++kk) {
            uint64_t da = desc(a_addr, mode ? ((a_addr >> 7) & 7) : 0, (uint32_t)sbo) + (uint64_t)(kk * 2);
            uint64_t db = desc(smem_u32(sB), 0) + (uint64_t)(kk * 2);
            uint32_t acc = kk != 0;
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                         :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(mbar)) : "memory");
    }
    mbar_wait(mbar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                       "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                     : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c * 32));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 128 + c * 32 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(128u) : "memory");
}
typedef CUresult (*Enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
    void *fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    Enc enc = (Enc)fp;
    const int RA = 240, RB = 128, K = 32;
    float *hA = (float *)malloc(RA * K * 4), *hB = (float *)malloc(RB * K * 4);
    for (int i = 0; i < RA * K; ++i) hA[i] = (float)((int)(((unsigned)i * 2654435761u) >> 20) % 17 - 8);
    for (int i = 0; i < RB * K; ++i) hB[i] = (float)((i * 5 + 1) % 13 - 6) * 0.5f;
    float *dA, *dB, *dO; cudaMalloc(&dA, RA * K * 4); cudaMalloc(&dB, RB * K * 4); cudaMalloc(&dO, 128 * 128 * 4);
    cudaMemcpy(dA, hA, RA * K * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, RB * K * 4, cudaMemcpyHostToDevice);
    CUtensorMap ta, tb;
    cuuint64_t dimsA[2] = {(cuuint64_t)K, (cuuint64_t)RA}, dimsB[2] = {(cuuint64_t)K, (cuuint64_t)RB}, str[1] = {(cuuint64_t)K * 4};
    cuuint32_t boxA[2] = {32, 240}, boxB[2] = {32, 128}, es[2] = {1, 1};
    enc(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dA, dimsA, str, boxA, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    enc(&tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dB, dimsB, str, boxB, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
    float *hO = (float *)malloc(128 * 128 * 4);
    const int sbos[3] = {1024, 1280, 2304};           // 8-row groups 8 / 10 / 18 rows apart (dense, 8-wide tile in a 10- or 18-wide halo)
    for (int si = 0; si < 3; ++si)
    for (int mode = 0; mode < 2; ++mode)
        for (int shift = 0; shift <= 22; ++shift) {
            const int sbo = sbos[si], grow = sbo / 128;
            if (shift + 15 * grow + 8 > RA) continue;
            cudaMemset(dO, 0, 128 * 128 * 4);
            k<<<1, 128, 60000>>>(ta, tb, dO, shift, mode, sbo);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("sbo %d mode %d shift %d: CUDA error %s\n", sbo, mode, shift, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(hO, dO, 128 * 128 * 4, cudaMemcpyDeviceToHost);
            double maxerr = 0;
            for (int m = 0; m < 128; ++m) for (int n = 0; n < 128; ++n) {
                const int row = shift + (m / 8) * grow + (m % 8);
                double s = 0; for (int kk = 0; kk < K; ++kk) s += (double)hA[row * K + kk] * hB[n * K + kk];
                maxerr = fmax(maxerr, fabs(s - hO[m * 128 + n]));
            }
            printf("sbo %4d mode %d (base_offset %s) shift %2d rows: max abs err %.3g %s\n", sbo, mode, mode ? "= (addr>>7)&7" : "= 0", shift, maxerr, maxerr < 1e-3 ? "OK" : "WRONG");
        }
    return 0;
}
