// Modulated convolution on 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a only.
//
// Replaces the dense contraction of ModulatedConv2d (reference layers.py:293-323): cuDNN grouped
// conv / grouped transposed conv over per-sample weight copies.  Here all samples share ONE weight
// matrix (the style scales the activations, the demodulation scales the outputs, see layers.py of this
// package), so every variant is an implicit GEMM   D[pixel, cout] = sum_{tap, cin} A[pixel+tap, cin] * W[cout, tap, cin]
//   M = batch * grid_h * grid_w pixels, N = cout, K = taps * cin
// and one kernel covers all of them through a tap list + an input stride + an output lattice:
//   plain 3x3 conv / its dgrad      9 taps, stride 1, dense output
//   transposed stride-2 conv        4 phase launches (4/2/2/1 taps), output lattice stride 2
//   dgrad of the transposed conv    9 taps, input stride 2 (TMA element strides)
//
// Data path per CTA (one 128-pixel x BLOCK_N tile of D, fp32 accumulator in TMEM):
//   warp 0   TMA producer: per (tap, 32-channel K block) one 4-D box {32 ch, TW, TH, TN} of the NHWC
//            activations -- the tap shift and the zero padding are done by the TMA unit (signed box
//            coordinates, out-of-bound fill) -- plus one 2-D box {32, BLOCK_N} of the weights, both
//            128-byte swizzled, into a STAGES-deep ring guarded by full/empty mbarriers;
//   warp 1   MMA issuer: one elected lane issues tcgen05.mma.cta_group::1.kind::tf32 (M128 x N x K8, 4 per
//            K block) on shared-memory descriptors, tcgen05.commit releases ring slots / publishes TMEM;
//   warps 2-5 epilogue: tcgen05.ld 32 lanes x 32 columns, fused StyledConv epilogue
//            (demodulate, style-map affine, noise, bias, leaky-ReLU * sqrt2, optional second output
//            pre-multiplied by the NEXT layer's style and rounded to tf32), 128-byte row stores.
// fp32 storage, TF32 multiplicands (both operands are rounded to tf32 with cvt.rna by the kernels that
// produce them), fp32 accumulation: the same arithmetic class as the reference's cuDNN path under
// torch's default cudnn.allow_tf32 = True.
#include <algorithm>
#include "common.cuh"
#include "tma_host.cuh"
#include <stdlib.h>

namespace sr {
namespace {

constexpr int kConvThreads = 192;
constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 32;                         // fp32 elements = one 128-byte swizzle row
constexpr int A_BYTES = BLOCK_M * BLOCK_K * 4;

// ------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// One lane polls, the warp re-converges: 32 lanes spinning on a barrier word compete with the tensor core's operand
// reads for the shared-memory port (measured on the wgrad kernel: +20 % time with all lanes polling).
__device__ __forceinline__ void mbar_wait_warp(uint64_t *bar, uint32_t parity, int one_lane = 0) {
    if (one_lane) {
        if ((threadIdx.x & 31) == 0) mbar_wait(bar, parity);
        __syncwarp();
    } else {
        mbar_wait(bar, parity);
    }
}
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap *map, const void *src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 :: "l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tcgen05_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], tf32 inputs, fp32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

// same with 16-bit operands (kind::f16: bf16 x bf16 -> fp32; K = 16 elements = the same 32 bytes per instruction)
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// two floats -> one 32-bit word of two bf16 (round to nearest even), `lo` in the low half = the lower address
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// instruction descriptor: tf32 formats (2) -> bf16 formats (1) for both operands
__device__ __forceinline__ uint32_t idesc_bf16(uint32_t idesc_tf32) { return idesc_tf32 - (1u << 7) - (1u << 10); }

// K-major operand tile, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor:
// start >> 4 at [0,14), LBO >> 4 at [16,30) (unused for swizzled K-major, 1), SBO >> 4 at [32,46),
// version 1 at [46,48), layout SWIZZLE_128B = 2 at [61,64)).
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__device__ __forceinline__ float round_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}

// Warp-uniform election of one issuing lane (the same lane every time).  The producer and MMA warps run their loops with
// all 32 lanes converged and elect only around the asynchronous instructions: inside an `if (lane == 0)` region the
// compiler cannot use the uniform datapath and wraps every UTMALDG / UTCHMMA operand in an ELECT / R2UR / BRA.U.ANY
// waterfall loop (~25 instructions per MMA; ncu: the single issuing thread, not the data, bounded the N = 128 layers).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ------------------------------------------------------------------------------------ kernel
constexpr int kMaxPhases = 4;

struct ConvPhase {                          // one tap list + output lattice (the 4 parity classes of a transposed conv)
    int num_taps;
    int tap_dy[9], tap_dx[9], tap_k0[9];    // input offset of a tap and its first K column in the weight matrix
    int grid_h, grid_w, tiles_x, tiles_y;
    int tile_begin;                         // first M tile of this phase inside an image group
    int tile_begin2;                        // same with every phase padded to an even tile count (CTA-pair kernel)
    long long out_offset;                   // lattice origin (y0 * row + x0 * pix), in floats
    long long noise_offset;
};

struct ConvKParams {
    int tw_log2, th_log2, tn_log2;          // tile = 2^tn images x 2^th rows x 2^tw columns = 128 pixels
    int batch;
    int op16;                               // 0: fp32 words holding tf32 operands (kind::tf32)   1: bf16 operands (kind::f16);
                                            // applies to the activation operand, the weights and the second output
    int kelems;                             // channels per 128-byte K block: 32 (tf32) or 64 (bf16)
    int kblocks_per_tap;                    // cin / kelems
    int num_phases, n_tiles, total_tiles;   // total_tiles = (sum of M tiles) * n_tiles
    int total_pairs;                        // CTA-pair kernel: (sum of padded M tiles / 2) * n_tiles
    // Tile order: image groups outermost, then phases, then the tiles of the phase.  With one image (or tn-image tile) per
    // group the four parity classes of a transposed conv visit the same input image back to back, so they read it from L2
    // instead of streaming the whole batch from DRAM once per phase (ncu, 256->128 @128^2, phase-major order: 2.15 GB
    // read for a 0.54 GB input, DRAM 57 % busy).  group_n = tiles_n reproduces the phase-major order.
    int group_n, tiles_n, tiles_per_group, tiles_per_group2;
    ConvPhase ph[kMaxPhases];
    int in_stride;
    int cout;
    long long out_img_stride, out_row_stride, out_pix_stride;   // in floats
    float *out, *out2;
    int epilogue;                           // 0: acc * rowscale   1: styled   2: styled, but `out` receives acc * rowscale (the
                                            //    value before the map affine / noise / bias / activation; out2 and ToRGB see y)
    const float *rowscale, *scale2, *bias, *noise, *noise_weight, *stylemap;
    long long noise_img_stride, noise_row_stride, noise_pix_stride;
    long long map_img_stride, map_plane_stride;
    float alpha, gain;
    const float *rgb_w;                     // fused ToRGB: [batch, 3, cout] per-sample 1x1 weights (or nullptr)
    float *rgb_out;                         // [batch, out_h, out_w, 3], += sum_c out[..., c] * rgb_w[n, k, c]
    int debug;                              // SR_CONV_DEBUG (profiling only): 1 = epilogue handshake only, 2 = no MMAs,
                                            // 4 = one lane polls the barriers, 8 = halo epilogue stores without TMA
};

// TMA-store descriptors of the output lattice(s): one per phase (the parity classes of a transposed conv write
// interleaved sub-lattices = strided views of the output), for `out` and for the optional second output.
// Box = the 32 rows of one epilogue warp x 32 channels, 128-byte swizzled.
struct alignas(64) ConvOutMaps {
    CUtensorMap out[kMaxPhases];
    CUtensorMap out2[kMaxPhases];
};
constexpr int kStageBufBytes = 32 * 128;               // one warp, one 32-channel chunk
constexpr int kEpiSmemBytes = 4 * 2 * kStageBufBytes;  // 4 epilogue warps x 2 buffers

struct TileCoord { int phase, gx0, gy0, n0, n_tile; bool skip; };

__device__ __forceinline__ TileCoord decode_tile(const ConvKParams &p, int T) {
    TileCoord c;
    c.n_tile = T % p.n_tiles;
    int mt = T / p.n_tiles;
    const int g = mt / p.tiles_per_group;
    mt -= g * p.tiles_per_group;
    c.phase = 0;
#pragma unroll
    for (int i = 1; i < kMaxPhases; ++i)
        if (i < p.num_phases && mt >= p.ph[i].tile_begin) c.phase = i;
    const ConvPhase &ph = p.ph[c.phase];
    mt -= ph.tile_begin;
    const int tx = mt % ph.tiles_x, ty = (mt / ph.tiles_x) % ph.tiles_y, tn = g * p.group_n + mt / (ph.tiles_x * ph.tiles_y);
    c.skip = tn >= p.tiles_n;                                   // last, partial image group
    c.gx0 = tx << p.tw_log2; c.gy0 = ty << p.th_log2; c.n0 = tn << p.tn_log2;
    return c;
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}

// Epilogue of one 128-row accumulator: wait for the MMA, tcgen05.ld 32 columns at a time, fused tail, then each warp
// stages its 32 rows x 32 channels (4 KB, 128-byte swizzled, conflict free) in shared memory and one lane hands the block
// to the TMA store unit: full 128-byte lines leave the SM (per-thread 16-byte global stores made the epilogue as long
// as the main loop on the K = 1152 layers), and the unit clips partial tiles and walks strided lattices by itself.
// Returns the three ToRGB partial sums of this thread's pixel in rgb[] and whether the pixel is inside the lattice.
// STAGED: the per-sample / per-channel vectors of the tile (demodulation, next-layer style, bias, ToRGB weights) were
// copied to shared memory by stage_tile_vectors() BEFORE the wait for the accumulator, so the chunk loop reads them with
// broadcast LDS.128 instead of stalling on dependent global loads (ncu: `long_sb` on those loads made the epilogue of a
// 128-channel tile longer than its main loop).  Requires all 128 rows of the tile to belong to one image (tn = 1).
constexpr int kVecKinds = 6;                            // rowscale, scale2, bias, rgb_w[0..2]
template <int BLOCK_N, int THREADS>
__device__ __forceinline__ void stage_tile_vectors(const ConvKParams &p, const TileCoord &tc, float *vb, int epi_tid)
{
    const int n = tc.n0 < p.batch ? tc.n0 : p.batch - 1;        // padded tile of an odd phase: any valid image
    const int ch0 = tc.n_tile * BLOCK_N;
    constexpr int Q = BLOCK_N / 4;
    for (int i = epi_tid; i < kVecKinds * Q; i += THREADS) {
        const int v = i / Q, c4 = i - v * Q;
        const float *src = nullptr;
        if (v == 0) src = p.rowscale ? p.rowscale + (long long)n * p.cout + ch0 : nullptr;
        else if (v == 1) src = p.out2 ? p.scale2 + (long long)n * p.cout + ch0 : nullptr;
        else if (v == 2) src = (p.epilogue >= 1 && p.bias) ? p.bias + ch0 : nullptr;
        else src = p.rgb_w ? p.rgb_w + ((long long)n * 3 + (v - 3)) * p.cout + ch0 : nullptr;
        const float fill = (v == 0) ? 1.0f : 0.0f;
        const float4 val = src ? __ldg(reinterpret_cast<const float4 *>(src) + c4) : make_float4(fill, fill, fill, fill);
        *reinterpret_cast<float4 *>(vb + v * BLOCK_N + 4 * c4) = val;
    }
}

template <int BLOCK_N, bool STAGED = false, bool OP16 = false>
__device__ __forceinline__ bool epilogue_tile(const ConvKParams &p, const ConvOutMaps &om, const ConvPhase &ph,
                                              const TileCoord &tc, int q, int lane, int tx, int ty, int tn, uint32_t tmem_acc,
                                              uint64_t *tmem_full, uint32_t acc_par, uint8_t *stage, uint32_t &stage_sel,
                                              float (&rgb)[3], long long &rgb_index, const float *vb = nullptr,
                                              int c_begin = 0, int c_end = BLOCK_N / 32)
{
    const int gx = tc.gx0 + tx, gy = tc.gy0 + ty, n = tc.n0 + tn;
    const bool valid = gx < ph.grid_w && gy < ph.grid_h && n < p.batch;
    const int nb = valid ? n : 0;                                     // per-sample vectors of a masked row: any valid row
    float pre_add = 0.0f, map_mul = 1.0f;
    if (p.epilogue >= 1 && valid) {
        const long long npix = (long long)gy * p.noise_row_stride + (long long)gx * p.noise_pix_stride + ph.noise_offset;
        if (p.noise) pre_add = __ldg(p.noise_weight) * __ldg(p.noise + (long long)n * p.noise_img_stride + npix);
        if (p.stylemap) {
            const float *m = p.stylemap + (long long)n * p.map_img_stride + npix;
            map_mul = __ldg(m);
            pre_add += __ldg(m + p.map_plane_stride);
        }
    }
    // first pixel of this warp's 32 rows (box origin of its TMA stores)
    const int row0 = q * 32;
    const int wx = tc.gx0 + (row0 & ((1 << p.tw_log2) - 1));
    const int wy = tc.gy0 + ((row0 >> p.tw_log2) & ((1 << p.th_log2) - 1));
    const int wn = tc.n0 + (row0 >> (p.tw_log2 + p.th_log2));
    rgb[0] = rgb[1] = rgb[2] = 0.0f;
    // STAGED: rows leave through shared memory and coalesced 16-byte global stores (lanes 8k..8k+7 write the 128 bytes of
    // one pixel) instead of TMA stores: nothing in the chunk loop waits for the TMA unit, which is busy with the loads.
    // rowoff[i] = element offset of row (lane / 8 + 4 i) of this warp's 32 rows, or -1 outside the lattice.
    long long rowoff[8];
    const bool direct = STAGED && (p.debug & 8) && !OP16;
    if (direct) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int R = q * 32 + (lane >> 3) + 4 * i;
            const int rx = tc.gx0 + (R & ((1 << p.tw_log2) - 1));
            const int ry = tc.gy0 + ((R >> p.tw_log2) & ((1 << p.th_log2) - 1));
            const int rn = tc.n0 + (R >> (p.tw_log2 + p.th_log2));
            const bool ok = rx < ph.grid_w && ry < ph.grid_h && rn < p.batch;
            rowoff[i] = ok ? (long long)rn * p.out_img_stride + (long long)ry * p.out_row_stride +
                             (long long)rx * p.out_pix_stride + ph.out_offset + 4 * (lane & 7) : -1;
        }
    }
    mbar_wait_warp(tmem_full, acc_par, p.debug & 4);
    tcgen05_fence_after();
    if (p.debug & 1) { rgb_index = 0; return false; }
#pragma unroll 1
    for (int c = c_begin; c < c_end; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
        const int ch0 = tc.n_tile * BLOCK_N + c * 32;
        const float4 *rs = p.rowscale ? reinterpret_cast<const float4 *>(p.rowscale + (long long)nb * p.cout + ch0) : nullptr;
        const float4 *s2 = p.out2 ? reinterpret_cast<const float4 *>(p.scale2 + (long long)nb * p.cout + ch0) : nullptr;
        const float4 *bs = (p.epilogue >= 1 && p.bias) ? reinterpret_cast<const float4 *>(p.bias + ch0) : nullptr;
        uint8_t *buf1 = stage + (stage_sel & 1) * kStageBufBytes;
        uint8_t *buf2 = stage + ((stage_sel + 1) & 1) * kStageBufBytes;
        // the store that last read buf1 (two stores ago) must have finished reading shared memory
        if (!direct) {
            if (lane == 0) tma_store_wait_read<1>();
            __syncwarp();
        }
        float y[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float v[4] = {__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                          __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])};
            if (STAGED) { const float4 s = *reinterpret_cast<const float4 *>(vb + c * 32 + 4 * j); v[0] *= s.x; v[1] *= s.y; v[2] *= s.z; v[3] *= s.w; }
            else if (rs) { const float4 s = __ldg(rs + j); v[0] *= s.x; v[1] *= s.y; v[2] *= s.z; v[3] *= s.w; }
            const float4 pre = make_float4(v[0], v[1], v[2], v[3]);      // demodulated conv output (epilogue 2 stores this)
            if (p.epilogue >= 1) {
                float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
                if (STAGED) b = *reinterpret_cast<const float4 *>(vb + 2 * BLOCK_N + c * 32 + 4 * j);
                else if (bs) b = __ldg(bs + j);
                const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float t = v[e] * map_mul + pre_add + bb[e];
                    v[e] = ((t > 0.0f) ? t : t * p.alpha) * p.gain;
                }
            }
            y[4 * j] = v[0]; y[4 * j + 1] = v[1]; y[4 * j + 2] = v[2]; y[4 * j + 3] = v[3];
            // 128-byte swizzle: 16-byte chunk j of row `lane` lives at chunk j ^ (lane % 8)
            *reinterpret_cast<float4 *>(buf1 + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                (p.epilogue == 2) ? pre : make_float4(v[0], v[1], v[2], v[3]);
            if (p.rgb_w && valid) {                 // ToRGB rides the epilogue: 3 dot products over the channel chunk
                const float *w0 = p.rgb_w + (long long)n * 3 * p.cout + ch0 + 4 * j;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float4 ww = STAGED ? *reinterpret_cast<const float4 *>(vb + (3 + k) * BLOCK_N + c * 32 + 4 * j)
                                             : __ldg(reinterpret_cast<const float4 *>(w0 + (long long)k * p.cout));
                    rgb[k] = fmaf(v[0], ww.x, fmaf(v[1], ww.y, fmaf(v[2], ww.z, fmaf(v[3], ww.w, rgb[k]))));
                }
            }
        }
        if (direct) {
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int rr = (lane >> 3) + 4 * i;
                const float4 o = *reinterpret_cast<const float4 *>(buf1 + rr * 128 + (((lane & 7) ^ (rr & 7)) << 4));
                if (rowoff[i] >= 0) st_stream4(p.out + rowoff[i] + ch0, o);
            }
        } else {
            fence_async_smem();
            __syncwarp();
            if (lane == 0) { tma_store_4d(&om.out[tc.phase], buf1, ch0, wx, wy, wn); tma_store_commit(); }
        }
        ++stage_sel;
        if (STAGED ? (p.out2 != nullptr) : (s2 != nullptr)) {
            if (!direct) {
                if (lane == 0) tma_store_wait_read<1>();
                __syncwarp();
            }
            if (OP16) {
                // bf16 operand for the next layer: 32 channels = 64 bytes per row, 64-byte swizzle (16-byte chunk q of row r
                // lives at chunk q ^ ((r >> 1) & 3): conflict free for the 8 lanes of a store wavefront)
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    uint32_t w[4];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int j = 2 * q4 + h;
                        const float4 s = STAGED ? *reinterpret_cast<const float4 *>(vb + BLOCK_N + c * 32 + 4 * j) : __ldg(s2 + j);
                        w[2 * h] = pack_bf16(y[4 * j] * s.x, y[4 * j + 1] * s.y);
                        w[2 * h + 1] = pack_bf16(y[4 * j + 2] * s.z, y[4 * j + 3] * s.w);
                    }
                    *reinterpret_cast<uint4 *>(buf2 + lane * 64 + ((q4 ^ ((lane >> 1) & 3)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 s = STAGED ? *reinterpret_cast<const float4 *>(vb + BLOCK_N + c * 32 + 4 * j) : __ldg(s2 + j);
                    *reinterpret_cast<float4 *>(buf2 + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                        make_float4(round_tf32(y[4 * j] * s.x), round_tf32(y[4 * j + 1] * s.y), round_tf32(y[4 * j + 2] * s.z),
                                    round_tf32(y[4 * j + 3] * s.w));
                }
            }
            if (direct) {
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int rr = (lane >> 3) + 4 * i;
                    const float4 o = *reinterpret_cast<const float4 *>(buf2 + rr * 128 + (((lane & 7) ^ (rr & 7)) << 4));
                    if (rowoff[i] >= 0) st_stream4(p.out2 + rowoff[i] + ch0, o);
                }
            } else {
                fence_async_smem();
                __syncwarp();
                if (lane == 0) { tma_store_4d(&om.out2[tc.phase], buf2, ch0, wx, wy, wn); tma_store_commit(); }
            }
            ++stage_sel;
        }
    }
    rgb_index = ((long long)n * p.map_plane_stride + (long long)gy * p.noise_row_stride + (long long)gx * p.noise_pix_stride +
                 ph.noise_offset) * 3;
    return valid;
}

// Persistent: gridDim.x CTAs (one per SM) walk the tile list round-robin.  The TMEM holds TWO accumulators, so the
// epilogue of tile i (TMEM -> registers -> fused tail -> global) overlaps the TMA/MMA main loop of tile i+1.
template <int BLOCK_N, int STAGES, bool OP16>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_igemm_tf32_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                       const __grid_constant__ ConvOutMaps out_maps,
                       const ConvKParams p)
{
    constexpr int B_BYTES = BLOCK_N * BLOCK_K * 4;
    constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BLOCK_N >> 3) << 17) |
                                ((uint32_t)(BLOCK_M >> 4) << 24);            // f32 accum, tf32 x tf32, K-major A and B
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;
    uint8_t *sB = smem + STAGES * A_BYTES;
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + STAGES * (A_BYTES + B_BYTES));
    uint64_t *empty_bar = full_bar + STAGES;
    uint64_t *tmem_full_bar = empty_bar + STAGES;      // [2]
    uint64_t *tmem_empty_bar = tmem_full_bar + 2;      // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty_bar + 2);
    uint8_t *epi_stage = smem + STAGES * (A_BYTES + B_BYTES) + 1024;   // 1024-byte aligned (swizzle), 8 KB per epilogue warp

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_b) : "memory");
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {                                   // TMEM: 2 accumulators of BLOCK_N fp32 columns x 128 lanes
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(tmem_slot)), "r"((uint32_t)(2 * BLOCK_N)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {                                   // ===== TMA producer warp (converged loop, elect_one issues)
        uint32_t s = 0, par = 0;
        for (int T = blockIdx.x; T < p.total_tiles; T += gridDim.x) {
            const TileCoord tc = decode_tile(p, T);
            if (tc.skip) continue;
            const ConvPhase &ph = p.ph[tc.phase];
            for (int t = 0; t < ph.num_taps; ++t) {
                const int cx = tc.gx0 * p.in_stride + ph.tap_dx[t], cy = tc.gy0 * p.in_stride + ph.tap_dy[t];
                for (int kc = 0; kc < p.kblocks_per_tap; ++kc) {
                    mbar_wait(&empty_bar[s], par ^ 1);
                    if (elect_one()) {
                        mbar_expect_tx(&full_bar[s], A_BYTES + B_BYTES);
                        tma_load_4d(sA + s * A_BYTES, &tmap_a, &full_bar[s], kc * (OP16 ? 64 : 32), cx, cy, tc.n0);
                        tma_load_2d(sB + s * B_BYTES, &tmap_b, &full_bar[s], ph.tap_k0[t] + kc * (OP16 ? 64 : 32), tc.n_tile * BLOCK_N);
                    }
                    __syncwarp();
                    if (++s == STAGES) { s = 0; par ^= 1; }
                }
            }
        }
    } else if (warp == 1) {                            // ===== MMA warp
        uint32_t s = 0, par = 0, lt = 0;
        const uint32_t sA_u32 = smem_u32(sA), sB_u32 = smem_u32(sB);
        for (int T = blockIdx.x; T < p.total_tiles; T += gridDim.x) {
            const TileCoord tc = decode_tile(p, T);
            if (tc.skip) continue;
            const int num_kb = p.ph[tc.phase].num_taps * p.kblocks_per_tap;
            const uint32_t acc = lt & 1, acc_par = (lt >> 1) & 1;
            mbar_wait(&tmem_empty_bar[acc], acc_par ^ 1);          // epilogue has drained this accumulator
            tcgen05_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&full_bar[s], par);
                tcgen05_fence_after();
                const uint64_t da = make_kmajor_sw128_desc(sA_u32 + s * A_BYTES);
                const uint64_t db = make_kmajor_sw128_desc(sB_u32 + s * B_BYTES);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < BLOCK_K / 8; ++k) {    // one instruction = 32 bytes along the swizzled row (8 tf32 / 16 bf16)
                        if (p.debug & 2) continue;
                        if (OP16) umma_f16(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc_bf16(kIdesc), (kb | k) != 0);
                        else umma_tf32(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), kIdesc, (kb | k) != 0);
                    }
                    tcgen05_commit(&empty_bar[s]);             // frees the ring slot once these MMAs have read it
                }
                __syncwarp();
                if (++s == STAGES) { s = 0; par ^= 1; }
            }
            if (elect_one()) tcgen05_commit(&tmem_full_bar[acc]);  // accumulator complete
            __syncwarp();
            ++lt;
        }
    } else {                                           // ===== epilogue: warp w reads TMEM lanes 32*(w%4)..+31
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int tx = row & ((1 << p.tw_log2) - 1);
        const int ty = (row >> p.tw_log2) & ((1 << p.th_log2) - 1);
        const int tn = row >> (p.tw_log2 + p.th_log2);
        uint32_t lt = 0, stage_sel = 0;
        uint8_t *my_stage = epi_stage + q * 2 * kStageBufBytes;
        for (int T = blockIdx.x; T < p.total_tiles; T += gridDim.x) {
            const TileCoord tc = decode_tile(p, T);
            if (tc.skip) continue;
            const ConvPhase &ph = p.ph[tc.phase];
            const uint32_t acc = lt & 1, acc_par = (lt >> 1) & 1;
            ++lt;
            float rgb[3];
            long long rgb_index;
            const bool valid = epilogue_tile<BLOCK_N, false, OP16>(p, out_maps, ph, tc, q, lane, tx, ty, tn, tmem_base + acc * BLOCK_N,
                                                      &tmem_full_bar[acc], acc_par, my_stage, stage_sel, rgb, rgb_index);
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);   // this warp no longer reads accumulator `acc`
            if (p.rgb_w && valid) {
                float *dst = p.rgb_out + rgb_index;
                atomicAdd(dst, rgb[0]); atomicAdd(dst + 1, rgb[1]); atomicAdd(dst + 2, rgb[2]);
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // all TMA stores of this warp landed
    }
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"((uint32_t)(2 * BLOCK_N)) : "memory");
    }
}


// ------------------------------------------------------------------------------------ CTA-pair variant (cta_group::2)
// Two CTAs of a cluster (one TPC) compute a 256-pixel x BLOCK_N tile with ONE tcgen05.mma.cta_group::2 stream issued
// by the leader: each CTA stages its own 128 pixel rows of A and HALF of the weight tile (BLOCK_N/2 rows), so the
// shared-memory operand traffic per SM drops from 128 to 96 B/clk at N = 128 (64 instead of 96 at N = 256) -- the limit
// the single-CTA kernel hits on the 128-channel layers.  Protocol: both producers' TMA loads (cta_group::2 form)
// complete on the LEADER's full barrier; the leader's tcgen05.commit multicasts to both CTAs' empty / tmem_full
// barriers; all eight epilogue warps arrive on the leader's tmem_empty barrier.  Only the leader arms a full barrier
// (one arrival + the bytes of BOTH CTAs): the peer's TMA bytes may land first, the transaction count then simply goes
// negative until the leader's expect_tx -- a per-stage remote mbarrier.arrive from the peer costs ~190 ns each.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;      // clears the CTA-rank bit of a shared::cluster address -> leader CTA

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_count_x() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t *bar) {      // arrive on the LEADER CTA's copy of `bar`
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" :: "r"(smem_u32(bar) & kPeerBitMask) : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tcgen05_commit_2sm(uint64_t *bar) {      // arrives on `bar` in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ TileCoord decode_tile_pair(const ConvKParams &p, int P, int rank) {
    TileCoord c;
    c.n_tile = P % p.n_tiles;
    int mt = 2 * (P / p.n_tiles);                               // first tile of the pair (pairs never straddle a phase)
    const int g = mt / p.tiles_per_group2;
    mt -= g * p.tiles_per_group2;
    c.phase = 0;
#pragma unroll
    for (int i = 1; i < kMaxPhases; ++i)
        if (i < p.num_phases && mt >= p.ph[i].tile_begin2) c.phase = i;
    const ConvPhase &ph = p.ph[c.phase];
    mt -= ph.tile_begin2;
    const int per_img = ph.tiles_x * ph.tiles_y;
    const int tn_first = g * p.group_n + mt / per_img;
    c.skip = tn_first >= p.tiles_n;                             // whole pair beyond the last image group
    mt += rank;                                                 // may run one past the real tiles of the phase ...
    const int tl = mt / per_img;                                // ... then tl == group_n: mask it (n0 >= batch)
    const int tx = mt % ph.tiles_x, ty = (mt / ph.tiles_x) % ph.tiles_y, tn = g * p.group_n + tl;
    c.gx0 = tx << p.tw_log2; c.gy0 = ty << p.th_log2;
    c.n0 = (tl >= p.group_n || tn >= p.tiles_n) ? p.batch : (tn << p.tn_log2);
    return c;
}

template <int BLOCK_N, int STAGES, bool OP16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kConvThreads, 1)
conv_igemm_tf32_2cta_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                       const __grid_constant__ ConvOutMaps out_maps,
                            const ConvKParams p)
{
    constexpr int BH_BYTES = (BLOCK_N / 2) * BLOCK_K * 4;                  // this CTA's half of the weight tile
    constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BLOCK_N >> 3) << 17) |
                                ((uint32_t)((2 * BLOCK_M) >> 4) << 24);     // M = 256 across the pair
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;
    uint8_t *sB = smem + STAGES * A_BYTES;
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + STAGES * (A_BYTES + BH_BYTES));
    uint64_t *empty_bar = full_bar + STAGES;
    uint64_t *tmem_full_bar = empty_bar + STAGES;      // [2]
    uint64_t *tmem_empty_bar = tmem_full_bar + 2;      // [2], used in the leader only
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty_bar + 2);
    uint8_t *epi_stage = smem + STAGES * (A_BYTES + BH_BYTES) + 1024;  // 1024-byte aligned (swizzle), 8 KB per epilogue warp
    float *vec_stage = reinterpret_cast<float *>(epi_stage + kEpiSmemBytes);   // [2][kVecKinds][BLOCK_N] (one-image tiles)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)(blockIdx.x & 1);            // __cluster_dims__(2,1,1)
    const int cluster = (int)(blockIdx.x >> 1), num_clusters = (int)(gridDim.x >> 1);

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_b) : "memory");
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster_sync_all();                                // barriers of both CTAs exist before any remote arrive
    if (warp == 1) {                                   // both CTAs, same warp id: 2 accumulators of BLOCK_N columns each
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(tmem_slot)), "r"((uint32_t)(2 * BLOCK_N)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {                                   // ===== TMA producer warp (both CTAs)
        uint32_t s = 0, par = 0;
        for (int P = cluster; P < p.total_pairs; P += num_clusters) {
            const TileCoord tc = decode_tile_pair(p, P, rank);
            if (tc.skip) continue;
            const ConvPhase &ph = p.ph[tc.phase];
            for (int t = 0; t < ph.num_taps; ++t) {
                const int cx = tc.gx0 * p.in_stride + ph.tap_dx[t], cy = tc.gy0 * p.in_stride + ph.tap_dy[t];
                for (int kc = 0; kc < p.kblocks_per_tap; ++kc) {
                    mbar_wait(&empty_bar[s], par ^ 1);
                    if (elect_one()) {
                        if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * (A_BYTES + BH_BYTES));   // bytes of BOTH CTAs
                        tma_load_4d_2sm(sA + s * A_BYTES, &tmap_a, &full_bar[s], kc * (OP16 ? 64 : 32), cx, cy, tc.n0);
                        tma_load_2d_2sm(sB + s * BH_BYTES, &tmap_b, &full_bar[s], ph.tap_k0[t] + kc * (OP16 ? 64 : 32),
                                        tc.n_tile * BLOCK_N + rank * (BLOCK_N / 2));
                    }
                    __syncwarp();
                    if (++s == STAGES) { s = 0; par ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {                               // ===== MMA warp (leader CTA only)
            uint32_t s = 0, par = 0, lt = 0;
            const uint32_t sA_u32 = smem_u32(sA), sB_u32 = smem_u32(sB);
            for (int P = cluster; P < p.total_pairs; P += num_clusters) {
                const TileCoord tc = decode_tile_pair(p, P, 0);
                if (tc.skip) continue;
                const int num_kb = p.ph[tc.phase].num_taps * p.kblocks_per_tap;
                const uint32_t acc = lt & 1, acc_par = (lt >> 1) & 1;
                mbar_wait(&tmem_empty_bar[acc], acc_par ^ 1);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[s], par);
                    tcgen05_fence_after();
                    const uint64_t da = make_kmajor_sw128_desc(sA_u32 + s * A_BYTES);
                    const uint64_t db = make_kmajor_sw128_desc(sB_u32 + s * BH_BYTES);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BLOCK_K / 8; ++k) {
                            if (p.debug & 2) continue;
                            if (OP16) umma_f16_2sm(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc_bf16(kIdesc), (kb | k) != 0);
                            else umma_tf32_2sm(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), kIdesc, (kb | k) != 0);
                        }
                        tcgen05_commit_2sm(&empty_bar[s]);
                    }
                    __syncwarp();
                    if (++s == STAGES) { s = 0; par ^= 1; }
                }
                if (elect_one()) tcgen05_commit_2sm(&tmem_full_bar[acc]);
                __syncwarp();
                ++lt;
            }
        }
    } else {                                           // ===== epilogue (both CTAs, each on its own 128 TMEM lanes)
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int tx = row & ((1 << p.tw_log2) - 1);
        const int ty = (row >> p.tw_log2) & ((1 << p.th_log2) - 1);
        const int tn = row >> (p.tw_log2 + p.th_log2);
        uint32_t lt = 0, stage_sel = 0;
        uint8_t *my_stage = epi_stage + q * 2 * kStageBufBytes;
        for (int P = cluster; P < p.total_pairs; P += num_clusters) {
            const TileCoord tc = decode_tile_pair(p, P, rank);
            if (tc.skip) continue;
            const ConvPhase &ph = p.ph[tc.phase];
            const uint32_t acc = lt & 1, acc_par = (lt >> 1) & 1;
            ++lt;
            float rgb[3];
            long long rgb_index;
            bool valid;
            if (p.tn_log2 == 0) {        // all rows of the tile belong to one image: per-tile vectors through shared memory
                float *vb = vec_stage + (acc & 1) * (kVecKinds * BLOCK_N);     // (ncu: the rowscale loads were the epilogue's stall)
                stage_tile_vectors<BLOCK_N, 128>(p, tc, vb, (int)threadIdx.x - 64);
                asm volatile("bar.sync 1, 128;" ::: "memory");
                valid = epilogue_tile<BLOCK_N, true, OP16>(p, out_maps, ph, tc, q, lane, tx, ty, tn, tmem_base + acc * BLOCK_N,
                                                     &tmem_full_bar[acc], acc_par, my_stage, stage_sel, rgb, rgb_index, vb);
            } else {
                valid = epilogue_tile<BLOCK_N, false, OP16>(p, out_maps, ph, tc, q, lane, tx, ty, tn, tmem_base + acc * BLOCK_N,
                                               &tmem_full_bar[acc], acc_par, my_stage, stage_sel, rgb, rgb_index);
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&tmem_empty_bar[acc]);
            if (p.rgb_w && valid) {
                float *dst = p.rgb_out + rgb_index;
                atomicAdd(dst, rgb[0]); atomicAdd(dst + 1, rgb[1]); atomicAdd(dst + 2, rgb[2]);
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    tcgen05_fence_before();
    cluster_sync_all();                                // nobody in the pair touches TMEM / remote barriers any more
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"((uint32_t)(2 * BLOCK_N)) : "memory");
    }
}

// ------------------------------------------------------------------------------------ halo variant of the CTA-pair kernel
// The im2col kernels above fetch one activation box per (tap, K block): every input pixel crosses the L2 -> SM path
// nine times for a 3x3 conv.  Measured (profiles/r1_ncu_full_summary.md): an SM sustains ~55 B/clk from L2, a
// 128-pixel x N=128 tile wants 94-128 B/clk, hence 52-56 % tensor-pipe utilisation on the 128-channel layers.
// Here a tile is 8 x 16 pixels of one image and the producer loads, per K block of 32 channels, ONE halo box
// {32 ch, 8 + rx, 16 + ry} (rows of 128 B, TMA 128-byte swizzle, zero fill outside the image).  Every tap is then a
// UMMA descriptor into that box: start address advanced by (dy * halo_w + dx) rows and stride-byte-offset = halo_w rows,
// so 8-row group g of the operand is tile row g shifted by the tap.  The swizzle XOR is a function of the absolute
// shared-memory address (profiles/r1_halo_descriptor_experiment.txt: any row shift and SBO = 1280 give exact results
// with base_offset = 0), so the shifted views read exactly what the TMA unit wrote.  Taps of a strided gather (dgrad of
// the stride-2 transposed conv) fall into up to four parity classes; each class has its own base-shifted tensor map and
// halo box ("group").  Operand bytes per SM and K block for a 3x3 conv drop from 9 x (16 + 8) KB to 22.5 + 9 x 8 KB.
constexpr int kHaloStageBytes = 23 * 1024;          // >= 10 x 18 rows x 128 B, multiple of the 1024-byte swizzle period
constexpr int kMaxGroups = 6;

struct HaloPhase {
    int num_groups;
    int g_map[kMaxGroups];                          // index into HaloMaps::a
    int g_sbo[kMaxGroups];                          // halo width in bytes (rows of 128 B) = UMMA stride byte offset
    int g_bytes[kMaxGroups];                        // box bytes (expect_tx)
    int g_xoff[kMaxGroups], g_yoff[kMaxGroups];     // box origin relative to the tile origin, in lattice units of the map
    int g_tap_begin[kMaxGroups], g_tap_end[kMaxGroups];
    int tap_k0[9];                                  // first K column of the tap in the weight matrix (regrouped order)
    int tap_aoff[9];                                // byte offset of the tap's first operand row inside the halo box
};
struct HaloParams { HaloPhase ph[kMaxPhases]; };
struct alignas(64) HaloMaps { CUtensorMap a[kMaxGroups]; };

__device__ __forceinline__ uint64_t make_kmajor_sw128_desc_sbo(uint32_t smem_addr, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// EPI_WARPS = 4 or 8 epilogue warps: with 8, two warps share each 32-lane TMEM quadrant and split the columns, which
// doubles the number of independent tcgen05.ld -> math -> TMA-store chains that drain an accumulator.
template <int BLOCK_N, int SA, int SB, int TPS, int EPI_WARPS, bool OP16>   // TPS = taps per weight stage (one barrier round trip per TPS taps)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 32 * EPI_WARPS, 1)
conv_halo_tf32_2cta_kernel(const __grid_constant__ HaloMaps amaps, const __grid_constant__ CUtensorMap tmap_b,
                           const __grid_constant__ ConvOutMaps out_maps, const ConvKParams p, const HaloParams hp)
{
    constexpr int BH_BYTES = (BLOCK_N / 2) * BLOCK_K * 4;
    constexpr int B_STAGE = TPS * BH_BYTES;
    constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BLOCK_N >> 3) << 17) |
                                ((uint32_t)((2 * BLOCK_M) >> 4) << 24);
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;
    uint8_t *sB = smem + SA * kHaloStageBytes;
    uint64_t *a_full = reinterpret_cast<uint64_t *>(smem + SA * kHaloStageBytes + SB * B_STAGE);
    uint64_t *a_empty = a_full + SA;
    uint64_t *b_full = a_empty + SA;
    uint64_t *b_empty = b_full + SB;
    uint64_t *tmem_full_bar = b_empty + SB;            // [2]
    uint64_t *tmem_empty_bar = tmem_full_bar + 2;      // [2], used in the leader only
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty_bar + 2);
    uint8_t *epi_stage = smem + SA * kHaloStageBytes + SB * B_STAGE + 1024;
    float *vec_stage = reinterpret_cast<float *>(epi_stage + EPI_WARPS * 2 * kStageBufBytes);   // [2][kVecKinds][BLOCK_N]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)(blockIdx.x & 1);            // __cluster_dims__(2,1,1): CTA rank in the pair
    const int cluster = (int)(blockIdx.x >> 1), num_clusters = (int)(gridDim.x >> 1);

    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < kMaxGroups; ++i) asm volatile("prefetch.tensormap [%0];" :: "l"(&amaps.a[i]) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_b) : "memory");
        for (int s = 0; s < SA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < SB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], 2 * EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(tmem_slot)), "r"((uint32_t)(2 * BLOCK_N)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {                                   // ===== TMA producer warp (both CTAs: own halo, own half of the weights)
        uint32_t sa = 0, pa = 0, sb = 0, pb = 0;
        for (int P = cluster; P < p.total_pairs; P += num_clusters) {
            const TileCoord tc = decode_tile_pair(p, P, rank);
            const HaloPhase &h = hp.ph[tc.phase];
            const int brow = tc.n_tile * BLOCK_N + rank * (BLOCK_N / 2);
            for (int kc = 0; kc < p.kblocks_per_tap; ++kc) {
                for (int g = 0; g < h.num_groups; ++g) {
                    mbar_wait_warp(&a_empty[sa], pa ^ 1, p.debug & 4);
                    if (elect_one()) {
                        if (rank == 0) mbar_expect_tx(&a_full[sa], 2 * (uint32_t)h.g_bytes[g]);
                        tma_load_4d_2sm(sA + sa * kHaloStageBytes, &amaps.a[h.g_map[g]], &a_full[sa], kc * (OP16 ? 64 : 32),
                                        tc.gx0 + h.g_xoff[g], tc.gy0 + h.g_yoff[g], tc.n0);
                    }
                    __syncwarp();
                    if (++sa == SA) { sa = 0; pa ^= 1; }
                    const int te = h.g_tap_end[g];
                    for (int t = h.g_tap_begin[g]; t < te; t += TPS) {
                        const int n = min(TPS, te - t);
                        mbar_wait_warp(&b_empty[sb], pb ^ 1, p.debug & 4);
                        if (elect_one()) {
                            if (rank == 0) mbar_expect_tx(&b_full[sb], 2 * (uint32_t)n * BH_BYTES);
#pragma unroll
                            for (int j = 0; j < TPS; ++j)
                                if (j < n)
                                    tma_load_2d_2sm(sB + sb * B_STAGE + j * BH_BYTES, &tmap_b, &b_full[sb],
                                                    h.tap_k0[t + j] + kc * (OP16 ? 64 : 32), brow);
                        }
                        __syncwarp();
                        if (++sb == SB) { sb = 0; pb ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {                               // ===== MMA warp (leader CTA only), one elected lane issues
            uint32_t sa = 0, pa = 0, sb = 0, pb = 0, lt = 0;
            const uint32_t sA_u32 = smem_u32(sA), sB_u32 = smem_u32(sB);
            for (int P = cluster; P < p.total_pairs; P += num_clusters, ++lt) {
                const TileCoord tc = decode_tile_pair(p, P, 0);
                const HaloPhase &h = hp.ph[tc.phase];
                const uint32_t acc = lt & 1, acc_par = (lt >> 1) & 1;
                mbar_wait_warp(&tmem_empty_bar[acc], acc_par ^ 1, p.debug & 4);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
                uint32_t accumulate = 0;
                for (int kc = 0; kc < p.kblocks_per_tap; ++kc) {
                    for (int g = 0; g < h.num_groups; ++g) {
                        mbar_wait_warp(&a_full[sa], pa, p.debug & 4);
                        tcgen05_fence_after();
                        const uint32_t a_base = sA_u32 + sa * kHaloStageBytes;
                        const uint32_t sbo = (uint32_t)h.g_sbo[g];
                        const int te = h.g_tap_end[g];
                        for (int t = h.g_tap_begin[g]; t < te; t += TPS) {
                            const int n = min(TPS, te - t);
                            mbar_wait_warp(&b_full[sb], pb, p.debug & 4);
                            tcgen05_fence_after();
                            const uint32_t b_base = sB_u32 + sb * B_STAGE;
                            if (elect_one()) {
#pragma unroll
                                for (int j = 0; j < TPS; ++j) {
                                    if (j < n) {
                                        const uint64_t da = make_kmajor_sw128_desc_sbo(a_base + (uint32_t)h.tap_aoff[t + j], sbo);
                                        const uint64_t db = make_kmajor_sw128_desc(b_base + j * BH_BYTES);
#pragma unroll
                                        for (int k = 0; k < BLOCK_K / 8; ++k) {
                                            if (!(p.debug & 2)) {
                                                if (OP16) umma_f16_2sm(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc_bf16(kIdesc), accumulate);
                                                else umma_tf32_2sm(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), kIdesc, accumulate);
                                            }
                                            accumulate = 1;
                                        }
                                    }
                                }
                                tcgen05_commit_2sm(&b_empty[sb]);
                            }
                            __syncwarp();
                            accumulate = 1;
                            if (++sb == SB) { sb = 0; pb ^= 1; }
                        }
                        if (elect_one()) tcgen05_commit_2sm(&a_empty[sa]);   // halo box free once all of its taps have been read
                        __syncwarp();
                        if (++sa == SA) { sa = 0; pa ^= 1; }
                    }
                }
                if (elect_one()) tcgen05_commit_2sm(&tmem_full_bar[acc]);
                __syncwarp();
            }
        }
    } else {                                           // ===== epilogue (both CTAs, each on its own 128 TMEM lanes)
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int tx = row & ((1 << p.tw_log2) - 1);
        const int ty = (row >> p.tw_log2) & ((1 << p.th_log2) - 1);
        const int tn = row >> (p.tw_log2 + p.th_log2);
        uint32_t lt = 0, stage_sel = 0;
        const int half = (warp - 2) >> 2;                                   // which half of the columns (EPI_WARPS = 8)
        constexpr int CHUNKS = BLOCK_N / 32 / (EPI_WARPS / 4);
        uint8_t *my_stage = epi_stage + (warp - 2) * 2 * kStageBufBytes;
        for (int P = cluster; P < p.total_pairs; P += num_clusters, ++lt) {
            const TileCoord tc = decode_tile_pair(p, P, rank);
            const ConvPhase &ph = p.ph[tc.phase];
            const uint32_t acc = lt & 1, acc_par = (lt >> 1) & 1;
            float rgb[3];
            long long rgb_index;
            // per-tile vectors -> shared memory (double buffered), published to the 4 epilogue warps by a named barrier;
            // the global loads overlap the wait for the accumulator
            float *vb = vec_stage + (lt & 1) * (kVecKinds * BLOCK_N);
            stage_tile_vectors<BLOCK_N, 32 * EPI_WARPS>(p, tc, vb, (int)threadIdx.x - 64);
            asm volatile("bar.sync 1, %0;" :: "n"(32 * EPI_WARPS) : "memory");
            const bool valid = epilogue_tile<BLOCK_N, true, OP16>(p, out_maps, ph, tc, q, lane, tx, ty, tn, tmem_base + acc * BLOCK_N,
                                                            &tmem_full_bar[acc], acc_par, my_stage, stage_sel, rgb, rgb_index, vb,
                                                            half * CHUNKS, (half + 1) * CHUNKS);
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&tmem_empty_bar[acc]);
            if (p.rgb_w && valid) {
                float *dst = p.rgb_out + rgb_index;
                atomicAdd(dst, rgb[0]); atomicAdd(dst + 1, rgb[1]); atomicAdd(dst + 2, rgb[2]);
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    tcgen05_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"((uint32_t)(2 * BLOCK_N)) : "memory");
    }
}

// ------------------------------------------------------------------------------------ wgrad
// dW[co, t, ci] += sum over (image, pixel) of G[n, gy*ga + gdy_t, gx*ga + gdx_t, co] * X[n, gy*xa + xdy_t, gx*xa + xdx_t, ci]
// i.e. D = A^T B with the reduction (K) running over pixels.  Both operands are read straight from the NHWC
// tensors, so they are MN-major for the tensor core: a TMA box {32 channels, TW, TH, TN} lands as 64 K-rows of
// 128 bytes (two 4-row swizzle atoms per K=8 instruction); M = 128 output channels = 4 such column blocks, LBO apart.
// Split-K over pixel tiles (gridDim.z) with fp32 atomic accumulation into the zero-initialised result.
// Two shapes: N = 256 input channels x 32 pixels per stage (4 stages; 96 B/clk of shared-memory operand traffic) when
// cin % 256 == 0, else N = 128 x 64 pixels per stage (3 stages).

struct WgradKParams {
    int tw_log2, th_log2, tn_log2;                     // K tile = 64 pixels
    int tiles_x, tiles_y, tiles_n;
    int ktiles_per_split;
    int n_tiles;                                       // cin / WG_N
    int g_stride, x_stride;
    int g_dy[9], g_dx[9], x_dy[9], x_dx[9], tap_out[9];
    int taps_total, cin;
    float *dw;                                         // [cout][taps_total][cin]
    int op16;                                          // 1: bf16 operands (64 channels per 128-byte column block, K = 16 pixels
                                                       // per instruction, plain 128-byte swizzle); 0: tf32 in fp32 words
};

// MN-major tf32 operands exist in one shared-memory layout only: 128-byte swizzle with 32-byte atomicity
// (cute::UMMA::LayoutType::SWIZZLE_128B_BASE32B = 1, TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): atoms of
// 4 K-rows x 128 B; LBO = distance to the next 32-element group along M/N, SBO = distance to the next 4 K-rows.
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
    d |= (uint64_t)(lbo_bytes >> 4) << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;
    return d;
}

// 16-bit MN-major operands use the plain 128-byte swizzle (cute::UMMA::LayoutType::SWIZZLE_128B = 2, TMA
// CU_TENSOR_MAP_SWIZZLE_128B): atoms of 8 K-rows x 128 B (64 channels); LBO = distance to the next 64-channel group along
// M/N, SBO = distance to the next 8 K-rows.
__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc16(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
    d |= (uint64_t)(lbo_bytes >> 4) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

template <int WG_N, int WG_BK, int WG_STAGES>
__global__ void __launch_bounds__(kConvThreads, 1)
wgrad_tf32_kernel(const __grid_constant__ CUtensorMap tmap_g, const __grid_constant__ CUtensorMap tmap_x,
                  const WgradKParams p)
{
    constexpr int WG_COLBLK_BYTES = WG_BK * 128;           // one 32-channel column block of a stage
    constexpr int NB = WG_N / 32;                          // column blocks of the B (activation) operand
    constexpr int WG_STAGE_BYTES = (4 + NB) * WG_COLBLK_BYTES;
    constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                                ((uint32_t)(WG_N >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);   // MN-major A and B
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + WG_STAGES * WG_STAGE_BYTES);
    uint64_t *empty_bar = full_bar + WG_STAGES;
    uint64_t *tmem_full_bar = empty_bar + WG_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tile = blockIdx.x / p.n_tiles, n_tile = blockIdx.x % p.n_tiles;
    const int tap = blockIdx.y;
    const int total_kt = p.tiles_x * p.tiles_y * p.tiles_n;
    const int kt0 = blockIdx.z * p.ktiles_per_split;
    const int kt1 = min(total_kt, kt0 + p.ktiles_per_split);
    const int num_kt = kt1 - kt0;                      // >= 1 by construction of the grid

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_g) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_x) : "memory");
        for (int s = 0; s < WG_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(tmem_slot)), "r"((uint32_t)WG_N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < num_kt; ++it) {
                const int s = it % WG_STAGES;
                const uint32_t ph = (it / WG_STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                const int kt = kt0 + it;
                const int tile_x = kt % p.tiles_x, tile_y = (kt / p.tiles_x) % p.tiles_y, tile_n = kt / (p.tiles_x * p.tiles_y);
                const int gx0 = tile_x << p.tw_log2, gy0 = tile_y << p.th_log2, n0 = tile_n << p.tn_log2;
                uint8_t *sa = smem + s * WG_STAGE_BYTES, *sb = sa + 4 * WG_COLBLK_BYTES;
                const int cb = p.op16 ? 64 : 32, gb = p.op16 ? 2 : 4, xb = p.op16 ? NB / 2 : NB;   // channels per block, blocks
                mbar_expect_tx(&full_bar[s], (gb + xb) * WG_COLBLK_BYTES);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (j < gb)
                        tma_load_4d(sa + j * WG_COLBLK_BYTES, &tmap_g, &full_bar[s], m_tile * BLOCK_M + j * cb,
                                    gx0 * p.g_stride + p.g_dx[tap], gy0 * p.g_stride + p.g_dy[tap], n0);
#pragma unroll
                for (int j = 0; j < NB; ++j)
                    if (j < xb)
                        tma_load_4d(sb + j * WG_COLBLK_BYTES, &tmap_x, &full_bar[s], n_tile * WG_N + j * cb,
                                    gx0 * p.x_stride + p.x_dx[tap], gy0 * p.x_stride + p.x_dy[tap], n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            for (int it = 0; it < num_kt; ++it) {
                const int s = it % WG_STAGES;
                const uint32_t ph = (it / WG_STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                tcgen05_fence_after();
                const uint32_t a0 = smem_u32(smem + s * WG_STAGE_BYTES), b0 = a0 + 4 * WG_COLBLK_BYTES;
                if (p.op16) {
#pragma unroll
                    for (int k = 0; k < WG_BK / 16; ++k)   // 16 pixels (two 8-row atoms) per instruction
                        umma_f16(tmem_base, make_mnmajor_sw128_desc16(a0 + k * 2048, WG_COLBLK_BYTES),
                                 make_mnmajor_sw128_desc16(b0 + k * 2048, WG_COLBLK_BYTES), idesc_bf16(kIdesc), (it | k) != 0);
                } else {
#pragma unroll
                    for (int k = 0; k < WG_BK / 8; ++k)    // 8 pixels (one swizzle atom of K rows) per instruction
                        umma_tf32(tmem_base, make_mnmajor_sw128_desc(a0 + k * 1024, WG_COLBLK_BYTES),
                                  make_mnmajor_sw128_desc(b0 + k * 1024, WG_COLBLK_BYTES), kIdesc, (it | k) != 0);
                }
                tcgen05_commit(&empty_bar[s]);
            }
            tcgen05_commit(tmem_full_bar);
        }
    } else {
        const int q = warp & 3;
        const int co = m_tile * BLOCK_M + q * 32 + lane;
        float *dst = p.dw + ((long long)co * p.taps_total + p.tap_out[tap]) * p.cin + n_tile * WG_N;
        mbar_wait(tmem_full_bar, 0);
        tcgen05_fence_after();
#pragma unroll 1
        for (int c = 0; c < WG_N / 32; ++c) {
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
#pragma unroll
            for (int j = 0; j < 32; j += 4)            // 128-bit vector reduction: a quarter of the L2 transactions
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(dst + c * 32 + j), "f"(__uint_as_float(r[j])),
                             "f"(__uint_as_float(r[j + 1])), "f"(__uint_as_float(r[j + 2])), "f"(__uint_as_float(r[j + 3])) : "memory");
        }
        tcgen05_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"((uint32_t)WG_N) : "memory");
    }
}


// CTA-pair wgrad: the pair accumulates a 256 (cout) x 256 (cin) block of one tap; each CTA stages its own 128 output
// channels of G and 128 of the 256 input channels of X (64 KB per 64-pixel stage instead of the 96 KB a single CTA would
// need for the same flops), one tcgen05.mma.cta_group::2 stream, barrier protocol as in conv_igemm_tf32_2cta_kernel.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kConvThreads, 1)
wgrad_tf32_2cta_kernel(const __grid_constant__ CUtensorMap tmap_g, const __grid_constant__ CUtensorMap tmap_x,
                       const WgradKParams p)
{
    constexpr int WG_BK = 64, WG_STAGES = 3, WG_N = 256;
    constexpr int WG_COLBLK_BYTES = WG_BK * 128;
    constexpr int WG_STAGE_BYTES = 8 * WG_COLBLK_BYTES;                  // 4 blocks of G + 4 blocks of X per CTA
    constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                                ((uint32_t)(WG_N >> 3) << 17) | ((uint32_t)((2 * BLOCK_M) >> 4) << 24);
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + WG_STAGES * WG_STAGE_BYTES);
    uint64_t *empty_bar = full_bar + WG_STAGES;
    uint64_t *tmem_full_bar = empty_bar + WG_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)(blockIdx.x & 1);            // __cluster_dims__(2,1,1)
    const int pair = blockIdx.x >> 1;
    const int m_tile = pair / p.n_tiles, n_tile = pair % p.n_tiles;      // in units of 256 channels
    const int tap = blockIdx.y;
    const int total_kt = p.tiles_x * p.tiles_y * p.tiles_n;
    const int kt0 = blockIdx.z * p.ktiles_per_split;
    const int num_kt = min(total_kt, kt0 + p.ktiles_per_split) - kt0;

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_g) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_x) : "memory");
        for (int s = 0; s < WG_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(tmem_slot)), "r"((uint32_t)WG_N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < num_kt; ++it) {
                const int s = it % WG_STAGES;
                const uint32_t ph = (it / WG_STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                const int kt = kt0 + it;
                const int tile_x = kt % p.tiles_x, tile_y = (kt / p.tiles_x) % p.tiles_y, tile_n = kt / (p.tiles_x * p.tiles_y);
                const int gx0 = tile_x << p.tw_log2, gy0 = tile_y << p.th_log2, n0 = tile_n << p.tn_log2;
                uint8_t *sa = smem + s * WG_STAGE_BYTES, *sb = sa + 4 * WG_COLBLK_BYTES;
                const int cb = p.op16 ? 64 : 32, nb = p.op16 ? 2 : 4;                     // channels per block, blocks per operand
                if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * 2 * nb * WG_COLBLK_BYTES);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (j >= nb) break;
                    tma_load_4d_2sm(sa + j * WG_COLBLK_BYTES, &tmap_g, &full_bar[s], m_tile * 256 + rank * 128 + j * cb,
                                    gx0 * p.g_stride + p.g_dx[tap], gy0 * p.g_stride + p.g_dy[tap], n0);
                    tma_load_4d_2sm(sb + j * WG_COLBLK_BYTES, &tmap_x, &full_bar[s], n_tile * 256 + rank * 128 + j * cb,
                                    gx0 * p.x_stride + p.x_dx[tap], gy0 * p.x_stride + p.x_dy[tap], n0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            for (int it = 0; it < num_kt; ++it) {
                const int s = it % WG_STAGES;
                const uint32_t ph = (it / WG_STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                tcgen05_fence_after();
                const uint32_t a0 = smem_u32(smem + s * WG_STAGE_BYTES), b0 = a0 + 4 * WG_COLBLK_BYTES;
                if (p.op16) {
#pragma unroll
                    for (int k = 0; k < WG_BK / 16; ++k)
                        umma_f16_2sm(tmem_base, make_mnmajor_sw128_desc16(a0 + k * 2048, WG_COLBLK_BYTES),
                                     make_mnmajor_sw128_desc16(b0 + k * 2048, WG_COLBLK_BYTES), idesc_bf16(kIdesc), (it | k) != 0);
                } else {
#pragma unroll
                    for (int k = 0; k < WG_BK / 8; ++k)
                        umma_tf32_2sm(tmem_base, make_mnmajor_sw128_desc(a0 + k * 1024, WG_COLBLK_BYTES),
                                      make_mnmajor_sw128_desc(b0 + k * 1024, WG_COLBLK_BYTES), kIdesc, (it | k) != 0);
                }
                tcgen05_commit_2sm(&empty_bar[s]);
            }
            tcgen05_commit_2sm(tmem_full_bar);
        }
    } else {
        const int q = warp & 3;
        const int co = m_tile * 256 + rank * 128 + q * 32 + lane;
        float *dst = p.dw + ((long long)co * p.taps_total + p.tap_out[tap]) * p.cin + n_tile * WG_N;
        mbar_wait(tmem_full_bar, 0);
        tcgen05_fence_after();
#pragma unroll 1
        for (int c = 0; c < WG_N / 32; ++c) {
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(dst + c * 32 + j), "f"(__uint_as_float(r[j])),
                             "f"(__uint_as_float(r[j + 1])), "f"(__uint_as_float(r[j + 2])), "f"(__uint_as_float(r[j + 3])) : "memory");
        }
    }
    tcgen05_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"((uint32_t)WG_N) : "memory");
    }
}

// CTA-pair wgrad for 128-wide channel blocks: the pair accumulates TWO taps of a 128 (cout) x 128 (cin) block with one
// M = 256 x N = 128 tcgen05.mma.cta_group::2 stream.  The M halves are the two taps: CTA r stages the 128 output channels
// of G shifted by ITS tap and 64 of the 128 input channels of the (common, unshifted) X tile.  Per SM that is 96 instead
// of 128 B/clk of shared-memory operand reads and 48 instead of 64 KB of TMA traffic per 64-pixel stage -- the two limits
// of the single-CTA N = 128 kernel (ncu: tensor pipe 60 %).  All taps must read X at the same offset; the host moves the
// tap shift of a plain 3x3 conv from X to G (sum over x-pixels of g[p - tap] * x[p], zero fill outside G).
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kConvThreads, 1)
wgrad_tf32_2cta_taps_kernel(const __grid_constant__ CUtensorMap tmap_g, const __grid_constant__ CUtensorMap tmap_x,
                            const WgradKParams p, int num_taps)
{
    constexpr int WG_BK = 64, WG_STAGES = 4, WG_N = 128;
    constexpr int WG_COLBLK_BYTES = WG_BK * 128;
    constexpr int WG_STAGE_BYTES = 6 * WG_COLBLK_BYTES;                  // 4 blocks of G (128 co) + 2 blocks of X (64 ci) per CTA
    constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                                ((uint32_t)(WG_N >> 3) << 17) | ((uint32_t)((2 * BLOCK_M) >> 4) << 24);
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + WG_STAGES * WG_STAGE_BYTES);
    uint64_t *empty_bar = full_bar + WG_STAGES;
    uint64_t *tmem_full_bar = empty_bar + WG_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)(blockIdx.x & 1);
    const int blk = blockIdx.x >> 1;
    const int m_tile = blk / p.n_tiles, n_tile = blk % p.n_tiles;        // in units of 128 channels
    const int tap_raw = 2 * blockIdx.y + rank;
    const bool tap_valid = tap_raw < num_taps;
    const int tap = tap_valid ? tap_raw : num_taps - 1;                   // odd tap count: the last pair's second half idles
    const int total_kt = p.tiles_x * p.tiles_y * p.tiles_n;
    const int kt0 = blockIdx.z * p.ktiles_per_split;
    const int num_kt = min(total_kt, kt0 + p.ktiles_per_split) - kt0;

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_g) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap_x) : "memory");
        for (int s = 0; s < WG_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(tmem_slot)), "r"((uint32_t)WG_N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < num_kt; ++it) {
                const int s = it % WG_STAGES;
                const uint32_t ph = (it / WG_STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                const int kt = kt0 + it;
                const int tile_x = kt % p.tiles_x, tile_y = (kt / p.tiles_x) % p.tiles_y, tile_n = kt / (p.tiles_x * p.tiles_y);
                const int gx0 = tile_x << p.tw_log2, gy0 = tile_y << p.th_log2, n0 = tile_n << p.tn_log2;
                uint8_t *sa = smem + s * WG_STAGE_BYTES, *sb = sa + 4 * WG_COLBLK_BYTES;
                const int cb = p.op16 ? 64 : 32, gb = p.op16 ? 2 : 4, xb = p.op16 ? 1 : 2;      // channels per block, blocks of G / X
                if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * (gb + xb) * WG_COLBLK_BYTES);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (j < gb)
                        tma_load_4d_2sm(sa + j * WG_COLBLK_BYTES, &tmap_g, &full_bar[s], m_tile * 128 + j * cb,
                                        gx0 * p.g_stride + p.g_dx[tap], gy0 * p.g_stride + p.g_dy[tap], n0);
#pragma unroll
                for (int j = 0; j < 2; ++j)
                    if (j < xb)
                        tma_load_4d_2sm(sb + j * WG_COLBLK_BYTES, &tmap_x, &full_bar[s], n_tile * 128 + rank * 64 + j * cb,
                                        gx0 * p.x_stride + p.x_dx[0], gy0 * p.x_stride + p.x_dy[0], n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            for (int it = 0; it < num_kt; ++it) {
                const int s = it % WG_STAGES;
                const uint32_t ph = (it / WG_STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                tcgen05_fence_after();
                const uint32_t a0 = smem_u32(smem + s * WG_STAGE_BYTES), b0 = a0 + 4 * WG_COLBLK_BYTES;
                if (p.op16) {
#pragma unroll
                    for (int k = 0; k < WG_BK / 16; ++k)
                        umma_f16_2sm(tmem_base, make_mnmajor_sw128_desc16(a0 + k * 2048, WG_COLBLK_BYTES),
                                     make_mnmajor_sw128_desc16(b0 + k * 2048, WG_COLBLK_BYTES), idesc_bf16(kIdesc), (it | k) != 0);
                } else {
#pragma unroll
                    for (int k = 0; k < WG_BK / 8; ++k)
                        umma_tf32_2sm(tmem_base, make_mnmajor_sw128_desc(a0 + k * 1024, WG_COLBLK_BYTES),
                                      make_mnmajor_sw128_desc(b0 + k * 1024, WG_COLBLK_BYTES), kIdesc, (it | k) != 0);
                }
                tcgen05_commit_2sm(&empty_bar[s]);
            }
            tcgen05_commit_2sm(tmem_full_bar);
        }
    } else {
        const int q = warp & 3;
        const int co = m_tile * 128 + q * 32 + lane;
        float *dst = p.dw + ((long long)co * p.taps_total + p.tap_out[tap]) * p.cin + n_tile * WG_N;
        mbar_wait(tmem_full_bar, 0);
        tcgen05_fence_after();
        if (tap_valid) {
#pragma unroll 1
            for (int c = 0; c < WG_N / 32; ++c) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(dst + c * 32 + j), "f"(__uint_as_float(r[j])),
                                 "f"(__uint_as_float(r[j + 1])), "f"(__uint_as_float(r[j + 2])), "f"(__uint_as_float(r[j + 3])) : "memory");
            }
        }
    }
    tcgen05_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"((uint32_t)WG_N) : "memory");
    }
}

// ------------------------------------------------------------------------------------ small helper kernels
// xs[b,p,c] = tf32(x[b,p,c] * s[b,c])  -- the modulated, tensor-core-ready copy of an NHWC activation
template <bool BF16>
__global__ void __launch_bounds__(256)
modulate_tf32_kernel(float *__restrict__ xs, const float *__restrict__ x, const float *__restrict__ s,
                     uint32_t n4, FastDiv c4_div, FastDiv img4_div, int c4)
{
    const uint32_t stride = gridDim.x * 256;
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n4; i += stride) {
        uint32_t q, cq, b, rem;
        c4_div.divmod(i, q, cq);                      // channel quad inside the pixel
        img4_div.divmod(i, b, rem);                   // image index
        const float4 v = ld_stream4(x + 4ull * i);
        const float4 m = s ? __ldg(reinterpret_cast<const float4 *>(s) + (size_t)b * c4 + cq) : make_float4(1.f, 1.f, 1.f, 1.f);
        if (BF16) {                                   // xs is a bfloat16 tensor: 4 channels = 8 bytes
            reinterpret_cast<uint2 *>(xs)[i] = make_uint2(pack_bf16(v.x * m.x, v.y * m.y), pack_bf16(v.z * m.z, v.w * m.w));
        } else {
            float4 o = make_float4(round_tf32(v.x * m.x), round_tf32(v.y * m.y), round_tf32(v.z * m.z), round_tf32(v.w * m.w));
            *reinterpret_cast<float4 *>(xs + 4ull * i) = o;
        }
    }
}

// out = (a + b) * scale (fp32) and, optionally, its GEMM-operand copy (tf32-rounded fp32 or bfloat16) in one pass: the
// residual sum of a Discriminator ResBlock (reference layers.py:390) feeding the next block's first conv
template <bool BF16>
__global__ void __launch_bounds__(256)
residual_combine_kernel(float *__restrict__ out, float *__restrict__ op, const float *__restrict__ a, const float *__restrict__ b,
                        float scale, uint32_t n4)
{
    const uint32_t stride = gridDim.x * 256;
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n4; i += stride) {
        const float4 u = ld_stream4(a + 4ull * i), v = ld_stream4(b + 4ull * i);
        const float4 o = make_float4((u.x + v.x) * scale, (u.y + v.y) * scale, (u.z + v.z) * scale, (u.w + v.w) * scale);
        *reinterpret_cast<float4 *>(out + 4ull * i) = o;
        if (op) {
            if (BF16) reinterpret_cast<uint2 *>(op)[i] = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
            else *reinterpret_cast<float4 *>(op + 4ull * i) = make_float4(round_tf32(o.x), round_tf32(o.y), round_tf32(o.z), round_tf32(o.w));
        }
    }
}

// Weight re-layout: reference [cout, cin, kh, kw] -> GEMM B operand [rows][taps][cols], scaled and rounded to tf32.
//   transpose = 0: rows = cout, cols = cin, tap t = ky*kw + kx                         (forward)
//   transpose = 1: rows = cin, cols = cout, tap t = (kh-1-ky)*kw + (kw-1-kx)           (dgrad of the plain conv)
//   transpose = 2: rows = cin, cols = cout, tap t = ky*kw + kx                         (transposed conv forward)
//   transpose = 3: rows = cout, cols = cin, tap t = ky*kw + kx (same as 0)             (dgrad of the transposed conv)
__global__ void __launch_bounds__(256)
weight_prep_kernel(float *__restrict__ dst, const float *__restrict__ w, float scale, int cout, int cin, int kh, int kw,
                   int transpose)
{
    const int taps = kh * kw;
    const int64_t total = (int64_t)cout * cin * taps;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        int64_t r = i;
        const int kx = (int)(r % kw); r /= kw;
        const int ky = (int)(r % kh); r /= kh;
        const int ci = (int)(r % cin);
        const int co = (int)(r / cin);
        const float v = round_tf32(w[i] * scale);
        int64_t o;
        if (transpose == 1) o = ((int64_t)ci * taps + ((kh - 1 - ky) * kw + (kw - 1 - kx))) * cout + co;
        else if (transpose == 2) o = ((int64_t)ci * taps + (ky * kw + kx)) * cout + co;
        else o = ((int64_t)co * taps + (ky * kw + kx)) * cin + ci;
        dst[o] = v;
    }
}

// One pass over a 3x3 weight: forward GEMM layout, transposed (dgrad) layout and the demodulation statistic together.
//   fwd[co][t][ci]  = tf32(scale * w[co][ci][t])
//   tr [ci][t'][co] = tf32(scale * w[co][ci][t]),  t' = taps-1-t (flip = 1, dgrad of the plain conv) or t (flip = 0)
//   wsq[co][ci]     = scale^2 * sum_t w[co][ci][t]^2                                   (any of the three may be NULL)
// A CTA owns a 32 x 32 (co, ci) block: the read is one contiguous 32*taps-float run per co, both writes are 128-byte
// rows (through a shared-memory transpose); the per-mode kernel above scatters 4-byte stores for the transposed layout.
constexpr int kWP = 32;
__device__ __forceinline__ void store_operand(float *base, int64_t idx, float v, bool bf16) {
    if (bf16) {
        uint16_t h;
        asm("cvt.rn.bf16.f32 %0, %1;" : "=h"(h) : "f"(v));
        reinterpret_cast<uint16_t *>(base)[idx] = h;
    } else {
        base[idx] = round_tf32(v);
    }
}

__device__ __forceinline__ void weight_prep_dual_body(float *__restrict__ fwd, float *__restrict__ tr, float *__restrict__ wsq,
                                                      const float *__restrict__ w, float scale, int cout, int cin, int taps,
                                                      int flip, int bf16, int bx, int by, float *wt)
{
    const int row = kWP * taps + 1;
    const int co0 = by * kWP, ci0 = bx * kWP;
    for (int idx = threadIdx.x; idx < kWP * kWP * taps; idx += 256) {
        const int r = idx / (kWP * taps), c = idx - r * (kWP * taps);          // c = ci_local * taps + t
        const int co = co0 + r, ci = ci0 + c / taps;
        wt[r * row + c] = (co < cout && ci < cin) ? __ldg(w + ((int64_t)co * cin + ci0) * taps + c) * scale : 0.0f;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (fwd) {                                            // rows (co, t), 32 consecutive ci
        for (int q = wid; q < kWP * taps; q += 8) {
            const int r = q / taps, t = q - r * taps;
            if (co0 + r < cout && ci0 + lane < cin)
                store_operand(fwd, ((int64_t)(co0 + r) * taps + t) * cin + ci0 + lane, wt[r * row + lane * taps + t], bf16 != 0);
        }
    }
    if (tr) {                                             // rows (ci, t'), 32 consecutive co
        for (int q = wid; q < kWP * taps; q += 8) {
            const int c = q / taps, t = q - c * taps;
            const int tt = flip ? taps - 1 - t : t;
            if (ci0 + c < cin && co0 + lane < cout)
                store_operand(tr, ((int64_t)(ci0 + c) * taps + tt) * cout + co0 + lane, wt[lane * row + c * taps + t], bf16 != 0);
        }
    }
    if (wsq) {
        for (int q = threadIdx.x; q < kWP * kWP; q += 256) {
            const int r = q >> 5, c = q & 31;
            float acc = 0.0f;
            for (int t = 0; t < taps; ++t) { const float v = wt[r * row + c * taps + t]; acc = fmaf(v, v, acc); }
            if (co0 + r < cout && ci0 + c < cin) wsq[(int64_t)(co0 + r) * cin + ci0 + c] = acc;
        }
    }
}

__global__ void __launch_bounds__(256)
weight_prep_dual_kernel(float *__restrict__ fwd, float *__restrict__ tr, float *__restrict__ wsq, const float *__restrict__ w,
                        float scale, int cout, int cin, int taps, int flip, int bf16)
{
    extern __shared__ float wt[];                         // [32 co][32 ci * taps + 1]
    weight_prep_dual_body(fwd, tr, wsq, w, scale, cout, cin, taps, flip, bf16, blockIdx.x, blockIdx.y, wt);
}

// every conv weight of a network in ONE launch (blockIdx.z = layer): 13 launches of 256 CTAs each were latency bound
// (28 us for 28 MB); together they fill the machine
struct WeightPrepTable { int n; sr_weight_prep_item item[SR_WEIGHT_PREP_MAX]; };

__global__ void __launch_bounds__(256)
weight_prep_multi_kernel(const __grid_constant__ WeightPrepTable tab, int bf16)
{
    extern __shared__ float wt[];
    const sr_weight_prep_item &it = tab.item[blockIdx.z];
    if ((int)blockIdx.x * kWP >= it.cin || (int)blockIdx.y * kWP >= it.cout) return;
    weight_prep_dual_body(reinterpret_cast<float *>(it.fwd), reinterpret_cast<float *>(it.tr), it.wsq, it.w, it.scale, it.cout, it.cin,
                          it.taps, it.flip_transposed, bf16, blockIdx.x, blockIdx.y, wt);
}

__global__ void __launch_bounds__(256)
weight_sq_backward_multi_kernel(const __grid_constant__ WeightPrepTable tab)
{
    const sr_weight_prep_item &it = tab.item[blockIdx.y];
    const uint32_t total = (uint32_t)it.cout * it.cin * it.taps;
    const float s2 = 2.0f * it.scale * it.scale;
    for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < total; i += gridDim.x * 256u)
        it.gw[i] = s2 * __ldg(it.w + i) * __ldg(it.g_wsq + i / (uint32_t)it.taps);
}

// ------------------------------------------------------------------------------------ host side
int ilog2_exact(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

// OP16 (bf16 operands, tcgen05 kind::f16) is a COMPILE-TIME parameter of the conv kernels: as a run-time branch around every
// MMA it cost the tf32 path a quarter of its speed (ncu, round 2: 128 -> 128 @ 256^2 at 56 % tensor pipe instead of 87 %,
// 512 -> 512 @ 64^2 at 72 % instead of 98 % -- the issuing thread bounds these loops).
template <int BLOCK_N, int STAGES, bool OP16>
int launch_conv(const CUtensorMap &ta, const CUtensorMap &tb, const ConvOutMaps &om, const ConvKParams &p, cudaStream_t st)
{
    constexpr int B_BYTES = BLOCK_N * BLOCK_K * 4;
    const size_t smem = 1024 + (size_t)STAGES * (A_BYTES + B_BYTES) + 1024 + kEpiSmemBytes;
    auto kern = conv_igemm_tf32_kernel<BLOCK_N, STAGES, OP16>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("conv: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
        configured = true;
    }
    const int grid = p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs;     // persistent: one CTA per SM
    kern<<<grid, kConvThreads, smem, st>>>(ta, tb, om, p);
    return SR_OK;
}

template <int BLOCK_N, int STAGES, bool OP16>
int launch_conv_2cta(const CUtensorMap &ta, const CUtensorMap &tb, const ConvOutMaps &om, const ConvKParams &p, cudaStream_t st)
{
    constexpr int BH_BYTES = (BLOCK_N / 2) * BLOCK_K * 4;
    const size_t smem = 1024 + (size_t)STAGES * (A_BYTES + BH_BYTES) + 1024 + kEpiSmemBytes + 2 * kVecKinds * BLOCK_N * sizeof(float);
    static_assert(1024 + STAGES * (A_BYTES + BH_BYTES) + 1024 + kEpiSmemBytes + 2 * kVecKinds * BLOCK_N * 4 <= 232448,
                  "CTA-pair conv kernel: shared memory budget");
    auto kern = conv_igemm_tf32_2cta_kernel<BLOCK_N, STAGES, OP16>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("conv: cudaFuncSetAttribute(2cta): %s", cudaGetErrorString(e)); return (int)e; }
        configured = true;
    }
    int clusters = p.total_pairs < kNumSMs / 2 ? p.total_pairs : kNumSMs / 2;
    kern<<<2 * clusters, kConvThreads, smem, st>>>(ta, tb, om, p);    // __cluster_dims__(2,1,1): CTA pairs
    return SR_OK;
}

template <int BLOCK_N, int SA, int SB, int TPS, int EPI_WARPS, bool OP16>
int launch_conv_halo(const HaloMaps &am, const CUtensorMap &tb, const ConvOutMaps &om, const ConvKParams &p, const HaloParams &hp,
                     cudaStream_t st)
{
    constexpr int BH_BYTES = (BLOCK_N / 2) * BLOCK_K * 4;
    const size_t smem = 1024 + (size_t)SA * kHaloStageBytes + (size_t)SB * TPS * BH_BYTES + 1024 +
                        (size_t)EPI_WARPS * 2 * kStageBufBytes + 2 * kVecKinds * BLOCK_N * sizeof(float);
    static_assert(1024 + SA * kHaloStageBytes + SB * TPS * BH_BYTES + 1024 + EPI_WARPS * 2 * kStageBufBytes +
                  2 * kVecKinds * BLOCK_N * 4 <= 232448, "halo kernel: shared memory budget");
    auto kern = conv_halo_tf32_2cta_kernel<BLOCK_N, SA, SB, TPS, EPI_WARPS, OP16>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("conv: cudaFuncSetAttribute(halo): %s", cudaGetErrorString(e)); return (int)e; }
        configured = true;
    }
    int clusters = p.total_pairs < kNumSMs / 2 ? p.total_pairs : kNumSMs / 2;
    kern<<<2 * clusters, 64 + 32 * EPI_WARPS, smem, st>>>(am, tb, om, p, hp);
    return SR_OK;
}

// Group the taps of every phase by parity class modulo the input stride and describe each class as a halo box
// (see conv_halo_tf32_2cta_kernel).  Returns false when the problem does not fit the halo kernel.
bool build_halo(const sr_conv_args *args, int count, EncodeTiledFn enc, HaloParams &hp, HaloMaps &am, int op16)
{
    const int esize = op16 ? 2 : 4, kelems = 128 / esize;
    const CUtensorMapDataType dtype = op16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    const sr_conv_args *a = args;
    const int s = a->in_stride;
    // Default: one (8 + rx)-wide box per parity class; operand groups then start on arbitrary 128-byte rows, which costs
    // nothing (scratch/umma_align_bench.cu: 64.0 clk per M128 x N128 x K8 for every start row / SBO).  SR_HALO_SPLITX=1
    // loads one 8-wide box per (class, dx) instead, so every group starts on a 1024-byte atom (experiment; slower).
    static const char *splitx_env = getenv("SR_HALO_SPLITX");
    const bool splitx = splitx_env && splitx_env[0] == '1';
    struct MapKey { int cy, cx, hw, hh; } keys[kMaxGroups];
    int num_maps = 0;
    for (int i = 0; i < kMaxPhases; ++i) {
        HaloPhase &h = hp.ph[i];
        const sr_conv_args *b = args + (i < count ? i : 0);
        h.num_groups = 0;
        int cls_y[kMaxGroups], cls_x[kMaxGroups], cls_q[kMaxGroups], minx[kMaxGroups], maxx[kMaxGroups], miny[kMaxGroups],
            maxy[kMaxGroups];
        int tap_group[9];
        for (int t = 0; t < b->num_taps; ++t) {
            const int cy = pos_mod_i(b->tap_dy[t], s), cx = pos_mod_i(b->tap_dx[t], s);
            const int qy = floor_div_i(b->tap_dy[t], s), qx = floor_div_i(b->tap_dx[t], s);
            int g = 0;
            for (; g < h.num_groups; ++g) if (cls_y[g] == cy && cls_x[g] == cx && (!splitx || cls_q[g] == qx)) break;
            if (g == h.num_groups) {
                if (g == kMaxGroups) return false;
                cls_y[g] = cy; cls_x[g] = cx; cls_q[g] = qx; minx[g] = maxx[g] = qx; miny[g] = maxy[g] = qy;
                ++h.num_groups;
            }
            if (qx < minx[g]) minx[g] = qx;
            if (qx > maxx[g]) maxx[g] = qx;
            if (qy < miny[g]) miny[g] = qy;
            if (qy > maxy[g]) maxy[g] = qy;
            tap_group[t] = g;
        }
        int pos = 0;
        for (int g = 0; g < h.num_groups; ++g) {
            const int hw = 8 + maxx[g] - minx[g], hh = 16 + maxy[g] - miny[g];
            if (hw * hh * 128 > kHaloStageBytes || hw > 256 || hh > 256) return false;
            int m = 0;
            for (; m < num_maps; ++m)
                if (keys[m].cy == cls_y[g] && keys[m].cx == cls_x[g] && keys[m].hw == hw && keys[m].hh == hh) break;
            if (m == num_maps) {
                if (m == kMaxGroups) return false;
                const int64_t lw = (a->in_w - cls_x[g] + s - 1) / s, lh = (a->in_h - cls_y[g] + s - 1) / s;
                if (lw < 1 || lh < 1) return false;
                const char *base = reinterpret_cast<const char *>(a->in) + ((int64_t)cls_y[g] * a->in_w + cls_x[g]) * a->cin * esize;
                cuuint64_t dims[4] = {(cuuint64_t)a->cin, (cuuint64_t)lw, (cuuint64_t)lh, (cuuint64_t)a->batch};
                cuuint64_t strides[3] = {(cuuint64_t)a->cin * esize * s, (cuuint64_t)a->in_w * a->cin * esize * s,
                                         (cuuint64_t)a->in_h * a->in_w * a->cin * esize};
                cuuint32_t box[4] = {(cuuint32_t)kelems, (cuuint32_t)hw, (cuuint32_t)hh, 1};
                cuuint32_t estr[4] = {1, 1, 1, 1};
                if (enc(&am.a[m], dtype, 4, const_cast<char *>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return false;
                keys[m] = {cls_y[g], cls_x[g], hw, hh};
                ++num_maps;
            }
            h.g_map[g] = m; h.g_sbo[g] = hw * 128; h.g_bytes[g] = hw * hh * 128;
            h.g_xoff[g] = minx[g]; h.g_yoff[g] = miny[g];
            h.g_tap_begin[g] = pos;
            for (int t = 0; t < b->num_taps; ++t) {
                if (tap_group[t] != g) continue;
                h.tap_k0[pos] = (int)(b->tap_w[t] * a->cin);
                h.tap_aoff[pos] = ((floor_div_i(b->tap_dy[t], s) - miny[g]) * hw + (floor_div_i(b->tap_dx[t], s) - minx[g])) * 128;
                ++pos;
            }
            h.g_tap_end[g] = pos;
        }
        for (int g = h.num_groups; g < kMaxGroups; ++g) {
            h.g_map[g] = 0; h.g_sbo[g] = 1024; h.g_bytes[g] = 0; h.g_xoff[g] = h.g_yoff[g] = 0; h.g_tap_begin[g] = h.g_tap_end[g] = 0;
        }
        for (int t = pos; t < 9; ++t) { h.tap_k0[t] = 0; h.tap_aoff[t] = 0; }
    }
    for (int m = num_maps; m < kMaxGroups; ++m) am.a[m] = am.a[0];
    return num_maps >= 1;
}

}  // namespace
}  // namespace sr

using namespace sr;

// `count` launches that share tensors / strides / epilogue and differ only in tap list, lattice size and lattice
// origin (the four parity classes of the stride-2 transposed conv) run as ONE persistent grid.
static int conv_igemm_multi(const sr_conv_args *args, int count, void *stream, int op16)
{
    SR_REQUIRE(args && count >= 1 && count <= kMaxPhases, "conv: 1..4 phases");
    const sr_conv_args *a = args;
    const int esize = op16 ? 2 : 4, kelems = 128 / esize;          // one K block = one 128-byte swizzle row of channels
    const CUtensorMapDataType op_dtype = op16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    SR_REQUIRE(a->in && a->weight && a->out, "conv: null tensor");
    SR_REQUIRE(a->cin >= kelems && a->cin % kelems == 0, "conv: cin must be a multiple of %d (got %lld)", kelems, (long long)a->cin);
    SR_REQUIRE(a->cout >= 128 && a->cout % 128 == 0, "conv: cout must be a multiple of 128 (got %lld)", (long long)a->cout);
    SR_REQUIRE(a->in_stride >= 1 && a->in_stride <= 8 && a->out_stride >= 1, "conv: bad strides");
    SR_REQUIRE(a->batch >= 1, "conv: empty problem");
    SR_REQUIRE((reinterpret_cast<uintptr_t>(a->in) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->weight) & 15) == 0 &&
               (reinterpret_cast<uintptr_t>(a->out) & 15) == 0, "conv: tensors must be 16-byte aligned");
    SR_REQUIRE(a->epilogue >= 0 && a->epilogue <= 2, "conv: unknown epilogue");
    SR_REQUIRE(!a->out2 || a->scale2, "conv: out2 needs scale2");
    SR_REQUIRE(!a->noise || a->noise_weight, "conv: noise needs noise_weight");
    int64_t max_gw = 0, max_gh = 0;
    for (int i = 0; i < count; ++i) {
        const sr_conv_args *b = args + i;
        SR_REQUIRE(b->num_taps >= 1 && b->num_taps <= 9, "conv: 1..9 taps");
        SR_REQUIRE(b->grid_h >= 1 && b->grid_w >= 1, "conv: empty lattice");
        SR_REQUIRE(b->in == a->in && b->weight == a->weight && b->out == a->out && b->out2 == a->out2 && b->cin == a->cin &&
                   b->cout == a->cout && b->batch == a->batch && b->in_h == a->in_h && b->in_w == a->in_w &&
                   b->out_h == a->out_h && b->out_w == a->out_w && b->in_stride == a->in_stride &&
                   b->out_stride == a->out_stride && b->epilogue == a->epilogue && b->taps_total == a->taps_total,
                   "conv: phases must share tensors, strides and epilogue");
        if (b->grid_w > max_gw) max_gw = b->grid_w;
        if (b->grid_h > max_gh) max_gh = b->grid_h;
    }
    EncodeTiledFn enc = encode_tiled();
    if (!enc) { set_error("conv: cuTensorMapEncodeTiled not available from the driver"); return SR_ERR_DRIVER; }
    cudaStream_t st = (cudaStream_t)stream;

    // Halo kernel (8 x 16 pixel tiles of one image, one activation box per K block instead of one per tap) whenever
    // the lattice is large enough for CTA pairs; SR_CONV_HALO=0 keeps the im2col kernels (A/B comparison).
    static const char *force_halo = getenv("SR_CONV_HALO");
    HaloParams hp;
    HaloMaps am;
    // Only where one box serves many taps (the plain 3x3 conv and its dgrad): for the 4/2/2/1-tap phases of the transposed
    // conv and for the strided gather the per-tap boxes of the im2col kernel were measured as fast or faster.
    bool use_halo = max_gw >= 8 && max_gh >= 16 && a->in_stride == 1 && count == 1 && a->num_taps >= 6 &&
                    !(force_halo && force_halo[0] == '0');
    if (force_halo && force_halo[0] == '2') use_halo = max_gw >= 8 && max_gh >= 16;      // experiment: halo everywhere
    if (use_halo) {
        long long mt2 = 0;
        for (int i = 0; i < count; ++i) {
            const long long t_ph = ((args[i].grid_w + 7) / 8) * ((args[i].grid_h + 15) / 16) * a->batch;
            mt2 += (t_ph + 1) / 2 * 2;
        }
        use_halo = mt2 / 2 * (a->cout / 128) >= 32 && build_halo(args, count, enc, hp, am, op16);
    }

    // tile shape: 16x8 pixels of one image, or several whole small images
    int tw, th, tn;
    if (use_halo) { tw = 8; th = 16; tn = 1; }
    else if (max_gw > 8) { tw = 16; th = 8; tn = 1; }
    else if (max_gw > 4) { tw = 8; th = (max_gh > 4) ? 8 : 4; tn = 128 / (tw * th); }
    else { tw = 4; th = 4; tn = 8; }

    ConvKParams p;
    p.tw_log2 = ilog2_exact(tw); p.th_log2 = ilog2_exact(th); p.tn_log2 = ilog2_exact(tn);
    p.batch = (int)a->batch;
    p.op16 = op16; p.kelems = kelems;
    p.kblocks_per_tap = (int)(a->cin / kelems);
    p.num_phases = count;
    const int tiles_n = (int)((a->batch + tn - 1) / tn);
    // multi-phase launches whose input does not fit in L2: one image per group (measured 0.66 vs 0.70 ms at 256->128
    // @128^2, 537 MB input; no gain or a small loss below that); SR_CONV_PHASE_MAJOR=1/0 forces phase- / group-major order
    static const char *pm_env = getenv("SR_CONV_PHASE_MAJOR");
    const double in_bytes = 4.0 * (double)a->batch * a->in_h * a->in_w * a->cin;
    bool group_major = count > 1 && in_bytes > 384e6;
    if (pm_env) group_major = count > 1 && pm_env[0] == '0';
    const int group_n = group_major ? 1 : tiles_n;
    long long m_tiles = 0, m_tiles2 = 0;
    for (int i = 0; i < kMaxPhases; ++i) {
        ConvPhase &ph = p.ph[i];
        const sr_conv_args *b = args + (i < count ? i : 0);
        ph.num_taps = b->num_taps;
        for (int t = 0; t < 9; ++t) { ph.tap_dy[t] = b->tap_dy[t]; ph.tap_dx[t] = b->tap_dx[t]; ph.tap_k0[t] = (int)(b->tap_w[t] * a->cin); }
        ph.grid_h = (int)b->grid_h; ph.grid_w = (int)b->grid_w;
        ph.tiles_x = (int)((b->grid_w + tw - 1) / tw);
        ph.tiles_y = (int)((b->grid_h + th - 1) / th);
        ph.tile_begin = (int)m_tiles;
        ph.tile_begin2 = (int)m_tiles2;
        ph.out_offset = ((long long)b->out_y0 * a->out_w + b->out_x0) * a->cout;
        ph.noise_offset = (long long)b->out_y0 * a->out_w + b->out_x0;
        if (i < count) {
            const long long t_ph = (long long)ph.tiles_x * ph.tiles_y * group_n;
            m_tiles += t_ph;
            m_tiles2 += (t_ph + 1) / 2 * 2;
        }
    }
    p.group_n = group_n; p.tiles_n = tiles_n;
    p.tiles_per_group = (int)m_tiles; p.tiles_per_group2 = (int)m_tiles2;
    const long long num_groups = (tiles_n + group_n - 1) / group_n;
    m_tiles *= num_groups; m_tiles2 *= num_groups;          // totals over the batch
    // few tiles (low resolutions): prefer 128-wide N tiles so more SMs take part
    int block_n = (a->cout % 256 == 0) ? 256 : 128;
    if (block_n == 256 && (use_halo ? m_tiles2 / 2 : m_tiles) * (a->cout / 256) < kNumSMs / 2) block_n = 128;
    p.n_tiles = (int)(a->cout / block_n);
    SR_REQUIRE(m_tiles * p.n_tiles < 0x7fffffffll, "conv: too many tiles");
    p.total_tiles = (int)(m_tiles * p.n_tiles);
    p.total_pairs = (int)(m_tiles2 / 2 * p.n_tiles);
    // CTA pairs (cta_group::2) halve the weight-tile traffic per SM; they need enough tiles to fill the 74 pairs.
    // SR_CONV_2CTA=0/1 forces the choice.
    static const char *force_2cta = getenv("SR_CONV_2CTA");
    bool use_2cta = p.total_pairs >= kNumSMs / 2;      // measured: 921 vs 839 TF/s (512 ch @64^2), 880 vs 783 (256 ch @128^2)
    if (force_2cta) use_2cta = force_2cta[0] == '1';
    if (use_halo) use_2cta = true;

    CUtensorMap ta, tb;
    {
        cuuint64_t dims[4] = {(cuuint64_t)a->cin, (cuuint64_t)a->in_w, (cuuint64_t)a->in_h, (cuuint64_t)a->batch};
        cuuint64_t strides[3] = {(cuuint64_t)a->cin * esize, (cuuint64_t)a->in_w * a->cin * esize,
                                 (cuuint64_t)a->in_h * a->in_w * a->cin * esize};
        const cuuint32_t s = (cuuint32_t)a->in_stride;
        cuuint32_t box[4] = {(cuuint32_t)kelems, (cuuint32_t)tw * s, (cuuint32_t)th * s, (cuuint32_t)tn};
        cuuint32_t estr[4] = {1, s, s, 1};
        CUresult r = enc(&ta, op_dtype, 4, const_cast<float *>(a->in), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("conv: cuTensorMapEncodeTiled(A) failed with %d", (int)r); return SR_ERR_DRIVER; }
    }
    {
        const cuuint64_t ktot = (cuuint64_t)a->taps_total * a->cin;
        cuuint64_t dims[2] = {ktot, (cuuint64_t)a->cout};
        cuuint64_t strides[1] = {ktot * esize};
        cuuint32_t box[2] = {(cuuint32_t)kelems, (cuuint32_t)(use_2cta ? block_n / 2 : block_n)};   // a CTA of a pair stages half the rows
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&tb, op_dtype, 2, const_cast<float *>(a->weight), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("conv: cuTensorMapEncodeTiled(B) failed with %d", (int)r); return SR_ERR_DRIVER; }
    }

    // output lattices as strided 4-D views {channels, grid_w, grid_h, batch}; box = one epilogue warp's 32 rows x 32 channels
    ConvOutMaps om;
    {
        int bw = tw, bh = 32 / tw, bn = 1;
        if (bh > th) { bn = bh / th; bh = th; }
        for (int i = 0; i < kMaxPhases; ++i) {
            const sr_conv_args *b = args + (i < count ? i : 0);
            for (int which = 0; which < 2; ++which) {
                char *base = reinterpret_cast<char *>(which == 0 ? a->out : a->out2);
                CUtensorMap *m = which == 0 ? &om.out[i] : &om.out2[i];
                if (!base) { om.out2[i] = om.out[i]; continue; }
                // `out` is always fp32; the second output (the next layer's operand) has the operand type: a warp's
                // 32 x 32-channel block is 128-byte rows (fp32, 128-byte swizzle) or 64-byte rows (bf16, 64-byte swizzle)
                const bool o16 = which == 1 && op16;
                const long long es = o16 ? 2 : 4;
                base += ((long long)b->out_y0 * a->out_w + b->out_x0) * a->cout * es;
                cuuint64_t dims[4] = {(cuuint64_t)a->cout, (cuuint64_t)b->grid_w, (cuuint64_t)b->grid_h, (cuuint64_t)a->batch};
                cuuint64_t strides[3] = {(cuuint64_t)(a->out_stride * a->cout * es), (cuuint64_t)(a->out_stride * a->out_w * a->cout * es),
                                         (cuuint64_t)(a->out_h * a->out_w * a->cout * es)};
                cuuint32_t box[4] = {32, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
                cuuint32_t estr[4] = {1, 1, 1, 1};
                CUresult r = enc(m, o16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, dims, strides,
                                 box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, o16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) { set_error("conv: cuTensorMapEncodeTiled(out) failed with %d", (int)r); return SR_ERR_DRIVER; }
            }
        }
    }
    SR_REQUIRE(!a->out2 || (reinterpret_cast<uintptr_t>(a->out2) & 15) == 0, "conv: out2 must be 16-byte aligned");

    p.in_stride = a->in_stride;
    p.cout = (int)a->cout;
    p.out_pix_stride = (long long)a->out_stride * a->cout;
    p.out_row_stride = (long long)a->out_stride * a->out_w * a->cout;
    p.out_img_stride = (long long)a->out_h * a->out_w * a->cout;
    p.out = a->out; p.out2 = a->out2;
    p.epilogue = a->epilogue;
    p.rowscale = a->rowscale; p.scale2 = a->scale2; p.bias = a->bias;
    p.noise = a->noise; p.noise_weight = a->noise_weight; p.stylemap = a->stylemap;
    // noise / stylemap are planar [*, out_h, out_w] and are addressed on the same output lattice
    p.noise_pix_stride = a->out_stride;
    p.noise_row_stride = (long long)a->out_stride * a->out_w;
    p.noise_img_stride = a->noise_batch_stride;
    p.map_plane_stride = (long long)a->out_h * a->out_w;
    p.map_img_stride = a->stylemap_batch_stride;
    p.alpha = a->alpha; p.gain = a->gain;
    p.rgb_w = a->rgb_weight; p.rgb_out = a->rgb_out;
    SR_REQUIRE(!p.rgb_w || p.rgb_out, "conv: rgb_weight needs rgb_out");
    static const char *debug_env = getenv("SR_CONV_DEBUG");
    p.debug = debug_env ? atoi(debug_env) : 0;
    if (p.rgb_w) {
        cudaError_t e = cudaMemsetAsync(p.rgb_out, 0, sizeof(float) * 3 * (size_t)(a->batch * a->out_h * a->out_w), st);
        if (e != cudaSuccess) { set_error("conv: memset(rgb_out): %s", cudaGetErrorString(e)); return (int)e; }
    }

    int rc;
    if (use_halo) {
        // 8 epilogue warps pay when a 128-channel tile also writes the pre-modulated second output (measured 0.93 ->
        // 0.84 ms at 128 ch / 256^2); elsewhere 4 warps and one more pipeline stage are as fast or faster.
        static const char *epi_env = getenv("SR_CONV_EPI_WARPS");          // A/B: force 4 or 8 epilogue warps
        bool epi8 = a->out2 != nullptr && block_n == 128;
        if (epi_env) epi8 = epi_env[0] == '8';
#define SR_HALO(N, A, B, T, E) (p.op16 ? launch_conv_halo<N, A, B, T, E, true>(am, tb, om, p, hp, st) \
                                       : launch_conv_halo<N, A, B, T, E, false>(am, tb, om, p, hp, st))
        if (block_n == 256) rc = epi8 ? SR_HALO(256, 2, 6, 1, 8) : SR_HALO(256, 2, 7, 1, 4);
        else rc = epi8 ? SR_HALO(128, 2, 4, 3, 8) : SR_HALO(128, 3, 4, 3, 4);
#undef SR_HALO
    } else if (use_2cta) {
        if (block_n == 256) rc = p.op16 ? launch_conv_2cta<256, 5, true>(ta, tb, om, p, st) : launch_conv_2cta<256, 5, false>(ta, tb, om, p, st);
        else rc = p.op16 ? launch_conv_2cta<128, 7, true>(ta, tb, om, p, st) : launch_conv_2cta<128, 7, false>(ta, tb, om, p, st);
    } else if (block_n == 256) rc = p.op16 ? launch_conv<256, 4, true>(ta, tb, om, p, st) : launch_conv<256, 4, false>(ta, tb, om, p, st);
    else rc = p.op16 ? launch_conv<128, 6, true>(ta, tb, om, p, st) : launch_conv<128, 6, false>(ta, tb, om, p, st);
    if (rc != SR_OK) return rc;
    count_launch();
    return check_launch("sr_conv_igemm_tf32");
}

extern "C" int sr_conv_igemm_multi_tf32(const sr_conv_args *args, int count, void *stream) { return conv_igemm_multi(args, count, stream, 0); }
extern "C" int sr_conv_igemm_tf32(const sr_conv_args *a, void *stream) { return conv_igemm_multi(a, 1, stream, 0); }
// bf16 operand form: `in`, `weight` and `out2` point to bfloat16 data of the same logical shapes; `out`, the epilogue
// vectors and the accumulation stay fp32.  cin % 64 == 0.  (BASELINE.json configs[3]: the bf16 train step.)
extern "C" int sr_conv_igemm_multi_bf16(const sr_conv_args *args, int count, void *stream) { return conv_igemm_multi(args, count, stream, 1); }

static int conv_wgrad(const sr_wgrad_args *a, void *stream, int op16)
{
    SR_REQUIRE(a && a->g && a->x && a->dw, "wgrad: null argument");
    const int esize = op16 ? 2 : 4;
    SR_REQUIRE(a->cout >= 128 && a->cout % 128 == 0 && a->cin >= 128 && a->cin % 128 == 0,
               "wgrad: cin and cout must be multiples of 128 (got %lld, %lld)", (long long)a->cin, (long long)a->cout);
    SR_REQUIRE(a->num_taps >= 1 && a->num_taps <= 9 && a->g_stride >= 1 && a->x_stride >= 1, "wgrad: bad taps/strides");
    SR_REQUIRE(a->batch >= 1 && a->grid_h >= 1 && a->grid_w >= 1, "wgrad: empty problem");
    EncodeTiledFn enc = encode_tiled();
    if (!enc) { set_error("wgrad: cuTensorMapEncodeTiled not available from the driver"); return SR_ERR_DRIVER; }
    cudaStream_t st = (cudaStream_t)stream;
    static const bool allow_wide = getenv("SR_WGRAD_WIDE") != nullptr;      // experimental N = 256 shape (measured slower in round 1)
    static const char *force_pair = getenv("SR_WGRAD_2CTA");
    bool pair = (a->cout % 256 == 0) && (a->cin % 256 == 0);               // CTA-pair kernel: 256 x 256 block per tap
    if (force_pair) pair = pair && force_pair[0] == '1';
    const bool wide = !pair && allow_wide && (a->cin % 256 == 0);           // N = 256, 32 pixels per stage
    // Two-tap CTA pairs for 128-wide channel blocks (wgrad_tf32_2cta_taps_kernel): every tap must read X at the same
    // offset.  True as given for the transposed conv (taps shift G); for a stride-1 conv whose taps shift X over an X of
    // the same size as G, the shift moves to G:  sum_p g[p] x[p + t] = sum_p' g[p' - t] x[p']  (zero fill outside G).
    static const char *taps2_env = getenv("SR_WGRAD_TAPS2");                // A/B: 0 = single-CTA N = 128 kernel
    int32_t tg_dy[9], tg_dx[9], tx_dy[9], tx_dx[9];
    bool taps2 = !pair && !wide && a->num_taps >= 2 && !(taps2_env && taps2_env[0] == '0');
    if (taps2) {
        bool x_same = true, g_zero = true;
        for (int t = 0; t < a->num_taps; ++t) {
            x_same &= a->x_dy[t] == a->x_dy[0] && a->x_dx[t] == a->x_dx[0];
            g_zero &= a->g_dy[t] == 0 && a->g_dx[t] == 0;
        }
        for (int t = 0; t < 9; ++t) { tg_dy[t] = a->g_dy[t]; tg_dx[t] = a->g_dx[t]; tx_dy[t] = a->x_dy[t]; tx_dx[t] = a->x_dx[t]; }
        if (!x_same) {
            const bool movable = g_zero && a->g_stride == 1 && a->x_stride == 1 && a->g_h == a->x_h && a->g_w == a->x_w &&
                                 a->grid_h == a->x_h && a->grid_w == a->x_w;
            if (movable) {
                for (int t = 0; t < a->num_taps; ++t) { tg_dy[t] = -a->x_dy[t]; tg_dx[t] = -a->x_dx[t]; tx_dy[t] = 0; tx_dx[t] = 0; }
            } else {
                taps2 = false;
            }
        }
    }
    int tw, th, tn;
    if (wide) {
        if (a->grid_w > 8) { tw = 16; th = 2; tn = 1; }
        else if (a->grid_w > 4) { tw = 8; th = 4; tn = 1; }
        else { tw = 4; th = 4; tn = 2; }
    } else {
        if (a->grid_w > 8) { tw = 16; th = 4; tn = 1; }
        else if (a->grid_w > 4) { tw = 8; th = 8; tn = 1; }
        else { tw = 4; th = 4; tn = 4; }
    }
    const int wg_n = (wide || pair) ? 256 : 128;
    CUtensorMap tg, tx;
    auto make_map = [&](CUtensorMap *m, const float *base, int64_t hh, int64_t ww, int64_t cc, int stride) -> int {
        cuuint64_t dims[4] = {(cuuint64_t)cc, (cuuint64_t)ww, (cuuint64_t)hh, (cuuint64_t)a->batch};
        cuuint64_t strides[3] = {(cuuint64_t)cc * esize, (cuuint64_t)ww * cc * esize, (cuuint64_t)hh * ww * cc * esize};
        const cuuint32_t s = (cuuint32_t)stride;
        cuuint32_t box[4] = {(cuuint32_t)(128 / esize), (cuuint32_t)tw * s, (cuuint32_t)th * s, (cuuint32_t)tn};
        cuuint32_t estr[4] = {1, s, s, 1};
        // MN-major operands: tf32 exists only with 32-byte swizzle atomicity, 16-bit types take the plain 128-byte swizzle
        CUresult r = enc(m, op16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                         const_cast<float *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         op16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("wgrad: cuTensorMapEncodeTiled failed with %d", (int)r); return SR_ERR_DRIVER; }
        return SR_OK;
    };
    int rc = make_map(&tg, a->g, a->g_h, a->g_w, a->cout, a->g_stride);
    if (rc != SR_OK) return rc;
    rc = make_map(&tx, a->x, a->x_h, a->x_w, a->cin, a->x_stride);
    if (rc != SR_OK) return rc;

    WgradKParams p;
    p.tw_log2 = ilog2_exact(tw); p.th_log2 = ilog2_exact(th); p.tn_log2 = ilog2_exact(tn);
    p.tiles_x = (int)((a->grid_w + tw - 1) / tw);
    p.tiles_y = (int)((a->grid_h + th - 1) / th);
    p.tiles_n = (int)((a->batch + tn - 1) / tn);
    p.n_tiles = (int)(a->cin / wg_n);
    const int mn_tiles = (int)(a->cout / (pair ? 256 : BLOCK_M)) * p.n_tiles;      // pair kernel: 256 x 256 blocks
    const long long total_kt = (long long)p.tiles_x * p.tiles_y * p.tiles_n;
    // Split K.  All CTAs of a launch run the same number of K tiles and one CTA (pair) owns an SM (pair), so the
    // launch executes in whole waves: time ~ ceil(CTAs / slots) * (K tiles per CTA + fixed cost), where the fixed cost
    // (TMEM alloc, pipeline fill, red.add epilogue) is worth about 16 K tiles.  Pick the split count that minimises it
    // (e.g. 9 taps x 33 splits = 297 CTAs is THREE waves of 148; 32 splits is two).
    const long long slots = (pair || taps2) ? kNumSMs / 2 : kNumSMs;
    const long long base_ctas = (long long)mn_tiles * (taps2 ? (a->num_taps + 1) / 2 : a->num_taps);
    long long splits = 1, best_cost = -1;
    const long long max_splits = total_kt < 4 * slots ? total_kt : 4 * slots;
    for (long long s = 1; s <= max_splits; ++s) {
        const long long per = (total_kt + s - 1) / s;
        const long long real = (total_kt + per - 1) / per;
        if (real != s) continue;                                  // same work distribution as a smaller s
        const long long waves = (base_ctas * s + slots - 1) / slots;
        const long long cost = waves * (per + 16);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; splits = s; }
        if (base_ctas * s > 6 * slots) break;
    }
    p.ktiles_per_split = (int)((total_kt + splits - 1) / splits);
    splits = (total_kt + p.ktiles_per_split - 1) / p.ktiles_per_split;
    p.g_stride = a->g_stride; p.x_stride = a->x_stride;
    for (int t = 0; t < 9; ++t) {
        p.g_dy[t] = a->g_dy[t]; p.g_dx[t] = a->g_dx[t]; p.x_dy[t] = a->x_dy[t]; p.x_dx[t] = a->x_dx[t];
        if (taps2) { p.g_dy[t] = tg_dy[t]; p.g_dx[t] = tg_dx[t]; p.x_dy[t] = tx_dy[t]; p.x_dx[t] = tx_dx[t]; }
        p.tap_out[t] = a->tap_out[t];
    }
    p.taps_total = (int)a->taps_total; p.cin = (int)a->cin;
    p.dw = a->dw;
    p.op16 = op16;
    if (a->zero_init) {
        cudaError_t e = cudaMemsetAsync(a->dw, 0, sizeof(float) * (size_t)a->cout * a->taps_total * a->cin, st);
        if (e != cudaSuccess) { set_error("wgrad: memset: %s", cudaGetErrorString(e)); return (int)e; }
    }
    const dim3 grid((unsigned)(pair ? 2 * mn_tiles : mn_tiles), (unsigned)a->num_taps, (unsigned)splits);
    static bool configured[2] = {false, false};
    static bool configured_pair = false, configured_taps2 = false;
    if (taps2) {
        const size_t smem = 1024 + (size_t)4 * 6 * 64 * 128 + 256;
        if (!configured_taps2) {
            cudaError_t e = cudaFuncSetAttribute(wgrad_tf32_2cta_taps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) { set_error("wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
            configured_taps2 = true;
        }
        const dim3 grid2((unsigned)(2 * mn_tiles), (unsigned)((a->num_taps + 1) / 2), (unsigned)splits);
        wgrad_tf32_2cta_taps_kernel<<<grid2, kConvThreads, smem, st>>>(tg, tx, p, (int)a->num_taps);
    } else if (pair) {
        const size_t smem = 1024 + (size_t)3 * 8 * 64 * 128 + 256;
        if (!configured_pair) {
            cudaError_t e = cudaFuncSetAttribute(wgrad_tf32_2cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) { set_error("wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
            configured_pair = true;
        }
        wgrad_tf32_2cta_kernel<<<grid, kConvThreads, smem, st>>>(tg, tx, p);
    } else if (wide) {
        auto kern = wgrad_tf32_kernel<256, 32, 4>;
        const size_t smem = 1024 + (size_t)4 * (4 + 8) * 32 * 128 + 256;
        if (!configured[0]) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) { set_error("wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
            configured[0] = true;
        }
        kern<<<grid, kConvThreads, smem, st>>>(tg, tx, p);
    } else {
        auto kern = wgrad_tf32_kernel<128, 64, 3>;
        const size_t smem = 1024 + (size_t)3 * (4 + 4) * 64 * 128 + 256;
        if (!configured[1]) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) { set_error("wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
            configured[1] = true;
        }
        kern<<<grid, kConvThreads, smem, st>>>(tg, tx, p);
    }
    count_launch();
    return check_launch("sr_conv_wgrad_tf32");
}

extern "C" int sr_conv_wgrad_tf32(const sr_wgrad_args *a, void *stream) { return conv_wgrad(a, stream, 0); }
// bf16 operand form: g and x point to bfloat16 NHWC tensors; dw stays fp32 (split-K reduction with fp32 atomics).
extern "C" int sr_conv_wgrad_bf16(const sr_wgrad_args *a, void *stream) { return conv_wgrad(a, stream, 1); }

static int modulate_any(float *xs, const float *x, const float *style, int64_t batch, int64_t pixels, int64_t channels,
                        void *stream, bool bf16)
{
    SR_REQUIRE(xs && x, "modulate: null pointer");
    SR_REQUIRE(channels % 4 == 0, "modulate: channels must be a multiple of 4");
    SR_REQUIRE((reinterpret_cast<uintptr_t>(xs) & 15) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
               (!style || (reinterpret_cast<uintptr_t>(style) & 15) == 0), "modulate: 16-byte alignment required");
    const int64_t n4 = batch * pixels * channels / 4;
    if (n4 == 0) return SR_OK;
    SR_REQUIRE(n4 < 0x7fffffffll, "modulate: tensor too large");
    int64_t blocks = (n4 + 255) / 256;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    if (bf16)
        modulate_tf32_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
            xs, x, style, (uint32_t)n4, FastDiv((uint32_t)(channels / 4)), FastDiv((uint32_t)(pixels * channels / 4)),
            (int)(channels / 4));
    else
        modulate_tf32_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
            xs, x, style, (uint32_t)n4, FastDiv((uint32_t)(channels / 4)), FastDiv((uint32_t)(pixels * channels / 4)),
            (int)(channels / 4));
    count_launch();
    return check_launch("sr_modulate_tf32");
}

extern "C" int sr_modulate_tf32(float *xs, const float *x, const float *style, int64_t batch, int64_t pixels,
                                int64_t channels, void *stream)
{
    return modulate_any(xs, x, style, batch, pixels, channels, stream, false);
}
// xs is a bfloat16 tensor [batch, pixels, channels] (round to nearest even)
extern "C" int sr_modulate_bf16(void *xs, const float *x, const float *style, int64_t batch, int64_t pixels,
                                int64_t channels, void *stream)
{
    return modulate_any(reinterpret_cast<float *>(xs), x, style, batch, pixels, channels, stream, true);
}

extern "C" int sr_conv_weight_prep_tf32(float *dst, const float *w, float scale, int64_t cout, int64_t cin, int kh, int kw,
                                        int transpose, void *stream)
{
    SR_REQUIRE(dst && w && cout > 0 && cin > 0 && kh > 0 && kw > 0, "weight_prep: bad arguments");
    SR_REQUIRE(transpose >= 0 && transpose <= 3, "weight_prep: transpose mode 0..3");
    const int64_t total = cout * cin * kh * kw;
    int64_t blocks = (total + 255) / 256;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    weight_prep_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(dst, w, scale, (int)cout, (int)cin, kh, kw, transpose);
    count_launch();
    return check_launch("sr_conv_weight_prep_tf32");
}

static int weight_prep_dual_any(float *fwd, float *tr, float *wsq, const float *w, float scale, int64_t cout, int64_t cin,
                                int taps, int flip_transposed, void *stream, int bf16)
{
    SR_REQUIRE(w && cout > 0 && cin > 0 && taps > 0 && taps <= 25, "weight_prep_dual: bad arguments");
    SR_REQUIRE(fwd || tr || wsq, "weight_prep_dual: nothing to produce");
    const dim3 grid((unsigned)((cin + kWP - 1) / kWP), (unsigned)((cout + kWP - 1) / kWP));
    const size_t smem = sizeof(float) * kWP * (kWP * taps + 1);
    weight_prep_dual_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(fwd, tr, wsq, w, scale, (int)cout, (int)cin, taps,
                                                                    flip_transposed, bf16);
    count_launch();
    return check_launch("sr_conv_weight_prep_dual_tf32");
}

extern "C" int sr_conv_weight_prep_dual_tf32(float *fwd, float *tr, float *wsq, const float *w, float scale, int64_t cout,
                                             int64_t cin, int taps, int flip_transposed, void *stream)
{
    return weight_prep_dual_any(fwd, tr, wsq, w, scale, cout, cin, taps, flip_transposed, stream, 0);
}
// fwd / tr are bfloat16 tensors of the same shapes; wsq stays fp32 (computed from the unrounded weights)
extern "C" int sr_conv_weight_prep_dual_bf16(void *fwd, void *tr, float *wsq, const float *w, float scale, int64_t cout,
                                             int64_t cin, int taps, int flip_transposed, void *stream)
{
    return weight_prep_dual_any(reinterpret_cast<float *>(fwd), reinterpret_cast<float *>(tr), wsq, w, scale, cout, cin, taps,
                                flip_transposed, stream, 1);
}

static int weight_prep_multi_any(const sr_weight_prep_item *items, int n, void *stream, int bf16)
{
    SR_REQUIRE(items && n >= 1 && n <= SR_WEIGHT_PREP_MAX, "weight_prep_multi: 1..%d layers", SR_WEIGHT_PREP_MAX);
    WeightPrepTable tab;
    tab.n = n;
    int gx = 0, gy = 0, taps_max = 0;
    for (int i = 0; i < n; ++i) {
        const sr_weight_prep_item &it = items[i];
        SR_REQUIRE(it.w && it.cout > 0 && it.cin > 0 && it.taps > 0 && it.taps <= 25 && (it.fwd || it.tr || it.wsq),
                   "weight_prep_multi: bad layer %d", i);
        tab.item[i] = it;
        gx = std::max(gx, (it.cin + kWP - 1) / kWP); gy = std::max(gy, (it.cout + kWP - 1) / kWP); taps_max = std::max(taps_max, it.taps);
    }
    const size_t smem = sizeof(float) * kWP * (kWP * taps_max + 1);
    static bool configured = false;
    if (!configured && smem > 48 * 1024) {
        cudaFuncSetAttribute(weight_prep_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(float) * kWP * (kWP * 25 + 1)));
        configured = true;
    }
    weight_prep_multi_kernel<<<dim3(gx, gy, n), 256, smem, (cudaStream_t)stream>>>(tab, bf16);
    count_launch();
    return check_launch("sr_conv_weight_prep_multi");
}

extern "C" int sr_conv_weight_prep_multi_tf32(const sr_weight_prep_item *items, int n, void *stream)
{
    return weight_prep_multi_any(items, n, stream, 0);
}
extern "C" int sr_conv_weight_prep_multi_bf16(const sr_weight_prep_item *items, int n, void *stream)
{
    return weight_prep_multi_any(items, n, stream, 1);
}
extern "C" int sr_weight_sq_backward_multi_f32(const sr_weight_prep_item *items, int n, void *stream)
{
    SR_REQUIRE(items && n >= 1 && n <= SR_WEIGHT_PREP_MAX, "weight_sq_backward_multi: 1..%d layers", SR_WEIGHT_PREP_MAX);
    WeightPrepTable tab;
    tab.n = n;
    for (int i = 0; i < n; ++i) {
        const sr_weight_prep_item &it = items[i];
        SR_REQUIRE(it.w && it.gw && it.g_wsq && it.cout > 0 && it.cin > 0 && it.taps > 0 &&
                   (int64_t)it.cout * it.cin * it.taps < (1ll << 31), "weight_sq_backward_multi: bad layer %d", i);
        tab.item[i] = it;
    }
    weight_sq_backward_multi_kernel<<<dim3(kNumSMs * 2, n), 256, 0, (cudaStream_t)stream>>>(tab);
    count_launch();
    return check_launch("sr_weight_sq_backward_multi_f32");
}

static int residual_combine_any(float *out, float *op, const float *a, const float *b, float scale, int64_t n, void *stream, bool bf16)
{
    SR_REQUIRE(out && a && b && n >= 0 && n % 4 == 0, "residual_combine: bad arguments (element count must be a multiple of 4)");
    SR_REQUIRE(((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(op) | reinterpret_cast<uintptr_t>(a) |
                 reinterpret_cast<uintptr_t>(b)) & 15) == 0, "residual_combine: 16-byte alignment required");
    const int64_t n4 = n / 4;
    if (n4 == 0) return SR_OK;
    SR_REQUIRE(n4 < 0x7fffffffll, "residual_combine: tensor too large");
    int64_t blocks = (n4 + 255) / 256;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    if (bf16) residual_combine_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(out, op, a, b, scale, (uint32_t)n4);
    else residual_combine_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(out, op, a, b, scale, (uint32_t)n4);
    count_launch();
    return check_launch("sr_residual_combine");
}

extern "C" int sr_residual_combine_tf32(float *out, float *operand, const float *a, const float *b, float scale, int64_t n,
                                        void *stream)
{
    return residual_combine_any(out, operand, a, b, scale, n, stream, false);
}
extern "C" int sr_residual_combine_bf16(float *out, void *operand, const float *a, const float *b, float scale, int64_t n,
                                        void *stream)
{
    return residual_combine_any(out, reinterpret_cast<float *>(operand), a, b, scale, n, stream, true);
}
