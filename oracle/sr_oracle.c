/* TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the hot path of WestlyPark/StyleRenderer, used as the parity oracle
 * for the sm_100a kernels in stylerenderer_b200/csrc.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this; the product never does.
 *
 * Pinning (see tests/test_oracle_pinning.py): checked against
 *   - the reference's only known-answer test, the 5x5 triangle of op/rasterize.py:83-107,
 *   - the reference's own CPU implementations compiled from /root/reference into oracle/_ref
 *     (rasterize_cpu, rasterize_cpu_backward, fused_bias_act_cpu) -- bit-exact,
 *   - `upfirdn2d_native` (op/upfirdn2d.py:159-200) via committed fixtures in tests/golden/.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared -o oracle/libsr_oracle.so oracle/sr_oracle.c -lm
 * (no -march / -ffast-math: every floating-point operation must round once, like the reference
 * host build)
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

/* ------------------------------------------------------------------ upfirdn2d ------------- */
/* Semantics of op/upfirdn2d.py:159-200 (`upfirdn2d_native`), equivalently the CUDA kernel
 * op/upfirdn2d_kernel.cu:127-204, for a [major, in_h, in_w, minor] tensor:
 *   1. zero-stuff by (up_y, up_x); 2. pad / crop by (pad_y0,pad_y1,pad_x0,pad_x1);
 *   3. TRUE convolution with taps[kh][kw] (taps flipped w.r.t. correlation, :186-187);
 *   4. keep every (down_y, down_x)-th sample; out = (in*up + pad0 + pad1 - k) / down + 1 (:197-198).
 * Written as a gather: out[oy,ox] = sum_{ky,kx} taps[ky][kx] * U[oy*down_y + (kh-1-ky) - pad_y0, ...]
 * where U is the zero-stuffed image and out-of-range / non-multiple-of-up samples are zero. */
void sr_oracle_upfirdn2d_f32(float *out, const float *in, const float *taps,
                             int64_t major, int64_t in_h, int64_t in_w, int64_t minor,
                             int64_t kh, int64_t kw, int64_t up_x, int64_t up_y,
                             int64_t down_x, int64_t down_y,
                             int64_t pad_x0, int64_t pad_x1, int64_t pad_y0, int64_t pad_y1)
{
    int64_t out_h = (in_h * up_y + pad_y0 + pad_y1 - kh) / down_y + 1;
    int64_t out_w = (in_w * up_x + pad_x0 + pad_x1 - kw) / down_x + 1;
    int64_t m, oy, ox, c, ky, kx;
    for (m = 0; m < major; ++m)
        for (oy = 0; oy < out_h; ++oy)
            for (ox = 0; ox < out_w; ++ox)
                for (c = 0; c < minor; ++c) {
                    /* F.conv2d with flipped taps visits the window row-major, (:186-188) */
                    float acc = 0.0f;
                    for (ky = 0; ky < kh; ++ky) {
                        int64_t uy = oy * down_y + ky - pad_y0;       /* row in the zero-stuffed image */
                        if (uy < 0 || uy >= in_h * up_y || uy % up_y) continue;
                        for (kx = 0; kx < kw; ++kx) {
                            int64_t ux = ox * down_x + kx - pad_x0;
                            if (ux < 0 || ux >= in_w * up_x || ux % up_x) continue;
                            acc += in[((m * in_h + uy / up_y) * in_w + ux / up_x) * minor + c]
                                 * taps[(kh - 1 - ky) * kw + (kw - 1 - kx)];
                        }
                    }
                    out[((m * out_h + oy) * out_w + ox) * minor + c] = acc;
                }
}

/* ------------------------------------------------------------------ fused_bias_act -------- */
/* op/fused_bias_act_kernel.cu:43-70 (`fused_bias_act_cpu`) == the CUDA kernel :14-42.
 * act*10+grad: 30 -> leaky-relu of (x+b); 31 -> pass/scale x by the sign of `ref`; 32 -> 0;
 * 1x -> linear.  `bias` / `ref` may be NULL ("empty tensor", op/fused_bias_act.cpp:10-11). */
void sr_oracle_fused_bias_act_f32(float *out, const float *x, const float *bias, const float *ref,
                                  int act, int grad, float alpha, float scale,
                                  int64_t size_x, int64_t step_b, int64_t size_b)
{
    int64_t i;
    for (i = 0; i < size_x; ++i) {
        float v = x[i], r = ref ? ref[i] : 0.0f, y;
        if (bias) v += bias[(i / step_b) % size_b];
        switch (act * 10 + grad) {
        default:
        case 10: case 11: y = v; break;
        case 12: case 32: y = 0.0f; break;
        case 30: y = (v > 0.0f) ? v : v * alpha; break;
        case 31: y = (r > 0.0f) ? v : v * alpha; break;
        }
        out[i] = y * scale;
    }
}

/* grad_bias of op/fused_act.py:33-38: dx summed over every axis but the channel axis.
 * Accumulated in double (order independent). */
void sr_oracle_bias_grad_f32(double *gb, const float *dx, int64_t size_x, int64_t step_b, int64_t size_b)
{
    int64_t i;
    for (i = 0; i < size_b; ++i) gb[i] = 0.0;
    for (i = 0; i < size_x; ++i) gb[(i / step_b) % size_b] += (double)dx[i];
}

/* ------------------------------------------------------------------ rasterizer ------------ */
#define REAL float
#define FN(name) name##_f32
#include "raster_body.inc"
#undef REAL
#undef FN

#define REAL double
#define FN(name) name##_f64
#include "raster_body.inc"
#undef REAL
#undef FN

/* ------------------------------------------------------------------ mesh vertex normals ---- */
/* `mesh_point_normal` (utils_3d.py:379-404) + `Normalize` L2 (layers.py:13-30):
 *   fn[f] = (v_b - v_a) x (v_c - v_a)   (elementwise products and differences, :384-388);
 *   for corner j = 0..2: vn += sparse.mm(I_j, fn)  -- a per-corner partial sum over the faces in list order, added to
 *   the running total (:390-403);  vn / max(sqrt(sum vn^2), eps).
 * The sparse product's internal summation order is not specified; the partial sums are taken here in face order. */
void sr_oracle_vertex_normals_f32(float *out, const float *v, const int64_t *tri, int64_t batch, int64_t nv, int64_t nf,
                                  float *scratch /* [nv*3] */, float eps)
{
    int64_t b, f, i, j;
    for (b = 0; b < batch; ++b) {
        const float *vb = v + b * nv * 3;
        float *o = out + b * nv * 3;
        for (i = 0; i < nv * 3; ++i) o[i] = 0.0f;
        for (j = 0; j < 3; ++j) {
            for (i = 0; i < nv * 3; ++i) scratch[i] = 0.0f;
            for (f = 0; f < nf; ++f) {
                const int64_t ia = tri[f * 3], ib = tri[f * 3 + 1], ic = tri[f * 3 + 2], id = tri[f * 3 + j];
                float ab[3], ac[3], n[3];
                int c;
                for (c = 0; c < 3; ++c) { ab[c] = vb[ib * 3 + c] - vb[ia * 3 + c]; ac[c] = vb[ic * 3 + c] - vb[ia * 3 + c]; }
                n[0] = ab[1] * ac[2] - ab[2] * ac[1];
                n[1] = ab[2] * ac[0] - ab[0] * ac[2];
                n[2] = ab[0] * ac[1] - ab[1] * ac[0];
                for (c = 0; c < 3; ++c) scratch[id * 3 + c] += n[c];
            }
            for (i = 0; i < nv * 3; ++i) o[i] += scratch[i];
        }
        for (i = 0; i < nv; ++i) {
            float x = o[i * 3], y = o[i * 3 + 1], z = o[i * 3 + 2];
            float nrm = sqrtf(x * x + y * y + z * z);
            if (nrm < eps) nrm = eps;
            o[i * 3] = x / nrm; o[i * 3 + 1] = y / nrm; o[i * 3 + 2] = z / nrm;
        }
    }
}
