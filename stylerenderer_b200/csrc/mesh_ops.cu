// Mesh front-end of the rasteriser: area-weighted vertex normals (reference utils_3d.py:379-404 `mesh_point_normal`).
//
// The reference gathers the three corners of every face, takes the un-normalised cross product (b-a) x (c-a) and
// scatter-adds it to the corners through THREE sparse matrix products whose index tensors are built from a Python
// `range(len(tri))` on the host every call (a 70 k element list at BFM size), then L2-normalises with a clamped norm.
// Here: one thread per (image, face) accumulates with float atomics into the zero-filled output, one thread per
// (image, vertex) normalises.  9 atomics per face into an L2-resident 0.4 MB/image array; the summation order is not
// deterministic (a vertex has ~6 incident faces), so parity with the oracle is to rounding, not bit-exact.
#include "common.cuh"

namespace sr {
namespace {

__global__ void __launch_bounds__(256)
face_normal_scatter_kernel(float *__restrict__ vn, const float *__restrict__ v, const int64_t *__restrict__ tri,
                           int64_t batch, int64_t nv, int64_t nf, int shared_f)
{
    const int64_t total = batch * nf;
    for (int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * 256) {
        const int64_t b = idx / nf, f = idx - b * nf;
        const int64_t *t = tri + (shared_f ? f : idx) * 3;
        const int64_t ia = __ldg(t), ib = __ldg(t + 1), ic = __ldg(t + 2);
        if (ia < 0 || ia >= nv || ib < 0 || ib >= nv || ic < 0 || ic >= nv) continue;
        const float *vb = v + b * nv * 3;
        const float ax = __ldg(vb + ia * 3), ay = __ldg(vb + ia * 3 + 1), az = __ldg(vb + ia * 3 + 2);
        const float abx = __ldg(vb + ib * 3) - ax, aby = __ldg(vb + ib * 3 + 1) - ay, abz = __ldg(vb + ib * 3 + 2) - az;
        const float acx = __ldg(vb + ic * 3) - ax, acy = __ldg(vb + ic * 3 + 1) - ay, acz = __ldg(vb + ic * 3 + 2) - az;
        // single-rounded products and differences like the reference's elementwise torch ops (no fused multiply-add)
        const float nx = __fsub_rn(__fmul_rn(aby, acz), __fmul_rn(abz, acy));
        const float ny = __fsub_rn(__fmul_rn(abz, acx), __fmul_rn(abx, acz));
        const float nz = __fsub_rn(__fmul_rn(abx, acy), __fmul_rn(aby, acx));
        float *o = vn + b * nv * 3;
        const int64_t ids[3] = {ia, ib, ic};
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            atomicAdd(o + ids[j] * 3 + 0, nx);
            atomicAdd(o + ids[j] * 3 + 1, ny);
            atomicAdd(o + ids[j] * 3 + 2, nz);
        }
    }
}

__global__ void __launch_bounds__(256)
normalize3_kernel(float *__restrict__ vn, int64_t count, float eps)
{
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < count; i += (int64_t)gridDim.x * 256) {
        float *p = vn + i * 3;
        const float x = p[0], y = p[1], z = p[2];
        // reference layers.py:19-22: norm = sqrt(sum v*v), clamp(min = eps), v / norm
        float n = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
        n = fmaxf(n, eps);
        p[0] = __fdiv_rn(x, n); p[1] = __fdiv_rn(y, n); p[2] = __fdiv_rn(z, n);
    }
}

// out[b,v,:] = in[b,v,:] . R[b] + t[b]   with pose[b] = the 3x4 matrix [ R | t ] row-major (R applied from the right, as in
// reference utils_3d.py:374-376: matmul(v, T[:, :3, :3]) + T[:, :3, 3]).  One thread per vertex; the 12 pose floats of an
// image are a broadcast load.
__global__ void __launch_bounds__(256)
pose_apply_kernel(float *__restrict__ out, const float *__restrict__ in, const float *__restrict__ pose, int64_t batch,
                  int64_t nv, int in_stride)
{
    const int64_t total = batch * nv;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int64_t b = i / nv;
        const float *T = pose + b * 12;
        const float *p = in + i * in_stride;
        const float x = p[0], y = p[1], z = p[2];
        float *o = out + i * 3;
#pragma unroll
        for (int j = 0; j < 3; ++j)      // column j of R: T[0][j], T[1][j], T[2][j]; translation T[j][3]
            o[j] = fmaf(z, __ldg(T + 8 + j), fmaf(y, __ldg(T + 4 + j), fmaf(x, __ldg(T + j), __ldg(T + 4 * j + 3))));
    }
}

}  // namespace
}  // namespace sr

using namespace sr;

extern "C" int sr_mesh_pose_apply_f32(float *out, const float *verts, const float *pose, int64_t batch, int64_t nv,
                                      int64_t vert_stride, void *stream)
{
    SR_REQUIRE(out && verts && pose, "pose_apply: null pointer");
    SR_REQUIRE(batch >= 0 && nv >= 0 && vert_stride >= 3, "pose_apply: bad sizes");
    if (batch == 0 || nv == 0) return SR_OK;
    int64_t blocks = (batch * nv + 255) / 256;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    pose_apply_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(out, verts, pose, batch, nv, (int)vert_stride);
    count_launch();
    return check_launch("sr_mesh_pose_apply_f32");
}

// The mesh front-end of GeneratorWithMap as ONE call (reference face_model.py:71-72 ->
// utils_3d.py:360-404 -> model.py:260-270): morphable-model output -> pose -> area-weighted vertex normals -> the normal
// map at every resolution of the generator, written as [b, 3, s, s] planes.  Forward only: no index / coefficient
// buffers are written (levels[i].ids / .bary may be NULL), which is the state of the training loop (the mesh is sampled
// under no_grad, reference train.py:249-251).  5 launches: pose, face scatter, normalise, triangle pass, resolve.
extern "C" int sr_mesh_normal_pyramid_f32(int64_t batch, int64_t nv, int64_t nf, const float *verts_in, int64_t vert_stride,
                                          const float *pose, const int64_t *tris, float *verts_out, float *normals,
                                          int n_levels, const sr_raster_level *levels, uint64_t *keys, float eps,
                                          void *stream)
{
    SR_REQUIRE(verts_in && tris && normals && levels && keys, "mesh_normal_pyramid: null pointer");
    SR_REQUIRE(pose == nullptr || verts_out != nullptr, "mesh_normal_pyramid: a pose needs verts_out");
    const float *v = verts_in;
    if (pose) {
        const int rc = sr_mesh_pose_apply_f32(verts_out, verts_in, pose, batch, nv, vert_stride, stream);
        if (rc != SR_OK) return rc;
        v = verts_out;
    } else {
        SR_REQUIRE(vert_stride == 3, "mesh_normal_pyramid: without a pose the vertices must be packed [b,n,3]");
    }
    int rc = sr_mesh_vertex_normals_f32(normals, v, tris, batch, nv, nf, 1, 1e-8f, stream);
    if (rc != SR_OK) return rc;
    return sr_rasterize_pyramid_maps_f32(batch, nv, nf, n_levels, levels, 0, 1, 0, v, tris, keys, eps, normals, 3, 1, stream);
}

extern "C" int sr_mesh_vertex_normals_f32(float *normals, const float *verts, const int64_t *tris, int64_t batch, int64_t nv,
                                          int64_t nf, int shared_f, float eps, void *stream)
{
    SR_REQUIRE(normals && verts && (tris || nf == 0), "vertex_normals: null pointer");
    SR_REQUIRE(batch >= 0 && nv >= 0 && nf >= 0, "vertex_normals: negative size");
    if (batch == 0 || nv == 0) return SR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(normals, 0, sizeof(float) * (size_t)(batch * nv * 3), st);
    if (e != cudaSuccess) { set_error("vertex_normals: memset: %s", cudaGetErrorString(e)); return (int)e; }
    if (nf > 0) {
        int64_t blocks = (batch * nf + 255) / 256;
        if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
        face_normal_scatter_kernel<<<(unsigned)blocks, 256, 0, st>>>(normals, verts, tris, batch, nv, nf, shared_f);
    }
    int64_t blocks = (batch * nv + 255) / 256;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    normalize3_kernel<<<(unsigned)blocks, 256, 0, st>>>(normals, batch * nv, eps);
    count_launch(nf > 0 ? 2 : 1);
    return check_launch("sr_mesh_vertex_normals_f32");
}
