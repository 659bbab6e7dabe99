// K10: nearest-hit ray casting of source keypoints onto the mesh (K11, the LM pose solve, is in
// pnp_lm.cu).
//
// K10 replaces Embree's rtcIntersect1 behind AcceleratedMesh::RayCast
// (/root/reference/cpp/ray_casting.cc:65-121) and the per-match loop of SolveFrame
// (/root/reference/cpp/tracker.cc:64-92): a stack-based traversal of a BVH built once per
// mesh, one thread per ray, nearest hit with tnear = 0, masked triangle = miss
// (ray_casting.cc:106-108), position = (1-u-v) p1 + u p2 + v p3 (geometry.h:17-19).
//
#include <cooperative_groups.h>

#include "common.cuh"
#include "geom.cuh"
#include "kernels.h"
#include "track_kernels.h"
#include "bvh.cuh"

namespace cg = cooperative_groups;

namespace pc {

__device__ __forceinline__ V3 barycentric_pos(const MeshView& mesh, int prim, float u, float v) {
    const uint32_t i0 = mesh.tris[3 * prim], i1 = mesh.tris[3 * prim + 1], i2 = mesh.tris[3 * prim + 2];
    const V3 p1 = v3(mesh.verts[3 * i0], mesh.verts[3 * i0 + 1], mesh.verts[3 * i0 + 2]);
    const V3 p2 = v3(mesh.verts[3 * i1], mesh.verts[3 * i1 + 1], mesh.verts[3 * i1 + 2]);
    const V3 p3 = v3(mesh.verts[3 * i2], mesh.verts[3 * i2 + 1], mesh.verts[3 * i2 + 2]);
    const float w = (float)(1.0 - (double)u - (double)v);     // geometry.h:18: (1.0 - u - v) is double
    return p1 * w + p2 * u + p3 * v;
}

__device__ __forceinline__ bool tri_masked(const MeshView& mesh, int prim) {
    return mesh.mask != nullptr && ((mesh.mask[prim >> 5] >> (prim & 31)) & 1u);
}

__global__ void __launch_bounds__(128) raycast_sources_kernel(MeshView mesh, const RaySource* __restrict__ srcs,
                                                              int nsrc, int total, int check_mask, M3 model_r,
                                                              V3 model_t, float* __restrict__ X_out,
                                                              float* __restrict__ x_out, uint8_t* __restrict__ valid,
                                                              uint32_t* __restrict__ prim_out,
                                                              float* __restrict__ uv_out, float* __restrict__ t_out,
                                                              float* __restrict__ pos_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int s = 0;
    while (s + 1 < nsrc && i >= srcs[s + 1].first) s++;
    const RaySource& S = srcs[s];
    const int j = i - S.first;
    const int kp = S.indices ? (int)S.indices[j] : j;
    const float px = S.keypoints[2 * kp], py = S.keypoints[2 * kp + 1];
    // GetRayObjectSpace (ray_casting.h:53-63): dir = inv(V*M)[3x3] * Unproject(pos)
    const V3 dc = V3{S.sgn * ((px - S.cx) / S.fx), S.sgn * ((py - S.cy) / S.fy), S.sgn};
    const V3 d = mul(S.dir_mat, dc);
    const HitRec h = bvh_nearest_hit(mesh.bvh, S.origin, d);
    bool ok = h.prim >= 0;
    if (ok && check_mask && tri_masked(mesh, h.prim)) ok = false;
    V3 pos = v3(0, 0, 0);
    if (ok) pos = barycentric_pos(mesh, h.prim, h.u, h.v);
    if (valid) valid[i] = ok ? 1 : 0;
    if (X_out) {                                  // tracker.cc:80-82: world = M3x3 * pos + Mt
        const V3 wv = ok ? mul(model_r, pos) + model_t : v3(0, 0, 0);
        X_out[3 * i] = wv.x; X_out[3 * i + 1] = wv.y; X_out[3 * i + 2] = wv.z;
    }
    if (x_out && S.targets) { x_out[2 * i] = S.targets[2 * j]; x_out[2 * i + 1] = S.targets[2 * j + 1]; }
    if (pos_out) { pos_out[3 * i] = pos.x; pos_out[3 * i + 1] = pos.y; pos_out[3 * i + 2] = pos.z; }
    if (prim_out) prim_out[i] = ok ? (uint32_t)h.prim : 0xFFFFFFFFu;
    if (uv_out) { uv_out[2 * i] = h.u; uv_out[2 * i + 1] = h.v; }
    if (t_out) t_out[i] = ok ? h.t : 0.f;
}

void launch_raycast_sources(const MeshView& mesh, const RaySource* srcs_dev, int nsrc, int total, int check_mask,
                            const float model[16], float* X_out, float* x_out, uint8_t* valid, uint32_t* prim_out,
                            float* uv_out, float* t_out, float* pos_out, cudaStream_t s) {
    if (total <= 0) return;
    M3 mr;
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) mr.m[r * 3 + c] = model[r * 4 + c];
    const V3 mt = v3(model[3], model[7], model[11]);
    raycast_sources_kernel<<<(total + 127) / 128, 128, 0, s>>>(mesh, srcs_dev, nsrc, total, check_mask, mr, mt, X_out,
                                                                x_out, valid, prim_out, uv_out, t_out, pos_out);
}

// ---- K10, device-resident: sources described by device pointers (fused analyze -> track chain) ----
// Every block first derives its source's ray frame (inverse of view * model, in double) from the
// source camera that an earlier PnP launch left in device memory, then casts one ray per thread.
__global__ void __launch_bounds__(128) raycast_resident_kernel(MeshView mesh, ResidentSources S,
                                                               float* __restrict__ X_out, float* __restrict__ x_out,
                                                               uint8_t* __restrict__ valid) {
    __shared__ RaySource rs;
    __shared__ int ok_src;
    const ResidentSource& src = S.s[blockIdx.y];
    if (threadIdx.x == 0) {
        const pc_camera_state cam = *src.cam;
        ok_src = make_ray_source(cam, S.model, rs) ? 1 : 0;
    }
    __syncthreads();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= S.cap) return;
    const int i = blockIdx.y * S.cap + j;
    const int rows = min(*src.rows, S.cap);
    bool ok = ok_src && j < rows;
    V3 wv = v3(0, 0, 0);
    float tx = 0.f, ty = 0.f;
    if (ok) {
        const int kp = (int)src.indices[j];
        const float px = src.keypoints[2 * kp], py = src.keypoints[2 * kp + 1];
        const V3 dc = V3{rs.sgn * ((px - rs.cx) / rs.fx), rs.sgn * ((py - rs.cy) / rs.fy), rs.sgn};
        const V3 d = mul(rs.dir_mat, dc);
        const HitRec h = bvh_nearest_hit(mesh.bvh, rs.origin, d);
        ok = h.prim >= 0 && !tri_masked(mesh, h.prim);
        if (ok) {                                 // tracker.cc:80-82: world = M3x3 * pos + Mt
            const V3 pos = barycentric_pos(mesh, h.prim, h.u, h.v);
            M3 mr;
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
                for (int c = 0; c < 3; c++) mr.m[r * 3 + c] = S.model[r * 4 + c];
            wv = mul(mr, pos) + v3(S.model[3], S.model[7], S.model[11]);
        }
        tx = src.targets[2 * j]; ty = src.targets[2 * j + 1];
    }
    valid[i] = ok ? 1 : 0;
    X_out[3 * i] = wv.x; X_out[3 * i + 1] = wv.y; X_out[3 * i + 2] = wv.z;
    x_out[2 * i] = tx; x_out[2 * i + 1] = ty;
}

void launch_raycast_resident(const MeshView& mesh, const ResidentSources& srcs, float* X_out, float* x_out,
                             uint8_t* valid, cudaStream_t s) {
    if (srcs.nsrc <= 0 || srcs.cap <= 0) return;
    dim3 grid((srcs.cap + 127) / 128, srcs.nsrc);
    raycast_resident_kernel<<<grid, 128, 0, s>>>(mesh, srcs, X_out, x_out, valid);
}

}  // namespace pc
