// Launcher interface of the track kernels (K10 ray cast, K11 PnP LM).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "geom.cuh"

namespace pc {

struct BvhView {
    const float4* nodes;   // 2 float4 per node: (bmin.xyz, left|first), (bmax.xyz, count); count>0 = leaf
    const float4* tris;    // 3 float4 per leaf-ordered triangle: (p1.xyz, prim id), (p2.xyz,0), (p3.xyz,0)
    int num_nodes;
};

struct MeshView {
    BvhView bvh;
    const float* verts;    // nv x 3
    const uint32_t* tris;  // nt x 3
    const uint32_t* mask;  // bitfield or nullptr
    int nv, nt;
};

// One source camera's rays: origin and direction matrix of GetRayObjectSpace
// (ray_casting.h:53-63), the source keypoints and which of them to cast.
struct RaySource {
    V3 origin;
    M3 dir_mat;
    float fx, fy, cx, cy, sgn;
    const float* keypoints;     // device, nk x 2
    const uint32_t* indices;    // device, rows (nullptr: cast keypoints[0..rows))
    const float* targets;       // device, rows x 2 (may be nullptr)
    int first;                  // offset of this source's rows in the output arrays
    int rows;
};

struct PnpParams {
    unsigned long long max_iterations;
    int loss_type;
    float loss_scale, gradient_tol, step_tol, initial_lambda, min_lambda, max_lambda;
    float max_inlier_error;
    int opt_f, opt_pp;
    Bounds bounds;
};

struct PnpResult {
    pc_bundle_stats stats;
    float inlier_ratio;
    int num_matches;
    int status;      // 0 ok, 1 not enough features
};

void launch_raycast_sources(const MeshView& mesh, const RaySource* srcs_dev, int nsrc, int total, int check_mask,
                            const float model[16], float* X_out, float* x_out, uint8_t* valid, uint32_t* prim_out,
                            float* uv_out, float* t_out, float* pos_out, cudaStream_t s);

void launch_pnp_lm(const float* X, const float* x, const float* w, const uint8_t* valid, int m, const PnpParams& prm,
                   pc_camera_state* cam_io, PnpResult* result, cudaStream_t s);

}  // namespace pc
