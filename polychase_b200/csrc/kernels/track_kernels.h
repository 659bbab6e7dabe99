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

// inverse of a general 4x4 (row-major) in double, Gauss-Jordan with partial pivoting; false if
// singular
PC_HD bool invert4x4(const double m[16], double inv[16]) {
    double a[4][8];
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) { a[r][c] = m[r * 4 + c]; a[r][4 + c] = r == c ? 1.0 : 0.0; }
    for (int col = 0; col < 4; col++) {
        int piv = col;
        for (int r = col + 1; r < 4; r++)
            if (fabs(a[r][col]) > fabs(a[piv][col])) piv = r;
        if (a[piv][col] == 0.0) return false;
        if (piv != col)
            for (int k = 0; k < 8; k++) { const double t = a[piv][k]; a[piv][k] = a[col][k]; a[col][k] = t; }
        const double d = 1.0 / a[col][col];
        for (int k = 0; k < 8; k++) a[col][k] *= d;
        for (int r = 0; r < 4; r++)
            if (r != col) {
                const double f = a[r][col];
                if (f != 0.0)
                    for (int k = 0; k < 8; k++) a[r][k] -= f * a[col][k];
            }
    }
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) inv[r * 4 + c] = a[r][4 + c];
    return true;
}

// GetRayObjectSpace (ray_casting.h:53-63): mat = (view * model)^-1, origin = mat.col(3),
// dir = mat[3x3] * Unproject(pos).  The reference inverts in float32 (Eigen general inverse);
// here the inverse is formed in double and rounded once.  Fills origin / dir_mat / intrinsics.
PC_HD bool make_ray_source(const pc_camera_state& cam, const float model[16], RaySource& s) {
    const Cam c = make_cam(cam);
    double view[16] = {c.R.m[0], c.R.m[1], c.R.m[2], c.t.x, c.R.m[3], c.R.m[4], c.R.m[5], c.t.y,
                       c.R.m[6], c.R.m[7], c.R.m[8], c.t.z, 0, 0, 0, 1};
    double vm[16], inv[16];
    for (int r = 0; r < 4; r++)
        for (int cc = 0; cc < 4; cc++) {
            double acc = 0;
            for (int k = 0; k < 4; k++) acc += view[r * 4 + k] * (double)model[k * 4 + cc];
            vm[r * 4 + cc] = acc;
        }
    if (!invert4x4(vm, inv)) return false;
    s.origin = V3{(float)inv[3], (float)inv[7], (float)inv[11]};
    for (int r = 0; r < 3; r++)
        for (int cc = 0; cc < 3; cc++) s.dir_mat.m[r * 3 + cc] = (float)inv[r * 4 + cc];
    s.fx = c.fx; s.fy = c.fy; s.cx = c.cx; s.cy = c.cy; s.sgn = c.sgn;
    return true;
}

// A source of the device-resident tracking chain: everything, including the row count and the
// source camera, is read from device memory when the kernel runs.
struct ResidentSource {
    const pc_camera_state* cam;  // device: pose of the source frame (written by an earlier PnP launch)
    const float* keypoints;      // device: source frame keypoints
    const uint32_t* indices;     // device: compacted flow rows (keypoint index)
    const float* targets;        // device: compacted flow rows (target position)
    const int* rows;             // device: number of rows
};
struct ResidentSources {
    ResidentSource s[8];
    int nsrc;
    int cap;                     // rows reserved per source in the output arrays
    float model[16];
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

// cam_in / cam_out / result are device pointers (cam_in may equal cam_out); nothing is read
// back, so launches chain on a stream without host round trips.
// Device-resident variant of K10 for the fused analyze->track chain: output row of (source s, row j)
// is s * cap + j; rows beyond a source's count are marked invalid.
void launch_raycast_resident(const MeshView& mesh, const ResidentSources& srcs, float* X_out, float* x_out,
                             uint8_t* valid, cudaStream_t s);

void launch_pnp_lm(const float* X, const float* x, const float* w, const uint8_t* valid, int m, const PnpParams& prm,
                   const pc_camera_state* cam_in, pc_camera_state* cam_out, PnpResult* result, cudaStream_t s);
int pnp_cluster_size();

}  // namespace pc
