// Camera / pose / quaternion / loss math shared by the track and refine kernels and their
// host-side drivers (all functions are __host__ __device__, float32 like the reference's
// `using Float = float`, /root/reference/cpp/eigen_typedefs.h:13).
//
//   CameraIntrinsics   /root/reference/cpp/pnp/types.h:18-198
//   Pose               /root/reference/cpp/pose.h:9-160
//   QuatStepPost       /root/reference/cpp/pnp/quaternion.h:11-20
//   robust losses      /root/reference/cpp/pnp/robust_loss.h:47-104
// Eigen conventions (toRotationMatrix, Quaternion(Matrix3), q * AngleAxis): SURVEY.md App. C.
#pragma once

#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <stdint.h>

#include "../../../include/polychase_b200.h"

namespace pc {

#define PC_HD __host__ __device__ __forceinline__

struct V3 { float x, y, z; };
struct M3 { float m[9]; };   // row-major

PC_HD V3 v3(float x, float y, float z) { return V3{x, y, z}; }
PC_HD V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
PC_HD V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
PC_HD V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
PC_HD V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
PC_HD float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
PC_HD V3 cross(V3 a, V3 b) { return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
PC_HD V3 mul(const M3& A, V3 v) {
    return V3{A.m[0] * v.x + A.m[1] * v.y + A.m[2] * v.z, A.m[3] * v.x + A.m[4] * v.y + A.m[5] * v.z,
              A.m[6] * v.x + A.m[7] * v.y + A.m[8] * v.z};
}
PC_HD V3 mul_t(const M3& A, V3 v) {   // A^T v
    return V3{A.m[0] * v.x + A.m[3] * v.y + A.m[6] * v.z, A.m[1] * v.x + A.m[4] * v.y + A.m[7] * v.z,
              A.m[2] * v.x + A.m[5] * v.y + A.m[8] * v.z};
}

// Eigen::Quaternionf::toRotationMatrix, q = (w,x,y,z)
PC_HD M3 quat_to_matrix(const float q[4]) {
    const float w = q[0], x = q[1], y = q[2], z = q[3];
    const float tx = 2.f * x, ty = 2.f * y, tz = 2.f * z;
    const float twx = tx * w, twy = ty * w, twz = tz * w;
    const float txx = tx * x, txy = ty * x, txz = tz * x;
    const float tyy = ty * y, tyz = tz * y, tzz = tz * z;
    M3 R;
    R.m[0] = 1.f - (tyy + tzz); R.m[1] = txy - twz; R.m[2] = txz + twy;
    R.m[3] = txy + twz; R.m[4] = 1.f - (txx + tzz); R.m[5] = tyz - twx;
    R.m[6] = txz - twy; R.m[7] = tyz + twx; R.m[8] = 1.f - (txx + tyy);
    return R;
}

// Eigen::Quaternionf(Matrix3f) -- Shepperd's method; q out = (w,x,y,z)
PC_HD void quat_from_matrix(const float m[9], float q[4]) {
    float t = m[0] + m[4] + m[8];
    if (t > 0.f) {
        t = sqrtf(t + 1.f);
        q[0] = 0.5f * t;
        t = 0.5f / t;
        q[1] = (m[7] - m[5]) * t;
        q[2] = (m[2] - m[6]) * t;
        q[3] = (m[3] - m[1]) * t;
    } else {
        int i = 0;
        if (m[4] > m[0]) i = 1;
        if (m[8] > m[i * 3 + i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrtf(m[i * 3 + i] - m[j * 3 + j] - m[k * 3 + k] + 1.f);
        q[1 + i] = 0.5f * t;
        t = 0.5f / t;
        q[0] = (m[k * 3 + j] - m[j * 3 + k]) * t;
        q[1 + j] = (m[j * 3 + i] + m[i * 3 + j]) * t;
        q[1 + k] = (m[k * 3 + i] + m[i * 3 + k]) * t;
    }
}

// q <- q * AngleAxis(|w|, w/|w|)   (no renormalisation, like the reference)
PC_HD void quat_step_post(const float q[4], V3 w, float out[4]) {
    const float angle = sqrtf(dot(w, w));
    if (angle > 0.f) {
        const V3 ax = V3{w.x / angle, w.y / angle, w.z / angle};
        const float half = 0.5f * angle;
        const float c = cosf(half), s = sinf(half);
        const float bw = c, bx = s * ax.x, by = s * ax.y, bz = s * ax.z;
        const float aw = q[0], axx = q[1], ay = q[2], az = q[3];
        out[0] = aw * bw - axx * bx - ay * by - az * bz;
        out[1] = aw * bx + axx * bw + ay * bz - az * by;
        out[2] = aw * by + ay * bw + az * bx - axx * bz;
        out[3] = aw * bz + az * bw + axx * by - ay * bx;
    } else {
        out[0] = q[0]; out[1] = q[1]; out[2] = q[2]; out[3] = q[3];
    }
}

struct Cam {           // unpacked pc_camera_state with cached rotation matrix
    float fx, fy, cx, cy, aspect;
    float sgn;         // +1 OpenCV, -1 OpenGL (types.h:95-98)
    M3 R;
    V3 t;
};

PC_HD Cam make_cam(const pc_camera_state& s) {
    Cam c;
    c.fx = s.fx; c.fy = s.fy; c.cx = s.cx; c.cy = s.cy; c.aspect = s.aspect_ratio;
    c.sgn = (s.convention != 0.f) ? 1.f : -1.f;
    c.R = quat_to_matrix(s.q);
    c.t = V3{s.t[0], s.t[1], s.t[2]};
    return c;
}
PC_HD bool is_behind(const Cam& c, V3 p) { return c.sgn > 0.f ? p.z < 0.f : p.z > 0.f; }   // types.h:129-132
PC_HD V3 unproject(const Cam& c, float x, float y) {                                          // types.h:95-98
    return V3{c.sgn * ((x - c.cx) / c.fx), c.sgn * ((y - c.cy) / c.fy), c.sgn};
}
PC_HD V3 cam_center(const Cam& c) { return -mul_t(c.R, c.t); }                                // pose.h:46

// Loss / Weight of the robust losses on the squared residual norm.
struct Loss {
    int kind;          // 0 trivial, 1 huber, 2 cauchy
    float thr, sq_thr, inv_sq_thr;
};
PC_HD Loss make_loss(int kind, float scale) {
    Loss l;
    l.kind = kind; l.thr = scale; l.sq_thr = scale * scale;
    l.inv_sq_thr = (float)(1.0 / (double)l.sq_thr);
    return l;
}
PC_HD float loss_value(const Loss& l, float r2) {
    if (l.kind == 0) return r2;
    if (l.kind == 1) {
        if (r2 <= l.sq_thr) return r2;
        const float r = sqrtf(r2);
        return (float)((double)l.thr * (2.0 * (double)r - (double)l.thr));
    }
    return l.sq_thr * log1pf(r2 * l.inv_sq_thr);
}
PC_HD float loss_weight(const Loss& l, float r2) {
    if (l.kind == 0) return 1.f;
    if (l.kind == 1) return r2 <= l.sq_thr ? 1.f : l.thr / sqrtf(r2);
    const float w = 1.f / (1.f + r2 * l.inv_sq_thr);
    return w > FLT_MIN ? w : FLT_MIN;
}

// CameraIntrinsics::GetBounds (types.h:156-192)
struct Bounds { float f_low, f_high, cx_low, cx_high, cy_low, cy_high; };
inline Bounds get_bounds(const pc_camera_state& s, float min_fov_deg = 15.f, float max_fov_deg = 160.f) {
    const float min_fov = (float)(min_fov_deg * M_PI / 180), max_fov = (float)(max_fov_deg * M_PI / 180);
    const float tmin = tanf(min_fov / 2), tmax = tanf(max_fov / 2);
    Bounds b;
    if (s.convention == 0.f) { b.f_low = -(s.width / 2.0f) / tmin; b.f_high = -(s.width / 2.0f) / tmax; }
    else { b.f_high = (s.width / 2.0f) / tmin; b.f_low = (s.width / 2.0f) / tmax; }
    b.cx_low = 0.f; b.cx_high = s.width; b.cy_low = 0.f; b.cy_high = s.height;
    return b;
}
PC_HD float clampf(float v, float lo, float hi) { return v < lo ? lo : (hi < v ? hi : v); }   // std::clamp

// PnPProblem::Step / RefinementProblemBase::Step (pnp_problem.h:101-131, refiner.cc:508-537)
PC_HD void camera_step(const pc_camera_state& in, const float* dp, bool opt_f, bool opt_pp, const Bounds& b,
                       pc_camera_state& out) {
    out = in;
    quat_step_post(in.q, V3{dp[0], dp[1], dp[2]}, out.q);
    out.t[0] = in.t[0] + dp[3]; out.t[1] = in.t[1] + dp[4]; out.t[2] = in.t[2] + dp[5];
    if (opt_f) {
        out.fy = in.fy + dp[6];
        out.fx = out.fy * out.aspect_ratio;
        out.fy = clampf(out.fy, b.f_low, b.f_high);
        out.fx = clampf(out.fx, b.f_low, b.f_high);
    }
    if (opt_pp) {
        out.cx = clampf(in.cx + dp[7], b.cx_low, b.cx_high);
        out.cy = clampf(in.cy + dp[8], b.cy_low, b.cy_high);
    }
}

}  // namespace pc
