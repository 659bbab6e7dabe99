// FindTransformation (/root/reference/cpp/pin_mode.cc:16-246) over the C ABI's PnP solve.
#include "pin_mode.h"

#include <cmath>
#include <cstring>

namespace pch {

namespace {

struct V3 {
    float x, y, z;
};
V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
float Dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
V3 Cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
float Norm(V3 a) { return std::sqrt(Dot(a, a)); }
V3 Normalized(V3 a) {
    const float n = Norm(a);
    return n > 0 ? a * (1.0f / n) : a;
}

// 4x4 inverse (Eigen's .inverse() of a general 4x4; formed in double, returned in float)
Mat4 Inverse(const Mat4& m) {
    double a[4][8];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            a[i][j] = m[4 * i + j];
            a[i][4 + j] = i == j;
        }
    for (int c = 0; c < 4; c++) {
        int p = c;
        for (int r = c + 1; r < 4; r++)
            if (std::fabs(a[r][c]) > std::fabs(a[p][c])) p = r;
        if (a[p][c] == 0.0) throw std::runtime_error("FindTransformation: singular matrix");
        if (p != c)
            for (int j = 0; j < 8; j++) std::swap(a[c][j], a[p][j]);
        const double d = a[c][c];
        for (int j = 0; j < 8; j++) a[c][j] /= d;
        for (int r = 0; r < 4; r++)
            if (r != c) {
                const double f = a[r][c];
                for (int j = 0; j < 8; j++) a[r][j] -= f * a[c][j];
            }
    }
    Mat4 o;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) o[4 * i + j] = (float)a[i][4 + j];
    return o;
}

V3 TransformPoint(const Mat4& m, V3 p) {   // Eigen::Affine3f(m) * p
    return {m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3], m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7],
            m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11]};
}

V3 Unproject(const CameraIntrinsics& k, Vec2 x) {   // types.h:95-98
    const float s = k.convention == CameraConvention::OpenCV ? 1.0f : -1.0f;
    return {s * ((x[0] - k.cx) / k.fx), s * ((x[1] - k.cy) / k.fy), s};
}

struct Ray {
    V3 origin, dir;
};

Ray GetRayWorldSpace(const SceneTransformations& scene, Vec2 pos) {   // ray_casting.h:65-74
    const Mat4 mat = Inverse(scene.view_matrix);
    const V3 d = Unproject(scene.intrinsics, pos);
    return {{mat[3], mat[7], mat[11]},
            {mat[0] * d.x + mat[1] * d.y + mat[2] * d.z, mat[4] * d.x + mat[5] * d.y + mat[6] * d.z,
             mat[8] * d.x + mat[9] * d.y + mat[10] * d.z}};
}

[[noreturn]] void BadTransType(TransformationType t) {
    throw std::runtime_error("Invalid trans_type value: " + std::to_string(static_cast<int>(t)));
}

// pin_mode.cc:16-108
SceneTransformations FindTransformationN(const std::vector<float>& object_points, const SceneTransformations& initial,
                                         const SceneTransformations& current, const PinUpdate& update,
                                         TransformationType trans_type, bool optimize_focal_length,
                                         bool optimize_principal_point) {
    const size_t n = object_points.size() / 3;
    PCH_CHECK(n > 2);
    // Step 1: the pins in the initial camera's space and their projections
    const Mat4 model_view = MatMul(initial.view_matrix, initial.model_matrix);
    std::vector<float> cam_pts(3 * n), img_pts(2 * n);
    const CameraIntrinsics& k = initial.intrinsics;
    for (size_t i = 0; i < n; i++) {
        const V3 p = TransformPoint(model_view, {object_points[3 * i], object_points[3 * i + 1], object_points[3 * i + 2]});
        cam_pts[3 * i] = p.x; cam_pts[3 * i + 1] = p.y; cam_pts[3 * i + 2] = p.z;
        // image_points_3d = p * K^T, then divide by the third component (To3x3ProjectionMatrix, types.h:52-61)
        const float u = k.fx * p.x + k.cx * p.z, v = k.fy * p.y + k.cy * p.z, w = p.z;
        img_pts[2 * i] = u / w;
        img_pts[2 * i + 1] = v / w;
    }
    img_pts[2 * update.pin_idx] = update.pos[0];                              // apply the update
    img_pts[2 * update.pin_idx + 1] = update.pos[1];
    // Step 2: start from the current transform so that transitions are smooth
    const Mat4 initial_pose = MatMul(MatMul(current.view_matrix, current.model_matrix), Inverse(model_view));
    CameraState cam{current.intrinsics, Pose::FromRt(initial_pose)};
    BundleOptions bundle_opts;
    bundle_opts.loss_type = BundleOptions::LossType::TRIVIAL;
    const pc_bundle_opts bo = ToAbi(bundle_opts);
    pc_camera_state cs = ToAbi(cam);
    pc_bundle_stats stats{};
    float inlier_ratio = 0;
    {
        auto dc = AcquireInteractiveContext();       // a UI drag must not wait for a running track / refine pass
        std::lock_guard<std::mutex> lk(dc->mtx);
        const int rc = pc_solve_pnp(dc->ctx, cam_pts.data(), img_pts.data(), nullptr, (int)n, &bo, 0.0f /* no inlier ratio */,
                                    optimize_focal_length, optimize_principal_point, &cs, &stats, &inlier_ratio);
        if (rc != PC_OK) ThrowPcError(dc->ctx, rc);
    }
    const CameraState result = FromAbi(cs);
    const std::array<float, 9> R = result.pose.R();
    const Vec3 t = result.pose.t;
    switch (trans_type) {
        case TransformationType::Model: {
            Mat4 nmv{};                                                        // [R * mv_R | R * mv_t + t]
            for (int i = 0; i < 3; i++) {
                for (int j = 0; j < 3; j++)
                    nmv[4 * i + j] = R[3 * i] * model_view[j] + R[3 * i + 1] * model_view[4 + j] + R[3 * i + 2] * model_view[8 + j];
                nmv[4 * i + 3] = R[3 * i] * model_view[3] + R[3 * i + 1] * model_view[7] + R[3 * i + 2] * model_view[11] + t[i];
            }
            nmv[15] = 1.0f;
            return SceneTransformations{MatMul(Inverse(initial.view_matrix), nmv), current.view_matrix, result.intrinsics};
        }
        case TransformationType::Camera: {
            Mat4 upd{};
            for (int i = 0; i < 3; i++) {
                for (int j = 0; j < 3; j++) upd[4 * i + j] = R[3 * i + j];
                upd[4 * i + 3] = t[i];
            }
            upd[15] = 1.0f;
            return SceneTransformations{current.model_matrix, MatMul(upd, initial.view_matrix), result.intrinsics};
        }
        default:
            BadTransType(trans_type);
    }
}

// pin_mode.cc:110-149
SceneTransformations FindTransformation1(const std::vector<float>& object_points, const SceneTransformations& scene,
                                         const PinUpdate& update, TransformationType trans_type) {
    PCH_CHECK(object_points.size() == 3);
    const Ray ray = GetRayWorldSpace(scene, update.pos);
    const V3 point_world = TransformPoint(scene.model_matrix, {object_points[0], object_points[1], object_points[2]});
    const float depth = Norm(point_world - ray.origin);
    const V3 translated = ray.origin + Normalized(ray.dir) * depth;
    const V3 translation = translated - point_world;
    Mat4 new_model = scene.model_matrix;
    new_model[3] += translation.x;
    new_model[7] += translation.y;
    new_model[11] += translation.z;
    switch (trans_type) {
        case TransformationType::Model:
            return SceneTransformations{new_model, scene.view_matrix, scene.intrinsics};
        case TransformationType::Camera:
            return SceneTransformations{scene.model_matrix,
                                        MatMul(scene.view_matrix, MatMul(new_model, Inverse(scene.model_matrix))),
                                        scene.intrinsics};
        default:
            BadTransType(trans_type);
    }
}

// pin_mode.cc:151-217
SceneTransformations FindTransformation2(const std::vector<float>& object_points, const SceneTransformations& scene,
                                         const PinUpdate& update, TransformationType trans_type) {
    PCH_CHECK(object_points.size() == 6);
    const Ray ray = GetRayWorldSpace(scene, update.pos);
    const Mat4 view_inv = Inverse(scene.view_matrix);
    const V3 camera_center{view_inv[3], view_inv[7], view_inv[11]};
    const uint32_t mi = update.pin_idx, ai = 1 - update.pin_idx;
    const V3 moving = TransformPoint(scene.model_matrix, {object_points[3 * mi], object_points[3 * mi + 1], object_points[3 * mi + 2]});
    const V3 anchor = TransformPoint(scene.model_matrix, {object_points[3 * ai], object_points[3 * ai + 1], object_points[3 * ai + 2]});
    const float depth = Norm(moving - ray.origin);
    const V3 translated_moving = ray.origin + Normalized(ray.dir) * depth;
    const V3 du = moving - anchor, dv = translated_moving - anchor;
    const V3 dn_unit = Normalized({view_inv[2], view_inv[6], view_inv[10]});
    const V3 du_unit = Normalized(du), dv_unit = Normalized(dv);
    const float angle = std::atan2(Dot(Cross(du_unit, dv_unit), dn_unit), Dot(du_unit, dv_unit));
    // Eigen::AngleAxisf(angle, dn_unit) as a rotation matrix
    const float c = std::cos(angle), s = std::sin(angle), ic = 1.0f - c;
    const float x = dn_unit.x, y = dn_unit.y, z = dn_unit.z;
    const float Rm[9] = {c + x * x * ic, x * y * ic - z * s, x * z * ic + y * s,
                         y * x * ic + z * s, c + y * y * ic, y * z * ic - x * s,
                         z * x * ic - y * s, z * y * ic + x * s, c + z * z * ic};
    // scaling around the anchor = moving the anchor along the view ray by 1 / scale (the reference's own FIXME applies)
    const float scale_inv = Norm(du) / Norm(dv);
    const V3 new_anchor = camera_center + (anchor - camera_center) * scale_inv;
    // update = Translation(new_anchor) * rot * Translation(-anchor)
    Mat4 upd{};
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) upd[4 * i + j] = Rm[3 * i + j];
    }
    const V3 ra{Rm[0] * anchor.x + Rm[1] * anchor.y + Rm[2] * anchor.z, Rm[3] * anchor.x + Rm[4] * anchor.y + Rm[5] * anchor.z,
                Rm[6] * anchor.x + Rm[7] * anchor.y + Rm[8] * anchor.z};
    upd[3] = new_anchor.x - ra.x;
    upd[7] = new_anchor.y - ra.y;
    upd[11] = new_anchor.z - ra.z;
    upd[15] = 1.0f;
    switch (trans_type) {
        case TransformationType::Model:
            return SceneTransformations{MatMul(upd, scene.model_matrix), scene.view_matrix, scene.intrinsics};
        case TransformationType::Camera:
            return SceneTransformations{scene.model_matrix, MatMul(scene.view_matrix, upd), scene.intrinsics};
        default:
            BadTransType(trans_type);
    }
}

}  // namespace

SceneTransformations FindTransformation(const std::vector<float>& object_points, const SceneTransformations& initial,
                                        const SceneTransformations& current, const PinUpdate& update,
                                        TransformationType trans_type, bool optimize_focal_length,
                                        bool optimize_principal_point) {
    const size_t n = object_points.size() / 3;
    PCH_CHECK(object_points.size() % 3 == 0);
    PCH_CHECK(update.pin_idx < n);                                              // pin_mode.cc:225
    switch (n) {
        case 1:
            return FindTransformation1(object_points, initial, update, trans_type);
        case 2:   // not entirely correct in the reference either: it starts from the current transform (:231-235)
            return FindTransformation2(object_points, current, update, trans_type);
        default:
            return FindTransformationN(object_points, initial, current, update, trans_type, optimize_focal_length,
                                       optimize_principal_point);
    }
}

}  // namespace pch
