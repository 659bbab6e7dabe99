// Host-side value types of the reference's public surface (no Eigen / OpenCV):
//   VideoInfo, OpticalFlowOptions      /root/reference/cpp/opticalflow.h:20-33
//   GFTTOptions                        /root/reference/cpp/feature_detection/gftt.h:5-21
//   CameraIntrinsics, CameraState,
//   BundleOptions, BundleStats         /root/reference/cpp/pnp/types.h:13-225
//   Pose                               /root/reference/cpp/pose.h:9-160
//   CameraTrajectory                   /root/reference/cpp/camera_trajectory.h:14-91
//   Mesh, SceneTransformations         /root/reference/cpp/geometry.h:52-168
//   RayHit, AcceleratedMesh            /root/reference/cpp/ray_casting.h:15-50
#pragma once

#include <array>
#include <cmath>
#include <cstdint>
#include <limits>
#include <memory>
#include <mutex>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/polychase_b200.h"

namespace pch {

// CHECK macros of the reference throw std::logic_error with file:line (utils.h:12-39, utils.cc:8-19)
[[noreturn]] inline void assert_fail(const char* expr, const char* file, int line, const char* func) {
    throw std::logic_error(std::string("[") + file + ":" + std::to_string(line) + " " + func +
                           "] Assertion failed: " + expr);
}
#define PCH_CHECK(expr)                                                      \
    do {                                                                     \
        if (!static_cast<bool>(expr)) ::pch::assert_fail(#expr, __FILE__, __LINE__, __func__); \
    } while (0)

using Mat4 = std::array<float, 16>;   // row-major
using Vec2 = std::array<float, 2>;
using Vec3 = std::array<float, 3>;
using Keypoints = std::vector<Vec2>;
using KeypointsIndices = std::vector<uint32_t>;
using FlowErrors = std::vector<float>;

constexpr int32_t kInvalidId = std::numeric_limits<int32_t>::max();   // database.h:13

struct VideoInfo {
    uint32_t width = 0, height = 0;
    int32_t first_frame = 0;
    uint32_t num_frames = 0;
};

struct GFTTOptions {
    double quality_level = 0.01;
    double min_distance = 5.0;
    int block_size = 3;
    int gradient_size = 3;
    int max_corners = 0;
    bool use_harris = false;
    double harris_k = 0.04;
    int grid_rows = 4;
    int grid_cols = 4;
};

struct OpticalFlowOptions {
    int window_size = 10;
    int max_level = 3;
    int term_max_iters = 30;
    double term_epsilon = 0.01;
    double min_eigen_threshold = 1e-4;
};

enum class CameraConvention { OpenGL, OpenCV };
enum class TransformationType { Camera, Model };

struct CameraIntrinsics {
    float fx = 0, fy = 0, cx = 0, cy = 0, aspect_ratio = 1, width = 0, height = 0;
    CameraConvention convention = CameraConvention::OpenGL;

    // types.h:31-50 (near/far are placeholders there too)
    Mat4 To4x4ProjectionMatrix() const {
        constexpr float f = 100.0f, n = 10.0f;
        constexpr float p22 = -(f + n) / (f - n), p23 = -2.0f * f * n / (f - n);
        return Mat4{fx, 0, cx, 0, 0, fy, cy, 0, 0, 0, p22, p23, 0, 0, 1.0f, 0};
    }
};

struct Pose {
    std::array<float, 4> q{1, 0, 0, 0};   // w, x, y, z
    Vec3 t{0, 0, 0};

    std::array<float, 9> R() const;
    Mat4 Rt4x4() const;
    static Pose FromRt(const Mat4& m);     // pose.h:133-136
};

struct CameraState {
    CameraIntrinsics intrinsics;
    Pose pose;
};

struct BundleOptions {
    size_t max_iterations = 100;
    size_t max_allowed_parallelism = 8;
    enum class LossType { TRIVIAL, HUBER, CAUCHY } loss_type = LossType::HUBER;
    float loss_scale = 1.0f;
    float gradient_tol = 1e-10f;
    float step_tol = 1e-8f;
    float initial_lambda = 1e-5f;
    float min_lambda = 1e-10f;
    float max_lambda = 1e10f;
    bool verbose = false;
};

struct BundleStats {
    size_t iterations = 0;
    float initial_cost = 0, cost = 0, lambda = 0;
    size_t invalid_steps = 0;
    float step_norm = 0, grad_norm = 0;
};

struct PnPResult {
    CameraState camera;
    BundleStats bundle_stats;
    float inlier_ratio = 0.0f;
};

class CameraTrajectory {
   public:
    CameraTrajectory() = default;
    CameraTrajectory(int32_t first_frame_id, size_t count) : states(count), first_frame_id(first_frame_id) {}
    bool IsValidFrame(int32_t frame_id) const { return Index(frame_id) < Count(); }
    bool IsFrameFilled(int32_t frame_id) const { return IsValidFrame(frame_id) && Get(frame_id).has_value(); }
    const std::optional<CameraState>& Get(int32_t frame_id) const {
        const size_t index = Index(frame_id);
        PCH_CHECK(index < Count());
        return states[index];
    }
    void Set(int32_t frame_id, const CameraState& state) {
        const size_t index = Index(frame_id);
        PCH_CHECK(index < Count());
        states[index] = state;
    }
    void Clear(int32_t frame_id) {
        const size_t index = Index(frame_id);
        PCH_CHECK(index < Count());
        states[index] = std::nullopt;
    }
    size_t Count() const { return states.size(); }
    int32_t FirstFrame() const { return first_frame_id; }
    int32_t LastFrame() const { return first_frame_id + (int32_t)states.size() - 1; }
    size_t Index(int32_t frame_id) const { return static_cast<size_t>(frame_id - first_frame_id); }

   private:
    std::vector<std::optional<CameraState>> states;
    int32_t first_frame_id = 0;
};

struct Bbox3 { Vec3 pmin, pmax; };

struct Mesh {                                   // geometry.h:52-152
    std::vector<float> vertices;                // nv x 3
    std::vector<uint32_t> triangles;            // nt x 3
    std::vector<uint32_t> masked_triangles;     // bitfield, padded to a multiple of 4 words
    Bbox3 bbox{};
    Mesh(std::vector<float> v, std::vector<uint32_t> t, std::vector<uint32_t> m);
    size_t NumVertices() const { return vertices.size() / 3; }
    size_t NumTriangles() const { return triangles.size() / 3; }
    bool IsTriangleMasked(uint32_t tri_idx) const;
    void MaskTriangle(uint32_t tri_idx);
    void UnmaskTriangle(uint32_t tri_idx);
    void ToggleMaskTriangle(uint32_t tri_idx);
};

struct SceneTransformations {
    Mat4 model_matrix{};
    Mat4 view_matrix{};
    CameraIntrinsics intrinsics;
};

struct RayHit {
    Vec3 pos{}, normal{};
    Vec2 barycentric_coordinate{};
    float t = 0;
    uint32_t primitive_id = 0;
};

// One GPU context shared by everything in the process that asks for the same device
// (POLYCHASE_DEVICE, default 0).  Calls on a context are serialised by `mtx`.
struct DeviceContext {
    pc_ctx* ctx = nullptr;
    std::mutex mtx;
    uint64_t mesh_epoch = 0;        // which AcceleratedMesh is currently uploaded
    ~DeviceContext();
};
std::shared_ptr<DeviceContext> AcquireDeviceContext(int max_width, int max_height, int max_features);
// The context of the interactive entry points (ray_cast from the UI thread, pinned frame buffers):
// separate from the solver context, so a UI ray cast never waits for a running track / refine pass
// (the reference's Embree ray cast is independent of its solvers too, ray_casting.cc:128-133).
std::shared_ptr<DeviceContext> AcquireInteractiveContext();
[[noreturn]] void ThrowPcError(pc_ctx* ctx, int code);

class AcceleratedMesh {                          // ray_casting.h:23-50; BVH lives on the GPU
   public:
    AcceleratedMesh(std::vector<float> vertices, std::vector<uint32_t> triangles,
                    std::vector<uint32_t> masked_triangles);
    AcceleratedMesh(const AcceleratedMesh&) = delete;
    AcceleratedMesh& operator=(const AcceleratedMesh&) = delete;
    const Mesh& Inner() const { return mesh_; }
    Mesh& InnerMut();    // the caller may edit the mask: every context re-uploads the mesh on its next Bind
    // Makes this mesh the context's current mesh (uploads + builds the BVH when needed).
    void Bind(DeviceContext& dc) const;
    std::optional<RayHit> RayCast(const SceneTransformations& scene, Vec2 pos, bool check_mask) const;

   private:
    Mesh mesh_;
    uint64_t epoch_;
};

// conversions to the C ABI records
pc_camera_state ToAbi(const CameraState& s);
CameraState FromAbi(const pc_camera_state& s);
pc_bundle_opts ToAbi(const BundleOptions& o);
BundleStats FromAbi(const pc_bundle_stats& s);
pc_gftt_opts ToAbi(const GFTTOptions& o);
pc_flow_opts ToAbi(const OpticalFlowOptions& o);

Mat4 MatMul(const Mat4& a, const Mat4& b);

}  // namespace pch
