#include "types.h"

#include <atomic>
#include <cstdlib>
#include <cstring>

namespace pch {

std::array<float, 9> Pose::R() const {      // Eigen::Quaternionf::toRotationMatrix
    const float w = q[0], x = q[1], y = q[2], z = q[3];
    const float tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const float twx = tx * w, twy = ty * w, twz = tz * w;
    const float txx = tx * x, txy = ty * x, txz = tz * x;
    const float tyy = ty * y, tyz = tz * y, tzz = tz * z;
    return {1 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1 - (txx + tzz), tyz - twx,
            txz - twy, tyz + twx, 1 - (txx + tyy)};
}

Mat4 Pose::Rt4x4() const {
    const auto r = R();
    return Mat4{r[0], r[1], r[2], t[0], r[3], r[4], r[5], t[1], r[6], r[7], r[8], t[2], 0, 0, 0, 1};
}

Pose Pose::FromRt(const Mat4& M) {            // Eigen::Quaternionf(Matrix3f), Shepperd's method
    const float m[9] = {M[0], M[1], M[2], M[4], M[5], M[6], M[8], M[9], M[10]};
    Pose p;
    float tr = m[0] + m[4] + m[8];
    if (tr > 0.f) {
        float s = std::sqrt(tr + 1.f);
        p.q[0] = 0.5f * s;
        s = 0.5f / s;
        p.q[1] = (m[7] - m[5]) * s;
        p.q[2] = (m[2] - m[6]) * s;
        p.q[3] = (m[3] - m[1]) * s;
    } else {
        int i = 0;
        if (m[4] > m[0]) i = 1;
        if (m[8] > m[i * 3 + i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        float s = std::sqrt(m[i * 3 + i] - m[j * 3 + j] - m[k * 3 + k] + 1.f);
        p.q[1 + i] = 0.5f * s;
        s = 0.5f / s;
        p.q[0] = (m[k * 3 + j] - m[j * 3 + k]) * s;
        p.q[1 + j] = (m[j * 3 + i] + m[i * 3 + j]) * s;
        p.q[1 + k] = (m[k * 3 + i] + m[i * 3 + k]) * s;
    }
    p.t = {M[3], M[7], M[11]};
    return p;
}

Mat4 MatMul(const Mat4& a, const Mat4& b) {
    Mat4 o{};
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) {
            float s = 0;
            for (int k = 0; k < 4; k++) s += a[r * 4 + k] * b[k * 4 + c];
            o[r * 4 + c] = s;
        }
    return o;
}

Mesh::Mesh(std::vector<float> v, std::vector<uint32_t> t, std::vector<uint32_t> m)
    : vertices(std::move(v)), triangles(std::move(t)), masked_triangles(std::move(m)) {
    const int mask_num_ints = (int)((NumTriangles() + 31) / 32);
    const int padded = mask_num_ints + (4 - mask_num_ints % 4) % 4;      // geometry.h:63-65
    if (masked_triangles.empty()) masked_triangles.assign(padded, 0u);
    PCH_CHECK((int)masked_triangles.size() >= padded);                  // geometry.h:72
    Vec3 pmin{std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
    Vec3 pmax{std::numeric_limits<float>::lowest(), std::numeric_limits<float>::lowest(),
              std::numeric_limits<float>::lowest()};
    for (size_t i = 0; i < NumVertices(); i++)
        for (int k = 0; k < 3; k++) {
            pmin[k] = std::min(pmin[k], vertices[3 * i + k]);
            pmax[k] = std::max(pmax[k], vertices[3 * i + k]);
        }
    bbox = {pmin, pmax};
}

bool Mesh::IsTriangleMasked(uint32_t tri) const {
    PCH_CHECK(tri / 32 < masked_triangles.size());
    return (masked_triangles[tri / 32] & (1u << (tri % 32))) != 0;
}
void Mesh::MaskTriangle(uint32_t tri) { PCH_CHECK(tri / 32 < masked_triangles.size()); masked_triangles[tri / 32] |= (1u << (tri % 32)); }
void Mesh::UnmaskTriangle(uint32_t tri) { PCH_CHECK(tri / 32 < masked_triangles.size()); masked_triangles[tri / 32] &= ~(1u << (tri % 32)); }
void Mesh::ToggleMaskTriangle(uint32_t tri) { PCH_CHECK(tri / 32 < masked_triangles.size()); masked_triangles[tri / 32] ^= (1u << (tri % 32)); }

void ThrowPcError(pc_ctx* ctx, int code) {
    const std::string msg = pc_last_error(ctx);
    if (code == PC_ERR_INVALID) throw std::logic_error(msg);            // the reference's CHECKs
    throw std::runtime_error(msg);
}

DeviceContext::~DeviceContext() {
    if (ctx) pc_destroy(ctx);
}

static int EnvDevice() {
    const char* e = std::getenv("POLYCHASE_DEVICE");
    return e ? std::atoi(e) : 0;
}

// The solver context (mesh, ray casts, PnP, BA) is process wide and small; analyze passes create
// their own context sized to the video.
std::shared_ptr<DeviceContext> AcquireDeviceContext(int max_width, int max_height, int max_features) {
    static std::mutex g_mtx;
    static std::weak_ptr<DeviceContext> g_solver;
    const bool solver = max_width <= 0;
    std::lock_guard<std::mutex> lk(g_mtx);
    if (solver) {
        if (auto sp = g_solver.lock()) return sp;
    }
    auto dc = std::make_shared<DeviceContext>();
    pc_limits lim{};
    lim.device = EnvDevice();
    lim.max_width = solver ? 64 : max_width;
    lim.max_height = solver ? 64 : max_height;
    lim.max_features = solver ? 1024 : max_features;
    const int rc = pc_create(&lim, &dc->ctx);
    if (rc != PC_OK) throw std::runtime_error(pc_last_error(nullptr));
    if (solver) g_solver = dc;
    return dc;
}

std::shared_ptr<DeviceContext> AcquireInteractiveContext() {
    static std::mutex g_mtx;
    static std::weak_ptr<DeviceContext> g_interactive;
    std::lock_guard<std::mutex> lk(g_mtx);
    if (auto sp = g_interactive.lock()) return sp;
    auto dc = std::make_shared<DeviceContext>();
    pc_limits lim{};
    lim.device = EnvDevice();
    lim.max_width = 64;
    lim.max_height = 64;
    lim.max_features = 1024;
    lim.ring_frames = 10;
    lim.pipeline_depth = 1;
    const int rc = pc_create(&lim, &dc->ctx);
    if (rc != PC_OK) throw std::runtime_error(pc_last_error(nullptr));
    g_interactive = dc;
    return dc;
}

static std::atomic<uint64_t> g_mesh_epoch{1};

AcceleratedMesh::AcceleratedMesh(std::vector<float> vertices, std::vector<uint32_t> triangles,
                                 std::vector<uint32_t> masked_triangles)
    : mesh_(std::move(vertices), std::move(triangles), std::move(masked_triangles)), epoch_(g_mesh_epoch++) {
    for (uint32_t idx : mesh_.triangles) PCH_CHECK(idx < mesh_.NumVertices());   // geometry.h:98
}

void AcceleratedMesh::Bind(DeviceContext& dc) const {
    if (dc.mesh_epoch == epoch_) return;
    const int rc = pc_mesh_set(dc.ctx, mesh_.vertices.data(), (int)mesh_.NumVertices(), mesh_.triangles.data(),
                               (int)mesh_.NumTriangles(), mesh_.masked_triangles.data(),
                               (int)mesh_.masked_triangles.size());
    if (rc != PC_OK) ThrowPcError(dc.ctx, rc);
    dc.mesh_epoch = epoch_;
}

Mesh& AcceleratedMesh::InnerMut() {
    epoch_ = g_mesh_epoch++;
    return mesh_;
}

std::optional<RayHit> AcceleratedMesh::RayCast(const SceneTransformations& scene, Vec2 pos, bool check_mask) const {
    auto dc = AcquireInteractiveContext();
    std::lock_guard<std::mutex> lk(dc->mtx);
    Bind(*dc);
    CameraState cs{scene.intrinsics, Pose::FromRt(scene.view_matrix)};
    const pc_camera_state cam = ToAbi(cs);
    uint8_t hit = 0;
    float p[3], uv[2], t = 0;
    uint32_t prim = 0;
    const int rc = pc_ray_cast(dc->ctx, scene.model_matrix.data(), &cam, pos.data(), 1, check_mask ? 1 : 0, &hit, p,
                               &prim, uv, &t);
    if (rc != PC_OK) ThrowPcError(dc->ctx, rc);
    if (!hit) return std::nullopt;
    RayHit h;
    h.pos = {p[0], p[1], p[2]};
    h.barycentric_coordinate = {uv[0], uv[1]};
    h.t = t;
    h.primitive_id = prim;
    // geometric normal of the hit triangle, normalised (ray_casting.cc:113-116)
    const uint32_t* tri = &mesh_.triangles[3 * prim];
    const float* a = &mesh_.vertices[3 * tri[0]];
    const float* b = &mesh_.vertices[3 * tri[1]];
    const float* c = &mesh_.vertices[3 * tri[2]];
    const float e1[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, e2[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
    float n[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
    const float len = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    if (len > 0) { n[0] /= len; n[1] /= len; n[2] /= len; }
    h.normal = {n[0], n[1], n[2]};
    return h;
}

pc_camera_state ToAbi(const CameraState& s) {
    pc_camera_state o{};
    o.fx = s.intrinsics.fx; o.fy = s.intrinsics.fy; o.cx = s.intrinsics.cx; o.cy = s.intrinsics.cy;
    o.aspect_ratio = s.intrinsics.aspect_ratio; o.width = s.intrinsics.width; o.height = s.intrinsics.height;
    o.convention = s.intrinsics.convention == CameraConvention::OpenCV ? 1.f : 0.f;
    for (int k = 0; k < 4; k++) o.q[k] = s.pose.q[k];
    for (int k = 0; k < 3; k++) o.t[k] = s.pose.t[k];
    o.filled = 1.f;
    return o;
}

CameraState FromAbi(const pc_camera_state& s) {
    CameraState o;
    o.intrinsics.fx = s.fx; o.intrinsics.fy = s.fy; o.intrinsics.cx = s.cx; o.intrinsics.cy = s.cy;
    o.intrinsics.aspect_ratio = s.aspect_ratio; o.intrinsics.width = s.width; o.intrinsics.height = s.height;
    o.intrinsics.convention = s.convention != 0.f ? CameraConvention::OpenCV : CameraConvention::OpenGL;
    for (int k = 0; k < 4; k++) o.pose.q[k] = s.q[k];
    for (int k = 0; k < 3; k++) o.pose.t[k] = s.t[k];
    return o;
}

pc_bundle_opts ToAbi(const BundleOptions& o) {
    pc_bundle_opts b{};
    b.max_iterations = o.max_iterations;
    b.max_allowed_parallelism = o.max_allowed_parallelism;
    b.loss_type = static_cast<int>(o.loss_type);
    b.loss_scale = o.loss_scale; b.gradient_tol = o.gradient_tol; b.step_tol = o.step_tol;
    b.initial_lambda = o.initial_lambda; b.min_lambda = o.min_lambda; b.max_lambda = o.max_lambda;
    b.verbose = o.verbose;
    return b;
}

BundleStats FromAbi(const pc_bundle_stats& s) {
    BundleStats o;
    o.iterations = s.iterations; o.initial_cost = s.initial_cost; o.cost = s.cost; o.lambda = s.lambda;
    o.invalid_steps = s.invalid_steps; o.step_norm = s.step_norm; o.grad_norm = s.grad_norm;
    return o;
}

pc_gftt_opts ToAbi(const GFTTOptions& o) {
    pc_gftt_opts g{};
    g.quality_level = o.quality_level; g.min_distance = o.min_distance; g.block_size = o.block_size;
    g.gradient_size = o.gradient_size; g.max_corners = o.max_corners; g.use_harris = o.use_harris;
    g.harris_k = o.harris_k; g.grid_rows = o.grid_rows; g.grid_cols = o.grid_cols;
    return g;
}

pc_flow_opts ToAbi(const OpticalFlowOptions& o) {
    pc_flow_opts f{};
    f.window_size = o.window_size; f.max_level = o.max_level; f.term_max_iters = o.term_max_iters;
    f.term_epsilon = o.term_epsilon; f.min_eigen_threshold = o.min_eigen_threshold;
    return f;
}

}  // namespace pch
