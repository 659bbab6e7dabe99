// C ABI: mesh (BVH), ray casting, PnP and per-frame tracking.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <vector>

#include "../kernels/track_kernels.h"
#include "context.h"
#include "mesh.h"

namespace pc {

// ---- host BVH build: top-down median split on the longest centroid axis, leaves <= 4 -----
namespace {

struct BuildTri {
    float bmin[3], bmax[3], c[3];
    int prim;
};

struct HostNode {
    float bmin[3];
    int first;       // leaf: first triangle; inner: index of the left child (right = left + 1)
    float bmax[3];
    int count;       // > 0 leaf
};

void build_recursive(std::vector<HostNode>& nodes, std::vector<BuildTri>& tris, int node, int lo, int hi) {
    HostNode& n = nodes[node];
    float cmin[3] = {INFINITY, INFINITY, INFINITY}, cmax[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int k = 0; k < 3; k++) { n.bmin[k] = INFINITY; n.bmax[k] = -INFINITY; }
    for (int i = lo; i < hi; i++)
        for (int k = 0; k < 3; k++) {
            n.bmin[k] = std::min(n.bmin[k], tris[i].bmin[k]);
            n.bmax[k] = std::max(n.bmax[k], tris[i].bmax[k]);
            cmin[k] = std::min(cmin[k], tris[i].c[k]);
            cmax[k] = std::max(cmax[k], tris[i].c[k]);
        }
    const int cnt = hi - lo;
    int axis = 0;
    if (cmax[1] - cmin[1] > cmax[axis] - cmin[axis]) axis = 1;
    if (cmax[2] - cmin[2] > cmax[axis] - cmin[axis]) axis = 2;
    if (cnt <= 4 || !(cmax[axis] - cmin[axis] > 0.f)) {
        if (cnt <= 8 || !(cmax[axis] - cmin[axis] > 0.f)) {   // degenerate clusters stay one (possibly fat) leaf
            n.first = lo;
            n.count = cnt;
            return;
        }
    }
    const int mid = lo + cnt / 2;
    std::nth_element(tris.begin() + lo, tris.begin() + mid, tris.begin() + hi,
                     [axis](const BuildTri& a, const BuildTri& b) { return a.c[axis] < b.c[axis]; });
    const int left = (int)nodes.size();
    nodes.push_back(HostNode{});
    nodes.push_back(HostNode{});
    nodes[node].first = left;
    nodes[node].count = 0;
    build_recursive(nodes, tris, left, lo, mid);
    build_recursive(nodes, tris, left + 1, mid, hi);
}

}  // namespace

void free_mesh(MeshData* m) {
    if (!m) return;
    cudaFree(m->d_nodes); cudaFree(m->d_tris4); cudaFree(m->d_verts); cudaFree(m->d_tris); cudaFree(m->d_mask);
    cudaFree(m->d_srcs); cudaFree(m->d_X); cudaFree(m->d_x); cudaFree(m->d_w); cudaFree(m->d_valid);
    cudaFree(m->d_kps); cudaFree(m->d_idx); cudaFree(m->d_tgt); cudaFree(m->d_cam); cudaFree(m->d_result);
    cudaFree(m->d_prim); cudaFree(m->d_uv); cudaFree(m->d_t); cudaFree(m->d_pos);
    delete m;
}

MeshView mesh_view(const MeshData* m) {
    MeshView v;
    v.bvh.nodes = m->d_nodes;
    v.bvh.tris = m->d_tris4;
    v.bvh.num_nodes = m->num_nodes;
    v.verts = m->d_verts;
    v.tris = m->d_tris;
    v.mask = m->d_mask;
    v.nv = m->nv;
    v.nt = m->nt;
    return v;
}

template <typename T>
static int ensure(pc_ctx* c, T*& p, size_t& cap, size_t need) {
    if (cap >= need) return PC_OK;
    cudaFree(p);
    p = nullptr;
    cap = 0;
    need = std::max<size_t>(need * 3 / 2, 256);
    PC_CUDA(c, cudaMalloc(&p, need * sizeof(T)));
    cap = need;
    return PC_OK;
}

int ensure_match_capacity(pc_ctx* c, MeshData* m, size_t rows) {
    int rc;
    if ((rc = ensure(c, m->d_X, m->cap_X, rows * 3))) return rc;
    if ((rc = ensure(c, m->d_x, m->cap_x, rows * 2))) return rc;
    if ((rc = ensure(c, m->d_w, m->cap_w, rows))) return rc;
    if ((rc = ensure(c, m->d_valid, m->cap_valid, rows))) return rc;
    if ((rc = ensure(c, m->d_prim, m->cap_prim, rows))) return rc;
    if ((rc = ensure(c, m->d_uv, m->cap_uv, rows * 2))) return rc;
    if ((rc = ensure(c, m->d_t, m->cap_t, rows))) return rc;
    if ((rc = ensure(c, m->d_pos, m->cap_pos, rows * 3))) return rc;
    return PC_OK;
}

// inverse of a general 4x4 (row-major) in double; false if singular
bool invert4x4(const double m[16], double inv[16]) {
    double a[4][8];
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) { a[r][c] = m[r * 4 + c]; a[r][4 + c] = r == c ? 1.0 : 0.0; }
    for (int col = 0; col < 4; col++) {
        int piv = col;
        for (int r = col + 1; r < 4; r++)
            if (fabs(a[r][col]) > fabs(a[piv][col])) piv = r;
        if (a[piv][col] == 0.0) return false;
        if (piv != col)
            for (int k = 0; k < 8; k++) std::swap(a[piv][k], a[col][k]);
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
// here the inverse is formed in double and rounded once.
bool make_ray_source(const pc_camera_state& cam, const float model[16], RaySource& s) {
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

PnpParams make_pnp_params(const pc_bundle_opts* o, float max_inlier_error, int opt_f, int opt_pp,
                          const pc_camera_state& cam) {
    PnpParams p;
    p.max_iterations = o->max_iterations;
    p.loss_type = o->loss_type;
    p.loss_scale = o->loss_scale;
    p.gradient_tol = o->gradient_tol;
    p.step_tol = o->step_tol;
    p.initial_lambda = o->initial_lambda;
    p.min_lambda = o->min_lambda;
    p.max_lambda = o->max_lambda;
    p.max_inlier_error = max_inlier_error;
    p.opt_f = opt_f;
    p.opt_pp = opt_pp;
    p.bounds = get_bounds(cam);     // solvers.cc:19-21: bounds from the initial intrinsics
    return p;
}

int validate_bundle_opts(pc_ctx* c, const pc_bundle_opts* o) {
    PC_CHECK(c, o != nullptr, "bundle options are required");
    if (o->loss_type < 0 || o->loss_type > 2)
        return fail(c, PC_ERR_INVALID, "Unknown loss type: " + std::to_string(o->loss_type));   // solvers.cc:67-70
    return PC_OK;
}

}  // namespace pc

using namespace pc;

extern "C" {

int pc_mesh_set(pc_ctx* c, const float* verts, int nv, const uint32_t* tris, int nt, const uint32_t* mask_bits,
                int n_mask_words) {
    PC_CUDA(c, cudaSetDevice(c->device));
    PC_CHECK(c, verts && tris && nv > 0 && nt > 0, "mesh needs vertices and triangles");
    for (int i = 0; i < nt * 3; i++)
        if (tris[i] >= (uint32_t)nv) return fail(c, PC_ERR_INVALID, "check failed: idx < vertices.rows()");   // geometry.h:98
    const int mask_ints = (nt + 31) / 32;
    const int mask_padded = mask_ints + (4 - mask_ints % 4) % 4;     // geometry.h:63-65
    if (mask_bits && n_mask_words > 0 && n_mask_words < mask_padded)
        return fail(c, PC_ERR_INVALID, "check failed: masked_triangles.rows() >= mask_num_ints_padded");     // :72
    if (c->mesh) { free_mesh(c->mesh); c->mesh = nullptr; }
    MeshData* m = new MeshData();
    c->mesh = m;
    m->nv = nv;
    m->nt = nt;
    std::vector<BuildTri> bt(nt);
    for (int i = 0; i < nt; i++) {
        BuildTri& t = bt[i];
        t.prim = i;
        for (int k = 0; k < 3; k++) {
            const float a = verts[3 * tris[3 * i] + k], b = verts[3 * tris[3 * i + 1] + k], d = verts[3 * tris[3 * i + 2] + k];
            t.bmin[k] = std::min(a, std::min(b, d));
            t.bmax[k] = std::max(a, std::max(b, d));
            t.c[k] = (a + b + d) * (1.f / 3.f);
        }
    }
    std::vector<HostNode> nodes;
    nodes.reserve(2 * nt);
    nodes.push_back(HostNode{});
    build_recursive(nodes, bt, 0, 0, nt);
    std::vector<float4> n4(nodes.size() * 2), t4((size_t)nt * 3);
    for (size_t i = 0; i < nodes.size(); i++) {
        n4[2 * i] = make_float4(nodes[i].bmin[0], nodes[i].bmin[1], nodes[i].bmin[2], 0.f);
        n4[2 * i + 1] = make_float4(nodes[i].bmax[0], nodes[i].bmax[1], nodes[i].bmax[2], 0.f);
        memcpy(&n4[2 * i].w, &nodes[i].first, 4);
        memcpy(&n4[2 * i + 1].w, &nodes[i].count, 4);
    }
    for (int i = 0; i < nt; i++) {
        const int p = bt[i].prim;
        for (int v = 0; v < 3; v++) {
            const uint32_t vi = tris[3 * p + v];
            t4[3 * i + v] = make_float4(verts[3 * vi], verts[3 * vi + 1], verts[3 * vi + 2], 0.f);
        }
        memcpy(&t4[3 * i].w, &p, 4);
    }
    m->num_nodes = (int)nodes.size();
    PC_CUDA(c, cudaMalloc(&m->d_nodes, n4.size() * sizeof(float4)));
    PC_CUDA(c, cudaMalloc(&m->d_tris4, t4.size() * sizeof(float4)));
    PC_CUDA(c, cudaMalloc(&m->d_verts, sizeof(float) * 3 * nv));
    PC_CUDA(c, cudaMalloc(&m->d_tris, sizeof(uint32_t) * 3 * nt));
    PC_CUDA(c, cudaMemcpy(m->d_nodes, n4.data(), n4.size() * sizeof(float4), cudaMemcpyHostToDevice));
    PC_CUDA(c, cudaMemcpy(m->d_tris4, t4.data(), t4.size() * sizeof(float4), cudaMemcpyHostToDevice));
    PC_CUDA(c, cudaMemcpy(m->d_verts, verts, sizeof(float) * 3 * nv, cudaMemcpyHostToDevice));
    PC_CUDA(c, cudaMemcpy(m->d_tris, tris, sizeof(uint32_t) * 3 * nt, cudaMemcpyHostToDevice));
    if (mask_bits && n_mask_words > 0) {
        PC_CUDA(c, cudaMalloc(&m->d_mask, sizeof(uint32_t) * n_mask_words));
        PC_CUDA(c, cudaMemcpy(m->d_mask, mask_bits, sizeof(uint32_t) * n_mask_words, cudaMemcpyHostToDevice));
    }
    // geometry.h:74-95 bounding box
    for (int k = 0; k < 3; k++) { m->bbox_min[k] = FLT_MAX; m->bbox_max[k] = -FLT_MAX; }
    for (int i = 0; i < nv; i++)
        for (int k = 0; k < 3; k++) {
            m->bbox_min[k] = std::min(m->bbox_min[k], verts[3 * i + k]);
            m->bbox_max[k] = std::max(m->bbox_max[k], verts[3 * i + k]);
        }
    PC_CUDA(c, cudaMalloc(&m->d_srcs, sizeof(RaySource) * 16));
    PC_CUDA(c, cudaMalloc(&m->d_cam, sizeof(pc_camera_state)));
    PC_CUDA(c, cudaMalloc(&m->d_result, sizeof(PnpResult)));
    return PC_OK;
}

int pc_ray_cast(pc_ctx* c, const float model[16], const pc_camera_state* cam, const float* pos, int n, int check_mask,
                uint8_t* hit_out, float* pos_out, uint32_t* prim_out, float* uv_out, float* t_out) {
    PC_CUDA(c, cudaSetDevice(c->device));
    MeshData* m = c->mesh;
    if (!m) return fail(c, PC_ERR_STATE, "no mesh set");
    PC_CHECK(c, model && cam && (n == 0 || pos), "bad arguments");
    if (n == 0) return PC_OK;
    int rc = ensure_match_capacity(c, m, n);
    if (rc) return rc;
    if ((rc = ensure(c, m->d_kps, m->cap_kps, (size_t)n * 2))) return rc;
    RaySource s{};
    if (!make_ray_source(*cam, model, s)) return fail(c, PC_ERR_INVALID, "view*model is singular");
    s.keypoints = m->d_kps;
    s.indices = nullptr;
    s.targets = nullptr;
    s.first = 0;
    s.rows = n;
    cudaStream_t st = c->compute;
    PC_CUDA(c, cudaMemcpyAsync(m->d_kps, pos, sizeof(float) * 2 * n, cudaMemcpyHostToDevice, st));
    PC_CUDA(c, cudaMemcpyAsync(m->d_srcs, &s, sizeof(s), cudaMemcpyHostToDevice, st));
    span_begin(c, KF_RAYCAST, st);
    launch_raycast_sources(mesh_view(m), m->d_srcs, 1, n, check_mask, model, nullptr, nullptr, m->d_valid, m->d_prim,
                           m->d_uv, m->d_t, m->d_pos, st);
    span_end(c, st);
    rc = check_launch(c, "raycast", 1);
    if (rc) return rc;
    if (hit_out) PC_CUDA(c, cudaMemcpyAsync(hit_out, m->d_valid, n, cudaMemcpyDeviceToHost, st));
    if (pos_out) PC_CUDA(c, cudaMemcpyAsync(pos_out, m->d_pos, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, st));
    if (prim_out) PC_CUDA(c, cudaMemcpyAsync(prim_out, m->d_prim, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, st));
    if (uv_out) PC_CUDA(c, cudaMemcpyAsync(uv_out, m->d_uv, sizeof(float) * 2 * n, cudaMemcpyDeviceToHost, st));
    if (t_out) PC_CUDA(c, cudaMemcpyAsync(t_out, m->d_t, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
    PC_CUDA(c, cudaStreamSynchronize(st));
    return PC_OK;
}

static int run_pnp(pc_ctx* c, MeshData* m, const float* dX, const float* dx, const float* dw, const uint8_t* dvalid,
                   int rows, const pc_bundle_opts* bo, float max_inlier_error, int opt_f, int opt_pp,
                   pc_camera_state* cam, pc_bundle_stats* stats, float* inlier_ratio, int* num_matches) {
    cudaStream_t st = c->compute;
    const PnpParams prm = make_pnp_params(bo, max_inlier_error, opt_f, opt_pp, *cam);
    PC_CUDA(c, cudaMemcpyAsync(m->d_cam, cam, sizeof(*cam), cudaMemcpyHostToDevice, st));
    span_begin(c, KF_PNP, st);
    launch_pnp_lm(dX, dx, dw, dvalid, rows, prm, m->d_cam, m->d_result, st);
    span_end(c, st);
    int rc = check_launch(c, "pnp", 1);
    if (rc) return rc;
    PnpResult res;
    pc_camera_state out;
    PC_CUDA(c, cudaMemcpyAsync(&res, m->d_result, sizeof(res), cudaMemcpyDeviceToHost, st));
    PC_CUDA(c, cudaMemcpyAsync(&out, m->d_cam, sizeof(out), cudaMemcpyDeviceToHost, st));
    PC_CUDA(c, cudaStreamSynchronize(st));
    if (num_matches) *num_matches = res.num_matches;
    if (res.status == 1)
        return fail(c, PC_ERR_NOT_ENOUGH_FEATURES, "Not enough features (" + std::to_string(res.num_matches) + " matches)");
    *cam = out;
    if (stats) *stats = res.stats;
    if (inlier_ratio) *inlier_ratio = res.inlier_ratio;
    return PC_OK;
}

// A context without a mesh still needs the small device scratch of MeshData for pc_solve_pnp.
static int ensure_scratch(pc_ctx* c) {
    if (c->mesh) return PC_OK;
    MeshData* m = new MeshData();
    c->mesh = m;
    PC_CUDA(c, cudaMalloc(&m->d_srcs, sizeof(RaySource) * 16));
    PC_CUDA(c, cudaMalloc(&m->d_cam, sizeof(pc_camera_state)));
    PC_CUDA(c, cudaMalloc(&m->d_result, sizeof(PnpResult)));
    return PC_OK;
}

int pc_solve_pnp(pc_ctx* c, const float* X, const float* x, const float* weights, int mrows,
                 const pc_bundle_opts* bo, float max_inlier_error, int opt_f, int opt_pp, pc_camera_state* cam,
                 pc_bundle_stats* stats, float* inlier_ratio) {
    PC_CUDA(c, cudaSetDevice(c->device));
    int rc = validate_bundle_opts(c, bo);
    if (rc) return rc;
    PC_CHECK(c, X && x && cam, "bad arguments");
    PC_CHECK(c, mrows >= 3, "object_points.rows() >= 3");      // solvers.cc:55
    if ((rc = ensure_scratch(c))) return rc;
    MeshData* m = c->mesh;
    if ((rc = ensure_match_capacity(c, m, mrows))) return rc;
    cudaStream_t st = c->compute;
    PC_CUDA(c, cudaMemcpyAsync(m->d_X, X, sizeof(float) * 3 * mrows, cudaMemcpyHostToDevice, st));
    PC_CUDA(c, cudaMemcpyAsync(m->d_x, x, sizeof(float) * 2 * mrows, cudaMemcpyHostToDevice, st));
    if (weights) PC_CUDA(c, cudaMemcpyAsync(m->d_w, weights, sizeof(float) * mrows, cudaMemcpyHostToDevice, st));
    return run_pnp(c, m, m->d_X, m->d_x, weights ? m->d_w : nullptr, nullptr, mrows, bo, max_inlier_error, opt_f,
                   opt_pp, cam, stats, inlier_ratio, nullptr);
}

int pc_track_frame(pc_ctx* c, const pc_match_source* srcs, int nsrc, const float model[16],
                   const pc_camera_state* init, const pc_bundle_opts* bo, int opt_f, int opt_pp,
                   pc_camera_state* out, pc_bundle_stats* stats, float* inlier_ratio, int* num_matches) {
    PC_CUDA(c, cudaSetDevice(c->device));
    int rc = validate_bundle_opts(c, bo);
    if (rc) return rc;
    MeshData* m = c->mesh;
    if (!m || !m->d_nodes) return fail(c, PC_ERR_STATE, "no mesh set");
    PC_CHECK(c, model && init && out && nsrc >= 0 && nsrc <= 16 && (nsrc == 0 || srcs), "bad arguments");
    size_t total_rows = 0, total_kps = 0;
    for (int i = 0; i < nsrc; i++) {
        PC_CHECK(c, srcs[i].rows >= 0 && srcs[i].nk >= 0, "negative sizes");
        total_rows += srcs[i].rows;
        total_kps += srcs[i].nk;
    }
    if (num_matches) *num_matches = 0;
    if (total_rows < 3) return fail(c, PC_ERR_NOT_ENOUGH_FEATURES, "Not enough features.");   // tracker.cc:95-97
    if ((rc = ensure_match_capacity(c, m, total_rows))) return rc;
    if ((rc = ensure(c, m->d_kps, m->cap_kps, total_kps * 2))) return rc;
    if ((rc = ensure(c, m->d_idx, m->cap_idx, total_rows))) return rc;
    if ((rc = ensure(c, m->d_tgt, m->cap_tgt, total_rows * 2))) return rc;
    cudaStream_t st = c->compute;
    std::vector<RaySource> rs(nsrc);
    size_t row0 = 0, kp0 = 0;
    for (int i = 0; i < nsrc; i++) {
        const pc_match_source& S = srcs[i];
        for (int k = 0; k < S.rows; k++)
            if (S.src_kps_indices[k] >= (uint32_t)S.nk)
                return fail(c, PC_ERR_INVALID, "flow row references a keypoint outside the source frame");
        if (!make_ray_source(S.camera, model, rs[i])) return fail(c, PC_ERR_INVALID, "view*model is singular");
        rs[i].keypoints = m->d_kps + kp0 * 2;
        rs[i].indices = m->d_idx + row0;
        rs[i].targets = m->d_tgt + row0 * 2;
        rs[i].first = (int)row0;
        rs[i].rows = S.rows;
        if (S.nk) PC_CUDA(c, cudaMemcpyAsync(m->d_kps + kp0 * 2, S.keypoints, sizeof(float) * 2 * S.nk, cudaMemcpyHostToDevice, st));
        if (S.rows) {
            PC_CUDA(c, cudaMemcpyAsync(m->d_idx + row0, S.src_kps_indices, sizeof(uint32_t) * S.rows, cudaMemcpyHostToDevice, st));
            PC_CUDA(c, cudaMemcpyAsync(m->d_tgt + row0 * 2, S.tgt_kps, sizeof(float) * 2 * S.rows, cudaMemcpyHostToDevice, st));
        }
        row0 += S.rows;
        kp0 += S.nk;
    }
    PC_CUDA(c, cudaMemcpyAsync(m->d_srcs, rs.data(), sizeof(RaySource) * nsrc, cudaMemcpyHostToDevice, st));
    span_begin(c, KF_RAYCAST, st);
    launch_raycast_sources(mesh_view(m), m->d_srcs, nsrc, (int)total_rows, 1, model, m->d_X, m->d_x, m->d_valid,
                           nullptr, nullptr, nullptr, nullptr, st);
    span_end(c, st);
    rc = check_launch(c, "raycast", 1);
    if (rc) return rc;
    pc_camera_state cam = *init;
    rc = run_pnp(c, m, m->d_X, m->d_x, nullptr, m->d_valid, (int)total_rows, bo, 12.0f /* tracker.cc:123 */, opt_f,
                 opt_pp, &cam, stats, inlier_ratio, num_matches);
    if (rc) return rc;
    *out = cam;
    return PC_OK;
}

}  // extern "C"
