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


// ---- fused analyze -> track chain ------------------------------------------------------------
// TrackCameraTrajectory (tracker.cc:133-192) walks the frames in order and SolveFrame (:36-131)
// needs, for frame f, the flows a -> f of every already posed frame a plus the pose of f-1 as the
// start.  In a forward sweep that is exactly what the streaming analyzer has just produced for
// the frame it pushed, so the ray cast and the LM solve are queued on a second (high priority)
// stream right behind the frame's LK batch: flow rows, source keypoints and source poses are all
// read from HBM where earlier launches left them, and the solved pose lands in a device ring that
// the next frames read.  The host only ever sees the results (pc_analyze_pop).
void free_track_chain(TrackChain* t) {
    if (!t) return;
    cudaFree(t->d_cams); cudaFree(t->d_X); cudaFree(t->d_x); cudaFree(t->d_valid);
    if (t->join) cudaEventDestroy(t->join);
    if (t->stream) cudaStreamDestroy(t->stream);
    delete t;
}

static inline int cam_slot(int32_t frame_id) { return ((frame_id % kCamRing) + kCamRing) % kCamRing; }

static bool chain_knows(const TrackChain* t, int32_t frame_id) {
    const int s = cam_slot(frame_id);
    return t->cam_known[s] && t->cam_frame[s] == frame_id;
}

int track_chain_enqueue(pc_ctx* c, Stage& st, int32_t frame_id, bool is_halo, int cap) {
    st.track_state = 0;
    TrackChain* t = c->track;
    if (!t || !t->on) return PC_OK;
    const int slot = cam_slot(frame_id);
    auto seed = t->seeds.find(frame_id);
    if (seed != t->seeds.end()) {                 // known pose (TrackSequence's frame_from, tracker.cc:205-207)
        *st.trk_cam_host = seed->second;
        t->seeds.erase(seed);
        PC_CUDA(c, cudaMemcpyAsync(t->d_cams + slot, st.trk_cam_host, sizeof(pc_camera_state), cudaMemcpyHostToDevice,
                                   t->stream));
        PC_CUDA(c, cudaEventRecord(st.tracked, t->stream));
        t->cam_frame[slot] = frame_id;
        t->cam_known[slot] = true;
        st.track_state = 2;
        return PC_OK;
    }
    t->cam_frame[slot] = frame_id;
    t->cam_known[slot] = false;
    if (is_halo) return PC_OK;
    // sources: the pairs (a -> f) of this stage whose source frame is posed, ascending a like
    // FindOpticalFlowsToImage returns them
    MeshData* m = c->mesh;
    ResidentSources rs{};
    rs.cap = cap;
    memcpy(rs.model, t->model, sizeof(rs.model));
    int32_t best_init = INT32_MIN;
    for (int k = st.num_pairs - 1; k >= 0; k--) {
        if (st.to[k] != frame_id || st.from[k] >= frame_id || !chain_knows(t, st.from[k])) continue;
        const FrameSlot* a = find_slot(c, st.from[k]);
        if (!a) continue;
        ResidentSource& s = rs.s[rs.nsrc++];
        s.cam = t->d_cams + cam_slot(st.from[k]);
        s.keypoints = a->kps;
        s.indices = st.dev[k].idx;
        s.targets = st.dev[k].tgt;
        s.rows = st.dev[k].count;
        best_init = std::max(best_init, st.from[k]);
    }
    if (rs.nsrc == 0) return PC_OK;               // nothing posed to track from: the frame stays unposed
    const size_t rows = (size_t)rs.nsrc * cap;
    if (t->cap_rows < rows) {
        PC_CUDA(c, cudaStreamSynchronize(t->stream));
        cudaFree(t->d_X); cudaFree(t->d_x); cudaFree(t->d_valid);
        t->d_X = t->d_x = nullptr; t->d_valid = nullptr; t->cap_rows = 0;
        const size_t want = (size_t)4 * std::max(cap, c->lim.max_features);
        PC_CUDA(c, cudaMalloc(&t->d_X, sizeof(float) * 3 * want));
        PC_CUDA(c, cudaMalloc(&t->d_x, sizeof(float) * 2 * want));
        PC_CUDA(c, cudaMalloc(&t->d_valid, want));
        t->cap_rows = want;
    }
    PC_CUDA(c, cudaStreamWaitEvent(t->stream, st.computed, 0));
    span_begin(c, KF_RAYCAST, t->stream);
    launch_raycast_resident(mesh_view(m), rs, t->d_X, t->d_x, t->d_valid, t->stream);
    span_end(c, t->stream);
    // the start of the solve is the previous frame's pose (tracker.cc:112-119)
    PnpParams prm = make_pnp_params(&t->bo, 12.0f /* tracker.cc:123 */, t->opt_f, t->opt_pp, pc_camera_state{});
    prm.bounds = t->bounds;
    span_begin(c, KF_PNP, t->stream);
    launch_pnp_lm(t->d_X, t->d_x, nullptr, t->d_valid, (int)rows, prm, t->d_cams + cam_slot(best_init),
                  t->d_cams + slot, st.trk_result_dev, t->stream);
    span_end(c, t->stream);
    int rc = check_launch(c, "track chain", 2);
    if (rc) return rc;
    PC_CUDA(c, cudaMemcpyAsync(st.trk_result_host, st.trk_result_dev, sizeof(PnpResult), cudaMemcpyDeviceToHost, t->stream));
    PC_CUDA(c, cudaMemcpyAsync(st.trk_cam_host, t->d_cams + slot, sizeof(pc_camera_state), cudaMemcpyDeviceToHost, t->stream));
    PC_CUDA(c, cudaEventRecord(st.tracked, t->stream));
    t->cam_known[slot] = true;
    st.track_state = 1;
    return PC_OK;
}

int track_chain_collect(pc_ctx* c, Stage& st, pc_frame_result* out) {
    out->tracked = 0;
    if (st.track_state == 0) return PC_OK;
    PC_CUDA(c, cudaEventSynchronize(st.tracked));
    out->tracked = st.track_state;
    out->camera = *st.trk_cam_host;
    if (st.track_state == 2) return PC_OK;
    const PnpResult& r = *st.trk_result_host;
    out->num_matches = r.num_matches;
    if (r.status == 1) {                          // tracker.cc:160-166
        out->tracked = 0;
        if (c->track) c->track->cam_known[cam_slot(st.frame_id)] = false;
        return fail(c, PC_ERR_NOT_ENOUGH_FEATURES,
                    "Could not track to frame: " + std::to_string(st.frame_id) + ". Not enough features.");
    }
    out->inlier_ratio = r.inlier_ratio;
    out->stats = r.stats;
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
    launch_pnp_lm(dX, dx, dw, dvalid, rows, prm, m->d_cam, m->d_cam, m->d_result, st);
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

int pc_analyze_track_begin(pc_ctx* c, const float model[16], const pc_bundle_opts* bo, int opt_f, int opt_pp) {
    PC_CUDA(c, cudaSetDevice(c->device));
    if (!c->analyzing) return fail(c, PC_ERR_STATE, "no analyze pass is open");
    PC_CHECK(c, !c->any_pushed, "start the track chain before the first push");
    int rc = validate_bundle_opts(c, bo);
    if (rc) return rc;
    PC_CHECK(c, model != nullptr, "model matrix is required");
    if (!c->mesh || !c->mesh->d_nodes) return fail(c, PC_ERR_STATE, "no mesh set");
    if (!c->track) {
        TrackChain* t = new TrackChain();
        c->track = t;
        int lo = 0, hi = 0;
        PC_CUDA(c, cudaDeviceGetStreamPriorityRange(&lo, &hi));
        PC_CUDA(c, cudaStreamCreateWithPriority(&t->stream, cudaStreamNonBlocking, hi));
        PC_CUDA(c, cudaEventCreateWithFlags(&t->join, cudaEventDisableTiming));
        PC_CUDA(c, cudaMalloc(&t->d_cams, sizeof(pc_camera_state) * kCamRing));
    }
    TrackChain* t = c->track;
    memcpy(t->model, model, sizeof(t->model));
    t->bo = *bo;
    t->opt_f = opt_f;
    t->opt_pp = opt_pp;
    t->seeds.clear();
    t->have_bounds = false;
    for (int i = 0; i < kCamRing; i++) { t->cam_known[i] = false; t->cam_frame[i] = 0; }
    for (auto& st : c->stages) {
        st.track_state = 0;
        if (st.tracked) continue;
        PC_CUDA(c, cudaMalloc(&st.trk_result_dev, sizeof(PnpResult)));
        PC_CUDA(c, cudaMallocHost(&st.trk_result_host, sizeof(PnpResult)));
        PC_CUDA(c, cudaMallocHost(&st.trk_cam_host, sizeof(pc_camera_state)));
        PC_CUDA(c, cudaEventCreateWithFlags(&st.tracked, cudaEventDisableTiming));
    }
    t->on = true;
    return PC_OK;
}

int pc_analyze_track_seed(pc_ctx* c, int32_t frame_id, const pc_camera_state* cam) {
    TrackChain* t = c->track;
    if (!c->analyzing || !t || !t->on) return fail(c, PC_ERR_STATE, "no track chain is open");
    PC_CHECK(c, cam != nullptr, "camera state is required");
    PC_CHECK(c, !c->any_pushed || frame_id > c->last_pushed, "seed a frame before it is pushed");
    if (!t->have_bounds) {
        t->bounds = get_bounds(*cam);
        t->have_bounds = true;
    }
    t->seeds[frame_id] = *cam;
    t->seeds[frame_id].filled = 1.f;
    return PC_OK;
}

}  // extern "C"
