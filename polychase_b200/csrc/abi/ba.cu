// C ABI: trajectory refinement (bundle adjustment) -- problem upload, cost, normal equations and
// the sparse Levenberg-Marquardt loop (RefineTrajectory / LevMarqSparseSolve,
// /root/reference/cpp/refiner.cc:649-690, /root/reference/cpp/pnp/lev_marq.h:492-588).
//
// The LM loop runs on the device (kernels/ba_lm.cu): parameters, candidate parameters and the loop state
// stay in HBM, the host enqueues one fixed launch sequence per iteration and synchronises once per
// callback.  Edge-sharded refine: with pc_ba_set_edge_shard every rank evaluates a
// contiguous share of the edges; the per-edge normal-equation blocks and the per-edge costs are
// all-gathered (NCCL, comm.cu) once per iteration each, after which every rank assembles, factors and
// decides identically -- bit-equal to the single-GPU solve, because every gathered value was produced by
// exactly one rank and the assembly order does not change.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../kernels/ba_kernels.h"
#include "comm.h"
#include "context.h"
#include "mesh.h"

namespace pc {

struct BAData {
    BAView v{};
    std::vector<void*> allocs;
    pc_camera_state* d_cams = nullptr;        // current parameters
    pc_camera_state* d_cams_new = nullptr;    // candidate parameters of the iteration
    float* d_scalars = nullptr;               // [0] cost [1] grad_norm [2] step_norm [4] llt_ok [5] cost_new
    float* d_expected = nullptr;              // nf partial sums of the expected cost change
    BALmState* d_state = nullptr;
    BALmState* h_state = nullptr;             // pinned mirror
    pc_camera_state* h_cams = nullptr;        // pinned staging of the trajectory
    uint8_t* d_edge_mask = nullptr;
    int opt_f = 0, opt_pp = 0;
    // edge sharding
    bool sharded = false;
    int edges_per_rank = 0;                   // chunk size of the all-gathers (the arrays are padded to world * chunk)
    int n_edges_padded = 0;
};

void free_ba(BAData* b) {
    if (!b) return;
    for (void* p : b->allocs) cudaFree(p);
    cudaFreeHost(b->h_state);
    cudaFreeHost(b->h_cams);
    delete b;
}

template <typename T>
static int dev_alloc(pc_ctx* c, BAData* b, T** out, size_t n) {
    void* p = nullptr;
    PC_CUDA(c, cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    b->allocs.push_back(p);
    *out = (T*)p;
    return PC_OK;
}

template <typename T>
static int dev_upload(pc_ctx* c, BAData* b, const T** out, const T* host, size_t n) {
    T* p = nullptr;
    int rc = dev_alloc(c, b, &p, n);
    if (rc) return rc;
    if (n) PC_CUDA(c, cudaMemcpy(p, host, n * sizeof(T), cudaMemcpyHostToDevice));
    *out = p;
    return PC_OK;
}

static int upload_traj(pc_ctx* c, BAData* b, const pc_camera_state* traj, pc_camera_state* dst) {
    memcpy(b->h_cams, traj, sizeof(pc_camera_state) * b->v.nf);
    PC_CUDA(c, cudaMemcpyAsync(dst, b->h_cams, sizeof(pc_camera_state) * b->v.nf, cudaMemcpyHostToDevice, c->compute));
    return PC_OK;
}

static BAView view_with(const BAData* b, const pc_camera_state* cams) {
    BAView v = b->v;
    v.cams = cams;
    return v;
}

// refresh + cost of `cams` -> *cost_out (device).  TotalCost (lev_marq.h:773-824): refresh the per-keypoint
// intersections (cache semantics of refiner.cc:323-350), then the per-edge normalised robust cost.
static int enqueue_cost(pc_ctx* c, BAData* b, const pc_camera_state* cams, const Loss& loss, const BALmState* st, int gate,
                        float* cost_out) {
    cudaStream_t s = c->compute;
    const BAView v = view_with(b, cams);
    span_begin(c, KF_BA, s);
    launch_ba_refresh_points(v, mesh_view(c->mesh), st, gate, s);
    if (b->sharded) {
        // own edges only, then every rank receives every edge's cost (each produced by exactly one rank)
        if (v.n_edges > 0) launch_ba_cost_edges(v, loss, st, gate, s);
        int rc = comm_allgather_inplace(c, v.edge_cost, (size_t)b->edges_per_rank, s);
        if (rc) return rc;
        launch_ba_cost_sum(v, st, gate, cost_out, s);
    } else {
        launch_ba_cost(v, loss, st, gate, cost_out, s);
    }
    span_end(c, s);
    return check_launch(c, "ba cost", 3);
}

// BuildNormalEquations (lev_marq.h:653-771) of `cams` into band / jtr / diag
static int enqueue_build(pc_ctx* c, BAData* b, const pc_camera_state* cams, const Loss& loss, const BALmState* st, int gate) {
    cudaStream_t s = c->compute;
    const BAView v = view_with(b, cams);
    span_begin(c, KF_BA, s);
    launch_ba_build(v, mesh_view(c->mesh), loss, st, gate, s);
    if (b->sharded) {
        int rc = comm_allgather_inplace(c, v.edge_pair, (size_t)b->edges_per_rank * v.pair_stride, s);
        if (rc) return rc;
    }
    launch_ba_assemble(v, st, gate, s);
    span_end(c, s);
    return check_launch(c, "ba build", 3);
}

}  // namespace pc

using namespace pc;

extern "C" {

int pc_ba_load(pc_ctx* c, const pc_ba_problem* pr) {
    PC_CUDA(c, cudaSetDevice(c->device));
    PC_CHECK(c, pr != nullptr, "problem is NULL");
    if (!c->mesh || !c->mesh->d_nodes) return fail(c, PC_ERR_STATE, "no mesh set");
    PC_CHECK(c, pr->num_frames > 2, "traj.Count() > 2");                         // refiner.cc:661
    PC_CHECK(c, pr->kp_offsets && pr->num_edges >= 0 && (pr->num_edges == 0 || pr->edges), "bad arrays");
    if (c->ba) {
        PC_CUDA(c, cudaStreamSynchronize(c->compute));
        free_ba(c->ba);
        c->ba = nullptr;
    }
    BAData* b = new BAData();
    c->ba = b;
    BAView& v = b->v;
    const int nf = pr->num_frames;
    v.nf = nf;
    v.opt_f = pr->optimize_focal_length != 0;
    v.opt_pp = pr->optimize_principal_point != 0;
    v.p = (v.opt_f || v.opt_pp) ? 9 : 6;                                          // refiner.cc:229-233
    b->opt_f = v.opt_f;
    b->opt_pp = v.opt_pp;
    v.n_kps = pr->kp_offsets[nf];
    v.n_edges = pr->num_edges;
    int n_rows = 0;
    for (int e = 0; e < pr->num_edges; e++) {
        const pc_ba_edge& ed = pr->edges[e];
        PC_CHECK(c, ed.src_frame_idx >= 0 && ed.src_frame_idx < nf && ed.tgt_frame_idx >= 0 && ed.tgt_frame_idx < nf,
                 "b1 < num_blocks && b2 < num_blocks");                            // lev_marq.h:435-436
        PC_CHECK(c, ed.src_frame_idx != ed.tgt_frame_idx, "b1 != b2");            // :437
        PC_CHECK(c, abs(ed.src_frame_idx - ed.tgt_frame_idx) < kBandBlocks,
                 "edges further than 8 frames apart are outside the banded solver (flows exist for +-1,2,4,8)");
        PC_CHECK(c, ed.rows >= 0 && ed.first_row >= 0, "bad edge rows");
        n_rows = std::max(n_rows, ed.first_row + ed.rows);
    }
    v.n_rows = n_rows;
    for (int i = 0; i < 16; i++) v.M[i] = pr->model[i];
    double Md[16], Mi[16];
    for (int i = 0; i < 16; i++) Md[i] = pr->model[i];
    if (!invert4x4(Md, Mi)) return fail(c, PC_ERR_INVALID, "model matrix is singular");
    for (int i = 0; i < 16; i++) v.Minv[i] = (float)Mi[i];
    // host-side derived tables
    std::vector<int> kp_frame(v.n_kps);
    for (int f = 0; f < nf; f++) {
        PC_CHECK(c, pr->kp_offsets[f + 1] >= pr->kp_offsets[f], "kp_offsets must be non-decreasing");
        for (int g = pr->kp_offsets[f]; g < pr->kp_offsets[f + 1]; g++) kp_frame[g] = f;
    }
    std::vector<uint8_t> referenced(v.n_kps, 0);
    std::vector<float> edge_weight(pr->num_edges);
    for (int e = 0; e < pr->num_edges; e++) {
        const pc_ba_edge& ed = pr->edges[e];
        const int nk = pr->kp_offsets[ed.src_frame_idx + 1] - pr->kp_offsets[ed.src_frame_idx];
        for (int r = 0; r < ed.rows; r++) {
            const uint32_t k = pr->src_kps_indices[ed.first_row + r];
            if (k >= (uint32_t)nk) return fail(c, PC_ERR_INVALID, "check failed: src_kps_indices[kp_idx] < src_kps.size()");
            referenced[pr->kp_offsets[ed.src_frame_idx] + k] = 1;
        }
        // FrameWeight (refiner.cc:250-257) of the source frame = EdgeWeight (:598-601)
        const int dist = std::min(ed.src_frame_idx, nf - 1 - ed.src_frame_idx);
        edge_weight[e] = 1.0f / ((float)dist + 1.0f);
    }
    std::vector<int> inc_off(nf + 1, 0), inc_edges;
    {
        std::vector<std::vector<int>> inc(nf);
        for (int e = 0; e < pr->num_edges; e++) {                 // ascending edge order per frame
            inc[pr->edges[e].src_frame_idx].push_back(e);
            inc[pr->edges[e].tgt_frame_idx].push_back(e);
        }
        for (int f = 0; f < nf; f++) {
            inc_edges.insert(inc_edges.end(), inc[f].begin(), inc[f].end());
            inc_off[f + 1] = (int)inc_edges.size();
        }
    }
    int rc;
    if ((rc = dev_upload(c, b, &v.kps, pr->keypoints, (size_t)v.n_kps * 2))) return rc;
    if ((rc = dev_upload(c, b, &v.kp_frame, kp_frame.data(), kp_frame.size()))) return rc;
    if ((rc = dev_upload(c, b, &v.kp_offsets, pr->kp_offsets, (size_t)nf + 1))) return rc;
    if ((rc = dev_upload(c, b, &v.referenced, referenced.data(), referenced.size()))) return rc;
    if ((rc = dev_upload(c, b, &v.edges, pr->edges, (size_t)pr->num_edges))) return rc;
    if ((rc = dev_upload(c, b, &v.edge_weight, edge_weight.data(), edge_weight.size()))) return rc;
    if ((rc = dev_upload(c, b, &v.src_idx, pr->src_kps_indices, (size_t)n_rows))) return rc;
    if ((rc = dev_upload(c, b, &v.tgt, pr->tgt_kps, (size_t)n_rows * 2))) return rc;
    if ((rc = dev_upload(c, b, &v.inc_offsets, inc_off.data(), inc_off.size()))) return rc;
    if ((rc = dev_upload(c, b, &v.inc_edges, inc_edges.data(), inc_edges.size()))) return rc;
    if ((rc = dev_alloc(c, b, &v.cache, (size_t)v.n_kps))) return rc;
    PC_CUDA(c, cudaMemset(v.cache, 0xFF, sizeof(uint32_t) * std::max(v.n_kps, 1)));   // kInvalidIndex (refiner.cc:239)
    if ((rc = dev_alloc(c, b, &v.pts, (size_t)v.n_kps * 3))) return rc;
    if ((rc = dev_alloc(c, b, &v.pt_valid, (size_t)v.n_kps))) return rc;
    PC_CUDA(c, cudaMemset(v.pt_valid, 0, std::max(v.n_kps, 1)));
    if ((rc = dev_alloc(c, b, &b->d_cams, (size_t)nf))) return rc;
    if ((rc = dev_alloc(c, b, &b->d_cams_new, (size_t)nf))) return rc;
    v.cams = b->d_cams;
    // per-edge arrays are padded so that an all-gather of equal chunks covers them (edge-sharded refine)
    const int world = std::max(1, comm_world(c));
    b->edges_per_rank = (pr->num_edges + world - 1) / world;
    b->n_edges_padded = std::max(1, b->edges_per_rank * world);
    if ((rc = dev_alloc(c, b, &v.edge_cost, (size_t)b->n_edges_padded))) return rc;
    PC_CUDA(c, cudaMemset(v.edge_cost, 0, sizeof(float) * b->n_edges_padded));
    const int NP = 2 * v.p;
    v.pair_stride = NP * (NP + 1) / 2 + NP + 1;
    if ((rc = dev_alloc(c, b, &v.edge_pair, (size_t)b->n_edges_padded * v.pair_stride))) return rc;
    PC_CUDA(c, cudaMemset(v.edge_pair, 0, sizeof(float) * (size_t)b->n_edges_padded * v.pair_stride));
    const size_t band_n = (size_t)nf * kBandBlocks * v.p * v.p;
    if ((rc = dev_alloc(c, b, &v.band, band_n))) return rc;
    if ((rc = dev_alloc(c, b, &v.lband, band_n))) return rc;
    if ((rc = dev_alloc(c, b, &v.jtr, (size_t)nf * v.p))) return rc;
    if ((rc = dev_alloc(c, b, &v.diag, (size_t)nf * v.p))) return rc;
    if ((rc = dev_alloc(c, b, &v.step, (size_t)nf * v.p))) return rc;
    if ((rc = dev_alloc(c, b, &v.tmp, (size_t)nf * v.p))) return rc;
    if ((rc = dev_alloc(c, b, &b->d_scalars, (size_t)8))) return rc;
    PC_CUDA(c, cudaMemset(b->d_scalars, 0, sizeof(float) * 8));
    v.scalars = b->d_scalars;
    if ((rc = dev_alloc(c, b, &b->d_expected, (size_t)nf))) return rc;
    if ((rc = dev_alloc(c, b, &b->d_state, (size_t)1))) return rc;
    if ((rc = dev_alloc(c, b, &b->d_edge_mask, (size_t)b->n_edges_padded))) return rc;
    v.edge_mask = nullptr;
    PC_CUDA(c, cudaMallocHost(&b->h_state, sizeof(BALmState)));
    PC_CUDA(c, cudaMallocHost(&b->h_cams, sizeof(pc_camera_state) * nf));
    return PC_OK;
}

int pc_ba_set_edge_shard(pc_ctx* c, int on) {
    PC_CUDA(c, cudaSetDevice(c->device));
    BAData* b = c->ba;
    if (!b) return fail(c, PC_ERR_STATE, "no refine problem loaded");
    if (!on) {
        b->sharded = false;
        b->v.edge_mask = nullptr;
        return PC_OK;
    }
    const int world = comm_world(c), rank = comm_rank(c);
    if (world < 1) return fail(c, PC_ERR_STATE, "pc_comm_init has not been called on this context");
    if (b->edges_per_rank * world != b->n_edges_padded || b->edges_per_rank * world < b->v.n_edges)
        return fail(c, PC_ERR_STATE, "the refine problem was loaded before pc_comm_init: load it again");
    std::vector<uint8_t> mask((size_t)b->n_edges_padded, 0);
    for (int e = rank * b->edges_per_rank; e < std::min((rank + 1) * b->edges_per_rank, b->v.n_edges); e++) mask[e] = 1;
    PC_CUDA(c, cudaMemcpy(b->d_edge_mask, mask.data(), mask.size(), cudaMemcpyHostToDevice));
    b->v.edge_mask = b->d_edge_mask;
    b->sharded = world > 1;
    if (!b->sharded) b->v.edge_mask = nullptr;
    return PC_OK;
}

int pc_ba_read_cache(pc_ctx* c, uint32_t* out, int cap) {
    PC_CUDA(c, cudaSetDevice(c->device));
    BAData* b = c->ba;
    if (!b) return fail(c, PC_ERR_STATE, "no refine problem loaded");
    if (cap < b->v.n_kps) return fail(c, PC_ERR_CAPACITY, "cache output buffer too small");
    PC_CUDA(c, cudaStreamSynchronize(c->compute));
    PC_CUDA(c, cudaMemcpy(out, b->v.cache, sizeof(uint32_t) * b->v.n_kps, cudaMemcpyDeviceToHost));
    return PC_OK;
}

int pc_ba_cost(pc_ctx* c, const pc_camera_state* traj, const pc_bundle_opts* bo, float* cost_out) {
    PC_CUDA(c, cudaSetDevice(c->device));
    BAData* b = c->ba;
    if (!b) return fail(c, PC_ERR_STATE, "no refine problem loaded");
    int rc = validate_bundle_opts(c, bo);
    if (rc) return rc;
    PC_CHECK(c, traj && cost_out, "bad arguments");
    const Loss loss = make_loss(bo->loss_type, bo->loss_scale);
    if ((rc = upload_traj(c, b, traj, b->d_cams))) return rc;
    if ((rc = enqueue_cost(c, b, b->d_cams, loss, nullptr, GATE_NONE, b->d_scalars))) return rc;
    PC_CUDA(c, cudaMemcpyAsync(cost_out, b->d_scalars, sizeof(float), cudaMemcpyDeviceToHost, c->compute));
    PC_CUDA(c, cudaStreamSynchronize(c->compute));
    return PC_OK;
}

int pc_ba_normal_equations(pc_ctx* c, const pc_camera_state* traj, const pc_bundle_opts* bo, float* JtJ_blocks,
                           float* Jtr) {
    PC_CUDA(c, cudaSetDevice(c->device));
    BAData* b = c->ba;
    if (!b) return fail(c, PC_ERR_STATE, "no refine problem loaded");
    int rc = validate_bundle_opts(c, bo);
    if (rc) return rc;
    PC_CHECK(c, traj != nullptr, "bad arguments");
    const Loss loss = make_loss(bo->loss_type, bo->loss_scale);
    if ((rc = upload_traj(c, b, traj, b->d_cams))) return rc;
    if ((rc = enqueue_build(c, b, b->d_cams, loss, nullptr, GATE_NONE))) return rc;
    const BAView& v = b->v;
    PC_CUDA(c, cudaStreamSynchronize(c->compute));
    if (JtJ_blocks)
        PC_CUDA(c, cudaMemcpy(JtJ_blocks, v.band, sizeof(float) * (size_t)v.nf * kBandBlocks * v.p * v.p, cudaMemcpyDeviceToHost));
    if (Jtr) PC_CUDA(c, cudaMemcpy(Jtr, v.jtr, sizeof(float) * (size_t)v.nf * v.p, cudaMemcpyDeviceToHost));
    return PC_OK;
}

// Solves the loaded normal equations (the last pc_ba_normal_equations) with damping lambda: step = -(A_damped)^-1 Jtr
// (ComputeStep, lev_marq.h:826-841).  Test / profiling entry point of K14; returns PC_ERR_STATE if the
// factorisation hits a non-positive pivot.
int pc_ba_solve_step(pc_ctx* c, float lambda, float* step_out, float* step_norm_out) {
    PC_CUDA(c, cudaSetDevice(c->device));
    BAData* b = c->ba;
    if (!b) return fail(c, PC_ERR_STATE, "no refine problem loaded");
    cudaStream_t s = c->compute;
    span_begin(c, KF_BA, s);
    launch_ba_solve(b->v, nullptr, lambda, s);
    span_end(c, s);
    int rc = check_launch(c, "ba solve", 1);
    if (rc) return rc;
    float sc[5];
    PC_CUDA(c, cudaMemcpyAsync(sc, b->d_scalars, sizeof(sc), cudaMemcpyDeviceToHost, s));
    PC_CUDA(c, cudaStreamSynchronize(s));
    if (sc[4] == 0.f) return fail(c, PC_ERR_STATE, "the damped normal equations are not positive definite");
    if (step_norm_out) *step_norm_out = sc[2];
    if (step_out) PC_CUDA(c, cudaMemcpy(step_out, b->v.step, sizeof(float) * (size_t)b->v.nf * b->v.p, cudaMemcpyDeviceToHost));
    return PC_OK;
}

int pc_ba_solve(pc_ctx* c, const pc_bundle_opts* bo, pc_camera_state* traj, pc_bundle_stats* stats_out,
                pc_ba_iter_cb cb, void* user) {
    PC_CUDA(c, cudaSetDevice(c->device));
    BAData* b = c->ba;
    if (!b) return fail(c, PC_ERR_STATE, "no refine problem loaded");
    int rc = validate_bundle_opts(c, bo);
    if (rc) return rc;
    PC_CHECK(c, traj != nullptr, "bad arguments");
    const BAView& v = b->v;
    const int nf = v.nf;
    for (int f = 0; f < nf; f++)
        if (traj[f].filled == 0.f) return fail(c, PC_ERR_INVALID, "check failed: traj.IsFrameFilled(frame)");   // refiner.cc:662-665
    const Loss loss = make_loss(bo->loss_type, bo->loss_scale);
    const Bounds bounds = get_bounds(traj[0]);                                      // refiner.cc:687
    cudaStream_t s = c->compute;

    // LevMarqSparseSolver::Solve (lev_marq.h:492-588): stats.cost = TotalCost(params); then the device loop
    if ((rc = upload_traj(c, b, traj, b->d_cams))) return rc;
    if ((rc = enqueue_cost(c, b, b->d_cams, loss, nullptr, GATE_NONE, b->d_scalars))) return rc;
    float cost0 = 0.f;
    PC_CUDA(c, cudaMemcpyAsync(&cost0, b->d_scalars, sizeof(float), cudaMemcpyDeviceToHost, s));
    PC_CUDA(c, cudaStreamSynchronize(s));
    BALmState& hs = *b->h_state;
    memset(&hs, 0, sizeof(hs));
    hs.gradient_tol = bo->gradient_tol;
    hs.step_tol = bo->step_tol;
    hs.min_lambda = bo->min_lambda;
    hs.max_lambda = bo->max_lambda;
    hs.max_iterations = bo->max_iterations;
    hs.cost = cost0;
    hs.initial_cost = cost0;
    hs.lambda = bo->initial_lambda;
    hs.v = 2.0f;
    hs.rebuild = 1;
    hs.grad_norm = -1.f;
    hs.step_norm = -1.f;
    hs.done = bo->max_iterations == 0 ? 1 : 0;
    PC_CUDA(c, cudaMemcpyAsync(b->d_state, &hs, sizeof(hs), cudaMemcpyHostToDevice, s));
    PC_CUDA(c, cudaStreamSynchronize(s));          // hs is reused as the read-back buffer below

    // Without a callback the loop state is looked at every kChunk iterations (iterations enqueued after the
    // loop has ended are launches that return at once); with one, after every iteration.
    const uint64_t kChunk = cb != nullptr ? 1 : 4;
    bool stopped_by_callback = false;
    pc_bundle_stats last_snap{};
    bool done = hs.done != 0;
    for (uint64_t it = 0; !done && it < bo->max_iterations + 1; it += kChunk) {
        for (uint64_t k = 0; k < kChunk; k++) {
            if ((rc = enqueue_build(c, b, b->d_cams, loss, b->d_state, GATE_BUILD))) return rc;
            span_begin(c, KF_BA, s);
            launch_ba_solve(v, b->d_state, 0.f, s);                                  // ComputeStep
            launch_ba_step(v, b->d_state, b->d_cams, b->d_cams_new, bounds, b->d_expected, s);
            span_end(c, s);
            if ((rc = check_launch(c, "ba solve", 2))) return rc;
            if ((rc = enqueue_cost(c, b, b->d_cams_new, loss, b->d_state, GATE_EVAL, b->d_scalars + 5))) return rc;
            span_begin(c, KF_BA, s);
            launch_ba_decide(v, b->d_state, b->d_cams, b->d_cams_new, b->d_expected, b->d_scalars + 5, s);
            span_end(c, s);
            if ((rc = check_launch(c, "ba decide", 1))) return rc;
        }
        PC_CUDA(c, cudaMemcpyAsync(&hs, b->d_state, sizeof(hs), cudaMemcpyDeviceToHost, s));
        PC_CUDA(c, cudaStreamSynchronize(s));
        done = hs.done != 0;
        if (cb != nullptr && hs.snap_valid) {                                         // lev_marq.h:576-580
            last_snap = hs.snap;
            if (!cb(&hs.snap, user)) {
                stopped_by_callback = true;
                break;
            }
        }
    }
    pc_bundle_stats stats{};
    if (stopped_by_callback) {
        stats = last_snap;                          // the `break` comes before ++iterations
    } else {
        stats.iterations = hs.iterations;
        stats.initial_cost = hs.initial_cost;
        stats.cost = hs.cost;
        stats.lambda = hs.lambda;
        stats.invalid_steps = hs.invalid_steps;
        stats.step_norm = hs.step_norm;
        stats.grad_norm = hs.grad_norm;
    }
    if (cb != nullptr) cb(&stats, user);
    PC_CUDA(c, cudaMemcpyAsync(b->h_cams, b->d_cams, sizeof(pc_camera_state) * nf, cudaMemcpyDeviceToHost, s));
    PC_CUDA(c, cudaStreamSynchronize(s));
    memcpy(traj, b->h_cams, sizeof(pc_camera_state) * nf);
    if (stats_out) *stats_out = stats;
    return PC_OK;
}

}  // extern "C"
