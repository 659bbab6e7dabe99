// C ABI: trajectory refinement (bundle adjustment) -- problem upload, cost, normal equations and
// the sparse Levenberg-Marquardt loop (RefineTrajectory / LevMarqSparseSolve,
// /root/reference/cpp/refiner.cc:649-690, /root/reference/cpp/pnp/lev_marq.h:492-588).
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../kernels/ba_kernels.h"
#include "context.h"
#include "mesh.h"

namespace pc {

struct BAData {
    BAView v{};
    std::vector<void*> allocs;
    std::vector<pc_camera_state> host_traj;
    pc_camera_state* d_cams = nullptr;
    float* d_scalars = nullptr;
    int opt_f = 0, opt_pp = 0;
};

void free_ba(BAData* b) {
    if (!b) return;
    for (void* p : b->allocs) cudaFree(p);
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

static int upload_traj(pc_ctx* c, BAData* b, const pc_camera_state* traj) {
    PC_CUDA(c, cudaMemcpyAsync(b->d_cams, traj, sizeof(pc_camera_state) * b->v.nf, cudaMemcpyHostToDevice, c->compute));
    return PC_OK;
}

// TotalCost (lev_marq.h:773-824): refresh the per-keypoint intersections (cache semantics of
// refiner.cc:323-350), then per-edge normalised robust cost.
static int ba_total_cost(pc_ctx* c, BAData* b, const pc_camera_state* traj, const Loss& loss, float* cost) {
    int rc = upload_traj(c, b, traj);
    if (rc) return rc;
    cudaStream_t st = c->compute;
    span_begin(c, KF_BA, st);
    launch_ba_refresh_points(b->v, mesh_view(c->mesh), st);
    launch_ba_cost(b->v, loss, st);
    span_end(c, st);
    rc = check_launch(c, "ba cost", 3);
    if (rc) return rc;
    PC_CUDA(c, cudaMemcpyAsync(cost, b->d_scalars, sizeof(float), cudaMemcpyDeviceToHost, st));
    PC_CUDA(c, cudaStreamSynchronize(st));
    return PC_OK;
}

static int ba_build(pc_ctx* c, BAData* b, const pc_camera_state* traj, const Loss& loss, float* grad_norm) {
    int rc = upload_traj(c, b, traj);
    if (rc) return rc;
    cudaStream_t st = c->compute;
    span_begin(c, KF_BA, st);
    launch_ba_build(b->v, mesh_view(c->mesh), loss, st);
    launch_ba_assemble(b->v, st);
    span_end(c, st);
    rc = check_launch(c, "ba build", 3);
    if (rc) return rc;
    if (grad_norm) {
        PC_CUDA(c, cudaMemcpyAsync(grad_norm, b->d_scalars + 1, sizeof(float), cudaMemcpyDeviceToHost, st));
        PC_CUDA(c, cudaStreamSynchronize(st));
    }
    return PC_OK;
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
    if (c->ba) { free_ba(c->ba); c->ba = nullptr; }
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
    for (int f = 0; f < nf; f++) {
        for (int e = 0; e < pr->num_edges; e++)
            if (pr->edges[e].src_frame_idx == f || pr->edges[e].tgt_frame_idx == f) inc_edges.push_back(e);
        inc_off[f + 1] = (int)inc_edges.size();
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
    v.cams = b->d_cams;
    if ((rc = dev_alloc(c, b, &v.edge_cost, (size_t)pr->num_edges))) return rc;
    const int NP = 2 * v.p;
    v.pair_stride = NP * (NP + 1) / 2 + NP + 1;
    if ((rc = dev_alloc(c, b, &v.edge_pair, (size_t)pr->num_edges * v.pair_stride))) return rc;
    const size_t band_n = (size_t)nf * kBandBlocks * v.p * v.p;
    if ((rc = dev_alloc(c, b, &v.band, band_n))) return rc;
    if ((rc = dev_alloc(c, b, &v.lband, band_n))) return rc;
    if ((rc = dev_alloc(c, b, &v.jtr, (size_t)nf * v.p))) return rc;
    if ((rc = dev_alloc(c, b, &v.diag, (size_t)nf * v.p))) return rc;
    if ((rc = dev_alloc(c, b, &v.step, (size_t)nf * v.p))) return rc;
    if ((rc = dev_alloc(c, b, &v.tmp, (size_t)nf * v.p))) return rc;
    if ((rc = dev_alloc(c, b, &b->d_scalars, (size_t)8))) return rc;
    v.scalars = b->d_scalars;
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
    if (!c->ba) return fail(c, PC_ERR_STATE, "no refine problem loaded");
    int rc = validate_bundle_opts(c, bo);
    if (rc) return rc;
    PC_CHECK(c, traj && cost_out, "bad arguments");
    const Loss loss = make_loss(bo->loss_type, bo->loss_scale);
    return ba_total_cost(c, c->ba, traj, loss, cost_out);
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
    rc = ba_build(c, b, traj, loss, nullptr);
    if (rc) return rc;
    const BAView& v = b->v;
    PC_CUDA(c, cudaStreamSynchronize(c->compute));
    if (JtJ_blocks)
        PC_CUDA(c, cudaMemcpy(JtJ_blocks, v.band, sizeof(float) * (size_t)v.nf * kBandBlocks * v.p * v.p, cudaMemcpyDeviceToHost));
    if (Jtr) PC_CUDA(c, cudaMemcpy(Jtr, v.jtr, sizeof(float) * (size_t)v.nf * v.p, cudaMemcpyDeviceToHost));
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
    const int nf = v.nf, p = v.p;
    for (int f = 0; f < nf; f++)
        if (traj[f].filled == 0.f) return fail(c, PC_ERR_INVALID, "check failed: traj.IsFrameFilled(frame)");   // refiner.cc:662-665
    const Loss loss = make_loss(bo->loss_type, bo->loss_scale);
    const Bounds bounds = get_bounds(traj[0]);                                      // refiner.cc:687
    cudaStream_t st = c->compute;
    std::vector<pc_camera_state> params(traj, traj + nf), params_new(nf);
    std::vector<float> step((size_t)nf * p);

    // LevMarqSparseSolver::Solve (lev_marq.h:492-588)
    pc_bundle_stats stats{};
    rc = ba_total_cost(c, b, params.data(), loss, &stats.cost);
    if (rc) return rc;
    stats.initial_cost = stats.cost;
    stats.grad_norm = -1.f;
    stats.step_norm = -1.f;
    stats.invalid_steps = 0;
    stats.lambda = bo->initial_lambda;
    float vfac = 2.0f;
    bool rebuild = true;
    for (stats.iterations = 0; stats.iterations < bo->max_iterations; ++stats.iterations) {
        if (rebuild) {
            rc = ba_build(c, b, params.data(), loss, &stats.grad_norm);
            if (rc) return rc;
            if (stats.grad_norm < bo->gradient_tol) break;
        }
        span_begin(c, KF_BA, st);
        launch_ba_solve(v, stats.lambda, st);                                       // ComputeStep
        span_end(c, st);
        rc = check_launch(c, "ba solve", 1);
        if (rc) return rc;
        float sc[5];
        PC_CUDA(c, cudaMemcpyAsync(sc, b->d_scalars, sizeof(sc), cudaMemcpyDeviceToHost, st));
        PC_CUDA(c, cudaStreamSynchronize(st));
        if (sc[4] == 0.f) {                                                          // factorisation failed
            stats.invalid_steps++;
            if (stats.lambda == bo->max_lambda) break;
            stats.lambda = std::min(bo->max_lambda, stats.lambda * vfac);
            vfac = 2 * vfac;
            rebuild = false;
            continue;
        }
        stats.step_norm = sc[2];
        if (stats.step_norm < bo->step_tol) break;
        PC_CUDA(c, cudaMemcpyAsync(step.data(), v.step, sizeof(float) * step.size(), cudaMemcpyDeviceToHost, st));
        PC_CUDA(c, cudaStreamSynchronize(st));
        // GlobalRefinementProblem::Step (refiner.cc:618-646): first and last cameras are constant
        params_new[0] = params[0];
        params_new[nf - 1] = params[nf - 1];
        for (int f = 1; f < nf - 1; f++)
            camera_step(params[f], &step[(size_t)f * p], b->opt_f, b->opt_pp, bounds, params_new[f]);
        float cost_new = 0.f;
        rc = ba_total_cost(c, b, params_new.data(), loss, &cost_new);
        if (rc) return rc;
        if (cost_new < stats.cost) {
            const float actual = cost_new - stats.cost;
            span_begin(c, KF_BA, st);
            launch_ba_expected_change(v, st);
            span_end(c, st);
            rc = check_launch(c, "ba expected change", 1);
            if (rc) return rc;
            float expected = 0.f;
            PC_CUDA(c, cudaMemcpyAsync(&expected, b->d_scalars + 3, sizeof(float), cudaMemcpyDeviceToHost, st));
            PC_CUDA(c, cudaStreamSynchronize(st));
            const float rho = actual / expected;
            if (rho > 0) {
                const float factor = (float)std::max(1.0 / 3.0, 1.0 - pow(2.0 * rho - 1.0, 3));   // Float factor
                stats.lambda = std::min(std::max(stats.lambda * factor, bo->min_lambda), bo->max_lambda);
            }
            std::swap(params, params_new);
            stats.cost = cost_new;
            vfac = 2;
            rebuild = true;
        } else {
            stats.invalid_steps++;
            if (stats.lambda == bo->max_lambda) break;
            stats.lambda = std::min(bo->max_lambda, stats.lambda * vfac);
            vfac = 2 * vfac;
            rebuild = false;
        }
        if (cb != nullptr && !cb(&stats, user)) break;
    }
    if (cb != nullptr) cb(&stats, user);
    memcpy(traj, params.data(), sizeof(pc_camera_state) * nf);
    if (stats_out) *stats_out = stats;
    return PC_OK;
}

}  // extern "C"
