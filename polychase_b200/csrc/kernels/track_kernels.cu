// K10 + K11: nearest-hit ray casting of source keypoints onto the mesh, and the robust
// Levenberg-Marquardt pose solve, both device resident.
//
// K10 replaces Embree's rtcIntersect1 behind AcceleratedMesh::RayCast
// (/root/reference/cpp/ray_casting.cc:65-121) and the per-match loop of SolveFrame
// (/root/reference/cpp/tracker.cc:64-92): a stack-based traversal of a BVH built once per
// mesh, one thread per ray, nearest hit with tnear = 0, masked triangle = miss
// (ray_casting.cc:106-108), position = (1-u-v) p1 + u p2 + v p3 (geometry.h:17-19).
//
// K11 replaces LevMarqDenseSolve<PnPProblem, Loss> (/root/reference/cpp/pnp/lev_marq.h:99-389,
// /root/reference/cpp/pnp/pnp_problem.h:52-131, /root/reference/cpp/pnp/solvers.cc:11-48):
// the whole LM loop (cost, normal equations, damping, 9x9 LLT, step, accept/reject, lambda
// schedule, termination tests) runs inside ONE kernel launch on one CTA -- the tracker is a
// sequential chain of small problems, so launch and sync latency, not bandwidth, bound it.
// Per-thread partial sums are float32; the cross-thread reduction is float64 (the reference's
// own sums are order-nondeterministic under TBB; this sits inside that band).
#include <cooperative_groups.h>

#include "common.cuh"
#include "geom.cuh"
#include "kernels.h"
#include "track_kernels.h"
#include "bvh.cuh"

namespace cg = cooperative_groups;

namespace pc {

__device__ __forceinline__ V3 barycentric_pos(const MeshView& mesh, int prim, float u, float v) {
    const uint32_t i0 = mesh.tris[3 * prim], i1 = mesh.tris[3 * prim + 1], i2 = mesh.tris[3 * prim + 2];
    const V3 p1 = v3(mesh.verts[3 * i0], mesh.verts[3 * i0 + 1], mesh.verts[3 * i0 + 2]);
    const V3 p2 = v3(mesh.verts[3 * i1], mesh.verts[3 * i1 + 1], mesh.verts[3 * i1 + 2]);
    const V3 p3 = v3(mesh.verts[3 * i2], mesh.verts[3 * i2 + 1], mesh.verts[3 * i2 + 2]);
    const float w = (float)(1.0 - (double)u - (double)v);     // geometry.h:18: (1.0 - u - v) is double
    return p1 * w + p2 * u + p3 * v;
}

__device__ __forceinline__ bool tri_masked(const MeshView& mesh, int prim) {
    return mesh.mask != nullptr && ((mesh.mask[prim >> 5] >> (prim & 31)) & 1u);
}

__global__ void __launch_bounds__(128) raycast_sources_kernel(MeshView mesh, const RaySource* __restrict__ srcs,
                                                              int nsrc, int total, int check_mask, M3 model_r,
                                                              V3 model_t, float* __restrict__ X_out,
                                                              float* __restrict__ x_out, uint8_t* __restrict__ valid,
                                                              uint32_t* __restrict__ prim_out,
                                                              float* __restrict__ uv_out, float* __restrict__ t_out,
                                                              float* __restrict__ pos_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int s = 0;
    while (s + 1 < nsrc && i >= srcs[s + 1].first) s++;
    const RaySource& S = srcs[s];
    const int j = i - S.first;
    const int kp = S.indices ? (int)S.indices[j] : j;
    const float px = S.keypoints[2 * kp], py = S.keypoints[2 * kp + 1];
    // GetRayObjectSpace (ray_casting.h:53-63): dir = inv(V*M)[3x3] * Unproject(pos)
    const V3 dc = V3{S.sgn * ((px - S.cx) / S.fx), S.sgn * ((py - S.cy) / S.fy), S.sgn};
    const V3 d = mul(S.dir_mat, dc);
    const HitRec h = bvh_nearest_hit(mesh.bvh, S.origin, d);
    bool ok = h.prim >= 0;
    if (ok && check_mask && tri_masked(mesh, h.prim)) ok = false;
    V3 pos = v3(0, 0, 0);
    if (ok) pos = barycentric_pos(mesh, h.prim, h.u, h.v);
    if (valid) valid[i] = ok ? 1 : 0;
    if (X_out) {                                  // tracker.cc:80-82: world = M3x3 * pos + Mt
        const V3 wv = ok ? mul(model_r, pos) + model_t : v3(0, 0, 0);
        X_out[3 * i] = wv.x; X_out[3 * i + 1] = wv.y; X_out[3 * i + 2] = wv.z;
    }
    if (x_out && S.targets) { x_out[2 * i] = S.targets[2 * j]; x_out[2 * i + 1] = S.targets[2 * j + 1]; }
    if (pos_out) { pos_out[3 * i] = pos.x; pos_out[3 * i + 1] = pos.y; pos_out[3 * i + 2] = pos.z; }
    if (prim_out) prim_out[i] = ok ? (uint32_t)h.prim : 0xFFFFFFFFu;
    if (uv_out) { uv_out[2 * i] = h.u; uv_out[2 * i + 1] = h.v; }
    if (t_out) t_out[i] = ok ? h.t : 0.f;
}

void launch_raycast_sources(const MeshView& mesh, const RaySource* srcs_dev, int nsrc, int total, int check_mask,
                            const float model[16], float* X_out, float* x_out, uint8_t* valid, uint32_t* prim_out,
                            float* uv_out, float* t_out, float* pos_out, cudaStream_t s) {
    if (total <= 0) return;
    M3 mr;
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) mr.m[r * 3 + c] = model[r * 4 + c];
    const V3 mt = v3(model[3], model[7], model[11]);
    raycast_sources_kernel<<<(total + 127) / 128, 128, 0, s>>>(mesh, srcs_dev, nsrc, total, check_mask, mr, mt, X_out,
                                                                x_out, valid, prim_out, uv_out, t_out, pos_out);
}

// ---- K11: dense LM, the whole loop in one launch on a cluster of PNP_CLUSTER CTAs ---------------
// The matches are striped over the cluster's threads; every reduction (cost, or the 45 + 9
// normal-equation sums) goes warp shuffle -> CTA shared memory -> distributed shared memory: each
// CTA publishes its partial, one cluster barrier, and every CTA adds the PNP_CLUSTER partials in
// rank order.  All CTAs therefore hold bit-identical totals and run the (tiny) 9x9 solve and the
// LM accept/reject logic redundantly: no second barrier, no global memory, no host round trip.
constexpr int PNP_THREADS = 512;
constexpr int PNP_CLUSTER = 8;       // portable cluster size
constexpr int PNP_NACC = 45 + 9;     // lower triangle of JtJ + Jtr

struct PnpShared {
    double red[32][PNP_NACC + 2];
    double part[2][PNP_NACC + 2];    // this CTA's partial sums, double-buffered across reductions
    double total[PNP_NACC + 2];
    int parity;
    float JtJ[81];
    float diag[9];
    float Jtr[9];
    float L[81];
    float step[9];
    pc_camera_state cam, cam_new;
    int flag;
};

__device__ __forceinline__ double block_reduce_many(PnpShared& sh, const float* vals, int n) {
    // reduces vals[0..n) over the whole cluster into sh.total[0..n) (double, same bits in every CTA)
    cg::cluster_group cluster = cg::this_cluster();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int k = 0; k < n; k++) {
        double v = (double)vals[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) sh.red[wid][k] = v;
    }
    __syncthreads();
    const int par = sh.parity;
    if (threadIdx.x < n) {
        double s = 0.0;
        for (int w = 0; w < PNP_THREADS / 32; w++) s += sh.red[w][threadIdx.x];
        sh.part[par][threadIdx.x] = s;
    }
    cluster.sync();                                  // partials of every CTA are visible
    if (threadIdx.x < n) {
        double s = 0.0;
        for (unsigned r = 0; r < cluster.num_blocks(); r++)
            s += *cluster.map_shared_rank(&sh.part[par][threadIdx.x], r);
        sh.total[threadIdx.x] = s;
    }
    if (threadIdx.x == 0) sh.parity = par ^ 1;       // the next reduction publishes into the other buffer
    __syncthreads();
    return 0.0;
}

// PnPProblem::Evaluate (pnp_problem.h:52-61) -> loss contribution of residual i
__device__ __forceinline__ void pnp_residual(const Cam& c, const float* X, const float* x, int i, float& rx,
                                             float& ry, bool& behind) {
    const V3 P = v3(X[3 * i], X[3 * i + 1], X[3 * i + 2]);
    const V3 Z = mul(c.R, P) + c.t;
    behind = is_behind(c, Z);
    rx = c.fx * Z.x / Z.z + c.cx - x[2 * i];
    ry = c.fy * Z.y / Z.z + c.cy - x[2 * i + 1];
}

__device__ float pnp_total_cost(PnpShared& sh, const pc_camera_state& cs, const Loss& loss, const float* X,
                                const float* x, const float* w, const uint8_t* valid, int m) {
    const Cam c = make_cam(cs);
    float acc[1] = {0.f};
    const int gtid = cg::this_cluster().block_rank() * PNP_THREADS + threadIdx.x;
    const int gthreads = cg::this_cluster().num_blocks() * PNP_THREADS;
    for (int i = gtid; i < m; i += gthreads) {
        if (valid && !valid[i]) continue;
        const float wi = w ? w[i] : 1.f;
        if (wi == 0.f) continue;                               // lev_marq.h:333-336
        float rx, ry;
        bool behind;
        pnp_residual(c, X, x, i, rx, ry, behind);
        float r2 = rx * rx + ry * ry;
        if (behind) r2 = INFINITY;                             // (FLT_MAX, FLT_MAX).squaredNorm() overflows
        acc[0] += wi * loss_value(loss, r2);
    }
    block_reduce_many(sh, acc, 1);
    return (float)sh.total[0];
}

__global__ void __launch_bounds__(PNP_THREADS, 1) pnp_lm_kernel(const float* __restrict__ X,
                                                                const float* __restrict__ x,
                                                                const float* __restrict__ w,
                                                                const uint8_t* __restrict__ valid, int m,
                                                                PnpParams prm, pc_camera_state* cam_io,
                                                                PnpResult* result) {
    __shared__ PnpShared sh;
    cg::cluster_group cluster = cg::this_cluster();
    const int tid = threadIdx.x;
    const bool lead = cluster.block_rank() == 0 && tid == 0;   // the one thread that writes results
    const int gtid = cluster.block_rank() * PNP_THREADS + tid;
    const int gthreads = cluster.num_blocks() * PNP_THREADS;
    if (tid == 0) sh.parity = 0;
    __syncthreads();
    // number of usable matches (rays that hit)
    {
        float cnt[1] = {0.f};
        for (int i = gtid; i < m; i += gthreads) cnt[0] += (!valid || valid[i]) ? 1.f : 0.f;
        block_reduce_many(sh, cnt, 1);
    }
    const int n_valid = (int)sh.total[0];
    __syncthreads();
    if (lead) {
        result->num_matches = n_valid;
        result->status = 0;
    }
    if (n_valid < 3) {                                         // tracker.cc:95-97 / solvers.cc:55
        if (lead) result->status = 1;
        cluster.sync();                                        // nobody leaves while a peer may still read its smem
        return;
    }
    // pnp_problem.h:33-34: intrinsics are only optimised with more than 3 points
    const bool opt_f = prm.opt_f && n_valid > 3, opt_pp = prm.opt_pp && n_valid > 3;
    const Loss loss = make_loss(prm.loss_type, prm.loss_scale);
    if (tid == 0) sh.cam = *cam_io;
    __syncthreads();

    float cost = pnp_total_cost(sh, sh.cam, loss, X, x, w, valid, m);
    const float initial_cost = cost;
    float lambda = prm.initial_lambda, v = 2.f;
    float grad_norm = -1.f, step_norm = -1.f;
    unsigned long long invalid_steps = 0, it = 0;
    bool rebuild = true;
    for (it = 0; it < prm.max_iterations; ++it) {
        if (rebuild) {
            // BuildNormalEquations (lev_marq.h:231-297) with PnPProblem::EvaluateWithJacobian
            const Cam c = make_cam(sh.cam);
            float acc[PNP_NACC];
#pragma unroll
            for (int k = 0; k < PNP_NACC; k++) acc[k] = 0.f;
            for (int i = gtid; i < m; i += gthreads) {
                if (valid && !valid[i]) continue;
                const float wi = w ? w[i] : 1.f;
                if (wi == 0.f) continue;
                const V3 P = v3(X[3 * i], X[3 * i + 1], X[3 * i + 2]);
                const V3 Z = mul(c.R, P) + c.t;
                const float iz = 1.f / Z.z;
                const float rx = c.fx * Z.x / Z.z + c.cx - x[2 * i];
                const float ry = c.fy * Z.y / Z.z + c.cy - x[2 * i + 1];
                // dz/dZ (types.h:79-85)
                const float a00 = c.fx * iz, a02 = -c.fx * Z.x / (Z.z * Z.z);
                const float a11 = c.fy * iz, a12 = -c.fy * Z.y / (Z.z * Z.z);
                // dRtZ_dR = R * Skew(-P) (pose.h:83-85)
                float dR[9];
                {
                    const float sx = -P.x, sy = -P.y, sz = -P.z;
                    // Skew(s) = [0 -sz sy; sz 0 -sx; -sy sx 0]
#pragma unroll
                    for (int r = 0; r < 3; r++) {
                        const float r0 = c.R.m[3 * r], r1 = c.R.m[3 * r + 1], r2 = c.R.m[3 * r + 2];
                        dR[3 * r + 0] = r1 * sz - r2 * sy;
                        dR[3 * r + 1] = -r0 * sz + r2 * sx;
                        dR[3 * r + 2] = r0 * sy - r1 * sx;
                    }
                }
                float J0[9], J1[9];
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    J0[k] = a00 * dR[k] + a02 * dR[6 + k];
                    J1[k] = a11 * dR[3 + k] + a12 * dR[6 + k];
                }
                J0[3] = a00; J0[4] = 0.f; J0[5] = a02;
                J1[3] = 0.f; J1[4] = a11; J1[5] = a12;
                J0[6] = opt_f ? c.aspect * Z.x / Z.z : 0.f;     // types.h:88-92
                J1[6] = opt_f ? Z.y / Z.z : 0.f;
                J0[7] = opt_pp ? 1.f : 0.f; J1[7] = 0.f;
                J0[8] = 0.f; J1[8] = opt_pp ? 1.f : 0.f;
                const float tw = wi * loss_weight(loss, rx * rx + ry * ry);   // lev_marq.h:266-268
                int k = 0;
#pragma unroll
                for (int r = 0; r < 9; r++)
#pragma unroll
                    for (int cc = 0; cc <= r; cc++) acc[k++] += tw * (J0[r] * J0[cc] + J1[r] * J1[cc]);
                const float wrx = tw * rx, wry = tw * ry;
#pragma unroll
                for (int r = 0; r < 9; r++) acc[45 + r] += J0[r] * wrx + J1[r] * wry;
            }
            block_reduce_many(sh, acc, PNP_NACC);
            if (tid == 0) {
                int k = 0;
                for (int r = 0; r < 9; r++)
                    for (int cc = 0; cc <= r; cc++) {
                        const float val = (float)sh.total[k++];
                        sh.JtJ[r * 9 + cc] = val;
                        sh.JtJ[cc * 9 + r] = val;
                    }
                float g2 = 0.f;
                for (int r = 0; r < 9; r++) {
                    sh.Jtr[r] = (float)sh.total[45 + r];
                    g2 += sh.Jtr[r] * sh.Jtr[r];
                    sh.diag[r] = fminf(fmaxf(sh.JtJ[r * 9 + r], 1e-6f), 1e32f);   // lev_marq.h:296
                }
                sh.step[0] = sqrtf(g2);     // stash grad norm
            }
            __syncthreads();
            grad_norm = sh.step[0];
            __syncthreads();
            if (grad_norm < prm.gradient_tol) break;
        }
        // ComputeStep (lev_marq.h:299-314): damp the diagonal, 9x9 LLT (lower), solve
        if (tid == 0) {
            float* L = sh.L;
            for (int r = 0; r < 9; r++)
                for (int cc = 0; cc < 9; cc++) L[r * 9 + cc] = sh.JtJ[r * 9 + cc];
            for (int r = 0; r < 9; r++) L[r * 9 + r] = sh.diag[r] * (float)(1.0 + (double)lambda);
            int ok = 1;
            for (int k = 0; k < 9 && ok; k++) {
                float xk = L[k * 9 + k];
                for (int j = 0; j < k; j++) xk -= L[k * 9 + j] * L[k * 9 + j];
                if (!(xk > 0.f)) { ok = 0; break; }
                xk = sqrtf(xk);
                L[k * 9 + k] = xk;
                for (int r = k + 1; r < 9; r++) {
                    float s = L[r * 9 + k];
                    for (int j = 0; j < k; j++) s -= L[r * 9 + j] * L[k * 9 + j];
                    L[r * 9 + k] = s / xk;
                }
            }
            if (ok) {
                float y[9];
                for (int r = 0; r < 9; r++) {
                    float s = sh.Jtr[r];
                    for (int j = 0; j < r; j++) s -= L[r * 9 + j] * y[j];
                    y[r] = s / L[r * 9 + r];
                }
                for (int r = 8; r >= 0; r--) {
                    float s = y[r];
                    for (int j = r + 1; j < 9; j++) s -= L[j * 9 + r] * y[j];
                    y[r] = s / L[r * 9 + r];
                }
                float n2 = 0.f;
                for (int r = 0; r < 9; r++) { sh.step[r] = -y[r]; n2 += y[r] * y[r]; }
                sh.L[0] = sqrtf(n2);   // stash step norm
            }
            sh.flag = ok;
        }
        __syncthreads();
        const int llt_ok = sh.flag;
        if (!llt_ok) {                                          // lev_marq.h:158-169
            invalid_steps++;
            if (lambda == prm.max_lambda) break;
            lambda = fminf(prm.max_lambda, lambda * v);
            v = 2.f * v;
            rebuild = false;
            __syncthreads();
            continue;
        }
        step_norm = sh.L[0];
        if (step_norm < prm.step_tol) break;
        if (tid == 0) camera_step(sh.cam, sh.step, opt_f, opt_pp, prm.bounds, sh.cam_new);   // pnp_problem.h:101-131
        __syncthreads();
        const float cost_new = pnp_total_cost(sh, sh.cam_new, loss, X, x, w, valid, m);
        if (cost_new < cost) {                                  // lev_marq.h:179-203
            if (tid == 0) {
                const float actual = cost_new - cost;
                // step^T (2 Jtr + JtJ_sym(undamped, clamped diag) step)
                float expected = 0.f;
                for (int r = 0; r < 9; r++) {
                    float s = 0.f;
                    for (int cc = 0; cc < 9; cc++)
                        s += (r == cc ? sh.diag[r] : sh.JtJ[r * 9 + cc]) * sh.step[cc];
                    expected += sh.step[r] * (2.f * sh.Jtr[r] + s);
                }
                const float rho = actual / expected;
                float lam = lambda;
                if (rho > 0.f) {
                    const float f = (float)fmax(1.0 / 3.0, 1.0 - pow(2.0 * (double)rho - 1.0, 3.0));   // Float factor
                    lam = fminf(fmaxf(lambda * f, prm.min_lambda), prm.max_lambda);
                }
                sh.L[1] = lam;
                sh.cam = sh.cam_new;
            }
            __syncthreads();
            lambda = sh.L[1];
            cost = cost_new;
            v = 2.f;
            rebuild = true;
            __syncthreads();
        } else {
            invalid_steps++;
            if (lambda == prm.max_lambda) break;
            lambda = fminf(prm.max_lambda, lambda * v);
            v = 2.f * v;
            rebuild = false;
        }
    }
    __syncthreads();
    // inlier ratio (solvers.cc:30-47)
    const Cam c = make_cam(sh.cam);
    float acc[1] = {0.f};
    if (prm.max_inlier_error > 0.f) {
        const float thr2 = prm.max_inlier_error * prm.max_inlier_error;
        for (int i = gtid; i < m; i += gthreads) {
            if (valid && !valid[i]) continue;
            float rx, ry;
            bool behind;
            pnp_residual(c, X, x, i, rx, ry, behind);
            float e2 = rx * rx + ry * ry;
            if (behind) e2 = INFINITY;
            if (e2 < thr2) acc[0] += 1.f;
        }
    }
    block_reduce_many(sh, acc, 1);
    if (lead) {
        *cam_io = sh.cam;
        result->stats.iterations = it;
        result->stats.initial_cost = initial_cost;
        result->stats.cost = cost;
        result->stats.lambda = lambda;
        result->stats.invalid_steps = invalid_steps;
        result->stats.step_norm = step_norm;
        result->stats.grad_norm = grad_norm;
        result->inlier_ratio = (float)sh.total[0] / (float)n_valid;
    }
    cluster.sync();                                            // keep every CTA's shared memory alive until all have read it
}

void launch_pnp_lm(const float* X, const float* x, const float* w, const uint8_t* valid, int m, const PnpParams& prm,
                   pc_camera_state* cam_io, PnpResult* result, cudaStream_t s) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(PNP_CLUSTER);
    cfg.blockDim = dim3(PNP_THREADS);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = PNP_CLUSTER;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, pnp_lm_kernel, X, x, w, valid, m, prm, cam_io, result);
}

}  // namespace pc
