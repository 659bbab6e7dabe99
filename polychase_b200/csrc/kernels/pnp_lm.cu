// K11: the robust Levenberg-Marquardt pose (+ intrinsics) solve of the tracker, one launch per
// frame, the whole LM loop on device.
//
// Replaces LevMarqDenseSolve<PnPProblem, Loss> (/root/reference/cpp/pnp/lev_marq.h:99-389,
// /root/reference/cpp/pnp/pnp_problem.h:52-131, /root/reference/cpp/pnp/solvers.cc:11-48).
//
// The tracker is a sequential chain of small problems (<= 8 sources x max_features matches, a
// dozen LM iterations each), so the kernel is built for LATENCY, not bandwidth:
//   * one thread-block cluster (16 CTAs where the device can place one, else 8) owns the problem.  The matches are staged
//     once into shared memory as structure-of-arrays (ray misses / zero weights folded into one
//     weight array) and never re-read from global memory;
//   * every reduction (cost, or the lower triangle of JtJ + Jtr) is a transposing warp reduction
//     in float64 (2 shuffles per value instead of 10: after five halving steps lane l holds the
//     warp total of value l), one shared-memory hop across the CTA's warps, one cluster barrier,
//     and a distributed-shared-memory read of the peers' partials in rank order -- every CTA ends
//     up with bit-identical totals;
//   * what follows a reduction (diagonal clamp, damping, the NPxNP LLT, the step, the camera
//     update, the accept/reject and lambda logic) is done REDUNDANTLY BY EVERY THREAD in registers,
//     fully unrolled: no broadcast, no second barrier, no single-thread shared-memory code on the
//     critical path.
// Per-thread partial sums are float32; cross-thread sums are float64 (the reference's own sums are
// order-nondeterministic under TBB; this sits inside that band).
#include <cooperative_groups.h>

#include "common.cuh"
#include "geom.cuh"
#include "kernels.h"
#include "track_kernels.h"

namespace cg = cooperative_groups;

namespace pc {

namespace {

constexpr int PNP_THREADS = 256;
constexpr int PNP_WARPS = PNP_THREADS / 32;
constexpr int PNP_MAXACC = 64;       // two groups of 32 values
constexpr unsigned FULL = 0xffffffffu;
constexpr int PNP_STAGE_ARRAYS = 6;  // X0 X1 X2 x0 x1 weight
constexpr unsigned PNP_MAX_CLUSTER = 16;

struct PnpShared {
    double warp_part[PNP_WARPS][PNP_MAXACC];
    double part[2][PNP_MAXACC];      // this CTA's partial sums, double-buffered across reductions
    double total[PNP_MAXACC];
};

__host__ __device__ constexpr int tri(int r, int c) { return r * (r + 1) / 2 + c; }   // r >= c

// Transposing warp reduction of 32 values: on return lane l holds sum over the warp of v[l]
// (returned value); v is clobbered.
__device__ __forceinline__ double warp_reduce32(double (&v)[32]) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const bool upper = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < o; i++) {
            const double send = upper ? v[i] : v[i + o];
            const double keep = upper ? v[i + o] : v[i];
            v[i] = keep + __shfl_xor_sync(FULL, send, o);
        }
    }
    return v[0];
}

// Sums vals[0..N) (one float per thread and value) over the whole cluster; afterwards
// sh.total[0..N) holds the totals (same bits in every CTA) and all threads may read them.
template <int N>
__device__ __forceinline__ void cluster_reduce(PnpShared& sh, int& parity, const float (&vals)[N]) {
    static_assert(N <= PNP_MAXACC, "too many accumulators");
    cg::cluster_group cluster = cg::this_cluster();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (N == 1) {
        double v = (double)vals[0];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
        if (lane == 0) sh.warp_part[wid][0] = v;
    } else {
#pragma unroll
        for (int g = 0; g < (N + 31) / 32; g++) {
            double v[32];
#pragma unroll
            for (int i = 0; i < 32; i++) v[i] = (g * 32 + i < N) ? (double)vals[(g * 32 + i < N) ? g * 32 + i : 0] : 0.0;
            const double r = warp_reduce32(v);
            if (g * 32 + lane < N) sh.warp_part[wid][g * 32 + lane] = r;
        }
    }
    __syncthreads();
    if (threadIdx.x < N) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < PNP_WARPS; w++) s += sh.warp_part[w][threadIdx.x];
        sh.part[parity][threadIdx.x] = s;
    }
    cluster.sync();                                  // partials of every CTA are visible
    if (threadIdx.x < N) {
        // all remote reads in flight at once (a rolled loop pays one DSMEM round trip per peer), then
        // the sum in rank order so that every CTA gets the same bits
        const unsigned nb = cluster.num_blocks();
        double part[PNP_MAX_CLUSTER];
#pragma unroll
        for (unsigned r = 0; r < PNP_MAX_CLUSTER; r++)
            part[r] = r < nb ? *cluster.map_shared_rank(&sh.part[parity][threadIdx.x], r) : 0.0;
        double s = 0.0;
#pragma unroll
        for (unsigned r = 0; r < PNP_MAX_CLUSTER; r++)
            if (r < nb) s += part[r];
        sh.total[threadIdx.x] = s;
    }
    parity ^= 1;                                     // the next reduction publishes into the other buffer
    __syncthreads();
}

struct Match { float X0, X1, X2, u, v, wt; };

}  // namespace

// NP = 6: pose only; NP = 9: pose + (fy, cx, cy) columns (zeroed unless enabled, pnp_problem.h:86-96)
template <int NP>
__global__ void __launch_bounds__(PNP_THREADS, 1)
pnp_lm_kernel(const float* __restrict__ X, const float* __restrict__ x, const float* __restrict__ w,
              const uint8_t* __restrict__ valid, int m, int per_thread, int staged, PnpParams prm,
              const pc_camera_state* __restrict__ cam_in, pc_camera_state* __restrict__ cam_out,
              PnpResult* __restrict__ result) {
    constexpr int NJ = NP * (NP + 1) / 2;
    constexpr int NACC = NJ + NP;
    __shared__ PnpShared sh;
    extern __shared__ float stage[];                 // PNP_STAGE_ARRAYS x (per_thread * PNP_THREADS)
    cg::cluster_group cluster = cg::this_cluster();
    const int tid = threadIdx.x;
    const unsigned rank = cluster.block_rank(), nblocks = cluster.num_blocks();
    const bool lead = rank == 0 && tid == 0;         // the one thread that writes results
    const int per_cta = per_thread * PNP_THREADS;
    int parity = 0;

    // match k of this thread: global row (k * nblocks + rank) * PNP_THREADS + tid
    auto fetch_global = [&](int k) {
        Match mt;
        const int i = (k * (int)nblocks + (int)rank) * PNP_THREADS + tid;
        mt.wt = 0.f; mt.X0 = mt.X1 = mt.X2 = mt.u = mt.v = 0.f;
        if (i < m && (!valid || valid[i])) {
            mt.wt = w ? w[i] : 1.f;
            mt.X0 = X[3 * i]; mt.X1 = X[3 * i + 1]; mt.X2 = X[3 * i + 2];
            mt.u = x[2 * i]; mt.v = x[2 * i + 1];
        }
        return mt;
    };
    auto fetch = [&](int k) {
        if (!staged) return fetch_global(k);
        Match mt;
        const int j = k * PNP_THREADS + tid;
        mt.X0 = stage[j]; mt.X1 = stage[per_cta + j]; mt.X2 = stage[2 * per_cta + j];
        mt.u = stage[3 * per_cta + j]; mt.v = stage[4 * per_cta + j]; mt.wt = stage[5 * per_cta + j];
        return mt;
    };

    pc_camera_state cam = *cam_in;
    const Loss loss = make_loss(prm.loss_type, prm.loss_scale);

    // ---- stage the matches, count the usable ones (rays that hit), initial cost --------------
    // (the weight == 0 skip is lev_marq.h:333-336)
    auto first_cost = [&](const pc_camera_state& cs) {
        const Cam c = make_cam(cs);
        float acc = 0.f, cnt = 0.f;
        for (int k = 0; k < per_thread; k++) {
            const Match mt = fetch_global(k);
            const int i = (k * (int)nblocks + (int)rank) * PNP_THREADS + tid;
            if (i < m && (!valid || valid[i])) cnt += 1.f;
            if (staged) {
                const int j = k * PNP_THREADS + tid;
                stage[j] = mt.X0; stage[per_cta + j] = mt.X1; stage[2 * per_cta + j] = mt.X2;
                stage[3 * per_cta + j] = mt.u; stage[4 * per_cta + j] = mt.v; stage[5 * per_cta + j] = mt.wt;
            }
            if (mt.wt == 0.f) continue;
            // PnPProblem::Evaluate (pnp_problem.h:52-61)
            const V3 Z = mul(c.R, v3(mt.X0, mt.X1, mt.X2)) + c.t;
            const float rx = c.fx * Z.x / Z.z + c.cx - mt.u;
            const float ry = c.fy * Z.y / Z.z + c.cy - mt.v;
            float r2 = rx * rx + ry * ry;
            if (is_behind(c, Z)) r2 = INFINITY;      // (FLT_MAX, FLT_MAX).squaredNorm() overflows
            acc += mt.wt * loss_value(loss, r2);
        }
        const float two[2] = {acc, cnt};
        cluster_reduce<2>(sh, parity, two);
        return (float)sh.total[0];
    };

    float cost = first_cost(cam);
    const int n_valid = (int)sh.total[1];
    if (lead) {
        result->num_matches = n_valid;
        result->status = 0;
    }
    if (n_valid < 3) {                               // tracker.cc:95-97 / solvers.cc:55
        if (lead) {
            result->status = 1;
            *cam_out = cam;                          // a chained successor still reads a defined pose
        }
        cluster.sync();                              // nobody leaves while a peer may still read its smem
        return;
    }
    // pnp_problem.h:33-34: intrinsics are only optimised with more than 3 points
    const bool opt_f = NP == 9 && prm.opt_f && n_valid > 3, opt_pp = NP == 9 && prm.opt_pp && n_valid > 3;

    const float initial_cost = cost;
    float lambda = prm.initial_lambda, v = 2.f;
    float grad_norm = -1.f, step_norm = -1.f;
    unsigned long long invalid_steps = 0, it = 0;
    bool rebuild = true;
    float A[NJ], diag[NP], Jtr[NP];                  // lower triangle of JtJ (undamped), clamped diagonal
#pragma unroll
    for (int k = 0; k < NJ; k++) A[k] = 0.f;
#pragma unroll
    for (int k = 0; k < NP; k++) { diag[k] = 0.f; Jtr[k] = 0.f; }

    // One pass over the matches at `cs`: the lower triangle of JtJ, Jtr (BuildNormalEquations,
    // lev_marq.h:231-297, with PnPProblem::EvaluateWithJacobian) AND the total cost, reduced together.
    // The LM loop evaluates every candidate step with it: when the step is accepted (the common case)
    // the normal equations of the next iteration are already there, so an iteration costs one pass
    // and one cluster reduction instead of two of each.  The sums are those of the separate passes.
    auto build_at = [&](const pc_camera_state& cs) {
            const Cam c = make_cam(cs);
            float acc[NACC + 1];
#pragma unroll
            for (int k = 0; k < NACC + 1; k++) acc[k] = 0.f;
            for (int k = 0; k < per_thread; k++) {
                const Match mt = fetch(k);
                if (mt.wt == 0.f) continue;
                const V3 P = v3(mt.X0, mt.X1, mt.X2);
                const V3 Z = mul(c.R, P) + c.t;
                const float iz = 1.f / Z.z;
                const float rx = c.fx * Z.x / Z.z + c.cx - mt.u;
                const float ry = c.fy * Z.y / Z.z + c.cy - mt.v;
                {                                            // TotalCost of the same point (as in first_cost)
                    float r2 = rx * rx + ry * ry;
                    if (is_behind(c, Z)) r2 = INFINITY;
                    acc[NACC] += mt.wt * loss_value(loss, r2);
                }
                // dz/dZ (types.h:79-85)
                const float a00 = c.fx * iz, a02 = -c.fx * Z.x / (Z.z * Z.z);
                const float a11 = c.fy * iz, a12 = -c.fy * Z.y / (Z.z * Z.z);
                // dRtZ_dR = R * Skew(-P) (pose.h:83-85); Skew(s) = [0 -sz sy; sz 0 -sx; -sy sx 0]
                float dR[9];
                {
                    const float sx = -P.x, sy = -P.y, sz = -P.z;
#pragma unroll
                    for (int r = 0; r < 3; r++) {
                        const float r0 = c.R.m[3 * r], r1 = c.R.m[3 * r + 1], r2 = c.R.m[3 * r + 2];
                        dR[3 * r + 0] = r1 * sz - r2 * sy;
                        dR[3 * r + 1] = -r0 * sz + r2 * sx;
                        dR[3 * r + 2] = r0 * sy - r1 * sx;
                    }
                }
                float J0[NP], J1[NP];
#pragma unroll
                for (int q = 0; q < 3; q++) {
                    J0[q] = a00 * dR[q] + a02 * dR[6 + q];
                    J1[q] = a11 * dR[3 + q] + a12 * dR[6 + q];
                }
                J0[3] = a00; J0[4] = 0.f; J0[5] = a02;
                J1[3] = 0.f; J1[4] = a11; J1[5] = a12;
                if (NP == 9) {
                    J0[NP - 3] = opt_f ? c.aspect * Z.x / Z.z : 0.f;     // types.h:88-92
                    J1[NP - 3] = opt_f ? Z.y / Z.z : 0.f;
                    J0[NP - 2] = opt_pp ? 1.f : 0.f; J1[NP - 2] = 0.f;
                    J0[NP - 1] = 0.f; J1[NP - 1] = opt_pp ? 1.f : 0.f;
                }
                const float tw = mt.wt * loss_weight(loss, rx * rx + ry * ry);   // lev_marq.h:266-268
                int q = 0;
#pragma unroll
                for (int r = 0; r < NP; r++)
#pragma unroll
                    for (int cc = 0; cc <= r; cc++) acc[q++] += tw * (J0[r] * J0[cc] + J1[r] * J1[cc]);
                const float wrx = tw * rx, wry = tw * ry;
#pragma unroll
                for (int r = 0; r < NP; r++) acc[NJ + r] += J0[r] * wrx + J1[r] * wry;
            }
            cluster_reduce<NACC + 1>(sh, parity, acc);
    };

    bool built = false;                              // sh.total holds the normal equations of `cam`
    for (it = 0; it < prm.max_iterations; ++it) {
        if (rebuild) {
            if (!built) build_at(cam);
            built = false;
            float g2 = 0.f;
#pragma unroll
            for (int k = 0; k < NJ; k++) A[k] = (float)sh.total[k];
#pragma unroll
            for (int r = 0; r < NP; r++) {
                Jtr[r] = (float)sh.total[NJ + r];
                g2 += Jtr[r] * Jtr[r];
                diag[r] = fminf(fmaxf(A[tri(r, r)], 1e-6f), 1e32f);     // lev_marq.h:296
            }
            grad_norm = sqrtf(g2);
            if (grad_norm < prm.gradient_tol) break;
        }
        // ComputeStep (lev_marq.h:299-314): damp the diagonal, NPxNP LLT (lower), solve
        float L[NJ], step[NP];
        bool ok = true;
        {
            const float damp = (float)(1.0 + (double)lambda);
#pragma unroll
            for (int k = 0; k < NP; k++) {
                float xk = diag[k] * damp;
#pragma unroll
                for (int j = 0; j < k; j++) xk -= L[tri(k, j)] * L[tri(k, j)];
                if (!(xk > 0.f)) ok = false;
                xk = sqrtf(xk);
                L[tri(k, k)] = xk;
#pragma unroll
                for (int r = k + 1; r < NP; r++) {
                    float s = A[tri(r, k)];
#pragma unroll
                    for (int j = 0; j < k; j++) s -= L[tri(r, j)] * L[tri(k, j)];
                    L[tri(r, k)] = s / xk;
                }
            }
        }
        if (!ok) {                                              // lev_marq.h:158-169
            invalid_steps++;
            if (lambda == prm.max_lambda) break;
            lambda = fminf(prm.max_lambda, lambda * v);
            v = 2.f * v;
            rebuild = false;
            continue;
        }
        {
            float y[NP];
#pragma unroll
            for (int r = 0; r < NP; r++) {
                float s = Jtr[r];
#pragma unroll
                for (int j = 0; j < r; j++) s -= L[tri(r, j)] * y[j];
                y[r] = s / L[tri(r, r)];
            }
#pragma unroll
            for (int r = NP - 1; r >= 0; r--) {
                float s = y[r];
#pragma unroll
                for (int j = r + 1; j < NP; j++) s -= L[tri(j, r)] * y[j];
                y[r] = s / L[tri(r, r)];
            }
            float n2 = 0.f;
#pragma unroll
            for (int r = 0; r < NP; r++) { step[r] = -y[r]; n2 += y[r] * y[r]; }
            step_norm = sqrtf(n2);
        }
        if (step_norm < prm.step_tol) break;
        pc_camera_state cam_new;
        {
            float dp[9];
#pragma unroll
            for (int r = 0; r < 9; r++) dp[r] = r < NP ? step[r < NP ? r : 0] : 0.f;
            camera_step(cam, dp, opt_f, opt_pp, prm.bounds, cam_new);          // pnp_problem.h:101-131
        }
        build_at(cam_new);                                      // cost of the step + (speculatively) its normal equations
        const float cost_new = (float)sh.total[NACC];
        if (cost_new < cost) {                                  // lev_marq.h:179-203
            const float actual = cost_new - cost;
            // step^T (2 Jtr + JtJ_sym(undamped, clamped diag) step)
            float expected = 0.f;
#pragma unroll
            for (int r = 0; r < NP; r++) {
                float s = 0.f;
#pragma unroll
                for (int cc = 0; cc < NP; cc++)
                    s += (r == cc ? diag[r] : (r > cc ? A[tri(r, cc)] : A[tri(cc, r)])) * step[cc];
                expected += step[r] * (2.f * Jtr[r] + s);
            }
            const float rho = actual / expected;
            if (rho > 0.f) {
                const double d = 2.0 * (double)rho - 1.0;
                const float f = (float)fmax(1.0 / 3.0, 1.0 - d * d * d);       // Float factor
                lambda = fminf(fmaxf(lambda * f, prm.min_lambda), prm.max_lambda);
            }
            cam = cam_new;
            cost = cost_new;
            v = 2.f;
            rebuild = true;
            built = true;                                       // sh.total: nothing has been reduced since
        } else {
            invalid_steps++;
            if (lambda == prm.max_lambda) break;
            lambda = fminf(prm.max_lambda, lambda * v);
            v = 2.f * v;
            rebuild = false;
        }
    }
    // inlier ratio (solvers.cc:30-47)
    {
        const Cam c = make_cam(cam);
        float acc[1] = {0.f};
        if (prm.max_inlier_error > 0.f) {
            const float thr2 = prm.max_inlier_error * prm.max_inlier_error;
            for (int k = 0; k < per_thread; k++) {
                const int i = (k * (int)nblocks + (int)rank) * PNP_THREADS + tid;
                if (i >= m || (valid && !valid[i])) continue;         // zero-weight matches still count here
                const Match mt = staged ? fetch(k) : fetch_global(k);
                const V3 Z = mul(c.R, v3(mt.X0, mt.X1, mt.X2)) + c.t;
                const float rx = c.fx * Z.x / Z.z + c.cx - mt.u;
                const float ry = c.fy * Z.y / Z.z + c.cy - mt.v;
                float e2 = rx * rx + ry * ry;
                if (is_behind(c, Z)) e2 = INFINITY;
                if (e2 < thr2) acc[0] += 1.f;
            }
        }
        cluster_reduce<1>(sh, parity, acc);
    }
    if (lead) {
        *cam_out = cam;
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

namespace {
int g_pnp_cluster = 0;        // 0 = not decided yet
size_t g_pnp_dyn_max = 0;
}  // namespace

int pnp_cluster_size() {
    if (g_pnp_cluster) return g_pnp_cluster;
    // 16 CTAs (the non-portable maximum, allowed on sm_100) when the device can place such a cluster,
    // else the portable 8: the solve is a latency chain, and the per-iteration pass over the matches
    // halves with twice the CTAs (16.3 vs 21.0 ms per 32 frames in the 4K pipeline)
    int want = 16;
    if (const char* e = getenv("PC_PNP_CLUSTER")) {
        const int vv = atoi(e);
        if (vv == 1 || vv == 2 || vv == 4 || vv == 8 || vv == 16) want = vv;
    }
    g_pnp_dyn_max = 160 * 1024;
    cudaFuncSetAttribute(pnp_lm_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g_pnp_dyn_max);
    cudaFuncSetAttribute(pnp_lm_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g_pnp_dyn_max);
    if (want > 8) {
        cudaError_t e1 = cudaFuncSetAttribute(pnp_lm_kernel<6>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        cudaError_t e2 = cudaFuncSetAttribute(pnp_lm_kernel<9>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e1 != cudaSuccess || e2 != cudaSuccess) { cudaGetLastError(); want = 8; }
        if (want > 8) {                                         // can a 16-CTA cluster be resident at all?
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(want);
            cfg.blockDim = dim3(PNP_THREADS);
            cfg.dynamicSmemBytes = 96 * 1024;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = want;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            int n6 = 0, n9 = 0;
            const cudaError_t q1 = cudaOccupancyMaxActiveClusters(&n6, pnp_lm_kernel<6>, &cfg);
            const cudaError_t q2 = cudaOccupancyMaxActiveClusters(&n9, pnp_lm_kernel<9>, &cfg);
            if (q1 != cudaSuccess || q2 != cudaSuccess || n6 < 1 || n9 < 1) { cudaGetLastError(); want = 8; }
        }
    }
    g_pnp_cluster = want;
    return want;
}

void launch_pnp_lm(const float* X, const float* x, const float* w, const uint8_t* valid, int m, const PnpParams& prm,
                   const pc_camera_state* cam_in, pc_camera_state* cam_out, PnpResult* result, cudaStream_t s) {
    const int cluster = pnp_cluster_size();
    const int group = cluster * PNP_THREADS;
    const int per_thread = (m + group - 1) / group;
    size_t dyn = (size_t)per_thread * PNP_THREADS * PNP_STAGE_ARRAYS * sizeof(float);
    int staged = 1;
    if (dyn > g_pnp_dyn_max) { staged = 0; dyn = 0; }         // very large problems stream from L2 instead
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cluster);
    cfg.blockDim = dim3(PNP_THREADS);
    cfg.dynamicSmemBytes = dyn;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (prm.opt_f || prm.opt_pp)
        cudaLaunchKernelEx(&cfg, pnp_lm_kernel<9>, X, x, w, valid, m, per_thread, staged, prm, cam_in, cam_out, result);
    else
        cudaLaunchKernelEx(&cfg, pnp_lm_kernel<6>, X, x, w, valid, m, per_thread, staged, prm, cam_in, cam_out, result);
}

}  // namespace pc
