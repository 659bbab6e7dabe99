// K14 + the Levenberg-Marquardt loop of the refiner on the device.
//
//   LevMarqSparseSolver::Solve       /root/reference/cpp/pnp/lev_marq.h:492-588 (the state machine, App. B of SURVEY.md)
//   ComputeStep (SimplicialLLT)      /root/reference/cpp/pnp/lev_marq.h:826-841
//   GlobalRefinementProblem::Step    /root/reference/cpp/refiner.cc:508-537,618-646
//
// The LM state (cost, lambda, the 2^k growth factor, rebuild / done flags, counters) lives in a BALmState
// record in HBM.  One LM iteration is a fixed sequence of launches
//     build + assemble (if `rebuild`)  ->  solve  ->  step  ->  refresh + cost  ->  decide
// in which every kernel first looks at the record and returns when the reference's loop would not have run it
// (a failed factorisation `continue`s past the evaluation, a converged loop has `break`-en).  The host only
// enqueues iterations and reads the record back when it has a callback to serve -- one synchronisation per
// callback instead of three per iteration plus a host-side parameter step.
//
// ba_solve_window_kernel: block-banded Cholesky (half-bandwidth 8 blocks) of the damped normal equations with
// the 9 x 9-block active window held in shared memory for the whole elimination: per pivot block one
// factorisation of the p x p diagonal block, the 8 panel blocks below it, the 36 trailing block updates and
// the forward substitution of the right-hand side, with the next block row streaming in from HBM meanwhile.
// The finished columns of L go out to HBM once and come back once for the backward substitution.
#include "ba_kernels.h"
#include "common.cuh"

namespace pc {

namespace {

constexpr int NB = kBandBlocks;       // 9: window rows / columns (ring indexed by frame % 9)
constexpr int SOLVE_THREADS = 256;

template <int P>
struct SolveSmem {
    float W[NB][NB][P * P];           // block (i, j), j <= i, i - j <= 8, at [i % 9][j % 9]
    float rhs[NB][P];                 // right-hand side rows of the window, forward-substituted in place
    float Lb[2][NB][P * P];           // backward pass: diagonal block + the 8 blocks below it (double buffered)
    float xs[NB][P];                  // backward pass: solution of the last 9 block rows
    float part[NB][P];
    double red[SOLVE_THREADS / 32];
    int ok;
};

__device__ __forceinline__ float damped(const BAView& v, int i, int k, int e, int p, float damp) {
    // element e of block (i, i-k) of the damped matrix: JtJ_diag * (1 + lambda) on the diagonal (lev_marq.h:828)
    const int a = e / p, b = e - a * p;
    if (k == 0 && a == b) return v.diag[i * p + a] * damp;
    return v.band[((size_t)i * NB + k) * p * p + e];
}

}  // namespace

template <int P>
__global__ void __launch_bounds__(SOLVE_THREADS) ba_solve_window_kernel(BAView v, BALmState* st, float lambda_in) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SolveSmem<P>& S = *reinterpret_cast<SolveSmem<P>*>(smem_raw);
    constexpr int PP = P * P;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int nf = v.nf;
    float lambda = lambda_in;
    if (st != nullptr) {                      // device-side LM loop: state decides whether this launch does anything
        if (st->done) return;
        if (tid == 0) st->skip = 0;
        if (st->rebuild) {                    // grad_norm = |Jtr| (lev_marq.h:503-508)
            double a = 0.0;
            for (int k = tid; k < nf * P; k += nt) a += (double)v.jtr[k] * (double)v.jtr[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if ((tid & 31) == 0) S.red[tid >> 5] = a;
            __syncthreads();
            double s = 0.0;
            for (int k = 0; k < nt / 32; k++) s += S.red[k];
            const float gn = (float)sqrt(s);
            __syncthreads();
            if (tid == 0) st->grad_norm = gn;
            if (gn < st->gradient_tol) {
                if (tid == 0) st->done = 1;
                return;
            }
        }
        lambda = st->lambda;
    }
    const float damp = (float)(1.0 + (double)lambda);
    float* Lg = v.lband;
    // ---- initial window: block rows 0..8 -----------------------------------------------------------------
    for (int idx = tid; idx < NB * NB * PP; idx += nt) {
        const int i = idx / (NB * PP), rem = idx - i * NB * PP, k = rem / PP, e = rem - k * PP;
        if (i < nf && i - k >= 0) S.W[i % NB][(i - k) % NB][e] = damped(v, i, k, e, P, damp);
    }
    for (int idx = tid; idx < NB * P; idx += nt) {
        const int i = idx / P;
        if (i < nf) S.rhs[i % NB][idx - i * P] = v.jtr[idx];
    }
    if (tid == 0) S.ok = 1;
    __syncthreads();
    float* y = v.tmp;
    for (int kf = 0; kf < nf; kf++) {
        const int nb = min(NB - 1, nf - 1 - kf);         // block rows below the pivot
        float* D = S.W[kf % NB][kf % NB];
        // -- P0: unblocked LLT of the pivot block + forward substitution of its right-hand side (one thread) --
        if (tid == 0) {
            bool ok = true;
            for (int c = 0; c < P && ok; c++) {
                float x = D[c * P + c];
                for (int j = 0; j < c; j++) x -= D[c * P + j] * D[c * P + j];
                if (!(x > 0.f)) { ok = false; break; }
                x = sqrtf(x);
                D[c * P + c] = x;
                const float ix = 1.f / x;
                for (int r = c + 1; r < P; r++) {
                    float s = D[r * P + c];
                    for (int j = 0; j < c; j++) s -= D[r * P + j] * D[c * P + j];
                    D[r * P + c] = s * ix;
                }
            }
            if (!ok) S.ok = 0;
            else {
                float* b = S.rhs[kf % NB];
                for (int r = 0; r < P; r++) {
                    float s = b[r];
                    for (int j = 0; j < r; j++) s -= D[r * P + j] * b[j];
                    b[r] = s / D[r * P + r];
                    y[kf * P + r] = b[r];
                }
            }
        }
        __syncthreads();
        if (!S.ok) break;
        // -- P1: panel X = B L_kk^-T, one thread per row; the finished column of L goes out to HBM --------------
        if (tid < nb * P) {
            const int i = kf + 1 + tid / P, a = tid % P;
            float* B = S.W[i % NB][kf % NB] + a * P;
            float row[P];
#pragma unroll
            for (int c = 0; c < P; c++) row[c] = B[c];
#pragma unroll
            for (int c = 0; c < P; c++) {
                float s = row[c];
#pragma unroll
                for (int j = 0; j < c; j++) s -= row[j] * D[c * P + j];
                row[c] = s / D[c * P + c];
            }
            float* out = Lg + ((size_t)i * NB + (i - kf)) * PP + a * P;
#pragma unroll
            for (int c = 0; c < P; c++) { B[c] = row[c]; out[c] = row[c]; }
        } else if (tid >= SOLVE_THREADS - PP) {
            const int e = tid - (SOLVE_THREADS - PP), a = e / P, b = e - a * P;
            Lg[((size_t)kf * NB) * PP + e] = b <= a ? D[e] : 0.f;
        }
        __syncthreads();
        // -- P2: trailing update, right-hand side update, and the block row that enters the window ----------------
        const int npairs = nb * (nb + 1) / 2;
        for (int idx = tid; idx < npairs * PP; idx += nt) {
            const int pr = idx / PP, e = idx - pr * PP, a = e / P, b = e - a * P;
            int ii = 0, acc = 0;
            while (acc + ii + 1 <= pr) { acc += ii + 1; ii++; }   // pr -> (ii, jj), jj <= ii
            const int jj = pr - acc;
            const int i = kf + 1 + ii, j = kf + 1 + jj;
            const float* Xi = S.W[i % NB][kf % NB] + a * P;
            const float* Xj = S.W[j % NB][kf % NB] + b * P;
            float s = 0.f;
#pragma unroll
            for (int q = 0; q < P; q++) s += Xi[q] * Xj[q];
            S.W[i % NB][j % NB][e] -= s;
        }
        if (tid < nb * P) {
            const int i = kf + 1 + tid / P, a = tid % P;
            const float* Xi = S.W[i % NB][kf % NB] + a * P;
            const float* yk = S.rhs[kf % NB];
            float s = 0.f;
#pragma unroll
            for (int q = 0; q < P; q++) s += Xi[q] * yk[q];
            S.rhs[i % NB][a] -= s;
        }
        __syncthreads();                                  // the pivot's row / column slots are free now
        const int in = kf + NB;                           // entering block row
        if (in < nf) {
            for (int idx = tid; idx < NB * PP; idx += nt) {
                const int k = idx / PP, e = idx - k * PP;   // block (in, in - k), k = 0..8: columns kf+1 .. kf+9
                S.W[in % NB][(in - k) % NB][e] = damped(v, in, k, e, P, damp);
            }
            if (tid < P) S.rhs[in % NB][tid] = v.jtr[in * P + tid];
        }
        __syncthreads();
    }
    if (!S.ok) {                                          // Eigen::NumericalIssue (lev_marq.h:510-521)
        if (tid == 0) {
            v.scalars[4] = 0.f;
            if (st != nullptr) {
                st->llt_ok = 0;
                st->invalid_steps++;
                if (st->lambda == st->max_lambda) st->done = 1;
                else {
                    st->lambda = fminf(st->max_lambda, st->lambda * st->v);
                    st->v = 2.f * st->v;
                    st->rebuild = 0;
                    st->skip = 1;
                    st->iterations++;                     // the loop's `continue`
                    if (st->iterations >= st->max_iterations) st->done = 1;
                }
            }
        }
        return;
    }
    // ---- backward substitution: x_i = L_ii^-T (y_i - sum_k L_(i+k,i)^T x_(i+k)) --------------------------------
    auto stage_blocks = [&](int i, int buf) {
        for (int idx = tid; idx < NB * PP; idx += nt) {
            const int k = idx / PP, e = idx - k * PP;
            if (k == 0) S.Lb[buf][0][e] = Lg[((size_t)i * NB) * PP + e];
            else if (i + k < nf) S.Lb[buf][k][e] = Lg[((size_t)(i + k) * NB + k) * PP + e];   // block (i+k, i)
        }
    };
    stage_blocks(nf - 1, (nf - 1) & 1);
    __syncthreads();
    float* x = v.step;
    for (int i = nf - 1; i >= 0; i--) {
        const int buf = i & 1;
        if (i > 0) stage_blocks(i - 1, buf ^ 1);          // next step's blocks stream in meanwhile
        const int kmax = min(NB - 1, nf - 1 - i);
        if (tid < kmax * P) {
            const int k = tid / P + 1, a = tid % P;
            const float* B = S.Lb[buf][k];
            const float* xk = S.xs[(i + k) % NB];
            float s = 0.f;
#pragma unroll
            for (int q = 0; q < P; q++) s += B[q * P + a] * xk[q];
            S.part[k][a] = s;
        }
        __syncthreads();
        if (tid == 0) {
            const float* Dm = S.Lb[buf][0];
            float* xi = S.xs[i % NB];
            float s[P];
#pragma unroll
            for (int r = 0; r < P; r++) {
                float a = y[i * P + r];
                for (int k = 1; k <= kmax; k++) a -= S.part[k][r];
                s[r] = a;
            }
#pragma unroll
            for (int r = P - 1; r >= 0; r--) {
                float a = s[r];
#pragma unroll
                for (int j = r + 1; j < P; j++) a -= Dm[j * P + r] * s[j];
                s[r] = a / Dm[r * P + r];
            }
#pragma unroll
            for (int r = 0; r < P; r++) { xi[r] = s[r]; x[i * P + r] = -s[r]; }   // step = -solve (lev_marq.h:838)
        }
        __syncthreads();
    }
    double a = 0.0;
    for (int k = tid; k < nf * P; k += nt) a += (double)x[k] * (double)x[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((tid & 31) == 0) S.red[tid >> 5] = a;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int k = 0; k < nt / 32; k++) s += S.red[k];
        const float sn = (float)sqrt(s);
        v.scalars[2] = sn;
        v.scalars[4] = 1.f;
        if (st != nullptr) {
            st->llt_ok = 1;
            st->step_norm = sn;
            if (sn < st->step_tol) st->done = 1;          // lev_marq.h:523-526
        }
    }
}

void launch_ba_solve(const BAView& v, BALmState* state, float lambda, cudaStream_t s) {
    if (v.p == 6) {
        static bool once6 = false;
        if (!once6) { cudaFuncSetAttribute(ba_solve_window_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SolveSmem<6>)); once6 = true; }
        ba_solve_window_kernel<6><<<1, SOLVE_THREADS, sizeof(SolveSmem<6>), s>>>(v, state, lambda);
    } else {
        static bool once9 = false;
        if (!once9) { cudaFuncSetAttribute(ba_solve_window_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SolveSmem<9>)); once9 = true; }
        ba_solve_window_kernel<9><<<1, SOLVE_THREADS, sizeof(SolveSmem<9>), s>>>(v, state, lambda);
    }
}

// ---- parameter step + the expected cost change -----------------------------------------------------------
// params_new = Step(params, step) for the interior cameras (refiner.cc:618-646: first and last are constant), and
// expected_cost_change = step^T (2 Jtr + JtJ step) with the clamped, undamped diagonal (lev_marq.h:541-545);
// one partial sum per camera row, added up by the decide kernel.
__global__ void __launch_bounds__(128) ba_step_kernel(BAView v, BALmState* st, const pc_camera_state* params,
                                                      pc_camera_state* params_new, Bounds bounds, float* expected_part) {
    if (st->done || st->skip) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v.nf) return;
    const int p = v.p, pp = p * p, nf = v.nf;
    if (i == 0 || i == nf - 1) params_new[i] = params[i];
    else {
        float dp[9];
        for (int a = 0; a < p; a++) dp[a] = v.step[i * p + a];
        for (int a = p; a < 9; a++) dp[a] = 0.f;
        pc_camera_state out;
        camera_step(params[i], dp, v.opt_f != 0, v.opt_pp != 0, bounds, out);
        params_new[i] = out;
    }
    double acc = 0.0;
    for (int a = 0; a < p; a++) {
        float s = 0.f;
        for (int k = 0; k < kBandBlocks && i - k >= 0; k++) {          // blocks (i, i-k)
            const float* B = v.band + ((size_t)i * kBandBlocks + k) * pp + a * p;
            const float* sv = v.step + (i - k) * p;
            for (int q = 0; q < p; q++) {
                const float m = (k == 0 && q == a) ? v.diag[i * p + a] : B[q];
                s += m * sv[q];
            }
        }
        for (int k = 1; k < kBandBlocks && i + k < nf; k++) {           // blocks (i+k, i)^T
            const float* B = v.band + ((size_t)(i + k) * kBandBlocks + k) * pp;
            const float* sv = v.step + (i + k) * p;
            for (int q = 0; q < p; q++) s += B[q * p + a] * sv[q];
        }
        acc += (double)v.step[i * p + a] * (double)(2.f * v.jtr[i * p + a] + s);
    }
    expected_part[i] = (float)acc;
}

void launch_ba_step(const BAView& v, BALmState* st, const pc_camera_state* params, pc_camera_state* params_new,
                    const Bounds& bounds, float* expected_part, cudaStream_t s) {
    ba_step_kernel<<<(v.nf + 127) / 128, 128, 0, s>>>(v, st, params, params_new, bounds, expected_part);
}

// ---- accept / reject (lev_marq.h:527-580) ----------------------------------------------------------------
__global__ void __launch_bounds__(256) ba_decide_kernel(BAView v, BALmState* st, pc_camera_state* params,
                                                        const pc_camera_state* params_new, const float* expected_part,
                                                        const float* cost_new_ptr) {
    __shared__ double red[8];
    __shared__ int s_accept;
    if (st->done || st->skip) {
        if (threadIdx.x == 0) st->snap_valid = 0;
        return;
    }
    double a = 0.0;
    for (int i = threadIdx.x; i < v.nf; i += blockDim.x) a += (double)expected_part[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double e = 0.0;
        for (int k = 0; k < 8; k++) e += red[k];
        const float cost_new = *cost_new_ptr;
        st->cost_new = cost_new;
        if (cost_new < st->cost) {
            const float actual = cost_new - st->cost;
            const float expected = (float)e;
            st->expected = expected;
            const float rho = actual / expected;
            if (rho > 0.f) {
                const float factor = (float)fmax(1.0 / 3.0, 1.0 - pow(2.0 * (double)rho - 1.0, 3.0));   // `const Float factor`
                st->lambda = fminf(fmaxf(st->lambda * factor, st->min_lambda), st->max_lambda);
            }
            st->cost = cost_new;
            st->v = 2.f;
            st->rebuild = 1;
            s_accept = 1;
        } else {
            st->invalid_steps++;
            s_accept = 0;
            if (st->lambda == st->max_lambda) st->done = 1;
            else {
                st->lambda = fminf(st->max_lambda, st->lambda * st->v);
                st->v = 2.f * st->v;
                st->rebuild = 0;
            }
        }
        if (!st->done) {
            // what the reference hands to the iteration callback (lev_marq.h:576-580), then ++iterations
            st->snap.iterations = st->iterations;
            st->snap.initial_cost = st->initial_cost;
            st->snap.cost = st->cost;
            st->snap.lambda = st->lambda;
            st->snap.invalid_steps = st->invalid_steps;
            st->snap.step_norm = st->step_norm;
            st->snap.grad_norm = st->grad_norm;
            st->snap_valid = 1;
            st->iterations++;
            if (st->iterations >= st->max_iterations) st->done = 1;
        } else {
            st->snap_valid = 0;
        }
    }
    __syncthreads();
    if (s_accept)                                              // *params = params_new
        for (int k = threadIdx.x; k < v.nf * 16; k += blockDim.x)
            reinterpret_cast<float*>(params)[k] = reinterpret_cast<const float*>(params_new)[k];
}

void launch_ba_decide(const BAView& v, BALmState* st, pc_camera_state* params, const pc_camera_state* params_new,
                      const float* expected_part, const float* cost_new, cudaStream_t s) {
    ba_decide_kernel<<<1, 256, 0, s>>>(v, st, params, params_new, expected_part, cost_new);
}

}  // namespace pc
