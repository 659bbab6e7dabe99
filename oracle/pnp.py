"""float32 restatement of the reference's per-frame pose solve -- TEST INFRASTRUCTURE.

PARITY UNPINNED: the reference ships no test or golden vector for this path and cannot be compiled
here (Eigen is absent); checked only by analytic-vs-numeric Jacobians, ground-truth recovery and the one
known-answer LLT example of cpp/examples/levmarq_ill_conditioned_float32_issue.cpp, whose recorded digits
no restated Eigen build variant reproduces (oracle/eigen_llt.py, tests/test_oracle_solvers.py).

  robust losses        /root/reference/cpp/pnp/robust_loss.h:47-104
  PnPProblem           /root/reference/cpp/pnp/pnp_problem.h:13-142
  LevMarqDenseSolver   /root/reference/cpp/pnp/lev_marq.h:99-389 (state machine: SURVEY App. B)
  SolvePnPIterative    /root/reference/cpp/pnp/solvers.cc:11-78
Sums are formed sequentially in residual order in float32 (what the reference does with
max_allowed_parallelism = 1)."""
from __future__ import annotations

from contextlib import contextmanager
from dataclasses import dataclass

import numpy as np

from .geometry import F, CameraState, Pose, quat_step_post, skew

TRIVIAL, HUBER, CAUCHY = 0, 1, 2
FLT_MIN = np.finfo(np.float32).tiny
FLT_MAX = np.finfo(np.float32).max


@dataclass
class BundleOptions:                        # types.h:200-215
    max_iterations: int = 100
    loss_type: int = HUBER
    loss_scale: float = 1.0
    gradient_tol: float = 1e-10
    step_tol: float = 1e-8
    initial_lambda: float = 1e-5
    min_lambda: float = 1e-10
    max_lambda: float = 1e10


@dataclass
class BundleStats:                          # types.h:217-225
    iterations: int = 0
    initial_cost: float = 0.0
    cost: float = 0.0
    lambda_: float = 0.0
    invalid_steps: int = 0
    step_norm: float = -1.0
    grad_norm: float = -1.0


class Loss:
    def __init__(self, kind: int, scale: float):
        self.kind = kind
        self.thr = F(scale)
        self.sq_thr = F(self.thr * self.thr)
        self.inv_sq_thr = F(1.0 / np.float64(self.sq_thr))

    def loss(self, r2):
        r2 = np.asarray(r2, F)
        with np.errstate(over="ignore", invalid="ignore"):
            if self.kind == TRIVIAL:
                return r2
            if self.kind == HUBER:
                r = np.sqrt(r2)
                big = (self.thr.astype(np.float64) * (2.0 * r.astype(np.float64) - self.thr)).astype(F)
                return np.where(r2 <= self.sq_thr, r2, big).astype(F)
            return (self.sq_thr * np.log1p(r2 * self.inv_sq_thr)).astype(F)

    def weight(self, r2):
        r2 = np.asarray(r2, F)
        with np.errstate(divide="ignore", invalid="ignore"):
            if self.kind == TRIVIAL:
                return np.ones_like(r2)
            if self.kind == HUBER:
                return np.where(r2 <= self.sq_thr, F(1.0), self.thr / np.sqrt(r2)).astype(F)
            return np.maximum(FLT_MIN, F(1.0) / (F(1.0) + r2 * self.inv_sq_thr)).astype(F)


_SUM = {"dtype": F, "rng": None}


@contextmanager
def summation(dtype=F, perm_seed=None):
    """How seq_sum accumulates inside the block.  Default: sequential float32 in residual order (the
    reference with max_allowed_parallelism = 1).  dtype=np.float64: correctly rounded sums (the
    order-independent value every float32 order scatters around).  perm_seed: sequential float32
    over a random permutation of the terms -- one sample of the reference's own nondeterminism (its
    TBB reductions combine thread-local partial sums in scheduling order, lev_marq.h:231-297,
    653-771); tests use a few seeds to measure that noise band."""
    old = dict(_SUM)
    _SUM["dtype"] = dtype
    _SUM["rng"] = None if perm_seed is None else np.random.default_rng(perm_seed)
    try:
        yield
    finally:
        _SUM.update(old)


def seq_sum(a, axis=0):
    """Accumulation along axis 0: sequential float32 (TBB with one thread) unless `summation` says otherwise."""
    assert axis == 0
    a = np.asarray(a, F)
    if not a.shape[0]:
        return np.zeros(a.shape[1:], F)
    if _SUM["rng"] is not None:
        a = a[_SUM["rng"].permutation(a.shape[0])]
    if _SUM["dtype"] is np.float64:
        return a.sum(axis=0, dtype=np.float64).astype(F)
    return np.add.accumulate(a, axis=0, dtype=F)[-1]


def llt_lower(A):
    """Eigen::LLT<Lower> on a float32 row-major matrix (lower triangle referenced), in Eigen 3.4's
    operation order (oracle/eigen_llt.py).  Returns (L, ok)."""
    from . import eigen_llt
    return eigen_llt.llt_lower(A)


def llt_solve(L, b):
    from . import eigen_llt
    return eigen_llt.llt_solve(L, b)


class PnPProblem:
    def __init__(self, x, X, weights, opt_f, opt_pp, bounds):
        self.x = np.asarray(x, F).reshape(-1, 2)
        self.X = np.asarray(X, F).reshape(-1, 3)
        self.w = None if weights is None or len(weights) == 0 else np.asarray(weights, F)
        m = len(self.x)
        self.opt_f = bool(opt_f) and m > 3          # pnp_problem.h:33-34
        self.opt_pp = bool(opt_pp) and m > 3
        self.bounds = bounds

    def residuals(self, cam: CameraState):          # Evaluate, pnp_problem.h:52-61
        Z = cam.pose.apply(self.X)
        behind = cam.intrinsics.is_behind(Z)
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            r = (cam.intrinsics.project(Z) - self.x).astype(F)
        r[behind] = FLT_MAX
        return r

    def residuals_jac(self, cam: CameraState):      # EvaluateWithJacobian, pnp_problem.h:63-99
        intr = cam.intrinsics
        R = cam.pose.R()
        Z = self.X
        RtZ = (Z @ R.T + cam.pose.t).astype(F)
        m = len(Z)
        z = intr.project(RtZ)
        res = (z - self.x).astype(F)
        x0, x1, x2 = RtZ[:, 0], RtZ[:, 1], RtZ[:, 2]
        dz = np.zeros((m, 2, 3), F)                 # types.h:79-85
        dz[:, 0, 0] = intr.fx / x2
        dz[:, 0, 2] = -intr.fx * x0 / (x2 * x2)
        dz[:, 1, 1] = intr.fy / x2
        dz[:, 1, 2] = -intr.fy * x1 / (x2 * x2)
        # dRtZ_dR = R * Skew(-Z)   (pose.h:83-85)
        sk = np.zeros((m, 3, 3), F)
        nz = -Z
        sk[:, 0, 1], sk[:, 0, 2] = -nz[:, 2], nz[:, 1]
        sk[:, 1, 0], sk[:, 1, 2] = nz[:, 2], -nz[:, 0]
        sk[:, 2, 0], sk[:, 2, 1] = -nz[:, 1], nz[:, 0]
        dR = np.einsum("ij,mjk->mik", R, sk).astype(F)
        J = np.zeros((m, 2, 9), F)
        J[:, :, 0:3] = np.einsum("mij,mjk->mik", dz, dR).astype(F)
        J[:, :, 3:6] = dz
        if self.opt_f:
            J[:, 0, 6] = intr.aspect_ratio * x0 / x2    # types.h:88-92
            J[:, 1, 6] = x1 / x2
        if self.opt_pp:
            J[:, 0, 7] = 1.0
            J[:, 1, 8] = 1.0
        return res, J

    def step(self, cam: CameraState, dp):           # pnp_problem.h:101-131
        new = cam.copy()
        dp = np.asarray(dp, F)
        new.pose.q = quat_step_post(cam.pose.q, dp[0:3])
        new.pose.t = (cam.pose.t + dp[3:6]).astype(F)
        b = self.bounds
        it, io = new.intrinsics, cam.intrinsics
        if self.opt_f:
            it.fy = F(io.fy + dp[6])
            it.fx = F(it.fy * it.aspect_ratio)
            it.fy = F(np.clip(it.fy, b["f_low"], b["f_high"]))
            it.fx = F(np.clip(it.fx, b["f_low"], b["f_high"]))
        if self.opt_pp:
            it.cx = F(np.clip(F(io.cx + dp[7]), b["cx_low"], b["cx_high"]))
            it.cy = F(np.clip(F(io.cy + dp[8]), b["cy_low"], b["cy_high"]))
        return new


def lm_state_machine(cost_fn, build_fn, solve_fn, step_fn, params, opts: BundleOptions, callback=None):
    """The LM loop shared by the dense and sparse solvers (lev_marq.h:132-228 / 492-588).
    build_fn(params) -> (JtJ handle, Jtr, diag);  solve_fn(handle, diag, lam, Jtr) -> step or None;
    returns (params, stats)."""
    st = BundleStats()
    st.cost = F(cost_fn(params))
    st.initial_cost = st.cost
    st.lambda_ = F(opts.initial_lambda)
    v = F(2.0)
    rebuild = True
    handle = Jtr = diag = None
    max_lambda, min_lambda = F(opts.max_lambda), F(opts.min_lambda)
    it = 0
    while it < opts.max_iterations:
        if rebuild:
            handle, Jtr, diag = build_fn(params)
            st.grad_norm = F(np.sqrt(F(np.dot(Jtr, Jtr))))
            if st.grad_norm < F(opts.gradient_tol):
                break
        step = solve_fn(handle, diag, st.lambda_, Jtr)
        if step is None:
            st.invalid_steps += 1
            if st.lambda_ == max_lambda:
                break
            st.lambda_ = min(max_lambda, F(st.lambda_ * v))
            v = F(2 * v)
            rebuild = False
            it += 1
            continue
        st.step_norm = F(np.sqrt(F(np.dot(step, step))))
        if st.step_norm < F(opts.step_tol):
            break
        params_new = step_fn(params, step)
        cost_new = F(cost_fn(params_new))
        if cost_new < st.cost:
            actual = F(cost_new - st.cost)
            expected = F(np.dot(step.astype(np.float64), 2.0 * Jtr.astype(np.float64)
                                + handle["mul"](step).astype(np.float64)))
            with np.errstate(divide="ignore", invalid="ignore"):
                rho = F(actual / expected)
            if rho > 0:
                factor = F(max(1.0 / 3.0, 1.0 - (2.0 * float(rho) - 1.0) ** 3))      # `const Float factor`
                st.lambda_ = F(np.clip(F(st.lambda_ * factor), min_lambda, max_lambda))
            params = params_new
            st.cost = cost_new
            v = F(2.0)
            rebuild = True
        else:
            st.invalid_steps += 1
            if st.lambda_ == max_lambda:
                break
            st.lambda_ = min(max_lambda, F(st.lambda_ * v))
            v = F(2 * v)
            rebuild = False
        if callback is not None and not callback(st):
            break
        it += 1
    st.iterations = it
    if callback is not None:
        callback(st)
    return params, st


def solve_pnp_iterative(X, x, weights, cam: CameraState, opts: BundleOptions, max_inlier_error=12.0,
                        optimize_focal_length=False, optimize_principal_point=False):
    """SolvePnPIterative (solvers.cc:11-78).  Returns (camera, stats, inlier_ratio)."""
    assert len(X) == len(x) and len(X) >= 3        # solvers.cc:54-55
    prob = PnPProblem(x, X, weights, optimize_focal_length, optimize_principal_point, cam.intrinsics.bounds())
    loss = Loss(opts.loss_type, opts.loss_scale)
    w = prob.w

    def cost_fn(c):                                  # TotalCost, lev_marq.h:316-356
        r = prob.residuals(c)
        with np.errstate(over="ignore", invalid="ignore"):
            r2 = (r[:, 0] * r[:, 0] + r[:, 1] * r[:, 1]).astype(F)
            l = loss.loss(r2)
        if w is not None:
            l = (w * l).astype(F)[w != 0]
        return seq_sum(l)

    def build_fn(c):                                 # BuildNormalEquations, lev_marq.h:231-297
        r, J = prob.residuals_jac(c)
        r2 = (r[:, 0] * r[:, 0] + r[:, 1] * r[:, 1]).astype(F)
        tw = loss.weight(r2)
        if w is not None:
            tw = (w * tw).astype(F)
        JtJ = seq_sum((tw[:, None, None] * np.einsum("mri,mrj->mij", J, J).astype(F)).astype(F))
        Jtr = seq_sum(np.einsum("mri,mr->mi", J, (tw[:, None] * r).astype(F)).astype(F))
        JtJ = np.tril(JtJ)
        diag = np.minimum(np.maximum(np.diag(JtJ), F(1e-6)), F(1e32)).astype(F)   # lev_marq.h:296

        def mul(s, A=JtJ, d=diag):                   # selfadjointView<Lower>() * step, undamped
            full = A + A.T - np.diag(np.diag(A))
            full = full.copy()
            np.fill_diagonal(full, d)
            return (full @ s).astype(F)
        return {"A": JtJ, "mul": mul}, Jtr, diag

    def solve_fn(handle, diag, lam, Jtr):            # ComputeStep, lev_marq.h:299-314
        A = handle["A"].copy()
        np.fill_diagonal(A, (diag * F(1.0 + np.float64(lam))).astype(F))
        L, ok = llt_lower(A)
        if not ok:
            return None
        return (-llt_solve(L, Jtr)).astype(F)

    cam_out, stats = lm_state_machine(cost_fn, build_fn, solve_fn, prob.step, cam.copy(), opts)
    inlier_ratio = F(0.0)
    if max_inlier_error > 0:                         # solvers.cc:30-47
        r = prob.residuals(cam_out)
        with np.errstate(over="ignore"):
            e2 = (r[:, 0] * r[:, 0] + r[:, 1] * r[:, 1]).astype(F)
        inlier_ratio = F(np.count_nonzero(e2 < F(max_inlier_error) ** 2)) / F(len(r))
    return cam_out, stats, inlier_ratio
