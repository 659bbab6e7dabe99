"""CPU checks of the Polychase-owned restatements (oracle/pnp.py, ba.py, track.py, geometry.py):
the reference's one known-answer vector, analytic-vs-numeric Jacobians, ground-truth recovery.
These parts have no reference golden vectors ("parity unpinned", oracle/__init__.py)."""
import numpy as np
import pytest

from oracle import ba as oba
from oracle import geometry as G
from oracle import pnp as opnp
from oracle import synth
from tests import helpers as H

F = np.float32


LM_FIXTURE_JTJ = np.array([
    [557551.4375, 0, 0, 0, 0, 0, 0, 0, 0],
    [296441.21875, 657639.8125, 0, 0, 0, 0, 0, 0, 0],
    [-4293.4072265625, -5085.32958984375, 364752.1875, 0, 0, 0, 0, 0, 0],
    [42399.52734375, 131392.296875, 31440.83984375, 70597.6328125, 0, 0, 0, 0, 0],
    [-27725.1328125, 44876.76953125, -105931.8828125, 0.0, 70597.6328125, 0, 0, 0, 0],
    [-43429.875, -83350.875, -62037.90625, -55166.17578125, 25584.125, 52518.796875, 0, 0, 0],
    [1993.02294921875, 3831.88916015625, 2867.069091796875, 2574.660400390625, -1193.505981445312,
     -2450.312255859375, 114.358093261719, 0, 0],
    [1947.6396484375, 6048.806640625, 1454.197631835938, 3295.457763671875, 0.0, -2574.660400390625,
     120.201538085938, 153.880432128906, 0],
    [-1262.969848632812, 2073.468505859375, -4891.965820312500, 0.0, 3295.457763671875, 1193.505981445312,
     -55.693786621094, 0.0, 153.880432128906]], F)
LM_FIXTURE_JTR = np.array([-2.338238716125, -4.207848548889, 3.598472595215, -1.105026721954, -1.491069078445,
                           0.368796110153, -0.017316624522, -0.051174595952, -0.068564474583], F)


def _fixture_outputs(base):
    """The two numbers the reference's example prints, computed in Eigen 3.4's operation order with
    the 9x9 matrix at 16-byte offset `base` (oracle/eigen_llt.py)."""
    from oracle import eigen_llt as E
    lam = F(1.5607382e-06)
    A = (LM_FIXTURE_JTJ + lam * np.eye(9, dtype=F)).astype(F)
    L, ok = E.llt_lower(A, base)
    assert ok
    step = (-E.llt_solve(L, LM_FIXTURE_JTR)).astype(F)
    r = (E.selfadjoint_lower_times(A, step) + LM_FIXTURE_JTR).astype(F)
    residual = F(np.sqrt(E.dot_fixed(r, r)))
    inner = (F(2) * LM_FIXTURE_JTR + E.selfadjoint_lower_times(LM_FIXTURE_JTJ, step)).astype(F)
    return float(residual), float(E.dot_fixed(step, inner)), step


def test_levmarq_known_answer_vector():
    """/root/reference/cpp/examples/levmarq_ill_conditioned_float32_issue.cpp:16-63 -- the reference's
    only numeric fixture: a 9x9 float32 LLT solve with condition number 4.4e10 whose comments record
    residual 0.0028946274 and expected cost change +0.000244110823.

    NOT REPRODUCED.  oracle/eigen_llt.py restates Eigen 3.4.0's operation order (unblocked LLT, panel
    triangular solves, packet reductions) for the build the reference's CMake describes (x86-64, no
    -march flags: SSE2, no FMA); that and 49 other variants (AVX packets, FMA contraction, every
    alignment of the matrix, no vectorisation) all give a residual of 5e-4 .. 1.4e-3 and a NEGATIVE
    expected cost change of -1.1e-4 .. -2.2e-4.  At this conditioning the recorded digits belong to
    one particular binary (compiler, ISA and Eigen version unknown), so the fixture pins the
    factorisation only to the extent asserted here: lambda is below half an ulp of every diagonal
    entry (the damped matrix IS JtJ), the float32 residual is 3 to 4 orders of magnitude above the
    float64 one, and the float32 expected cost change is unreliable (it deviates from the float64
    value -1.87e-4 by 4 .. 40 % depending on the alignment, which is what lev_marq.h:189-197 guards against)."""
    lam = F(1.5607382e-06)
    assert np.array_equal(np.diag(LM_FIXTURE_JTJ) + lam, np.diag(LM_FIXTURE_JTJ))
    full = (LM_FIXTURE_JTJ + LM_FIXTURE_JTJ.T - np.diag(np.diag(LM_FIXTURE_JTJ))).astype(np.float64)
    s64 = -np.linalg.solve(full, LM_FIXTURE_JTR.astype(np.float64))
    e64 = float(s64 @ (2 * LM_FIXTURE_JTR + full @ s64))
    r64 = float(np.linalg.norm(full @ s64 + LM_FIXTURE_JTR))
    assert e64 < 0 and r64 < 1e-6
    assert np.linalg.cond(full) > 1e10
    seen, dev = set(), []
    for base in (0, 4, 8, 12):
        residual, expected, step = _fixture_outputs(base)
        seen.add((residual, expected))
        assert 3e-4 < residual < 5e-3                      # reference records 0.0028946274
        assert abs(expected) < 1e-3                        # reference records +0.000244110823
        dev.append(abs(expected - e64) / abs(e64))
        assert np.abs(step - s64).max() > 1e-2 * np.abs(s64).max()
    assert len(seen) > 1                                   # the outputs depend on where the matrix sits in memory
    assert max(dev) > 0.05                                 # float32 is unreliable here: the documented issue


def test_eigen_ordered_llt_solves_well_conditioned_systems():
    """The Eigen-ordered factorisation / solves (oracle/eigen_llt.py) against float64 on SPD systems
    of the sizes the dense solver uses (6 and 9 parameters)."""
    from oracle import eigen_llt as E
    rng = np.random.default_rng(0)
    for n in (6, 9, 17):
        B = rng.normal(size=(n + 4, n))
        A = (B.T @ B + 0.1 * np.eye(n)).astype(F)
        b = rng.normal(size=n).astype(F)
        L, ok = E.llt_lower(A)
        assert ok
        assert np.allclose(L.astype(np.float64) @ L.astype(np.float64).T, A, rtol=0, atol=1e-5 * np.abs(A).max())
        x = E.llt_solve(L, b)
        assert np.allclose(x, np.linalg.solve(A.astype(np.float64), b), rtol=1e-3, atol=1e-4)
        if n <= 9:
            assert np.allclose(E.selfadjoint_lower_times(np.tril(A), x), A.astype(np.float64) @ x, atol=1e-4)
    L, ok = E.llt_lower(np.array([[1, 0], [2, 1]], F))       # indefinite -> NumericalIssue (lev_marq.h:305-309)
    assert not ok


def test_quaternion_helpers():
    rng = np.random.default_rng(0)
    for _ in range(20):
        R = synth.rot_xyz(*rng.uniform(-3, 3, 3))
        q = G.quat_from_matrix(R.astype(F))
        assert np.allclose(G.quat_to_matrix(q), R, atol=2e-6)
    q = G.quat_step_post(np.array([1, 0, 0, 0], F), np.array([0, 0, np.pi / 2], F))
    assert np.allclose(G.quat_to_matrix(q), synth.rot_xyz(0, 0, np.pi / 2), atol=1e-6)
    assert np.array_equal(G.quat_step_post(q, np.zeros(3, F)), q)      # quaternion.h:17-19


def _numeric_jac(fun, params, eps):
    base = fun(params)
    cols = []
    for k in range(len(params)):
        d = np.zeros(len(params))
        d[k] = eps[k]
        cols.append((fun(params + d) - fun(params - d)) / (2 * eps[k]))
    return base, np.stack(cols, -1)


def test_pnp_jacobian_matches_finite_differences():
    rng = np.random.default_rng(1)
    clip = synth.Clip(320, 240, 2, seed=1)
    cam = H.oracle_cam(clip, 1)
    X = np.stack([rng.uniform(-1, 1, 8), rng.uniform(-1, 1, 8), rng.uniform(-.1, .1, 8)], 1).astype(F)
    x = np.zeros((8, 2), F)
    prob = opnp.PnPProblem(x, X, None, True, True, cam.intrinsics.bounds())
    _, J = prob.residuals_jac(cam)

    def fun(dp):
        c = cam.copy()
        # float64 evaluation of the same parameterisation (right-multiplicative rotation step)
        w = dp[:3]
        ang = np.linalg.norm(w)
        R = cam.pose.R().astype(np.float64)
        if ang > 0:
            k = w / ang
            Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
            R = R @ (np.eye(3) + np.sin(ang) * Kx + (1 - np.cos(ang)) * Kx @ Kx)
        t = cam.pose.t.astype(np.float64) + dp[3:6]
        fy = float(cam.intrinsics.fy) + dp[6]
        fx = fy * float(cam.intrinsics.aspect_ratio)
        cx, cy = float(cam.intrinsics.cx) + dp[7], float(cam.intrinsics.cy) + dp[8]
        Z = X.astype(np.float64) @ R.T + t
        return np.stack([fx * Z[:, 0] / Z[:, 2] + cx, fy * Z[:, 1] / Z[:, 2] + cy], -1)

    _, Jn = _numeric_jac(fun, np.zeros(9), np.array([1e-5] * 3 + [1e-4] * 3 + [1e-2] * 3))
    assert np.allclose(J, Jn, rtol=2e-3, atol=2e-2)


def test_pnp_recovers_ground_truth_all_losses():
    rng = np.random.default_rng(2)
    clip = synth.Clip(640, 480, 3, seed=1)
    gt = H.oracle_cam(clip, 2)
    X = np.stack([rng.uniform(-2, 2, 400), rng.uniform(-1.5, 1.5, 400), rng.uniform(-.1, .1, 400)], 1).astype(F)
    x = gt.intrinsics.project(gt.pose.apply(X))
    for loss in (opnp.TRIVIAL, opnp.HUBER, opnp.CAUCHY):
        cam, st, inl = opnp.solve_pnp_iterative(X, x, None, H.perturb(gt, rng), opnp.BundleOptions(loss_type=loss))
        dq, dt = H.pose_close(gt, cam)
        assert dq < 1e-4 and dt < 1e-4 and inl == 1.0 and st.cost < 1e-3


def test_pnp_behind_camera_points_give_infinite_cost():
    clip = synth.Clip(320, 240, 1, seed=1)
    cam = H.oracle_cam(clip, 0)
    X = np.array([[0, 0, -50.0], [0.1, 0, 0], [0, 0.1, 0], [0.1, 0.1, 0]], F)   # first point behind
    prob = opnp.PnPProblem(np.zeros((4, 2), F), X, None, False, False, cam.intrinsics.bounds())
    r = prob.residuals(cam)
    assert r[0, 0] == np.finfo(np.float32).max                                 # pnp_problem.h:55-58


def _small_ba(seed=0, nf=6, n=60, opt=False):
    rng = np.random.default_rng(seed)
    clip = synth.Clip(320, 240, nf, seed=3)
    verts, tris = H.bumpy_mesh(clip, quads=4, amp=0.03)
    from oracle import raycast
    traj = [H.oracle_cam(clip, k) for k in range(nf)]
    kps = [np.stack([rng.uniform(30, 290, n), rng.uniform(30, 210, n)], 1).astype(F) for _ in range(nf)]
    edges = []
    model = np.eye(4, dtype=F)
    for a in range(nf):
        for d in (-2, -1, 1, 2):
            b = a + d
            if 0 <= b < nf:
                o, dd = raycast.ray_object_space(model, traj[a].pose.Rt4x4(), traj[a].intrinsics, kps[a])
                hit, pos, _, _, _ = raycast.ray_cast(verts, tris, None, o, dd)
                tgt = traj[b].intrinsics.project(traj[b].pose.apply(pos)) + rng.normal(0, 0.2, (n, 2)).astype(F)
                idx = np.nonzero(hit)[0].astype(np.uint32)
                edges.append(oba.Edge(a, b, idx, tgt[hit]))
    prob = oba.RefineProblem(kps, edges, verts, tris, None, model, opt, opt, traj[0].intrinsics.bounds())
    return prob, traj, rng


def test_ba_jacobian_matches_finite_differences():
    prob, traj, rng = _small_ba()
    loss = opnp.Loss(opnp.TRIVIAL, 1.0)
    pert = [traj[0]] + [H.perturb(c, rng, 0.05, 0.004) for c in traj[1:-1]] + [traj[-1]]
    prob.total_cost(pert, loss)                       # fills the primitive cache
    res, Js, Jt, ok = prob.residuals_jac(pert)
    # numeric derivative of residual i wrt the 6 pose parameters of its source / target camera
    i = int(np.nonzero(ok & (prob.r_src > 0) & (prob.r_src < prob.nf - 1)
                       & (prob.r_tgt > 0) & (prob.r_tgt < prob.nf - 1))[0][5])
    for which, Jan in (("src", Js[i]), ("tgt", Jt[i])):
        f = prob.r_src[i] if which == "src" else prob.r_tgt[i]
        cols = []
        for k in range(6):
            out = []
            for sgn in (+1, -1):
                dp = np.zeros(prob.nf * prob.p, F)
                eps = 1e-3 if k < 3 else 2e-3
                dp[f * prob.p + k] = sgn * eps
                t2 = prob.step(pert, dp)
                r2, _, _, _ = prob.residuals_jac(t2)
                out.append(r2[i].astype(np.float64))
            cols.append((out[0] - out[1]) / (2 * eps))
        Jn = np.stack(cols, -1)
        assert np.allclose(Jan[:, :6], Jn, rtol=3e-2, atol=0.5), (which, Jan[:, :6], Jn)


def test_ba_refines_towards_ground_truth():
    prob, traj, rng = _small_ba(seed=1)
    pert = [traj[0]] + [H.perturb(c, rng, 0.05, 0.004) for c in traj[1:-1]] + [traj[-1]]
    out, st = oba.refine_trajectory(prob, pert, opnp.BundleOptions(loss_type=opnp.CAUCHY, max_iterations=25))
    assert st.cost < 0.5 * st.initial_cost
    before = max(H.pose_close(traj[k], pert[k])[1] for k in range(1, len(traj) - 1))
    after = max(H.pose_close(traj[k], out[k])[1] for k in range(1, len(traj) - 1))
    assert after < 0.5 * before
    assert np.array_equal(out[0].pose.t, traj[0].pose.t) and np.array_equal(out[-1].pose.q, traj[-1].pose.q)


def test_ba_cache_semantics():
    """The primitive-id cache is filled by cost evaluations and consulted by the Jacobian pass
    (refiner.cc:323-350,391-393): before any cost evaluation every residual is dropped."""
    prob, traj, _ = _small_ba(seed=2)
    res, Js, Jt, ok = prob.residuals_jac(traj)
    assert not ok.any()
    prob.total_cost(traj, opnp.Loss(opnp.HUBER, 1.0))
    res, Js, Jt, ok = prob.residuals_jac(traj)
    assert ok.mean() > 0.95
