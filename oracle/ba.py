"""float32 restatement of the reference's trajectory refiner -- TEST INFRASTRUCTURE.

PARITY UNPINNED: no reference test or golden vector exists for the refiner and it cannot be compiled here
(Eigen, Embree absent); checked by Jacobian finite differences and ground-truth recovery only.

  RefinementProblemBase::Evaluate / EvaluateWithJacobian   /root/reference/cpp/refiner.cc:274-506
  GlobalRefinementProblem (edges, weights, Step)            /root/reference/cpp/refiner.cc:250-257,578-647
  LevMarqSparseSolver (normal equations, cost, LM loop)     /root/reference/cpp/pnp/lev_marq.h:391-871
  ray/plane and ray/triangle intersections                  /root/reference/cpp/ray_casting.h:76-179
The sparse SimplicialLLT is replaced by a dense Cholesky of the same matrix (equal up to
rounding).  Vectorised over all residuals; per-edge and global sums are sequential float32
like the reference with one thread."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import pnp, raycast
from .geometry import F, OPENCV, quat_step_post

INVALID = np.uint32(0xFFFFFFFF)


@dataclass
class Edge:
    src: int                     # frame index in the segment
    tgt: int
    src_kps_indices: np.ndarray  # (n,) indices into the src frame's (filtered) keypoints
    tgt_kps: np.ndarray          # (n,2)


class RefineProblem:
    def __init__(self, keypoints, edges, verts, tris, mask_bits, model, opt_f=False, opt_pp=False, bounds=None):
        self.kps = [np.asarray(k, F).reshape(-1, 2) for k in keypoints]
        self.nf = len(self.kps)
        self.edges = edges
        self.verts = np.asarray(verts, F)
        self.tris = np.asarray(tris, np.int64)
        self.mask_bits = None if mask_bits is None else np.asarray(mask_bits, np.uint32)
        self.M = np.asarray(model, F).reshape(4, 4)
        self.Minv = np.linalg.inv(self.M.astype(np.float64)).astype(F)
        self.opt_f, self.opt_pp = bool(opt_f), bool(opt_pp)
        self.p = 9 if (opt_f or opt_pp) else 6                  # refiner.cc:229-233
        self.bounds = bounds
        self.cache = [np.full(len(k), INVALID, np.uint32) for k in self.kps]   # refiner.cc:235-242
        self.offs = np.concatenate([[0], np.cumsum([len(k) for k in self.kps])]).astype(np.int64)
        # flattened residual table
        self.e_id = np.concatenate([np.full(len(e.src_kps_indices), i, np.int64) for i, e in enumerate(edges)])
        self.r_src = np.concatenate([np.full(len(e.src_kps_indices), e.src, np.int64) for e in edges])
        self.r_tgt = np.concatenate([np.full(len(e.src_kps_indices), e.tgt, np.int64) for e in edges])
        self.r_kp = np.concatenate([np.asarray(e.src_kps_indices, np.int64) for e in edges])
        self.r_tgt_pt = np.concatenate([np.asarray(e.tgt_kps, F).reshape(-1, 2) for e in edges]).astype(F)
        self.r_gkp = self.offs[self.r_src] + self.r_kp
        self.referenced = np.zeros(self.offs[-1], bool)
        self.referenced[self.r_gkp] = True
        self.kp_frame = np.concatenate([np.full(len(k), i, np.int64) for i, k in enumerate(self.kps)])
        self.all_kps = np.concatenate(self.kps).astype(F)

    def frame_weight(self, idx):                                 # refiner.cc:250-257
        d = min(idx, self.nf - 1 - idx)
        return F(1.0) / (F(d) + F(1.0))

    def edge_weight(self, e):                                    # refiner.cc:598-601
        return self.frame_weight(self.edges[e].src)

    def is_gt(self, idx):                                        # refiner.cc:268-271
        return idx == 0 or idx == self.nf - 1

    # -- rays of every referenced keypoint (they depend on the source camera only) -------
    def _rays(self, traj):
        g = np.nonzero(self.referenced)[0]
        fr = self.kp_frame[g]
        o_w = np.zeros((len(g), 3), F)
        d_w = np.zeros((len(g), 3), F)
        for f in np.unique(fr):
            cam = traj[f]
            sel = fr == f
            dc = cam.intrinsics.unproject(self.all_kps[g[sel]])                 # refiner.cc:306-307
            R = cam.pose.R()
            o_w[sel] = cam.pose.center()                                        # :309
            d_w[sel] = (dc @ R).astype(F)                                        # Derotate = R^T d  (:310)
        return g, o_w, d_w

    def _refresh_points(self, traj):
        """First half of Evaluate (refiner.cc:306-355): cached-triangle test, ray-cast fallback,
        cache update.  Returns object-space points per global keypoint + validity."""
        g, o_w, d_w = self._rays(traj)
        Mi = self.Minv
        oh = (np.concatenate([o_w, np.ones((len(g), 1), F)], 1) @ Mi.T).astype(F)
        o_o = (oh[:, :3] / oh[:, 3:4]).astype(F)                                # hnormalized (:315-316)
        d_o = (d_w @ Mi[:3, :3].T).astype(F)                                    # :317
        pts = np.zeros((self.offs[-1], 3), F)
        valid = np.zeros(self.offs[-1], bool)
        flat_cache = np.concatenate(self.cache)
        prim = flat_cache[g]
        found = np.zeros(len(g), bool)
        have = prim != INVALID
        if have.any():                                                          # :326-334
            t = self.tris[prim[have].astype(np.int64)]
            p1, p2, p3 = self.verts[t[:, 0]], self.verts[t[:, 1]], self.verts[t[:, 2]]
            hit, P = _mt_single(o_o[have], d_o[have], p1, p2, p3)
            idx = np.nonzero(have)[0]
            found[idx[hit]] = True
            pts[g[idx[hit]]] = P[hit]
        need = ~found
        if need.any():                                                          # :336-346
            h, P, pr, _, _ = raycast.ray_cast(self.verts, self.tris, self.mask_bits, o_o[need], d_o[need], True)
            idx = np.nonzero(need)[0]
            flat_cache[g[idx]] = np.where(h, pr, INVALID)                       # :341,349
            pts[g[idx[h]]] = P[h]
            found[idx[h]] = True
        valid[g] = found
        for f in range(self.nf):
            self.cache[f] = flat_cache[self.offs[f]:self.offs[f + 1]].copy()
        return pts, valid

    def residuals(self, traj):
        """Evaluate for every residual: (res (n,2), valid (n,))."""
        pts, valid = self._refresh_points(traj)
        Po = pts[self.r_gkp]
        ok = valid[self.r_gkp].copy()
        M = self.M
        Ph = (np.concatenate([Po, np.ones((len(Po), 1), F)], 1) @ M.T).astype(F)
        with np.errstate(divide="ignore", invalid="ignore"):
            Pw = (Ph[:, :3] / Ph[:, 3:4]).astype(F)                             # :352-353
        res = np.zeros((len(Po), 2), F)
        for f in np.unique(self.r_tgt):
            sel = self.r_tgt == f
            cam = traj[f]
            Pc = cam.pose.apply(Pw[sel])                                        # :354
            beh = cam.intrinsics.is_behind(Pc)                                  # :356-358
            with np.errstate(divide="ignore", invalid="ignore"):
                res[sel] = cam.intrinsics.project(Pc) - self.r_tgt_pt[sel]
            ok[np.nonzero(sel)[0][beh]] = False
        return res, ok

    def total_cost(self, traj, loss):                            # lev_marq.h:773-824
        res, ok = self.residuals(traj)
        r2 = (res[:, 0] * res[:, 0] + res[:, 1] * res[:, 1]).astype(F)
        l = loss.loss(r2)
        terms = np.zeros(len(self.edges), F)
        for e in range(len(self.edges)):
            ew = self.edge_weight(e)
            sel = (self.e_id == e) & ok
            n = int(sel.sum())
            ec = pnp.seq_sum(l[sel]) if n else F(0.0)
            if n:
                ec = F(ec / F(n))
            terms[e] = F(ew * ec)
        return F(pnp.seq_sum(terms))

    def residuals_jac(self, traj):
        """EvaluateWithJacobian for every residual (refiner.cc:363-506).
        Returns res (n,2), J_src (n,2,p), J_tgt (n,2,p), ok (n,)."""
        n = len(self.r_gkp)
        p = self.p
        flat_cache = np.concatenate(self.cache)
        prim = flat_cache[self.r_gkp]
        ok = prim != INVALID                                                     # :391-393
        prim_safe = np.where(ok, prim, 0).astype(np.int64)
        t = self.tris[prim_safe]
        p1, p2, p3 = self.verts[t[:, 0]], self.verts[t[:, 1]], self.verts[t[:, 2]]
        M, Mi = self.M, self.Minv
        plane_pt = ((np.concatenate([p1, np.ones((n, 1), F)], 1) @ M.T)[:, :3]).astype(F)        # :422-423
        nrm_o = np.cross((p2 - p1).astype(F), (p3 - p1).astype(F)).astype(F)
        plane_n = (nrm_o @ Mi[:3, :3]).astype(F)                                 # Minv^T(3x3) * n  (:424-428)
        res = np.zeros((n, 2), F)
        Js = np.zeros((n, 2, p), F)
        Jt = np.zeros((n, 2, p), F)
        src_pt = self.all_kps[self.r_gkp]
        I3 = np.eye(3, dtype=F)
        for fs in np.unique(self.r_src):
            cs = traj[fs]
            selS = self.r_src == fs
            it = cs.intrinsics
            s = F(1.0) if it.convention == OPENCV else F(-1.0)
            Rs = cs.pose.R()
            origin = cs.pose.center()
            dO_dR = _skew(origin)                                                # pose.h:53-55
            dO_dt = (-Rs.T).astype(F)
            for ft in np.unique(self.r_tgt[selS]):
                ct = traj[ft]
                sel = np.nonzero(selS & (self.r_tgt == ft) & ok)[0]
                if len(sel) == 0:
                    continue
                x = src_pt[sel]
                dirCam = it.unproject(x)                                         # types.h:100-125
                dI = np.zeros((len(sel), 3, 3), F)
                dI[:, 0, 0] = s * (it.cx - x[:, 0]) / (it.fy * it.fy * it.aspect_ratio)
                dI[:, 0, 1] = -s / it.fx
                dI[:, 1, 0] = s * (it.cy - x[:, 1]) / (it.fy * it.fy)
                dI[:, 1, 2] = -s / it.fy
                dirW = (dirCam @ Rs).astype(F)                                   # R^T d
                dDirW_dR = _skew_batch(dirW)
                nvec = plane_n[sel]
                d_dot_n = np.sum(dirW * nvec, -1, dtype=F)                       # ray_casting.h:90
                p0 = np.sum((plane_pt[sel] - origin).astype(F) * nvec, -1, dtype=F)
                with np.errstate(divide="ignore", invalid="ignore"):
                    tt = (p0.astype(np.float64) / d_dot_n.astype(np.float64)).astype(F)
                    X = (origin + dirW * tt[:, None]).astype(F)                  # :99
                    outer = (dirW[:, :, None] * nvec[:, None, :]).astype(F)
                    dX_dO = (I3 - outer / d_dot_n[:, None, None]).astype(F)      # :102-104
                    dX_dD = (dX_dO * tt[:, None, None]).astype(F)                # :106-109
                Rt = ct.pose.R()
                XCam = (X @ Rt.T + ct.pose.t).astype(F)
                beh = ct.intrinsics.is_behind(XCam)                              # refiner.cc:444-446
                itt = ct.intrinsics
                with np.errstate(divide="ignore", invalid="ignore"):
                    pp = itt.project(XCam)
                    x0, x1, x2 = XCam[:, 0], XCam[:, 1], XCam[:, 2]
                    dp = np.zeros((len(sel), 2, 3), F)
                    dp[:, 0, 0] = itt.fx / x2
                    dp[:, 0, 2] = -itt.fx * x0 / (x2 * x2)
                    dp[:, 1, 1] = itt.fy / x2
                    dp[:, 1, 2] = -itt.fy * x1 / (x2 * x2)
                    dpi = np.zeros((len(sel), 2, 3), F)
                    dpi[:, 0, 0] = itt.aspect_ratio * x0 / x2
                    dpi[:, 0, 1] = 1.0
                    dpi[:, 1, 0] = x1 / x2
                    dpi[:, 1, 2] = 1.0
                res[sel] = (pp - self.r_tgt_pt[sel]).astype(F)
                dp_dX = np.einsum("nij,jk->nik", dp, Rt).astype(F)               # :456
                if not self.is_gt(fs):                                           # refiner.cc:611
                    A = (np.einsum("nij,jk->nik", dX_dO, dO_dR)
                         + np.einsum("nij,njk->nik", dX_dD, dDirW_dR)).astype(F)
                    Js[sel, :, 0:3] = np.einsum("nij,njk->nik", dp_dX, A).astype(F)
                    Js[sel, :, 3:6] = np.einsum("nij,njk,kl->nil", dp_dX, dX_dO, dO_dt).astype(F)
                    if p == 9:
                        B = np.einsum("nij,njk,kl,nlm->nim", dp_dX, dX_dD, Rs.T.astype(F), dI).astype(F)
                        if not self.opt_f:
                            B[:, :, 0] = 0
                        if not self.opt_pp:
                            B[:, :, 1:3] = 0
                        Js[sel, :, 6:9] = B
                if not self.is_gt(ft):                                           # refiner.cc:612
                    dXCam_dR = np.einsum("ij,njk->nik", Rt, _skew_batch(-X)).astype(F)
                    Jt[sel, :, 0:3] = np.einsum("nij,njk->nik", dp, dXCam_dR).astype(F)
                    Jt[sel, :, 3:6] = dp
                    if p == 9:
                        B = dpi.copy()
                        if not self.opt_f:
                            B[:, :, 0] = 0
                        if not self.opt_pp:
                            B[:, :, 1:3] = 0
                        Jt[sel, :, 6:9] = B
                ok[sel[beh]] = False
        return res, Js, Jt, ok

    def normal_equations(self, traj, loss):                      # lev_marq.h:653-771
        p, nf = self.p, self.nf
        res, Js, Jt, ok = self.residuals_jac(traj)
        r2 = (res[:, 0] * res[:, 0] + res[:, 1] * res[:, 1]).astype(F)
        lw = loss.weight(r2)
        acc = pnp._SUM["dtype"]                                  # float32 like the reference unless a test asks for exact sums
        A = np.zeros((nf * p, nf * p), acc)
        g = np.zeros(nf * p, acc)
        for e, ed in enumerate(self.edges):
            ew = self.edge_weight(e)
            if ew == 0:
                continue
            sel = np.nonzero((self.e_id == e) & ok)[0]
            if len(sel):
                J = np.concatenate([Js[sel], Jt[sel]], axis=2)                   # (n,2,2p)
                wgt = (ew * F(1.0) * lw[sel]).astype(F)
                JtJ = pnp.seq_sum((np.einsum("nri,nrj->nij", J, J).astype(F) * wgt[:, None, None]).astype(F))
                Jtr = pnp.seq_sum(np.einsum("nri,nr->ni", J, (wgt[:, None] * res[sel]).astype(F)).astype(F))
                JtJ = (JtJ / F(len(sel))).astype(F)
                Jtr = (Jtr / F(len(sel))).astype(F)
            else:
                JtJ = np.zeros((2 * p, 2 * p), F)
                Jtr = np.zeros(2 * p, F)
            b1, b2 = ed.src * p, ed.tgt * p
            A[b1:b1 + p, b1:b1 + p] += np.tril(JtJ[:p, :p])
            A[b2:b2 + p, b2:b2 + p] += np.tril(JtJ[p:, p:])
            if b1 > b2:
                A[b1:b1 + p, b2:b2 + p] += JtJ[:p, p:]
            else:
                A[b2:b2 + p, b1:b1 + p] += JtJ[p:, :p]
            g[b1:b1 + p] += Jtr[:p]
            g[b2:b2 + p] += Jtr[p:]
        return A.astype(F), g.astype(F)

    def step(self, traj, dp):                                    # refiner.cc:508-537,618-646
        out = [c.copy() for c in traj]
        p = self.p
        b = self.bounds
        for f in range(1, self.nf - 1):
            d = np.asarray(dp[f * p:(f + 1) * p], F)
            c = out[f]
            c.pose.q = quat_step_post(traj[f].pose.q, d[0:3])
            c.pose.t = (traj[f].pose.t + d[3:6]).astype(F)
            it = c.intrinsics
            if self.opt_f:
                it.fy = F(it.fy + d[6])
                it.fx = F(it.fy * it.aspect_ratio)
                it.fy = F(np.clip(it.fy, b["f_low"], b["f_high"]))
                it.fx = F(np.clip(it.fx, b["f_low"], b["f_high"]))
            if self.opt_pp:
                it.cx = F(np.clip(F(it.cx + d[7]), b["cx_low"], b["cx_high"]))
                it.cy = F(np.clip(F(it.cy + d[8]), b["cy_low"], b["cy_high"]))
        return out


def _skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]], F)


def _skew_batch(v):
    n = len(v)
    s = np.zeros((n, 3, 3), F)
    s[:, 0, 1], s[:, 0, 2] = -v[:, 2], v[:, 1]
    s[:, 1, 0], s[:, 1, 2] = v[:, 2], -v[:, 0]
    s[:, 2, 0], s[:, 2, 1] = -v[:, 1], v[:, 0]
    return s


def _mt_single(o, d, p1, p2, p3):
    """Row-wise Moller-Trumbore (ray i vs triangle i), ray_casting.h:125-179."""
    eps = F(1e-10)
    e1, e2 = (p2 - p1).astype(F), (p3 - p1).astype(F)
    rxe2 = np.cross(d, e2).astype(F)
    det = np.sum(e1 * rxe2, -1, dtype=F)
    ok = ~((det > -eps) & (det < eps))
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = (F(1.0) / det).astype(F)
        s = (o - p1).astype(F)
        u = (inv * np.sum(s * rxe2, -1, dtype=F)).astype(F)
        ok &= ~((u < 0) | (u > 1))
        sxe1 = np.cross(s, e1).astype(F)
        v = (inv * np.sum(d * sxe1, -1, dtype=F)).astype(F)
        ok &= ~((v < 0) | (u + v > 1))
        t = (inv * np.sum(e2 * sxe1, -1, dtype=F)).astype(F)
        ok &= ~(t < 0)
        P = (o + d * t[:, None]).astype(F)
    return ok, P


def refine_trajectory(problem: RefineProblem, traj, opts: pnp.BundleOptions, callback=None):
    """RefineTrajectory + LevMarqSparseSolve (refiner.cc:649-690, lev_marq.h:492-588)."""
    assert len(traj) > 2                                         # refiner.cc:661
    loss = pnp.Loss(opts.loss_type, opts.loss_scale)
    if problem.bounds is None:
        problem.bounds = traj[0].intrinsics.bounds()             # refiner.cc:687

    def cost_fn(t):
        return problem.total_cost(t, loss)

    def build_fn(t):
        A, g = problem.normal_equations(t, loss)
        diag = np.minimum(np.maximum(np.diag(A), F(1e-6)), F(1e32)).astype(F)   # lev_marq.h:770

        def mul(s, A=A, d=diag):
            full = (A + A.T - np.diag(np.diag(A))).astype(F)
            np.fill_diagonal(full, d)
            return (full @ s).astype(F)
        return {"A": A, "mul": mul}, g, diag

    def solve_fn(handle, diag, lam, Jtr):                        # lev_marq.h:826-841
        A = handle["A"].copy()
        np.fill_diagonal(A, (diag * F(1.0 + np.float64(lam))).astype(F))
        full = (np.tril(A) + np.tril(A, -1).T).astype(F)
        try:
            L = np.linalg.cholesky(full)
        except np.linalg.LinAlgError:
            return None
        y = np.linalg.solve(L.astype(np.float64), Jtr.astype(np.float64))
        x = np.linalg.solve(L.T.astype(np.float64), y)
        return (-x).astype(F)

    return pnp.lm_state_machine(cost_fn, build_fn, solve_fn, problem.step, [c.copy() for c in traj], opts, callback)
