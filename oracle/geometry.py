"""float32 restatement of the reference's camera / pose / quaternion math -- TEST
INFRASTRUCTURE.  Sources: /root/reference/cpp/pnp/types.h:18-198 (CameraIntrinsics),
/root/reference/cpp/pose.h:9-160 (Pose), /root/reference/cpp/pnp/quaternion.h:11-20
(QuatStepPost).  Eigen semantics (toRotationMatrix, matrix->quaternion, q*AngleAxis) follow
SURVEY.md Appendix C.  Quaternions are stored (w, x, y, z) as the Python surface shows them
(polychase_pybind.cc:224-232)."""
from __future__ import annotations

from dataclasses import dataclass, field, replace

import numpy as np

F = np.float32
OPENGL, OPENCV = 0, 1


@dataclass
class Intrinsics:
    fx: float
    fy: float
    cx: float
    cy: float
    aspect_ratio: float
    width: float
    height: float
    convention: int = OPENGL

    def f32(self):
        return replace(self, fx=F(self.fx), fy=F(self.fy), cx=F(self.cx), cy=F(self.cy),
                       aspect_ratio=F(self.aspect_ratio), width=F(self.width), height=F(self.height))

    def project(self, X):                       # types.h:65-67
        X = np.asarray(X, F)
        return np.stack([self.fx * X[..., 0] / X[..., 2] + self.cx,
                         self.fy * X[..., 1] / X[..., 2] + self.cy], axis=-1).astype(F)

    def unproject(self, x):                     # types.h:95-98
        x = np.asarray(x, F)
        s = F(1.0) if self.convention == OPENCV else F(-1.0)
        return (s * np.stack([(x[..., 0] - self.cx) / self.fx, (x[..., 1] - self.cy) / self.fy,
                              np.ones_like(x[..., 0])], axis=-1)).astype(F)

    def is_behind(self, X):                     # types.h:129-132
        X = np.asarray(X, F)
        return X[..., 2] < 0 if self.convention == OPENCV else X[..., 2] > 0

    def bounds(self, min_fov_deg=15.0, max_fov_deg=160.0):   # types.h:156-192
        min_fov = F(F(min_fov_deg) * np.pi / 180)
        max_fov = F(F(max_fov_deg) * np.pi / 180)
        tmin, tmax = F(np.tan(F(min_fov / F(2)))), F(np.tan(F(max_fov / F(2))))
        hw = F(self.width) / F(2.0)
        if self.convention == OPENGL:
            f_low, f_high = F(-hw / tmin), F(-hw / tmax)
        else:
            f_high, f_low = F(hw / tmin), F(hw / tmax)
        return dict(f_low=f_low, f_high=f_high, cx_low=F(0), cx_high=F(self.width), cy_low=F(0),
                    cy_high=F(self.height))


def quat_to_matrix(q):                          # Eigen::Quaternion::toRotationMatrix
    w, x, y, z = [F(v) for v in q]
    tx, ty, tz = F(2) * x, F(2) * y, F(2) * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return np.array([[F(1) - (tyy + tzz), txy - twz, txz + twy],
                     [txy + twz, F(1) - (txx + tzz), tyz - twx],
                     [txz - twy, tyz + twx, F(1) - (txx + tyy)]], F)


def quat_from_matrix(m):                        # Eigen::Quaternion(Matrix3) (Shepperd)
    m = np.asarray(m)
    dt = m.dtype.type
    t = m[0, 0] + m[1, 1] + m[2, 2]
    q = np.zeros(4, m.dtype)                    # w, x, y, z
    if t > 0:
        t = np.sqrt(t + dt(1))
        q[0] = dt(0.5) * t
        t = dt(0.5) / t
        q[1] = (m[2, 1] - m[1, 2]) * t
        q[2] = (m[0, 2] - m[2, 0]) * t
        q[3] = (m[1, 0] - m[0, 1]) * t
    else:
        i = 0
        if m[1, 1] > m[0, 0]:
            i = 1
        if m[2, 2] > m[i, i]:
            i = 2
        j = (i + 1) % 3
        k = (j + 1) % 3
        t = np.sqrt(m[i, i] - m[j, j] - m[k, k] + dt(1))
        q[1 + i] = dt(0.5) * t
        t = dt(0.5) / t
        q[0] = (m[k, j] - m[j, k]) * t
        q[1 + j] = (m[j, i] + m[i, j]) * t
        q[1 + k] = (m[k, i] + m[i, k]) * t
    return q


def quat_mul(a, b):
    aw, ax, ay, az = [F(v) for v in a]
    bw, bx, by, bz = [F(v) for v in b]
    return np.array([aw * bw - ax * bx - ay * by - az * bz,
                     aw * bx + ax * bw + ay * bz - az * by,
                     aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx], F)


def quat_step_post(q, w_delta):                 # quaternion.h:11-20
    w_delta = np.asarray(w_delta, F)
    angle = F(np.sqrt(F(np.dot(w_delta, w_delta))))
    if angle > 0:
        axis = (w_delta / angle).astype(F)
        half = F(0.5) * angle
        aa = np.array([np.cos(half), *(np.sin(half) * axis)], F)
        return quat_mul(q, aa)
    return np.asarray(q, F).copy()


def skew(v):                                    # pose.h:150-158
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]], F)


@dataclass
class Pose:
    q: np.ndarray = field(default_factory=lambda: np.array([1, 0, 0, 0], F))
    t: np.ndarray = field(default_factory=lambda: np.zeros(3, F))

    def R(self):
        return quat_to_matrix(self.q)

    def Rt4x4(self):
        m = np.eye(4, dtype=F)
        m[:3, :3] = self.R()
        m[:3, 3] = self.t
        return m

    def apply(self, p):                          # pose.h:42-44 (R p + t)
        return (np.asarray(p, F) @ self.R().T + self.t).astype(F)

    def center(self):                            # pose.h:46
        return (-(self.R().T @ self.t)).astype(F)

    @staticmethod
    def from_Rt(mat):                            # pose.h:133-136
        mat = np.asarray(mat, F)
        return Pose(quat_from_matrix(mat[:3, :3]).astype(F), mat[:3, 3].astype(F).copy())


@dataclass
class CameraState:
    intrinsics: Intrinsics
    pose: Pose

    def copy(self):
        return CameraState(replace(self.intrinsics), Pose(self.pose.q.copy(), self.pose.t.copy()))
