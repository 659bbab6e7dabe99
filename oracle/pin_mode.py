"""float32 restatement of FindTransformationN -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Source: /root/reference/cpp/pin_mode.cc:16-108.  Three and more pins: the object points are taken to the
camera space of the INITIAL scene, projected, the dragged pin's image point is replaced, and
SolvePnPIterative (oracle/pnp.py) with the trivial loss finds the extra rigid motion, started from
current * initial^-1.  Parity unpinned: the reference holds no vector for this function."""
from __future__ import annotations

import numpy as np

from . import pnp
from .geometry import CameraState, F, Intrinsics, Pose

MODEL, CAMERA = 0, 1


def find_transformation_n(object_points, init_model, init_view, init_intr: Intrinsics, cur_model, cur_view,
                          cur_intr: Intrinsics, pin_idx: int, pos, trans_type: int, optimize_focal_length=False,
                          optimize_principal_point=False):
    """Returns (model_matrix, view_matrix, intrinsics) as pin_mode.cc:79-107 builds them."""
    P = np.asarray(object_points, F)
    assert len(P) > 2                                                   # pin_mode.cc:22
    init_model, init_view = np.asarray(init_model, F), np.asarray(init_view, F)
    cur_model, cur_view = np.asarray(cur_model, F), np.asarray(cur_view, F)
    mv = (init_view @ init_model).astype(F)                             # :29-32
    R0, t0 = mv[:3, :3], mv[:3, 3]
    Pc = ((P @ R0.T).astype(F) + t0).astype(F)                          # :34-36
    k = init_intr.f32()
    K3 = np.array([[k.fx, 0, k.cx], [0, k.fy, k.cy], [0, 0, 1]], F)     # To3x3ProjectionMatrix, types.h:69-79
    ip3 = (Pc @ K3.T).astype(F)                                         # :38-39
    x = (ip3[:, :2] / ip3[:, 2:3]).astype(F)                            # :40-44
    x[pin_idx] = np.asarray(pos, F)                                     # :47
    init_pose = ((cur_view @ cur_model).astype(F) @ np.linalg.inv(mv).astype(F)).astype(F)   # :51-54
    cam = CameraState(cur_intr.f32(), Pose.from_Rt(init_pose))
    opts = pnp.BundleOptions(loss_type=pnp.TRIVIAL)                     # :65-66
    cam, stats, _ = pnp.solve_pnp_iterative(Pc, x, None, cam, opts, max_inlier_error=0.0,
                                            optimize_focal_length=optimize_focal_length,
                                            optimize_principal_point=optimize_principal_point)
    R, t = cam.pose.R().astype(F), cam.pose.t.astype(F)
    if trans_type == MODEL:                                             # :79-91
        nmv = np.eye(4, dtype=F)
        nmv[:3, :3] = (R @ R0).astype(F)
        nmv[:3, 3] = ((R @ t0).astype(F) + t).astype(F)
        return (np.linalg.inv(init_view).astype(F) @ nmv).astype(F), cur_view, cam.intrinsics, stats
    U = np.eye(4, dtype=F)                                              # :92-101
    U[:3, :3] = R
    U[:3, 3] = t
    return cur_model, (U @ init_view).astype(F), cam.intrinsics, stats
