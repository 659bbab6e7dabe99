"""float32 restatement of the reference's sequential tracker -- TEST INFRASTRUCTURE.

PARITY UNPINNED beyond oracle/pnp.py's own checks (no reference fixtures for the tracker).

  SolveFrame              /root/reference/cpp/tracker.cc:36-131
  TrackCameraTrajectory   /root/reference/cpp/tracker.cc:133-192
  TrackSequence           /root/reference/cpp/tracker.cc:194-213
`flows` maps (from, to) -> (src_kps_indices, tgt_kps, flow_errors); `keypoints` maps frame ->
(N,2).  Sources are visited in ascending image_id_from, the order SQLite returns them for
FindOpticalFlowsToImage."""
from __future__ import annotations

import numpy as np

from . import pnp, raycast
from .geometry import F, CameraState, Pose


def gather_matches(keypoints, flows, traj, frame_id, model, verts, tris, mask_bits):
    """tracker.cc:43-93.  traj: dict frame -> CameraState (filled frames only)."""
    Xs, xs = [], []
    model = np.asarray(model, F).reshape(4, 4)
    for src in sorted(f for (f, t) in flows if t == frame_id):
        if src not in traj:                                       # :48-50
            continue
        idx, tgt, _ = flows[(src, frame_id)]
        if len(idx) == 0:
            continue
        cam = traj[src]
        kp = np.asarray(keypoints[src], F)[np.asarray(idx, np.int64)]
        origin, dirs = raycast.ray_object_space(model, cam.pose.Rt4x4(), cam.intrinsics, kp)   # :69-77
        hit, pos, _, _, _ = raycast.ray_cast(verts, tris, mask_bits, origin, dirs, True)
        world = (pos @ model[:3, :3].T + model[:3, 3]).astype(F)                                # :80-82
        Xs.append(world[hit])
        xs.append(np.asarray(tgt, F).reshape(-1, 2)[hit])
    if not Xs:
        return np.zeros((0, 3), F), np.zeros((0, 2), F)
    return np.concatenate(Xs), np.concatenate(xs)


def solve_frame(keypoints, flows, traj, frame_id, model, verts, tris, mask_bits, opts, opt_f=False, opt_pp=False):
    X, x = gather_matches(keypoints, flows, traj, frame_id, model, verts, tris, mask_bits)
    if len(X) < 3:                                                # :95-97
        return None
    if frame_id in traj:                                          # :111-119
        init = traj[frame_id]
    elif frame_id - 1 in traj:
        init = traj[frame_id - 1]
    elif frame_id + 1 in traj:
        init = traj[frame_id + 1]
    else:
        init = CameraState(None, Pose())
    return pnp.solve_pnp_iterative(X, x, None, init, opts, 12.0, opt_f, opt_pp) + (len(X),)


def track_sequence(keypoints, flows, frame_from, frame_to, start: CameraState, model, verts, tris, mask_bits,
                   opts: pnp.BundleOptions, opt_f=False, opt_pp=False, callback=None):
    """tracker.cc:133-213.  Returns dict frame -> (CameraState, BundleStats, inlier_ratio)."""
    traj = {frame_from: start.copy()}
    out = {}
    d = 1 if frame_from < frame_to else -1
    f = frame_from + d
    while f != frame_to + d:
        r = solve_frame(keypoints, flows, traj, f, model, verts, tris, mask_bits, opts, opt_f, opt_pp)
        if r is None:
            raise RuntimeError(f"Could not track to frame: {f}. Not enough features.")    # :162-166
        cam, stats, inl, m = r
        out[f] = (cam, stats, inl, m)
        if callback is not None and not callback(f, cam, stats, inl):
            return out
        traj[f] = cam
        f += d
    return out
