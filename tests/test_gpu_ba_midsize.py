"""Refine parity at a mid-size configuration -- 960x540, 24 frames, 1000 features/frame (~180 k residual
rows, 162 edges), the addon's OpenGL camera, the bench clip's motion law -- against the oracle
(oracle/ba.py), plus the behaviour VERDICT r1 asked about: does refine move the trajectory towards or
away from the ground truth?

Finding (profiles/r2_lk_track_error_by_skip.txt): round 1's configs[4] run refined AWAY from the truth
(max |t| error 0.062 -> 0.109) because its synthetic clip moved the 4K image by up to 89 px per 8 frames
-- outside the capture range of a 4-level pyramid with 10x10 windows -- so 54 % of the skip-4 and 96 % of
the skip-8 LK tracks were wrong by more than 3 px while still reporting status 1; the robust optimum of
those flows is not the true trajectory (the true trajectory itself cost 226 there).  With the motion the
survey specifies (<~ 20 px at skip 8, polychase_b200/synth.py::survey_speed) the flows are good and refine
converges onto the ground truth; this test pins that."""
import numpy as np
import pytest

from oracle import geometry as G
from oracle import pnp as opnp
from oracle import synth
from polychase_b200 import synth as psynth
from tests import helpers as H
from tests import test_gpu_track_refine as T

pytestmark = pytest.mark.gpu
F = np.float32


@pytest.fixture(scope="module")
def mid_scene(ctx_small):
    w, h, NF, first = 960, 540, 24, 5
    conv = G.OPENGL
    clip = synth.Clip(w, h, NF, seed=2, first_frame=first, speed=psynth.survey_speed(3840) * 4)   # 4K's px motion at 960 wide
    kps, flows = T.analyze_clip(ctx_small, clip, NF, 1000, conv, first=first)
    verts, tris = H.bumpy_mesh(clip, quads=8, amp=0.02)
    return dict(clip=clip, kps=kps, flows=flows, verts=verts, tris=tris, NF=NF, w=w, h=h, conv=conv, first=first)


def test_midsize_cost_and_normal_equations(ctx_small, mid_scene):
    prob = T.check_cost_and_normal_equations(ctx_small, mid_scene, False, False, 2, seed=11)
    assert len(prob.r_gkp) > 150_000 and len(prob.edges) == 8 * 24 - 30


def test_midsize_refine_matches_oracle_and_approaches_ground_truth(ctx_small, mid_scene):
    from polychase_b200 import capi
    sc = mid_scene
    NF, first = sc["NF"], sc["first"]
    # configs[4]'s perturbation: 0.2 degrees, 0.5 % of the depth on the interior frames
    model, traj, abi_edges, want, wst, band = T.ba_oracle_with_band(sc, 21, False, False, 12, seeds=(1,), rot_deg=0.2,
                                                                    trans=0.005 * sc["clip"].depth)
    ctx_small.mesh_set(sc["verts"], sc["tris"])
    ctx_small.ba_load([sc["kps"][first + k] for k in range(NF)], abi_edges, model, False, False)
    bo = capi.default_bundle(loss_type=2, max_iterations=12)
    truth = [T.cam_of(sc, first + k) for k in range(NF)]
    truth_cost = ctx_small.ba_cost([H.to_abi(c) for c in truth], bo)
    got, gst = ctx_small.ba_solve([H.to_abi(c) for c in traj], bo)
    assert abs(gst.initial_cost - wst.initial_cost) <= T.RTOL * abs(wst.initial_cost)
    assert abs(gst.cost - wst.cost) <= max(T.RTOL, band["cost"]) * abs(wst.cost), (gst.cost, wst.cost, band)
    for k in range(NF):
        dq, dt = H.pose_close(want[k], H.from_abi(got[k]))
        assert dq <= max(T.RTOL, band["q"]) and dt <= max(T.RTOL, band["t"]), (k, dq, dt, band)
    # towards the truth, not away from it
    before = max(H.pose_close(truth[k], traj[k])[1] for k in range(NF))
    after = max(H.pose_close(truth[k], H.from_abi(got[k]))[1] for k in range(NF))
    assert after < 0.05 * before, (before, after)
    assert gst.cost <= 1.02 * truth_cost and gst.cost < 0.05 * gst.initial_cost, (gst.cost, truth_cost, gst.initial_cost)
