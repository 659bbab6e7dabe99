"""Multi-GPU sharding on the device path: two contexts on ONE GPU play ranks 0
and 1 of a 40-frame clip -- polychase_b200.shard.shard_range for the partition, pc_analyze_set_halo
for the previous shard's last 8 frames, pc_analyze_track_seed for the poses handed across the shard
boundary -- and every row a shard owns, and every pose it solves, must equal the unsharded pass
(/root/reference/cpp/opticalflow.cc:209-321 over the whole clip, tracker.cc:133-213)."""
import numpy as np
import pytest

from oracle import synth
from polychase_b200 import shard
from tests import helpers as H

pytestmark = pytest.mark.gpu
F = np.float32


def _run(ctx, clip, first, count, halo, seeds, model, verts, tris, max_corners):
    """Analyze + fused track of frames [first - halo, first + count).  Returns kps, flows, poses."""
    from polychase_b200 import capi
    ctx.mesh_set(verts, tris)
    ctx.analyze_begin(clip.width, clip.height, first - halo, count + halo, capi.default_gftt(max_corners=max_corners))
    ctx.analyze_set_halo(halo)
    ctx.analyze_track_begin(model, capi.default_bundle(loss_type=2))
    for f, cam in seeds.items():
        ctx.analyze_track_seed(f, cam)
    kps, flows, poses, tracked = {}, {}, {}, {}

    def take(r):
        f = r["frame_id"]
        kps[f] = r["keypoints"]
        for (a, b, rows, idx, tgt, err) in r["pairs"]:
            flows[(a, b)] = (idx, tgt, err)
        tracked[f] = r["tracked"]
        if r["tracked"]:
            poses[f] = r["camera"]

    for k in range(first - halo, first + count):
        ctx.analyze_push(k, clip.rgb(k))
        if ctx.analyze_pending() >= 4:
            take(ctx.analyze_pop())
    while ctx.analyze_pending():
        take(ctx.analyze_pop())
    ctx.analyze_end()
    return kps, flows, poses, tracked


def test_sharded_pass_equals_unsharded_pass(ctx_small):
    from polychase_b200 import capi
    w, h, NF, first, mc = 480, 352, 40, 3, 400
    clip = synth.Clip(w, h, NF, seed=6, first_frame=first)
    verts, tris = H.bumpy_mesh(clip, quads=8, amp=0.03)
    model = np.eye(4, dtype=F)
    start = H.to_abi(H.oracle_cam(clip, first))
    # the whole clip on one context
    kps0, flows0, poses0, tracked0 = _run(ctx_small, clip, first, NF, 0, {first: start}, model, verts, tris, mc)
    assert len(flows0) == 8 * NF - 30 and all(tracked0[f] == 1 for f in range(first + 1, first + NF))
    world = 2
    seen_pairs = set()
    for rank in range(world):
        s, c = shard.shard_range(first, NF, world, rank)
        halo = shard.halo_frames(first, s)
        assert halo == (0 if rank == 0 else 8)
        # poses handed across the boundary: the previous shard's last `halo` solved poses (rank 0: the scene's start pose)
        seeds = {first: start} if rank == 0 else {f: poses0[f] for f in range(s - halo, s)}
        with capi.Context(max_width=w, max_height=h, max_features=1024) as ctx:
            kps, flows, poses, tracked = _run(ctx, clip, s, c, halo, seeds, model, verts, tris, mc)
        want_pairs = set(shard.owned_pairs(first, NF, s, c))
        assert set(flows) == want_pairs                      # halo frames emit no rows; owned pairs all present
        assert not (want_pairs & seen_pairs)
        seen_pairs |= want_pairs
        for f in range(s - halo, s + c):                     # detector on halo frames == the owning shard's
            assert np.array_equal(kps[f], kps0[f])
        for pr in want_pairs:                                # bit for bit across the shard boundary
            for a, b in zip(flows[pr], flows0[pr]):
                assert np.array_equal(a, b), pr
        for f in range(s - halo, s):
            assert tracked[f] == 2                           # seeded halo poses come back as given
        for f in range(max(s, first + 1), s + c):
            assert tracked[f] == 1
            g, o = poses[f], poses0[f]
            assert np.abs(np.array(g.q[:]) - np.array(o.q[:])).max() <= 1e-6
            assert np.abs(np.array(g.t[:]) - np.array(o.t[:])).max() <= 1e-6 * max(1.0, np.abs(np.array(o.t[:])).max())
    assert len(seen_pairs) == 8 * NF - 30


def test_trajectory_allgather_roundtrip_layout():
    """The record the stitch collective moves is the 64-byte pc_camera_state (16 floats)."""
    from polychase_b200 import capi
    import ctypes as C
    assert C.sizeof(capi.CameraState) == 4 * shard.CAMERA_STATE_FLOATS
