"""Multi-GPU paths through the C ABI's own NCCL layer (csrc/abi/comm.cu), 2 ranks on 2 GPUs:
  * pc_traj_allgather -- the path's one collective: per-rank trajectory segments stitched;
  * edge-sharded refine (section 8f.4): pc_ba_set_edge_shard + pc_ba_solve on every rank must reproduce the
    single-GPU pc_ba_solve bit for bit (each per-edge block / cost is produced by exactly one rank and gathered).
Skipped on a box with fewer than two GPUs (the round-end suite runs on one); run with `gpurun --gpus 2`."""
import os
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
F = np.float32


def _worker(rank, world, uid_path, out_path):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import time
    from oracle import synth
    from polychase_b200 import capi, shard
    from tests import helpers as H
    from oracle import geometry as G
    w, h, NF, mc = 480, 352, 20, 400
    clip = synth.Clip(w, h, NF, seed=6)
    verts, tris = H.bumpy_mesh(clip, quads=8, amp=0.03)
    with capi.Context(device=rank, max_width=w, max_height=h, max_features=1024) as ctx:
        # the communicator: rank 0 makes the id, the others read it (any side channel works)
        if rank == 0:
            uid = capi.comm_unique_id()
            with open(uid_path + ".tmp", "wb") as f:
                f.write(uid)
            os.replace(uid_path + ".tmp", uid_path)
        else:
            t0 = time.time()
            while not os.path.exists(uid_path):
                assert time.time() - t0 < 120
                time.sleep(0.01)
            uid = open(uid_path, "rb").read()
        ctx.comm_init(world, rank, uid)
        # ---- trajectory all-gather: rank r contributes its shard's poses -----------------------------------
        counts = [shard.shard_range(0, NF, world, r)[1] for r in range(world)]
        s, c = shard.shard_range(0, NF, world, rank)
        local = [H.to_abi(H.oracle_cam(clip, k, G.OPENGL)) for k in range(s, s + c)]
        full = ctx.traj_allgather(local, counts)
        ok_gather = len(full) == NF
        for k in range(NF):
            want = H.to_abi(H.oracle_cam(clip, k, G.OPENGL))
            ok_gather = ok_gather and bytes(full[k]) == bytes(want)
        # ---- edge-sharded refine == single-GPU refine ------------------------------------------------------
        kps, flows = {}, {}
        ctx.analyze_begin(w, h, 0, NF, capi.default_gftt(max_corners=mc))
        for k in range(NF):
            ctx.analyze_push(k, clip.rgb(k))
            if ctx.analyze_pending() >= 3:
                r = ctx.analyze_pop()
                kps[r["frame_id"]] = r["keypoints"]
                flows.update({(a, b): (idx, tgt) for (a, b, rows, idx, tgt, err) in r["pairs"]})
        while ctx.analyze_pending():
            r = ctx.analyze_pop()
            kps[r["frame_id"]] = r["keypoints"]
            flows.update({(a, b): (idx, tgt) for (a, b, rows, idx, tgt, err) in r["pairs"]})
        ctx.analyze_end()
        rng = np.random.default_rng(5)
        traj = [H.oracle_cam(clip, k) for k in range(NF)]
        for k in range(1, NF - 1):
            traj[k] = H.perturb(traj[k], rng, rot_deg=0.05, trans=0.004)
        edges = [(a, b, flows[(a, b)][0], flows[(a, b)][1]) for (a, b) in sorted(flows) if len(flows[(a, b)][0])]
        ctx.mesh_set(verts, tris)
        bo = capi.default_bundle(loss_type=2, max_iterations=15)
        atraj = [H.to_abi(cm) for cm in traj]
        ctx.ba_load([kps[k] for k in range(NF)], edges, np.eye(4, dtype=F), False, False)
        single, sst = ctx.ba_solve(atraj, bo)                        # every edge on this GPU
        ctx.ba_load([kps[k] for k in range(NF)], edges, np.eye(4, dtype=F), False, False)
        ctx.ba_set_edge_shard(True)
        costs = []
        sharded, hst = ctx.ba_solve(atraj, bo, callback=lambda st: costs.append(float(st.cost)) or True)
        same = all(bytes(a) == bytes(b) for a, b in zip(single, sharded))
        ctx.comm_destroy()
    np.save(out_path % rank, np.array([int(ok_gather), int(same), int(sst.iterations), int(hst.iterations),
                                        float(sst.cost), float(hst.cost), float(sst.initial_cost), len(edges)], np.float64))


def test_two_rank_allgather_and_edge_sharded_refine():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    tmp = tempfile.mkdtemp()
    uid_path = os.path.join(tmp, "nccl_id")
    out_path = os.path.join(tmp, "rank%d.npy")
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, 2, uid_path, out_path)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    res = [np.load(out_path % r) for r in range(2)]
    for r in res:
        assert r[0] == 1, "pc_traj_allgather result differs from the per-rank segments"
        assert r[1] == 1, "edge-sharded refine is not bit-equal to the single-GPU refine"
        assert r[2] == r[3] and r[2] >= 2 and r[4] == r[5] and r[5] < 0.5 * r[6]
    assert np.array_equal(res[0], res[1])            # both ranks end in the same state
