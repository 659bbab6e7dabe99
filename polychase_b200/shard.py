"""Multi-GPU sharding of the hot path: frames shard as one contiguous
sub-sequence per GPU; the only collective is one all-gather of packed CameraState records
(64 B per frame, include/polychase_b200.h pc_camera_state) that stitches the per-GPU trajectory
segments before the global refine.  No data-path collective is needed for Analyze: every
directed pair (a, b) is owned by the shard that holds its later frame max(a, b), and each shard
re-prepares the previous shard's last 8 frames as a halo."""
from __future__ import annotations

from typing import List, Tuple

SKIPS = (1, 2, 4, 8)      # |image_skips|, /root/reference/cpp/opticalflow.cc:76-77
HALO = max(SKIPS)
CAMERA_STATE_FLOATS = 16


def shard_range(first_frame: int, num_frames: int, world: int, rank: int) -> Tuple[int, int]:
    """(start, count) of rank's contiguous sub-sequence; the remainder goes to the first ranks."""
    base, rem = divmod(num_frames, world)
    count = base + (1 if rank < rem else 0)
    start = first_frame + rank * base + min(rank, rem)
    return start, count


def halo_frames(first_frame: int, start: int) -> int:
    """Frames of the previous shard this shard must prepare (pyramid + keypoints) as partners."""
    return min(HALO, start - first_frame)


def owned_pairs(first_frame: int, num_frames: int, start: int, count: int) -> List[Tuple[int, int]]:
    """Directed pairs whose later frame lies in [start, start+count)."""
    out = []
    last = first_frame + num_frames
    for j in range(start, start + count):
        for d in SKIPS:
            i = j - d
            if i >= first_frame and j < last:
                out.append((i, j))
                out.append((j, i))
    return out


def allgather_trajectory(local, counts: List[int]):
    """Stitches per-rank trajectory segments.  `local`: float32 tensor (counts[rank], 16) on the
    rank's device (CUDA -> NCCL over NVLink, CPU -> gloo).  Returns (sum(counts), 16)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    assert len(counts) == world and local.shape[1] == CAMERA_STATE_FLOATS
    pad = max(counts)
    buf = torch.zeros((pad, CAMERA_STATE_FLOATS), dtype=torch.float32, device=local.device)
    buf[: local.shape[0]] = local
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return torch.cat([o[:n] for o, n in zip(out, counts)], dim=0)
