"""Per-GPU pinned host->device copy rate with every rank copying at once, for three placements of the page-locked
buffer: wherever the process's default policy puts it, bound to the GPU's NUMA node, bound to the other node.
Run under torchrun (one rank per GPU).  Rank 0 prints one JSON line."""
import ctypes
import json
import os

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
libc = ctypes.CDLL(None, use_errno=True)
SYS_set_mempolicy, MPOL_DEFAULT, MPOL_PREFERRED, MPOL_BIND = 238, 0, 1, 2


def set_policy(mode, node):
    mask = ctypes.c_ulong(0 if node is None else 1 << node)
    rc = libc.syscall(SYS_set_mempolicy, mode, ctypes.byref(mask), 64)
    return rc if rc == 0 else -ctypes.get_errno()


def read(path):
    try:
        return open(path).read().strip()
    except Exception as e:
        return repr(e)


bdf = torch.cuda.get_device_properties(local).pci_bus_id if hasattr(torch.cuda.get_device_properties(local), "pci_bus_id") else None
import subprocess
q = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local)], capture_output=True, text=True).stdout.strip()
sysfs = "/sys/bus/pci/devices/" + q.lower().replace("00000000:", "0000:") + "/numa_node"
gpu_node = read(sysfs)
info = {"rank": rank, "bus": q, "gpu_numa_node": gpu_node, "cpus_allowed": read("/proc/self/status").split("Cpus_allowed_list:")[1].split("\n")[0].strip(),
        "mems_allowed": read("/proc/self/status").split("Mems_allowed_list:")[1].split("\n")[0].strip()}
frame = 3840 * 2160 * 3
n_frames = 16
dst = torch.empty(frame, dtype=torch.uint8, device="cuda")
streams = [torch.cuda.Stream() for _ in range(4)]


def rate(buf):
    cut = [(frame * i // 4) // 256 * 256 for i in range(4)] + [frame]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0.record()
    for s in streams:
        s.wait_event(ev0)
    for k in range(3 * n_frames):
        src = buf[(k % n_frames) * frame:(k % n_frames + 1) * frame]
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                dst[cut[i]:cut[i + 1]].copy_(src[cut[i]:cut[i + 1]], non_blocking=True)
    for s in streams:
        e = torch.cuda.Event()
        e.record(s)
        torch.cuda.current_stream().wait_event(e)
    ev1.record()
    torch.cuda.synchronize()
    return 3 * n_frames * frame / (ev0.elapsed_time(ev1) * 1e-3) / 1e9


try:
    node = int(gpu_node)
except ValueError:
    node = -1
for label, mode, nd in (("default", MPOL_DEFAULT, None), ("gpu_node", MPOL_BIND, node if node >= 0 else None),
                        ("other_node", MPOL_BIND, (1 - node) if node in (0, 1) else None)):
    if label != "default" and nd is None:
        info[label] = None
        continue
    rc = set_policy(mode, nd)
    buf = torch.empty(frame * n_frames, dtype=torch.uint8)
    buf.fill_(7)                                   # first touch under the policy
    cudart = torch.cuda.cudart()
    rcr = cudart.cudaHostRegister(buf.data_ptr(), buf.numel(), 0)
    rate(buf)
    info[label] = {"policy_rc": rc, "register_rc": int(rcr), "gbs": round(rate(buf), 2)}
    cudart.cudaHostUnregister(buf.data_ptr())
    set_policy(MPOL_DEFAULT, None)
    del buf
if world > 1:
    allinfo = [None] * world
    dist.all_gather_object(allinfo, info)
else:
    allinfo = [info]
if rank == 0:
    out = {"ranks": allinfo, "nodes": {n: read(f"/sys/devices/system/node/{n}/cpulist") for n in sorted(os.listdir("/sys/devices/system/node")) if n.startswith("node")},
           "nproc": os.cpu_count()}
    for k in ("default", "gpu_node", "other_node"):
        vals = [r[k]["gbs"] for r in allinfo if r.get(k)]
        if vals:
            out["sum_gbs_" + k] = round(sum(vals), 1)
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
