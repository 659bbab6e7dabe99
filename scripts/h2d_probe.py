"""Raw pinned host->device copy rates on this box: one 4K RGB frame (24.9 MB) per copy, 1 / 2 / 4 streams (slices of the
frame on several copy engines), and the same from a 1 GB buffer.  Prints one JSON line."""
import json
import subprocess
import torch

frame = 3840 * 2160 * 3
out = {"frame_bytes": frame}
dst = torch.empty(frame, dtype=torch.uint8, device="cuda")
src = torch.empty(frame, dtype=torch.uint8).pin_memory()


def rate(n_streams, reps=40):
    streams = [torch.cuda.Stream() for _ in range(n_streams)]
    cut = [(frame * i // n_streams) // 256 * 256 for i in range(n_streams)] + [frame]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for s in streams:
        s.wait_event(ev0)
    for _ in range(reps):
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                dst[cut[i]:cut[i + 1]].copy_(src[cut[i]:cut[i + 1]], non_blocking=True)
    for s in streams:
        ev = torch.cuda.Event()
        ev.record(s)
        torch.cuda.current_stream().wait_event(ev)
    ev1.record()
    torch.cuda.synchronize()
    return reps * frame / (ev0.elapsed_time(ev1) * 1e-3) / 1e9


for n in (1, 2, 4):
    rate(n, 5)
    out[f"gbs_{n}_streams"] = round(rate(n), 2)
big = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
dbig = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
dbig.copy_(big, non_blocking=True)
torch.cuda.synchronize()
ev0.record()
dbig.copy_(big, non_blocking=True)
ev1.record()
torch.cuda.synchronize()
out["gbs_1GiB_copy"] = round((1 << 30) / (ev0.elapsed_time(ev1) * 1e-3) / 1e9, 2)
for cmd in (["nvidia-smi", "--query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max", "--format=csv,noheader"],
            ["numactl", "--hardware"], ["nproc"]):
    try:
        out[" ".join(cmd[:2])] = subprocess.run(cmd, capture_output=True, text=True, timeout=20).stdout.strip()[:600]
    except Exception as e:
        out[" ".join(cmd[:2])] = repr(e)
print(json.dumps(out))
