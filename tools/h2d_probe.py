"""Probe of the host->device path on the GPU box: pinned copy bandwidth (one big copy, per-pair copies through vrd_h2d_pairs),
NUMA affinity of the GPU, and the chunk timeline of forward_test with pinned inputs."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vrdone_b200 import MaskVRD, synth

def bw(label, fn, nbytes, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t)
    print(f"{label}: {nbytes / best / 1e9:.1f} GB/s ({1e3 * best:.1f} ms)")

print("cpu count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
try:
    import pynvml
    pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
    n = (os.cpu_count() + 63) // 64
    mask = pynvml.nvmlDeviceGetCpuAffinity(h, n)
    cores = [i * 64 + b for i, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
    print("gpu0 cpu affinity:", len(cores), cores[:4], "...", cores[-4:])
    print("pcie gen", pynvml.nvmlDeviceGetCurrPcieLinkGeneration(h), "width", pynvml.nvmlDeviceGetCurrPcieLinkWidth(h))
except Exception as e:
    cores = None; print("nvml:", e)
os.system("nvidia-smi topo -m 2>/dev/null | head -20; numactl -H 2>/dev/null | head; lscpu | grep -i -E 'numa|socket|model name'")
dev = torch.device("cuda:0")
N = 1 << 30
for tag in ("default", "gpu-local cores"):
    if tag != "default":
        if not cores: break
        os.sched_setaffinity(0, set(cores) & os.sched_getaffinity(0) or os.sched_getaffinity(0))
    src = torch.empty(N, dtype=torch.uint8).pin_memory(); src.fill_(1)
    dst = torch.empty(N, dtype=torch.uint8, device=dev)
    bw(f"[{tag}] one 1 GiB pinned copy", lambda: dst.copy_(src, non_blocking=True), N)
    back = torch.empty(N, dtype=torch.uint8).pin_memory()
    bw(f"[{tag}] one 1 GiB D2H copy", lambda: back.copy_(dst, non_blocking=True), N)
    del src, dst, back

cfg = synth.load_config("vidor")
torch.manual_seed(0)
model = MaskVRD(cfg["model_config"], dev).eval().to(dev)
model._config_eval(cfg["inference_config"])
video = synth.synthetic_video(cfg, 0, n_tracklets=40, n_frames=1200)
pinned = dict(video)
pinned["so_features_list"] = [t.t().contiguous().pin_memory().t() for t in video["so_features_list"]]
nbytes = sum(t.numel() * 4 for t in video["so_features_list"])
model(pinned); model(pinned)
ops = model._ops
feats = pinned["so_features_list"]
meta = np.array([f.data_ptr() for f in feats], dtype=np.int64)
sizes = np.array([f.numel() * 4 for f in feats], dtype=np.int64)
offs = np.cumsum((sizes + 255) // 256 * 256) - (sizes + 255) // 256 * 256
buf = torch.empty(int(offs[-1] + sizes[-1] + 256), dtype=torch.uint8, device=dev)
st = torch.cuda.Stream()
bw("per-pair copies via vrd_h2d_pairs", lambda: ops.h2d_pairs(meta, sizes, buf, offs, st), nbytes)
t = time.perf_counter(); ops.h2d_pairs(meta, sizes, buf, offs, st); print("  host enqueue of", len(feats), "copies: %.2f ms" % (1e3 * (time.perf_counter() - t))); torch.cuda.synchronize()
for rows in (24576, 36864, 49152, 73728, 196608):
    model.h2d_chunk_rows = rows
    model(pinned)
    ts = []
    for _ in range(3):
        torch.cuda.synchronize(); t = time.perf_counter(); out = model(pinned); ts.append(1e3 * (time.perf_counter() - t)); del out
    print(f"h2d_chunk_rows {rows}: forward wall ms {[round(x, 1) for x in ts]}  stats {({k: round(v, 1) for k, v in model.last_stats.items()})}")
dv = {k: ([t.cuda() for t in v] if isinstance(v, list) else v) for k, v in video.items()}
model(dv)
ts = []
for _ in range(3):
    torch.cuda.synchronize(); t = time.perf_counter(); out = model(dv); ts.append(1e3 * (time.perf_counter() - t)); del out
print(f"device-resident: forward wall ms {[round(x, 1) for x in ts]}  stats {({k: round(v, 1) for k, v in model.last_stats.items()})}")
