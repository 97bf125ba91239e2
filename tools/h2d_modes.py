"""Host->device copy throughput of one video's pinned pair tensors under different issue patterns (CUDA events, no compute).
Usage: python -m tools.h2d_modes [video index in the cfg2 set]"""
import sys
import time
import numpy as np
import torch
from vrdone_b200 import synth
from vrdone_b200.cuda_ops import CudaOps

idx = int(sys.argv[1]) if len(sys.argv) > 1 else 0
cfg = synth.load_config("vidor")
s, nf, nt = synth.cfg2_video_set(10, 0)[idx]
v = synth.synthetic_video(cfg, s, n_tracklets=nt, n_frames=nf)
feats = v["so_features_list"]
sizes = np.array([f.numel() * 4 for f in feats], dtype=np.int64)
total = int(sizes.sum())
gap = 256
arena = torch.empty(total + gap * len(feats), dtype=torch.uint8, pin_memory=True)
offs = np.cumsum(sizes + gap) - (sizes + gap)
src = arena.data_ptr() + offs
dst = torch.empty(total + 256 * len(feats), dtype=torch.uint8, device="cuda")
doffs = (np.cumsum((sizes + 255) // 256 * 256) - (sizes + 255) // 256 * 256).astype(np.int64)
ops = CudaOps()
streams = [torch.cuda.Stream() for _ in range(4)]
print(f"video {idx}: {len(feats)} pairs, {total / 1e9:.2f} GB, mean copy {sizes.mean() / 1e6:.2f} MB")


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        fn()
        host = time.perf_counter() - t0
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, host


def per_pair_one_stream():
    ops.h2d_pairs(np.ascontiguousarray(src), np.ascontiguousarray(sizes), dst, np.ascontiguousarray(doffs), torch.cuda.current_stream())


def per_pair_n_streams(n):
    def f():
        cur = torch.cuda.current_stream()
        e = torch.cuda.Event(); e.record(cur)
        for k in range(n):
            streams[k].wait_event(e)
            sel = np.arange(k, len(feats), n)
            ops.h2d_pairs(np.ascontiguousarray(src[sel]), np.ascontiguousarray(sizes[sel]), dst, np.ascontiguousarray(doffs[sel]), streams[k])
            d = torch.cuda.Event(); d.record(streams[k]); cur.wait_event(d)
    return f


def one_big():
    dst[: total].copy_(arena[: total], non_blocking=True)


for name, fn in (("per-pair copies, one stream", per_pair_one_stream), ("per-pair copies, 2 streams", per_pair_n_streams(2)),
                 ("per-pair copies, 4 streams", per_pair_n_streams(4)), ("one copy of the whole arena", one_big)):
    ms, host = timed(fn)
    print(f"{name:32s} {ms:8.2f} ms  {total / ms / 1e6:6.1f} GB/s   host issue {1e3 * host:6.2f} ms")
