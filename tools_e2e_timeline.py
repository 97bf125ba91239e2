"""Timeline of one forward_test with pinned host inputs: per chunk, when its copies and kernels start/end on the device
(CUDA events on the copy and compute streams) and when the host issued them."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from vrdone_b200 import MaskVRD, synth

cfg = synth.load_config("vidor")
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = MaskVRD(cfg["model_config"], dev).eval().to(dev)
model._config_eval(cfg["inference_config"])
video = synth.synthetic_video(cfg, 0, n_tracklets=40, n_frames=1200)
pinned = dict(video)
pinned["so_features_list"] = [t.t().contiguous().pin_memory().t() for t in video["so_features_list"]]
for _ in range(3):
    model(pinned)
eng = model._get_engine()
rec = []
t0 = [0.0]
orig_stage, orig_fwd = model._issue_copies, eng.forward_packed

def ev(stream):
    e = torch.cuda.Event(enable_timing=True); e.record(stream); return e

def stage(ops, plan, slot):
    cs = model._copy_stream
    h0 = time.perf_counter()
    if model._pack_done[slot] is not None:
        cs.wait_event(model._pack_done[slot])
    a = ev(cs)
    r = orig_stage(ops, plan, slot)
    b = ev(cs)
    rec.append(("copy", len(plan["idx"]), a, b, 1e3 * (h0 - t0[0]), 1e3 * (time.perf_counter() - t0[0])))
    return r

def fwd(lay, *a, **k):
    cur = torch.cuda.current_stream()
    h0 = time.perf_counter()
    s = ev(cur)
    r = orig_fwd(lay, *a, **k)
    e = ev(cur)
    rec.append(("compute", lay.B, s, e, 1e3 * (h0 - t0[0]), 1e3 * (time.perf_counter() - t0[0])))
    return r

model._issue_copies = stage
eng.forward_packed = fwd
for rep in range(2):
    rec.clear()
    torch.cuda.synchronize()
    base = ev(torch.cuda.current_stream())
    model._copy_stream.wait_event(base)
    t0[0] = time.perf_counter()
    out = model(pinned)
    wall = 1e3 * (time.perf_counter() - t0[0])
    torch.cuda.synchronize()
    print(f"--- rep {rep}: forward wall {wall:.1f} ms, stats {({k: round(v, 1) for k, v in model.last_stats.items()})}")
    for kind, n, a, b, h0, h1 in rec:
        print(f"{kind:8s} pairs {n:4d}  device {base.elapsed_time(a):6.1f} -> {base.elapsed_time(b):6.1f} ms   host issue {h0:6.1f} -> {h1:6.1f} ms")
