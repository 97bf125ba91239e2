"""Timeline of one forward_test with pinned host inputs: per chunk, when its copies and kernels start/end on the device
(CUDA events on the copy and compute streams) and when the host issued them."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vrdone_b200 import MaskVRD, synth

cfg = synth.load_config("vidor")
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = MaskVRD(cfg["model_config"], dev).eval().to(dev)
model._config_eval(cfg["inference_config"])
video = synth.synthetic_video(cfg, 0, n_tracklets=40, n_frames=1200)
pinned = dict(video)
pinned["so_features_list"] = [t.t().contiguous().pin_memory().t() for t in video["so_features_list"]]
import os as _os
if _os.environ.get('H2D_CHUNK'): model.h2d_chunk_rows = int(_os.environ['H2D_CHUNK'])
if _os.environ.get('H2D_EDGE'): model.h2d_edge_rows = int(_os.environ['H2D_EDGE'])
for _ in range(3):
    model(pinned)
eng = model._get_engine()
rec = []
t0 = [0.0]
orig_prep, orig_fwd, orig_pred = model._prepare_chunk, eng.backbone, eng.predict

def ev(stream):
    e = torch.cuda.Event(enable_timing=True); e.record(stream); return e

def prep(ops, feats, lens, tpads, chunk, ci, dev_, cur, any_host):
    cs = model._copy_stream if any_host else cur
    h0 = time.perf_counter()
    a = ev(cs)
    r = orig_prep(ops, feats, lens, tpads, chunk, ci, dev_, cur, any_host)
    b = ev(cs)
    rec.append(("copy", chunk[1] - chunk[0], a, b, 1e3 * (h0 - t0[0]), 1e3 * (time.perf_counter() - t0[0])))
    return r

def fwd(lay, *a, **k):
    cur = torch.cuda.current_stream()
    h0 = time.perf_counter()
    s = ev(cur)
    r = orig_fwd(lay, *a, **k)
    e = ev(cur)
    rec.append(("compute", lay.B, s, e, 1e3 * (h0 - t0[0]), 1e3 * (time.perf_counter() - t0[0])))
    return r

model._prepare_chunk = prep
eng.backbone = fwd

def pred(lay, *a, **k):
    cur = torch.cuda.current_stream()
    h0 = time.perf_counter(); s_ = ev(cur); r = orig_pred(lay, *a, **k); e_ = ev(cur)
    rec.append(("predict", lay.B, s_, e_, 1e3 * (h0 - t0[0]), 1e3 * (time.perf_counter() - t0[0])))
    return r

eng.predict = pred
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
