"""Where do Python GC pauses happen during forward_test? (GPU box)"""
import gc, os, sys, time, traceback
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vrdone_b200 import MaskVRD, synth
cfg = synth.load_config("vidor")
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = MaskVRD(cfg["model_config"], dev).eval().to(dev)
model._config_eval(cfg["inference_config"])
videos = [synth.synthetic_video(cfg, s, n_tracklets=40, n_frames=1200) for s in (0, 1)]
dvs = [{k: ([t.cuda() for t in v] if isinstance(v, list) else v) for k, v in video.items()} for video in videos]
for dv in dvs: model(dv)
print("freeze count", gc.get_freeze_count(), "thresholds", gc.get_threshold(), "tracked", len(gc.get_objects()))
t0 = [0.0]
def cb(phase, info):
    if phase == "start":
        t0[0] = time.perf_counter()
        cb.stack = traceback.extract_stack(limit=8)[:-1]
        cb.counts = gc.get_count()
    else:
        d = 1e3 * (time.perf_counter() - t0[0])
        if d > 1.0:
            print(f"  gc gen{info['generation']} {d:.1f} ms collected={info['collected']} counts_before={cb.counts} at " +
                  " <- ".join(f"{os.path.basename(f.filename)}:{f.lineno}" for f in reversed(cb.stack[-4:])))
gc.callbacks.append(cb)
for s in range(6):
    torch.cuda.synchronize(); t = time.perf_counter()
    out = model(dvs[s % 2])
    t1 = time.perf_counter()
    del out
    t2 = time.perf_counter()
    print(f"step {s}: forward {1e3*(t1-t):.1f} ms, del {1e3*(t2-t1):.1f} ms, stats {({k: round(v, 1) for k, v in model.last_stats.items()})}")
