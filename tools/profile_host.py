"""Host-side profile of one forward (cProfile) + GPU/host time split.  Usage: python -m tools.profile_host [tracklets]"""
import cProfile
import pstats
import sys
import time

import torch

from vrdone_b200 import MaskVRD, synth

n_trk = int(sys.argv[1]) if len(sys.argv) > 1 else 16
cfg = synth.load_config("vidor")
torch.manual_seed(0)
model = MaskVRD(cfg["model_config"], "cuda").eval().to("cuda")
model._config_eval(cfg["inference_config"])
if len(sys.argv) > 2:
    model.max_rows = int(sys.argv[2])
video = synth.synthetic_video(cfg, 0, n_tracklets=n_trk, n_frames=1200)
dv = {k: ([t.cuda() for t in v] if isinstance(v, list) else (v.cuda() if torch.is_tensor(v) else v)) for k, v in video.items()}
print("pairs", len(video["sids"]), "frames", sum(int(f.shape[1]) for f in video["so_features_list"]))
for _ in range(3):
    model(dv)
torch.cuda.synchronize()
for rep in range(2):
    t0 = time.perf_counter()
    feats = dv["so_features_list"]
    lens = [int(f.shape[1]) for f in feats]
    from vrdone_b200.layout import reference_padded_lengths
    tp = reference_padded_lengths(lens, cfg["model_config"])
    t1 = time.perf_counter()
    r = model.run_network(feats, tp, 6)
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    out = model(dv)
    t4 = time.perf_counter()
    print(f"tpads {1e3*(t1-t0):.1f} ms | run_network enqueue {1e3*(t2-t1):.1f} ms | gpu drain {1e3*(t3-t2):.1f} ms | full forward {1e3*(t4-t3):.1f} ms")
pr = cProfile.Profile()
pr.enable()
model(dv)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(30)
