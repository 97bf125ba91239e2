"""One device-resident forward of a cfg2 video, a few times (for ncu).  Usage: python -m tools.prof_video [video index in the cfg2 set] [reps]"""
import sys
import torch
from vrdone_b200 import MaskVRD, synth

idx = int(sys.argv[1]) if len(sys.argv) > 1 else 3
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = synth.load_config("vidor")
torch.manual_seed(0)
model = MaskVRD(cfg["model_config"], "cuda").eval().to("cuda")
model._config_eval(cfg["inference_config"])
s, nf, nt = synth.cfg2_video_set(10, 0)[idx]
v = synth.synthetic_video(cfg, s, n_tracklets=nt, n_frames=nf)
dv = {k: ([t.cuda() for t in x] if isinstance(x, list) else (x.cuda() if torch.is_tensor(x) else x)) for k, x in v.items()}
del v
for _ in range(reps):
    out = model(dv)
torch.cuda.synchronize()
print("pairs", len(dv["sids"]), "triplets", None if out is None else len(out["triplets"]))
