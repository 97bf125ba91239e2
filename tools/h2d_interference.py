"""Does a concurrent host->device copy stream slow the kernels down (L2 pollution / HBM contention)?  Network-only time per
video with the pair features resident in HBM, alone and while a side stream keeps copying pinned host memory to the device."""
import sys
import time
import torch
from vrdone_b200 import MaskVRD, synth
from vrdone_b200.layout import reference_padded_lengths

cfg = synth.load_config("vidor")
torch.manual_seed(0)
model = MaskVRD(cfg["model_config"], "cuda").eval().to("cuda")
model._config_eval(cfg["inference_config"])
v = synth.synthetic_video(cfg, 0, n_tracklets=40, n_frames=1200)
feats = [f.cuda() for f in v["so_features_list"]]
tp = reference_padded_lengths([int(f.shape[1]) for f in feats], cfg["model_config"])
side = torch.cuda.Stream()
src = torch.empty(160 << 20, dtype=torch.uint8, pin_memory=True)
dst = [torch.empty(160 << 20, dtype=torch.uint8, device="cuda") for _ in range(2)]


def run(n, copies):
    for _ in range(2):
        model.run_network(feats, tp, model.topk)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        if copies:
            with torch.cuda.stream(side):
                for j in range(copies):
                    dst[j % 2].copy_(src, non_blocking=True)
        model.run_network(feats, tp, model.topk)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print("alone            : %.2f ms per video" % run(8, 0))
print("with 4 x 160 MB  : %.2f ms per video (0.64 GB copied per video)" % run(8, 4))
print("with 7 x 160 MB  : %.2f ms per video (1.1 GB copied per video)" % run(8, 7))
print("alone            : %.2f ms per video" % run(8, 0))
# the same 1.1 GB per video into a small destination that stays in L2 (no DRAM write-back of the copied data)
src = torch.empty(8 << 20, dtype=torch.uint8, pin_memory=True)
dst = [torch.empty(8 << 20, dtype=torch.uint8, device="cuda") for _ in range(2)]
print("with 140 x 8 MB  : %.2f ms per video (1.1 GB copied per video, 16 MB of destinations)" % run(8, 140))
# device-to-device copies of the same volume (no PCIe)
src = torch.empty(160 << 20, dtype=torch.uint8, device="cuda")
dst = [torch.empty(160 << 20, dtype=torch.uint8, device="cuda") for _ in range(2)]
print("with 7 x 160 MB D2D: %.2f ms per video" % run(8, 7))
