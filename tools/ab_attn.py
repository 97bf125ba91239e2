"""Times the SOS full-attention kernel alone on the q / k / v shapes of a few cfg2 videos (CUDA events, per video layout)."""
import sys, json
import torch
from vrdone_b200 import synth
from vrdone_b200.cuda_ops import CudaOps
from vrdone_b200.layout import PackLayout, reference_padded_lengths

cfg = synth.load_config("vidor")
mc = cfg["model_config"]
ops = CudaOps()
specs = synth.cfg2_video_set(10, 0)
st = cfg["dataset_config"]["feat_stride"]
res = {}
tot = 0.0
for vi in (3, 0, 6, 1, 5):
    s, nf, nt = specs[vi]
    _, _, durs, *_ = synth._tracklets(cfg, s, nt, nf, features=False, split_rng=True)
    pr = synth._overlapping_pairs(durs, st)
    lens = [len(range(0, min(durs[a][1], durs[b][1]) - max(durs[a][0], durs[b][0]), st)) for a, b in pr]
    lay = PackLayout(lens, reference_padded_lengths(lens, mc), 4, "cuda")
    l0 = lay.levels[0]
    R, C = l0.R, mc["embd_dim"]
    q = (torch.randn(R, C, device="cuda") * 0.5).to(torch.bfloat16)
    k = (torch.randn(R, C, device="cuda") * 0.5).to(torch.bfloat16)
    v = torch.randn(R, C, device="cuda").to(torch.bfloat16)
    o = torch.empty_like(q)
    for _ in range(3):
        ops.full_attn(q, k, v, o, l0, mc["fuse_head"])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        ops.full_attn(q, k, v, o, l0, mc["fuse_head"])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    fl = 4.0 * C * sum(float(x) * x for x in lens)
    res[vi] = {"pairs": len(lens), "us": round(ms * 1e3, 1), "tflops": round(fl / ms / 1e9, 1)}
    tot += ms
print(json.dumps({"per_video": res, "total_us": round(tot * 1e3, 1)}))
