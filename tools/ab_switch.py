"""A/B of launcher switches on the same device-resident videos in ONE process (vrd_set_option): network-only time of a few cfg2
videos per configuration, interleaved rounds.  Usage: python -m tools.ab_switch [n_videos] [rounds] [max_rows list]"""
import json
import sys
import torch
from vrdone_b200 import MaskVRD, synth
from vrdone_b200.layout import reference_padded_lengths

n_videos = int(sys.argv[1]) if len(sys.argv) > 1 else 4
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = synth.load_config("vidor")
mc = cfg["model_config"]
torch.manual_seed(0)
model = MaskVRD(mc, "cuda").eval().to("cuda")
model._config_eval(cfg["inference_config"])
specs = synth.cfg2_video_set(10, 0)
order = [3, 0, 6, 1, 4, 5, 2, 7, 8, 9][:n_videos]          # the long-pair video first
vids = []
for i in order:
    s, nf, nt = specs[i]
    v = synth.synthetic_video(cfg, s, n_tracklets=nt, n_frames=nf)
    feats = [f.cuda() for f in v["so_features_list"]]
    lens = [int(f.shape[1]) for f in feats]
    vids.append((feats, reference_padded_lengths(lens, mc)))
    del v
pairs = sum(len(f) for f, _ in vids)
ops = None


def run_once():
    for feats, tp in vids:
        model.run_network(feats, tp, model.topk)


def timed(reps=2):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run_once()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


run_once()
ops = model._ops
configs = [("base: spec=0 pdl=0 dw=2", {"gemm_spec": 0, "pdl": 0, "dw_cfg": 2}), ("spec=1", {"gemm_spec": 1, "pdl": 0, "dw_cfg": 2}),
           ("spec=1 pdl=1", {"gemm_spec": 1, "pdl": 1, "dw_cfg": 2}), ("spec=1 dw=4", {"gemm_spec": 1, "pdl": 0, "dw_cfg": 4}), ("spec=1 dw=5", {"gemm_spec": 1, "pdl": 0, "dw_cfg": 5})]
if len(sys.argv) > 3 and sys.argv[3] == "one":
    configs = [("default", {})]
if len(sys.argv) > 3 and sys.argv[3] == "pdl":
    configs = [configs[1], configs[2]]
if len(sys.argv) > 3 and sys.argv[3] == "dw":
    configs = [configs[1], configs[3], configs[4]]
if len(sys.argv) > 3 and sys.argv[3] == "short":
    configs = configs[:2]
res = {name: [] for name, _ in configs}
ref = None
for r in range(rounds):
    for name, opt in configs:
        for k, v in opt.items():
            ops.set_option(k, v)
        run_once()
        res[name].append(round(timed(), 3))
out = {"pairs": pairs, "videos": order, "ms_per_pass": res, "best": {k: min(v) for k, v in res.items()}}
ops.set_option("pdl", 0)
ops.set_option("dw_cfg", 2)
# the specialised GEMM epilogues do the same arithmetic: bit-identical logits
outs = {}
for sp in (0, 1):
    ops.set_option("gemm_spec", sp)
    r = model.run_network(vids[0][0], vids[0][1], model.topk)
    torch.cuda.synchronize()
    outs[sp] = r["logits"].float().clone()
out["gemm_spec_bit_identical"] = bool(torch.equal(outs[0], outs[1]))
# numerical agreement of the two dwconv variants on the first video
model.max_rows = 196608
outs = {}
for dw in (2, 4):
    ops.set_option("dw_cfg", dw)
    r = model.run_network(vids[0][0], vids[0][1], model.topk)
    torch.cuda.synchronize()
    outs[dw] = r["logits"].float().clone()
out["dw4_vs_dw2_logits_maxabs"] = float((outs[4] - outs[2]).abs().max())
out["logits_scale"] = float(outs[2].abs().max())
print(json.dumps(out))
