"""GPU debug aid (not a test): compares engine intermediates ("taps") with the oracle's on a small case and prints
where the first divergence appears.  Usage: python -m tools.gpu_debug [config] [fp32|bf16]"""
import sys

import torch

from oracle import maskvrd_oracle as O
from tests import helpers as H
from vrdone_b200 import synth
from vrdone_b200.layout import PackLayout, reference_padded_lengths


def rows_to_padded(t, lay, streams, T):
    """[streams*R, C] rows -> list over streams of (B, C, T_l) padded tensors (zeros beyond valid)."""
    outs = []
    for s in range(streams):
        x = torch.zeros(lay.B, t.shape[1], T)
        for i in range(lay.B):
            r0, L = s * lay.R + int(lay.off[i]), int(lay.len[i])
            x[i, :, :L] = t[r0:r0 + L].t()
        outs.append(x)
    return outs


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "vidvrd"
    precision = sys.argv[2] if len(sys.argv) > 2 else "fp32"
    fix = H.network_fixture(name)
    cfg, model, sd = H.seeded_model(name, fix["wseed"], precision=precision)
    mc = cfg["model_config"]
    T = mc["max_seq_len"]
    lens = [l for l in fix["lens"] if l <= T][:6]
    feats = synth.pair_features(mc, lens, 77)
    x, m = O.pad_batch(feats, T)
    taps = {}
    with torch.no_grad():
        ref = O.mask_vrd(x, m, sd, mc, taps)
    model.to("cuda")
    eng = model._get_engine()
    eng.taps = {}
    r = model.run_network([f.cuda() for f in feats], [T] * len(lens), cfg["inference_config"]["topk"], want_masks=True)
    torch.cuda.synchronize()
    mf = m.float()
    pairs = [("so_in", ["s_in", "o_in"])] + \
        [x for i in range(mc["backbone_arch"][1]) for x in ((f"so_stem{i}", [f"s_stem{i}", f"o_stem{i}"]), (f"so_sos{i}", [f"s_sos{i}", f"o_sos{i}"]))] + \
        [(f"e{i}", [f"e{i}"]) for i in range(4)] + [(f"fpn{l}", [f"fpn{l}"]) for l in (3, 2, 1, 0)] + [("mask_features", ["mask_features"])]
    print(f"{'tap':16s} {'max|diff|':>12s} {'scale':>10s}")
    for ename, onames in pairs:
        t, lay, streams = eng.taps[ename]
        Tl = T >> lay.level
        got = rows_to_padded(t, lay, streams, Tl)
        for g, on in zip(got, onames):
            o = taps[on]
            mask_l = m[:, :, ::(1 << lay.level)].float()
            d = ((g - o) * mask_l).abs().max().item()
            print(f"{on:16s} {d:12.3e} {o.abs().max().item():10.3e}")
    for j in range(mc["predictor"]["num_layers"]):
        t = eng.taps[f"dec{j}"][0]
        o = taps[f"dec{j}"]   # (B, 256, Q)
        g = t.view(len(lens), -1, t.shape[1]).transpose(1, 2)
        print(f"{'dec%d' % j:16s} {(g - o).abs().max().item():12.3e} {o.abs().max().item():10.3e}")
    print(f"{'logits':16s} {(r['logits'].cpu() - ref['pred_logits']).abs().max().item():12.3e} {ref['pred_logits'].abs().max().item():10.3e}")
    for i, mk in enumerate(r["masks"]):
        d = (mk.t().cpu() - ref["pred_masks"][i][:, :lens[i]]).abs().max().item()
        print(f"{'masks[%d] L=%d' % (i, lens[i]):16s} {d:12.3e}")


if __name__ == "__main__":
    main()
