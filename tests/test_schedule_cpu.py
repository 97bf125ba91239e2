"""The engine's varlen schedule (layout, pad-column corrections, weight folding) driven by the CPU emulation of the
kernels must reproduce the padded reference forward (through the golden fixtures)."""
import pytest
import torch

from tests import helpers as H
from tests.emu_ops import EmuOps
from vrdone_b200 import synth
from vrdone_b200.engine import Engine, PackedWeights
from vrdone_b200.layout import PackLayout, reference_padded_lengths


@pytest.mark.parametrize("name,idx", [("vidvrd", None), ("vidor", [1, 3, 7, 8]), ("vidor_local", [2, 3]), ("vidor_x", [2, 3])])
def test_emulated_schedule_matches_golden(name, idx):
    fix = H.network_fixture(name)
    cfg, model, sd = H.seeded_model(name, fix["wseed"])
    mc = cfg["model_config"]
    feats = synth.pair_features(mc, fix["lens"], fix["xseed"])
    idx = list(range(len(feats))) if idx is None else idx
    lens = [fix["lens"][i] for i in idx]
    tpads = [fix["tpads"][i] for i in idx]
    lay = PackLayout(lens, tpads, 4, "cpu")
    eng = Engine(PackedWeights(sd, mc, "cpu", torch.float32), EmuOps())
    with torch.no_grad():
        out = eng.forward_packed(lay, [feats[i] for i in idx], None, cfg["inference_config"]["topk"], want_masks=True)
    l0 = lay.levels[0]
    for j, i in enumerate(idx):
        assert H.rel_err(out["logits"][j], fix["pred_logits"][i]) < 2e-5
        r0, L = int(l0.off[j]), lens[j]
        assert H.rel_err(out["masks"][r0:r0 + L].t(), fix["pred_masks"][i]) < 2e-5
        # binarised masks / first-last frames are bit-exact against the reference's sigmoid > 0.5
        act = torch.sigmoid(fix["pred_masks"][i]) > 0.5
        for q in range(act.shape[0]):
            nz = torch.nonzero(act[q]).flatten()
            exp = [int(nz[0]), int(nz[-1])] if nz.numel() else [-1, -1]
            assert out["first_last"][j, q].tolist() == exp
    probs = torch.softmax(torch.stack([fix["pred_logits"][i] for i in idx]), -1)[..., 1:]
    ids = torch.topk(probs, cfg["inference_config"]["topk"], dim=-1).indices + 1
    assert torch.equal(out["topk_ids"].long(), ids)


def test_layout_invariants():
    mc = synth.load_config("vidor")["model_config"]
    lens = [512, 1, 2, 3, 700, 64, 65]
    tp = reference_padded_lengths(lens, mc)
    assert tp == [512, 512, 512, 512, 704, 512, 512]
    lay = PackLayout(lens, tp, 4, "cpu")
    for l, lev in enumerate(lay.levels):
        assert lev.R % 128 == 0
        rs = lev.row_seq
        assert rs[0] == -1
        for i, L in enumerate(lens):
            ll = (L + (1 << l) - 1) >> l
            assert int(lev.len[i]) == ll
            off = int(lev.off[i])
            assert torch.all(rs[off:off + ll] == i) and rs[off - 1] == -1 and rs[off + ll] == -1
            assert int(lev.haspad[i]) == int(ll < tp[i] >> l)
    # 200-pair slices decide the long padding independently
    lens = [600] + [10] * 199 + [900]
    tp = reference_padded_lengths(lens, mc)
    assert tp[0] == 640 and tp[1] == 512 and tp[200] == 960


def test_emulated_bf16_schedule_uses_fused_layernorm_epilogues():
    """The bf16 schedule routes the embedding convs through gemm_ln and the encoder blocks' attention projection through gemm_res_ln
    (LayerNorm as the GEMM's epilogue); with the CPU emulation of those ops it still reproduces the reference within bf16 accuracy."""
    fix = H.network_fixture("vidor")
    cfg, model, sd = H.seeded_model("vidor", fix["wseed"])
    mc = cfg["model_config"]
    feats = synth.pair_features(mc, fix["lens"], fix["xseed"])
    idx = [1, 3]
    lens = [fix["lens"][i] for i in idx]
    tpads = [fix["tpads"][i] for i in idx]
    lay = PackLayout(lens, tpads, 4, "cpu")
    ops = EmuOps()
    eng = Engine(PackedWeights(sd, mc, "cpu", torch.bfloat16), ops)
    with torch.no_grad():
        out = eng.forward_packed(lay, [feats[i] for i in idx], None, cfg["inference_config"]["topk"], want_masks=True)
    n_conv, n_stem, n_branch = mc["backbone_arch"]
    assert ops.calls.count("gemm_ln") == n_conv                      # one stream of embedding convs (no clip features in vidor)
    assert ops.calls.count("gemm_res_ln") == n_stem + n_branch       # one per encoder block
    for j, i in enumerate(idx):
        assert H.rel_err(out["logits"][j], fix["pred_logits"][i]) < 4e-2
