"""Parameter containers that reproduce the reference checkpoint schema (names, shapes, default init).

The reference model is an ``nn.Module`` tree whose ``state_dict`` keys are the checkpoint contract
(/root/reference/eval.py:117-134).  The classes below are *containers only*:
they own parameters with the same names, shapes and default initialisation as the reference's
sub-modules (models/blocks.py:63-158, 656-744, 992-1068, 1134-1149; models/local_transformer.py:
13-31, 69-142, 289-376, 625-768, 838-954; models/backbones.py:12-146, 250-320; models/fpns.py:145-226;
models/predictor.py:16-83) and have no forward of their own -- the forward pass is the CUDA engine in
``engine.py``.
"""
from __future__ import annotations

import math

import torch
from torch import nn


class _NoForward(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover - containers are never called
        raise RuntimeError("vrdone_b200 parameter containers have no forward; call MaskVRD(input_data)")


class ChannelNorm(_NoForward):
    """LayerNorm over channels of (B, C, T): weight/bias shaped (1, C, 1)."""

    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(1, c, 1))
        self.bias = nn.Parameter(torch.zeros(1, c, 1))


class PathScale(_NoForward):
    """AffineDropPath: learnable per-channel ``scale`` (1, C, 1), init 1e-4, applied in eval."""

    def __init__(self, c, init=1e-4):
        super().__init__()
        self.scale = nn.Parameter(init * torch.ones(1, c, 1))


class MaskedConv(_NoForward):
    """``.conv`` = Conv1d with odd kernel and same padding; bias (if any) starts at zero."""

    def __init__(self, cin, cout, k, stride=1, groups=1, bias=True):
        super().__init__()
        self.stride = stride
        self.conv = nn.Conv1d(cin, cout, k, stride, k // 2, 1, groups, bias)
        if bias:
            nn.init.zeros_(self.conv.bias)


class PointwiseMLP(_NoForward):
    """ConvMLP: ``layers.{i}`` = Conv1d(k=1), GELU between, zero biases."""

    def __init__(self, cin, hidden, cout, n):
        super().__init__()
        dims = [cin] + [hidden] * (n - 1) + [cout]
        self.layers = nn.ModuleList(nn.Conv1d(a, b, 1) for a, b in zip(dims[:-1], dims[1:]))


class PlainAttention(_NoForward):
    """MaskedMHA(_QKV): query/key/value/proj 1x1 only."""

    def __init__(self, c, n_head):
        super().__init__()
        self.n_head = n_head
        self.key = nn.Conv1d(c, c, 1)
        self.query = nn.Conv1d(c, c, 1)
        self.value = nn.Conv1d(c, c, 1)
        self.proj = nn.Conv1d(c, c, 1)


class ConvAttention(_NoForward):
    """(Local)MaskedMHCA(_QKV): depthwise conv + LN + 1x1 for each of q/k/v, then proj."""

    def __init__(self, c, n_head, q_kernel=3, kv_kernel=3, stride=1, window=None):
        super().__init__()
        self.n_head, self.window, self.stride = n_head, window, stride
        self.query_conv = MaskedConv(c, c, q_kernel, stride, groups=c, bias=False)
        self.query_norm = ChannelNorm(c)
        self.key_conv = MaskedConv(c, c, kv_kernel, stride, groups=c, bias=False)
        self.key_norm = ChannelNorm(c)
        self.value_conv = MaskedConv(c, c, kv_kernel, stride, groups=c, bias=False)
        self.value_norm = ChannelNorm(c)
        self.key = nn.Conv1d(c, c, 1)
        self.query = nn.Conv1d(c, c, 1)
        self.value = nn.Conv1d(c, c, 1)
        self.proj = nn.Conv1d(c, c, 1)


def _mlp(c, hidden):
    # indices 0 and 3 carry parameters, as in the reference's Sequential(conv, act, drop, conv, drop)
    return nn.Sequential(nn.Conv1d(c, hidden, 1), nn.GELU(), nn.Identity(), nn.Conv1d(hidden, c, 1), nn.Identity())


class EncoderBlock(_NoForward):
    """TransformerBlock: ln1, windowed conv-attention, ln2, 4x MLP, two path scales."""

    def __init__(self, c, n_head, window, stride, path_drop):
        super().__init__()
        self.ln1 = ChannelNorm(c)
        self.ln2 = ChannelNorm(c)
        self.attn = ConvAttention(c, n_head, stride=stride, window=window)
        self.mlp = _mlp(c, 4 * c)
        if path_drop > 0.0:
            self.drop_path_attn = PathScale(c)
            self.drop_path_mlp = PathScale(c)
        else:
            raise NotImplementedError("droppath == 0 (no AffineDropPath scale) is not used by any shipped config")


class DecoderLayer(_NoForward):
    """MaskedConvTransformerDecoderLayer: self-attention + cross-attention (+ FFN)."""

    def __init__(self, c, n_head, hidden=None, path_drop=0.1, qx_stride=0, kv_stride=1, with_ffn=True,
                 use_local=False, window=None):
        super().__init__()
        if path_drop <= 0.0:
            raise NotImplementedError("path_drop == 0 is not used by any shipped config")
        self.with_ffn = with_ffn
        self.ln1 = ChannelNorm(c)
        self.ln2 = ChannelNorm(c)
        win = window if use_local else None
        if qx_stride == 0:
            if use_local:
                raise NotImplementedError("windowed attention without conv (fuse_qx_stride == 0 with use_local)")
            self.self_attn = PlainAttention(c, n_head)
            q_kernel = 1
        else:
            if qx_stride != 1:
                raise NotImplementedError("strided SOS attention (fuse_qx_stride > 1)")
            self.self_attn = ConvAttention(c, n_head, window=win)
            q_kernel = 3
        if kv_stride != 1:
            raise NotImplementedError("cross-attention kv stride other than 1")
        self.multihead_attn = ConvAttention(c, n_head, q_kernel=q_kernel, kv_kernel=3, window=win)
        self.drop_path_attn1 = PathScale(c)
        self.drop_path_attn2 = PathScale(c)
        if with_ffn:
            self.ln3 = ChannelNorm(c)
            self.mlp = _mlp(c, hidden if hidden is not None else 4 * c)
            self.drop_path_mlp = PathScale(c)


def _zero_conv_biases(module):
    for m in module.modules():
        if isinstance(m, (nn.Conv1d, nn.Linear)) and m.bias is not None:
            nn.init.zeros_(m.bias)


class Backbone(_NoForward):
    def __init__(self, cfg, with_clip):
        super().__init__()
        c, nv = cfg["embd_dim"], cfg["visual_dim"]
        n_conv, n_stem, n_branch = cfg["backbone_arch"]
        win = cfg["n_mha_win_size"]
        if win <= 1:
            raise NotImplementedError("n_mha_win_size <= 1 (full attention in the stem) is not used by any shipped config")
        if cfg["use_abs_pe"] or cfg["use_rel_pe"]:
            raise NotImplementedError("position encodings are disabled in every shipped config")
        if not cfg["embd_with_ln"] or cfg["fuse_ks"] != 1 or cfg["embd_kernel_size"] != 3:
            raise NotImplementedError("only embd_with_ln=True, fuse_ks=1, embd_kernel_size=3 are supported")
        self.visual_embd = nn.ModuleList(
            MaskedConv(nv if i == 0 else c, c, 3, bias=False) for i in range(n_conv))
        self.visual_embd_norm = nn.ModuleList(ChannelNorm(c) for _ in range(n_conv))
        self.bbox_entity_embd = MaskedConv(cfg["bbox_entity_dim"], c, 3)
        self.bbox_entity_norm = ChannelNorm(c)
        self.visual_bbox_fuse = PointwiseMLP(2 * c, c, c, 2)
        self.stem = nn.ModuleList(EncoderBlock(c, cfg["n_head"], win, 1, cfg["droppath"]) for _ in range(n_stem))

        def sos():
            return DecoderLayer(c, cfg["fuse_head"], path_drop=cfg["fuse_path_drop"], qx_stride=cfg["fuse_qx_stride"],
                                kv_stride=cfg["fuse_kv_stride"], with_ffn=False, use_local=cfg["use_local"], window=win)

        self.s_attn = nn.ModuleList(sos() for _ in range(n_stem))
        self.o_attn = nn.ModuleList(sos() for _ in range(n_stem))
        self.s_fuse_norm = ChannelNorm(c)
        self.o_fuse_norm = ChannelNorm(c)
        self.so_fuse = PointwiseMLP(2 * c, c, c, 2)
        self.bbox_so_embd = MaskedConv(cfg["bbox_so_dim"], c, 3)
        self.so_visual_bbox_fuse = PointwiseMLP(2 * c, c, c, 2)
        self.branch = nn.ModuleList(
            EncoderBlock(c, cfg["n_head"], win, cfg["scale_factor"], cfg["droppath"]) for _ in range(n_branch))
        if with_clip:
            nc = cfg["clip_dim"]
            self.clip_embd = nn.ModuleList(MaskedConv(nc if i == 0 else c, c, 3, bias=False) for i in range(n_conv))
            self.clip_embd_norm = nn.ModuleList(ChannelNorm(c) for _ in range(n_conv))
            self.visual_clip_fuse = PointwiseMLP(2 * c, c, c, 2)
        _zero_conv_biases(self)


class Neck(_NoForward):
    """FPN1D_Fuse: per-level input norm, lateral 1x1 (+norm) below the top, depthwise smoothing conv
    (+norm); the top level's conv is grouped with ``fpn_dim`` groups; ``mask_features`` depthwise+bias."""

    def __init__(self, cfg):
        super().__init__()
        cin, c = cfg["embd_dim"], cfg["fpn_dim"]
        n = cfg["backbone_arch"][-1] + 1
        if cfg["fpn_start_level"] != 0 or not cfg["fpn_with_ln"] or not cfg["fpn_norm_first"]:
            raise NotImplementedError("only fpn_start_level=0, fpn_with_ln=True, fpn_norm_first=True are supported")
        self.input_norms = nn.ModuleList(ChannelNorm(cin) for _ in range(n))
        self.lateral_convs = nn.ModuleList(
            [MaskedConv(cin, c, 1, bias=False) for _ in range(n - 1)] + [None])
        self.lateral_norms = nn.ModuleList([ChannelNorm(c) for _ in range(n - 1)] + [None])
        self.fpn_convs = nn.ModuleList(
            [MaskedConv(c, c, 3, groups=c, bias=False) for _ in range(n - 1)] + [MaskedConv(cin, c, 3, groups=c, bias=False)])
        self.fpn_norms = nn.ModuleList(ChannelNorm(c) for _ in range(n))
        self.mask_features = MaskedConv(c, c, 3, groups=c)


class _Decoder(_NoForward):
    def __init__(self, pc):
        super().__init__()
        c = pc["n_embd"]
        self.layers = nn.ModuleList(
            DecoderLayer(c, pc["n_head"], hidden=pc["n_hidden"], path_drop=pc["path_pdrop"], qx_stride=pc["n_qx_stride"],
                         kv_stride=pc["n_kv_stride"], with_ffn=True) for _ in range(pc["num_layers"]))
        self.norm = ChannelNorm(c)


class _DecoderOnly(_NoForward):
    def __init__(self, pc):
        super().__init__()
        self.decoder = _Decoder(pc)
        _zero_conv_biases(self)


class Predictor(_NoForward):
    def __init__(self, pc):
        super().__init__()
        c = pc["n_embd"]
        if pc["n_qx_stride"] != 0 or pc["n_kv_stride"] != 1:
            raise NotImplementedError("predictor supports n_qx_stride=0, n_kv_stride=1 (all shipped configs)")
        self.transformer = _DecoderOnly(pc)
        self.query_embed = nn.Embedding(pc["num_queries"], c)
        self.input_norm = ChannelNorm(pc["n_input"])
        if pc["n_input"] == c and not pc["enforce_input_project"]:
            raise NotImplementedError("predictor without input projection")
        self.input_proj = nn.Conv1d(pc["n_input"], c, 1)
        nn.init.zeros_(self.input_proj.bias)
        self.class_embed = nn.Conv1d(c, pc["num_classes"] + 1, 1)
        p = pc["cls_prior_prob"]
        nn.init.constant_(self.class_embed.bias, -math.log((1 - p) / p))
        self.mask_embed = PointwiseMLP(c, c, c, 3)
        _zero_conv_biases(self.mask_embed)
