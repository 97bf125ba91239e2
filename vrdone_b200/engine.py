"""Host-side schedule of the MaskVRD network forward over packed varlen rows.

This file decides WHICH kernels run in which order on which buffers; the arithmetic lives in the CUDA
kernels behind ``ops`` (``cuda_ops.CudaOps`` -> C-ABI in ``csrc/``).  It follows the data flow of the
reference forward (models/backbones.py:154-248 / 323-436, models/blocks.py:1070-1080,
models/local_transformer.py:807-835, models/fpns.py:229-257, models/predictor.py:85-115) re-expressed
for the token-major varlen layout of ``layout.py``.

Weight re-parameterisations done once per checkpoint, all exact algebra:
  * AffineDropPath ``scale`` is folded into the rows of the preceding 1x1 conv (proj / mlp.3),
  * 1/sqrt(head_dim) is folded into the query projection,
  * k=3 conv weights (out, in, 3) are stored tap-major as [out, 3*in] for the row-shifted GEMM,
  * the constant that the first pad column feeds into ``visual_embd[1]`` / ``clip_embd[1]`` (ReLU of the
    previous LayerNorm bias, appendix B) becomes a per-row correction vector ``W[:, :, 2] @ relu(beta)``.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

from .layout import LevelLayout, PackLayout

ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2
ROW_BUCKET = 8192


class PackedWeights:
    """Kernel-layout copies of a reference-schema ``state_dict`` on ``device``.

    ``adt`` is the dtype of GEMM operands (torch.float32 or torch.bfloat16); LayerNorm parameters, biases,
    depthwise weights and everything consumed by CUDA-core kernels stay fp32."""

    def __init__(self, sd: Dict[str, torch.Tensor], mc: dict, device, adt: torch.dtype):
        self.mc, self.adt, self.device = mc, adt, device
        self.t: Dict[str, torch.Tensor] = {}
        f32 = {k: v.detach().to(torch.float32).cpu() for k, v in sd.items()}
        self._src = f32
        c = mc["embd_dim"]
        clip = bool(mc.get("with_clip_feature", False))
        n_conv, n_stem, n_branch = mc["backbone_arch"]
        bb = "backbone"
        embeds = [("visual_embd", "visual_embd_norm")] + ([("clip_embd", "clip_embd_norm")] if clip else [])
        for conv, norm in embeds:
            for i in range(n_conv):
                w = f32[f"{bb}.{conv}.{i}.conv.weight"]
                self._gemm_w(f"{bb}.{conv}.{i}", w.permute(0, 2, 1).reshape(w.shape[0], -1), None)
                self._ln(f"{bb}.{norm}.{i}")
                if i > 0:
                    beta = f32[f"{bb}.{norm}.{i - 1}.bias"].flatten()
                    self._put(f"{bb}.{conv}.{i}.corr", w[:, :, 2] @ torch.relu(beta))
        if clip:
            self._pw_mlp(f"{bb}.visual_clip_fuse", 2)
        for name in ("bbox_entity_embd", "bbox_so_embd"):
            w = f32[f"{bb}.{name}.conv.weight"]
            self._put(f"{bb}.{name}.w", w.permute(2, 1, 0).reshape(-1, w.shape[0]))  # [3*cin, N], rows tap-major
            self._put(f"{bb}.{name}.b", f32[f"{bb}.{name}.conv.bias"])
        self._ln(f"{bb}.bbox_entity_norm")
        for name in ("visual_bbox_fuse", "so_fuse", "so_visual_bbox_fuse"):
            self._pw_mlp(f"{bb}.{name}", 2)
        for i in range(n_stem):
            self._encoder_block(f"{bb}.stem.{i}", mc["n_head"])
            for side in ("s_attn", "o_attn"):
                p = f"{bb}.{side}.{i}"
                self._ln(p + ".ln1")
                self._ln(p + ".ln2")
                self._conv_attention(p + ".self_attn", mc["fuse_head"], f32[p + ".drop_path_attn1.scale"])
                self._conv_attention(p + ".multihead_attn", mc["fuse_head"], f32[p + ".drop_path_attn2.scale"])
        self._ln(f"{bb}.s_fuse_norm")
        self._ln(f"{bb}.o_fuse_norm")
        for i in range(n_branch):
            self._encoder_block(f"{bb}.branch.{i}", mc["n_head"])
        n_lev = n_branch + 1
        for l in range(n_lev):
            self._ln(f"neck.input_norms.{l}")
            self._ln(f"neck.fpn_norms.{l}")
            w = f32[f"neck.fpn_convs.{l}.conv.weight"]
            # tap-major [3, channels]; the top level's grouped weight (F, 2, 3) is indexed by INPUT channel 2c+j
            self._put(f"neck.fpn_convs.{l}.w", w.reshape(-1, 3).t())
            if l < n_lev - 1:
                self._ln(f"neck.lateral_norms.{l}")
                self._gemm_w(f"neck.lateral_convs.{l}", f32[f"neck.lateral_convs.{l}.conv.weight"][:, :, 0], None)
        self._put("neck.mask_features.w", f32["neck.mask_features.conv.weight"].reshape(-1, 3).t())
        self._put("neck.mask_features.b", f32["neck.mask_features.conv.bias"])
        pc = mc["predictor"]
        self._ln("predictor.input_norm")
        self._gemm_w("predictor.input_proj", f32["predictor.input_proj.weight"][:, :, 0], f32["predictor.input_proj.bias"])
        self._put("predictor.query_pos", f32["predictor.query_embed.weight"])   # [Q, 256]
        hs_scale = 1.0 / math.sqrt(pc["n_embd"] // pc["n_head"])
        for j in range(pc["num_layers"]):
            p = f"predictor.transformer.decoder.layers.{j}"
            for n in ("ln1", "ln2", "ln3"):
                self._ln(f"{p}.{n}")
            sa = p + ".self_attn"
            sc1 = f32[p + ".drop_path_attn1.scale"].flatten()
            self._gemm_w(sa + ".query", f32[sa + ".query.weight"][:, :, 0] * hs_scale, f32[sa + ".query.bias"] * hs_scale)
            self._gemm_w(sa + ".key", f32[sa + ".key.weight"][:, :, 0], f32[sa + ".key.bias"])
            self._gemm_w(sa + ".value", f32[sa + ".value.weight"][:, :, 0], f32[sa + ".value.bias"])
            self._gemm_w(sa + ".proj", f32[sa + ".proj.weight"][:, :, 0] * sc1[:, None], f32[sa + ".proj.bias"] * sc1)
            self._conv_attention(p + ".multihead_attn", pc["n_head"], f32[p + ".drop_path_attn2.scale"])
            self._mlp(p, f32[p + ".drop_path_mlp.scale"])
        self._ln("predictor.transformer.decoder.norm")
        # class head: N (= K+1) padded up to a multiple of 64 with zero rows so it fits the GEMM N tile
        wc, bc = f32["predictor.class_embed.weight"][:, :, 0], f32["predictor.class_embed.bias"]
        self.n_cls = wc.shape[0]
        self.n_cls_pad = (self.n_cls + 63) // 64 * 64
        wcp = torch.zeros(self.n_cls_pad, wc.shape[1])
        wcp[: self.n_cls] = wc
        bcp = torch.zeros(self.n_cls_pad)
        bcp[: self.n_cls] = bc
        self._gemm_w("predictor.class_embed", wcp, bcp)
        self._pw_mlp("predictor.mask_embed", 3)
        del self._src

    # -- helpers ------------------------------------------------------------------------------
    def _put(self, name, t, dtype=torch.float32):
        self.t[name] = t.to(dtype).contiguous().to(self.device)

    def _gemm_w(self, name, w, b):
        self._put(name + ".W", w, self.adt)
        if b is not None:
            self._put(name + ".b", b)

    def _ln(self, p):
        self._put(p + ".g", self._src[p + ".weight"].flatten())
        self._put(p + ".be", self._src[p + ".bias"].flatten())

    def _pw_mlp(self, p, n):
        for i in range(n):
            self._gemm_w(f"{p}.{i}", self._src[f"{p}.layers.{i}.weight"][:, :, 0], self._src[f"{p}.layers.{i}.bias"])

    def _conv_attention(self, p, n_head, out_scale):
        s = self._src
        c = s[p + ".query.weight"].shape[0]
        q_scale = 1.0 / math.sqrt(c // n_head)
        sc = out_scale.flatten()
        for n in ("query", "key", "value"):
            w = s[f"{p}.{n}_conv.conv.weight"]
            self._put(f"{p}.{n}_conv.w", w.reshape(w.shape[0], -1).t())   # [3, C] tap-major ([1, C] for the predictor's query conv)
            self._ln(f"{p}.{n}_norm")
            f = q_scale if n == "query" else 1.0
            self._gemm_w(f"{p}.{n}", s[f"{p}.{n}.weight"][:, :, 0] * f, s[f"{p}.{n}.bias"] * f)
        self._gemm_w(p + ".proj", s[p + ".proj.weight"][:, :, 0] * sc[:, None], s[p + ".proj.bias"] * sc)

    def _mlp(self, p, out_scale):
        s = self._src
        sc = out_scale.flatten()
        self._gemm_w(p + ".mlp.0", s[p + ".mlp.0.weight"][:, :, 0], s[p + ".mlp.0.bias"])
        self._gemm_w(p + ".mlp.3", s[p + ".mlp.3.weight"][:, :, 0] * sc[:, None], s[p + ".mlp.3.bias"] * sc)

    def _encoder_block(self, p, n_head):
        self._ln(p + ".ln1")
        self._ln(p + ".ln2")
        self._conv_attention(p + ".attn", n_head, self._src[p + ".drop_path_attn.scale"])
        self._mlp(p, self._src[p + ".drop_path_mlp.scale"])


class Engine:
    """Runs the network for one packed batch of pairs.  ``ops`` provides the kernels (see cuda_ops.CudaOps)."""

    def __init__(self, weights: PackedWeights, ops):
        self.w, self.ops, self.mc = weights, ops, weights.mc
        self.adt, self.device = weights.adt, weights.device
        self.taps: Optional[Dict[str, torch.Tensor]] = None   # set to {} to record intermediates (tests only)

    # -- buffer helpers -------------------------------------------------------------------------
    def _buf(self, rows, cols, dtype):
        # rows are rounded up to a bucket so that the caching allocator sees the same block sizes for every chunk / video
        # (no cudaMalloc / cudaFree churn in steady state); kernels only ever touch the first ``rows`` rows.
        bucket = (rows + ROW_BUCKET - 1) // ROW_BUCKET * ROW_BUCKET
        return torch.empty(bucket, cols, dtype=dtype, device=self.device)[:rows]

    @property
    def fuse_embed_ln(self) -> bool:
        """LayerNorm + ReLU of the embedding convs as the GEMM's epilogue (the launcher option ``embed_ln``, shared with the native
        schedule so that both run the same kernels); kernel providers without options (the CPU emulation of the tests) fuse."""
        get = getattr(self.ops, "get_option", None)
        return hasattr(self.ops, "gemm_ln") and (get is None or get("embed_ln") != 0)

    def _W(self, name):
        return self.w.t[name]

    def _b(self, name):
        return self.w.t.get(name)

    def _gemm(self, a, wname, out, lay=None, streams=1, **kw):
        self.ops.gemm(a, self._W(wname + ".W"), out, bias=self._b(wname + ".b"), lay=lay, streams=streams, **kw)
        return out

    def _tap(self, name, t, lay, streams=1):
        if self.taps is not None:
            self.taps[name] = (t.detach().float().cpu().clone(), lay, streams)

    # -- sub-networks ---------------------------------------------------------------------------
    def _attention_qkv(self, p, srcs, lay_in, lay_out, stride, streams):
        """srcs: list of (x, pre_ln_prefix or None, [branch names]) -> dict name -> projected [rows, C] adt."""
        ops, C = self.ops, srcs[0][0].shape[1]
        rows = streams * lay_out.R
        out = {}
        for x, pre, names in srcs:
            pre_gb = (self._W(pre + ".g"), self._W(pre + ".be")) if pre is not None else None
            branches, bufs = [], {}
            for n, use_pre in names:
                bufs[n] = self._buf(rows, C, self.adt)
                branches.append((self._W(f"{p}.{n}_conv.w"), use_pre, self._W(f"{p}.{n}_norm.g"),
                                 self._W(f"{p}.{n}_norm.be"), bufs[n]))
            ops.dwconv_ln(x, lay_in, lay_out, stride, pre_gb, branches, streams)
            for n, _ in names:
                out[n] = self._gemm(bufs[n], f"{p}.{n}", self._buf(rows, C, self.adt), lay_out, streams)
        return out

    def _encoder_block(self, x, p, lay_in: LevelLayout, lay_out: LevelLayout, stride, streams, n_head, window):
        ops, C = self.ops, x.shape[1]
        rows = streams * lay_out.R
        qkv = self._attention_qkv(p + ".attn", [(x, p + ".ln1", [("query", True), ("key", True), ("value", True)])],
                                  lay_in, lay_out, stride, streams)
        a = self._buf(rows, C, self.adt)
        ops.window_attn(qkv["query"], qkv["key"], qkv["value"], a, lay_out, n_head, window // 2, streams)
        if stride == 1:
            skip = x
        else:
            skip = self._buf(rows, C, torch.float32)
            ops.maxpool_skip(x, lay_in, lay_out, skip)
        y = self._buf(rows, C, torch.float32)
        h = self._buf(rows, C, self.adt)
        if self.adt == torch.bfloat16 and C == 512 and hasattr(ops, "gemm_res_ln"):
            # bf16 path: projection + residual + the LayerNorm that feeds the MLP in one launch
            ops.gemm_res_ln(a, self._W(p + ".attn.proj.W"), y, h, (self._W(p + ".ln2.g"), self._W(p + ".ln2.be")),
                            bias=self._b(p + ".attn.proj.b"), res1=skip, lay=lay_out, streams=streams)
        else:
            self._gemm(a, p + ".attn.proj", y, lay_out, streams, res1=skip)
            ops.layernorm(y, self._W(p + ".ln2.g"), self._W(p + ".ln2.be"), h, relu=False, lay=lay_out, streams=streams)
        h2 = self._gemm(h, p + ".mlp.0", self._buf(rows, 4 * C, self.adt), lay_out, streams, act=ACT_GELU)
        return self._gemm(h2, p + ".mlp.3", self._buf(rows, C, torch.float32), lay_out, streams, res1=y)

    def _sos_layer(self, tgt, mem, p, lay: LevelLayout, out, n_head, window):
        """out = tgt + decoder_layer(tgt, mem)  (the reference adds the layer output, which already
        contains tgt, back onto tgt: backbones.py:217-221)."""
        ops, C, R = self.ops, tgt.shape[1], lay.R

        def attend(q, k, v):
            a = self._buf(R, C, self.adt)
            if window is None:
                ops.full_attn(q, k, v, a, lay, n_head)
            else:
                ops.window_attn(q, k, v, a, lay, n_head, window // 2, 1)
            return a

        sa = p + ".self_attn"
        qkv = self._attention_qkv(sa, [(tgt, p + ".ln1", [("query", True), ("key", True), ("value", False)])],
                                  lay, lay, 1, 1)
        a = attend(qkv["query"], qkv["key"], qkv["value"])
        tgt1 = self._gemm(a, sa + ".proj", self._buf(R, C, torch.float32), lay, 1, res1=tgt)
        ca = p + ".multihead_attn"
        qkv = self._attention_qkv(ca, [(tgt1, p + ".ln2", [("query", True)]),
                                       (mem, None, [("key", False), ("value", False)])], lay, lay, 1, 1)
        a = attend(qkv["query"], qkv["key"], qkv["value"])
        self._gemm(a, ca + ".proj", out, lay, 1, res1=tgt1, res2=tgt)

    def _embed(self, x, conv, norm, lay0, out):
        """2 x (k=3 conv-as-GEMM -> LN -> ReLU) on stacked s/o rows; result written into ``out`` (a column slice)."""
        ops, w, C = self.ops, self.w, self.mc["embd_dim"]
        n_conv = self.mc["backbone_arch"][0]
        rows = 2 * lay0.R
        for i in range(n_conv):
            corr = self._b(f"backbone.{conv}.{i}.corr")
            nx = out if i == n_conv - 1 else self._buf(rows, C, self.adt)
            ln = (self._W(f"backbone.{norm}.{i}.g"), self._W(f"backbone.{norm}.{i}.be"))
            if self.fuse_embed_ln and self.adt == torch.bfloat16 and C == 512:
                # bf16 path: LayerNorm + ReLU are the GEMM's epilogue (the tile spans the 512-channel row); no fp32 conv output
                ops.gemm_ln(x, self._W(f"backbone.{conv}.{i}.W"), nx, ln, bias=self._b(f"backbone.{conv}.{i}.b"), taps=3, corr=corr,
                            relu=True, lay=lay0, streams=2)
            else:
                e = self._buf(rows, C, torch.float32)
                self._gemm(x, f"backbone.{conv}.{i}", e, lay0, 2, taps=3, corr=corr)
                ops.layernorm(e, ln[0], ln[1], nx, relu=True, lay=lay0, streams=2)
            x = nx
        return out

    # -- full network ---------------------------------------------------------------------------
    def forward_packed(self, lay: PackLayout, pair_ptrs, pair_strides, topk: int, want_masks: bool = False, after_pack=None,
                       token_major: bool = False):
        """pair_ptrs: int64[B] device addresses of the fp32 (C, L_i) inputs; pair_strides: int64[B, 2] element strides
        (channel, time).  Returns dict(logits [B,Q,K+1] f32, topk_scores/topk_ids [B,Q,topk], first_last [B,Q,2] int32,
        masks [R0, Q] f32 or None)."""
        e_top, mf = self.backbone(lay, pair_ptrs, pair_strides, after_pack=after_pack, token_major=token_major)
        return self.predict(lay, e_top, mf, topk, want_masks)

    def predict(self, lay, e_top, mf, topk: int, want_masks: bool = False):
        """Query decoder + heads over the coarsest level ``e_top`` [R_top, C] and the mask features ``mf`` [R_0, F] of a batch
        whose levels 0 and top are described by ``lay`` (a PackLayout, or a MergedLayout over several backbone chunks)."""
        if hasattr(self.ops, "bind_stream"):
            self.ops.bind_stream()
        return self._predictor(e_top, mf, lay, topk, want_masks)

    def backbone(self, lay: PackLayout, pair_ptrs, pair_strides, after_pack=None, token_major: bool = False):
        """Pack + embeddings + stem / SOS + fuse + pyramid + FPN of one chunk of pairs -> (coarsest level [R_top, C] fp32, mask
        features [R_0, fpn_dim] fp32)."""
        ops, mc, w = self.ops, self.mc, self.w
        if hasattr(ops, "bind_stream"):
            ops.bind_stream()
        C, adt = mc["embd_dim"], self.adt
        nv, nbe, nbs = mc["visual_dim"], mc["bbox_entity_dim"], mc["bbox_so_dim"]
        clip = bool(mc.get("with_clip_feature", False))
        nc = mc["clip_dim"] if clip else 0
        n_conv, n_stem, n_branch = mc["backbone_arch"]
        win = mc["n_mha_win_size"]
        L = lay.levels
        l0, R0 = L[0], L[0].R
        bb = "backbone"

        # 1. gather + split the ragged (C, L) pair tensors into token-major operand matrices
        vis = self._buf(2 * R0, nv, adt)
        clp = self._buf(2 * R0, nc, adt) if clip else None
        bso = self._buf(R0, 8, torch.float32)
        bent = self._buf(2 * R0, 8, torch.float32)
        ops.pack_pairs(pair_ptrs, pair_strides, l0, nv, nc, nbs, nbe, vis, clp, bso, bent, token_major=token_major)
        if after_pack is not None:
            after_pack()     # the pair tensors (or their staging buffer) are no longer needed once this kernel has run

        # 2. embedding convs, entity-box embedding, fuse MLPs  (s rows then o rows; weights are shared)
        cat = self._buf(2 * R0, 2 * C, adt)
        if clip:
            vcat = self._buf(2 * R0, 2 * C, adt)
            self._embed(vis, "visual_embd", "visual_embd_norm", l0, vcat[:, :C])
            self._embed(clp, "clip_embd", "clip_embd_norm", l0, vcat[:, C:])
            h = self._gemm(vcat, bb + ".visual_clip_fuse.0", self._buf(2 * R0, C, adt), l0, 2, act=ACT_GELU)
            self._gemm(h, bb + ".visual_clip_fuse.1", cat[:, :C], l0, 2)
        else:
            self._embed(vis, "visual_embd", "visual_embd_norm", l0, cat[:, :C])
        ops.small_conv(bent, nbe, self._W(bb + ".bbox_entity_embd.w"), self._W(bb + ".bbox_entity_embd.b"),
                       (self._W(bb + ".bbox_entity_norm.g"), self._W(bb + ".bbox_entity_norm.be")), True,
                       cat[:, C:], l0, 2)
        h = self._gemm(cat, bb + ".visual_bbox_fuse.0", self._buf(2 * R0, C, adt), l0, 2, act=ACT_GELU)
        x = self._gemm(h, bb + ".visual_bbox_fuse.1", self._buf(2 * R0, C, torch.float32), l0, 2)
        self._tap("so_in", x, l0, 2)

        # 3. stem blocks (shared weights, both streams at once) interleaved with subject-object synergy layers
        sos_window = win if mc["use_local"] else None
        for i in range(n_stem):
            x = self._encoder_block(x, f"{bb}.stem.{i}", l0, l0, 1, 2, mc["n_head"], win)
            self._tap(f"so_stem{i}", x, l0, 2)
            s, o = x[:R0], x[R0:]
            xn = self._buf(2 * R0, C, torch.float32)
            self._sos_layer(s, o, f"{bb}.s_attn.{i}", l0, xn[:R0], mc["fuse_head"], sos_window)
            self._sos_layer(o, s, f"{bb}.o_attn.{i}", l0, xn[R0:], mc["fuse_head"], sos_window)
            x = xn
            self._tap(f"so_sos{i}", x, l0, 2)

        # 4. fuse the two streams and the relative-box embedding into one
        cat2 = self._buf(R0, 2 * C, adt)
        ops.layernorm(x[:R0], self._W(bb + ".s_fuse_norm.g"), self._W(bb + ".s_fuse_norm.be"), cat2[:, :C], relu=False,
                      lay=l0, streams=1)
        ops.layernorm(x[R0:], self._W(bb + ".o_fuse_norm.g"), self._W(bb + ".o_fuse_norm.be"), cat2[:, C:], relu=False,
                      lay=l0, streams=1)
        h = self._gemm(cat2, bb + ".so_fuse.0", self._buf(R0, C, adt), l0, 1, act=ACT_GELU)
        cat3 = self._buf(R0, 2 * C, adt)
        self._gemm(h, bb + ".so_fuse.1", cat3[:, :C], l0, 1)
        ops.small_conv(bso, nbs, self._W(bb + ".bbox_so_embd.w"), self._W(bb + ".bbox_so_embd.b"), None, False,
                       cat3[:, C:], l0, 1)
        h = self._gemm(cat3, bb + ".so_visual_bbox_fuse.0", self._buf(R0, C, adt), l0, 1, act=ACT_GELU)
        e = [self._gemm(h, bb + ".so_visual_bbox_fuse.1", self._buf(R0, C, torch.float32), l0, 1)]

        # 5. stride-2 pyramid
        for i in range(n_branch):
            e.append(self._encoder_block(e[i], f"{bb}.branch.{i}", L[i], L[i + 1], mc["scale_factor"], 1,
                                         mc["n_head"], win))
        for i, f in enumerate(e):
            self._tap(f"e{i}", f, L[i])

        # 6. top-down FPN -> mask features at full temporal resolution
        F = mc["fpn_dim"]
        top = n_branch
        y = self._buf(L[top].R, F, torch.float32)
        ops.fpn_top(e[top], L[top], (self._W(f"neck.input_norms.{top}.g"), self._W(f"neck.input_norms.{top}.be")),
                    self._W(f"neck.fpn_convs.{top}.w"),
                    (self._W(f"neck.fpn_norms.{top}.g"), self._W(f"neck.fpn_norms.{top}.be")), y)
        self._tap(f"fpn{top}", y, L[top])
        for l in range(top - 1, -1, -1):
            n = self._buf(L[l].R, C, adt)
            ops.layernorm(e[l], self._W(f"neck.input_norms.{l}.g"), self._W(f"neck.input_norms.{l}.be"), n, relu=False,
                          lay=L[l], streams=1)
            cur = self._gemm(n, f"neck.lateral_convs.{l}", self._buf(L[l].R, F, torch.float32), L[l], 1)
            yl = self._buf(L[l].R, F, torch.float32)
            ops.fpn_level(cur, y, L[l], L[l + 1],
                          (self._W(f"neck.lateral_norms.{l}.g"), self._W(f"neck.lateral_norms.{l}.be")),
                          self._W(f"neck.fpn_norms.{l + 1}.be"), self._W(f"neck.fpn_convs.{l}.w"),
                          (self._W(f"neck.fpn_norms.{l}.g"), self._W(f"neck.fpn_norms.{l}.be")), yl)
            y = yl
            self._tap(f"fpn{l}", y, L[l])
        mf = self._buf(R0, F, torch.float32)
        ops.mask_features(y, l0, self._W("neck.fpn_norms.0.be"), self._W("neck.mask_features.w"),
                          self._W("neck.mask_features.b"), mf)
        self._tap("mask_features", mf, l0)

        return e[top], mf

    def _predictor(self, e_top, mf, lay: PackLayout, topk, want_masks):
        ops, mc, adt = self.ops, self.mc, self.adt
        pc = mc["predictor"]
        D, Q, nh, n_layers = pc["n_embd"], pc["num_queries"], pc["n_head"], pc["num_layers"]
        lt = lay.levels[-1]
        B = lay.B
        n = self._buf(lt.R, e_top.shape[1], adt)
        ops.layernorm(e_top, self._W("predictor.input_norm.g"), self._W("predictor.input_norm.be"), n, relu=False,
                      lay=lt, streams=1)
        src = self._gemm(n, "predictor.input_proj", self._buf(lt.R, D, torch.float32), lt, 1)
        self._tap("src", src, lt)
        MQ = (B * Q + 127) // 128 * 128
        tgt = torch.zeros(MQ, D, dtype=torch.float32, device=self.device)
        pos = self._W("predictor.query_pos")
        for j in range(n_layers):
            p = f"predictor.transformer.decoder.layers.{j}"
            # self-attention among the Q queries of each pair: q = k = LN1(tgt) + pos, v = tgt
            qk = self._buf(MQ, D, adt)
            ops.query_ln(tgt, (self._W(p + ".ln1.g"), self._W(p + ".ln1.be")), pos, Q, B * Q, None, None, qk)
            tv = self._buf(MQ, D, adt)
            ops.query_ln(tgt, None, None, Q, B * Q, None, None, tv)
            sa = p + ".self_attn"
            q = self._gemm(qk, sa + ".query", self._buf(MQ, D, adt))
            k = self._gemm(qk, sa + ".key", self._buf(MQ, D, adt))
            v = self._gemm(tv, sa + ".value", self._buf(MQ, D, adt))
            a = self._buf(MQ, D, adt)
            ops.query_self_attn(q, k, v, a, B, Q, nh)
            tgt1 = self._gemm(a, sa + ".proj", self._buf(MQ, D, torch.float32), res1=tgt)
            # cross-attention to the coarsest pyramid level
            ca = p + ".multihead_attn"
            kv = self._attention_qkv(ca, [(src, None, [("key", False), ("value", False)])], lt, lt, 1, 1)
            h = self._buf(MQ, D, adt)
            ops.query_ln(tgt1, (self._W(p + ".ln2.g"), self._W(p + ".ln2.be")), pos, Q, B * Q,
                         self._W(ca + ".query_conv.w"), (self._W(ca + ".query_norm.g"), self._W(ca + ".query_norm.be")), h)
            q = self._gemm(h, ca + ".query", self._buf(MQ, D, adt))
            a = self._buf(MQ, D, adt)
            ops.query_cross_attn(q, kv["key"], kv["value"], a, lt, Q, nh)
            tgt2 = self._gemm(a, ca + ".proj", self._buf(MQ, D, torch.float32), res1=tgt1)
            # FFN
            h = self._buf(MQ, D, adt)
            ops.query_ln(tgt2, (self._W(p + ".ln3.g"), self._W(p + ".ln3.be")), None, Q, B * Q, None, None, h)
            h2 = self._gemm(h, p + ".mlp.0", self._buf(MQ, pc["n_hidden"], adt), act=ACT_GELU)
            tgt = self._gemm(h2, p + ".mlp.3", self._buf(MQ, D, torch.float32), res1=tgt2)
            if self.taps is not None:
                self.taps[f"dec{j}"] = (tgt[: B * Q].detach().float().cpu().clone(), None, 1)
        hs = self._buf(MQ, D, adt)
        ops.query_ln(tgt, (self._W("predictor.transformer.decoder.norm.g"), self._W("predictor.transformer.decoder.norm.be")),
                     None, Q, B * Q, None, None, hs)
        logits = self._gemm(hs, "predictor.class_embed", self._buf(MQ, self.w.n_cls_pad, torch.float32))
        m = self._gemm(hs, "predictor.mask_embed.0", self._buf(MQ, D, adt), act=ACT_GELU)
        m = self._gemm(m, "predictor.mask_embed.1", self._buf(MQ, D, adt), act=ACT_GELU)
        me = self._gemm(m, "predictor.mask_embed.2", self._buf(MQ, D, torch.float32))
        l0 = lay.levels[0]
        masks = self._buf(l0.R, Q, torch.float32) if want_masks else None
        first_last = torch.empty(B, Q, 2, dtype=torch.int32, device=self.device)
        ops.mask_logits(me, mf, l0, Q, masks, first_last)
        scores = torch.empty(B * Q, topk, dtype=torch.float32, device=self.device)
        ids = torch.empty(B * Q, topk, dtype=torch.int32, device=self.device)
        ops.softmax_topk(logits, B * Q, self.w.n_cls, topk, scores, ids)
        return {"logits": logits[: B * Q, : self.w.n_cls].view(B, Q, -1), "topk_scores": scores.view(B, Q, topk),
                "topk_ids": ids.view(B, Q, topk), "first_last": first_last, "masks": masks}
