"""CPU oracle for the MaskVRD inference hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain functional PyTorch (CPU, fp32 or fp64) restatement of the reference forward of
``MaskVRD`` on *padded* batches, written from the reference's behaviour, driven only by a
``state_dict`` with the reference's parameter names and the
``model_config`` dict.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module; the product package
``vrdone_b200`` never does.

Parity status: the reference ships no tests, golden vectors or fixtures for this path
(SURVEY.md section 4), so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF run in the
build container: ``tests/golden/make_golden.py`` imports ``/root/reference`` read-only, runs
``MaskVRD._mask_vrd`` / ``forward_test`` on seeded inputs and stores inputs, weights and outputs as
small fixtures under ``tests/golden/``; ``tests/test_oracle.py`` checks this file against them
(and against the live reference whenever ``/root/reference`` is present).

Reference lines each function follows (paths relative to /root/reference):
  layer_norm            models/blocks.py:143-158
  masked_conv           models/blocks.py:91-113
  conv_mlp              models/blocks.py:57-61
  window_attention      models/blocks.py:920-989, models/local_transformer.py:553-623
  full_attention        models/local_transformer.py:144-187 and 33-67
  transformer_block     models/blocks.py:1070-1080
  decoder_layer         models/local_transformer.py:773-835
  backbone              models/backbones.py:154-248 and 323-436
  neck                  models/fpns.py:229-257
  predictor             models/predictor.py:85-115, models/local_transformer.py:875-976
  mask_vrd              models/maskvrd.py:161-167
  preprocessing         models/maskvrd.py:363-414
  forward_test          models/maskvrd.py:201-337
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

EPS = 1e-5


# ----------------------------------------------------------------------------------------------
# primitive layers
# ----------------------------------------------------------------------------------------------
def layer_norm(x, sd, p):
    """Channel LayerNorm of a (B, C, T) tensor, biased variance, eps inside the sqrt."""
    mu = x.mean(dim=1, keepdim=True)
    r = x - mu
    var = (r * r).mean(dim=1, keepdim=True)
    y = r / torch.sqrt(var + EPS)
    return y * sd[p + ".weight"] + sd[p + ".bias"]


def masked_conv(x, mask, sd, p, stride=1, groups=1):
    """Conv1d (odd kernel, 'same' padding) then multiply by the (down-sampled) mask."""
    w = sd[p + ".conv.weight"]
    b = sd.get(p + ".conv.bias", None)
    y = F.conv1d(x, w, b, stride=stride, padding=w.shape[-1] // 2, groups=groups)
    if stride > 1:
        m = mask[:, :, ::stride]  # nearest down-sampling of a length-T mask to T/stride
    else:
        m = mask
    return y * m.to(y.dtype), m


def conv1x1(x, sd, p):
    return F.conv1d(x, sd[p + ".weight"], sd.get(p + ".bias", None))


def conv_mlp(x, sd, p, n_layers):
    for i in range(n_layers):
        x = conv1x1(x, sd, f"{p}.layers.{i}")
        if i < n_layers - 1:
            x = F.gelu(x)
    return x


def _heads(x, n_head):
    b, c, t = x.shape
    return x.view(b, n_head, c // n_head, t).transpose(2, 3)  # (B, nh, T, hs)


def _merge(x):
    b, nh, t, hs = x.shape
    return x.transpose(2, 3).reshape(b, nh * hs, t)


def banded_softmax_pv(q, k, v, key_mask, w):
    """softmax over keys j with |i-j| <= w, 0 <= j < T; invalid keys get -1e4 (as the reference does)
    and rows of invalid queries are zeroed.  q, k, v: (B, nh, T, hs); key_mask: (B, 1, T) bool."""
    b, nh, t, hs = q.shape
    kp = F.pad(k, (0, 0, w, w))
    vp = F.pad(v, (0, 0, w, w))
    ku = kp.unfold(2, 2 * w + 1, 1)  # (B, nh, T, hs, 2w+1)
    vu = vp.unfold(2, 2 * w + 1, 1)
    att = torch.einsum("bhtd,bhtdj->bhtj", q, ku)
    idx = torch.arange(t)[:, None] + torch.arange(-w, w + 1)[None, :]  # (T, 2w+1) absolute key index
    in_range = (idx >= 0) & (idx < t)
    km = F.pad(key_mask[:, 0, :], (w, w), value=False).unfold(1, 2 * w + 1, 1)  # (B, T, 2w+1)
    att = att + (~km)[:, None].to(att.dtype) * (-1e4)
    att = att.masked_fill(~in_range[None, None], float("-inf"))
    att = F.softmax(att, dim=-1)
    att = att.masked_fill(~key_mask[:, 0, :, None][:, None], 0.0)
    return torch.einsum("bhtj,bhtdj->bhtd", att, vu)


def dense_softmax_pv(q, k, v, key_mask):
    att = q @ k.transpose(-2, -1)
    att = att.masked_fill(~key_mask[:, :, None, :], float("-inf"))
    att = F.softmax(att, dim=-1)
    return att @ (v * key_mask[:, :, :, None].to(v.dtype))


def conv_attention(q_in, k_in, v_in, q_mask, kv_mask, sd, p, n_head, stride=1, window=None):
    """depthwise conv -> mask -> LN -> 1x1 for each of q/k/v, attention (windowed if ``window``), proj, mask."""
    q, qm = masked_conv(q_in, q_mask, sd, p + ".query_conv", stride=stride, groups=q_in.shape[1])
    q = layer_norm(q, sd, p + ".query_norm")
    k, km = masked_conv(k_in, kv_mask, sd, p + ".key_conv", stride=stride, groups=k_in.shape[1])
    k = layer_norm(k, sd, p + ".key_norm")
    v, _ = masked_conv(v_in, kv_mask, sd, p + ".value_conv", stride=stride, groups=v_in.shape[1])
    v = layer_norm(v, sd, p + ".value_norm")
    q = _heads(conv1x1(q, sd, p + ".query"), n_head)
    k = _heads(conv1x1(k, sd, p + ".key"), n_head)
    v = _heads(conv1x1(v, sd, p + ".value"), n_head)
    q = q * (1.0 / math.sqrt(q.shape[-1]))
    if window is not None:
        o = banded_softmax_pv(q, k, v, km, window // 2)
    else:
        o = dense_softmax_pv(q, k, v, km)
    o = conv1x1(_merge(o), sd, p + ".proj") * qm.to(o.dtype)
    return o, qm


def plain_attention(q_in, k_in, v_in, q_mask, kv_mask, sd, p, n_head):
    q = _heads(conv1x1(q_in, sd, p + ".query"), n_head)
    k = _heads(conv1x1(k_in, sd, p + ".key"), n_head)
    v = _heads(conv1x1(v_in, sd, p + ".value"), n_head)
    q = q * (1.0 / math.sqrt(q.shape[-1]))
    o = dense_softmax_pv(q, k, v, kv_mask)
    return conv1x1(_merge(o), sd, p + ".proj") * q_mask.to(o.dtype)


def transformer_block(x, mask, sd, p, n_head, window, stride):
    h = layer_norm(x, sd, p + ".ln1")
    a, m = conv_attention(h, h, h, mask, mask, sd, p + ".attn", n_head, stride=stride, window=window)
    mf = m.to(x.dtype)
    skip = x if stride == 1 else F.max_pool1d(x, stride + 1, stride=stride, padding=(stride + 1) // 2)
    y = skip * mf + sd[p + ".drop_path_attn.scale"] * a
    h = layer_norm(y, sd, p + ".ln2")
    h = conv1x1(F.gelu(conv1x1(h, sd, p + ".mlp.0")), sd, p + ".mlp.3") * mf
    return y + sd[p + ".drop_path_mlp.scale"] * h, m


def decoder_layer(tgt, memory, tgt_mask, mem_mask, sd, p, n_head, query_pos=None, conv_self=True,
                  window=None, with_ffn=False):
    """self-attention (q = k = LN1(tgt) [+pos], v = tgt), cross-attention (q = LN2(tgt) [+pos],
    k = v = memory), optional FFN -- each with residual*mask + scale*out."""
    h = layer_norm(tgt, sd, p + ".ln1")
    qk = h if query_pos is None else h + query_pos
    if conv_self:
        a, m = conv_attention(qk, qk, tgt, tgt_mask, tgt_mask, sd, p + ".self_attn", n_head, window=window)
    else:
        a, m = plain_attention(qk, qk, tgt, tgt_mask, tgt_mask, sd, p + ".self_attn", n_head), tgt_mask
    mf = m.to(tgt.dtype)
    tgt = tgt * mf + sd[p + ".drop_path_attn1.scale"] * a
    h = layer_norm(tgt, sd, p + ".ln2")
    q = h if query_pos is None else h + query_pos
    a, m = conv_attention(q, memory, memory, tgt_mask, mem_mask, sd, p + ".multihead_attn", n_head, window=window)
    mf = m.to(tgt.dtype)
    tgt = tgt * mf + sd[p + ".drop_path_attn2.scale"] * a
    if with_ffn:
        h = layer_norm(tgt, sd, p + ".ln3")
        h = conv1x1(F.gelu(conv1x1(h, sd, p + ".mlp.0")), sd, p + ".mlp.3") * mf
        tgt = tgt + sd[p + ".drop_path_mlp.scale"] * h
    return tgt


# ----------------------------------------------------------------------------------------------
# backbone / neck / predictor
# ----------------------------------------------------------------------------------------------
def backbone(x, mask, sd, cfg, taps=None):
    nv, nbe, nbs = cfg["visual_dim"], cfg["bbox_entity_dim"], cfg["bbox_so_dim"]
    clip = bool(cfg.get("with_clip_feature", False))
    nc = cfg["clip_dim"] if clip else 0
    n_embd_layers, n_stem, n_branch = cfg["backbone_arch"]
    n_head, n_fuse_head = cfg["n_head"], cfg["fuse_head"]
    win = cfg["n_mha_win_size"]
    assert x.shape[1] == 2 * nv + 2 * nc + nbs + 2 * nbe
    c0 = 2 * nv + 2 * nc
    streams = [x[:, :nv], x[:, nv:2 * nv]]
    clips = [x[:, 2 * nv:2 * nv + nc], x[:, 2 * nv + nc:c0]] if clip else None
    bbox_so = x[:, c0:c0 + nbs]
    bbox_ent = [x[:, c0 + nbs:c0 + nbs + nbe], x[:, c0 + nbs + nbe:]]
    mf = mask.to(x.dtype)
    p = "backbone"

    def embed(f, conv_p, norm_p):
        for i in range(n_embd_layers):
            f, _ = masked_conv(f, mask, sd, f"{p}.{conv_p}.{i}")
            f = F.relu(layer_norm(f, sd, f"{p}.{norm_p}.{i}"))
        return f

    streams = [embed(f, "visual_embd", "visual_embd_norm") for f in streams]
    if clip:
        clips = [embed(f, "clip_embd", "clip_embd_norm") for f in clips]
        streams = [conv_mlp(torch.cat([f, c], 1), sd, p + ".visual_clip_fuse", 2) * mf
                   for f, c in zip(streams, clips)]
    boxes = []
    for f in bbox_ent:
        f, _ = masked_conv(f, mask, sd, p + ".bbox_entity_embd")
        boxes.append(F.relu(layer_norm(f, sd, p + ".bbox_entity_norm")))
    s, o = [conv_mlp(torch.cat([f, b], 1), sd, p + ".visual_bbox_fuse", 2) * mf for f, b in zip(streams, boxes)]
    if taps is not None:
        taps["s_in"], taps["o_in"] = s, o

    sos_window = win if cfg["use_local"] else None
    for i in range(n_stem):
        s, _ = transformer_block(s, mask, sd, f"{p}.stem.{i}", n_head, win, 1)
        o, _ = transformer_block(o, mask, sd, f"{p}.stem.{i}", n_head, win, 1)
        if taps is not None:
            taps[f"s_stem{i}"], taps[f"o_stem{i}"] = s, o
        s_mut = decoder_layer(s, o, mask, mask, sd, f"{p}.s_attn.{i}", n_fuse_head, window=sos_window)
        o_mut = decoder_layer(o, s, mask, mask, sd, f"{p}.o_attn.{i}", n_fuse_head, window=sos_window)
        s, o = s + s_mut, o + o_mut
        if taps is not None:
            taps[f"s_sos{i}"], taps[f"o_sos{i}"] = s, o

    s = layer_norm(s, sd, p + ".s_fuse_norm")
    o = layer_norm(o, sd, p + ".o_fuse_norm")
    so = conv_mlp(torch.cat([s, o], 1), sd, p + ".so_fuse", 2) * mf
    b, _ = masked_conv(bbox_so, mask, sd, p + ".bbox_so_embd")
    e = conv_mlp(torch.cat([so, b], 1), sd, p + ".so_visual_bbox_fuse", 2) * mf
    feats, masks = [e], [mask]
    for i in range(n_branch):
        e, mask = transformer_block(e, mask, sd, f"{p}.branch.{i}", n_head, win, cfg["scale_factor"])
        feats.append(e)
        masks.append(mask)
    if taps is not None:
        for i, f in enumerate(feats):
            taps[f"e{i}"] = f
    return feats, masks


def neck(feats, masks, sd, taps=None):
    n = len(feats)
    y = None
    for l in range(n - 1, -1, -1):
        x = layer_norm(feats[l], sd, f"neck.input_norms.{l}")
        if l == n - 1:
            w = sd[f"neck.fpn_convs.{l}.conv.weight"]
            y, _ = masked_conv(x, masks[l], sd, f"neck.fpn_convs.{l}", groups=w.shape[0])
        else:
            cur, _ = masked_conv(x, masks[l], sd, f"neck.lateral_convs.{l}")
            cur = layer_norm(cur, sd, f"neck.lateral_norms.{l}")
            y = cur + y.repeat_interleave(2, dim=2)  # nearest x2 up-sampling
            y, _ = masked_conv(y, masks[l], sd, f"neck.fpn_convs.{l}", groups=y.shape[1])
        y = layer_norm(y, sd, f"neck.fpn_norms.{l}")
        if taps is not None:
            taps[f"fpn{l}"] = y
    out, _ = masked_conv(y, masks[0], sd, "neck.mask_features", groups=y.shape[1])
    return out


def predictor(x, mask_features, mask, output_mask, sd, cfg, taps=None):
    pc = cfg["predictor"]
    n_layers, n_head = pc["num_layers"], pc["n_head"]
    src = conv1x1(layer_norm(x, sd, "predictor.input_norm"), sd, "predictor.input_proj") * mask.to(x.dtype)
    b = src.shape[0]
    qpos = sd["predictor.query_embed.weight"].t()[None].expand(b, -1, -1)  # (B, 256, Q)
    tgt = torch.zeros_like(qpos)
    tmask = torch.ones(b, 1, qpos.shape[-1], dtype=torch.bool)
    for i in range(n_layers):
        tgt = decoder_layer(tgt, src, tmask, mask, sd, f"predictor.transformer.decoder.layers.{i}", n_head,
                            query_pos=qpos, conv_self=False, with_ffn=True)
        if taps is not None:
            taps[f"dec{i}"] = tgt
    hs = layer_norm(tgt, sd, "predictor.transformer.decoder.norm")
    logits = conv1x1(hs, sd, "predictor.class_embed").transpose(1, 2)            # (B, Q, K+1)
    membed = conv_mlp(hs, sd, "predictor.mask_embed", 3).transpose(1, 2)         # (B, Q, 256)
    pm = torch.einsum("bqc,bcm->bqm", membed, mask_features)
    pm = pm.masked_fill(~output_mask, -10.0)
    if taps is not None:
        taps["hs"], taps["mask_embed"] = hs, membed
    return {"pred_logits": logits, "pred_masks": pm, "output_mask": output_mask}


def mask_vrd(x, mask, sd, cfg, taps=None):
    """x: (B, C, T) float, mask: (B, 1, T) bool -> dict(pred_logits (B,Q,K+1), pred_masks (B,Q,T), output_mask)."""
    feats, masks = backbone(x, mask, sd, cfg, taps)
    mf = neck(feats, masks, sd, taps)
    if taps is not None:
        taps["mask_features"] = mf
    return predictor(feats[-1], mf, masks[-1], masks[0], sd, cfg, taps)


# ----------------------------------------------------------------------------------------------
# batching + post-processing
# ----------------------------------------------------------------------------------------------
def max_div_factor(cfg) -> int:
    """Largest (fpn stride x 2*(win//2)) over the pyramid levels (maskvrd.py:57-63)."""
    n_levels = cfg["backbone_arch"][-1] + 1
    w = cfg["n_mha_win_size"]
    best = 1
    for l in range(cfg["fpn_start_level"], n_levels):
        s = cfg["scale_factor"] ** l
        best = max(best, s * (w // 2) * 2 if w > 1 else s)
    return best


def padded_lengths(lengths: List[int], cfg) -> List[int]:
    """Padded length the reference gives every pair of a list of pair lengths: pairs are taken in
    slices of ``max_so_pair``; short pairs (L <= max_seq_len) pad to max_seq_len, long pairs to the
    slice's longest length rounded up to ``max_div_factor``."""
    msl, mdf, chunk = cfg["max_seq_len"], max_div_factor(cfg), cfg["max_so_pair"]
    out = []
    for s in range(0, len(lengths), chunk):
        sl = lengths[s:s + chunk]
        longest = max([msl] + [l for l in sl if l > msl])
        t_long = (longest + mdf - 1) // mdf * mdf
        out += [msl if l <= msl else t_long for l in sl]
    return out


def pad_batch(feats: List[torch.Tensor], t_pad: int):
    x = feats[0].new_zeros(len(feats), feats[0].shape[0], t_pad)
    lens = torch.tensor([f.shape[1] for f in feats])
    for i, f in enumerate(feats):
        x[i, :, :f.shape[1]] = f
    mask = (torch.arange(t_pad)[None, :] < lens[:, None])[:, None, :]
    return x, mask


def network_outputs(feats_list: List[torch.Tensor], sd, cfg, batch: int = 16):
    """Per-pair (logits (Q,K+1), masks (Q,L)) with the reference's padding semantics, in input order."""
    lens = [f.shape[1] for f in feats_list]
    tpads = padded_lengths(lens, cfg)
    logits: List[Optional[torch.Tensor]] = [None] * len(lens)
    masks: List[Optional[torch.Tensor]] = [None] * len(lens)
    groups: Dict[int, List[int]] = {}
    for i, t in enumerate(tpads):
        groups.setdefault(t, []).append(i)
    for t, ids in groups.items():
        for s in range(0, len(ids), batch):
            sub = ids[s:s + batch]
            x, m = pad_batch([feats_list[i] for i in sub], t)
            out = mask_vrd(x, m, sd, cfg)
            for j, i in enumerate(sub):
                logits[i] = out["pred_logits"][j]
                masks[i] = out["pred_masks"][j][:, :lens[i]]
    return logits, masks


def decode_triplets(logits, masks, input_data, infer_cfg):
    """softmax -> top-k classes (1..K) per query; per (pair, query, k): binarise the mask with
    sigmoid > 0.5, first/last active frame -> duration; drop short ones; rank by mean score."""
    topk, n_max = infer_cfg["topk"], infer_cfg["n_max_pair"]
    stride, min_frames = infer_cfg["feat_stride"], infer_cfg["pred_min_frames"]
    rows = []
    probs = F.softmax(torch.stack(logits, 0), dim=-1)
    scores, cats = torch.topk(probs[..., 1:], k=topk, dim=-1)
    cats = cats + 1
    for i, (sid, oid) in enumerate(zip(input_data["sids"].tolist(), input_data["oids"].tolist())):
        s_dur, o_dur = input_data["traj_durations"][sid].tolist(), input_data["traj_durations"][oid].tolist()
        so_start, so_end = max(s_dur[0], o_dur[0]), min(s_dur[1], o_dur[1])
        off = int(input_data["so_offset"][i])
        active = torch.sigmoid(masks[i]) > 0.5                   # (Q, L)
        for q in range(active.shape[0]):
            nz = torch.nonzero(active[q]).flatten()
            if nz.numel() == 0:
                continue
            a, b = int(nz[0]) * stride + off, int(nz[-1]) * stride + off + 1
            assert a >= 0 and b <= so_end - so_start
            if b - a < min_frames:
                continue
            for k in range(topk):
                rows.append((i, sid, oid, q, k, so_start + a, so_start + b, a, b))
    if not rows:
        return None
    trip = torch.tensor([[int(input_data["cat_ids"][r[1]]), int(cats[r[0], r[3], r[4]]), int(input_data["cat_ids"][r[2]])]
                         for r in rows])
    sc = torch.stack([torch.stack([input_data["cat_scores"][r[1]], scores[r[0], r[3], r[4]],
                                   input_data["cat_scores"][r[2]]]) for r in rows])
    avg = sc.mean(-1)
    order = torch.argsort(avg, descending=True)[:n_max].tolist()
    out = {"triplets": [], "triple_scores": [], "triple_scores_avg": [], "so_trajs": [], "pred_durations": [],
           "so_tids": []}
    for j in order:
        i, sid, oid, q, k, d0, d1, a, b = rows[j]
        s_dur, o_dur = input_data["traj_durations"][sid].tolist(), input_data["traj_durations"][oid].tolist()
        so_start = max(s_dur[0], o_dur[0])
        st = input_data["bboxes_list"][sid][so_start - s_dur[0] + a: so_start - s_dur[0] + b]
        ot = input_data["bboxes_list"][oid][so_start - o_dur[0] + a: so_start - o_dur[0] + b]
        out["triplets"].append(trip[j].tolist())
        out["triple_scores"].append(sc[j].tolist())
        out["triple_scores_avg"].append(float(avg[j]))
        out["so_trajs"].append([st.tolist(), ot.tolist()])
        out["pred_durations"].append([d0, d1])
        out["so_tids"].append([sid, oid])
    return out


def forward_test(input_data, sd, cfg, infer_cfg):
    """Restatement of MaskVRD.forward_test: dict of per-pair tensors in -> dict of Python lists out (or None)."""
    with torch.no_grad():
        logits, masks = network_outputs(input_data["so_features_list"], sd, cfg)
        return decode_triplets(logits, masks, input_data, infer_cfg)
