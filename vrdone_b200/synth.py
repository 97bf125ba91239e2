"""Seeded synthetic inputs and weight initialisations.

Everything here is deterministic given a seed and runs on the CPU RNG, so the build container, the
test fixtures and the GPU box produce identical tensors.
"""
from __future__ import annotations

import os
from typing import Dict, List

import torch
import yaml

CONFIG_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "configs")
CONFIG_NAMES = ("vidvrd", "vidor", "vidor_local", "vidor_x")


def load_config(name: str) -> dict:
    """Return {'model_config', 'inference_config', 'dataset_config'}; ``with_clip_feature`` is copied from the
    dataset section into the model section, as the reference CLI does (eval.py:53)."""
    with open(os.path.join(CONFIG_DIR, name + ".yaml")) as f:
        cfg = yaml.safe_load(f)
    cfg["model_config"]["with_clip_feature"] = bool(cfg["dataset_config"].get("with_clip_feature", False))
    return cfg


def input_channels(mc: dict) -> int:
    nc = mc["clip_dim"] if mc.get("with_clip_feature", False) else 0
    return 2 * mc["visual_dim"] + 2 * nc + mc["bbox_so_dim"] + 2 * mc["bbox_entity_dim"]


def stress_state_dict(sd: Dict[str, torch.Tensor], seed: int) -> Dict[str, torch.Tensor]:
    """A seeded initialisation under which every branch of the network matters (the default init has
    1e-4 path scales and constant class logits, SURVEY.md section 7 hard-part 1).  Keys are visited in
    sorted order so the result depends only on (names, shapes, seed)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k in sorted(sd.keys()):
        v = sd[k]
        shape = tuple(v.shape)
        if k == "empty_weight":
            out[k] = v.clone()
        elif k.endswith(".scale"):
            out[k] = 0.3 + 0.7 * torch.rand(shape, generator=g)
        elif "norm" in k or ".ln" in k:
            if k.endswith(".weight"):
                out[k] = 0.5 + torch.rand(shape, generator=g)
            else:
                out[k] = 0.5 * torch.randn(shape, generator=g)
        elif k.endswith("query_embed.weight"):
            out[k] = torch.randn(shape, generator=g)
        elif k.endswith(".bias"):
            out[k] = 0.1 * torch.randn(shape, generator=g)
        else:  # conv weights: uniform with the fan-in bound of the default Conv1d init
            fan_in = shape[1] * shape[2]
            bound = 1.0 / fan_in ** 0.5
            out[k] = (2 * torch.rand(shape, generator=g) - 1) * bound
    return out


def pair_features(mc: dict, lengths: List[int], seed: int) -> List[torch.Tensor]:
    """One (C, L) fp32 tensor per pair, N(0,1) visual/CLIP channels and O(1) geometry channels, delivered the way
    the reference data loader does: a transposed view of an (L, C)-contiguous buffer (vidor.py:708-711)."""
    g = torch.Generator().manual_seed(seed)
    c = input_channels(mc)
    return [torch.randn(l, c, generator=g).permute(1, 0) for l in lengths]


def _geometry(sb, ob, w, h):
    """5-d subject/object relative and 8-d per-entity box features (formulas of utils/misc.py:158-217)."""
    def cwh(b):
        return (b[:, 0] + b[:, 2]) / 2, (b[:, 1] + b[:, 3]) / 2, b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]

    sx, sy, sw, sh = cwh(sb)
    ox, oy, ow, oh = cwh(ob)
    rel = torch.stack([(sx - ox) / ox, (sy - oy) / oy, torch.log(sw / ow), torch.log(sh / oh),
                       torch.log((sw * sh) / (ow * oh))], 1)

    def ent(b):
        n = b.clone()
        n[:, 0::2] /= w
        n[:, 1::2] /= h
        cols = []
        for v in cwh(n):
            d = v[1:] - v[:-1]
            first = d[:1] - (d[1:2] - d[:1]) if d.numel() > 1 else d[:1]
            cols += [v, torch.cat([first, d])]
        return torch.stack(cols, 1)

    return rel, ent(sb), ent(ob)


def cfg2_video_set(n_videos: int = 10, seed: int = 0):
    """SURVEY.md section 8d cfg2 as a fixed, stratified video set: frame counts F in {900, 1200, 1800, 3600} in the proportions
    .3 / .4 / .2 / .1 (exactly, for multiples of 10 videos -- a plain draw of 10 misses the 3600-frame videos, the only ones with
    pairs longer than max_seq_len, 35 % of the time), tracklet counts N ~ U[20, 60].  Returns [(video seed, n_frames, n_tracklets)],
    to be handed to ``synthetic_video`` / ``synthetic_tracklet_video``."""
    g = torch.Generator().manual_seed(10007 * (seed + 1))
    quota = [(900, .3), (1200, .4), (1800, .2), (3600, .1)]
    frames: List[int] = []
    for f, p in quota:
        frames += [f] * int(round(p * n_videos))
    while len(frames) < n_videos:
        frames.append(1200)
    frames = frames[:n_videos]
    order = torch.randperm(n_videos, generator=g).tolist()
    return [(1000 * (seed + 1) + i, frames[order[i]], int(torch.randint(20, 61, (1,), generator=g))) for i in range(n_videos)]


def _tracklets(cfg: dict, seed: int, n_tracklets: int = None, n_frames: int = None, features: bool = True, split_rng: bool = False):
    """Seeded per-tracklet data (durations, boxes, visual / CLIP features); the RNG is returned so that callers draw the
    remaining per-video quantities in the historical order (the golden fixtures depend on it).  ``split_rng``: the features come
    from a second generator, so that ``features=False`` (structure only: durations and boxes, for cost models) yields the same
    durations as the full draw."""
    mc, dc = cfg["model_config"], cfg["dataset_config"]
    g = torch.Generator().manual_seed(seed)
    gf = torch.Generator().manual_seed(seed + 0x5EED) if split_rng else g
    assert features or split_rng, "structure-only draws need split_rng (the features share the structure's generator otherwise)"
    stride = dc.get("feat_stride", 1)
    clip = mc.get("with_clip_feature", False)
    vid_w, vid_h = 1280.0, 720.0

    def randint(lo, hi):
        return int(torch.randint(lo, hi, (1,), generator=g))

    if n_frames is None:
        if stride == 1:
            n_frames = 150
        else:
            n_frames = [900, 1200, 1800, 3600][int(torch.multinomial(torch.tensor([.3, .4, .2, .1]), 1, generator=g))]
    if n_tracklets is None:
        n_tracklets = 6 if stride == 1 else randint(20, 61)
    durs, boxes, vis, clips = [], [], [], []
    for _ in range(n_tracklets):
        if stride == 1:
            start = randint(0, 60)
            length = min(randint(30, 150), n_frames - start)
        else:
            length = randint(60, n_frames + 1)
            start = randint(0, n_frames - length + 1)
        durs.append([start, start + length])
        wh = 20 + 200 * torch.rand(length, 2, generator=g)
        xy = torch.rand(length, 2, generator=g) * (torch.tensor([vid_w, vid_h]) - wh - 2) + 1
        boxes.append(torch.cat([xy, xy + wh], 1))
        if features:
            vis.append(torch.randn(length, mc["visual_dim"], generator=gf))
            if clip:
                clips.append(torch.randn(length, mc["clip_dim"], generator=gf))
    return g, n_tracklets, durs, boxes, vis, clips, (vid_w, vid_h)


def pair_lengths(durs, stride):
    """Sub-sampled lengths of the pairs ``_overlapping_pairs`` keeps, in the same order."""
    return [len(range(0, min(durs[s][1], durs[o][1]) - max(durs[s][0], durs[o][0]), stride)) for s, o in _overlapping_pairs(durs, stride)]


def _overlapping_pairs(durs, stride):
    """Ordered (subject, object) tracklet pairs the synthetic loader keeps: temporal overlap of at least ``min_frames`` frames
    and at least two sub-sampled frames."""
    min_frames = 5 if stride > 1 else 2
    out = []
    for s in range(len(durs)):
        for o in range(len(durs)):
            if s == o:
                continue
            a, b = max(durs[s][0], durs[o][0]), min(durs[s][1], durs[o][1])
            if b - a < min_frames or len(range(0, b - a, stride)) < 2:
                continue
            out.append((s, o))
    return out


def synthetic_tracklet_video(cfg: dict, seed: int, n_tracklets: int = None, n_frames: int = None, name: str = None,
                             split_rng: bool = False) -> dict:
    """The same synthetic video as ``synthetic_video`` BEFORE the data loader's pair construction: the contract of the input of
    the reference's ``_val_getitem`` (dataloaders/vidor.py:556-571) with ``sids`` / ``oids`` already enumerated."""
    mc, dc = cfg["model_config"], cfg["dataset_config"]
    g, n_tracklets, durs, boxes, vis, clips, wh = _tracklets(cfg, seed, n_tracklets, n_frames, split_rng=split_rng)
    pairs = _overlapping_pairs(durs, dc.get("feat_stride", 1))
    n_cat = 35 if mc["num_classes"] > 100 else 80
    out = {
        "video_name": name or f"synthetic_{seed}",
        "video_wh": wh,
        "sids": torch.tensor([p[0] for p in pairs], dtype=torch.int64),
        "oids": torch.tensor([p[1] for p in pairs], dtype=torch.int64),
        "cat_ids": torch.randint(1, n_cat + 1, (n_tracklets,), generator=g),
        "cat_scores": 0.4 + 0.6 * torch.rand(n_tracklets, generator=g),
        "traj_durations": torch.tensor(durs, dtype=torch.int64),
        "bboxes_list": boxes,
        "visual_features_list": vis,
    }
    if mc.get("with_clip_feature", False):
        out["clip_features_list"] = clips
    return out


def synthetic_video(cfg: dict, seed: int, n_tracklets: int = None, n_frames: int = None, name: str = None, only_pairs=None) -> dict:
    """A synthetic ``input_data`` dict with the reference data loader's contract: every
    ordered pair of tracklets with enough temporal overlap, features sub-sampled with ``feat_stride``.  ``only_pairs``: keep
    just these pair indices (same tensors as the full video's, without building the others: bounded CPU samples)."""
    mc, dc, ic = cfg["model_config"], cfg["dataset_config"], cfg["inference_config"]
    stride = dc.get("feat_stride", 1)
    clip = mc.get("with_clip_feature", False)
    g, n_tracklets, durs, boxes, vis, clips, (vid_w, vid_h) = _tracklets(cfg, seed, n_tracklets, n_frames)
    n_cat = 35 if mc["num_classes"] > 100 else 80
    sids, oids, feats, offs = [], [], [], []
    only = None if only_pairs is None else set(int(i) for i in only_pairs)
    for pi, (s, o) in enumerate(_overlapping_pairs(durs, stride)):
        if only is not None and pi not in only:
            continue
        a, b = max(durs[s][0], durs[o][0]), min(durs[s][1], durs[o][1])
        ss, os_ = a - durs[s][0], a - durs[o][0]
        sl_s = slice(ss, ss + b - a, stride)
        sl_o = slice(os_, os_ + b - a, stride)
        rel, es, eo = _geometry(boxes[s][sl_s], boxes[o][sl_o], vid_w, vid_h)
        parts = [vis[s][sl_s], vis[o][sl_o]]
        if clip:
            parts += [clips[s][sl_s], clips[o][sl_o]]
        parts += [rel, es, eo]
        feats.append(torch.cat(parts, -1).permute(1, 0))
        sids.append(s)
        oids.append(o)
        offs.append(0)
    return {
        "video_name": name or f"synthetic_{seed}",
        "sids": torch.tensor(sids, dtype=torch.int64),
        "oids": torch.tensor(oids, dtype=torch.int64),
        "cat_ids": torch.randint(1, n_cat + 1, (n_tracklets,), generator=g),
        "cat_scores": 0.4 + 0.6 * torch.rand(n_tracklets, generator=g),
        "traj_durations": torch.tensor(durs, dtype=torch.int64),
        "bboxes_list": boxes,
        "so_features_list": feats,
        "so_offset": torch.tensor(offs, dtype=torch.int64),
    }


def with_duplicates(trk: dict, cfg: dict, seed: int, n_dup: int = 6) -> dict:
    """A tracklet-level video with ``n_dup`` near-duplicate tracklets inserted at random positions, so that the data loader's
    duplicate-tracklet vIoU filter (dataloaders/vidor.py:583-641) has work to do: copies of existing tracklets over a
    sub-interval (dropped by the filter when listed after their source: rule 1; their source is dropped when the longer
    copy comes later: rule 2), with box jitter small (vIoU ~0.97) or large (~0.8, survives), some with another category.
    ``sids`` / ``oids`` are re-enumerated over the new tracklet list."""
    g = torch.Generator().manual_seed(1000 + seed)
    mc, dc = cfg["model_config"], cfg["dataset_config"]
    clip = "clip_features_list" in trk
    durs = [list(map(int, d)) for d in trk["traj_durations"].tolist()]
    items = [dict(dur=durs[i], box=trk["bboxes_list"][i], vis=trk["visual_features_list"][i],
                  clip=trk["clip_features_list"][i] if clip else None, cat=int(trk["cat_ids"][i]), score=float(trk["cat_scores"][i]))
             for i in range(len(durs))]
    n0 = len(items)
    for d in range(n_dup):
        src = items[int(torch.randint(0, n0, (1,), generator=g))]
        a, b = src["dur"]
        length = b - a
        lo = int(torch.randint(0, max(1, length // 3), (1,), generator=g))
        hi = length - int(torch.randint(0, max(1, length // 3), (1,), generator=g))
        if d % 3 == 0:
            lo, hi = 0, length                                     # same duration: either rule may fire, depending on the order
        jitter = 1.5 if d % 4 != 3 else 14.0                        # pixels; the large one keeps vIoU below 0.9
        box = src["box"][lo:hi] + jitter * (torch.rand(hi - lo, 4, generator=g) - 0.5)
        item = dict(dur=[a + lo, a + hi], box=box, vis=torch.randn(hi - lo, mc["visual_dim"], generator=g),
                    clip=torch.randn(hi - lo, mc["clip_dim"], generator=g) if clip else None,
                    cat=src["cat"] if d % 5 != 4 else src["cat"] % 30 + 1, score=float(0.4 + 0.6 * torch.rand(1, generator=g)))
        items.insert(int(torch.randint(0, len(items) + 1, (1,), generator=g)), item)
    durs = [it["dur"] for it in items]
    pairs = _overlapping_pairs(durs, dc.get("feat_stride", 1))
    out = dict(trk)
    out.update({
        "sids": torch.tensor([p[0] for p in pairs], dtype=torch.int64),
        "oids": torch.tensor([p[1] for p in pairs], dtype=torch.int64),
        "cat_ids": torch.tensor([it["cat"] for it in items], dtype=torch.int64),
        "cat_scores": torch.tensor([it["score"] for it in items], dtype=torch.float32),
        "traj_durations": torch.tensor(durs, dtype=torch.int64),
        "bboxes_list": [it["box"] for it in items],
        "visual_features_list": [it["vis"] for it in items],
    })
    if clip:
        out["clip_features_list"] = [it["clip"] for it in items]
    return out
