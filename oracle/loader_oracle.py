"""CPU oracle for the data loader's per-video pair construction  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates ``VidOR._val_getitem`` of the reference (dataloaders/vidor.py:556-734; the VidVRD loader, dataloaders/vidvrd.py:552-716,
is the same algorithm): box clamping, the duplicate-tracklet vIoU filter, the pair loop with its two drop rules, the sub-sampled
feature gather and the box-geometry features.  It is the checker of SURVEY.md section 8f rows 1 and 2 (tracklet-level entry point and
the device-side vIoU filter); only ``tests/`` may import it.

Parity status: PINNED against outputs of the reference itself.  ``tests/golden/make_golden.py`` imports
``/root/reference/dataloaders/vidor.py`` read-only, calls the unmodified ``VidOR._val_getitem`` on seeded synthetic tracklet
videos (with injected near-duplicate tracklets so that both removal rules fire) and stores the surviving pairs, their lengths
and a checksum of every pair tensor in ``tests/golden/loader_*.json``; ``tests/test_oracle.py`` checks this file against them and,
when ``/root/reference`` exists, against the live reference.

Reference lines each function follows (paths relative to /root/reference):
  clamp_boxes           dataloaders/vidor.py:572-580
  viou_sums             dataloaders/vidor.py:609-631
  duplicate_filter      dataloaders/vidor.py:583-641
  surviving_pairs       dataloaders/vidor.py:643-650
  val_getitem           dataloaders/vidor.py:556-734
  geometry              utils/misc.py:158-217
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch

TO_REMOVE = 1      # dataloaders/vidor.py:16


def clamp_boxes(bboxes_list: List[torch.Tensor], video_wh) -> List[torch.Tensor]:
    w, h = video_wh
    out = []
    for b in bboxes_list:
        b = b.clone()
        b[:, 0] = torch.clamp(b[:, 0], 0)
        b[:, 1] = torch.clamp(b[:, 1], 0)
        b[:, 2] = torch.clamp(b[:, 2], None, w - 1)
        b[:, 3] = torch.clamp(b[:, 3], None, h - 1)
        out.append(b)
    return out


def viou_sums(b_bbox: torch.Tensor, r_bbox: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(intersection volume, base volume, ref volume) of two box sequences over the same frames."""
    area_b = (b_bbox[:, 2] - b_bbox[:, 0] + TO_REMOVE) * (b_bbox[:, 3] - b_bbox[:, 1] + TO_REMOVE)
    area_r = (r_bbox[:, 2] - r_bbox[:, 0] + TO_REMOVE) * (r_bbox[:, 3] - r_bbox[:, 1] + TO_REMOVE)
    lt = torch.max(b_bbox[:, :2], r_bbox[:, :2])
    rb = torch.min(b_bbox[:, 2:], r_bbox[:, 2:])
    wh = (rb - lt + TO_REMOVE).clamp(min=0.0)
    return (wh[:, 0] * wh[:, 1]).sum(), area_b.sum(), area_r.sum()


def duplicate_filter(bboxes_list, traj_durations, cat_ids, viou_threshold: float = 0.9) -> List[bool]:
    """Greedy removal of duplicate tracklets.  For base < ref of the same category with overlapping durations: ref is dropped
    when inter / vol(ref) > threshold and base covers ref in time; otherwise base is dropped (and its scan ends) when
    inter / vol(base) > threshold and ref covers base.  Refs already dropped are skipped; a dropped base still scans."""
    n = len(bboxes_list)
    valid = [True] * n
    durs = [[int(traj_durations[i][0]), int(traj_durations[i][1])] for i in range(n)]
    for base in range(n):
        for ref in range(base + 1, n):
            if not valid[ref] or int(cat_ids[base]) != int(cat_ids[ref]):
                continue
            bd, rd = durs[base], durs[ref]
            if rd[0] >= bd[1] or rd[1] <= bd[0]:
                continue
            s, e = max(bd[0], rd[0]), min(bd[1], rd[1])
            inter, vol_b, vol_r = viou_sums(bboxes_list[base][s - bd[0]: e - bd[0]], bboxes_list[ref][s - rd[0]: e - rd[0]])
            if inter / vol_r > viou_threshold and bd[0] <= rd[0] and bd[1] >= rd[1]:
                valid[ref] = False
            elif inter / vol_b > viou_threshold and rd[0] <= bd[0] and rd[1] >= bd[1]:
                valid[base] = False
                break
    return valid


def surviving_pairs(sids: torch.Tensor, oids: torch.Tensor, valid: List[bool]) -> torch.Tensor:
    """Mask over the candidate pairs whose subject and object both survive."""
    v = torch.tensor(valid, dtype=torch.bool)
    return v[sids] & v[oids]


def geometry(sb: torch.Tensor, ob: torch.Tensor, w: float, h: float):
    """5-d relative and 8-d per-entity box features (utils/misc.py:158-217)."""
    def cwh(b):
        return (b[:, 0] + b[:, 2]) / 2, (b[:, 1] + b[:, 3]) / 2, b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]

    sx, sy, sw, sh = cwh(sb)
    ox, oy, ow, oh = cwh(ob)
    rel = torch.stack([(sx - ox) / ox, (sy - oy) / oy, torch.log(sw / ow), torch.log(sh / oh), torch.log((sw * sh) / (ow * oh))], 1)

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


def val_getitem(data: Dict, feat_stride: int, stride_offset: int = 0, proposal_min_frames: int = 0,
                viou_threshold: float = 0.9, with_clip: bool = False) -> Dict:
    """The eval-time item the reference loader hands to ``MaskVRD.forward`` (``{}`` when nothing survives)."""
    w, h = data["video_wh"]
    boxes = clamp_boxes(data["bboxes_list"], (w, h))
    durs = data["traj_durations"]
    valid = duplicate_filter(boxes, durs, data["cat_ids"], viou_threshold)
    keep = surviving_pairs(data["sids"], data["oids"], valid)
    sids, oids = data["sids"][keep], data["oids"][keep]
    if len(sids) == 0:
        return {}
    vis = data["visual_features_list"]
    clips = data.get("clip_features_list") if with_clip else None
    feats, offs, ok = [], [], []
    for s, o in zip(sids.tolist(), oids.tolist()):
        a, b = max(int(durs[s][0]), int(durs[o][0])), min(int(durs[s][1]), int(durs[o][1]))
        n, sd, od = b - a, a - int(durs[s][0]), a - int(durs[o][0])
        s_feat = vis[s][sd: n + sd]
        if s_feat.shape[0] < proposal_min_frames:
            ok.append(False)
            continue
        s_feat = s_feat[stride_offset::feat_stride]
        if s_feat.shape[0] < 2:
            ok.append(False)
            continue
        sel = slice(stride_offset, None, feat_stride)
        o_feat = vis[o][od: n + od][sel]
        sb, ob = boxes[s][sd: n + sd][sel], boxes[o][od: n + od][sel]
        rel, es, eo = geometry(sb, ob, w, h)
        parts = [s_feat, o_feat]
        if clips is not None:
            parts += [clips[s][sd: n + sd][sel], clips[o][od: n + od][sel]]
        feats.append(torch.cat(parts + [rel, es, eo], -1).permute(1, 0))
        offs.append(stride_offset)
        ok.append(True)
    okt = torch.tensor(ok, dtype=torch.bool)
    sids, oids = sids[okt], oids[okt]
    if len(sids) == 0:
        return {}
    return {"sids": sids, "oids": oids, "cat_ids": data["cat_ids"], "cat_scores": data["cat_scores"],
            "traj_durations": durs, "bboxes_list": boxes, "so_features_list": feats,
            "so_offset": torch.tensor(offs, dtype=torch.int64), "valid_tracklets": valid}
