"""Generates the golden fixtures under tests/golden/ by running the UNMODIFIED reference (imported read-only from
/root/reference) on seeded inputs.  Run in the build container only:  python tests/golden/make_golden.py

Weights and inputs are not stored: they are re-derived from the seeds with vrdone_b200.synth (CPU RNG, deterministic);
a checksum of both is stored so that RNG drift is detected instead of silently invalidating the fixtures.
"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from vrdone_b200 import MaskVRD, synth  # noqa: E402
from vrdone_b200.layout import reference_padded_lengths  # noqa: E402

CASES = {
    "vidvrd": dict(lens=[96, 95, 49, 7, 2, 89, 97, 109], wseed=11, xseed=12),
    "vidor": dict(lens=[512, 511, 505, 121, 127, 128, 64, 2, 3, 513, 640], wseed=21, xseed=22),
    "vidor_local": dict(lens=[512, 509, 33, 200, 576], wseed=31, xseed=32),
    "vidor_x": dict(lens=[512, 510, 17, 3, 130], wseed=41, xseed=42),
}
VIDEO_CASES = {"vidvrd": dict(wseed=11, vseed=0), "vidor": dict(wseed=21, vseed=3, n_tracklets=5, n_frames=700),
               # round 2: the SOS-windowed and CLIP configs, and a VidOR video with more than max_so_pair pairs whose 200-pair
               # slices each hold long pairs (L > max_seq_len: the reference's long-batch padding, maskvrd.py:364-379)
               "vidor_local": dict(wseed=31, vseed=6, n_tracklets=5, n_frames=700),
               "vidor_x": dict(wseed=41, vseed=7, n_tracklets=4, n_frames=600),
               "vidor_long": dict(config="vidor", wseed=21, vseed=10, n_tracklets=16, n_frames=2600)}
# default (timing) initialisation: torch.manual_seed(0) + the module's own init, as bench.py uses it
DEFAULT_INIT_CASES = {"vidor_default": dict(config="vidor", lens=[37, 128, 300, 512, 600], xseed=52)}


def load_reference():
    sys.path.insert(0, "/root/reference")
    for m in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
        del sys.modules[m]
    from models.maskvrd import MaskVRD as Ref
    sys.path.pop(0)
    return Ref


def checksum(tensors):
    return float(sum(t.double().abs().sum() for t in tensors))


def build_ref(Ref, name, wseed):
    cfg = synth.load_config(name)
    mc = cfg["model_config"]
    ref = Ref(mc, "cpu").eval()
    ref._config_eval(cfg["inference_config"])
    ours = MaskVRD(mc, "cpu")
    assert set(ours.state_dict().keys()) == set(ref.state_dict().keys())
    sd = synth.stress_state_dict(ours.state_dict(), wseed)
    ref.load_state_dict(sd, strict=True)
    return cfg, ref, sd


LOADER_CASES = {
    "vidor": [dict(seed=0, n_tracklets=10, n_frames=600, n_dup=8), dict(seed=1, n_tracklets=14, n_frames=900, n_dup=10),
              dict(seed=2, n_tracklets=7, n_frames=400, n_dup=0)],
    "vidor_x": [dict(seed=3, n_tracklets=8, n_frames=500, n_dup=6)],
    "vidvrd": [dict(seed=4, n_tracklets=8, n_frames=150, n_dup=6)],
}


def loader_fixtures():
    """the unmodified ``VidOR._val_getitem`` (dataloaders/vidor.py:556-734; vidvrd.py:552-716 is the same
    code) on seeded synthetic tracklet videos with injected near-duplicates.  The method only reads four attributes of its
    dataset object, so it is called on a stand-in that carries them (constructing the dataset needs the annotation files)."""
    import types
    sys.path.insert(0, "/root/reference")
    for m in [k for k in sys.modules if k == "utils" or k.startswith("utils.") or k.startswith("dataloaders")]:
        del sys.modules[m]
    import dataloaders.vidor as V
    import dataloaders.vidvrd as VV
    sys.path.pop(0)
    for name, cases in LOADER_CASES.items():
        cfg = synth.load_config(name)
        dc = cfg["dataset_config"]
        fixes = []
        for case in cases:
            trk = synth.synthetic_tracklet_video(cfg, case["seed"], n_tracklets=case["n_tracklets"], n_frames=case["n_frames"])
            if case["n_dup"]:
                trk = synth.with_duplicates(trk, cfg, case["seed"], case["n_dup"])
            stand_in = types.SimpleNamespace(with_clip_feature="clip_features_list" in trk, feat_stride=dc.get("feat_stride", 1),
                                             random_stride=False, stride_offset=0,
                                             proposal_min_frames=dc.get("proposal_min_frames", 0))
            if name == "vidvrd":      # the ImageNet-VidVRD loader's own copy of the method (dataloaders/vidvrd.py:552-716)
                out = VV.VidVRD._test_getitem(stand_in, trk, viou_threshold=0.9)
            else:
                out = V.VidOR._val_getitem(stand_in, trk, viou_threshold=0.9)
            fixes.append({**case, "n_candidate_pairs": len(trk["sids"]), "n_tracklets_total": len(trk["bboxes_list"]),
                          "inputs_checksum": checksum(trk["visual_features_list"] + trk["bboxes_list"]),
                          "proposal_min_frames": stand_in.proposal_min_frames, "feat_stride": stand_in.feat_stride,
                          "sids": out["sids"].tolist(), "oids": out["oids"].tolist(), "so_offset": out["so_offset"].tolist(),
                          "lens": [int(f.shape[1]) for f in out["so_features_list"]],
                          "pair_checksums": [float(f.double().abs().sum()) for f in out["so_features_list"]],
                          "boxes_checksum": checksum(out["bboxes_list"])})
            print(name, case, "loader fixture:", len(trk["sids"]), "candidate pairs ->", len(out["sids"]), "kept")
        with open(os.path.join(HERE, f"loader_{name}.json"), "w") as f:
            json.dump({"config": name, "viou_threshold": 0.9, "cases": fixes}, f)


def main():
    if "--loader-only" in sys.argv:
        return loader_fixtures()
    # --only name[,name...]: regenerate just these network / video fixtures (existing files of the others stay untouched)
    only = set(sys.argv[sys.argv.index("--only") + 1].split(",")) if "--only" in sys.argv else None
    if not only:
        loader_fixtures()
    Ref = load_reference()
    torch.manual_seed(0)
    for name, case in CASES.items():
        if only and name not in only:
            continue
        cfg, ref, sd = build_ref(Ref, name, case["wseed"])
        mc = cfg["model_config"]
        feats = synth.pair_features(mc, case["lens"], case["xseed"])
        tpads = reference_padded_lengths(case["lens"], mc)
        logits, masks = [], []
        with torch.no_grad():
            for f, t in zip(feats, tpads):   # one pair per reference call: outputs depend only on (pair, T_pad)
                x = torch.zeros(1, f.shape[0], t)
                x[0, :, : f.shape[1]] = f
                m = (torch.arange(t) < f.shape[1])[None, None]
                r = ref._mask_vrd(x, m)
                logits.append(r["pred_logits"][0].clone())
                masks.append(r["pred_masks"][0][:, : f.shape[1]].clone())
        fix = {"config": name, "lens": case["lens"], "tpads": tpads, "wseed": case["wseed"], "xseed": case["xseed"],
               "weights_checksum": checksum(sd.values()), "inputs_checksum": checksum(feats),
               "pred_logits": torch.stack(logits), "pred_masks": masks,
               "schema": {k: list(v.shape) for k, v in sd.items()}}
        torch.save(fix, os.path.join(HERE, f"network_{name}.pt"))
        print(name, "network fixture:", fix["pred_logits"].shape, "logit std", float(fix["pred_logits"].std()))
    for name, case in DEFAULT_INIT_CASES.items():
        if only and name not in only:
            continue
        cfg = synth.load_config(case["config"])
        mc = cfg["model_config"]
        torch.manual_seed(0)
        ours = MaskVRD(mc, "cpu")
        sd = {k: v.detach().clone() for k, v in ours.state_dict().items()}
        ref = Ref(mc, "cpu").eval()
        ref.load_state_dict(sd, strict=True)
        feats = synth.pair_features(mc, case["lens"], case["xseed"])
        tpads = reference_padded_lengths(case["lens"], mc)
        logits, masks = [], []
        with torch.no_grad():
            for f, t in zip(feats, tpads):
                x = torch.zeros(1, f.shape[0], t)
                x[0, :, : f.shape[1]] = f
                m = (torch.arange(t) < f.shape[1])[None, None]
                r = ref._mask_vrd(x, m)
                logits.append(r["pred_logits"][0].clone())
                masks.append(r["pred_masks"][0][:, : f.shape[1]].clone())
        fix = {"config": case["config"], "lens": case["lens"], "tpads": tpads, "xseed": case["xseed"], "init": "torch.manual_seed(0)",
               "weights_checksum": checksum(sd.values()), "inputs_checksum": checksum(feats),
               "pred_logits": torch.stack(logits), "pred_masks": masks}
        torch.save(fix, os.path.join(HERE, f"network_{name}.pt"))
        print(name, "default-init network fixture: logit std", float(fix["pred_logits"].std()))
    for name, case in VIDEO_CASES.items():
        if only and name not in only:
            continue
        cfg, ref, sd = build_ref(Ref, case.get("config", name), case["wseed"])
        kw = {k: v for k, v in case.items() if k in ("n_tracklets", "n_frames")}
        video = synth.synthetic_video(cfg, case["vseed"], **kw)
        with torch.no_grad():
            out = ref(video)
        lens = [int(f.shape[1]) for f in video["so_features_list"]]
        if out is not None:   # keep the fixture small: trajectories are stored as (n_frames, checksum) per triplet
            out["so_trajs"] = [[len(t[0]), float(torch.tensor(t).double().sum())] for t in out["so_trajs"]]
        fix = {"config": case.get("config", name), **case, "n_pairs": len(lens), "lens": lens, "inputs_checksum": checksum(video["so_features_list"]),
               "weights_checksum": checksum(sd.values()), "output": out}
        with open(os.path.join(HERE, f"video_{name}.json"), "w") as f:
            json.dump(fix, f)
        print(name, "video fixture:", len(lens), "pairs ->", None if out is None else len(out["triplets"]), "triplets")


if __name__ == "__main__":
    main()
