"""Shared test helpers: golden fixtures, seeded models, reference loader (build container only)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"

from vrdone_b200 import MaskVRD, synth  # noqa: E402


def have_reference() -> bool:
    return os.path.isdir(os.path.join(REFERENCE, "models"))


def load_reference_class():
    """Import the reference MaskVRD from /root/reference in isolation (its top-level packages are called
    ``models`` / ``utils``)."""
    sys.path.insert(0, REFERENCE)
    try:
        for m in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            del sys.modules[m]
        from models.maskvrd import MaskVRD as Ref
    finally:
        sys.path.remove(REFERENCE)
    return Ref


def checksum(tensors):
    return float(sum(t.double().abs().sum() for t in tensors))


def network_fixture(name):
    return torch.load(os.path.join(GOLDEN, f"network_{name}.pt"), weights_only=False)


def video_fixture(name):
    with open(os.path.join(GOLDEN, f"video_{name}.json")) as f:
        return json.load(f)


def seeded_model(name, wseed, device="cpu", precision="fp32"):
    """Our MaskVRD with the seeded stress initialisation of the fixtures."""
    cfg = synth.load_config(name)
    model = MaskVRD(cfg["model_config"], device)
    sd = synth.stress_state_dict(model.state_dict(), wseed)
    model.load_state_dict(sd, strict=True)
    model.eval()
    model._config_eval(cfg["inference_config"])
    model.set_precision(precision)
    return cfg, model, sd


def rel_err(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-12))


def loader_fixture(name):
    with open(os.path.join(GOLDEN, f"loader_{name}.json")) as f:
        return json.load(f)


def loader_case_video(cfg, case):
    """The seeded tracklet-level video of a loader fixture case."""
    trk = synth.synthetic_tracklet_video(cfg, case["seed"], n_tracklets=case["n_tracklets"], n_frames=case["n_frames"])
    if case["n_dup"]:
        trk = synth.with_duplicates(trk, cfg, case["seed"], case["n_dup"])
    return trk
