"""Drop-in boundary checks that need no GPU: checkpoint schema, C-ABI exports, loud failure without CUDA."""
import ctypes
import os
import re

import pytest
import torch

from tests import helpers as H
from vrdone_b200 import MaskVRD, synth


@pytest.mark.parametrize("name", ["vidvrd", "vidor", "vidor_local", "vidor_x"])
def test_state_dict_schema_matches_reference_fixture(name):
    fix = H.network_fixture(name)
    model = MaskVRD(synth.load_config(name)["model_config"], "cpu")
    ours = {k: list(v.shape) for k, v in model.state_dict().items()}
    assert ours == fix["schema"]


@pytest.mark.skipif(not H.have_reference(), reason="/root/reference not present (GPU box)")
def test_state_dict_round_trip_with_live_reference():
    Ref = H.load_reference_class()
    mc = synth.load_config("vidor_x")["model_config"]
    ref, ours = Ref(mc, "cpu"), MaskVRD(mc, "cpu")
    ours.load_state_dict(ref.state_dict(), strict=True)       # reference checkpoint -> ours
    ref.load_state_dict(ours.state_dict(), strict=True)       # and back
    # default init statistics follow the reference (path scales 1e-4, zero conv biases, class prior bias)
    sd = ours.state_dict()
    assert float(sd["backbone.stem.0.drop_path_attn.scale"].mean()) == pytest.approx(1e-4)
    assert float(sd["predictor.class_embed.bias"][0]) == pytest.approx(-4.59511985, rel=1e-6)
    assert float(sd["backbone.so_fuse.layers.0.bias"].abs().max()) == 0.0


def test_cabi_library_exports_every_declared_symbol():
    from vrdone_b200 import build, cuda_ops
    lib_path = build.build()
    header = open(os.path.join(H.ROOT, "include", "vrdone_b200.h")).read()
    declared = set(re.findall(r"\b(vrd_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 19
    lib = ctypes.CDLL(lib_path)
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in include/vrdone_b200.h but not exported"
    assert declared == set(cuda_ops.exported_symbols())
    assert cuda_ops.load_library().vrd_abi_version() == 2


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_fails_loudly_without_cuda():
    cfg, model, _ = H.seeded_model("vidvrd", 1)
    video = synth.synthetic_video(cfg, 0)
    with pytest.raises(RuntimeError, match="CUDA"):
        model(video)


def test_training_mode_is_out_of_scope():
    cfg, model, _ = H.seeded_model("vidvrd", 1)
    model.train()
    with pytest.raises(NotImplementedError):
        model({})
