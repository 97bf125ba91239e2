"""End-to-end parity of the CUDA path through the drop-in module against the golden fixtures (outputs of the reference),
the oracle, and size-independent properties at benchmark scale."""
import numpy as np
import pytest
import torch

from oracle import maskvrd_oracle as O
from tests import helpers as H
from vrdone_b200 import synth
from vrdone_b200.layout import reference_padded_lengths

pytestmark = pytest.mark.gpu

NAMES = ["vidvrd", "vidor", "vidor_local", "vidor_x"]
# Tolerances (relative to the tensor's max magnitude).  fp32 path: BASELINE.json asks for 1e-3 on logits and mask
# probabilities.  bf16 path: operands of every GEMM are rounded to bf16 (8-bit mantissa, ~4e-3 relative per operand)
# through ~30 chained GEMM+LN layers; the bound is <= 2x the largest error measured over the four stress-init fixtures on
# a B200 (profiles/r2_parity.md lists the measured values per config).
TOL = {"fp32": (1e-3, 1e-3), "bf16": (1.3e-2, 7e-3)}      # measured bf16 maxima: 6.6e-3 (logits), 3.5e-3 (mask probabilities)
MEASURED = {}     # (test, case, precision) -> measured errors / rates, dumped to gpurun_out/parity_measured.json at session end


def _record(key, **vals):
    import json, os
    MEASURED["/".join(key)] = {k: (round(float(v), 6) if isinstance(v, float) else v) for k, v in vals.items()}
    os.makedirs(os.path.join(H.ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(H.ROOT, "gpurun_out", "parity_measured.json"), "w") as f:
        json.dump(MEASURED, f, indent=1, sort_keys=True)


def run_fixture(name, precision):
    fix = H.network_fixture(name)
    cfg, model, sd = H.seeded_model(name, fix["wseed"], precision=precision)
    model.to("cuda")
    feats = [f.cuda() for f in synth.pair_features(cfg["model_config"], fix["lens"], fix["xseed"])]
    r = model.run_network(feats, fix["tpads"], cfg["inference_config"]["topk"], want_masks=True)
    torch.cuda.synchronize()
    return fix, cfg, model, r


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name", NAMES)
def test_network_matches_reference_golden(name, precision):
    fix, cfg, model, r = run_fixture(name, precision)
    tol_l, tol_m = TOL[precision]
    logits = r["logits"].cpu()
    err_l = H.rel_err(logits, fix["pred_logits"])
    err_m = max(float((torch.sigmoid(m.t().cpu()) - torch.sigmoid(fix["pred_masks"][i])).abs().max()) for i, m in enumerate(r["masks"]))
    k = cfg["inference_config"]["topk"]
    ids_ref = torch.topk(torch.softmax(fix["pred_logits"], -1)[..., 1:], k, -1).indices + 1
    flips = total = 0
    for i, m in enumerate(r["masks"]):
        a, b = torch.sigmoid(m.t().cpu()) > 0.5, torch.sigmoid(fix["pred_masks"][i]) > 0.5
        flips += int((a != b).sum()); total += a.numel()
    topk_mismatch = float((r["topk_ids"].cpu().long() != ids_ref).float().mean())
    _record(("network", name, precision), logits_rel=err_l, mask_prob_abs=err_m, topk_mismatch_rate=topk_mismatch,
            mask_flips=flips, mask_elems=total)
    assert err_l < tol_l
    for i, m in enumerate(r["masks"]):
        pm, ref = torch.sigmoid(m.t().cpu()), torch.sigmoid(fix["pred_masks"][i])
        assert float((pm - ref).abs().max()) < tol_m, f"pair {i} (L={fix['lens'][i]})"
    # integer outputs are bit-exact GIVEN EQUAL LOGITS: recompute them on the CPU from the GPU's own logits
    probs = torch.softmax(logits, -1)[..., 1:]
    k = cfg["inference_config"]["topk"]
    ref_ids = torch.topk(probs, k, -1).indices + 1
    assert torch.equal(r["topk_ids"].cpu().long(), ref_ids)
    for i, m in enumerate(r["masks"]):
        act = torch.sigmoid(m.cpu()) > 0.5
        for q in range(act.shape[1]):
            nz = torch.nonzero(act[:, q]).flatten()
            exp = [int(nz[0]), int(nz[-1])] if nz.numel() else [-1, -1]
            assert r["first_last"][i, q].tolist() == exp
    # and against the reference's own logits they agree wherever the reference is not borderline (stress init: logit std ~0.6).
    # BASELINE.md section 2 gives the reference's own bf16-vs-fp64 figures for comparison: 24 / 8181 mask flips (2.9e-3).
    if precision == "fp32":
        assert topk_mismatch < 0.005 and flips <= max(1, total // 2000)
    else:
        # measured: top-k entries that differ from the reference's 1.0 % (vidor, k = 6 of 50) .. 6.1 % (vidvrd, k = 8 of 132);
        # mask flips <= 1.0e-3 of the frames (the reference's own bf16 run: 2.9e-3)
        assert topk_mismatch < 0.12 and flips <= max(2, total // 400)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_network_matches_reference_golden_default_init(precision):
    """The timing initialisation of bench.py (torch.manual_seed(0), path scales 1e-4, near-constant class logits): logits and
    mask probabilities against the reference; top-k order is noise at this init (the reference's own bf16
    run mismatches 166 / 216 top-k entries against fp64), so only its rate is recorded."""
    fix = H.network_fixture("vidor_default")
    cfg = synth.load_config(fix["config"])
    from vrdone_b200 import MaskVRD
    torch.manual_seed(0)
    model = MaskVRD(cfg["model_config"], "cpu").eval()
    assert H.checksum(model.state_dict().values()) == pytest.approx(fix["weights_checksum"], rel=1e-12)
    model._config_eval(cfg["inference_config"])
    model.set_precision(precision).to("cuda")
    feats = [f.cuda() for f in synth.pair_features(cfg["model_config"], fix["lens"], fix["xseed"])]
    k = cfg["inference_config"]["topk"]
    r = model.run_network(feats, fix["tpads"], k, want_masks=True)
    logits = r["logits"].cpu()
    err_l = H.rel_err(logits, fix["pred_logits"])
    err_m = max(float((torch.sigmoid(m.t().cpu()) - torch.sigmoid(fix["pred_masks"][i])).abs().max()) for i, m in enumerate(r["masks"]))
    ids_ref = torch.topk(torch.softmax(fix["pred_logits"], -1)[..., 1:], k, -1).indices + 1
    flips = sum(int(((torch.sigmoid(m.t().cpu()) > 0.5) != (torch.sigmoid(fix["pred_masks"][i]) > 0.5)).sum()) for i, m in enumerate(r["masks"]))
    _record(("network_default_init", "vidor", precision), logits_rel=err_l, mask_prob_abs=err_m,
            topk_mismatch_rate=float((r["topk_ids"].cpu().long() != ids_ref).float().mean()), mask_flips=flips,
            mask_elems=sum(m.numel() for m in r["masks"]))
    tol_l, tol_m = TOL[precision]
    assert err_l < tol_l and err_m < tol_m


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name", ["vidvrd", "vidor", "vidor_local", "vidor_x", "vidor_long"])
def test_forward_test_matches_reference_golden_video(name, precision):
    """``model(input_data)`` against the result dict of the UNMODIFIED reference on the same seeded video (fixtures made by
    tests/golden/make_golden.py): all four configs, plus a VidOR video with > max_so_pair pairs and long pairs (L > max_seq_len)
    in every 200-pair slice.  fp32: ranked triplets identical apart from borderline ties; bf16: the agreement rate is
    recorded and bounded from below (near-tied candidates swap ranks when logits move by ~1e-2)."""
    fix = H.video_fixture(name)
    cfg, model, sd = H.seeded_model(fix["config"], fix["wseed"], precision=precision)
    model.to("cuda")
    kw = {k: fix[k] for k in ("n_tracklets", "n_frames") if k in fix}
    video = synth.synthetic_video(cfg, fix["vseed"], **kw)
    assert [int(f.shape[1]) for f in video["so_features_list"]] == fix["lens"]
    dev_video = {k: ([t.cuda() for t in v] if isinstance(v, list) else (v.cuda() if torch.is_tensor(v) else v))
                 for k, v in video.items()}
    out = model(dev_video)
    ref = fix["output"]
    same = [a == b and c == d and e == f for a, b, c, d, e, f in
            zip(out["triplets"], ref["triplets"], out["pred_durations"], ref["pred_durations"], out["so_tids"], ref["so_tids"])]
    # set agreement (order-free): a reference triplet counts as found when the same (tids, triplet, duration) is reported at all
    key = lambda o, i: (tuple(o["so_tids"][i]), tuple(o["triplets"][i]), tuple(o["pred_durations"][i]))
    ours_set = {key(out, i) for i in range(len(out["triplets"]))}
    found = np.mean([key(ref, i) in ours_set for i in range(len(ref["triplets"]))])
    score_err = float(np.abs(np.array(out["triple_scores_avg"])[: len(ref["triplets"])] - np.array(ref["triple_scores_avg"])[: len(out["triplets"])]).max())
    _record(("forward_test", name, precision), n_triplets=len(out["triplets"]), n_ref=len(ref["triplets"]),
            same_rank_rate=float(np.mean(same)), found_rate=float(found), ranked_score_abs=score_err)
    if precision == "fp32":
        assert len(out["triplets"]) == len(ref["triplets"])
        assert np.mean(same) > 0.98, "ranked triplets differ from the reference beyond borderline ties"
        assert score_err < 2e-3
    else:
        # measured: every reference triplet is reported for 4 of the 5 videos (94.7 % for vidor_local), rank-for-rank agreement
        # 68-94 % (near-tied candidates swap), mean scores of equal ranks within 2.2e-4
        # below n_max_pair the NUMBER of triplets is the number of candidates that pass pred_min_frames: a mask flip at a duration's
        # edge (bf16: ~1e-3 of the frames) moves it by one
        assert abs(len(out["triplets"]) - len(ref["triplets"])) <= max(1, len(ref["triplets"]) // 50)
        assert found > 0.9 and score_err < 1e-3
    for t, (n, chk), ok in zip(out["so_trajs"], ref["so_trajs"], same):
        if ok:
            assert len(t[0]) == n and abs(float(torch.tensor(t).double().sum()) - chk) < 1e-3 * max(1.0, abs(chk))


def test_mask_vrd_tensor_boundary():
    fix = H.network_fixture("vidvrd")
    cfg, model, sd = H.seeded_model("vidvrd", fix["wseed"], precision="fp32")
    model.to("cuda")
    feats = synth.pair_features(cfg["model_config"], fix["lens"][:6], fix["xseed"])
    x, m = O.pad_batch(feats, 96)
    out = model._mask_vrd(x.cuda(), m.cuda())
    ref = O.mask_vrd(x, m, sd, cfg["model_config"])
    assert H.rel_err(out["pred_logits"].cpu(), ref["pred_logits"]) < 1e-3
    assert float((torch.sigmoid(out["pred_masks"].cpu()) - torch.sigmoid(ref["pred_masks"])).abs().max()) < 1e-3
    assert torch.equal(out["pred_masks"].cpu() == -10.0, ref["pred_masks"] == -10.0)


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_full_size_properties(precision):
    """BASELINE config 2 scale (a synthetic VidOR video, O(1000) pairs): results must not depend on how pairs are packed."""
    cfg = synth.load_config("vidor")
    from vrdone_b200 import MaskVRD
    torch.manual_seed(0)
    model = MaskVRD(cfg["model_config"], "cuda").eval().to("cuda")
    model._config_eval(cfg["inference_config"])
    model.set_precision(precision)
    n_trk = 40 if precision == "bf16" else 12
    video = synth.synthetic_video(cfg, 1, n_tracklets=n_trk, n_frames=1200)
    feats = [f.cuda() for f in video["so_features_list"]]
    lens = [int(f.shape[1]) for f in feats]
    tpads = reference_padded_lengths(lens, cfg["model_config"])
    k = cfg["inference_config"]["topk"]
    a = model.run_network(feats, tpads, k)
    b = model.run_network(feats, tpads, k)                                   # idempotence
    for key in ("logits", "topk_ids", "first_last", "topk_scores"):
        assert torch.equal(a[key], b[key]), key
    model.max_rows = 8192                                                     # different chunking
    c = model.run_network(feats, tpads, k)
    perm = torch.randperm(len(feats), generator=torch.Generator().manual_seed(0)).tolist()   # different row placement
    d = model.run_network([feats[i] for i in perm], [tpads[i] for i in perm], k)
    inv = torch.empty(len(perm), dtype=torch.long)
    inv[torch.tensor(perm)] = torch.arange(len(perm))
    for key in ("logits", "topk_ids", "first_last"):
        assert torch.equal(a[key], c[key]), key
        assert torch.equal(a[key], d[key][inv.cuda()]), key
    assert torch.isfinite(a["logits"]).all()
    fl = a["first_last"].cpu()
    L = torch.tensor(lens)[:, None]
    assert bool(((fl[..., 0] <= fl[..., 1]) & (fl[..., 1] < L) & ((fl[..., 0] >= 0) | (fl[..., 1] == -1))).all())


def test_host_resident_inputs_match_device_resident():
    """forward(input_data) with pinned HOST pair features (read by the pack kernel over PCIe), with pageable host features
    (copied first) and with device features must give identical results."""
    fix = H.video_fixture("vidvrd")
    cfg, model, sd = H.seeded_model("vidvrd", fix["wseed"], precision="bf16")
    model.to("cuda")
    video = synth.synthetic_video(cfg, fix["vseed"])
    on_dev = dict(video)
    on_dev["so_features_list"] = [t.cuda() for t in video["so_features_list"]]
    pinned = dict(video)
    pinned["so_features_list"] = [t.t().contiguous().pin_memory().t() for t in video["so_features_list"]]
    dense_pinned = dict(video)
    dense_pinned["so_features_list"] = [t.contiguous().pin_memory() for t in video["so_features_list"]]
    arena = dict(video)        # all pairs back to back in one pinned buffer: staged with one copy per run
    feats = video["so_features_list"]
    buf = torch.empty(sum(f.numel() for f in feats), pin_memory=True)
    views, pos = [], 0
    for f in feats:
        v = buf[pos:pos + f.numel()].view(f.shape[1], f.shape[0])
        v.copy_(f.t())
        views.append(v.t())
        pos += f.numel()
    arena["so_features_list"] = views
    a, b, c, d, e = model(on_dev), model(pinned), model(video), model(dense_pinned), model(arena)
    for other in (b, c, d, e):
        assert other["triplets"] == a["triplets"] and other["pred_durations"] == a["pred_durations"]
        assert other["triple_scores"] == a["triple_scores"] and other["so_trajs"] == a["so_trajs"]


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
@pytest.mark.parametrize("name", NAMES)
def test_native_schedule_equals_python_schedule(name, precision):
    """The C++ backbone schedule (csrc/engine.cu) issues the same kernels in the same order as engine.Engine.backbone:
    results must be bit-identical."""
    fix = H.network_fixture(name)
    cfg, model, sd = H.seeded_model(name, fix["wseed"], precision=precision)
    model.to("cuda")
    feats = [f.cuda() for f in synth.pair_features(cfg["model_config"], fix["lens"], fix["xseed"])]
    k = cfg["inference_config"]["topk"]
    model.use_native = True
    a = model.run_network(feats, fix["tpads"], k, want_masks=True)
    n_native = model._ops.launches
    model.use_native = False
    b = model.run_network(feats, fix["tpads"], k, want_masks=True)
    assert model._ops.launches - n_native == n_native, "both schedules must issue the same number of launches"
    for key in ("logits", "topk_ids", "first_last", "topk_scores"):
        assert torch.equal(a[key], b[key]), key
    for ma, mb in zip(a["masks"], b["masks"]):
        assert torch.equal(ma, mb)


@pytest.mark.parametrize("name,kw", [("vidvrd", {}), ("vidor", dict(n_tracklets=10, n_frames=600)), ("vidor_x", dict(n_tracklets=6, n_frames=400))])
def test_tracklet_pack_matches_loader_built_pairs(name, kw):
    """the pack stage fed from per-tracklet arrays must produce the operand rows the pair lists produce
    (feature rows bit-exact; box-geometry channels to float rounding: logf / division order on the device vs the CPU)."""
    from vrdone_b200.layout import PackLayout
    cfg, model, sd = H.seeded_model(name, 7, precision="fp32")
    model.to("cuda")
    mc = cfg["model_config"]
    trk = synth.synthetic_tracklet_video(cfg, 4, **kw)
    vid = synth.synthetic_video(cfg, 4, **kw)
    st = cfg["dataset_config"]["feat_stride"]
    eng = model._get_engine()
    nat = model._native
    keep, L, s_off, o_off = model.pair_table(trk["traj_durations"].numpy(), trk["sids"].numpy(), trk["oids"].numpy(), st, 0, 0)
    lens = L.tolist()
    tpads = reference_padded_lengths(lens, mc)
    lay = PackLayout(lens, tpads, model.n_levels, "cuda")
    import numpy as np
    n_frames = np.array([v.shape[0] for v in trk["visual_features_list"]])
    base = np.cumsum(n_frames) - n_frames
    tab = np.zeros((len(lens), 4), dtype=np.int32)
    tab[:, 0] = base[trk["sids"].numpy()] + s_off
    tab[:, 1] = base[trk["oids"].numpy()] + o_off
    tab[:, 2] = st
    vis_all = torch.cat(trk["visual_features_list"]).cuda()
    clip_all = torch.cat(trk["clip_features_list"]).cuda() if "clip_features_list" in trk else None
    boxes_all = torch.cat(trk["bboxes_list"]).cuda()
    a_top, a_mf = nat.backbone_tracklets(lay, vis_all, clip_all, boxes_all, torch.from_numpy(tab).cuda(), trk["video_wh"])
    feats = [f.cuda() for f in vid["so_features_list"]]
    ptrs = torch.tensor([f.data_ptr() for f in feats], dtype=torch.int64).cuda()
    strides = torch.tensor([[f.stride(0), f.stride(1)] for f in feats], dtype=torch.int64).cuda()
    b_top, b_mf = nat.backbone(lay, ptrs, strides, token_major=True)
    torch.cuda.synchronize()
    assert H.rel_err(a_top, b_top) < 2e-5 and H.rel_err(a_mf, b_mf) < 2e-5


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_forward_tracklets_matches_forward_on_pair_lists(precision):
    cfg, model, sd = H.seeded_model("vidor", 21, precision=precision)
    model.to("cuda")
    kw = dict(n_tracklets=8, n_frames=700)
    trk = synth.synthetic_tracklet_video(cfg, 3, **kw)
    vid = synth.synthetic_video(cfg, 3, **kw)
    dev_vid = {k: ([t.cuda() for t in v] if isinstance(v, list) else (v.cuda() if torch.is_tensor(v) else v)) for k, v in vid.items()}
    a = model(dev_vid)
    b = model.forward_tracklets(trk, cfg["dataset_config"])
    assert len(a["triplets"]) == len(b["triplets"])
    same = [x == y and u == v and p == q for x, y, u, v, p, q in
            zip(a["triplets"], b["triplets"], a["pred_durations"], b["pred_durations"], a["so_tids"], b["so_tids"])]
    # the box-geometry channels differ in the last bit (logf / division on the device vs the CPU); on the bf16 path that can
    # flip the rounding of an embedding and with it the order of near-tied candidates
    assert np.mean(same) > (0.98 if precision == "fp32" else 0.9)
    assert np.allclose(np.array(a["triple_scores_avg"]), np.array(b["triple_scores_avg"]), atol=1e-3 if precision == "fp32" else 1e-2)
    for ta, tb, ok in zip(a["so_trajs"], b["so_trajs"], same):
        if ok:
            assert ta == tb


def test_forward_returns_none_without_candidates_and_handles_many_slices():
    """More pairs than max_so_pair (several 200-pair slices with their own long-pair padding) against the oracle, and the
    ``None`` result when no candidate passes ``pred_min_frames`` (reference maskvrd.py:311-312)."""
    cfg, model, sd = H.seeded_model("vidvrd", 11, precision="fp32")
    model.to("cuda")
    video = synth.synthetic_video(cfg, 5, n_tracklets=17, n_frames=150)
    assert len(video["sids"]) > 200 and max(int(f.shape[1]) for f in video["so_features_list"]) > cfg["model_config"]["max_seq_len"]
    dev_video = {k: ([t.cuda() for t in v] if isinstance(v, list) else (v.cuda() if torch.is_tensor(v) else v)) for k, v in video.items()}
    out = model(dev_video)
    ref = O.forward_test(video, sd, cfg["model_config"], cfg["inference_config"])
    assert len(out["triplets"]) == len(ref["triplets"])
    same = [x == y and u == v and p == q for x, y, u, v, p, q in
            zip(out["triplets"], ref["triplets"], out["pred_durations"], ref["pred_durations"], out["so_tids"], ref["so_tids"])]
    assert np.mean(same) > 0.98
    model.pred_min_frames = 10 ** 6
    assert model(dev_video) is None


def test_pipelined_run_videos_equals_blocking_calls():
    """runner.run_videos keeps two videos in flight (device-resident, pinned-host and tracklet-level inputs mixed sizes): every
    result must equal the blocking ``model(video)`` call's, in input order, including ``None`` results."""
    from vrdone_b200 import runner
    cfg, model, sd = H.seeded_model("vidor", 21, precision="bf16")
    model.to("cuda")
    model.h2d_chunk_rows, model.h2d_edge_rows = 4096, 1024          # several chunks per video: staging buffers are reused across videos
    vids = [synth.synthetic_video(cfg, s, n_tracklets=n, n_frames=f) for s, n, f in ((1, 8, 700), (2, 5, 300), (3, 9, 900), (4, 6, 500))]

    def dev(v):
        return {k: ([t.cuda() for t in x] if isinstance(x, list) else (x.cuda() if torch.is_tensor(x) else x)) for k, x in v.items()}

    def pinned(v):
        return {k: ([t.t().contiguous().pin_memory().t() for t in x] if k == "so_features_list" else x) for k, x in v.items()}

    keys = ("triplets", "triple_scores", "triple_scores_avg", "so_trajs", "pred_durations", "so_tids")
    for make in (dev, pinned):
        inputs = [make(v) for v in vids] * 2
        blocking = [model(v) for v in inputs]
        for depth in (1, 2, 3):
            piped = list(runner.run_videos(model, iter(inputs), depth=depth))
            assert len(piped) == len(blocking)
            for a, b in zip(blocking, piped):
                assert all(a[k] == b[k] for k in keys)
    trk = [synth.synthetic_tracklet_video(cfg, s, n_tracklets=n, n_frames=f) for s, n, f in ((1, 8, 700), (2, 5, 300), (3, 9, 900))]
    blocking = [model.forward_tracklets(t, cfg["dataset_config"]) for t in trk]
    piped = list(runner.run_videos(model, trk, dataset_config=cfg["dataset_config"]))
    for a, b in zip(blocking, piped):
        assert all(a[k] == b[k] for k in keys)
    model.pred_min_frames = 10 ** 6
    assert list(runner.run_videos(model, [dev(v) for v in vids[:3]])) == [None, None, None]


def _viou_inputs(trk):
    import numpy as np
    from oracle import loader_oracle as LO
    boxes = LO.clamp_boxes(trk["bboxes_list"], trk["video_wh"])
    n_frames = np.array([b.shape[0] for b in boxes])
    base = torch.from_numpy(np.cumsum(n_frames) - n_frames).to(torch.int32).cuda()
    return boxes, torch.cat(boxes).cuda(), base, trk["traj_durations"].to(torch.int32).cuda(), trk["cat_ids"].to(torch.int32).cuda()


@pytest.mark.parametrize("name", ["vidor", "vidor_x", "vidvrd"])
def test_viou_filter_kernel_matches_loader_oracle(name):
    """volumes, rule decisions and the greedy scan of the device filter against the loader oracle (which is
    pinned to the reference's ``_val_getitem``), on the fixture videos with injected near-duplicates."""
    from oracle import loader_oracle as LO
    from vrdone_b200.cuda_ops import CudaOps
    ops = CudaOps()
    cfg = synth.load_config(name)
    for case in H.loader_fixture(name)["cases"]:
        trk = H.loader_case_video(cfg, case)
        boxes, boxes_d, base, durs, cats = _viou_inputs(trk)
        valid, flags, sums = ops.viou_filter(boxes_d, base, durs, cats, 0.9, torch.cuda.current_stream(), want_sums=True)
        torch.cuda.synchronize()
        assert valid.bool().tolist() == LO.duplicate_filter(boxes, trk["traj_durations"], trk["cat_ids"], 0.9)
        n, d, n_checked = len(boxes), trk["traj_durations"].tolist(), 0
        for b in range(n):
            for r in range(b + 1, n):
                if int(trk["cat_ids"][b]) != int(trk["cat_ids"][r]) or d[r][0] >= d[b][1] or d[r][1] <= d[b][0]:
                    assert int(flags[b, r]) == 0
                    continue
                s, e = max(d[b][0], d[r][0]), min(d[b][1], d[r][1])
                ref = LO.viou_sums(boxes[b][s - d[b][0]: e - d[b][0]].double(), boxes[r][s - d[r][0]: e - d[r][0]].double())
                got = sums[b, r].cpu()
                assert torch.allclose(got, torch.stack(ref), rtol=1e-6), (b, r)      # fp32 per-frame terms, fp64 sums
                n_checked += 1
        assert n_checked > 0 or case["n_dup"] == 0


def test_viou_filter_rules_on_device():
    from oracle import loader_oracle as LO
    from vrdone_b200.cuda_ops import CudaOps
    ops = CudaOps()
    box = torch.tensor([[100.0, 100.0, 200.0, 220.0]])
    for lens, shifts, durs, cats in (
            ([100, 40, 40, 40, 50, 50, 30, 60], [0, 0, 0, 500.0, 1.0, 1.0, 2.0, 2.0],
             [[0, 100], [10, 50], [10, 50], [10, 50], [200, 250], [200, 250], [310, 340], [300, 360]], [1, 1, 2, 1, 3, 3, 4, 4]),
            ([30, 60, 30], [0, 0, 0], [[10, 40], [0, 60], [10, 40]], [1, 1, 1]),
            ([5], [0], [[0, 5]], [7])):
        trk = {"bboxes_list": [box.repeat(n, 1) + s for n, s in zip(lens, shifts)], "video_wh": (1280.0, 720.0),
               "traj_durations": torch.tensor(durs), "cat_ids": torch.tensor(cats)}
        boxes, boxes_d, base, d, c = _viou_inputs(trk)
        valid, _, _ = ops.viou_filter(boxes_d, base, d, c, 0.9, torch.cuda.current_stream())
        assert valid.bool().tolist() == LO.duplicate_filter(boxes, trk["traj_durations"], trk["cat_ids"], 0.9)


@pytest.mark.parametrize("name", ["vidor", "vidor_x"])
def test_forward_tracklets_with_duplicate_filter_matches_loader_then_forward(name):
    """The whole 8f rows 1-2 path: forward_tracklets(viou_threshold=0.9) on raw tracklets == forward on the item the loader
    oracle builds (duplicate filter + pair construction on the CPU), host- and device-resident tracklet inputs."""
    from oracle import loader_oracle as LO
    fix = H.loader_fixture(name)
    cfg, model, sd = H.seeded_model(name, 21, precision="fp32")
    model.to("cuda")
    case = fix["cases"][0]
    trk = H.loader_case_video(cfg, case)
    item = LO.val_getitem(trk, case["feat_stride"], 0, case["proposal_min_frames"], 0.9, with_clip="clip_features_list" in trk)
    assert item["sids"].tolist() == case["sids"]
    dev_item = {k: ([t.cuda() for t in v] if isinstance(v, list) and torch.is_tensor(v[0]) else (v.cuda() if torch.is_tensor(v) else v))
                for k, v in item.items() if k != "valid_tracklets"}
    a = model(dev_item)
    dc = dict(cfg["dataset_config"], viou_threshold=0.9, proposal_min_frames=case["proposal_min_frames"])
    dev_trk = {k: ([t.cuda() for t in v] if isinstance(v, list) else (v.cuda() if torch.is_tensor(v) else v)) for k, v in trk.items()}
    for inp in (trk, dev_trk):
        b = model.forward_tracklets(inp, dc)
        assert len(a["triplets"]) == len(b["triplets"])
        same = [x == y and u == v and p == q for x, y, u, v, p, q in
                zip(a["triplets"], b["triplets"], a["pred_durations"], b["pred_durations"], a["so_tids"], b["so_tids"])]
        assert np.mean(same) > 0.98            # box-geometry channels differ in the last bit (device logf / division)
        assert np.allclose(np.array(a["triple_scores_avg"]), np.array(b["triple_scores_avg"]), atol=1e-3)
        for ta, tb, ok in zip(a["so_trajs"], b["so_trajs"], same):
            if ok:
                assert ta == tb
    unfiltered = model.forward_tracklets(trk, cfg["dataset_config"])
    removed = {i for i, v in enumerate(item["valid_tracklets"]) if not v}
    assert removed and not any(s in removed or o in removed for s, o in b["so_tids"])
    assert unfiltered is not None


def test_inputs_may_be_dropped_right_after_submit():
    """host pair features are read by raw cudaMemcpyAsync calls that torch's pinned allocator knows nothing
    about; ``submit()`` must keep them alive until the result is ready, because a ``for v in loader: submit(v)`` loop rebinds
    ``v`` at once and the pinned blocks would be recycled (and refilled) under the DMA."""
    import gc
    import weakref
    from vrdone_b200 import runner
    cfg, model, sd = H.seeded_model("vidor", 21, precision="bf16")
    model.to("cuda")
    model.h2d_chunk_rows, model.h2d_edge_rows = 4096, 1024
    base = synth.synthetic_video(cfg, 2, n_tracklets=9, n_frames=900)

    def pinned_copy():
        v = dict(base)
        v["so_features_list"] = [t.t().contiguous().pin_memory().t() for t in base["so_features_list"]]
        return v

    expect = model(pinned_copy())
    v = pinned_copy()
    ref0 = weakref.ref(v["so_features_list"][0])
    pending = model.submit(v)
    del v
    gc.collect()
    assert ref0() is not None, "submit() must hold the host tensors until the copies have completed"
    # recycle pinned memory aggressively while the DMA may still be running: same-sized blocks are handed out again at once
    junk = [torch.full_like(t.t().contiguous(), float("nan")).pin_memory() for t in base["so_features_list"][:64]]
    out = pending.result()
    assert pending._keep is None       # ... and lets go of them once the result is out
    del junk
    keys = ("triplets", "triple_scores", "so_trajs", "pred_durations", "so_tids")
    assert all(out[k] == expect[k] for k in keys)
    # and through the pipelined loop with a generator that keeps no reference of its own
    outs = list(runner.run_videos(model, (pinned_copy() for _ in range(3))))
    assert all(o[k] == expect[k] for o in outs for k in keys)


def test_lazy_trajs_and_eval_format_on_gpu():
    """SURVEY 8f row 3 end to end on the device path: ``lazy_trajs`` results equal the eager ones, and the reference's
    per-video conversion (utils/evaluate.py:38-73) consumes both."""
    from vrdone_b200.eval_format import EvaluationFormatConvertor
    from vrdone_b200.maskvrd import LazyTrajs
    cfg, model, sd = H.seeded_model("vidor", 21, precision="bf16")
    model.to("cuda")
    video = synth.synthetic_video(cfg, 3, n_tracklets=6, n_frames=700, name="0001_3598080384")
    dev_video = {k: ([t.cuda() for t in v] if isinstance(v, list) else (v.cuda() if torch.is_tensor(v) else v)) for k, v in video.items()}
    eager = model(dev_video)
    model.lazy_trajs = True
    lazy = model(dev_video)
    model.lazy_trajs = False
    assert isinstance(lazy["so_trajs"], LazyTrajs) and isinstance(eager["so_trajs"], list)
    for k in ("triplets", "triple_scores", "triple_scores_avg", "pred_durations", "so_tids"):
        assert eager[k] == lazy[k]
    assert lazy["so_trajs"] == eager["so_trajs"]
    conv = EvaluationFormatConvertor("vidor")
    a = conv.to_eval_format_pr(video["video_name"], eager)
    b = conv.to_eval_format_pr(video["video_name"], lazy)
    assert list(a) == ["3598080384"] and a == b
    rel = a["3598080384"][0]
    assert set(rel) == {"triplet", "duration", "score", "sub_traj", "obj_traj"}
    assert len(rel["sub_traj"]) == rel["duration"][1] - rel["duration"][0] and len(rel["sub_traj"][0]) == 4
    arr = EvaluationFormatConvertor("vidor", trajs="array").to_eval_format_pr(video["video_name"], lazy)["3598080384"]
    assert all(np.array_equal(np.asarray(x["sub_traj"], dtype=np.float32), y["sub_traj"]) for x, y in zip(a["3598080384"], arr))
    assert conv.to_eval_format_pr(video["video_name"], None) == {"3598080384": []}


@pytest.mark.parametrize("name,kw", [("vidor", dict(n_tracklets=14, n_frames=900)), ("vidvrd", dict(n_tracklets=17, n_frames=150)),
                                     ("vidor_x", dict(n_tracklets=6, n_frames=500))])
def test_device_ranking_equals_host_ranking(name, kw):
    """csrc/rank.cu (candidate filter, fp32 mean score, top-n_max_pair with ties to the earlier candidate) against the dense numpy
    ranking of ``_decode`` on the same kernel outputs: identical result dicts, for device- and host-resident inputs, the tracklet
    entry point, private / shared box lists, and the ``None`` result."""
    cfg, model, sd = H.seeded_model(name, 21, precision="bf16")
    model.to("cuda")
    video = synth.synthetic_video(cfg, 5, **kw)
    dev_video = {k: ([t.cuda() for t in v] if isinstance(v, list) else (v.cuda() if torch.is_tensor(v) else v)) for k, v in video.items()}
    keys = ("triplets", "triple_scores", "triple_scores_avg", "so_trajs", "pred_durations", "so_tids")
    model.device_rank = False
    dense = model(dev_video)
    model.device_rank = True
    assert dense is not None and len(dense["triplets"]) == min(model.n_max_pair, len(dense["triplets"]))
    for inp in (dev_video, video):
        for private in (False, True):
            model.private_box_lists = private
            out = model(inp)
            assert all(out[k] == dense[k] for k in keys)
    model.private_box_lists = False
    if name != "vidvrd":
        trk = synth.synthetic_tracklet_video(cfg, 5, **kw)
        model.device_rank = False
        dense_t = model.forward_tracklets(trk, cfg["dataset_config"])
        model.device_rank = True
        out_t = model.forward_tracklets(trk, cfg["dataset_config"])
        assert all(out_t[k] == dense_t[k] for k in keys)
    # fewer kept candidates than n_max_pair, and none at all
    model.n_max_pair = 1000
    model.pred_min_frames = int(0.9 * max(int(f.shape[1]) for f in video["so_features_list"]) * cfg["inference_config"]["feat_stride"])
    model.device_rank = False
    few_dense = model(dev_video)
    model.device_rank = True
    few = model(dev_video)
    assert (few is None) == (few_dense is None)
    if few is not None:
        assert len(few["triplets"]) <= 1000 and all(few[k] == few_dense[k] for k in keys)
    model.pred_min_frames = 10 ** 6
    assert model(dev_video) is None
