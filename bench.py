#!/usr/bin/env python
"""Benchmark of the MaskVRD inference hot path: relation pairs/sec (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config vidor] [--precision bf16|fp32]

Workload (BASELINE.json configs[1]): a fixed, seeded SET of synthetic VidOR-shaped videos --
frame counts F in {900, 1200, 1800, 3600} (.3/.4/.2/.1), N ~ U[20, 60] tracklets, every ordered pair with temporal overlap,
feat_stride 4; the 3600-frame video carries pairs longer than max_seq_len (the reference's long-batch path, O(L^2) SOS
attention).  A "step" is ONE PASS of ``model(input_data)`` (network + heads epilogue + ranking + triplet decoding) over every
video of the set.  ``--tracklets T --frames F`` selects the fixed-shape round-1 workload instead.  Prints ONE JSON line
(rank 0); details go to stderr.  For N > 1 launch with torchrun: every rank processes its own copy of the set (weak scaling,
no collective on the data path), time = max over ranks.

  value        whole-job pairs/s, pair features already resident in HBM (CUDA events around the K steps)
  e2e          the same through the public API with HOST (pinned) pair features: H2D copies + D2H of results inside;
               .blocking_value = the reference's eval loop (one blocking model(input_data) after the other);
               .tracklet_api_value = MaskVRD.forward_tracklets with host tracklet features
  roofline     dominant kernel (tcgen05 GEMM): algorithmic FLOPs of its launches / their CUDA-event durations vs the measured
               bf16 peak; per-kernel time table; HBM fractions of the memory-bound kernels; attention TFLOP/s
  parity       GPU outputs vs the oracle on the cpu_baseline sample (same pairs, same timing initialisation)
  sweep        BASELINE config 5: a fixed set of distinct videos sharded by LPT over the ranks (strong scaling) through
               runner.run_sharded + gather_object, inside the timed region
  cpu_baseline the oracle port of the reference forward on the host cores, on a bounded random sample of the same workload
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from vrdone_b200 import synth, runner  # noqa: E402

METRIC = "relation pairs/sec, VrdONE forward"
UNIT = "pairs/s"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--config", default="vidor", choices=list(synth.CONFIG_NAMES))
    p.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    p.add_argument("--set-videos", type=int, default=10, help="videos in the cfg2 set (one step = one pass over the set)")
    p.add_argument("--set-seed", type=int, default=0)
    p.add_argument("--extra-seeds", default="", help="comma-separated set seeds: `value` (HBM-resident, pipelined loop) is also measured on "
                                                      "the cfg2 sets drawn with these seeds and reported under `seeds` (SURVEY 8d: seeds {0,1,2})")
    p.add_argument("--tracklets", type=int, default=None, help="fixed-shape workload: tracklets per video (with --frames)")
    p.add_argument("--frames", type=int, default=None)
    p.add_argument("--videos", type=int, default=2, help="fixed-shape workload: distinct videos per step")
    p.add_argument("--cpu-pairs", type=int, default=40, help="pairs in the bounded CPU-baseline sample")
    p.add_argument("--sweep-videos", type=int, default=48, help="videos of the config-5 strong-scaling sweep (0: skip)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-extras", action="store_true", help="skip the blocking / tracklet / network-only / sweep legs")
    return p.parse_args()


def log(*a):
    if int(os.environ.get("RANK", "0")) == 0:
        print(*a, file=sys.stderr, flush=True)


def peaks():
    """Roofline denominators: MEASURED_PEAKS.json (driver-written) or the fallback B200_PROFILING.md states."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        if all(k in d for k in ("bf16_tflops_sustained", "bf16_tflops", "hbm_gbs")):
            return d["bf16_tflops_sustained"], d["bf16_tflops"], d["hbm_gbs"], "measured"
    return 1400.0, 1590.0, 6650.0, "fallback"      # /opt/skills/guides/B200_PROFILING.md: 1.59 PF burst / ~1.4 sustained, 6.65 TB/s


def gemm_traffic(default_workload: bool):
    """DRAM bytes per GEMM launch (dram__bytes_read.sum + dram__bytes_write.sum, mean over the launches of one forward) from the
    committed ncu capture of the default workload; None otherwise."""
    for name in ("r2_gemm_traffic.json", "r1_gemm_traffic.json"):
        path = os.path.join(ROOT, "profiles", name)
        if default_workload and os.path.exists(path):
            with open(path) as f:
                d = json.load(f)
            return d["traffic_bytes_per_launch"], name
    return None, None


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._stop_evt = index, [], set(), None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        # NVML queries contend with the CUDA driver for a few ms each: a handful of samples per timed region, not a busy poll
        self._stop_evt.wait(0.02)
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, n in names.items():
                    if r & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        if not self.samples and self.nv is not None:      # a region shorter than the first sampling delay: sample at its end
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            except Exception:
                pass
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def video_specs(args):
    """[(video seed, n_frames, n_tracklets)] of one step."""
    if args.tracklets is not None or args.frames is not None:
        return [(v, args.frames or 1200, args.tracklets or 40) for v in range(args.videos)]
    return synth.cfg2_video_set(args.set_videos, args.set_seed)


def workload_name(args, cfg):
    st = cfg["dataset_config"]["feat_stride"]
    if args.tracklets is not None or args.frames is not None:
        return (f"{args.config}.yaml forward_test, synthetic VidOR-shaped videos: {args.tracklets or 40} tracklets x {args.frames or 1200} "
                f"frames, feat_stride {st}, all ordered overlapping pairs; one step = {args.videos} videos")
    return (f"{args.config}.yaml forward_test, SURVEY 8d cfg2 set: {args.set_videos} synthetic VidOR-shaped videos, F in "
            f"{{900,1200,1800,3600}} frames (.3/.4/.2/.1), N ~ U[20,60] tracklets, feat_stride {st}, all ordered overlapping pairs "
            f"(set seed {args.set_seed}); one step = one pass over the set")


def to_device(video, dev):
    return {k: ([t.to(dev) for t in v] if isinstance(v, list) else (v.to(dev) if torch.is_tensor(v) else v)) for k, v in video.items()}


def pin_pairs(video):
    """Pinned host copy of the pair features in the data loader's layout -- every pair a (C, L) view of its own (L, C)-contiguous
    block (reference vidor.py:708-711; DataLoader(pin_memory=True) preserves strides) -- carved out of ONE pinned allocation per
    video with a gap between pairs, so that pairs are NOT adjacent in memory and travel as one copy each, exactly like separately
    pinned tensors (12 k cudaHostAlloc calls per set would only slow the benchmark's setup down)."""
    feats = video["so_features_list"]
    gap = 64                                                  # floats (256 bytes): keeps pairs from being coalesced into runs
    total = sum(f.numel() + gap for f in feats)
    arena = torch.empty(total, dtype=torch.float32, pin_memory=True)
    views, pos = [], 0
    for f in feats:
        C, L = f.shape
        v = arena[pos:pos + L * C].view(L, C)
        v.copy_(f.t())
        views.append(v.t())
        pos += L * C + gap
    out = dict(video)
    out["so_features_list"] = views
    return out


def sample_pairs(video, idx):
    """``input_data`` restricted to the pairs ``idx`` (tracklet-level entries stay)."""
    sub = dict(video)
    sub["so_features_list"] = [video["so_features_list"][i] for i in idx]
    for k in ("sids", "oids", "so_offset"):
        sub[k] = video[k][torch.as_tensor(idx, dtype=torch.long)]
    return sub


def cpu_sample(host_videos, n_pairs, seed):
    """A bounded random sample of the workload for the CPU legs: ``n_pairs`` pairs drawn without replacement from the pairs of
    ALL videos of the step, grouped by video (the oracle, like the reference, works video by video)."""
    g = np.random.default_rng(seed)
    owners = np.concatenate([np.full(len(v["sids"]), i) for i, v in enumerate(host_videos)])
    local = np.concatenate([np.arange(len(v["sids"])) for v in host_videos])
    pick = np.sort(g.choice(len(owners), size=min(n_pairs, len(owners)), replace=False))
    return [(int(i), local[pick[owners[pick] == i]].tolist()) for i in np.unique(owners[pick])]


def cpu_forward(cfg, sd, host_videos, sample, keep_outputs=False):
    """Oracle port of the reference forward_test over a sample: ONE batched network pass over all sampled pairs (the reference
    runs a slice of up to 200 pairs per network call, maskvrd.py:201-240; the oracle works in sub-batches of 16), then the
    reference's candidate loop + ranking per video.  Returns (pairs, seconds, [(video, pair indices, logits, masks)])."""
    from oracle import maskvrd_oracle as O
    mc = cfg["model_config"]
    subs = [(vi, idx, sample_pairs(host_videos[vi], idx)) for vi, idx in sample]
    feats = [f for _, _, sub in subs for f in sub["so_features_list"]]
    t0 = time.perf_counter()
    with torch.no_grad():
        logits, masks = O.network_outputs(feats, sd, mc)
        pos = 0
        for vi, idx, sub in subs:
            O.decode_triplets(logits[pos:pos + len(idx)], masks[pos:pos + len(idx)], sub, cfg["inference_config"])
            pos += len(idx)
    dt = time.perf_counter() - t0
    outs, pos = [], 0
    if keep_outputs:
        for vi, idx, sub in subs:
            outs.append((vi, idx, logits[pos:pos + len(idx)], masks[pos:pos + len(idx)]))
            pos += len(idx)
    return len(feats), dt, outs


def default_state_dict(mc):
    from vrdone_b200 import MaskVRD
    torch.manual_seed(0)
    return {k: v.detach().clone() for k, v in MaskVRD(mc, "cpu").state_dict().items()}


def parity_block(model, cfg, dev_videos, outs):
    """GPU (benchmarked precision) vs oracle on the CPU sample: same pairs, same weights, the oracle's padded lengths."""
    from oracle import maskvrd_oracle as O
    mc = cfg["model_config"]
    k = cfg["inference_config"]["topk"]
    lmax = mmax = 0.0
    ref_all, mis, tot_ids, flips, tot_m, n = [], 0, 0, 0, 0, 0
    all_lens = [int(dev_videos[vi]["so_features_list"][i].shape[1]) for vi, idx, _, _ in outs for i in idx]
    all_tpads = O.padded_lengths(all_lens, mc)         # the oracle padded the merged sample: long pairs to ITS longest pair
    pos = 0
    for vi, idx, logits, masks in outs:
        feats = [dev_videos[vi]["so_features_list"][i] for i in idx]
        r = model.run_network(feats, all_tpads[pos:pos + len(idx)], k, want_masks=True)
        pos += len(idx)
        torch.cuda.synchronize()
        ref = torch.stack(logits)
        got = r["logits"].cpu()
        lmax = max(lmax, float((got - ref).abs().max()))
        ref_all.append(ref)
        ids_ref = torch.topk(torch.softmax(ref, -1)[..., 1:], k, -1).indices + 1
        mis += int((r["topk_ids"].cpu().long() != ids_ref).sum()); tot_ids += ids_ref.numel()
        for m, rm in zip(r["masks"], masks):
            a, b = torch.sigmoid(m.t().cpu()), torch.sigmoid(rm)
            mmax = max(mmax, float((a - b).abs().max()))
            flips += int(((a > 0.5) != (b > 0.5)).sum()); tot_m += a.numel()
        n += len(idx)
    scale = max(float(torch.cat(ref_all).abs().max()), 1e-12)
    return {"pairs": n, "logits_rel": lmax / scale, "mask_prob_abs": mmax, "topk_mismatch": mis, "topk_entries": tot_ids,
            "mask_flips": flips, "mask_elems": tot_m, "init": "torch.manual_seed(0) default init (near-constant class logits: top-k order "
            "is noise, BASELINE.md section 2: the reference's own bf16 run mismatches 166/216)", "vs": "oracle port, fp32"}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = synth.load_config(args.config)
    mc = cfg["model_config"]
    threads = os.cpu_count() or 1
    workload = workload_name(args, cfg)
    specs = video_specs(args)
    default_workload = (args.config == "vidor" and args.precision == "bf16" and args.tracklets is None and args.frames is None
                        and args.set_videos == 10 and args.set_seed == 0)

    if args.impl == "reference":
        # the reference's own CPU implementation of the path (oracle port: the reference is a Python package that cannot travel),
        # all host threads, each step a fresh bounded random sample of the SAME workload
        if rank != 0:
            return
        torch.set_num_threads(threads)
        st = cfg["dataset_config"].get("feat_stride", 1)
        counts = [len(synth._overlapping_pairs(synth._tracklets(cfg, s, nt, nf)[2], st)) for s, nf, nt in specs]
        sd = default_state_dict(mc)
        n_step = max(8, args.cpu_pairs // 2)

        def one_step(seed):     # only the sampled pairs' features are built (the full set is 11 GB)
            sample = cpu_sample([{"sids": range(c)} for c in counts], n_step, seed)
            vids = {vi: synth.synthetic_video(cfg, specs[vi][0], n_tracklets=specs[vi][2], n_frames=specs[vi][1], only_pairs=idx)
                    for vi, idx in sample}
            return cpu_forward(cfg, sd, vids, [(vi, list(range(len(idx)))) for vi, idx in sample])

        for w in range(max(0, min(args.warmup, 1))):
            one_step(1000 + w)
        tot_pairs = tot_t = 0.0
        for s in range(args.steps):
            n, dt, _ = one_step(s)
            tot_pairs += n
            tot_t += dt
        val = tot_pairs / tot_t
        sample = (f"{n_step} pairs per step drawn at random (seeded, without replacement) from all pairs of the step's videos, padded as the "
                  f"reference pads them (short pairs to max_seq_len, long pairs to the sample's longest)")
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": workload, "sample": sample},
                          "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
                          "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    from vrdone_b200 import MaskVRD
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL announces its version on stdout at the first communicator: keep stdout for the ONE JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    torch.manual_seed(0)
    model = MaskVRD(mc, dev).eval().to(dev)
    model._config_eval(cfg["inference_config"])
    model.set_precision(args.precision)
    Q, topk = mc["predictor"]["num_queries"], cfg["inference_config"]["topk"]

    t_setup = time.perf_counter()
    host_videos, pinned, dev_videos = [], [], []
    for s, nf, nt in specs:      # weak scaling: every rank works on its own copy of the SAME videos
        v = synth.synthetic_video(cfg, s, n_tracklets=nt, n_frames=nf)
        pv = pin_pairs(v)
        pinned.append(pv)
        dev_videos.append(to_device(pv, dev))
        host_videos.append(pv)          # the pinned copy doubles as the host copy (CPU legs read it)
        del v
    lens_all = [[int(f.shape[1]) for f in v["so_features_list"]] for v in host_videos]
    n_pairs = [len(l) for l in lens_all]
    frames = [sum(l) for l in lens_all]
    flops = [runner.video_cost(args.config, l) for l in lens_all]
    in_bytes = [sum(f.numel() * 4 for f in v["so_features_list"]) for v in host_videos]
    long_pairs = [sum(1 for x in l if x > mc["max_seq_len"]) for l in lens_all]
    log(f"[bench] rank {rank}: {len(specs)} videos, {sum(n_pairs)} pairs, {sum(frames)} valid frames, {sum(in_bytes) / 1e9:.2f} GB of pair "
        f"features per step, {sum(flops) / 1e12:.1f} TFLOP per step; setup {time.perf_counter() - t_setup:.1f} s")

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    step_ms = []

    def timed(videos, steps, pipelined=True, dataset_config=None):
        """K steps (passes over ``videos``) between two CUDA events.  ``pipelined``: through ``runner.run_videos`` (two videos in
        flight: the decode of one overlaps the kernels of the next; every result is produced and dropped inside the timed
        region); otherwise one synchronous ``model(v)`` call after the other, as the reference's eval loop does.  Returns
        (ms max over ranks, whole-job pairs)."""
        sync_all()
        step_ms.clear()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        tw = time.perf_counter()
        if pipelined:
            n_v = len(videos)
            for i, out in enumerate(runner.run_videos(model, (videos[j % n_v] for j in range(steps * n_v)), dataset_config=dataset_config)):
                del out     # dropping the result's Python objects is part of the step
                if (i + 1) % n_v == 0:
                    step_ms.append(round(1e3 * (time.perf_counter() - tw), 1))
                    tw = time.perf_counter()
        else:
            for s in range(steps):
                for v in videos:
                    # host-resident ``v``: the module moves the pair features (inside the timed region)
                    out = model(v) if dataset_config is None else model.forward_tracklets(v, dataset_config)
                    del out
                step_ms.append(round(1e3 * (time.perf_counter() - tw), 1))
                tw = time.perf_counter()
        e1.record()
        sync_all()
        return reduce_max(e0.elapsed_time(e1)), float(world * steps * sum(n_pairs))

    def warm(videos, dataset_config=None, passes=None):
        """W untimed steps through BOTH loops: the pipelined one keeps two videos' pinned read-back / layout buffers alive at
        once, and the first cudaHostAlloc of those costs ~100 ms."""
        passes = args.warmup if passes is None else passes
        for s in range(max(1, (passes + 1) // 2)):
            for v in videos:
                model(v) if dataset_config is None else model.forward_tracklets(v, dataset_config)
        for out in runner.run_videos(model, (videos[j % len(videos)] for j in range(max(1, passes // 2) * len(videos))),
                                     dataset_config=dataset_config):
            del out

    ops = None
    warm(dev_videos)
    ops = model._ops
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = ops.launches
    ms, pairs = timed(dev_videos, args.steps)
    launches = ops.launches - l0
    steps_value = list(step_ms)
    host_pipe = {k: round(v, 2) for k, v in model.last_stats.items()}
    clocks = sampler.stop()
    log(f"[bench] value: {pairs / ms * 1e3:.0f} pairs/s, {ms / args.steps:.1f} ms/step, per-step wall ms {steps_value}, host stats of the last "
        f"video {host_pipe}")

    warm(pinned)                        # staging buffers, pinned upload blocks and the copy streams are created on first use
    ms_e2e, pairs_e2e = timed(pinned, args.steps)
    steps_e2e = list(step_ms)
    log(f"[bench] e2e: {pairs_e2e / ms_e2e * 1e3:.0f} pairs/s, per-step wall ms {steps_e2e}, host stats {model.last_stats}")

    extras = not args.no_extras
    e2e = {"value": pairs_e2e / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(sum(in_bytes)),
           "d2h_bytes_per_step": int(len(specs) * 4 * (4 + 6 * cfg["inference_config"]["n_max_pair"])),
           "api": "runner.run_videos(model, videos): submit video i+1, then wait for + decode video i (two videos in flight; results "
                  "identical to [model(v) for v in videos])"}
    out_extra = {}
    if extras:
        ms_sync, pairs_sync = timed(dev_videos, args.steps, pipelined=False)
        host_sync = {k: round(v, 2) for k, v in model.last_stats.items()}
        ms_e2e_sync, pairs_e2e_sync = timed(pinned, args.steps, pipelined=False)
        e2e["blocking_value"] = pairs_e2e_sync / (ms_e2e_sync * 1e-3)
        out_extra["blocking_call"] = {"note": "one blocking model(input_data) call after the other (the reference's eval loop, eval.py:140-152)",
                                      "value": pairs_sync / (ms_sync * 1e-3), "e2e": e2e["blocking_value"], "unit": UNIT}
        log(f"[bench] blocking: {out_extra['blocking_call']}, host stats {host_sync}")

        # the same videos through the tracklet-level entry point (per-tracklet features cross PCIe once; the pair
        # gather and the box geometry run on the device).  Reported beside the headline numbers, never instead of them.
        trk_videos = []
        for s, nf, nt in specs:
            t = synth.synthetic_tracklet_video(cfg, s, n_tracklets=nt, n_frames=nf)
            for key in ("visual_features_list", "clip_features_list", "bboxes_list"):
                if key in t:
                    t[key] = [x.pin_memory() for x in t[key]]
            trk_videos.append(t)
        trk_bytes = sum(x.numel() * 4 for t in trk_videos for key in ("visual_features_list", "clip_features_list", "bboxes_list") if key in t
                        for x in t[key])
        dc = cfg["dataset_config"]
        warm(trk_videos, dc, passes=2)
        ms_trk, pairs_trk = timed(trk_videos, args.steps, dataset_config=dc)
        ms_trk_sync, _ = timed(trk_videos, args.steps, pipelined=False, dataset_config=dc)
        e2e.update({"tracklet_api_value": pairs_trk / (ms_trk * 1e-3), "tracklet_api_blocking_value": pairs_trk / (ms_trk_sync * 1e-3),
                    "tracklet_api_h2d_bytes_per_step": int(trk_bytes)})
        log(f"[bench] tracklet api: {e2e['tracklet_api_value']:.0f} pairs/s pipelined, {e2e['tracklet_api_blocking_value']:.0f} blocking")
        del trk_videos

        # network only (the `_mask_vrd`-equivalent part, no ranking / decode): kernels of all steps back to back
        from vrdone_b200.layout import reference_padded_lengths
        net_in = [(v["so_features_list"], reference_padded_lengths(l, mc)) for v, l in zip(dev_videos, lens_all)]
        for feats_v, tp_v in net_in:
            model.run_network(feats_v, tp_v, model.topk)
        sync_all()
        net_sampler = ClockSampler(local_rank)       # this loop keeps the GPU at 100 % duty: power capping shows up here first
        net_sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(args.steps):
            for feats_v, tp_v in net_in:
                model.run_network(feats_v, tp_v, model.topk)
        e1.record()
        sync_all()
        net_clocks = net_sampler.stop()
        ms_net = reduce_max(e0.elapsed_time(e1))
        out_extra["network_only"] = {"value": world * args.steps * sum(n_pairs) / (ms_net * 1e-3), "unit": UNIT,
                                     "valid_frames_per_s": world * args.steps * sum(frames) / (ms_net * 1e-3),
                                     "ms_per_step": ms_net / args.steps, "clocks": net_clocks}
        log(f"[bench] network only: {out_extra['network_only']}")

    # other draws of the cfg2 set (different videos, pair counts and lengths), same loop, per-step wall times kept
    seeds_out = None
    if args.extra_seeds and args.tracklets is None and args.frames is None:
        seeds_out = {str(args.set_seed): {"value": pairs / (ms * 1e-3), "pairs_per_step": int(sum(n_pairs)), "ms_per_step": ms / args.steps,
                                          "step_wall_ms": steps_value}}
        for sd in [int(x) for x in args.extra_seeds.split(",") if x.strip() != ""]:
            if sd == args.set_seed:
                continue
            vids = []
            for s_, nf, nt in synth.cfg2_video_set(args.set_videos, sd):
                vids.append(to_device(synth.synthetic_video(cfg, s_, n_tracklets=nt, n_frames=nf), dev))
            n_sd = sum(len(v["sids"]) for v in vids)
            warm(vids, passes=2)
            sync_all()
            step_ms.clear()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            tw = time.perf_counter()
            for i, out_v in enumerate(runner.run_videos(model, (vids[j % len(vids)] for j in range(args.steps * len(vids))))):
                del out_v
                if (i + 1) % len(vids) == 0:
                    step_ms.append(round(1e3 * (time.perf_counter() - tw), 1))
                    tw = time.perf_counter()
            e1.record()
            sync_all()
            ms_sd = reduce_max(e0.elapsed_time(e1))
            seeds_out[str(sd)] = {"value": world * args.steps * n_sd / (ms_sd * 1e-3), "pairs_per_step": int(n_sd),
                                  "ms_per_step": ms_sd / args.steps, "step_wall_ms": list(step_ms)}
            del vids
            torch.cuda.empty_cache()
        vals = sorted(v["value"] for v in seeds_out.values())
        seeds_out["median_value"] = vals[len(vals) // 2]
        log(f"[bench] seeds: {seeds_out}")

    # roofline pass: a CUDA-event pair around every launch (the per-operator Python schedule: same kernels, same order)
    sustained, burst, hbm, src = peaks()
    model.use_native = False
    n_prof = min(args.steps, 2)
    ops.start_timing()
    for s in range(n_prof):
        for v in dev_videos:
            model(v)
    torch.cuda.synchronize(dev)
    prof = ops.stop_timing()
    model.use_native = True
    gemm = prof.get("vrd_gemm", {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "n": 0})
    by_shape = {k.split(": ", 1)[1]: {"ms_per_step": round(v["ms"] / n_prof, 3), "launches_per_step": round(v["n"] / n_prof, 1),
                                       "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 1)}
                for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]) if ": " in k}
    log("[bench] gemm by shape: " + json.dumps(by_shape))
    prof = {k: v for k, v in prof.items() if ": " not in k}
    total_ms = sum(p["ms"] for p in prof.values()) or 1.0
    ach = gemm["flops"] / (gemm["ms"] * 1e-3) / 1e12 if gemm["ms"] > 0 else 0.0
    traffic, traffic_src = gemm_traffic(default_workload)
    hbm_kernels = {k[4:]: {"ms_per_step": round(v["ms"] / n_prof, 3), "gbs": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 0),
                           "frac": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9 / hbm, 3)}
                   for k, v in prof.items() if v.get("bytes", 0.0) > 0 and k != "vrd_gemm" and v["ms"] > 0}
    attn = prof.get("vrd_full_attn")
    whole_tflops = args.steps * world * sum(flops) / (ms * 1e-3) / 1e12 / world      # per GPU
    roofline = {"bound": "tensor", "kernel": "gemm_tcgen05_kernel", "achieved": ach, "peak": sustained, "unit": "TFLOP/s",
                "frac": ach / sustained, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": f"{src}: bf16 dense sustained {sustained} (burst {burst}) TFLOP/s, HBM copy {hbm} GB/s",
                "launches_per_step": gemm["n"] / n_prof, "avg_launch_us": 1e3 * gemm["ms"] / max(1, gemm["n"]),
                "share_of_step": gemm["ms"] / total_ms,
                "per_kernel_ms": {k[4:]: round(v["ms"] / n_prof, 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])},
                "hbm_kernels": hbm_kernels,
                "whole_path_tflops_per_gpu": whole_tflops, "whole_path_frac": whole_tflops / sustained}
    if attn is not None and attn["ms"] > 0:
        roofline["full_attention"] = {"ms_per_step": round(attn["ms"] / n_prof, 3), "tflops": round(attn["flops"] / (attn["ms"] * 1e-3) / 1e12, 1),
                                      "frac": round(attn["flops"] / (attn["ms"] * 1e-3) / 1e12 / sustained, 3)}

    # BASELINE config 5: fixed set of distinct videos, sharded by LPT over the ranks, results gathered on rank 0 -- all timed
    sweep = None
    if extras and args.sweep_videos > 0:
        sweep = run_sweep(args, cfg, model, dev, rank, world, sync_all, reduce_max)

    parity = cpu = None
    if not args.no_cpu_baseline and rank == 0 and world == 1:      # the CPU legs run at N = 1 only (torchrun pins OMP to one thread)
        torch.set_num_threads(threads)
        sd = default_state_dict(mc)
        sample = cpu_sample(host_videos, args.cpu_pairs, 12345)
        n, dt, outs = cpu_forward(cfg, sd, host_videos, sample, keep_outputs=True)
        cpu = {"value": n / dt, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{n} pairs drawn at random (seeded) from all pairs of the step's videos, oracle port of forward_test ({dt:.1f} s)"}
        parity = parity_block(model, cfg, dev_videos, outs)
        parity["precision"] = args.precision

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    out = {"metric": METRIC, "value": pairs / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": args.precision if args.precision == "bf16" else "f32", "data": "synthetic",
           "config": {"workload": workload, "videos_per_step": len(specs), "pairs_per_step": n_pairs, "long_pairs_per_step": long_pairs,
                      "max_pair_len": max(max(l) for l in lens_all), "valid_frames_per_step": int(sum(frames)),
                      "tflop_per_step": round(sum(flops) / 1e12, 2),
                      "l2": "inputs larger than L2 (pair features of one step: %.2f GB)" % (sum(in_bytes) / 1e9),
                      "weights": "random init (torch.manual_seed(0))", "parallelism": f"dp{world} by video",
                      "step_wall_ms": {"value": steps_value, "e2e": steps_e2e}},
           "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
           "valid_frames_per_s": world * args.steps * sum(frames) / (ms * 1e-3)}
    out.update(out_extra)
    if seeds_out is not None:
        out["seeds"] = seeds_out
    if sweep is not None:
        out["sweep"] = sweep
    if parity is not None:
        out["parity"] = parity
    if cpu is not None:
        out["cpu_baseline"] = cpu
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_sweep(args, cfg, model, dev, rank, world, sync_all, reduce_max):
    """BASELINE config 5 (scaled stand-in: ``--sweep-videos`` distinct videos for the 835 of VidOR validation, drawn as cfg2):
    strong scaling.  The videos are tracklet-level host inputs (the contract that scales: 11x fewer bytes than pair lists),
    sharded longest-processing-time-first on the FLOP model, every rank runs the pipelined loop on its shard
    (``runner.run_sharded``) and rank 0 gathers all results (``gather_object``); everything is inside the timed region."""
    import torch.distributed as dist
    dc = cfg["dataset_config"]
    st = dc.get("feat_stride", 1)
    specs = synth.cfg2_video_set(args.sweep_videos, args.set_seed + 7)
    # costs need the durations only (cheap); features are generated for the rank's own shard
    metas, costs, npairs = [], [], []
    for s, nf, nt in specs:
        _, _, durs, *_ = synth._tracklets(cfg, s, nt, nf, features=False, split_rng=True)
        pr = synth._overlapping_pairs(durs, st)
        lens = [len(range(0, min(durs[a][1], durs[b][1]) - max(durs[a][0], durs[b][0]), st)) for a, b in pr]
        costs.append(runner.video_cost(args.config, lens))
        npairs.append(len(lens))
    shards = runner.shard_videos(costs, world)
    videos = [None] * len(specs)
    for i in shards[rank]:
        s, nf, nt = specs[i]
        t = synth.synthetic_tracklet_video(cfg, s, n_tracklets=nt, n_frames=nf, split_rng=True)
        for key in ("visual_features_list", "clip_features_list", "bboxes_list"):
            if key in t:
                t[key] = [x.pin_memory() for x in t[key]]
        videos[i] = t
    for i in shards[rank][:2]:          # warm-up (buffers of this input kind)
        model.forward_tracklets(videos[i], dc)
    tw = time.perf_counter()
    runner.gather_results({})           # warm-up of the gather path (NCCL sets up its send/recv channels on first use)
    gather_warm_ms = 1e3 * (time.perf_counter() - tw)
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    res = runner.run_sharded(videos, costs, model=model, dataset_config=dc, gather=False)
    e1.record()
    torch.cuda.synchronize(dev)
    my_ms = e0.elapsed_time(e1)
    tg = time.perf_counter()
    merged = runner.gather_results(res)
    gather_ms = 1e3 * (time.perf_counter() - tg)
    wall_ms = 1e3 * (time.perf_counter() - t0)
    sync_all()
    ms_max = reduce_max(my_ms)
    ms_sum = my_ms
    if world > 1:
        t = torch.tensor([my_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms_sum = float(t)
    wall_max = reduce_max(wall_ms)
    if rank == 0:
        assert len(merged) == len(specs)
    n_trip = sum(len(r["triplets"]) for r in merged.values() if r is not None) if rank == 0 else 0
    return {"config": f"BASELINE configs[4] scaled: {len(specs)} distinct cfg2 videos (stand-in for the 835 of VidOR validation), tracklet-level "
                      f"host inputs, LPT shards over {world} rank(s), results gathered on rank 0",
            "scaling": "strong", "videos": len(specs), "pairs": int(sum(npairs)), "value": sum(npairs) / (ms_max * 1e-3), "unit": UNIT,
            "ms_device_max_over_ranks": ms_max, "imbalance_max_over_mean": ms_max / (ms_sum / world),
            "cost_imbalance_lpt": max(sum(costs[i] for i in sh) for sh in shards) / (sum(costs) / world),
            "gather_ms_rank0": gather_ms, "gather_first_use_ms_rank0": gather_warm_ms, "wall_ms_incl_gather_max_over_ranks": wall_max,
            "value_incl_gather": sum(npairs) / (wall_max * 1e-3), "triplets_gathered": n_trip}


if __name__ == "__main__":
    main()
