#!/usr/bin/env python
"""Benchmark of the MaskVRD inference hot path: relation pairs/sec (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config vidor] [--precision bf16|fp32]

A "step" is one pass of the hot path (``model(input_data)`` = network + heads epilogue + triplet decoding) over one
synthetic VidOR-shaped video (tens of tracklets, every ordered pair with temporal overlap,
feat_stride 4).  Prints ONE JSON line (rank 0).  For N > 1 launch with torchrun; every rank processes its own videos
(weak scaling, no collective on the data path) and the time is the max over ranks.

  value        whole-job pairs/s with the pair features already resident in HBM (CUDA events around the K steps)
  e2e          the same through the public API with HOST (pinned) pair features: H2D copies + D2H of results inside
  roofline     the tcgen05 GEMM kernel: algorithmic FLOPs of its launches / their CUDA-event durations vs the measured
               bf16 peak in MEASURED_PEAKS.json (second pass over the same steps with per-launch events)
  cpu_baseline the oracle port of the reference forward on the host cores, on a bounded sample of the same workload
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from vrdone_b200 import synth, runner  # noqa: E402

METRIC = "relation pairs/sec, VrdONE forward"
UNIT = "pairs/s"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--config", default="vidor", choices=list(synth.CONFIG_NAMES))
    p.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    p.add_argument("--videos", type=int, default=2, help="distinct synthetic videos cycled through the steps")
    p.add_argument("--tracklets", type=int, default=40)
    p.add_argument("--frames", type=int, default=1200)
    p.add_argument("--cpu-pairs", type=int, default=48, help="pairs in the bounded CPU-baseline sample")
    p.add_argument("--no-cpu-baseline", action="store_true")
    return p.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d.get("bf16_tflops_sustained", 1400.0), d.get("bf16_tflops", 1590.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1400.0, 1590.0, 6650.0, "fallback"


def gemm_traffic(args):
    """DRAM bytes per GEMM launch (dram__bytes_read.sum + dram__bytes_write.sum, averaged over the 118 launches of one forward)
    from the committed ncu capture of the DEFAULT workload; None for any other workload or precision."""
    path = os.path.join(ROOT, "profiles", "r1_gemm_traffic.json")
    default = args.config == "vidor" and args.precision == "bf16" and args.tracklets == 40 and args.frames == 1200
    if not (default and os.path.exists(path)):
        return None
    with open(path) as f:
        return json.load(f)["traffic_bytes_per_launch"]


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


def make_videos(cfg, args, rank):
    vids = []
    for v in range(args.videos):
        # weak scaling: every rank works on its own copy of the SAME synthetic videos, so per-GPU work is identical for all N
        vids.append(synth.synthetic_video(cfg, v, n_tracklets=args.tracklets, n_frames=args.frames))
    return vids


def to_device(video, dev):
    return {k: ([t.to(dev) for t in v] if isinstance(v, list) else (v.to(dev) if torch.is_tensor(v) else v)) for k, v in video.items()}


def pin(video):
    """Pinned host copy of the pair features, keeping the data loader's layout: a (C, L) view of an (L, C)-contiguous buffer
    (reference vidor.py:708-711; DataLoader(pin_memory=True) preserves strides)."""
    return {k: ([t.t().contiguous().pin_memory().t() if k == "so_features_list" else t for t in v] if isinstance(v, list) else v)
            for k, v in video.items()}


def pin_arena(video):
    """The same pair features pinned in ONE arena per video, pairs back to back in token-major order (what a loader that
    allocates a video's pinned memory once would hand over): back-to-back pairs are staged with one copy per chunk."""
    feats = video["so_features_list"]
    total = sum(f.numel() for f in feats)
    arena = torch.empty(total, dtype=torch.float32, pin_memory=True)
    views, pos = [], 0
    for f in feats:
        C, L = f.shape
        v = arena[pos:pos + L * C].view(L, C)
        v.copy_(f.t())
        views.append(v.t())
        pos += L * C
    out = dict(video)
    out["so_features_list"] = views
    return out


def cpu_baseline(cfg, video, n_pairs, threads):
    """Oracle port of the reference forward on the host cores, on the first ``n_pairs`` pairs of a video."""
    from oracle import maskvrd_oracle as O
    from vrdone_b200 import MaskVRD
    torch.set_num_threads(threads)
    mc = cfg["model_config"]
    torch.manual_seed(0)
    sd = {k: v.detach().clone() for k, v in MaskVRD(mc, "cpu").state_dict().items()}
    n = min(n_pairs, len(video["so_features_list"]))
    sub = dict(video)
    sub["so_features_list"] = video["so_features_list"][:n]
    sub["sids"], sub["oids"], sub["so_offset"] = video["sids"][:n], video["oids"][:n], video["so_offset"][:n]
    t0 = time.perf_counter()
    O.forward_test(sub, sd, mc, cfg["inference_config"])
    dt = time.perf_counter() - t0
    return n / dt, n, dt


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = synth.load_config(args.config)
    threads = os.cpu_count() or 1
    workload = (f"{args.config}.yaml forward_test, synthetic VidOR-shaped videos: {args.tracklets} tracklets x {args.frames} frames, "
                f"feat_stride {cfg['dataset_config']['feat_stride']}, all ordered overlapping pairs")

    if args.impl == "reference":
        if rank != 0:
            return
        video = synth.synthetic_video(cfg, 0, n_tracklets=args.tracklets, n_frames=args.frames)
        n = max(8, args.cpu_pairs // 2)
        for _ in range(max(0, min(args.warmup, 1))):
            cpu_baseline(cfg, video, n, threads)
        tot_pairs = tot_t = 0.0
        for _ in range(args.steps):
            v, npairs, dt = cpu_baseline(cfg, video, n, threads)
            tot_pairs += npairs
            tot_t += dt
        val = tot_pairs / tot_t
        sample = f"first {n} pairs of synthetic video seed 0 per step (reference padding: short pairs to max_seq_len)"
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
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    model = MaskVRD(cfg["model_config"], dev).eval().to(dev)
    model._config_eval(cfg["inference_config"])
    model.set_precision(args.precision)
    host_videos = make_videos(cfg, args, rank)
    pinned = [pin(v) for v in host_videos]
    dev_videos = [to_device(v, dev) for v in host_videos]
    n_pairs = [len(v["sids"]) for v in host_videos]
    flops = [runner.video_cost(args.config, [int(f.shape[1]) for f in v["so_features_list"]]) for v in host_videos]
    frames = [sum(int(f.shape[1]) for f in v["so_features_list"]) for v in host_videos]
    in_bytes = [sum(f.numel() * 4 for f in v["so_features_list"]) for v in host_videos]

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    gc_ms = [0.0]
    gc_t0 = [0.0]

    gc_gen = {}

    def gc_cb(phase, info):
        if phase == "start":
            gc_t0[0] = time.perf_counter()
        else:
            d = 1e3 * (time.perf_counter() - gc_t0[0])
            gc_ms[0] += d
            g = gc_gen.setdefault(info["generation"], [0, 0.0])
            g[0] += 1
            g[1] += d
    import gc
    gc.callbacks.append(gc_cb)
    step_wall = []
    free_ms = []

    def timed(videos, steps, h2d, pipelined=True, dataset_config=None):
        """K steps between two CUDA events.  ``pipelined``: through ``runner.run_videos`` (two videos in flight: the decode of
        one overlaps the kernels of the next; every result is produced and dropped inside the timed region); otherwise one
        synchronous ``model(v)`` call after the other, as the reference's eval loop does."""
        sync_all()
        gc_ms[0] = 0.0
        gc_gen.clear()
        step_wall.clear()
        free_ms.clear()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pairs = 0
        if pipelined:
            tw = time.perf_counter()
            for out in runner.run_videos(model, (videos[s % len(videos)] for s in range(steps)), dataset_config=dataset_config):
                del out     # dropping the ~2*10^5 Python objects of the result is part of the step
                step_wall.append(round(1e3 * (time.perf_counter() - tw), 1))
                tw = time.perf_counter()
            pairs = sum(n_pairs[s % len(videos)] for s in range(steps))
        else:
            for s in range(steps):
                v = videos[s % len(videos)]
                tw = time.perf_counter()
                # h2d: ``v`` holds pinned HOST pair features; the module moves them (inside the timed region)
                out = model(v) if dataset_config is None else model.forward_tracklets(v, dataset_config)
                tm = time.perf_counter()
                del out
                step_wall.append(round(1e3 * (time.perf_counter() - tw), 1))
                free_ms.append(round(1e3 * (time.perf_counter() - tm), 1))
                pairs += n_pairs[s % len(videos)]
        e1.record()
        sync_all()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            p = torch.tensor([float(pairs)], device=dev)
            dist.all_reduce(p, op=dist.ReduceOp.SUM)
            ms, pairs = float(t), float(p)
        return ms, pairs

    def warm(videos, dataset_config=None):
        """W untimed steps through BOTH loops: the pipelined one keeps two videos' pinned read-back / layout buffers alive at
        once, and the first cudaHostAlloc of those costs ~100 ms."""
        for s in range(args.warmup):
            v = videos[s % len(videos)]
            model(v) if dataset_config is None else model.forward_tracklets(v, dataset_config)
        for out in runner.run_videos(model, (videos[s % len(videos)] for s in range(max(args.warmup, 3))), dataset_config=dataset_config):
            del out

    warm(dev_videos)
    ops = model._ops
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = ops.launches
    ms, pairs = timed(dev_videos, args.steps, h2d=False)
    launches = ops.launches - l0
    wall_pipe = list(step_wall)
    host_pipe = {k: round(v, 2) for k, v in model.last_stats.items()}
    ms_sync, pairs_sync = timed(dev_videos, args.steps, h2d=False, pipelined=False)
    host_hbm = {k: round(v, 2) for k, v in model.last_stats.items()}
    host_hbm["python_gc_ms_per_step"] = round(gc_ms[0] / args.steps, 2)
    host_hbm["python_gc_by_generation"] = {str(k): [v[0], round(v[1], 1)] for k, v in gc_gen.items()}
    host_hbm["forward_wall_ms_each_step"] = list(step_wall)
    host_hbm["result_free_ms_each_step"] = list(free_ms)
    clocks = sampler.stop()
    warm(pinned)                        # staging buffers, pinned upload blocks and the copy stream are created on first use
    ms_e2e, pairs_e2e = timed(pinned, args.steps, h2d=True)
    wall_pipe_e2e = list(step_wall)
    host_pipe_e2e = {k: round(v, 2) for k, v in model.last_stats.items()}
    ms_e2e_sync, pairs_e2e_sync = timed(pinned, args.steps, h2d=True, pipelined=False)
    arena = [pin_arena(v) for v in host_videos]
    warm(arena)
    ms_arena, pairs_arena = timed(arena, args.steps, h2d=True)
    del arena
    host_e2e = {k: round(v, 2) for k, v in model.last_stats.items()}
    host_e2e["forward_wall_ms_each_step"] = list(step_wall)

    # the same videos through the tracklet-level entry point (per-tracklet features cross PCIe once; the pair
    # gather and the box geometry run on the device).  Reported beside the headline numbers, never instead of them.
    trk_videos = []
    for v in range(args.videos):
        t = synth.synthetic_tracklet_video(cfg, v, n_tracklets=args.tracklets, n_frames=args.frames)
        for key in ("visual_features_list", "clip_features_list", "bboxes_list"):
            if key in t:
                t[key] = [x.pin_memory() for x in t[key]]
        trk_videos.append(t)
    trk_bytes = [sum(x.numel() * 4 for key in ("visual_features_list", "clip_features_list", "bboxes_list") if key in t for x in t[key])
                 for t in trk_videos]
    warm(trk_videos, cfg["dataset_config"])
    ms_trk, pairs_trk = timed(trk_videos, args.steps, h2d=True, dataset_config=cfg["dataset_config"])
    ms_trk_sync, _ = timed(trk_videos, args.steps, h2d=True, pipelined=False, dataset_config=cfg["dataset_config"])
    # the same loops with so_trajs returned as a lazy sequence (no eager per-frame box lists)
    model.lazy_trajs = True
    ms_lazy, pairs_lazy = timed(dev_videos, args.steps, h2d=False)
    ms_lazy_e2e, _ = timed(pinned, args.steps, h2d=True)
    model.lazy_trajs = False

    # network only (the `_mask_vrd`-equivalent part, no triplet decode): kernels of all steps back to back
    from vrdone_b200.layout import reference_padded_lengths
    net_in = []
    for v in dev_videos:
        lens_v = [int(f.shape[1]) for f in v["so_features_list"]]
        net_in.append((v["so_features_list"], reference_padded_lengths(lens_v, cfg["model_config"])))
    for feats_v, tp_v in net_in:
        model.run_network(feats_v, tp_v, model.topk)
    sync_all()
    net_sampler = ClockSampler(local_rank)       # this loop keeps the GPU at 100 % duty: power capping shows up here first
    net_sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(args.steps):
        feats_v, tp_v = net_in[s % len(net_in)]
        model.run_network(feats_v, tp_v, model.topk)
    e1.record()
    sync_all()
    net_clocks = net_sampler.stop()
    ms_net = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_net], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_net = float(t)
    pairs_net = world * sum(n_pairs[s % len(n_pairs)] for s in range(args.steps))
    frames_net = world * sum(frames[s % len(frames)] for s in range(args.steps))

    # roofline pass: the same steps with a CUDA-event pair around every launch of the dominant (GEMM) kernel
    model.use_native = False        # same kernels, same order, issued one by one from Python so that each launch can be timed
    ops.start_timing()
    for s in range(args.steps):
        model(dev_videos[s % len(dev_videos)])
    torch.cuda.synchronize(dev)
    prof = ops.stop_timing()
    model.use_native = True
    sustained, burst, hbm, src = peaks()
    gemm = prof.get("vrd_gemm", {"ms": 0.0, "flops": 0.0, "n": 0})
    by_shape = {k.split(": ", 1)[1]: {"ms_per_step": round(v["ms"] / args.steps, 3), "launches_per_step": v["n"] / args.steps,
                                       "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 1)}
                for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]) if ": " in k}
    prof = {k: v for k, v in prof.items() if ": " not in k}
    total_ms = sum(p["ms"] for p in prof.values()) or 1.0
    ach = gemm["flops"] / (gemm["ms"] * 1e-3) / 1e12 if gemm["ms"] > 0 else 0.0
    peak = sustained if args.precision == "bf16" else 75.0
    roofline = {"bound": "tensor", "kernel": "gemm_tcgen05_kernel" if args.precision == "bf16" else "gemm_simt_kernel",
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": gemm_traffic(args),
                "peak_source": f"{src} bf16 dense, sustained (burst {burst})" if args.precision == "bf16" else "nominal fp32 CUDA-core",
                "launches": gemm["n"], "avg_launch_us": 1e3 * gemm["ms"] / max(1, gemm["n"]),
                "share_of_step": gemm["ms"] / total_ms,
                "per_kernel_ms": {k: round(v["ms"] / args.steps, 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])},
                "gemm_by_shape": by_shape,
                "whole_path_algorithmic_tflops": sum(flops[s % len(flops)] for s in range(args.steps)) / (ms * 1e-3) / 1e12}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    out = {"metric": METRIC, "value": pairs / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": args.precision if args.precision == "bf16" else "f32", "data": "synthetic",
           "config": {"workload": workload, "pairs_per_step": n_pairs, "valid_frames_per_step": frames,
                      "l2": "inputs larger than L2 (pair features of one video: %.2f GB)" % (in_bytes[0] / 1e9),
                      "weights": "random init (torch.manual_seed(0))", "parallelism": f"dp{world} by video"},
           "e2e": {"value": pairs_e2e / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(sum(in_bytes) / len(in_bytes)),
                   "d2h_bytes_per_step": int(sum(n * cfg["model_config"]["predictor"]["num_queries"] * (8 * cfg["inference_config"]["topk"] + 8)
                                                 for n in n_pairs) / len(n_pairs))},
           "e2e_arena": {"value": pairs_arena / (ms_arena * 1e-3), "unit": UNIT,
                         "note": "as e2e, but every video's pair features pinned in ONE arena (pairs back to back): one copy per chunk"},
           "e2e_tracklet_api": {"value": pairs_trk / (ms_trk * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(sum(trk_bytes) / len(trk_bytes)),
                                "note": "MaskVRD.forward_tracklets: host tracklet features in, triplets out (SURVEY 8f row 1)"},
           "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
           "api": "runner.run_videos(model, videos): submit video i+1, then wait for + decode video i (two videos in flight)",
           "sync_call": {"note": "one blocking model(input_data) call after the other (the reference's eval loop), no overlap of "
                                 "host decode and device work",
                         "value": pairs_sync / (ms_sync * 1e-3), "e2e": pairs_e2e_sync / (ms_e2e_sync * 1e-3),
                         "e2e_tracklet_api": pairs_trk / (ms_trk_sync * 1e-3), "unit": UNIT},
           "lazy_trajs": {"note": "model.lazy_trajs = True: so_trajs is a LazyTrajs sequence (SURVEY 8f row 3); same pipelined loops",
                          "value": pairs_lazy / (ms_lazy * 1e-3), "e2e": pairs_lazy / (ms_lazy_e2e * 1e-3), "unit": UNIT},
           "network_only": {"note": "run_network (everything up to the compact per-(pair, query) records, no host decode), HBM-resident",
                            "value": pairs_net / (ms_net * 1e-3), "unit": UNIT, "valid_frames_per_s": frames_net / (ms_net * 1e-3),
                            "ms_per_step": ms_net / args.steps, "clocks": net_clocks},
           "valid_frames_per_s": world * sum(frames[s % len(frames)] for s in range(args.steps)) / (ms * 1e-3),
           "step_wall_ms_pipelined": {"hbm_resident": wall_pipe, "e2e": wall_pipe_e2e},
           "host_ms_last_step_pipelined": {"hbm_resident": host_pipe, "e2e": host_pipe_e2e},
           "host_ms_last_step": host_hbm, "host_ms_last_step_e2e": host_e2e}
    if not args.no_cpu_baseline:
        v, n, dt = cpu_baseline(cfg, host_videos[0], args.cpu_pairs, threads)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                               "sample": f"first {n} pairs of video 0 through the oracle port of forward_test ({dt:.1f} s)"}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
