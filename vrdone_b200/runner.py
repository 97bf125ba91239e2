"""Multi-GPU driver for the hot path: videos (and their pairs) are independent, so work is sharded host-side with
no collective on the data path.  One process per GPU; results are Python objects gathered on rank 0.

The reference runs inference single-process on cuda:0 (eval.py:83, 140-152); this is the 1..8 GPU version of that loop.
"""
from __future__ import annotations

from collections import deque
from typing import Callable, Dict, Iterable, Iterator, List, Optional, Sequence

import torch
import torch.distributed as dist

# algorithmic FLOP model of one pair with valid length L: a0*L + b0*L^2 + a1*ceil(L/2) + a2*ceil(L/4) + a3*ceil(L/8) + c
FLOP_COEFF = {
    "vidvrd": (58.174e6, 16384.0, 6.579e6, 6.579e6, 7.678e6, 70.56e6),
    "vidor": (58.190e6, 16384.0, 6.583e6, 6.583e6, 7.682e6, 70.18e6),
    "vidor_local": (58.338e6, 0.0, 6.583e6, 6.583e6, 7.682e6, 70.18e6),
    "vidor_x": (67.628e6, 16384.0, 6.583e6, 6.583e6, 7.686e6, 78.02e6),
}


def pair_flops(config_name: str, length: int) -> float:
    a0, b0, a1, a2, a3, c = FLOP_COEFF[config_name]
    L = int(length)
    return a0 * L + b0 * L * L + a1 * ((L + 1) // 2) + a2 * ((L + 3) // 4) + a3 * ((L + 7) // 8) + c


def video_cost(config_name: str, lengths: Sequence[int]) -> float:
    return float(sum(pair_flops(config_name, l) for l in lengths))


def shard_videos(costs: Sequence[float], world_size: int) -> List[List[int]]:
    """Longest-processing-time-first partition of video indices over ranks (deterministic)."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads = [0.0] * world_size
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda j: (loads[j], j))
        shards[r].append(i)
        loads[r] += costs[i]
    return [sorted(s) for s in shards]


def run_videos(model, videos: Iterable[dict], depth: int = 2, dataset_config: Optional[dict] = None) -> Iterator[object]:
    """The reference's eval loop ``for proposal in loader: out = model(proposal)`` (eval.py:140-152) with ``depth`` videos in
    flight: video i + 1 is enqueued (copies, kernels, asynchronous read-back) before video i is waited for and decoded, so the
    host-side decode of one video overlaps the device work of the next.  Yields exactly what ``model(video)`` returns, in
    input order.  With ``dataset_config`` the videos are tracklet-level inputs (``MaskVRD.submit_tracklets``)."""
    pending = deque()
    for v in videos:
        pending.append(model.submit(v) if dataset_config is None else model.submit_tracklets(v, dataset_config))
        if len(pending) >= depth:
            yield pending.popleft().result()
    while pending:
        yield pending.popleft().result()


def _rank_world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def run_sharded(videos: Sequence[dict], costs: Sequence[float], fn: Optional[Callable[[dict], object]] = None, model=None,
                dataset_config: Optional[dict] = None, gather: bool = True, lazy_transport: bool = True) -> Dict[int, object]:
    """Every rank processes its shard -- with ``fn(video)`` one video after the other, or with ``model`` through the pipelined
    ``run_videos`` loop (two videos in flight per GPU); rank 0 returns {video index: result} for all videos (other ranks:
    their own; every rank with ``gather=False``, to be merged later with ``gather_results``).  Only the entries of a rank's
    own shard of ``videos`` are touched.  Works with or without an initialised process group (world size 1).
    ``lazy_transport`` (multi-rank runs with ``model``): ``so_trajs`` travels as a ``LazyTrajs`` of float32 box arrays instead
    of nested Python lists -- pickling 10^7 Python floats per rank cost more than the forward itself (measured: 1.4 s for 24
    videos) -- and compares / indexes like the reference's lists on arrival (``.materialise()`` for the lists themselves)."""
    assert (fn is None) != (model is None), "give either fn or model"
    rank, world = _rank_world()
    mine = shard_videos(costs, world)[rank]
    if model is not None:
        was_lazy = getattr(model, "lazy_trajs", False)
        if lazy_transport and world > 1:
            model.lazy_trajs = True
        try:
            local = dict(zip(mine, run_videos(model, (videos[i] for i in mine), dataset_config=dataset_config)))
        finally:
            model.lazy_trajs = was_lazy
    else:
        local = {i: fn(videos[i]) for i in mine}
    return gather_results(local) if gather else local


def gather_results(local: Dict[int, object]) -> Dict[int, object]:
    """Host-side gather of the per-rank result dicts on rank 0 (``gather_object``: the results are Python lists; there is no
    collective on the data path).  Other ranks get their own dict back."""
    rank, world = _rank_world()
    if world == 1:
        return local
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(local if rank != 0 else None, gathered, dst=0)      # rank 0's own results need no pickling round trip
    if rank != 0:
        return local
    merged: Dict[int, object] = dict(local)
    for part in gathered[1:]:
        merged.update(part)
    return merged
