"""N > 1 host logic on CPU: LPT sharding + gather over a world_size-2 gloo group gives the same results as one rank."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vrdone_b200 import runner


def _fake_forward(video):
    # deterministic stand-in for model(video): depends only on the video's own content
    return {"name": video["video_name"], "n": len(video["lens"]), "sum": int(sum(video["lens"]))}


def _videos():
    g = torch.Generator().manual_seed(0)
    vids = []
    for i in range(11):
        n = int(torch.randint(1, 40, (1,), generator=g))
        vids.append({"video_name": f"v{i}", "lens": torch.randint(2, 700, (n,), generator=g).tolist()})
    return vids


class _FakeModel:
    """Stand-in with the submit / result protocol of MaskVRD (the pipelined per-rank loop)."""

    class _Pending:
        def __init__(self, video):
            self.video = video

        def result(self):
            return _fake_forward(self.video)

    lazy_trajs = False

    class _PendingTrajs(_Pending):
        """A result that carries ``so_trajs`` the way MaskVRD does: nested lists, or LazyTrajs when ``lazy_trajs`` is on."""

        def __init__(self, video, lazy):
            super().__init__(video)
            self.lazy = lazy

        def result(self):
            import numpy as np
            from vrdone_b200.maskvrd import LazyTrajs
            out = _fake_forward(self.video)
            boxes = np.arange(4 * 12, dtype=np.float32).reshape(12, 4) + len(self.video["lens"])
            views = [(boxes[0:3], boxes[2:5]), (boxes[4:12], boxes[0:8])]
            out["so_trajs"] = LazyTrajs(views) if self.lazy else [[a.tolist(), b.tolist()] for a, b in views]
            return out

    def __init__(self, with_trajs=False):
        self.with_trajs = with_trajs

    def submit(self, video):
        return self._PendingTrajs(video, self.lazy_trajs) if self.with_trajs else self._Pending(video)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    vids = _videos()
    costs = [runner.video_cost("vidor", v["lens"]) for v in vids]
    out = runner.run_sharded(vids, costs, _fake_forward)
    piped = runner.run_sharded(vids, costs, model=_FakeModel())
    m = _FakeModel(with_trajs=True)
    lazy = runner.run_sharded(vids, costs, model=m)                # trajectories travel as LazyTrajs (float32 arrays)
    assert m.lazy_trajs is False                                   # ... and the model's own setting is restored
    if rank == 0:
        from vrdone_b200.maskvrd import LazyTrajs
        assert all(isinstance(r["so_trajs"], LazyTrajs) for r in lazy.values())
        lazy = {i: dict(r, so_trajs=r["so_trajs"].materialise()) for i, r in lazy.items()}
        q.put((out, piped, lazy))
    dist.barrier()
    dist.destroy_process_group()


def test_lpt_partition_is_balanced_and_complete():
    costs = [runner.video_cost("vidor", v["lens"]) for v in _videos()]
    for world in (1, 2, 4, 8):
        shards = runner.shard_videos(costs, world)
        assert sorted(i for s in shards for i in s) == list(range(len(costs)))
        loads = [sum(costs[i] for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(costs)
    assert abs(runner.pair_flops("vidor", 512) - 37.18e9) / 37.18e9 < 2e-3      # SURVEY.md section 8d examples
    assert abs(runner.pair_flops("vidvrd", 96) - 6.37e9) / 6.37e9 < 2e-3


def test_two_rank_gloo_matches_single_rank():
    vids = _videos()
    costs = [runner.video_cost("vidor", v["lens"]) for v in vids]
    single = runner.run_sharded(vids, costs, _fake_forward)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged, piped, lazy = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert merged == single and piped == single
    assert lazy == runner.run_sharded(vids, costs, model=_FakeModel(with_trajs=True))      # one rank: plain nested lists
    assert runner.run_sharded(vids, costs, model=_FakeModel()) == single


def test_run_videos_keeps_order_and_depth():
    """Host logic of the pipelined eval loop: results come back in input order, at most ``depth`` videos are in flight, and a
    video is waited for only after the next one has been submitted."""
    from vrdone_b200 import runner
    log = []

    class Pending:
        def __init__(self, v):
            self.v = v

        def result(self):
            log.append(("result", self.v))
            return self.v * 10

    class Model:
        def submit(self, v):
            log.append(("submit", v))
            return Pending(v)

        def submit_tracklets(self, v, dc):
            log.append(("submit_trk", v))
            return Pending(v)

    assert list(runner.run_videos(Model(), range(4), depth=2)) == [0, 10, 20, 30]
    assert log[:4] == [("submit", 0), ("submit", 1), ("result", 0), ("submit", 2)]
    in_flight = peak = 0
    for what, _ in log:
        in_flight += 1 if what == "submit" else -1
        peak = max(peak, in_flight)
    assert peak == 2
    log.clear()
    assert list(runner.run_videos(Model(), range(3), depth=1)) == [0, 10, 20]
    assert [w for w, _ in log] == ["submit", "result"] * 3
    assert list(runner.run_videos(Model(), [5], depth=2, dataset_config={})) == [50]
    assert list(runner.run_videos(Model(), [], depth=2)) == []
