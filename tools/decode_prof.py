"""Times the host-side decode stages (pure CPU) to find where the time goes on a given box."""
import gc
import time

import numpy as np
import torch

from vrdone_b200 import MaskVRD, synth

cfg = synth.load_config("vidor")
model = MaskVRD(cfg["model_config"], "cpu").eval()
model._config_eval(cfg["inference_config"])
video = synth.synthetic_video(cfg, 0, n_tracklets=40, n_frames=1200)
B, Q, k = len(video["sids"]), 9, 6
lens = np.array([int(f.shape[1]) for f in video["so_features_list"]])
g = np.random.default_rng(0)
scores = g.random((B, Q, k), dtype=np.float32)
cats = g.integers(1, 51, (B, Q, k)).astype(np.int32)
first = (g.random((B, Q)) * lens[:, None] * 0.3).astype(np.int32)
last = np.minimum(lens[:, None] - 1, first + (g.random((B, Q)) * lens[:, None] * 0.7).astype(np.int32))
fl = np.stack([first, last], -1).astype(np.int32)
print("torch threads", torch.get_num_threads(), "pairs", B)
for label, dis in (("gc on", False), ("gc off", True)):
    if dis:
        gc.disable()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        model._decode(scores, cats, fl, video)
        ts.append(1e3 * (time.perf_counter() - t0))
    print(label, "decode ms:", [round(t, 1) for t in ts])
gc.enable()
# stage timers
self = model
t = [time.perf_counter()]
for rep in range(3):
    t = [time.perf_counter()]
    x = np.sort(np.random.default_rng(1).random(70000).astype(np.float32)); t.append(time.perf_counter())
    y = np.partition(x, 69800); t.append(time.perf_counter())
    z = [video["bboxes_list"][i % 40].numpy()[10:200].tolist() for i in range(400)]; t.append(time.perf_counter())
    w = [float(v) for v in x[:20000]]; t.append(time.perf_counter())
    print("sort70k, partition, 400 tolist, 20k float():", [round(1e3 * (b - a), 2) for a, b in zip(t[:-1], t[1:])])
