"""Throughput of the PUBLIC python call on real-shaped images: `HeadDetector.predict_batch` over pinned-free
numpy frames (720p by default) -> list of PredictionResult with materialised heads: device letterbox,
network, select/NMS, FLAME decode, D2H, result objects.  This is the drop-in path a user of the reference
switches to (bench.py times the BASELINE workload, whose inputs are already 640x640 uint8).

    python tools/bench_api.py [batch] [h] [w] [iters]"""
import json
import os
import sys
import time
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from head_detector_b200 import HeadDetector, arch, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
H = int(sys.argv[2]) if len(sys.argv) > 2 else 720
W = int(sys.argv[3]) if len(sys.argv) > 3 else 1280
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 10
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    det = HeadDetector(weights=arch.synthetic_weights(0), batch_size=B)
boxes, scores = synth.engineered_heads(B, det.model.A, 640, 8, seed=7)
det.model.set_override(boxes.cuda(), scores.cuda())
det.model.autotune(3)
rng = np.random.default_rng(0)
frames = [rng.integers(0, 256, (H, W, 3), dtype=np.uint8) for _ in range(B)]
stages = {}


def timed(label, fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    stages[label] = stages.get(label, 0.0) + time.perf_counter() - t0
    return r


for it in range(iters + 2):
    if it == 2:
        stages.clear()
        torch.cuda.synchronize()
        t_all = time.perf_counter()
    batch, xf, caches = timed("letterbox (H2D of raw frames + kernel)", lambda: det._prepare_batch(frames))
    out = timed("network + select/NMS + FLAME decode", lambda: det.detect_batch(batch, 0.5, xf))
    heads = timed("D2H + result objects", lambda: det._parse_batch(out, caches))
torch.cuda.synchronize()
dt = (time.perf_counter() - t_all) / iters
n_heads = sum(len(h) for h in heads)
print(json.dumps({"api": "HeadDetector.predict_batch stages", "batch": B, "frame": [H, W], "heads_per_batch": n_heads, "ms_per_batch": dt * 1e3,
                  "images_per_s": B / dt, "stage_ms": {k: v / iters * 1e3 for k, v in stages.items()},
                  "h2d_bytes_per_batch": B * H * W * 3, "d2h_bytes_per_batch": n_heads * (5023 * 12 + 413 * 4 + 36)}))
