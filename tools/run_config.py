"""Runs one BASELINE config shape through the full device path (graph replay) and prints images/s.
usage: python tools/run_config.py B S heads_per_image [steps]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from head_detector_b200 import arch, synth  # noqa: E402
from head_detector_b200.engine import Engine  # noqa: E402

B, S, heads = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 10
eng = Engine(arch.synthetic_weights(0), B, S, sparse_heads=True)
eng.input.copy_(synth.synthetic_images(B, S, 0).cuda())
boxes, scores = synth.engineered_heads(B, eng.A, S, heads, per_cluster=12 if S == 640 else 40, seed=7)
eng.set_override(boxes.cuda(), scores.cuda())
eng.autotune(5)
for _ in range(3):
    eng.run_device()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    eng.run_device()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(json.dumps({"batch": B, "image_size": S, "anchors": eng.A, "heads_total": int(eng.head_offsets[-1]), "ms_per_step": ms,
                  "images_per_s": B / ms * 1e3, "launches": eng.launch_count}))
