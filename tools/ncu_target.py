"""Small target for ncu: one warm-up + one eager step of the whole path (B=32, 640)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from head_detector_b200 import arch, synth  # noqa: E402
from head_detector_b200.engine import Engine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
eng = Engine(arch.synthetic_weights(0), B, 640)
img = synth.synthetic_images(B, 640, 0).cuda()
boxes, scores = synth.engineered_heads(B, eng.A, 640, 8, seed=7)
eng.set_override(boxes.cuda(), scores.cuda())
for _ in range(2):
    eng.forward(img)
    eng.postprocess(0.5, 0.5, 1000)
torch.cuda.synchronize()
print("heads", int(eng.head_offsets[-1]))
