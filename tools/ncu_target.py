"""Target for ncu: builds the engine (B=32, 640), optionally autotunes, warms up, then runs ONE eager
step of the whole path between cudaProfilerStart/Stop (use `ncu --profile-from-start off`)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from head_detector_b200 import arch, synth  # noqa: E402
from head_detector_b200.engine import Engine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
tune = len(sys.argv) > 2 and sys.argv[2] == "tuned"
eng = Engine(arch.synthetic_weights(0), B, 640)  # VGGHEADS_B200_SPARSE_HEADS=1 selects the two-phase (sparse heads) plan
img = synth.synthetic_images(B, 640, 0).cuda()
boxes, scores = synth.engineered_heads(B, eng.A, 640, 8, seed=7)
eng.set_override(boxes.cuda(), scores.cuda())
if tune:
    eng.autotune(5)
for _ in range(2):
    eng.forward(img)
    eng.postprocess(0.5, 0.5, 1000)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.forward(img)
eng.postprocess(0.5, 0.5, 1000)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("heads", int(eng.head_offsets[-1]))
for i, op in enumerate(eng.plan.ops):
    print(i, op.label, eng.op_config(i) if op.kind == 1 else "")
