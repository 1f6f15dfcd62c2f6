"""Experiment: device throughput with one batch in flight vs two (two engines on two streams, so that
the select/NMS/FLAME tail of batch i overlaps the stem/stage-1 kernels of batch i+1)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from head_detector_b200 import arch, synth  # noqa: E402
from head_detector_b200.engine import Engine  # noqa: E402

B, S, steps = 32, 640, int(sys.argv[1]) if len(sys.argv) > 1 else 40
w = arch.synthetic_weights(0)
engs = [Engine(w, B, S) for _ in range(3)]
boxes, scores = synth.engineered_heads(B, engs[0].A, S, 8, seed=7)
bd, sd = boxes.cuda(), scores.cuda()
imgs = [synth.synthetic_images(B, S, i).cuda() for i in range(4)]
for e in engs:
    e.set_override(bd, sd)
    e.autotune(5)
streams = [torch.cuda.Stream() for _ in engs]


def run(n_eng):
    for k in range(4):
        with torch.cuda.stream(streams[k % n_eng]):
            engs[k % n_eng].input.copy_(imgs[k % 4], non_blocking=True)
            engs[k % n_eng].run_device()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams:
        s.wait_event(e0)
    for k in range(steps):
        with torch.cuda.stream(streams[k % n_eng]):
            engs[k % n_eng].input.copy_(imgs[k % 4], non_blocking=True)
            engs[k % n_eng].run_device()
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return ms, B / ms * 1e3


for n in (1, 2, 3, 1, 2, 3):
    ms, ips = run(n)
    print(f"{n} batch(es) in flight: {ms:.3f} ms/step, {ips:.0f} img/s", flush=True)
