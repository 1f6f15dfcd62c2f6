"""autotune + per-op profile -> gpurun_out/ops_<tag>.txt (with the chosen configs)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from head_detector_b200 import _lib, arch, synth
from head_detector_b200.engine import Engine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
tag = sys.argv[2] if len(sys.argv) > 2 else "tuned"
tune = (sys.argv[3] != "0") if len(sys.argv) > 3 else True
eng = Engine(arch.synthetic_weights(0), B, 640)
eng.input.copy_(synth.synthetic_images(B, 640, 0).cuda())
boxes, scores = synth.engineered_heads(B, eng.A, 640, 8, seed=7)
eng.set_override(boxes.cuda(), scores.cuda())
if tune:
    eng.autotune(5)
eng.profile(iters=2)
rows = eng.profile(iters=10)
tot = sum(t for _, t, _ in rows); conv_ms = sum(t for _, t, f in rows if f); conv_fl = sum(f for _, t, f in rows if f)
with open(os.path.join(ROOT, "gpurun_out", f"ops_{tag}.txt"), "w") as f:
    def log(s):
        print(s); f.write(s + "\n")
    log(f"B={B} total {tot:.3f} ms/step ({B / tot * 1e3:.0f} img/s eager), conv {conv_ms:.3f} ms = {conv_fl / conv_ms / 1e9:.1f} TFLOP/s")
    for i, ((label, t, fl), op) in enumerate(zip(rows, list(eng.plan.ops) + [None] * 4)):
        extra = ""
        if op is not None and op.kind == _lib.OP_CONV:
            c = eng.op_config(i)
            extra = f"k{op.k} s{op.stride} cin{op.src[2]:5d} cout{op.cout:5d} in{eng.plan.bufs[op.src[0]][0]:4d} mt{c['mt']} st{c['stages']} bn{c['block_n']} bk{c['bk']} tile{c['tw']}x{c['th']}"
        log(f"{label:28s} {t * 1e3:9.1f} us {fl / t / 1e9 if t > 0 else 0:8.1f} TF/s {100 * t / tot:5.1f}%  {extra}")
