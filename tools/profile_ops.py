"""Per-op device times of one step (CUDA-event pairs, eager) -> gpurun_out/ops_<tag>.txt."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from head_detector_b200 import arch, synth  # noqa: E402
from head_detector_b200.engine import Engine  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 640
    tag = sys.argv[3] if len(sys.argv) > 3 else "r2"
    flags = sys.argv[4:]
    eng = Engine(arch.synthetic_weights(0), B, S, sparse_heads="dense" not in flags)
    eng.input.copy_(synth.synthetic_images(B, S, 0).cuda())
    boxes, scores = synth.engineered_heads(B, eng.A, S, 8, seed=7)
    eng.set_override(boxes.cuda(), scores.cuda())
    if "untuned" not in flags:
        eng.autotune(5)
    eng.profile(iters=2)
    rows = eng.profile(iters=10)
    tot = sum(t for _, t, _ in rows)
    conv_ms = sum(t for _, t, f in rows if f)
    conv_fl = sum(f for _, t, f in rows if f)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"ops_{tag}.txt"), "w") as f:
        def log(s):
            print(s)
            f.write(s + "\n")
        log(f"B={B} S={S} total {tot:.3f} ms/step ({B / tot * 1e3:.0f} img/s eager), conv {conv_ms:.3f} ms = {conv_fl / conv_ms / 1e9:.1f} TFLOP/s")
        for (label, t, fl), op in zip(rows, list(eng.plan.ops) + [None] * 4):
            extra = ""
            if op is not None and fl:
                r = eng.plan.bufs[op.src[0]][0]
                c = eng.op_config(list(eng.plan.ops).index(op))
                extra = f"k{op.k} s{op.stride} cin{op.src[2]:5d} cout{op.cout:5d} in{r:4d} mt{c['mt']} st{c['stages']} bn{c['block_n']} bk{c['bk']} tile{c['tw']}x{c['th']}"
            log(f"{label:28s} {t * 1e3:9.1f} us {fl / t / 1e9 if t > 0 else 0:8.1f} TF/s {100 * t / tot:5.1f}%  {extra}")


if __name__ == "__main__":
    main()
