"""Per-buffer error of a forced kernel variant against the CPU interpretation of the plan (debug aid).
    VGGHEADS_B200_SWAP=1 VGGHEADS_B200_XR=0 python tools/debug_variants.py [B] [S] [tune]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import plan_emulator as pe
from head_detector_b200 import synth
from head_detector_b200.engine import Engine
from oracle import net_oracle as no
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
S = int(sys.argv[2]) if len(sys.argv) > 2 else 640
eng = Engine(no.synthetic_weights(6), B, S)
if len(sys.argv) > 3 and sys.argv[3] == "tune":
    eng.autotune(3)
img = synth.synthetic_images(B, S, seed=12)
try:
    eng.forward(img.cuda())
    torch.cuda.synchronize()
except Exception as e:
    print("forward failed:", e)
    for i, op in enumerate(eng.plan.ops):
        if op.kind == 1:
            print(i, op.label, eng.op_config(i))
    sys.exit(1)
with torch.no_grad():
    ref = pe.run_plan(eng.packed, img, emulate_bf16=True)
writer = {}
for i, op in enumerate(eng.plan.ops):
    writer.setdefault(op.dst[0] if op.kind != 2 else op.src[0], []).append((i, op.label))
for name, bi in eng.plan.buf_names.items():
    got, want = eng.read_buffer(name), ref[bi]
    scale = want.abs().max().item() + 1e-6
    err = (got - want).abs()
    flag = "BAD" if err.max().item() > 2 ** -5 * scale + 1e-4 else "ok "
    ops = ", ".join(f"{l}:{eng.op_config(i)}" for i, l in writer.get(bi, []) if eng.plan.ops[i].kind == 1) if flag == "BAD" else ""
    print(f"{flag} {name:20s} max {err.max().item():9.4f} mean {err.mean().item():9.5f} scale {scale:9.3f} {ops}")
