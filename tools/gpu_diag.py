"""GPU bring-up diagnostic: runs the small network once and checks EVERY op in isolation - each op's
expected output is computed on the CPU from the inputs the GPU actually saw - so a broken kernel
configuration (k, stride, BK, transpose, residual ...) is pinpointed instead of smeared downstream.
Writes gpurun_out/diag.txt."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch  # noqa: E402

import plan_emulator as pe  # noqa: E402
from head_detector_b200 import _lib  # noqa: E402
from head_detector_b200.engine import Engine  # noqa: E402
from oracle import net_oracle as no  # noqa: E402


def main(S=128, B=2):
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    out = open(os.path.join(ROOT, "gpurun_out", "diag.txt"), "w")

    def log(*a):
        line = " ".join(str(x) for x in a)
        print(line)
        out.write(line + "\n")
        out.flush()

    log("device", torch.cuda.get_device_name(0), "cap", torch.cuda.get_device_capability(0))
    eng = Engine(no.synthetic_weights(3), B, S)
    torch.manual_seed(0)
    img = torch.randint(0, 256, (B, S, S, 3), dtype=torch.uint8)
    eng.forward(img.cuda())
    torch.cuda.synchronize()
    log("forward ok; anchors", eng.A)
    names = {i: n for n, i in eng.plan.buf_names.items()}
    gpu = [eng.read_buffer(names[i]) for i in range(len(eng.plan.bufs))]
    # buffers written more than once per forward cannot be checked post hoc
    writes = {}
    for op in eng.plan.ops:
        if op.kind == _lib.OP_SPP:
            continue
        key = (op.dst[0], op.dst[1])
        writes[key] = writes.get(key, 0) + 1
    multi = {k[0] for k, v in writes.items() if v > 1}
    n_bad = 0
    for op, m in zip(eng.plan.ops, eng.packed.op_meta):
        if op.dst[0] in multi or op.src[0] in multi or (op.res is not None and op.res[0] in multi):
            status = "skip(reused buffer)"
            log(f"{op.label:28s} {status}")
            continue
        work = [g.clone() for g in gpu]
        with torch.no_grad():
            pe.apply_op(eng.packed, op, m, work, img, True)
        d = op.dst[0]
        err = (work[d] - gpu[d]).abs()
        scale = work[d].abs().max().item() + 1e-9
        ok = err.max().item() <= 2 ** -6 * scale + 1e-5
        n_bad += (not ok)
        bk = 64 if op.src[2] % 64 == 0 else 32
        log(f"{op.label:28s} k{op.k} s{op.stride} cin{op.src[2]:5d} cout{op.cout:5d} bk{bk} up{op.up} res{int(op.res is not None)} "
            f"max_err {err.max().item():.4g} mean {err.mean().item():.3g} scale {scale:.3g} {'OK' if ok else 'BAD'}")
    log("bad ops:", n_bad)
    return n_bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
