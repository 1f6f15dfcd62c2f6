"""CPU interpreter of a head_detector_b200.arch plan + packed weights (test infrastructure).

Executes exactly what the CUDA executor is told to do (op list, channel slices, packed K-major
matrices, residuals, pixel-shuffle) with plain torch ops, so that plan wiring and weight packing
can be checked against the oracle network on a machine without a GPU."""
import numpy as np
import torch
import torch.nn.functional as F

from head_detector_b200 import _lib, arch


def bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


def act_round(dtype_name):
    """Rounding of a stored activation for plan.act_dtype ("bf16" | "fp16")."""
    dt = arch.ACT_DTYPES[dtype_name]
    return lambda t: t.to(dt).to(torch.float32)


def apply_op(pk, op, m, bufs, images_u8=None, emulate_bf16=True, patches=None):
    """Execute ONE plan op on the given NHWC float buffers (in place).  `patches[level]` = list of (image, y, x)
    survivor anchors of head level `level` (1-based) for the sparse-heads ops."""
    P = pk.plan
    rnd = act_round(P.act_dtype) if emulate_bf16 else (lambda t: t)   # `emulate_bf16`: emulate the 16-bit storage (either format)
    wdt = arch.ACT_DTYPES[P.act_dtype]
    if op.kind in (_lib.OP_PATCH_GATHER, _lib.OP_PATCH_MASK):
        PT, PC = arch.PATCH, arch.PATCH_C
        sb, so, C = op.src
        db, do = op.dst
        for pi, (b, y, x) in enumerate(patches[op.level]):
            feat_h, feat_w = (bufs[sb].shape[1], bufs[sb].shape[2]) if op.kind == _lib.OP_PATCH_GATHER else P.bufs[P.reg_buf[op.level - 1]][:2]
            for r in range(PT):
                for c in range(PT):
                    yy, xx = y - PC + r, x - PC + c
                    inside = 0 <= yy < feat_h and 0 <= xx < feat_w
                    if op.kind == _lib.OP_PATCH_GATHER:
                        bufs[db][0, pi * PT + r, c, do:do + C] = bufs[sb][b, yy, xx, so:so + C] if inside else 0.0
                    elif not inside:
                        bufs[db][0, pi * PT + r, c, do:do + C] = 0.0
        return
    if op.kind == _lib.OP_STEM:  # im2col of the uint8 image: [B,S/2,S/2,32] = 27 taps (ky,kx,c) + 5 zeros
        x = images_u8.permute(0, 3, 1, 2).float()
        cols = F.unfold(x, kernel_size=3, padding=1, stride=2)            # [B, c*9 + ky*3 + kx, L]
        Bn, _, L = cols.shape
        Ho = images_u8.shape[1] // 2
        cols = cols.view(Bn, 3, 9, L).permute(0, 3, 2, 1).reshape(Bn, Ho, Ho, 27)   # -> (ky,kx,c) order
        bufs[op.dst[0]][..., :27] = cols
        bufs[op.dst[0]][..., 27:] = 0
    elif op.kind == _lib.OP_STEM_CONV:  # fused stem: the same im2col rows times the packed [n_pad][32] taps, + bias, ReLU
        x = images_u8.permute(0, 3, 1, 2).float()
        cols = F.unfold(x, kernel_size=3, padding=1, stride=2)
        Bn, _, L = cols.shape
        Ho = images_u8.shape[1] // 2
        cols = cols.view(Bn, 3, 9, L).permute(0, 3, 2, 1).reshape(Bn, Ho, Ho, 27)
        wts = torch.from_numpy(pk.weights[m["w_off"]:m["w_off"] + m["n_pad"] * 32].view(np.int16).copy()).view(wdt).float().reshape(m["n_pad"], 32)
        y = cols @ wts[:op.cout, :27].T + torch.from_numpy(pk.bias[m["b_off"]:m["b_off"] + op.cout])
        bufs[op.dst[0]][..., :op.cout] = rnd(F.relu(y) if op.relu else y)
    elif op.kind == _lib.OP_SPP:
        b = bufs[op.src[0]]
        C = op.src[2]
        x = b[..., :C].permute(0, 3, 1, 2)
        for i, k in enumerate((5, 9, 13), start=1):
            b[..., i * C:(i + 1) * C] = F.max_pool2d(x, k, 1, k // 2).permute(0, 2, 3, 1)
    else:
        sb, so, cin = op.src
        x = bufs[sb][..., so:so + cin].permute(0, 3, 1, 2)
        wts = torch.from_numpy(pk.weights[m["w_off"]:m["w_off"] + m["n_pad"] * m["k_total"]].view(np.int16).copy()).view(wdt).float()
        W = wts.reshape(m["n_pad"], op.k, op.k, cin).permute(0, 3, 1, 2)
        bv = torch.from_numpy(pk.bias[m["b_off"]:m["b_off"] + m["n_pad"]])
        y = F.conv2d(x, W, bv, stride=op.stride, padding=op.k // 2)[:, :op.cout]
        if op.relu:
            y = F.relu(y)
        if op.res is not None:
            rb, ro, _ = op.res
            y = y + m["alpha"] * bufs[rb][..., ro:ro + op.cout].permute(0, 3, 1, 2)
        y = y.permute(0, 2, 3, 1)
        db, do = op.dst
        if not P.bufs[db][3]:
            y = rnd(y)
        if op.up:
            co = op.up_cout
            for sub in range(4):
                bufs[db][:, (sub >> 1)::2, (sub & 1)::2, do:do + co] = y[..., sub * co:(sub + 1) * co]
        else:
            bufs[db][..., do:do + op.cout] = y


def run_plan(pk: arch.PackedNet, images_u8: torch.Tensor, emulate_bf16: bool = True, patches=None):
    """images_u8 [B,S,S,3] uint8 -> list of NHWC float buffers.  Two-phase (sparse heads) plans run their dense ops
    only unless `patches` = {level: [(image, y, x), ...]} names the survivor anchors the patch ops work on."""
    B = images_u8.shape[0]
    P = pk.plan
    bufs = [torch.zeros(1 if i in P.stack_bufs else B, h, w, c) for i, (h, w, c, _) in enumerate(P.bufs)]
    n_ops = len(P.ops) if (patches is not None or P.n_dense_ops is None) else P.n_dense_ops
    for op, m in list(zip(P.ops, pk.op_meta))[:n_ops]:
        apply_op(pk, op, m, bufs, images_u8, emulate_bf16, patches)
    return bufs


def raw_to_oracle_layout(pk, bufs, level):
    """reg [B,68,H,W], cls [B,1,H,W], towers dict as the oracle's head_level returns them."""
    reg = bufs[pk.plan.reg_buf[level]].permute(0, 3, 1, 2)
    fl = bufs[pk.plan.flame_buf[level]].permute(0, 3, 1, 2)
    t = {tw: fl[:, arch.RAW_ROW_OFF[tw]:arch.RAW_ROW_OFF[tw] + oc] for tw, _, oc in arch.TOWERS}
    return reg[:, :68], reg[:, 68:69], t
