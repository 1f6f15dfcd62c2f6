"""YoloHeads_L (VGGHeads_L) as an execution plan for libvggheads_b200.

Architecture source: the reference's arch spec
`yolo_head_training/configs/arch_params/yolo_heads_l_arch_params.yaml:1-141` (widths, depths),
`yolo_head_training/yolo_head/yolo_head_dfl_head.py:23-135` (per-level head) and the YoloNAS
building blocks of super_gradients (stem / stage / CSP / SPP / up- / down-stage; SURVEY.md
Appendix A.1).  Everything is in DEPLOY form: each QARepVGG block or Conv-BN-ReLU is one conv +
bias + ReLU (Appendix A.2); `fold_qarepvgg` / `fold_conv_bn` / `deploy_from_unfused` below produce that form
from as-trained tensors.

What this module adds on top of the spec is the B200 data layout:
  * activations are NHWC bf16 buffers; every conv writes straight into a channel slice of its
    consumer's buffer, so no torch.cat of the reference graph is ever materialised;
  * convs that read the same tensor are fused along N (CSP conv1||conv2, head stems, first tower
    layers), tiny parallel convs are fused as block-diagonal GEMMs (pred convs, 32-wide towers);
  * the CSP concat is stored as [a | b | b1..bn] and conv3's K columns are permuted at pack time;
  * weights are packed K-major [Cout_pad][tap][Cin] bf16 for 2-D TMA boxes.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _lib

STEM_OUT = 48
BACKBONE = [(96, 2, 96), (192, 3, 128), (384, 5, 256), (768, 2, 512)]  # (out, blocks, hidden) yaml:12-38
NECKS = {"neck1": (192, 4, 128), "neck2": (96, 4, 128), "neck3": (192, 4, 128), "neck4": (384, 4, 256)}  # yaml:52-88
HEADS = [(96, 128, 8), (192, 256, 16), (384, 512, 32)]  # (in, bbox_inter, stride) yaml:96-138
FLAME_INTER = 256
TOWERS = [("shape", 256, 128), ("expr", 128, 64), ("rot", 32, 6), ("jaw", 32, 3), ("scale", 32, 1), ("transl", 32, 3)]
REG_ROWS, FLAME_ROWS = 80, 208  # 68+1 (+pad) ; 128+64+6+3+3+1 (+pad)
# Sparse heads: the FLAME branch of a head level (pose stem -> 3 tower layers -> output convs) is only ever read at the
# anchors that survive NMS.  Its receptive field there is 7x7 pixels of the level's feature map, so the branch can run
# on PATCH x PATCH windows gathered around the survivors (window origin = anchor - PATCH_C) instead of the whole map.
PATCH, PATCH_C = 8, 3
# raw flame row layout: [shape128 | expr64 | rot6 | jaw3 | transl3 | scale1 | pad3]
RAW_ROW_OFF = {"shape": 0, "expr": 128, "rot": 192, "jaw": 198, "transl": 201, "scale": 204}


@dataclass
class Part:
    """One reference conv inside a (possibly fused) GEMM: rows [row, row+cout) of the packed matrix,
    reading logical input channels [log, log+n) from physical columns [phys, phys+n) per segment."""
    name: str
    row: int
    cout: int
    segs: List[Tuple[int, int, int]]  # (phys_col, logical_cin, n)
    transposed: bool = False


@dataclass
class Op:
    kind: int
    src: Tuple[int, int, int] = (0, 0, 0)   # (buf, coff, cin)
    dst: Tuple[int, int] = (0, 0)           # (buf, coff)
    cout: int = 0                           # stored channels
    k: int = 1
    stride: int = 1
    relu: int = 1
    up: int = 0
    up_cout: int = 0
    res: Optional[Tuple[int, int, str]] = None  # (buf, coff, alpha weight name)
    parts: List[Part] = field(default_factory=list)
    n_pad: int = 0
    label: str = ""
    lane: int = 0  # independent branches of the graph run on separate lanes (CUDA streams / graph branches)
    level: int = 0  # survivor-patch ops (sparse heads): 1 + head level whose patch stack the op works on; 0 = dense op


class Plan:
    def __init__(self, image_size: int):
        assert image_size % 32 == 0
        self.S = image_size
        self.bufs: List[Tuple[int, int, int, int]] = []
        self.buf_names: Dict[str, int] = {}
        self.ops: List[Op] = []
        self.reg_buf: List[int] = []
        self.flame_buf: List[int] = []
        self.lane = 0  # lane given to ops appended from now on
        self.level = 0  # level tag given to ops appended from now on (sparse heads)
        self.stack_bufs: set = set()  # buffers that are ONE stacked image [H,W,C] (survivor patches), not [B,H,W,C]
        self.n_dense_ops: Optional[int] = None  # sparse heads: ops [0, n_dense_ops) run before select/NMS, the rest after
        self.patch_cap = 0  # sparse heads: patch capacity per level (batch * keep_top_k)
        self.split = False  # parity mode (split_plan): bf16 buffers hold six planes [h|m|h|m|h|l] per 32-channel granule
        self.act_dtype = "bf16"  # storage format of the 16-bit activation buffers and of the packed weights (ACT_DTYPES)

    # -- helpers
    def buf(self, name, res, C, fp32=0):
        self.bufs.append((res, res, C, fp32))
        self.buf_names[name] = len(self.bufs) - 1
        return len(self.bufs) - 1

    def patch_buf(self, name, C, fp32=0):
        """Stack of `patch_cap` survivor patches of PATCH x PATCH pixels, stored as one tall image
        [patch_cap * PATCH, PATCH, C]: a 3x3 conv over the stack is right wherever its window stays inside one patch."""
        self.bufs.append((self.patch_cap * PATCH, PATCH, C, fp32))
        self.buf_names[name] = len(self.bufs) - 1
        self.stack_bufs.add(len(self.bufs) - 1)
        return len(self.bufs) - 1

    def conv(self, label, src, dst, parts, k=1, stride=1, relu=1, res=None, up=0, up_cout=0, cout=None):
        rows = max(p.row + p.cout for p in parts) if not up else 4 * up_cout
        stored = cout if cout is not None else rows
        assert stored % 16 == 0, (label, stored)
        self.ops.append(Op(_lib.OP_CONV, src, dst, stored, k, stride, relu, up, up_cout, res, parts, 0, label, self.lane, self.level))

    def simple(self, name, src, dst, cout, k=1, stride=1, relu=1, res=None, cin_logical=None):
        cin = src[2] if cin_logical is None else cin_logical
        self.conv(name, src, dst, [Part(name, 0, cout, [(0, 0, cin)])], k, stride, relu, res)

    def csp(self, name, src, dst, cout, n, hid, ci, res_):
        """YoloNASCSPLayer; physical concat [a | b | b1..bn] (ci) or [a->bn | b]."""
        R = res_
        cat = self.buf(name + ".cat", R, hid * (2 + (n if ci else 0)))
        tmp = self.buf(name + ".tmp", R, hid)
        self.conv(name + ".conv1|conv2", src, (cat, 0),
                  [Part(name + ".conv1", 0, hid, [(0, 0, src[2])]), Part(name + ".conv2", hid, hid, [(0, 0, src[2])])])
        if ci:
            t = (cat, 0)
            for j in range(n):
                self.simple(f"{name}.b{j}.cv1", (t[0], t[1], hid), (tmp, 0), hid, k=3)
                o = (cat, (2 + j) * hid)
                self.simple(f"{name}.b{j}.cv2", (tmp, 0, hid), o, hid, k=3, res=(t[0], t[1], f"{name}.b{j}.alpha"))
                t = o
            segs = [(0, 0, hid), (hid, (n + 1) * hid, hid)] + [((1 + j) * hid, j * hid, hid) for j in range(1, n + 1)]
            self.conv(name + ".conv3", (cat, 0, hid * (2 + n)), dst, [Part(name + ".conv3", 0, cout, segs)])
        else:
            pp = [self.buf(name + ".t0", R, hid), self.buf(name + ".t1", R, hid)]
            t = (cat, 0)
            for j in range(n):
                self.simple(f"{name}.b{j}.cv1", (t[0], t[1], hid), (tmp, 0), hid, k=3)
                o = (cat, 0) if j == n - 1 else (pp[j % 2], 0)
                self.simple(f"{name}.b{j}.cv2", (tmp, 0, hid), o, hid, k=3, res=(t[0], t[1], f"{name}.b{j}.alpha"))
                t = o
            self.simple(name + ".conv3", (cat, 0, 2 * hid), dst, cout)


# 16-bit storage formats of the throughput mode.  fp16 (default): 11 significant bits - the precision class of the TF32
# convolutions the reference's own GPU path runs by default, and the format the released model was trained in (AMP:
# yolo_head_training/configs/yolo_heads_l.yaml:21 `mixed_precision: True`), so its activations are known to fit the
# range.  bf16: 8 significant bits, fp32's range.  Same kernels and tensor-core rate either way.
ACT_DTYPES = {"bf16": torch.bfloat16, "fp16": torch.float16}


def default_act_dtype() -> str:
    import os

    v = os.environ.get("VGGHEADS_B200_ACT", "fp16").lower()
    if v not in ACT_DTYPES:
        raise ValueError(f"VGGHEADS_B200_ACT must be one of {sorted(ACT_DTYPES)}, got {v!r}")
    return v


def build_plan(image_size: int = 640, sparse_heads: Optional[Tuple[int, int]] = None, fused_stem: bool = True,
               act_dtype: str = "bf16") -> Plan:
    """`sparse_heads=(batch, keep_top_k)` builds the two-phase plan: the dense ops end with the box branch; the FLAME
    branch of every level follows select/NMS and runs on survivor patches (see PATCH).  `fused_stem=False` keeps the
    two-op stem (uint8 im2col buffer + Cin = 32 1x1 conv on the tensor-core GEMM kernel) that the parity mode splits.
    `act_dtype`: "bf16" | "fp16" (ACT_DTYPES)."""
    assert act_dtype in ACT_DTYPES, act_dtype
    P = Plan(image_size)
    P.act_dtype = act_dtype
    if sparse_heads is not None:
        P.patch_cap = int(sparse_heads[0]) * int(sparse_heads[1])
    S = image_size
    if fused_stem:   # one kernel: uint8 window -> im2col fragments in registers -> MMA -> bias + ReLU -> 64-channel rows
        stem = P.buf("stem", S // 2, 64)    # 48 real channels + 16 zeros (64-wide K blocks downstream)
        P.ops.append(Op(_lib.OP_STEM_CONV, (0, 0, 32), (stem, 0), 64, 1, 1, 1, parts=[Part("stem", 0, STEM_OUT, [(0, 0, 27)])], label="stem"))
    else:
        cols = P.buf("stem.cols", S // 2, 32)   # im2col rows of the uint8 image: 27 taps (ky,kx,c) + 5 zeros
        P.ops.append(Op(_lib.OP_STEM, (0, 0, 3), (cols, 0), 32, 3, 2, 0, label="stem.im2col"))
        stem = P.buf("stem", S // 2, 64)
        P.conv("stem", (cols, 0, 32), (stem, 0), [Part("stem", 0, STEM_OUT, [(0, 0, 27)])], cout=64)
    prev, prev_c, res = stem, 64, S // 2
    feats = []
    for i, (cout, n, hid) in enumerate(BACKBONE, start=1):
        res //= 2
        down = P.buf(f"stage{i}.down", res, cout)
        logical_cin = STEM_OUT if i == 1 else prev_c
        P.conv(f"stage{i}.down", (prev, 0, prev_c), (down, 0), [Part(f"stage{i}.down", 0, cout, [(0, 0, logical_cin)])], k=3, stride=2)
        out = P.buf(f"c{i + 1}" if i < 4 else "stage4.out", res, cout)
        P.csp(f"stage{i}.csp", (down, 0, cout), (out, 0), cout, n, hid, True, res)
        feats.append((out, cout, res))
        prev, prev_c = out, cout
    (c2, c2c, r4), (c3, c3c, r8), (c4, c4c, r16), (c5s, c5c, r32) = feats
    spp = P.buf("spp.cat", r32, 4 * 384)
    P.simple("spp.cv1", (c5s, 0, 768), (spp, 0), 384)
    P.ops.append(Op(_lib.OP_SPP, (spp, 0, 384), (spp, 384), 0, label="spp.pool"))
    c5 = P.buf("c5", r32, 768)
    P.simple("spp.cv2", (spp, 0, 1536), (c5, 0), 768)

    n3_in = P.buf("neck3.in", r16, 192)   # [down(96) | i2(96)]
    n4_in = P.buf("neck4.in", r32, 384)   # [down(192) | i1(192)]

    def up_stage(name, low, low_c, low_res, s1, s1c, s2, s2c, inter_dst, out_name):
        out, n, hid = NECKS[name]
        R = low_res * 2
        P.simple(name + ".reduce", (low, 0, low_c), inter_dst, out)
        fuse = P.buf(name + ".fuse_in", R, 3 * out)
        P.conv(name + ".up", (inter_dst[0], inter_dst[1], out), (fuse, 0),
               [Part(name + ".up", 0, 4 * out, [(0, 0, out)], transposed=True)], relu=0, up=1, up_cout=out)
        # the two skip branches only depend on backbone features: separate lanes let them fill the
        # tails of the (much earlier) backbone kernels
        P.lane = 13
        P.simple(name + ".skip1", (s1, 0, s1c), (fuse, out), out)
        P.lane = 14
        s2r = P.buf(name + ".s2r", R * 2, out)
        P.simple(name + ".skip2_reduce", (s2, 0, s2c), (s2r, 0), out)
        P.simple(name + ".skip2_down", (s2r, 0, out), (fuse, 2 * out), out, k=3, stride=2)
        P.lane = 0
        y = P.buf(name + ".y", R, out)
        P.simple(name + ".fuse", (fuse, 0, 3 * out), (y, 0), out)
        o = P.buf(out_name, R, out)
        P.csp(name + ".csp", (y, 0, out), (o, 0), out, n, hid, False, R)
        return o

    x = up_stage("neck1", c5, 768, r32, c4, c4c, c3, c3c, (n4_in, 192), "neck1.out")
    p3 = up_stage("neck2", x, 192, r16, c3, c3c, c2, c2c, (n3_in, 96), "p3")
    P.simple("neck3.down", (p3, 0, 96), (n3_in, 0), 96, k=3, stride=2)
    p4 = P.buf("p4", r16, 192)
    P.csp("neck3.csp", (n3_in, 0, 192), (p4, 0), 192, NECKS["neck3"][1], NECKS["neck3"][2], False, r16)
    P.simple("neck4.down", (p4, 0, 192), (n4_in, 0), 192, k=3, stride=2)
    p5 = P.buf("p5", r32, 384)
    P.csp("neck4.csp", (n4_in, 0, 384), (p5, 0), 384, NECKS["neck4"][1], NECKS["neck4"][2], False, r32)

    def flame_branch(h, src, R_buf, lanes):
        """towers0 -> 2 x (shape | expr | transf) -> output convs, reading the pose-stem activations `src`
        (buf, coff); buffers come from R_buf(name, C, fp32)."""
        lane_a, lane_b, lane_c = lanes
        P.lane = lane_a
        t_prev = R_buf(h + ".t0", 512)
        parts, row = [], 0
        for tower, inter, _ in TOWERS:
            parts.append(Part(f"{h}.{tower}.0", row, inter, [(0, 0, FLAME_INTER)]))
            row += inter
        P.conv(h + ".towers0", (src[0], src[1], FLAME_INTER), (t_prev, 0), parts, k=3)
        for i in (1, 2):
            if P.level:  # the conv padding of the dense graph: activations outside the image are zero for the next 3x3
                P.lane = lane_a
                P.ops.append(Op(_lib.OP_PATCH_MASK, (t_prev, 0, 512), (t_prev, 0), 512, label=f"{h}.mask.t{i - 1}", lane=lane_a, level=P.level))
            t_next = R_buf(f"{h}.t{i}", 512)
            P.lane = lane_a
            P.simple(f"{h}.shape.{i}", (t_prev, 0, 256), (t_next, 0), 256, k=3)
            P.lane = lane_b
            P.simple(f"{h}.expr.{i}", (t_prev, 256, 128), (t_next, 256), 128, k=3)
            P.lane = lane_c
            P.conv(f"{h}.transf.{i}", (t_prev, 384, 128), (t_next, 384),
                   [Part(f"{h}.{tw}.{i}", 32 * q, 32, [(32 * q, 0, 32)]) for q, tw in enumerate(("rot", "jaw", "scale", "transl"))], k=3)
            t_prev = t_next
        P.lane = lane_a
        fl = R_buf(h + ".flame_raw", FLAME_ROWS, 1)
        col = {"shape": (0, 256), "expr": (256, 128), "rot": (384, 32), "jaw": (416, 32), "scale": (448, 32), "transl": (480, 32)}
        P.conv(h + ".flame_out", (t_prev, 0, 512), (fl, 0),
               [Part(f"{h}.{tw}.out", RAW_ROW_OFF[tw], oc, [(col[tw][0], 0, col[tw][1])]) for tw, _, oc in TOWERS],
               relu=0, cout=FLAME_ROWS)
        P.flame_buf.append(fl)

    sparse = sparse_heads is not None
    for l, ((cin, bb, stride), feat) in enumerate(zip(HEADS, (p3, p4, p5)), start=1):
        h, R = f"head{l}", S // stride
        lane_a, lane_b, lane_c, lane_d = 1 + 4 * (l - 1), 2 + 4 * (l - 1), 3 + 4 * (l - 1), 4 + 4 * (l - 1)
        P.lane = lane_a
        if sparse:   # dense part: the box branch only
            st = P.buf(h + ".stems", R, bb)
            P.conv(h + ".bbox_stem", (feat, 0, cin), (st, 0), [Part(h + ".bbox_stem", 0, bb, [(0, 0, cin)])])
        else:
            st = P.buf(h + ".stems", R, bb + FLAME_INTER)
            P.conv(h + ".stems", (feat, 0, cin), (st, 0),
                   [Part(h + ".bbox_stem", 0, bb, [(0, 0, cin)]), Part(h + ".pose_stem", bb, FLAME_INTER, [(0, 0, cin)])])
        P.lane = lane_d   # box branch: cls|reg -> preds
        cr = P.buf(h + ".clsreg", R, 2 * bb)
        P.conv(h + ".cls|reg", (st, 0, bb), (cr, 0),
               [Part(h + ".cls_conv", 0, bb, [(0, 0, bb)]), Part(h + ".reg_conv", bb, bb, [(0, 0, bb)])], k=3)
        reg = P.buf(h + ".reg_raw", R, REG_ROWS, fp32=1)
        P.conv(h + ".preds", (cr, 0, 2 * bb), (reg, 0),
               [Part(h + ".reg_pred", 0, 68, [(bb, 0, bb)]), Part(h + ".cls_pred", 68, 1, [(0, 0, bb)])], relu=0, cout=REG_ROWS)
        P.reg_buf.append(reg)
        if not sparse:
            flame_branch(h, (st, bb), lambda name, C, fp32=0, R=R: P.buf(name, R, C, fp32), (lane_a, lane_b, lane_c))
    if sparse:
        P.lane = 0
        P.n_dense_ops = len(P.ops)
        for l, ((cin, bb, stride), feat) in enumerate(zip(HEADS, (p3, p4, p5)), start=1):
            h = f"head{l}"
            lanes = (1 + 4 * (l - 1), 2 + 4 * (l - 1), 3 + 4 * (l - 1))
            P.level, P.lane = l, lanes[0]
            pin = P.patch_buf(h + ".patch_in", cin)
            P.ops.append(Op(_lib.OP_PATCH_GATHER, (feat, 0, cin), (pin, 0), cin, label=h + ".patch_gather", lane=lanes[0], level=l))
            ps = P.patch_buf(h + ".pose_stem", FLAME_INTER)
            P.conv(h + ".pose_stem", (pin, 0, cin), (ps, 0), [Part(h + ".pose_stem", 0, FLAME_INTER, [(0, 0, cin)])])
            P.ops.append(Op(_lib.OP_PATCH_MASK, (ps, 0, FLAME_INTER), (ps, 0), FLAME_INTER, label=h + ".mask.pose_stem", lane=lanes[0], level=l))
            flame_branch(h, (ps, 0), lambda name, C, fp32=0: P.patch_buf(name, C, fp32), lanes)
        P.level = 0
    P.lane = 0
    return P


SPLIT_PLANES = 6
# weight term multiplying activation plane q of [x_h | x_m | x_h | x_m | x_h | x_l]:  0 = w_h, 1 = w_m, 2 = w_l
SPLIT_WEIGHT_TERM = (0, 0, 1, 1, 2, 0)


def split_plan(P: Plan) -> Plan:
    """Parity mode (include/vggheads_b200.h `vgh_net_desc.split`): the same dense plan over activations stored as three
    bf16 terms.  Logical channel c of a bf16 buffer lives at 192*(c//32) + 32*plane + c%32; channel offsets / widths of
    every op are scaled by 6; fp32 buffers (raw head outputs) are untouched.  Pure bookkeeping - the arithmetic change
    is in the epilogue (three-term split store) and in `pack` (weight terms along K)."""
    assert P.n_dense_ops is None, "parity mode runs the dense-heads plan"
    assert P.act_dtype == "bf16", "parity mode splits values into bf16 terms"
    P.split = True
    is_bf16 = [not fp32 for (_, _, _, fp32) in P.bufs]
    P.bufs = [(h, w, c * SPLIT_PLANES if not fp32 else c, fp32) for (h, w, c, fp32) in P.bufs]
    for op in P.ops:
        assert op.kind != _lib.OP_STEM_CONV, "parity mode splits the two-op stem: build_plan(fused_stem=False)"
        if op.kind == _lib.OP_STEM:
            continue
        if op.kind == _lib.OP_SPP:   # the kernel works in logical channels
            continue
        assert op.src[1] % 32 == 0 and op.src[2] % 32 == 0 and op.dst[1] % 32 == 0, op.label
        op.src = (op.src[0], op.src[1] * SPLIT_PLANES, op.src[2] * SPLIT_PLANES)
        if is_bf16[op.dst[0]]:
            op.dst = (op.dst[0], op.dst[1] * SPLIT_PLANES)
        if op.res is not None:
            op.res = (op.res[0], op.res[1] * SPLIT_PLANES, op.res[2])
    return P


def split_terms(w: torch.Tensor):
    """fp32 -> (h, m, l) bf16-representable fp32 tensors with h + m + l == w (exactly, in fp32)."""
    h = w.to(torch.bfloat16).float()
    m = (w - h).to(torch.bfloat16).float()
    l = (w - h - m).to(torch.bfloat16).float()
    return h, m, l


def _auto_block_n(cout: int, up: int, up_cout: int) -> int:
    lim = up_cout if up else cout
    if lim <= 256:
        return lim
    for n in range(256, 15, -16):
        if lim % n == 0:
            return n
    return 16


@dataclass
class PackedNet:
    plan: Plan
    weights: np.ndarray   # uint16: bf16 or fp16 bits (plan.act_dtype)
    bias: np.ndarray      # float32
    op_meta: List[dict]


def pack(plan: Plan, w: Dict[str, torch.Tensor]) -> PackedNet:
    """Pack deploy-form weights ({name}.w [Cout,Cin,k,k], {name}.b) for the plan."""
    wchunks, bchunks, meta = [], [], []
    w_off = b_off = 0
    for op in plan.ops:
        if op.kind not in (_lib.OP_CONV, _lib.OP_STEM_CONV):
            meta.append({})
            continue
        cin, taps = op.src[2] // (SPLIT_PLANES if plan.split else 1), op.k * op.k   # logical channels per tap
        block_n = _auto_block_n(op.cout, op.up, op.up_cout)
        n_pad = (op.cout + block_n - 1) // block_n * block_n
        if not op.up:  # operand-swapped kernel: G channel groups of gw <= 128, each fetched as a 128-row box
            G = (op.cout + 127) // 128
            if op.cout % G == 0:
                n_pad = max(n_pad, (op.cout // G) * (G - 1) + 128)
        M = torch.zeros(n_pad, taps, cin, dtype=torch.float32)
        bvec = torch.zeros(n_pad, dtype=torch.float32)
        for p in op.parts:
            wt, bt = w[p.name + ".w"].float(), w[p.name + ".b"].float()
            if p.transposed:  # ConvTranspose2d weight [Cin, Cout, 2, 2]; row = (dy*2+dx)*Cout + co
                co = wt.shape[1]
                for dy in range(2):
                    for dx in range(2):
                        sub = dy * 2 + dx
                        M[sub * co:(sub + 1) * co, 0, :] = wt[:, :, dy, dx].T
                        bvec[sub * co:(sub + 1) * co] = bt
                continue
            if p.name == "stem":  # 3x3x3 taps flattened (ky,kx,c) into the K columns; /255 of detector.py:51 folded in
                M[p.row:p.row + p.cout, 0, 0:27] = wt.permute(0, 2, 3, 1).reshape(p.cout, 27) / 255.0
                bvec[p.row:p.row + p.cout] = bt
                continue
            assert wt.shape[0] == p.cout and wt.shape[2] == op.k, (p.name, tuple(wt.shape), p.cout, op.k)
            assert sum(s[2] for s in p.segs) == wt.shape[1], (p.name, p.segs, wt.shape)
            wk = wt.permute(0, 2, 3, 1).reshape(p.cout, taps, wt.shape[1])  # [co][tap][ci]
            for phys, log, n in p.segs:
                M[p.row:p.row + p.cout, :, phys:phys + n] = wk[:, :, log:log + n]
            bvec[p.row:p.row + p.cout] = bt
        alpha = float(w[op.res[2]]) if op.res is not None else 0.0
        if plan.split:   # K columns per 32-channel granule: [w_h | w_h | w_m | w_m | w_l | w_h] against [x_h | x_m | x_h | x_m | x_h | x_l]
            terms = split_terms(M)
            G = M.reshape(n_pad, taps, cin // 32, 1, 32)
            M = torch.cat([terms[t].reshape_as(G) for t in SPLIT_WEIGHT_TERM], dim=3).reshape(n_pad, taps, cin * SPLIT_PLANES)
            cin = cin * SPLIT_PLANES
        q = M.reshape(-1).to(ACT_DTYPES[plan.act_dtype])
        if not bool(torch.isfinite(q).all()):
            raise ValueError(f"{op.label}: weights do not fit {plan.act_dtype}; use act_dtype='bf16'")
        chunk = q.view(torch.int16).numpy().view(np.uint16)
        pad = (-chunk.size) % 64  # keep every TMA base address 128-byte aligned
        if pad:
            chunk = np.concatenate([chunk, np.zeros(pad, dtype=np.uint16)])
        wchunks.append(chunk)
        bchunks.append(bvec.numpy())
        meta.append(dict(w_off=w_off, b_off=b_off, n_pad=n_pad, k_total=taps * cin, block_n=block_n, alpha=alpha))
        w_off += chunk.size
        b_off += n_pad
    return PackedNet(plan, np.concatenate(wchunks), np.concatenate(bchunks), meta)


def conv_names() -> List[Tuple[str, int, int, int, bool]]:
    """(name, k, cin, cout, is_transpose) of the 191 deploy-form convs, derived from the plan."""
    out = [("stem", 3, 3, STEM_OUT, False)]
    for op in build_plan(640).ops:
        for p in op.parts:
            if p.name == "stem":
                continue
            cin = sum(s[2] for s in p.segs)
            out.append((p.name, 2 if p.transposed else op.k, cin, p.cout // 4 if p.transposed else p.cout, p.transposed))
    return out


def synthetic_weights(seed: int = 0, bias_std: float = 0.02) -> Dict[str, torch.Tensor]:
    """Seeded random-init deploy-form weights (no checkpoint is reachable offline): He-normal,
    second-moment preserving through the residual bottlenecks (alpha 0.5, cv2 gain 1), pred convs
    gain 1, cls_pred.bias = -log(99) (yolo_head_dfl_head.py:188-190)."""
    g = torch.Generator().manual_seed(seed)
    w: Dict[str, torch.Tensor] = {}
    for name, k, cin, cout, tr in conv_names():
        if tr:
            w[name + ".w"] = torch.randn(cin, cout, 2, 2, generator=g) * math.sqrt(1.0 / cin)
        else:
            lin = name.endswith("_pred") or name.endswith(".out") or (name.endswith(".cv2") and ".csp.b" in name)
            w[name + ".w"] = torch.randn(cout, cin, k, k, generator=g) * math.sqrt((1.0 if lin else 2.0) / (cin * k * k))
        w[name + ".b"] = torch.randn(cout, generator=g) * bias_std
        if name.endswith(".cls_pred"):
            w[name + ".b"] = torch.full((cout,), -math.log(99.0))
        if name.endswith(".cv2") and ".csp.b" in name:
            w[name[:-4] + ".alpha"] = torch.tensor(0.5)
    return w


# ------------------------------------------------------------------------------------------ re-parameterisation
BN_EPS = 1e-6  # yolo_heads_l_arch_params.yaml:139


def _bn_scale_shift(sd: Dict[str, torch.Tensor], prefix: str, eps: float = BN_EPS):
    s = sd[prefix + ".weight"].double() / torch.sqrt(sd[prefix + ".running_var"].double() + eps)
    return s, sd[prefix + ".bias"].double() - sd[prefix + ".running_mean"].double() * s


def fold_conv_bn(sd: Dict[str, torch.Tensor], name: str, eps: float = BN_EPS):
    """Conv(bias=False)+BN (+ReLU) -> conv weight/bias.  Keys: `{name}.conv.weight`, `{name}.bn.*`."""
    s, t = _bn_scale_shift(sd, name + ".bn", eps)
    return (sd[name + ".conv.weight"].double() * s[:, None, None, None]).float(), t.float()


def fold_qarepvgg(sd: Dict[str, torch.Tensor], name: str, eps: float = BN_EPS):
    """QARepVGG block -> one 3x3 conv + bias:
    W = post_bn( bn3(W3) + alpha * pad(W1) + I ),  b likewise.  Keys: `{name}.conv3.weight`, `{name}.bn3.*`,
    `{name}.conv1.weight/.bias`, optional `{name}.alpha`, `{name}.post_bn.*`; the identity branch exists iff
    the block is square with stride 1 and `{name}.residual` (bool tensor / flag) is true or absent-but-square."""
    w3 = sd[name + ".conv3.weight"].double()
    s3, t3 = _bn_scale_shift(sd, name + ".bn3", eps)
    alpha = sd[name + ".alpha"].double() if (name + ".alpha") in sd else torch.tensor(1.0, dtype=torch.float64)
    w = w3 * s3[:, None, None, None] + alpha * torch.nn.functional.pad(sd[name + ".conv1.weight"].double(), [1, 1, 1, 1])
    b = t3 + alpha * sd[name + ".conv1.bias"].double()
    residual = bool(sd[name + ".residual"]) if (name + ".residual") in sd else False
    if residual:
        assert w.shape[0] == w.shape[1], "identity branch needs a square block"
        idx = torch.arange(w.shape[0])
        w[idx, idx, 1, 1] += 1.0
    sp, tp = _bn_scale_shift(sd, name + ".post_bn", eps)
    return (w * sp[:, None, None, None]).float(), (b * sp + tp).float()


def deploy_from_unfused(sd: Dict[str, torch.Tensor], eps: float = BN_EPS) -> Dict[str, torch.Tensor]:
    """As-trained tensors (per layer name of `conv_names()`) -> the deploy-form dict `pack()` takes.
    A layer may be given as a QARepVGG block (`{name}.conv3.weight` ...), a Conv-BN (`{name}.conv.weight` +
    `{name}.bn.*`) or an already plain conv (`{name}.w` / `{name}.b`, or torch style `{name}.weight/.bias`).
    Mapping super_gradients' own state_dict keys onto these names is the remaining step of SURVEY 8 f3
    (super_gradients is not available offline to verify key names against)."""
    out: Dict[str, torch.Tensor] = {}
    for name, k, cin, cout, tr in conv_names():
        if name + ".conv3.weight" in sd:
            out[name + ".w"], out[name + ".b"] = fold_qarepvgg(sd, name, eps)
        elif name + ".conv.weight" in sd:
            out[name + ".w"], out[name + ".b"] = fold_conv_bn(sd, name, eps)
        elif name + ".w" in sd:
            out[name + ".w"], out[name + ".b"] = sd[name + ".w"].float(), sd[name + ".b"].float()
        elif name + ".weight" in sd:
            out[name + ".w"], out[name + ".b"] = sd[name + ".weight"].float(), sd[name + ".bias"].float()
        else:
            raise KeyError(f"no tensors for layer {name!r}")
        exp = (cin, cout, 2, 2) if tr else (cout, cin, k, k)
        if tuple(out[name + ".w"].shape) != exp:
            raise ValueError(f"{name}: weight shape {tuple(out[name + '.w'].shape)} != {exp}")
        if name.endswith(".cv2") and ".csp.b" in name:
            a = name[:-4] + ".alpha"
            out[a] = sd[a].float().reshape(()) if a in sd else torch.tensor(1.0)
    return out


def total_macs(image_size: int = 640) -> int:
    """Algorithmic MACs per image of the deploy-form network (83.34 G @640)."""
    P = build_plan(image_size)
    tot = 3 * STEM_OUT * 9 * (image_size // 2) ** 2
    for op in P.ops:
        if op.kind != _lib.OP_CONV:
            continue   # (the stem, fused or not, is counted above with its true K = 27)
        src_res = P.bufs[op.src[0]][0]
        out_res = src_res if op.up else src_res // op.stride
        for p in op.parts:
            if p.name == "stem":
                continue
            cin = sum(s[2] for s in p.segs)
            tot += cin * p.cout * out_res * out_res * (1 if p.transposed else op.k * op.k)
    return tot
