"""CPU oracle (plain torch fp32) for the VGGHeads_L conv network.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED at the super_gradients boundary: the released weights (`vgg_heads_l.trcd`,
HF hub) and `super_gradients>=3.7` (yolo_head_training/requirements.txt:1) are not available
offline and the reference holds no numeric fixture for the conv network
(yolo_head_training/tests/test_models.py:29-36 asserts nothing).  This file restates the
architecture from the reference's own spec and sources:

  * widths / depths / module kinds .. yolo_head_training/configs/arch_params/yolo_heads_l_arch_params.yaml:1-141
  * per-level head ................... yolo_head_training/yolo_head/yolo_head_dfl_head.py:23-186
  * multi-level decode ............... yolo_head_training/yolo_head/yolo_head_ndfl_heads.py:117-175,206-235
  * top-k decoding module ............ yolo_head_training/yolo_head/yolo_heads.py:44-86
  * YoloNAS stem/stage/CSP/SPP/up/down-stage and QARepVGG semantics: super_gradients (third
    party, restated from its published module definitions)

Two forms are provided:
  * `QARepVGGUnfused` + `fold_qarepvgg` / `fold_conv_bn`: the as-trained multi-branch blocks and
    the re-parameterisation algebra (Appendix A.2), checked against each other in tests;
  * `DeployNet`: the deploy (fully folded) network as a pure function of a {name: tensor}
    weight dict - the form the CUDA path executes and is compared with layer by layer.

Weight naming contract (shared with head_detector_b200/arch.py):  "<layer>.w" [Cout,Cin,k,k],
"<layer>.b" [Cout]; bottleneck shortcut scale "<csp>.b<j>.alpha" []; conv-transpose
"<stage>.up.w" [Cin,Cout,2,2].
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

BN_EPS = 1e-6  # yolo_heads_l_arch_params.yaml:139
REG_MAX = 16   # yaml:93
STRIDES = (8, 16, 32)

# (out_channels, num_blocks, hidden_channels) per backbone stage - yaml:12-38
BACKBONE_STAGES = [(96, 2, 96), (192, 3, 128), (384, 5, 256), (768, 2, 512)]
STEM_OUT = 48          # yaml:8
# neck stages - yaml:52-88: (out, num_blocks, hidden)
NECK = {"neck1": (192, 4, 128), "neck2": (96, 4, 128), "neck3": (192, 4, 128), "neck4": (384, 4, 256)}
# per-level head: (in_channels, bbox_inter_channels) - yaml:96-138 ; flame_inter=256, towers 256/128/32, outs 128/64
HEADS = [(96, 128), (192, 256), (384, 512)]
FLAME_INTER, SHAPE_INTER, EXPR_INTER, TRANSF_INTER = 256, 256, 128, 32
SHAPE_OUT, EXPR_OUT = 128, 64
TOWERS = [("shape", SHAPE_INTER, SHAPE_OUT), ("expr", EXPR_INTER, EXPR_OUT), ("rot", TRANSF_INTER, 6),
          ("jaw", TRANSF_INTER, 3), ("scale", TRANSF_INTER, 1), ("transl", TRANSF_INTER, 3)]


# ----------------------------------------------------------------------------- unfused blocks + folding
class QARepVGGUnfused(nn.Module):
    """y = ReLU(post_bn(BN(conv3x3(x)) + alpha*(conv1x1(x)+b1) + [x]))  (SURVEY Appendix A.1)."""

    def __init__(self, cin, cout, stride=1, residual=True, use_alpha=False):
        super().__init__()
        self.conv3 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(cout, eps=BN_EPS)
        self.conv1 = nn.Conv2d(cin, cout, 1, stride, 0, bias=True)
        self.alpha = nn.Parameter(torch.tensor(1.0)) if use_alpha else 1.0
        self.residual = residual and cin == cout and stride == 1
        self.post_bn = nn.BatchNorm2d(cout, eps=BN_EPS)

    def forward(self, x):
        y = self.bn3(self.conv3(x)) + self.alpha * self.conv1(x)
        if self.residual:
            y = y + x
        return F.relu(self.post_bn(y))


def randomize_bn_(bn: nn.BatchNorm2d, g: torch.Generator):
    with torch.no_grad():
        bn.weight.copy_(0.5 + torch.rand(bn.weight.shape, generator=g))
        bn.bias.copy_(0.1 * torch.randn(bn.bias.shape, generator=g))
        bn.running_mean.copy_(0.1 * torch.randn(bn.bias.shape, generator=g))
        bn.running_var.copy_(0.5 + torch.rand(bn.bias.shape, generator=g))


def fold_conv_bn(w: torch.Tensor, bn: nn.BatchNorm2d) -> Tuple[torch.Tensor, torch.Tensor]:
    s = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    return w * s[:, None, None, None], bn.bias - bn.running_mean * s


def fold_qarepvgg(m: QARepVGGUnfused) -> Tuple[torch.Tensor, torch.Tensor]:
    """Appendix A.2: one 3x3 conv + bias equivalent to the three branches and both BNs (eval mode)."""
    w, b = fold_conv_bn(m.conv3.weight, m.bn3)
    alpha = m.alpha if isinstance(m.alpha, float) else m.alpha.detach()
    w = w + alpha * F.pad(m.conv1.weight, [1, 1, 1, 1])
    b = b + alpha * m.conv1.bias
    if m.residual:
        eye = torch.zeros_like(w)
        idx = torch.arange(w.shape[0])
        eye[idx, idx, 1, 1] = 1.0
        w = w + eye
    s = m.post_bn.weight / torch.sqrt(m.post_bn.running_var + m.post_bn.eps)
    return (w * s[:, None, None, None]).detach(), ((b - m.post_bn.running_mean) * s + m.post_bn.bias).detach()


# ----------------------------------------------------------------------------- deploy-form network
class DeployNet:
    """Deploy-form YoloHeads_L as a function of a weight dict.  `forward` returns what the traced
    reference model returns: boxes [B,A,4], scores [B,A,1], flame [B,A,413]
    (yolo_head_ndfl_heads.py:174-175).  `taps` (optional dict) receives named intermediates."""

    def __init__(self, weights: Dict[str, torch.Tensor], act_round=None):
        self.w = weights
        # act_round: optional callable applied to every stored activation (e.g. bf16 rounding) so the
        # oracle can mimic the storage precision of the CUDA path while keeping fp32 accumulation.
        self.rnd = act_round if act_round is not None else (lambda t: t)

    # --- primitives
    def conv(self, name, x, stride=1, relu=True):
        w = self.w[name + ".w"]
        y = F.conv2d(x, w, self.w[name + ".b"], stride=stride, padding=w.shape[-1] // 2)
        return self.rnd(F.relu(y) if relu else y)

    def csp(self, name, x, n, concat_intermediates):
        a = self.conv(name + ".conv1", x)
        b = self.conv(name + ".conv2", x)
        outs = [a]
        for j in range(n):
            t = outs[-1]
            y = self.conv(f"{name}.b{j}.cv2", self.conv(f"{name}.b{j}.cv1", t))
            outs.append(self.rnd(self.w[f"{name}.b{j}.alpha"] * t + y))
        cat = torch.cat((outs if concat_intermediates else outs[-1:]) + [b], dim=1)
        return self.conv(name + ".conv3", cat)

    def up_stage(self, name, x, skip1, skip2, n):
        xi = self.conv(name + ".reduce", x)
        u = self.rnd(F.conv_transpose2d(xi, self.w[name + ".up.w"], self.w[name + ".up.b"], stride=2))
        s1 = self.conv(name + ".skip1", skip1)
        s2 = self.conv(name + ".skip2_down", self.conv(name + ".skip2_reduce", skip2), stride=2)
        y = self.conv(name + ".fuse", torch.cat([u, s1, s2], dim=1))
        return xi, self.csp(name + ".csp", y, n, False)

    def down_stage(self, name, x, skip, n):
        y = self.conv(name + ".down", x, stride=2)
        return self.csp(name + ".csp", torch.cat([y, skip], dim=1), n, False)

    def head_level(self, name, x):
        pose = self.conv(name + ".pose_stem", x)
        bbox = self.conv(name + ".bbox_stem", x)
        cls = self.conv(name + ".cls_pred", self.conv(name + ".cls_conv", bbox), relu=False)
        reg = self.conv(name + ".reg_pred", self.conv(name + ".reg_conv", bbox), relu=False)
        outs = {}
        for tower, _, _ in TOWERS:
            t = pose
            for i in range(3):
                t = self.conv(f"{name}.{tower}.{i}", t)
            outs[tower] = self.conv(f"{name}.{tower}.out", t, relu=False)
        return reg, cls, outs

    # --- whole graph
    def features(self, x, taps=None):
        x = self.conv("stem", x, stride=2)
        feats = []
        for i, (_, n, _) in enumerate(BACKBONE_STAGES, start=1):
            x = self.conv(f"stage{i}.down", x, stride=2)
            x = self.csp(f"stage{i}.csp", x, n, True)
            feats.append(x)
        c2, c3, c4, c5 = feats
        h = self.conv("spp.cv1", c5)
        pools = [F.max_pool2d(h, k, 1, k // 2) for k in (5, 9, 13)]
        c5 = self.conv("spp.cv2", torch.cat([h] + pools, dim=1))
        i1, x = self.up_stage("neck1", c5, c4, c3, NECK["neck1"][1])
        i2, p3 = self.up_stage("neck2", x, c3, c2, NECK["neck2"][1])
        p4 = self.down_stage("neck3", p3, i2, NECK["neck3"][1])
        p5 = self.down_stage("neck4", p4, i1, NECK["neck4"][1])
        if taps is not None:
            taps.update(c2=c2, c3=c3, c4=c4, c5=c5, p3=p3, p4=p4, p5=p5)
        return p3, p4, p5

    def raw_heads(self, feats):
        return [self.head_level(f"head{l + 1}", f) for l, f in enumerate(feats)]

    def forward(self, x, taps=None):
        raw = self.raw_heads(self.features(x, taps))
        if taps is not None:
            taps["raw"] = raw
        return decode_heads(raw)


def anchor_points(sizes: List[Tuple[int, int]]):
    """yolo_head_ndfl_heads.py:206-235: (x+0.5, y+0.5) row-major per level, level-major concat."""
    pts, strides = [], []
    for (h, w), s in zip(sizes, STRIDES):
        ys, xs = torch.meshgrid(torch.arange(h) + 0.5, torch.arange(w) + 0.5, indexing="ij")
        pts.append(torch.stack([xs, ys], dim=-1).reshape(-1, 2).float())
        strides.append(torch.full((h * w, 1), float(s)))
    return torch.cat(pts), torch.cat(strides)


def decode_heads(raw):
    """DFL expectation, sigmoid, distance2bbox, FLAME post-ops and the channel rotation
    (yolo_head_dfl_head.py:162-184, yolo_head_ndfl_heads.py:137-172)."""
    regs, clss, flames, sizes = [], [], [], []
    proj = torch.linspace(0, REG_MAX, REG_MAX + 1).reshape(1, REG_MAX + 1, 1, 1)
    for reg, cls, t in raw:
        b, _, h, w = reg.shape
        sizes.append((h, w))
        r = reg.reshape(b, 4, REG_MAX + 1, h * w).permute(0, 2, 3, 1)
        regs.append((F.softmax(r, dim=1) * proj).sum(1))          # [B, HW, 4]
        clss.append(cls.reshape(b, 1, h * w))
        shape = F.pad(torch.tanh(t["shape"]) * 3, (0, 0, 0, 0, 0, 300 - SHAPE_OUT))
        expr = F.pad(torch.tanh(t["expr"]) * 3, (0, 0, 0, 0, 0, 100 - EXPR_OUT))
        scale = torch.exp(t["scale"]) / 0.05
        flames.append(torch.cat([shape, expr, t["rot"], t["jaw"], t["transl"], scale], dim=1).flatten(2))
    pts, strides = anchor_points(sizes)
    scores = torch.cat(clss, dim=-1).permute(0, 2, 1).sigmoid()     # [B, A, 1]
    d = torch.cat(regs, dim=1)                                      # [B, A, 4] (l, t, r, b)
    boxes = torch.cat([pts - d[..., :2], pts + d[..., 2:]], dim=-1) * strides
    c = torch.cat(flames, dim=-1)                                   # [B, 413, A], head order [..rot6|jaw3..]
    # from_3dmm reads [jaw3|rot6] at 400..408, to_3dmm writes [rot6|jaw3]: a rotation by 3 of 400..408
    jaw_read, rot_read = c[:, 400:403], c[:, 403:409]
    transl = c[:, 409:412].clone()
    transl[:, 0:2] += (pts * strides).T[None]
    scale = c[:, 412:413] * strides[:, 0][None, None]
    out = torch.cat([c[:, :400], rot_read, jaw_read, transl, scale], dim=1)
    return boxes, scores, out.permute(0, 2, 1).contiguous()


def topk_decode(boxes, scores, flame, k=1000):
    """VGGHeadDecodingModule.forward (yolo_heads.py:44-86): per-image sorted top-k + gather."""
    idx = torch.topk(scores, dim=1, k=k, largest=True, sorted=True).indices[..., 0]
    g = lambda t: torch.gather(t, 1, idx[..., None].expand(-1, -1, t.shape[-1]))
    return g(boxes), g(scores), g(flame)


# ----------------------------------------------------------------------------- layer enumeration (independent count)
def conv_layer_list() -> List[Tuple[str, int, int, int, int, int]]:
    """(name, k, stride, cin, cout, out_stride_wrt_input) of every conv in deploy form; used by the
    oracle tests to re-derive SURVEY's 191 convs / 83.34 GMAC figure independently of the product."""
    L = [("stem", 3, 2, 3, STEM_OUT, 2)]
    cin, os_ = STEM_OUT, 2

    def csp(name, cin, cout, n, hid, ci, os_):
        L.append((name + ".conv1", 1, 1, cin, hid, os_)); L.append((name + ".conv2", 1, 1, cin, hid, os_))
        for j in range(n):
            L.append((f"{name}.b{j}.cv1", 3, 1, hid, hid, os_)); L.append((f"{name}.b{j}.cv2", 3, 1, hid, hid, os_))
        L.append((name + ".conv3", 1, 1, hid * (2 + (n if ci else 0)), cout, os_))

    for i, (cout, n, hid) in enumerate(BACKBONE_STAGES, start=1):
        os_ *= 2
        L.append((f"stage{i}.down", 3, 2, cin, cout, os_))
        csp(f"stage{i}.csp", cout, cout, n, hid, True, os_)
        cin = cout
    L.append(("spp.cv1", 1, 1, 768, 384, 32)); L.append(("spp.cv2", 1, 1, 1536, 768, 32))

    def up(name, c_low, c_s1, c_s2, os_out):
        out, n, hid = NECK[name]
        L.append((name + ".reduce", 1, 1, c_low, out, os_out * 2)); L.append((name + ".up", 2, 2, out, out, os_out))
        L.append((name + ".skip1", 1, 1, c_s1, out, os_out)); L.append((name + ".skip2_reduce", 1, 1, c_s2, out, os_out // 2))
        L.append((name + ".skip2_down", 3, 2, out, out, os_out)); L.append((name + ".fuse", 1, 1, 3 * out, out, os_out))
        csp(name + ".csp", out, out, n, hid, False, os_out)

    def down(name, c, c_skip, os_out):
        out, n, hid = NECK[name]
        L.append((name + ".down", 3, 2, c, out // 2, os_out))
        csp(name + ".csp", out // 2 + c_skip, out, n, hid, False, os_out)

    up("neck1", 768, 384, 192, 16); up("neck2", 192, 192, 96, 8); down("neck3", 96, 96, 16); down("neck4", 192, 192, 32)
    for l, ((cin_h, bb), s) in enumerate(zip(HEADS, STRIDES), start=1):
        h = f"head{l}"
        L.append((h + ".pose_stem", 1, 1, cin_h, FLAME_INTER, s)); L.append((h + ".bbox_stem", 1, 1, cin_h, bb, s))
        L.append((h + ".cls_conv", 3, 1, bb, bb, s)); L.append((h + ".reg_conv", 3, 1, bb, bb, s))
        L.append((h + ".reg_pred", 1, 1, bb, 4 * (REG_MAX + 1), s)); L.append((h + ".cls_pred", 1, 1, bb, 1, s))
        for tower, inter, outc in TOWERS:
            c = FLAME_INTER
            for i in range(3):
                L.append((f"{h}.{tower}.{i}", 3, 1, c, inter, s)); c = inter
            L.append((f"{h}.{tower}.out", 1, 1, inter, outc, s))
    return L


def total_macs(size: int = 640) -> int:
    tot = 0
    for name, k, s, cin, cout, os_ in conv_layer_list():
        hw = (size // os_) ** 2
        tot += cin * cout * hw * (1 if name.endswith(".up") else k * k)
    return tot


def synthetic_weights(seed: int = 0, bias_std: float = 0.02) -> Dict[str, torch.Tensor]:
    """Seeded deploy-form weights: He-normal folded convs, small biases, cls_pred.bias = -log(99)
    (yolo_head_dfl_head.py:188-190).  Conditioning: the bottleneck shortcut scale is alpha = 0.5 and
    the second conv of every bottleneck uses gain 1 instead of 2, which keeps the second moment of
    the activations ~constant through the 40 residual bottlenecks (alpha = 1 with He gain grows it
    ~2.6x per bottleneck and overflows exp() in the scale head).  Same work, finite numbers."""
    g = torch.Generator().manual_seed(seed)
    w: Dict[str, torch.Tensor] = {}
    for name, k, s, cin, cout, _ in conv_layer_list():
        if name.endswith(".up"):
            w[name + ".w"] = torch.randn(cin, cout, 2, 2, generator=g) * math.sqrt(1.0 / cin)
        else:
            gain = 1.0 if name.endswith("_pred") or name.endswith(".out") or (name.endswith(".cv2") and ".csp.b" in name) else 2.0
            w[name + ".w"] = torch.randn(cout, cin, k, k, generator=g) * math.sqrt(gain / (cin * k * k))
        w[name + ".b"] = torch.randn(cout, generator=g) * bias_std
        if name.endswith(".cls_pred"):
            w[name + ".b"] = torch.full((cout,), -math.log(99.0))
        if name.endswith(".cv2") and ".b" in name:
            w[name[: -len(".cv2")] + ".alpha"] = torch.tensor(0.5)
    return w
