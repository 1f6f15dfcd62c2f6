"""The AS-TRAINED (unfused) YoloHeads_L with the module / parameter names of the reference's checkpoint, as plain torch.
TEST INFRASTRUCTURE ONLY.

Why: the released model is a `torch.jit.trace` of a super_gradients `CustomizableDetector` (head_detector/detector.py:25-30,
yolo_head_training/yolo_head/exportable_mesh_model.py:439-442).  Neither the blob nor super_gradients is reachable
offline, so this module restates the detector in its training-time form - multi-branch QARepVGG blocks, separate
BatchNorms, the reference's head / decode code - with the SAME attribute names, so that

  * `torch.jit.trace(YoloHeadsL(...))` is a synthetic `vgg_heads_l.trcd` the UNMODIFIED reference `HeadDetector`
    runs on (oracle/make_golden.py, CPU reference timing), and
  * `state_dict()` has the key names head_detector_b200/weights.py maps from (f3 loader tests).

Sources: widths / depths yolo_heads_l_arch_params.yaml:1-141; heads yolo_head_dfl_head.py:23-186 and
yolo_head_ndfl_heads.py:52-235 (restated here as `DFLHead` / `NDFLHeads` with identical attribute names and checked
output-for-output against the reference classes in tests/test_oracle_sg.py when /root/reference is present);
backbone / neck: super_gradients' YoloNAS modules ([3P-MEM]) built on oracle/shims/super_gradients/_impl.py.
"""
from __future__ import annotations

import math
import os
import sys
import torch
import torch.nn.functional as F
from torch import nn

_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")
if _SHIMS not in sys.path:
    sys.path.insert(0, _SHIMS)
from super_gradients._impl import ConvBNReLU, QARepVGGBlock, Residual  # noqa: E402

BN_EPS = 1e-6
BACKBONE = [(96, 2, 96), (192, 3, 128), (384, 5, 256), (768, 2, 512)]
NECKS = {"neck1": (192, 4, 128), "neck2": (96, 4, 128), "neck3": (192, 4, 128), "neck4": (384, 4, 256)}
HEADS = [(96, 128, 8), (192, 256, 16), (384, 512, 32)]


class Conv(nn.Module):
    """YoloNAS `Conv`: conv(bias=False) + bn + ReLU."""

    def __init__(self, cin, cout, k, stride=1):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, stride, k // 2, bias=False)
        self.bn = nn.BatchNorm2d(cout)
        self.act = nn.ReLU()

    def forward(self, x):
        return self.act(self.bn(self.conv(x)))


class Bottleneck(nn.Module):
    """YoloNASBottleneck: alpha * x + cv2(cv1(x))."""

    def __init__(self, c):
        super().__init__()
        self.cv1 = QARepVGGBlock(c, c)
        self.cv2 = QARepVGGBlock(c, c)
        self.shortcut = Residual()
        self.alpha = nn.Parameter(torch.tensor([1.0]))

    def forward(self, x):
        return self.alpha * self.shortcut(x) + self.cv2(self.cv1(x))


class CSPLayer(nn.Module):
    def __init__(self, cin, cout, n, hidden, concat_intermediates):
        super().__init__()
        self.conv1 = Conv(cin, hidden, 1)
        self.conv2 = Conv(cin, hidden, 1)
        self.conv3 = Conv(hidden * (2 + (n if concat_intermediates else 0)), cout, 1)
        self.bottlenecks = nn.Sequential(*[Bottleneck(hidden) for _ in range(n)])
        self.concat_intermediates = concat_intermediates

    def forward(self, x):
        x1 = self.conv1(x)
        outs = [x1]
        for b in self.bottlenecks:
            outs.append(b(outs[-1]))
        x2 = self.conv2(x)
        keep = outs if self.concat_intermediates else outs[-1:]
        return self.conv3(torch.cat((*keep, x2), dim=1))


class Stem(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = QARepVGGBlock(cin, cout, stride=2, use_residual_connection=False)

    def forward(self, x):
        return self.conv(x)


class Stage(nn.Module):
    def __init__(self, cin, cout, n, hidden):
        super().__init__()
        self.downsample = QARepVGGBlock(cin, cout, stride=2, use_residual_connection=False)
        self.blocks = CSPLayer(cout, cout, n, hidden, True)

    def forward(self, x):
        return self.blocks(self.downsample(x))


class SPP(nn.Module):
    def __init__(self, cin, cout, k=(5, 9, 13)):
        super().__init__()
        self.cv1 = Conv(cin, cin // 2, 1)
        self.cv2 = Conv(cin // 2 * (len(k) + 1), cout, 1)
        self.m = nn.ModuleList([nn.MaxPool2d(kernel_size=x, stride=1, padding=x // 2) for x in k])

    def forward(self, x):
        x = self.cv1(x)
        return self.cv2(torch.cat([x] + [m(x) for m in self.m], 1))


class Backbone(nn.Module):
    def __init__(self):
        super().__init__()
        self.stem = Stem(3, 48)
        cin = 48
        for i, (cout, n, hid) in enumerate(BACKBONE, start=1):
            setattr(self, f"stage{i}", Stage(cin, cout, n, hid))
            cin = cout
        self.context_module = SPP(768, 768)

    def forward(self, x):
        x = self.stem(x)
        c2 = self.stage1(x)
        c3 = self.stage2(c2)
        c4 = self.stage3(c3)
        c5 = self.context_module(self.stage4(c4))
        return c2, c3, c4, c5


class UpStage(nn.Module):
    def __init__(self, c_low, c_skip1, c_skip2, out, n, hidden):
        super().__init__()
        self.reduce_skip1 = Conv(c_skip1, out, 1)
        self.reduce_skip2 = Conv(c_skip2, out, 1)
        self.conv = Conv(c_low, out, 1)
        self.upsample = nn.ConvTranspose2d(out, out, kernel_size=2, stride=2)
        self.downsample = Conv(out, out, 3, 2)
        self.reduce_after_concat = Conv(3 * out, out, 1)
        self.blocks = CSPLayer(out, out, n, hidden, False)

    def forward(self, x, skip1, skip2):
        skip1 = self.reduce_skip1(skip1)
        skip2 = self.reduce_skip2(skip2)
        x_inter = self.conv(x)
        x = self.upsample(x_inter)
        skip2 = self.downsample(skip2)
        x = self.reduce_after_concat(torch.cat([x, skip1, skip2], 1))
        return x_inter, self.blocks(x)


class DownStage(nn.Module):
    def __init__(self, c, c_skip, out, n, hidden):
        super().__init__()
        self.conv = Conv(c, out // 2, 3, 2)
        self.blocks = CSPLayer(out // 2 + c_skip, out, n, hidden, False)

    def forward(self, x, skip):
        return self.blocks(torch.cat([self.conv(x), skip], 1))


class Neck(nn.Module):
    def __init__(self):
        super().__init__()
        self.neck1 = UpStage(768, 384, 192, *NECKS["neck1"])
        self.neck2 = UpStage(192, 192, 96, *NECKS["neck2"])
        self.neck3 = DownStage(96, 96, *NECKS["neck3"])
        self.neck4 = DownStage(192, 192, *NECKS["neck4"])

    def forward(self, feats):
        c2, c3, c4, c5 = feats
        i1, x = self.neck1(c5, c4, c3)
        i2, p3 = self.neck2(x, c3, c2)
        p4 = self.neck3(p3, i2)
        p5 = self.neck4(p4, i1)
        return p3, p4, p5


# ------------------------------------------------------------------------------------------ heads (restated; names of the reference classes)
class DFLHead(nn.Module):
    """yolo_head_dfl_head.py:23-186 for shared_stem=False, first_conv_group_size=0 (yaml:96-138)."""

    def __init__(self, cin, bbox_inter, stride, flame_inter=256, shape_inter=256, expr_inter=128, transf_inter=32, shape_out=128, expr_out=64, reg_max=16):
        super().__init__()
        self.pose_stem = ConvBNReLU(cin, flame_inter, kernel_size=1, stride=1, padding=0, bias=False)
        self.bbox_stem = ConvBNReLU(cin, bbox_inter, kernel_size=1, stride=1, padding=0, bias=False)
        self.cls_convs = nn.Sequential(ConvBNReLU(bbox_inter, bbox_inter, kernel_size=3, stride=1, padding=1, bias=False))
        self.reg_convs = nn.Sequential(ConvBNReLU(bbox_inter, bbox_inter, kernel_size=3, stride=1, padding=1, bias=False))
        self.reg_pred = nn.Conv2d(bbox_inter, 4 * (reg_max + 1), 1, 1, 0)
        self.cls_pred = nn.Conv2d(bbox_inter, 1, 1, 1, 0)

        def tower(inter, out):
            layers, c = [], flame_inter
            for _ in range(3):
                layers.append(QARepVGGBlock(c, inter, use_residual_connection=False, use_alpha=True))
                c = inter
            layers.append(nn.Conv2d(inter, out, 1, 1, 0))
            return nn.Sequential(*layers)

        self.flame_shape_pred = tower(shape_inter, shape_out)
        self.flame_expression_pred = tower(expr_inter, expr_out)
        self.flame_rotation_pred = tower(transf_inter, 6)
        self.flame_jaw_pred = tower(transf_inter, 3)
        self.flame_scale_pred = tower(transf_inter, 1)
        self.flame_translation_pred = tower(transf_inter, 3)
        self.stride = stride
        torch.nn.init.constant_(self.cls_pred.bias, -math.log((1 - 1e-2) / 1e-2))

    def towers_raw(self, pose):
        """pre-activation tower outputs (what the CUDA path stores as its raw FLAME rows)."""
        return {"shape": self.flame_shape_pred(pose), "expr": self.flame_expression_pred(pose), "rot": self.flame_rotation_pred(pose),
                "jaw": self.flame_jaw_pred(pose), "transl": self.flame_translation_pred(pose), "scale": self.flame_scale_pred(pose)}

    def forward(self, x):
        pose, bbox = self.pose_stem(x), self.bbox_stem(x)
        cls_output = self.cls_pred(self.cls_convs(bbox))
        reg_output = self.reg_pred(self.reg_convs(bbox))
        t = self.towers_raw(pose)
        shape = F.pad(t["shape"].tanh() * 3, (0, 0, 0, 0, 0, 300 - t["shape"].size(1)))
        expr = F.pad(t["expr"].tanh() * 3, (0, 0, 0, 0, 0, 100 - t["expr"].size(1)))
        flame = torch.cat([shape, expr, t["rot"], t["jaw"], t["transl"], t["scale"].exp() / 0.05], dim=1)
        return reg_output, cls_output, flame


class NDFLHeads(nn.Module):
    """yolo_head_ndfl_heads.py:117-175 (the tracing branch: returns boxes, scores, flame)."""

    def __init__(self, reg_max=16):
        super().__init__()
        self.reg_max = reg_max
        self.register_buffer("proj_conv", torch.linspace(0, reg_max, reg_max + 1).reshape([1, reg_max + 1, 1, 1]), persistent=False)
        for l, (cin, bb, stride) in enumerate(HEADS, start=1):
            setattr(self, f"head{l}", DFLHead(cin, bb, stride))
        self.fpn_strides = tuple(s for _, _, s in HEADS)

    def forward(self, feats):
        cls_list, reg_list, flame_list, pts, strides = [], [], [], [], []
        for i, feat in enumerate(feats):
            b, _, h, w = feat.shape
            reg, cls, flame = getattr(self, f"head{i + 1}")(feat)
            r = torch.permute(reg.reshape([-1, 4, self.reg_max + 1, h * w]), [0, 2, 3, 1])
            reg_list.append(F.softmax(r, dim=1).mul(self.proj_conv).sum(1))
            cls_list.append(cls.reshape([b, -1, h * w]))
            flame_list.append(flame.flatten(2))
            sy, sx = torch.meshgrid(torch.arange(end=h) + 0.5, torch.arange(end=w) + 0.5, indexing="ij")
            pts.append(torch.stack([sx, sy], dim=-1).to(feat.dtype).reshape([-1, 2]))
            strides.append(torch.full([h * w, 1], self.fpn_strides[i], dtype=feat.dtype))
        pts, strides = torch.cat(pts), torch.cat(strides)
        scores = torch.permute(torch.cat(cls_list, dim=-1), [0, 2, 1]).sigmoid()
        d = torch.cat(reg_list, dim=1)
        boxes = torch.cat([pts - d[..., :2], pts + d[..., 2:]], dim=-1) * strides
        c = torch.cat(flame_list, dim=-1)                      # [B, 413, A] in head order
        # FlameParams.from_3dmm reads [shape|expr|jaw3|rot6|transl|scale]; to_3dmm_tensor writes [..|rot6|jaw3|..]
        jaw, rot = c[:, 400:403], c[:, 403:409]
        transl = c[:, 409:412].clone()
        transl[:, 0:2] = transl[:, 0:2] + (pts * strides).T[None]
        scale = c[:, 412:413] * strides[None, None, :, 0]
        flame = torch.cat([c[:, :400], rot, jaw, transl, scale], dim=1)
        return boxes, scores, flame.permute(0, 2, 1)


class YoloHeadsL(nn.Module):
    """`CustomizableDetector` layout: .backbone / .neck / .heads.  forward(x [B,3,S,S] in [0,1]) -> (boxes, scores, flame)."""

    def __init__(self, heads: nn.Module = None):
        super().__init__()
        self.backbone = Backbone()
        self.neck = Neck()
        self.heads = heads if heads is not None else NDFLHeads()
        for m in self.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.eps = BN_EPS   # `bn_eps` yaml:139, applied to every BatchNorm by CustomizableDetector

    def features(self, x):
        return self.neck(self.backbone(x))

    def forward(self, x):
        return self.heads(self.features(x))


class Traceable(nn.Module):
    """What the reference exports: `ConvertableCompletePipelineModel(model, pre, post)` keeps the detector under `.model`."""

    def __init__(self, model):
        super().__init__()
        self.model = model

    def forward(self, x):
        return self.model(x)


# ------------------------------------------------------------------------------------------ seeded parameters
@torch.no_grad()
def seed_parameters_(net: nn.Module, seed: int = 0, alpha: float = 0.5) -> nn.Module:
    """Deterministic, well-conditioned parameters: He-normal convs, BatchNorm affine / statistics drawn near identity,
    bottleneck shortcut scale `alpha`, tower alphas 1, cls_pred.bias = -log(99) (yolo_head_dfl_head.py:188-190).
    Pure functions of the torch CPU generator: reproducible on every machine."""
    g = torch.Generator().manual_seed(seed)
    for name, m in net.named_modules():
        if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
            fan_in = m.weight.shape[1] * m.weight.shape[2] * m.weight.shape[3] if isinstance(m, nn.Conv2d) else m.weight.shape[0]
            linear = name.endswith(("reg_pred", "cls_pred", "_pred.3", "upsample", "branch_1x1"))
            gain = 1.0 if linear else 2.0
            if "bottlenecks" in name and ".cv2." in name:
                gain = 0.5
            if name.endswith("branch_1x1"):
                gain = 0.25
            m.weight.copy_(torch.randn(m.weight.shape, generator=g) * math.sqrt(gain / fan_in))
            if m.bias is not None:
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.02)
            if name.endswith("cls_pred"):
                m.bias.fill_(-math.log(99.0))
        elif isinstance(m, nn.BatchNorm2d):
            m.weight.copy_(0.8 + 0.4 * torch.rand(m.weight.shape, generator=g))
            m.bias.copy_(0.05 * torch.randn(m.bias.shape, generator=g))
            m.running_mean.copy_(0.05 * torch.randn(m.bias.shape, generator=g))
            m.running_var.copy_(0.8 + 0.4 * torch.rand(m.bias.shape, generator=g))
            if name.endswith("post_bn") and "bottlenecks" in name:
                m.weight.mul_(0.6)   # the identity branch of the inner blocks adds the input's power; keep the chain's second moment ~flat
    for name, p in net.named_parameters():
        if name.endswith("alpha"):
            p.fill_(alpha if "bottlenecks" in name and name.count("cv") == 0 else 1.0)
    return net


def build(seed: int = 0, heads: nn.Module = None) -> YoloHeadsL:
    return seed_parameters_(YoloHeadsL(heads), seed).eval()


def trace_to(path: str, net: YoloHeadsL, image_size: int = 640) -> str:
    """Synthetic `vgg_heads_l.trcd`: torch.jit.trace of the detector wrapped like the reference's export pipeline."""
    wrapped = Traceable(net).eval()
    with torch.no_grad():
        ts = torch.jit.trace(wrapped, torch.zeros(1, 3, image_size, image_size), check_trace=False)
    ts.save(path)
    return path
