"""Real-weight loader: the reference's checkpoint formats -> the deploy-form weight dict `arch.pack` takes.

The reference ships its network as a TorchScript trace `vgg_heads_l.trcd` (head_detector/detector.py:25-30,
`torch.jit.load`), produced by `ExportableMeshEstimationModel.export` (yolo_head_training/yolo_head/
exportable_mesh_model.py:439-442) from a super_gradients `CustomizableDetector` (backbone / neck / heads).  The trace keeps
the as-trained multi-branch blocks (`YoloHeads.prep_model_for_conversion` only caches anchors,
yolo_heads.py:136-144), so the loader re-parameterises here (`bn_eps` = 1e-6 from
yolo_heads_l_arch_params.yaml:139):

  QARepVGG   W = post_bn( bn(W3) + alpha * pad(W1) + [I] ),  b likewise        (`branch_3x3.conv/bn`, `branch_1x1`, `alpha`, `post_bn`)
  Conv-BN    W = bn(W)                                                          (`conv` + `bn`, or `seq.conv` + `seq.bn` in the heads)
  plain conv / ConvTranspose2d as they are                                      (`weight`, `bias`)

Key names: the head modules are the reference's own classes (yolo_head_dfl_head.py:72-126: `pose_stem`, `bbox_stem`,
`cls_convs.0`, `reg_convs.0`, `reg_pred`, `cls_pred`, `flame_{shape,expression,rotation,jaw,scale,translation}_pred.{0..3}`);
backbone / neck names follow super_gradients' YoloNAS modules (`backbone.stem.conv`, `backbone.stage{i}.downsample`,
`backbone.stage{i}.blocks.{conv1,conv2,conv3,bottlenecks.{j}.{cv1,cv2,alpha}}`, `backbone.context_module.{cv1,cv2}`,
`neck.neck{1,2}.{conv,upsample,reduce_skip1,reduce_skip2,downsample,reduce_after_concat,blocks}`,
`neck.neck{3,4}.{conv,blocks}`) - restated from the library ([3P-MEM], not verifiable offline); a checkpoint whose
names differ fails loudly with the list of missing / unexpected keys instead of loading garbage.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch

from . import arch

BN_EPS = arch.BN_EPS
_TOWER_ATTR = {"shape": "flame_shape_pred", "expr": "flame_expression_pred", "rot": "flame_rotation_pred", "jaw": "flame_jaw_pred",
               "scale": "flame_scale_pred", "transl": "flame_translation_pred"}


def layer_map() -> List[Tuple[str, str, str]]:
    """(our layer name, checkpoint module prefix, kind) for the 191 convs; kind in
    {"qarep", "qarep_res", "convbn", "seqconvbn", "conv", "convT"}."""
    out: List[Tuple[str, str, str]] = [("stem", "backbone.stem.conv", "qarep")]

    def csp(ours, theirs, n):
        for c in ("conv1", "conv2", "conv3"):
            out.append((f"{ours}.{c}", f"{theirs}.{c}", "convbn"))
        for j in range(n):
            out.append((f"{ours}.b{j}.cv1", f"{theirs}.bottlenecks.{j}.cv1", "qarep_res"))
            out.append((f"{ours}.b{j}.cv2", f"{theirs}.bottlenecks.{j}.cv2", "qarep_res"))

    for i, (_, n, _) in enumerate(arch.BACKBONE, start=1):
        out.append((f"stage{i}.down", f"backbone.stage{i}.downsample", "qarep"))
        csp(f"stage{i}.csp", f"backbone.stage{i}.blocks", n)
    out += [("spp.cv1", "backbone.context_module.cv1", "convbn"), ("spp.cv2", "backbone.context_module.cv2", "convbn")]
    for nk in ("neck1", "neck2"):
        p = f"neck.{nk}"
        out += [(f"{nk}.reduce", f"{p}.conv", "convbn"), (f"{nk}.up", f"{p}.upsample", "convT"), (f"{nk}.skip1", f"{p}.reduce_skip1", "convbn"),
                (f"{nk}.skip2_reduce", f"{p}.reduce_skip2", "convbn"), (f"{nk}.skip2_down", f"{p}.downsample", "convbn"),
                (f"{nk}.fuse", f"{p}.reduce_after_concat", "convbn")]
        csp(f"{nk}.csp", f"{p}.blocks", arch.NECKS[nk][1])
    for nk in ("neck3", "neck4"):
        out.append((f"{nk}.down", f"neck.{nk}.conv", "convbn"))
        csp(f"{nk}.csp", f"neck.{nk}.blocks", arch.NECKS[nk][1])
    for l in (1, 2, 3):
        h, p = f"head{l}", f"heads.head{l}"
        out += [(f"{h}.bbox_stem", f"{p}.bbox_stem", "seqconvbn"), (f"{h}.pose_stem", f"{p}.pose_stem", "seqconvbn"),
                (f"{h}.cls_conv", f"{p}.cls_convs.0", "seqconvbn"), (f"{h}.reg_conv", f"{p}.reg_convs.0", "seqconvbn"),
                (f"{h}.reg_pred", f"{p}.reg_pred", "conv"), (f"{h}.cls_pred", f"{p}.cls_pred", "conv")]
        for tower, _, _ in arch.TOWERS:
            for i in range(3):
                out.append((f"{h}.{tower}.{i}", f"{p}.{_TOWER_ATTR[tower]}.{i}", "qarep"))
            out.append((f"{h}.{tower}.out", f"{p}.{_TOWER_ATTR[tower]}.3", "conv"))
    return out


def bottleneck_alphas() -> List[Tuple[str, str]]:
    """(our `<csp>.b<j>.alpha`, checkpoint `...bottlenecks.<j>.alpha`): the shortcut scale of every YoloNAS bottleneck."""
    out = []
    for ours, theirs, kind in layer_map():
        if kind == "qarep_res" and ours.endswith(".cv2"):
            out.append((ours[:-4] + ".alpha", theirs[:-4] + ".alpha"))
    return out


def _bn(sd, p, eps):
    s = sd[p + ".weight"].double() / torch.sqrt(sd[p + ".running_var"].double() + eps)
    return s, sd[p + ".bias"].double() - sd[p + ".running_mean"].double() * s


class _Used(dict):
    """state_dict view that records which keys were read (for the unexpected-key report)."""

    def __init__(self, sd):
        super().__init__(sd)
        self.used = set()

    def __getitem__(self, k):
        self.used.add(k)
        return super().__getitem__(k)

    def has(self, k):
        return super().__contains__(k)


def _fold_qarep(sd: _Used, p: str, residual: bool, eps: float):
    if not sd.has(p + ".branch_3x3.conv.weight") and sd.has(p + ".rbr_reparam.weight"):
        # already fused by the exporter: one conv (+ the post-BN when only partially fused)
        w, b = sd[p + ".rbr_reparam.weight"].double(), sd[p + ".rbr_reparam.bias"].double()
        if sd.has(p + ".post_bn.weight"):
            s, t = _bn(sd, p + ".post_bn", eps)
            w, b = w * s[:, None, None, None], b * s + t
        return w.float(), b.float()
    s3, t3 = _bn(sd, p + ".branch_3x3.bn", eps)
    w = sd[p + ".branch_3x3.conv.weight"].double() * s3[:, None, None, None]
    alpha = sd[p + ".alpha"].double().reshape(()) if sd.has(p + ".alpha") else torch.tensor(1.0, dtype=torch.float64)
    w = w + alpha * torch.nn.functional.pad(sd[p + ".branch_1x1.weight"].double(), [1, 1, 1, 1])
    b = t3 + (alpha * sd[p + ".branch_1x1.bias"].double() if sd.has(p + ".branch_1x1.bias") else 0.0)
    if residual:
        if w.shape[0] != w.shape[1]:
            raise ValueError(f"{p}: identity branch on a non-square block {tuple(w.shape)}")
        idx = torch.arange(w.shape[0])
        w[idx, idx, 1, 1] += 1.0
    if sd.has(p + ".post_bn.weight"):
        sp, tp = _bn(sd, p + ".post_bn", eps)
        w, b = w * sp[:, None, None, None], b * sp + tp
    return w.float(), b.float()


def _strip_prefix(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """The trace wraps the detector (`model.` inside ConvertableCompletePipelineModel, exportable_mesh_model.py:421-427);
    checkpoints may carry `net.` / `ema_net.` / `module.`.  Everything before the first `backbone.` / `neck.` / `heads.` goes."""
    out = {}
    for k, v in sd.items():
        for root in ("backbone.", "neck.", "heads."):
            i = k.find(root)
            if i == 0 or (i > 0 and k[i - 1] == "."):
                out[k[i:]] = v
                break
    return out


_IGNORED_SUFFIXES = (".num_batches_tracked", ".rbr_reparam.weight", ".rbr_reparam.bias")
_IGNORED_KEYS = ("heads.proj_conv", "heads.anchor_points", "heads.stride_tensor")


def deploy_from_state_dict(state_dict: Dict[str, torch.Tensor], eps: float = BN_EPS, strict: bool = True) -> Dict[str, torch.Tensor]:
    """state_dict of the reference's YoloHeads_L (any wrapper prefix) -> {"<layer>.w", "<layer>.b", "<csp>.b<j>.alpha"}.
    Raises KeyError listing missing / unexpected keys (strict) and ValueError on a shape mismatch."""
    sd = _Used(_strip_prefix(state_dict))
    if not sd:
        raise KeyError("no `backbone.` / `neck.` / `heads.` entries in the state_dict - not a YoloHeads checkpoint")
    shapes = {name: ((cin, cout, 2, 2) if tr else (cout, cin, k, k)) for name, k, cin, cout, tr in arch.conv_names()}
    out: Dict[str, torch.Tensor] = {}
    missing: List[str] = []
    for ours, p, kind in layer_map():
        try:
            if kind in ("qarep", "qarep_res"):
                w, b = _fold_qarep(sd, p, kind == "qarep_res", eps)
            elif kind in ("convbn", "seqconvbn"):
                q = p + (".seq" if kind == "seqconvbn" else "")
                s, t = _bn(sd, q + ".bn", eps)
                w, b = (sd[q + ".conv.weight"].double() * s[:, None, None, None]).float(), t.float()
            else:
                w, b = sd[p + ".weight"].float(), sd[p + ".bias"].float()
        except KeyError as ex:
            missing.append(str(ex.args[0]))
            continue
        if tuple(w.shape) != shapes[ours]:
            raise ValueError(f"{p}: weight shape {tuple(w.shape)} != {shapes[ours]} expected for layer {ours!r}")
        out[ours + ".w"], out[ours + ".b"] = w.contiguous(), b.contiguous()
    for ours, theirs in bottleneck_alphas():
        out[ours] = sd[theirs].float().reshape(()) if sd.has(theirs) else torch.tensor(1.0)
    if missing:
        raise KeyError(f"checkpoint lacks {len(missing)} tensors of YoloHeads_L, e.g. {missing[:6]}")
    if strict:
        extra = [k for k in sd.keys() if k not in sd.used and not k.endswith(_IGNORED_SUFFIXES) and k not in _IGNORED_KEYS]
        if extra:
            raise KeyError(f"checkpoint has {len(extra)} tensors this architecture does not use, e.g. {extra[:6]}")
    return out


def load_checkpoint(path: str, strict: bool = True) -> Dict[str, torch.Tensor]:
    """`vgg_heads_l.trcd` (TorchScript, what the reference downloads), a torch-saved state_dict / checkpoint
    ({"net": ...} / {"ema_net": ...} / {"state_dict": ...}), or a torch-saved deploy-form dict -> deploy-form dict."""
    try:
        sd = torch.jit.load(path, map_location="cpu").state_dict()
    except RuntimeError:
        obj = torch.load(path, map_location="cpu", weights_only=True)
        for key in ("ema_net", "net", "state_dict", "model"):
            if isinstance(obj, dict) and key in obj and isinstance(obj[key], dict):
                obj = obj[key]
                break
        if isinstance(obj, dict) and "stem.w" in obj:   # already deploy form
            return {k: v for k, v in obj.items()}
        sd = obj
    return deploy_from_state_dict(sd, strict=strict)


def resolve(weights, model: str = "vgg_heads_l") -> Dict[str, torch.Tensor]:
    """`HeadDetector(weights=...)`: a path, a deploy-form dict, a state_dict, or the literal "synthetic"."""
    if weights is None:
        raise FileNotFoundError(
            f"no weights for {model!r}: the reference downloads vgg_heads_l.trcd from the HF hub (detector.py:25-30), which is "
            "unreachable offline.  Pass weights=<path to vgg_heads_l.trcd / a state_dict checkpoint>, set $VGGHEADS_B200_WEIGHTS, "
            "or ask for seeded random weights explicitly with weights='synthetic' (detections are then meaningless).")
    if isinstance(weights, str):
        if weights == "synthetic":
            return arch.synthetic_weights(0)
        return load_checkpoint(weights)
    if isinstance(weights, dict):
        if "stem.w" in weights:
            return weights
        return deploy_from_state_dict(weights)
    raise TypeError(f"weights must be a path, a dict or 'synthetic', got {type(weights).__name__}")
