"""Loads the reference's OWN head / decode / top-k modules (unmodified files under
/root/reference/yolo_head_training/yolo_head) with the import stand-ins of oracle/shims.  TEST INFRASTRUCTURE ONLY;
needs /root/reference, so it runs in the build container only (oracle/make_golden.py, tests that skip without it).

The package's own `__init__.py` imports the whole training stack (datasets, albumentations, losses ...), so a bare
namespace package named `yolo_head` is registered instead and the needed sub-modules are imported from their files."""
import importlib
import os
import sys
import types

REF_PKG = "/root/reference/yolo_head_training/yolo_head"
HERE = os.path.dirname(os.path.abspath(__file__))

# yolo_heads_l_arch_params.yaml:90-138 (heads section), as python data
HEADS_YAML = dict(
    num_classes=1, reg_max=16,
    heads_list=[
        {"YoloHeadsDFLHead": dict(bbox_inter_channels=bb, flame_inter_channels=256, flame_shape_out_channels=128, flame_expression_out_channels=64,
                                  flame_shape_inter_channels=256, flame_expression_inter_channels=128, flame_transformation_inter_channels=32,
                                  flame_regression_blocks=3, shared_stem=False, width_mult=1, first_conv_group_size=0, stride=s, reg_max=16)}
        for bb, s in ((128, 8), (256, 16), (512, 32))],
)
HEAD_IN_CHANNELS = (96, 192, 384)


def available() -> bool:
    return os.path.isdir(REF_PKG)


def load():
    """-> module namespace with YoloHeadsDFLHead, YoloHeadsNDFLHeads, VGGHeadDecodingModule (the reference classes)."""
    shims = os.path.join(HERE, "shims")
    if shims not in sys.path:
        sys.path.insert(0, shims)
    import super_gradients  # noqa: F401  (installs the placeholder finder)

    if "yolo_head" not in sys.modules:
        pkg = types.ModuleType("yolo_head")
        pkg.__path__ = [REF_PKG]
        sys.modules["yolo_head"] = pkg
    ns = types.SimpleNamespace()
    ns.YoloHeadsDFLHead = importlib.import_module("yolo_head.yolo_head_dfl_head").YoloHeadsDFLHead
    ns.YoloHeadsNDFLHeads = importlib.import_module("yolo_head.yolo_head_ndfl_heads").YoloHeadsNDFLHeads
    ns.VGGHeadDecodingModule = importlib.import_module("yolo_head.yolo_heads").VGGHeadDecodingModule
    ns.flame = importlib.import_module("yolo_head.flame")
    return ns


def build_heads(ns=None):
    """The reference's YoloHeadsNDFLHeads for YoloHeads_L (yaml:90-138), eval mode, bn eps 1e-6 (yaml:139)."""
    import copy

    import torch

    ns = ns or load()
    heads = ns.YoloHeadsNDFLHeads(num_classes=1, in_channels=HEAD_IN_CHANNELS, heads_list=copy.deepcopy(HEADS_YAML["heads_list"]), reg_max=16)
    for m in heads.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.eps = 1e-6   # CustomizableDetector applies `bn_eps` to every BatchNorm of the model
    return heads.eval()
