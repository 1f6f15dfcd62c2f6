"""The few super_gradients definitions the reference's head / decode code actually EXECUTES, restated from the
library's published module definitions ([3P-MEM]).  TEST INFRASTRUCTURE ONLY.

Parameter / sub-module names follow the library so that state_dict keys of the reference modules have their real names
(`seq.conv.weight`, `seq.bn.*`; `branch_3x3.conv.weight`, `branch_3x3.bn.*`, `branch_1x1.weight/.bias`, `alpha`,
`post_bn.*`) - the names head_detector_b200/weights.py maps from."""
import math
from typing import List, Tuple, Union

import torch
from torch import nn

_REGISTRY = {}


def _register(name=None):
    def deco(cls):
        _REGISTRY[name or cls.__name__] = cls
        return cls
    return deco


register_detection_module = _register
register_model = _register


def width_multiplier(original, factor, divisor=None):
    if divisor is None:
        return int(original * factor)
    return math.ceil(int(original * factor) / divisor) * divisor


class HpmStruct:
    def __init__(self, **entries):
        self.__dict__.update(entries)

    def to_dict(self):
        return dict(self.__dict__)


def torch_version_is_greater_or_equal(major, minor):
    v = torch.__version__.split("+")[0].split(".")
    return (int(v[0]), int(v[1])) >= (major, minor)


def infer_model_device(model):
    try:
        return next(model.parameters()).device
    except StopIteration:
        return None


def infer_model_dtype(model):
    try:
        return next(model.parameters()).dtype
    except StopIteration:
        return None


class BaseDetectionModule(nn.Module):
    def __init__(self, in_channels):
        super().__init__()
        self.in_channels = in_channels


class SupportsReplaceNumClasses:
    pass


class SupportsInputShapeCheck:
    pass


class AbstractPoseEstimationPostPredictionCallback:
    pass


class CustomizableDetector(nn.Module):
    """Placeholder base: the reference's YoloHeads class statement needs a base to exist; it is never instantiated here
    (the backbone / neck come from the library's registry, which is not available)."""

    def __init__(self, *a, **k):
        super().__init__()


class ConvBNReLU(nn.Module):
    """super_gradients.modules.ConvBNReLU: `seq` = conv (+ bn) (+ ReLU)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 padding_mode="zeros", use_normalization=True, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True,
                 device=None, dtype=None, use_activation=True, inplace=False):
        super().__init__()
        self.seq = nn.Sequential()
        self.seq.add_module("conv", nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding,
                                              dilation=dilation, groups=groups, bias=bias, padding_mode=padding_mode))
        if use_normalization:
            self.seq.add_module("bn", nn.BatchNorm2d(out_channels, eps=eps, momentum=momentum, affine=affine, track_running_stats=track_running_stats))
        if use_activation:
            self.seq.add_module("act", nn.ReLU(inplace=inplace))

    def forward(self, x):
        return self.seq(x)


class Residual(nn.Module):
    def forward(self, x):
        return x


class QARepVGGBlock(nn.Module):
    """super_gradients.modules.QARepVGGBlock, training-time (multi-branch) form:
    y = act(post_bn(bn(conv3x3(x)) + alpha * (conv1x1(x) + b) + [x]))."""

    def __init__(self, in_channels, out_channels, stride=1, dilation=1, groups=1, activation_type=nn.ReLU, activation_kwargs=None,
                 se_type=nn.Identity, se_kwargs=None, build_residual_branches=True, use_residual_connection=True, use_alpha=False,
                 use_1x1_bias=True, use_post_bn=True):
        super().__init__()
        self.in_channels, self.out_channels, self.stride = in_channels, out_channels, stride
        self.nonlinearity = activation_type(**(activation_kwargs or {}))
        self.se = se_type(**(se_kwargs or {}))
        self.branch_3x3 = nn.Sequential()
        self.branch_3x3.add_module("conv", nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=stride, padding=dilation,
                                                     groups=groups, bias=False, dilation=dilation))
        self.branch_3x3.add_module("bn", nn.BatchNorm2d(num_features=out_channels))
        self.branch_1x1 = nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=stride, padding=0, groups=groups, bias=use_1x1_bias)
        if use_residual_connection:
            assert out_channels == in_channels and stride == 1
            self.identity = Residual()
        else:
            self.identity = None
        self.alpha = nn.Parameter(torch.tensor([1.0]), requires_grad=True) if use_alpha else 1.0
        self.post_bn = nn.BatchNorm2d(num_features=out_channels) if use_post_bn else nn.Identity()

    def forward(self, inputs):
        id_out = 0.0 if self.identity is None else self.identity(inputs)
        x = self.branch_3x3(inputs) + self.alpha * self.branch_1x1(inputs) + id_out
        return self.se(self.nonlinearity(self.post_bn(x)))


class DetectionModulesFactory:
    """`factory.get({TypeName: {kwargs}})` / `insert_module_param` as the reference's NDFL heads use them
    (yolo_head_ndfl_heads.py:86-95)."""

    @staticmethod
    def insert_module_param(conf, name, value):
        conf = {k: dict(v) for k, v in (conf.to_dict() if isinstance(conf, HpmStruct) else dict(conf)).items()}
        (type_name,) = conf.keys()
        conf[type_name][name] = value
        return conf

    def get(self, conf):
        conf = conf.to_dict() if isinstance(conf, HpmStruct) else dict(conf)
        (type_name, kwargs), = conf.items()
        return _REGISTRY[type_name](**kwargs)


def batch_distance2bbox(points, distance, max_shapes=None):
    """super_gradients.training.utils.bbox_utils.batch_distance2bbox: (l, t, r, b) distances -> xyxy."""
    lt, rb = torch.split(distance, 2, dim=-1)
    x1y1 = -lt + points
    x2y2 = rb + points
    return torch.cat([x1y1, x2y2], dim=-1)


def generate_anchors_for_grid_cell(feats, fpn_strides, grid_cell_size=5.0, grid_cell_offset=0.5, dtype=torch.float):
    """pp_yolo_head.generate_anchors_for_grid_cell (used by the reference only outside tracing, for training targets)."""
    anchors, anchor_points, num_anchors_list, stride_tensor = [], [], [], []
    device = feats[0].device
    for feat, stride in zip(feats, fpn_strides):
        _, _, h, w = feat.shape
        cell_half = grid_cell_size * stride * 0.5
        shift_x = (torch.arange(end=w) + grid_cell_offset) * stride
        shift_y = (torch.arange(end=h) + grid_cell_offset) * stride
        shift_y, shift_x = torch.meshgrid(shift_y, shift_x, indexing="ij")
        anchor = torch.stack([shift_x - cell_half, shift_y - cell_half, shift_x + cell_half, shift_y + cell_half], dim=-1).to(dtype=dtype)
        anchor_point = torch.stack([shift_x, shift_y], dim=-1).to(dtype=dtype)
        anchors.append(anchor.reshape([-1, 4]))
        anchor_points.append(anchor_point.reshape([-1, 2]))
        num_anchors_list.append(len(anchors[-1]))
        stride_tensor.append(torch.full([num_anchors_list[-1], 1], stride, dtype=dtype))
    return (torch.cat(anchors).to(device), torch.cat(anchor_points).to(device), num_anchors_list, torch.cat(stride_tensor).to(device))
