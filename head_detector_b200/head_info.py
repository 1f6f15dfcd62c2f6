"""Result / parameter containers of the public API - same names, fields and slicing as the
reference's head_detector/head_info.py:9-107 so that user code reading `.heads[i].bbox`,
`.flame_params.rotation`, `.head_pose.yaw` ... keeps working unchanged."""
from collections import namedtuple
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np
import torch
from torch import Tensor

Bbox = namedtuple("Bbox", ["x", "y", "w", "h"])
RPY = namedtuple("RPY", ["roll", "pitch", "yaw"])

# widths of the blocks of the 413-float vector, in the order `from_3dmm` reads them
FLAME_CONSTS = {"shape": 300, "expression": 100, "rotation": 6, "jaw": 3, "eyeballs": 0, "neck": 0, "translation": 3, "scale": 1}
_READ_ORDER = ("shape", "expression", "jaw", "rotation", "eyeballs", "neck", "translation", "scale")
_WRITE_ORDER = ("shape", "expression", "rotation", "jaw", "eyeballs", "neck", "translation", "scale")


@dataclass
class HeadMetadata:
    bbox: Bbox
    score: float
    flame_params: object
    vertices_3d: np.ndarray
    head_pose: RPY


@dataclass
class FlameParams:
    shape: Tensor
    expression: Tensor
    rotation: Tensor
    translation: Tensor
    scale: Tensor
    jaw: Tensor
    eyeballs: Tensor
    neck: Tensor

    @classmethod
    def from_3dmm(cls, tensor_3dmm: Tensor, constants: Optional[Dict[str, int]] = None, zero_expr: bool = False) -> "FlameParams":
        """tensor_3dmm [B, 413, ...] -> views, READ order [shape|expr|jaw|rot6d|eyes|neck|transl|scale]
        (reference head_info.py:45-89; note jaw precedes rotation when reading)."""
        constants = FLAME_CONSTS if constants is None else constants
        expected = sum(constants.values())
        if tensor_3dmm.size(1) != expected:
            raise ValueError(f"Invalid number of parameters. Expected: {expected}. Got: {tensor_3dmm.size(1)}.")
        fields, at = {}, 0
        for key in _READ_ORDER:
            fields[key] = tensor_3dmm[:, at: at + constants[key]]
            at += constants[key]
        if zero_expr:
            fields["expression"] = torch.zeros_like(fields["expression"])
        return cls(**fields)

    def to_3dmm_tensor(self) -> Tensor:
        """WRITE order puts rotation before jaw (reference head_info.py:91-107), so
        to_3dmm_tensor(from_3dmm(x)) rotates channels 400..408 - kept bug-compatible."""
        return torch.cat([getattr(self, key) for key in _WRITE_ORDER], dim=1)
