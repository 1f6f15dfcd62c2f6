"""FLAME decode on the B200: drop-in for the reference's head_detector/flame.py.

`FLAMELayer` keeps the reference's buffer names (flame.py:43-95) but its forward pass and
`reproject_spatial_vertices` (flame.py:179-208) run the fused CUDA kernel behind
`vgh_flame_decode` - there is no torch/CPU fallback."""
import ctypes as C
import os
from typing import Optional, Tuple

import numpy as np
import torch
from torch import Tensor

from . import _lib
from .head_info import FLAME_CONSTS, FlameParams

MESH_OFFSET_Z = 0.05
_ASSET = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets", "flame_generic.npz")


class FLAMELayer(torch.nn.Module):
    def __init__(self, consts=None, batch_size: int = 1, flame_path: Optional[str] = None) -> None:
        super().__init__()
        self.flame_constants = FLAME_CONSTS if consts is None else consts
        z = np.load(flame_path or _ASSET)
        self.register_buffer("faces_tensor", torch.from_numpy(z["faces"].astype(np.int64)))
        for name in ("v_template", "shapedirs", "J_regressor", "posedirs", "lbs_weights"):
            self.register_buffer(name, torch.from_numpy(np.ascontiguousarray(z[name], dtype=np.float32)))
        self.register_buffer("parents", torch.from_numpy(z["parents"]))
        self._handle = None

    # -- native handle (created on first use; needs a CUDA device)
    def handle(self):
        if self._handle is None:
            if not torch.cuda.is_available():
                raise RuntimeError("head_detector_b200.FLAMELayer needs a CUDA device (no CPU fallback)")
            h = C.c_void_p()
            arrs = [self.v_template, self.shapedirs, self.posedirs, self.J_regressor, self.lbs_weights]
            host = [a.detach().cpu().contiguous() for a in arrs]
            _lib.check(_lib.lib().vgh_flame_create(*[C.c_void_p(a.data_ptr()) for a in host], C.byref(h)), "vgh_flame_create")
            self._handle = h
        return self._handle

    def __del__(self):
        if getattr(self, "_handle", None) is not None and _lib._lib is not None:
            _lib.lib().vgh_flame_destroy(self._handle)
            self._handle = None

    def decode(self, flame_params: Tensor, xform: Optional[Tensor] = None, live=(300, 100)) -> Tuple[Tensor, Tensor, Tensor]:
        """[N,413] (cuda fp32) -> (model-space vertices [N,5023,3], R [N,3,3], projected [N,5023,3])."""
        p = flame_params.detach().to(device="cuda", dtype=torch.float32).contiguous()
        n = p.shape[0]
        verts = torch.empty(n, _lib.NUM_VERTS, 3, device="cuda")
        rot = torch.empty(n, 3, 3, device="cuda")
        proj = torch.empty(n, _lib.NUM_VERTS, 3, device="cuda")
        xf = None if xform is None else xform.to(device="cuda", dtype=torch.float32).contiguous()
        _lib.check(_lib.lib().vgh_flame_decode(self.handle(), p.data_ptr(), n, int(live[0]), int(live[1]),
                                               None if xf is None else xf.data_ptr(), verts.data_ptr(), rot.data_ptr(),
                                               proj.data_ptr(), _lib.stream_ptr()), "vgh_flame_decode")
        return verts, rot, proj

    def forward(self, flame_params: FlameParams, zero_rot: bool = False, zero_jaw: bool = False) -> Tensor:
        """flame.py:122-169.  zero_rot=True returns model-space vertices (incl. +0.05 z);
        otherwise the 6D rotation is applied (scale 1, translation 0)."""
        n = flame_params.shape.shape[0]
        p = torch.zeros(n, _lib.NUM_PARAMS, device="cuda")
        p[:, 0:flame_params.shape.shape[1]] = flame_params.shape
        p[:, 300:300 + flame_params.expression.shape[1]] = flame_params.expression
        if not zero_jaw and flame_params.jaw.shape[1] == 3:
            p[:, 400:403] = flame_params.jaw
        p[:, 403:409] = flame_params.rotation
        p[:, 412] = 1.0
        verts, _, proj = self.decode(p)
        return verts if zero_rot else proj


def reproject_spatial_vertices(flame: FLAMELayer, flame_params: Tensor, to_2d: bool = True, subset_indexes=None):
    """Same signature and return as the reference (flame.py:179-208):
    (vertices [N,5023,3], rotation_mat [N,3,3], projected [N,5023,2|3])."""
    shape = flame_params.size()
    if flame_params.size(0) == 0:
        dev = flame_params.device
        return (torch.zeros((0, _lib.NUM_VERTS, 3), device=dev),
                torch.eye(3, device=dev).unsqueeze(0).expand(0, 3, 3),
                torch.zeros((0, _lib.NUM_VERTS, 2 if to_2d else 3), device=dev))
    if flame_params.size(-1) != _lib.NUM_PARAMS:
        raise ValueError(f"Invalid number of parameters. Expected: {_lib.NUM_PARAMS}. Got: {flame_params.size(-1)}.")
    vertices, rotation_mat, projected = flame.decode(flame_params.reshape(-1, _lib.NUM_PARAMS))
    if subset_indexes is not None:
        projected = projected[:, subset_indexes]
    if to_2d:
        projected = projected[..., :2]
    projected = projected.view(*shape[:-1], *projected.size()[-2:]).contiguous()
    return vertices, rotation_mat, projected
