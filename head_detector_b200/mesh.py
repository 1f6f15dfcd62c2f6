"""Consumers of the decoded head meshes, on the device: the PNCC image of
`PredictionResult.get_pncc()` (reference: head_detector/pncc_processor.py:59-73 over the CPU rasteriser Sim3DR) and
`refined_head_bbox` (head_detector/utils.py:26-35).  The tables PNCCProcessor.__init__ rebuilds for every
PredictionResult in the reference (a Python filter over 9976 faces, ~0.09 s) are precomputed assets here and uploaded once."""
import os
from typing import Sequence

import numpy as np
import torch

from . import _lib

_ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets", "flame_indices.npz")
_tables = None
_dev = {}


def tables():
    """face / head_indices / head_w_ears vertex subsets, landmark triangles, PNCC triangles and NCC colours (numpy)."""
    global _tables
    if _tables is None:
        z = np.load(_ASSETS)
        _tables = {k: z[k] for k in z.files}
    return _tables


def _device_table(name: str, dtype) -> torch.Tensor:
    key = (name, torch.cuda.current_device())
    if key not in _dev:
        _dev[key] = torch.from_numpy(np.ascontiguousarray(tables()[name])).to(device="cuda", dtype=dtype).contiguous()
    return _dev[key]


def _require_cuda(what: str):
    if not torch.cuda.is_available():
        raise RuntimeError(f"head_detector_b200.{what} needs a CUDA device (no CPU fallback)")


def pncc_image(height: int, width: int, vertices: Sequence[np.ndarray]) -> np.ndarray:
    """PNCCProcessor.__call__: uint8 [H,W,3] with every head painted in order (later heads on top, z-buffer inside a head).
    `vertices`: per head [5023,3] image-space (HeadMetadata.vertices_3d); NOT modified (the reference flips z in place)."""
    _require_cuda("mesh.pncc_image")
    n = len(vertices)
    img = torch.zeros(height, width, 3, dtype=torch.uint8, device="cuda")
    if n == 0:
        return img.cpu().numpy()
    v = torch.from_numpy(np.ascontiguousarray(np.stack([np.asarray(x, dtype=np.float32) for x in vertices]))).cuda()
    return pncc_image_device(v, height, width, out=img).cpu().numpy()


def pncc_image_device(verts: torch.Tensor, height: int, width: int, out: torch.Tensor = None) -> torch.Tensor:
    """Device form: verts [n,5023,3] cuda fp32 (e.g. Engine.head_verts) -> uint8 cuda [H,W,3]."""
    _require_cuda("mesh.pncc_image_device")
    v = verts.detach().to(device="cuda", dtype=torch.float32).contiguous()
    tri, col = _device_table("pncc_triangles", torch.int32), _device_table("ncc_colors", torch.float32)
    img = out if out is not None else torch.zeros(height, width, 3, dtype=torch.uint8, device="cuda")
    keys = torch.empty(height * width, dtype=torch.int64, device="cuda")
    _lib.check(_lib.lib().vgh_pncc_render(v.data_ptr(), v.shape[0], tri.data_ptr(), tri.shape[0], col.data_ptr(), height, width,
                                          img.data_ptr(), keys.data_ptr(), _lib.stream_ptr()), "vgh_pncc_render")
    return img


def refined_head_bboxes(verts: torch.Tensor) -> torch.Tensor:
    """refined_head_bbox for n heads: verts [n,5023,3] cuda -> int32 cuda [n,4] = (x, y, w, h)."""
    _require_cuda("mesh.refined_head_bboxes")
    v = verts.detach().to(device="cuda", dtype=torch.float32).contiguous()
    idx = _device_table("head_indices", torch.int32)
    out = torch.empty(v.shape[0], 4, dtype=torch.int32, device="cuda")
    _lib.check(_lib.lib().vgh_head_bbox(v.data_ptr(), v.shape[0], idx.data_ptr(), idx.shape[0], out.data_ptr(), _lib.stream_ptr()), "vgh_head_bbox")
    return out
