"""Post-processing entry points with the reference's signatures (head_detector/utils.py)."""
from typing import Tuple

import numpy as np
import torch
import torch.nn.functional as F
from scipy.spatial.transform import Rotation

from . import _lib
from .head_info import RPY


def select_nms_indices(boxes_xyxy: torch.Tensor, scores: torch.Tensor, confidence_threshold=0.5, iou_threshold=0.5,
                       top_k=1000, keep_top_k=100) -> Tuple[torch.Tensor, torch.Tensor]:
    """Batched: boxes [B,A,4], scores [B,A] (cuda fp32) -> (keep_idx [B,keep_top_k] int32 anchor ids,
    -1 padded; keep_cnt [B] int32).  One kernel launch, one CTA per image."""
    b = boxes_xyxy.detach().to(device="cuda", dtype=torch.float32).contiguous()
    s = scores.detach().to(device="cuda", dtype=torch.float32).reshape(b.shape[0], -1).contiguous()
    B, A = s.shape
    idx = torch.empty(B, keep_top_k, dtype=torch.int32, device="cuda")
    cnt = torch.empty(B, dtype=torch.int32, device="cuda")
    _lib.check(_lib.lib().vgh_select_nms(b.data_ptr(), s.data_ptr(), B, A, float(confidence_threshold), float(iou_threshold),
                                         int(top_k), int(keep_top_k), idx.data_ptr(), cnt.data_ptr(), None, None,
                                         _lib.stream_ptr()), "vgh_select_nms")
    return idx, cnt


def nms(boxes_xyxy, scores, flame_params, confidence_threshold: float = 0.5, iou_threshold: float = 0.5,
        top_k: int = 1000, keep_top_k: int = 100):
    """Drop-in for head_detector/utils.py:159-194: returns (boxes [n,4], scores [n], flame [n,413]) of the
    FIRST image only, exactly like the reference (its `return` sits inside the batch loop)."""
    idx, cnt = select_nms_indices(boxes_xyxy[:1], scores[:1], confidence_threshold, iou_threshold, top_k, keep_top_k)
    keep = idx[0, : int(cnt[0])].long()
    dev = keep.device
    return (boxes_xyxy[0].detach().float().to(dev)[keep], scores[0].detach().float().to(dev).reshape(-1)[keep],
            flame_params[0].detach().float().to(dev)[keep])


def rot_mat_from_6dof(v: torch.Tensor) -> torch.Tensor:
    """utils.py:120-128 (host-side helper for API objects; the decode kernel has its own copy)."""
    assert v.shape[-1] == 6
    v = v.view(-1, 6)
    vx, vy = v[..., :3].clone(), v[..., 3:].clone()
    b1 = F.normalize(vx, dim=-1)
    b3 = F.normalize(torch.cross(b1, vy, dim=-1), dim=-1)
    b2 = -torch.cross(b1, b3, dim=1)
    return torch.stack((b1, b2, b3), dim=-1)


def limit_angle(angle, pi=180.0):
    if angle < -pi:
        angle = angle + (-2 * (int(angle / pi) // 2)) * pi
    if angle > pi:
        angle = angle - (2 * ((int(angle / pi) + 1) // 2)) * pi
    return angle


def rpy_from_rotations(rot_mats: np.ndarray):
    """Vectorised utils.py:146-151 for [N,3,3] rotation matrices -> list of RPY."""
    if len(rot_mats) == 0:
        return []
    ang = Rotation.from_matrix(np.transpose(np.asarray(rot_mats, dtype=np.float64), (0, 2, 1))).as_euler("xyz", degrees=True)
    return [RPY(*map(limit_angle, (a[2], a[0] - 180, a[1]))) for a in ang]


def calculate_rpy(flame_params) -> RPY:
    return rpy_from_rotations(rot_mat_from_6dof(flame_params.rotation).numpy()[:1])[0]
