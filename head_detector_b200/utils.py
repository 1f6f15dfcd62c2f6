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


IMAGE_SIZE = 640


def extend_bbox(bbox, offset=0.1) -> np.ndarray:
    """utils.py:38-68: grow [x, y, w, h] by `offset` (one value, (w, h) or (left, right, top, bottom)) of its size per side."""
    x, y, w, h = bbox
    if isinstance(offset, tuple):
        left, right, top, bottom = offset if len(offset) == 4 else (offset[0], offset[0], offset[1], offset[1])
    else:
        left = right = top = bottom = offset
    return np.array([x - w * left, y - h * top, w * (1.0 + right + left), h * (1.0 + top + bottom)]).astype("int32")


def extend_to_rect(bbox) -> np.ndarray:
    """utils.py:71-78: the longer side on both axes, centred on the shorter one."""
    x, y, w, h = bbox
    if w > h:
        return np.array([x, y - (w - h) // 2, w, w])
    return np.array([x - (h - w) // 2, y, h, h])


def flame_params_skull_center(flame_params, image: np.ndarray) -> Tuple[int, int]:
    """utils.py:81-92 (letterbox geometry of detector.py:41-50 at the fixed IMAGE_SIZE, as the reference has it)."""
    h, w = image.shape[:2]
    scale = IMAGE_SIZE / max(h, w)
    new_h, new_w = (IMAGE_SIZE, int(w * IMAGE_SIZE / h)) if h > w else (int(h * IMAGE_SIZE / w), IMAGE_SIZE)
    c = (flame_params.translation / scale)[0].numpy()
    return int(c[0] - (IMAGE_SIZE - new_w)), int(c[1] - (IMAGE_SIZE - new_h))


def get_rotation_mat(img: np.ndarray, img_center, angle):
    """utils.py:95-108: rotation about `img_center` with the canvas grown to the rotated bounds."""
    import cv2

    height, width = img.shape[:2]
    m = cv2.getRotationMatrix2D(img_center, angle, 1.0)
    abs_cos, abs_sin = abs(m[0, 0]), abs(m[0, 1])
    bound_w, bound_h = int(height * abs_sin + width * abs_cos), int(height * abs_cos + width * abs_sin)
    m[0, 2] += bound_w / 2 - img_center[0]
    m[1, 2] += bound_h / 2 - img_center[1]
    return m, (bound_w, bound_h)


def vertically_align(img: np.ndarray, vertices: np.ndarray, flame_params, roll: float):
    """utils.py:111-119: rotate the frame so that the head stands upright; vertices follow."""
    import cv2

    m, bounds = get_rotation_mat(img, flame_params_skull_center(flame_params, img), roll)
    upright = cv2.warpAffine(img, m, bounds, flags=cv2.INTER_LINEAR)
    return upright, np.hstack([vertices[:, :2], np.ones((vertices.shape[0], 1))]) @ m.T


def refined_head_bbox(vertices: np.ndarray):
    """utils.py:26-35 (host form; the batched device form is mesh.refined_head_bboxes)."""
    from .head_info import Bbox
    from .mesh import tables

    pts = np.asarray(vertices)[tables()["head_indices"]]
    x, y, x1, y1 = (int(v) for v in (pts[:, 0].min(), pts[:, 1].min(), pts[:, 0].max(), pts[:, 1].max()))
    return Bbox(x=x, y=y, w=x1 - x, h=y1 - y)


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
