"""CPU oracle for confidence-select + top-k + NMS.  TEST INFRASTRUCTURE ONLY.

Restates /root/reference/head_detector/utils.py:159-194 (`nms`) and its batched twin
/root/reference/yolo_head_training/yolo_head/yolo_heads_post_prediction_callback.py:55-97,
returning the ordered ORIGINAL anchor ids that survive (SURVEY.md section 7 gotcha 3) instead
of gathered rows.  `torchvision.ops.nms` is third-party arithmetic (torchvision~=0.15.2,
requirements.txt:2): greedy suppression in descending score order, suppress iff
IoU > threshold, IoU = inter / (area_i + area_j - inter) with every fp32 op rounded
separately.  That published algorithm is restated below in numpy float32 (numpy never
contracts to FMA), and `tests/test_oracle_nms.py` pins it against torchvision itself and
against tests/golden/nms_ref_cases.npz (outputs of the unmodified reference run here).

Tie rule: equal scores are ordered by lower anchor id first (stable descending sort); the
reference leaves ties implementation-defined, the synthetic inputs are tie-free.
"""
from __future__ import annotations

import numpy as np


def select_candidates(scores: np.ndarray, conf_thr: float, top_k: int) -> np.ndarray:
    """Anchor ids with score >= thr; if more than top_k, the top_k best; score-descending."""
    scores = np.asarray(scores, dtype=np.float32).reshape(-1)
    ids = np.nonzero(scores >= np.float32(conf_thr))[0]
    order = np.argsort(-scores[ids], kind="stable")  # stable => lower id first on ties
    ids = ids[order]
    return ids[:top_k] if ids.size > top_k else ids


def greedy_nms(boxes: np.ndarray, iou_thr: float) -> np.ndarray:
    """Greedy NMS over boxes ALREADY in descending-score order; returns kept positions."""
    b = np.asarray(boxes, dtype=np.float32)
    n = b.shape[0]
    if n == 0:
        return np.zeros(0, dtype=np.int64)
    x1, y1, x2, y2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    areas = (x2 - x1) * (y2 - y1)
    dead = np.zeros(n, dtype=bool)
    keep = []
    thr = np.float32(iou_thr)
    zero = np.float32(0)
    for i in range(n):
        if dead[i]:
            continue
        keep.append(i)
        if i + 1 == n:
            break
        xx1 = np.maximum(x1[i], x1[i + 1:])
        yy1 = np.maximum(y1[i], y1[i + 1:])
        xx2 = np.minimum(x2[i], x2[i + 1:])
        yy2 = np.minimum(y2[i], y2[i + 1:])
        w = np.maximum(zero, xx2 - xx1)
        h = np.maximum(zero, yy2 - yy1)
        inter = w * h
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = inter / (areas[i] + areas[i + 1:] - inter)
        dead[i + 1:] |= ovr > thr
    return np.asarray(keep, dtype=np.int64)


def select_nms(boxes, scores, conf_thr=0.5, iou_thr=0.5, top_k=1000, keep_top_k=100) -> np.ndarray:
    """One image: ordered surviving anchor ids (<= keep_top_k)."""
    boxes = np.asarray(boxes, dtype=np.float32)
    cand = select_candidates(scores, conf_thr, top_k)
    kept = greedy_nms(boxes[cand], iou_thr)
    return cand[kept][:keep_top_k]


def select_nms_batch(boxes, scores, **kw):
    """Batched semantics of YoloHeadsPostPredictionCallback: every image independently."""
    return [select_nms(b, s, **kw) for b, s in zip(boxes, scores)]
