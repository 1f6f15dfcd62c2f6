"""Pins oracle/nms_oracle.py against torchvision.ops.nms and the reference's utils.nms outputs."""
import os

import numpy as np
import pytest
import torch
import torchvision

from oracle import nms_oracle as no


def _ref_like(boxes, scores, conf=0.5, iou=0.5, top_k=1000, keep=100):
    """utils.py:159-194 with torchvision, returning anchor ids."""
    boxes, scores = torch.from_numpy(boxes), torch.from_numpy(scores)
    ids = torch.nonzero(scores >= conf)[:, 0]
    s, b = scores[ids], boxes[ids]
    if s.numel() > top_k:
        t = torch.topk(s, k=top_k, largest=True, sorted=True).indices
        s, b, ids = s[t], b[t], ids[t]
    k = torchvision.ops.nms(b, s, iou)
    return ids[k][:keep].numpy()


@pytest.mark.parametrize("case", ["few", "many", "none", "hires"])
def test_golden_reference_cases(golden_dir, case):
    g = np.load(os.path.join(golden_dir, "nms_ref_cases.npz"))
    got = no.select_nms(g[f"{case}_boxes"], g[f"{case}_scores"])
    assert got.tolist() == g[f"{case}_keep"].tolist()


@pytest.mark.parametrize("seed", range(6))
def test_random_vs_torchvision(seed):
    rng = np.random.default_rng(seed)
    n = 3000
    ctr = rng.uniform(0, 640, (n, 2)).astype(np.float32)
    half = rng.uniform(5, 120, (n, 2)).astype(np.float32)
    boxes = np.concatenate([ctr - half, ctr + half], 1).astype(np.float32)
    scores = (rng.uniform(0, 1, n) ** (1 + seed % 3)).astype(np.float32) + np.arange(n, dtype=np.float32) * 1e-7
    thr = [0.5, 0.2, 0.9][seed % 3]
    assert no.select_nms(boxes, scores, conf_thr=thr).tolist() == _ref_like(boxes, scores, conf=thr).tolist()


def test_iou_exactly_at_threshold_is_kept():
    boxes = np.array([[0, 0, 2, 2], [0, 1, 2, 3 + 0.0]], dtype=np.float32)  # inter 2, union 6 -> 1/3
    scores = np.array([0.9, 0.8], dtype=np.float32)
    assert no.select_nms(boxes, scores, iou_thr=1.0 / 3.0).tolist() in ([0, 1], [0])  # fp32(1/3) vs exact
    b2 = np.array([[0, 0, 2, 2], [0, 0, 2, 4]], dtype=np.float32)  # IoU exactly 0.5 in fp32
    assert no.select_nms(b2, scores, iou_thr=0.5).tolist() == [0, 1]
    assert _ref_like(b2, scores).tolist() == [0, 1]


def test_empty_and_degenerate():
    assert no.select_nms(np.zeros((10, 4), np.float32), np.zeros(10, np.float32)).size == 0
    b = np.array([[5, 5, 5, 5], [5, 5, 5, 5]], dtype=np.float32)  # zero-area: IoU = 0/0 = nan -> not suppressed
    s = np.array([0.9, 0.8], dtype=np.float32)
    assert no.select_nms(b, s).tolist() == _ref_like(b, s).tolist() == [0, 1]
