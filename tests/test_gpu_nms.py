"""CUDA select+NMS (vgh_select_nms) vs the oracle: kept ORIGINAL anchor ids must be bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import nms_oracle as no
from oracle.make_golden import clustered_anchors

pytestmark = pytest.mark.gpu


def _gpu(boxes, scores, **kw):
    from head_detector_b200.utils import select_nms_indices

    idx, cnt = select_nms_indices(torch.as_tensor(boxes).cuda(), torch.as_tensor(scores).cuda(), **kw)
    idx, cnt = idx.cpu().numpy(), cnt.cpu().numpy()
    out = []
    for b in range(idx.shape[0]):
        assert (idx[b, cnt[b]:] == -1).all()
        out.append(idx[b, :cnt[b]].tolist())
    return out


@pytest.mark.parametrize("case", ["few", "many", "none", "hires"])
def test_golden_reference_cases(golden_dir, case):
    g = np.load(os.path.join(golden_dir, "nms_ref_cases.npz"))
    got = _gpu(g[f"{case}_boxes"][None], g[f"{case}_scores"][None])[0]
    assert got == g[f"{case}_keep"].tolist()


def test_reference_signature_first_image_only(golden_dir):
    from head_detector_b200.utils import nms

    g = np.load(os.path.join(golden_dir, "nms_ref_cases.npz"))
    boxes = torch.from_numpy(np.stack([g["few_boxes"], g["many_boxes"]])).cuda()
    scores = torch.from_numpy(np.stack([g["few_scores"], g["many_scores"]]))[..., None].cuda()
    tag = torch.zeros(2, 8400, 413, device="cuda")
    tag[:, :, 0] = torch.arange(8400, device="cuda")
    b, s, f = nms(boxes, scores, tag)
    assert f[:, 0].long().tolist() == g["few_keep"].tolist()
    assert np.allclose(s.cpu().numpy(), g["few_keep_scores"])
    assert b.shape == (len(g["few_keep"]), 4)


@pytest.mark.parametrize("seed,thr,iou", [(0, 0.5, 0.5), (1, 0.2, 0.5), (2, 0.9, 0.3), (3, 0.05, 0.7), (4, 0.5, 0.0)])
def test_random_batches_bit_exact(seed, thr, iou):
    rng = np.random.default_rng(seed)
    B, n = 5, 8400
    ctr = rng.uniform(0, 640, (B, n, 2)).astype(np.float32)
    half = rng.uniform(3, 150, (B, n, 2)).astype(np.float32)
    boxes = np.concatenate([ctr - half, ctr + half], 2).astype(np.float32)
    scores = (rng.uniform(0, 1, (B, n)) ** (1 + seed)).astype(np.float32) + np.arange(n, dtype=np.float32) * 1e-7
    got = _gpu(boxes, scores, confidence_threshold=thr, iou_threshold=iou)
    for b in range(B):
        assert got[b] == no.select_nms(boxes[b], scores[b], conf_thr=thr, iou_thr=iou).tolist()


def test_ties_break_by_lower_anchor_id():
    boxes = np.tile(np.array([[10, 10, 50, 50]], np.float32), (64, 1))
    boxes[:, 0] += np.arange(64) * 100  # disjoint boxes -> nothing suppressed
    boxes[:, 2] += np.arange(64) * 100
    scores = np.full(64, 0.75, np.float32)
    assert _gpu(boxes[None], scores[None])[0] == no.select_nms(boxes, scores).tolist() == list(range(64))


def test_negative_scores_with_non_positive_threshold():
    """Raw logits instead of probabilities: scores of both signs, threshold <= 0 (round 1 ordered keys by raw float bits,
    which sorts negative floats upside down)."""
    rng = np.random.default_rng(5)
    boxes, _ = clustered_anchors(8400, 30, 10, seed=9, bg_hi=0.9)
    boxes = boxes.numpy()
    scores = (rng.random(8400, dtype=np.float32) * 2 - 1).astype(np.float32)
    scores += np.arange(8400, dtype=np.float32) * 1e-7
    for thr in (-0.5, 0.0, -2.0):
        got = _gpu(boxes[None], scores[None], confidence_threshold=thr)[0]
        want = no.select_nms(boxes, scores, conf_thr=thr).tolist()
        assert got == want and len(want) > 0, thr


def test_topk_and_keep_limits():
    boxes, scores = clustered_anchors(33600, 60, 30, seed=8, size=1280.0, bg_hi=0.7)
    boxes, scores = boxes.numpy(), scores.numpy()
    assert (scores >= 0.5).sum() > 1000
    for top_k, keep in ((1000, 100), (1024, 300), (37, 5)):
        got = _gpu(boxes[None], scores[None], top_k=top_k, keep_top_k=keep)[0]
        assert got == no.select_nms(boxes, scores, top_k=top_k, keep_top_k=keep).tolist()


def test_kept_rows_gathered():
    from head_detector_b200 import _lib

    boxes, scores = clustered_anchors(8400, 8, 12, seed=3)
    b, s = boxes.cuda()[None].contiguous(), scores.cuda()[None].contiguous()
    idx = torch.empty(1, 100, dtype=torch.int32, device="cuda")
    cnt = torch.empty(1, dtype=torch.int32, device="cuda")
    kb = torch.zeros(1, 100, 4, device="cuda")
    ks = torch.zeros(1, 100, device="cuda")
    _lib.check(_lib.lib().vgh_select_nms(b.data_ptr(), s.data_ptr(), 1, 8400, 0.5, 0.5, 1000, 100, idx.data_ptr(), cnt.data_ptr(),
                                         kb.data_ptr(), ks.data_ptr(), _lib.stream_ptr()))
    n = int(cnt[0])
    keep = idx[0, :n].long()
    assert torch.equal(kb[0, :n], b[0][keep]) and torch.equal(ks[0, :n], s[0][keep])


def test_bad_arguments_raise():
    from head_detector_b200.utils import select_nms_indices

    with pytest.raises(RuntimeError):
        select_nms_indices(torch.zeros(1, 10, 4).cuda(), torch.zeros(1, 10).cuda(), top_k=5000)
