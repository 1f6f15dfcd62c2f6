"""Device letterbox (`vgh_letterbox`) against cv2 (the reference's own calls,
detector.py:47-50), the numpy oracle and the golden outputs of the unmodified reference.  uint8 work:
every comparison is bit-exact."""
import hashlib
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cv2_letterbox(img, S=640):
    import cv2

    h, w = img.shape[:2]
    new_h, new_w = (S, int(w * S / h)) if h > w else (int(h * S / w), S)
    r = cv2.resize(img, (new_w, new_h), interpolation=cv2.INTER_LANCZOS4)
    pad_w, pad_h = S - r.shape[1], S - r.shape[0]
    r = cv2.copyMakeBorder(r, pad_h // 2, pad_h - pad_h // 2, pad_w // 2, pad_w - pad_w // 2, cv2.BORDER_CONSTANT, value=127)
    return r, (pad_w // 2, pad_h // 2), S / max(h, w)


def _rand(shape, seed):
    return np.random.default_rng(seed).integers(0, 256, tuple(shape) + (3,), dtype=np.uint8)


def test_mixed_batch_matches_cv2_and_oracle():
    from head_detector_b200.preprocess import letterbox_batch
    from oracle import letterbox_oracle as lo

    shapes = [(480, 640), (720, 1280), (1280, 720), (333, 500), (100, 57), (641, 640), (37, 41), (640, 640), (1500, 900), (1, 1), (3, 1000)]
    imgs = [_rand(s, i) for i, s in enumerate(shapes)]
    out, xf = letterbox_batch(imgs, 640)
    out = out.cpu().numpy()
    for i, im in enumerate(imgs):
        want, pad, scale = _cv2_letterbox(im)
        assert (int(xf[i, 0]), int(xf[i, 1])) == pad and float(xf[i, 2]) == np.float32(scale), shapes[i]
        assert np.array_equal(out[i], want), shapes[i]
    for i in (3, 4):
        assert np.array_equal(out[i], lo.transform_image(imgs[i])[0])


def test_reference_golden_cases(golden_dir):
    from head_detector_b200.preprocess import letterbox_batch

    g = np.load(os.path.join(golden_dir, "letterbox_ref_cases.npz"))
    imgs = [_rand((int(h), int(w)), int(g["seed0"]) + i) for i, (h, w) in enumerate(g["shapes"])]
    g1 = np.load(os.path.join(golden_dir, "letterbox_ref.npz"))
    imgs.append(np.random.default_rng(int(g1["seed"])).integers(0, 256, tuple(g1["shape"]), dtype=np.uint8))
    out, xf = letterbox_batch(imgs, 640)
    out = out.cpu().numpy()
    for i in range(len(g["shapes"])):
        assert (int(xf[i, 0]), int(xf[i, 1])) == tuple(g[f"pad_{i}"]) and float(xf[i, 2]) == np.float32(g[f"scale_{i}"])
        assert np.array_equal(out[i][::53, ::47], g[f"probe_{i}"])
        assert np.array_equal(np.frombuffer(hashlib.sha1(out[i].tobytes()).digest(), dtype=np.uint8), g[f"sha1_{i}"])
    assert np.array_equal(np.frombuffer(hashlib.sha1(out[-1].tobytes()).digest(), dtype=np.uint8), g1["sha1"])


def test_saturating_patterns_and_other_sizes():
    from head_detector_b200.preprocess import letterbox_batch

    yy, xx = np.mgrid[0:300, 0:420]
    imgs = []
    for pat in (((xx + yy) % 2) * 255, (xx % 2) * 255, (yy % 3 == 0) * 255, np.full_like(xx, 255), np.zeros_like(xx)):
        im = np.repeat(pat[..., None], 3, axis=2).astype(np.uint8)
        im[..., 1] = 255 - im[..., 1]
        imgs.append(im)
    for S in (640, 1280, 96):
        out, _ = letterbox_batch(imgs, S)
        for o, im in zip(out.cpu().numpy(), imgs):
            assert np.array_equal(o, _cv2_letterbox(im, S)[0])


def test_full_batch_properties_1080p():
    """BASELINE batch (32) of 1080x1920 frames: idempotence (a letterboxed frame is a fixed point), exact
    border, and checksums of two frames against cv2."""
    from head_detector_b200.preprocess import letterbox_batch

    imgs = [_rand((1080, 1920), 50 + i) for i in range(32)]
    out, xf = letterbox_batch(imgs, 640)
    assert out.shape == (32, 640, 640, 3) and torch.all(xf[:, 1] == 140) and torch.all(xf[:, 0] == 0)
    border = torch.tensor([127, 0, 0], dtype=torch.uint8, device="cuda")
    assert torch.all(out[:, :140] == border) and torch.all(out[:, 500:] == border)
    again, xf2 = letterbox_batch(list(out.cpu().numpy()), 640)
    assert torch.equal(again, out) and torch.all(xf2 == torch.tensor([0.0, 0.0, 1.0]))
    for i in (0, 31):
        assert np.array_equal(out[i].cpu().numpy(), _cv2_letterbox(imgs[i])[0])


def test_empty_batch_and_errors():
    from head_detector_b200.preprocess import letterbox_batch

    out, xf = letterbox_batch([], 640)
    assert out.shape == (0, 640, 640, 3) and xf.shape == (0, 3)
    with pytest.raises(RuntimeError, match="empty resized extent"):
        letterbox_batch([np.zeros((2000, 2, 3), np.uint8)], 640)  # cv2.resize raises on this one too
    with pytest.raises(ValueError):
        letterbox_batch([np.zeros((20, 20), np.uint8)], 640)
    with pytest.raises(ValueError):
        letterbox_batch([np.zeros((20, 20, 3), np.float32)], 640)
    out, _ = letterbox_batch([np.zeros((20, 30, 4), np.uint8)], 640)  # alpha channel dropped like `_transform_image`
    assert out.shape == (1, 640, 640, 3)


def test_head_detector_device_and_host_letterbox_agree():
    from head_detector_b200 import HeadDetector

    det = HeadDetector(image_size=128, batch_size=3, weights="synthetic")
    imgs = [_rand((90, 160), 1), _rand((200, 100), 2)]
    det._device_letterbox = True
    fd, xd, cd = det._prepare_batch(imgs)
    det._device_letterbox = False
    fh, xh, ch = det._prepare_batch(imgs)
    assert torch.equal(fd, fh) and torch.equal(xd, xh) and cd == ch
    res = det.predict_batch(imgs)
    assert len(res) == 2
