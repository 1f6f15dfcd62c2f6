"""Letterbox: the numpy oracle and the product's shared host/device arithmetic
(csrc/letterbox_core.h, run on the CPU through tests/harness) against cv2 and against the unmodified
reference's `_transform_image` (golden fixtures).  Everything here is integer work: bit-exact."""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np
import pytest

from oracle import letterbox_oracle as lo

HERE = os.path.dirname(os.path.abspath(__file__))
SHAPES = [(480, 640), (720, 1280), (333, 500), (100, 57), (641, 640), (37, 41), (640, 640), (1500, 900)]


def _cv2_letterbox(img, S=640):
    """Literal transcription of detector.py:40-52 with cv2 (the reference's own calls)."""
    import cv2

    h, w = img.shape[:2]
    new_h, new_w = (S, int(w * S / h)) if h > w else (int(h * S / w), S)
    r = cv2.resize(img, (new_w, new_h), interpolation=cv2.INTER_LANCZOS4)
    pad_w, pad_h = S - r.shape[1], S - r.shape[0]
    r = cv2.copyMakeBorder(r, pad_h // 2, pad_h - pad_h // 2, pad_w // 2, pad_w - pad_w // 2, cv2.BORDER_CONSTANT, value=127)
    return r, (pad_w // 2, pad_h // 2), S / max(h, w)


@pytest.fixture(scope="module")
def harness():
    """Compiles the CPU harness around letterbox_core.h (the functions the CUDA kernel calls)."""
    out_dir = os.path.join(HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "liblb_harness.so")
    src = os.path.join(HERE, "harness", "letterbox_host.cpp")
    core = os.path.join(os.path.dirname(HERE), "head_detector_b200", "csrc", "letterbox_core.h")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(core)):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", src, "-o", so], check=True)
    lib = C.CDLL(so)
    lib.lb_host_letterbox.restype = C.c_int
    lib.lb_host_letterbox.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.lb_host_axis_tables.restype = None
    lib.lb_host_axis_tables.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    return lib


def _harness_letterbox(lib, img, S=640):
    img = np.ascontiguousarray(img)
    out = np.empty((S, S, 3), np.uint8)
    geom = np.zeros(4, np.int32)
    rc = lib.lb_host_letterbox(img.ctypes.data, img.shape[0], img.shape[1], S, out.ctypes.data, geom.ctypes.data)
    return rc, out, geom


@pytest.mark.parametrize("shape", SHAPES)
def test_oracle_matches_cv2(shape):
    img = np.random.default_rng(shape[0] * 7 + shape[1]).integers(0, 256, shape + (3,), dtype=np.uint8)
    want, pad, scale = _cv2_letterbox(img)
    got, gpad, gscale = lo.transform_image(img)
    assert gpad == pad and gscale == scale
    assert np.array_equal(got, want)


def test_oracle_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "letterbox_ref_cases.npz"))
    for i, (h, w) in enumerate(g["shapes"]):
        src = np.random.default_rng(int(g["seed0"]) + i).integers(0, 256, (int(h), int(w), 3), dtype=np.uint8)
        img, pad, scale = lo.transform_image(src)
        assert tuple(pad) == tuple(g[f"pad_{i}"]) and scale == float(g[f"scale_{i}"])
        assert np.array_equal(img[::53, ::47], g[f"probe_{i}"])
        assert np.array_equal(np.frombuffer(hashlib.sha1(img.tobytes()).digest(), dtype=np.uint8), g[f"sha1_{i}"])


@pytest.mark.parametrize("shape", SHAPES)
def test_kernel_arithmetic_on_cpu_matches_cv2(harness, shape):
    """The exact functions the CUDA kernel executes, compiled for the host: same bits as cv2."""
    img = np.random.default_rng(shape[0] * 7 + shape[1]).integers(0, 256, shape + (3,), dtype=np.uint8)
    want, pad, _ = _cv2_letterbox(img)
    rc, got, geom = _harness_letterbox(harness, img)
    assert rc == 0 and (geom[2], geom[3]) == pad
    assert np.array_equal(got, want)


def test_kernel_arithmetic_extreme_patterns(harness):
    """Saturation paths: checkerboards / stripes drive the Lanczos lobes past [0,255]."""
    yy, xx = np.mgrid[0:300, 0:420]
    for pat in (((xx + yy) % 2) * 255, (xx % 2) * 255, (yy % 3 == 0) * 255, np.full_like(xx, 255), np.zeros_like(xx)):
        img = np.repeat(pat[..., None], 3, axis=2).astype(np.uint8)
        img[..., 1] = 255 - img[..., 1]
        want, _, _ = _cv2_letterbox(img)
        rc, got, _ = _harness_letterbox(harness, img)
        assert rc == 0 and np.array_equal(got, want)
        assert np.array_equal(lo.transform_image(img)[0], want)


def test_axis_tables_match_oracle_and_int32_never_wraps(harness):
    worst = 0
    for src, dst in [(640, 640), (1280, 640), (500, 640), (57, 364), (3000, 640), (41, 640), (640, 639)]:
        ofs = np.zeros(dst, np.int32)
        coef = np.zeros((dst, 8), np.int16)
        harness.lb_host_axis_tables(src, dst, ofs.ctypes.data, coef.ctypes.data)
        o, c = lo.axis_tables(src, dst)
        assert np.array_equal(ofs, o) and np.array_equal(coef, c)
        c = c.astype(np.int64)
        pos, neg = np.clip(c, 0, None).sum(1).max(), -np.clip(c, None, 0).sum(1).min()
        worst = max(worst, 255 * (pos * pos + neg * neg))  # rows chosen to maximise |vertical sum|
    assert worst + (1 << 21) < 2 ** 31


def test_empty_extent_is_an_error(harness):
    img = np.zeros((2000, 2, 3), np.uint8)  # int(2 * 640 / 2000) == 0: cv2.resize raises
    rc, _, _ = _harness_letterbox(harness, img)
    assert rc != 0
    with pytest.raises(ValueError):
        lo.transform_image(img)


def test_random_shapes_oracle_and_kernel_arithmetic_vs_cv2(harness):
    """Seeded sweep over odd shapes (up- and down-scaling, extreme aspect ratios, tiny images) and two letterbox
    sizes: numpy oracle == cv2 == the product's device arithmetic run on the host."""
    rng = np.random.default_rng(2024)
    for _ in range(12):
        h, w = int(rng.integers(1, 900)), int(rng.integers(1, 900))
        S = int(rng.choice([96, 320]))
        new_h, new_w = (S, int(w * S / h)) if h > w else (int(h * S / w), S)
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        if new_h <= 0 or new_w <= 0:
            rc, _, _ = _harness_letterbox(harness, img, S)
            assert rc != 0
            continue
        want, pad, scale = _cv2_letterbox(img, S)
        got, gpad, gscale = lo.transform_image(img, S)
        assert gpad == pad and gscale == scale and np.array_equal(got, want), (h, w, S)
        rc, dev, geom = _harness_letterbox(harness, img, S)
        assert rc == 0 and (geom[2], geom[3]) == pad and np.array_equal(dev, want), (h, w, S)
