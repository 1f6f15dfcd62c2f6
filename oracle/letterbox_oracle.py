"""CPU oracle of the letterbox pre-processing.  TEST INFRASTRUCTURE ONLY.

Restates `HeadDetector._transform_image` (reference head_detector/detector.py:40-52).  The resize
arithmetic is third party: `cv2.resize(..., INTER_LANCZOS4)` of OpenCV (the reference pins
opencv-contrib-python-headless==4.9.0.80, requirements.txt:3; this container has opencv 4.13).  Its
published algorithm for 8-bit images (modules/imgproc/src/resize.cpp: `interpolateLanczos4`,
`resizeGeneric_` with `HResizeLanczos4<uchar,int,short>` / `VResizeLanczos4<uchar,int,short,
FixedPtCast<int,uchar,22>>`) is restated here in numpy:

  * per destination index d: fx = float32((d + 0.5) * scale - 0.5), s = floor(fx), fx -= s;
    8 float32 weights `interpolateLanczos4(fx)`, rounded half-to-even to int16 at scale 2^11;
  * horizontal pass: exact int32 sums over taps s-3 .. s+4, indices clamped (border replicate);
  * vertical pass: exact int32 sums, then (v + 2^21) >> 22 saturated to uint8.

Pinned by: cv2 itself (tests/test_oracle_letterbox.py compares on seeded images of many shapes) and by
outputs of the unmodified reference `_transform_image` run in the build container
(tests/golden/letterbox_ref.npz, letterbox_ref_cases.npz; generator oracle/make_golden.py)."""
import math
from typing import Tuple

import numpy as np

_S45 = 0.70710678118654752440084436210485
_CS = ((1, 0), (-_S45, -_S45), (0, 1), (_S45, -_S45), (-1, 0), (_S45, _S45), (0, -1), (-_S45, _S45))
COEF_BITS = 11


def lanczos4_weights(x) -> np.ndarray:
    """OpenCV `interpolateLanczos4` (float32 x in [0,1)) -> 8 float32 weights."""
    f32 = np.float32
    x3 = f32(f32(x) + f32(3))                 # `x+3` is evaluated in float
    y0 = -float(x3) * math.pi * 0.25
    s0, c0 = math.sin(y0), math.cos(y0)
    co = np.zeros(8, f32)
    total = f32(0)
    for i in range(8):
        y0_ = f32(x3 - f32(i))
        if abs(y0_) >= f32(1e-6):
            y = -float(y0_) * math.pi * 0.25
            co[i] = f32((_CS[i][0] * s0 + _CS[i][1] * c0) / (y * y))
        else:
            co[i] = f32(1e30)
        total = f32(total + co[i])
    inv = f32(f32(1.0) / total)
    return (co * inv).astype(f32)


def axis_tables(src: int, dst: int) -> Tuple[np.ndarray, np.ndarray]:
    """First tap index [dst] (int, may be out of range: clamp on use) and int weights [dst,8]."""
    scale = 1.0 / (dst / src)
    ofs = np.zeros(dst, np.int64)
    coef = np.zeros((dst, 8), np.int64)
    for d in range(dst):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(math.floor(f))
        f = np.float32(f - np.float32(s))
        w = lanczos4_weights(f) * np.float32(1 << COEF_BITS)
        ofs[d] = s - 3
        coef[d] = np.clip(np.rint(w), -32768, 32767).astype(np.int64)
    return ofs, coef


def resize_lanczos4(img: np.ndarray, new_w: int, new_h: int) -> np.ndarray:
    """cv2.resize(img, (new_w, new_h), interpolation=cv2.INTER_LANCZOS4) for uint8 HWC images."""
    assert img.dtype == np.uint8 and img.ndim == 3
    h, w = img.shape[:2]
    xo, xa = axis_tables(w, new_w)
    yo, ya = axis_tables(h, new_h)
    xi = np.clip(xo[:, None] + np.arange(8)[None], 0, w - 1)
    yi = np.clip(yo[:, None] + np.arange(8)[None], 0, h - 1)
    src = img.astype(np.int64)
    hor = np.zeros((h, new_w, img.shape[2]), np.int64)
    for k in range(8):                       # tap by tap: keeps the temporaries small
        hor += src[:, xi[:, k], :] * xa[None, :, k, None]
    ver = np.zeros((new_h, new_w, img.shape[2]), np.int64)
    for k in range(8):
        ver += hor[yi[:, k]] * ya[:, k, None, None]
    assert np.abs(hor).max(initial=0) < 2 ** 31 and np.abs(ver).max(initial=0) < 2 ** 31  # OpenCV's int32 never wraps
    return np.clip((ver + (1 << (2 * COEF_BITS - 1))) >> (2 * COEF_BITS), 0, 255).astype(np.uint8)


def transform_image(image: np.ndarray, image_size: int = 640):
    """detector.py:40-52 up to (and excluding) the tensor conversion: (uint8 [S,S,3], (pad_x, pad_y), scale)."""
    S = image_size
    h, w = image.shape[:2]
    if h > w:
        new_h, new_w = S, int(w * S / h)
    else:
        new_h, new_w = int(h * S / w), S
    scale = S / max(image.shape[:2])
    if new_h <= 0 or new_w <= 0:
        raise ValueError("empty resized extent (cv2.resize raises)")
    image = resize_lanczos4(image, new_w, new_h)
    pad_w, pad_h = S - new_w, S - new_h
    # cv2.copyMakeBorder(BORDER_CONSTANT, value=127): the Python scalar becomes cv::Scalar(127, 0, 0, 0), so the
    # border pixel of the RGB frame is (127, 0, 0) - a quirk of the reference that is part of its output
    out = np.zeros((S, S, image.shape[2]), np.uint8)
    out[..., 0] = 127
    out[pad_h // 2: pad_h // 2 + new_h, pad_w // 2: pad_w // 2 + new_w] = image
    return out, (pad_w // 2, pad_h // 2), scale
