"""Device-side `HeadDetector._transform_image` (reference head_detector/detector.py:40-52): Lanczos-4
longest-side resize + centred border, bit-exact with the reference's cv2 calls, for a whole batch of
differently sized images in one kernel launch (`vgh_letterbox`).  Host work left: packing the raw
image bytes into one pinned buffer for a single H2D copy."""
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib


def _as_rgb_u8(image: np.ndarray) -> np.ndarray:
    if image.ndim != 3 or image.shape[2] < 3:
        raise ValueError(f"expected an HxWx3 uint8 image, got shape {tuple(image.shape)}")
    if image.dtype != np.uint8:
        raise ValueError(f"expected a uint8 image, got {image.dtype}")
    return np.ascontiguousarray(image[..., :3])


def letterbox_geometry(h: int, w: int, image_size: int) -> Tuple[Tuple[int, int], Tuple[int, int], float]:
    """((new_h, new_w), (pad_x, pad_y), scale) exactly as detector.py:41-52 computes them."""
    S = image_size
    new_h, new_w = (S, int(w * S / h)) if h > w else (int(h * S / w), S)
    return (new_h, new_w), ((S - new_w) // 2, (S - new_h) // 2), S / max(h, w)


def letterbox_batch(images: Sequence[np.ndarray], image_size: int = 640, out: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """n HxWx3 uint8 RGB images (any sizes) -> (uint8 cuda [n,S,S,3], float32 cpu [n,3] = pad_x, pad_y, scale).

    `out` may be a preallocated cuda uint8 tensor whose first n frames are written (e.g. the engine's
    staging input).  Runs on the current CUDA stream."""
    if not torch.cuda.is_available():
        raise RuntimeError("head_detector_b200.preprocess needs a CUDA device (no CPU fallback; the host path is cv2 in HeadDetector._transform_image)")
    imgs = [_as_rgb_u8(np.asarray(im)) for im in images]
    n, S = len(imgs), int(image_size)
    if out is None:
        out = torch.empty(n, S, S, 3, dtype=torch.uint8, device="cuda")
    if not (out.is_cuda and out.dtype == torch.uint8 and out.is_contiguous() and out.shape[0] >= n and tuple(out.shape[1:]) == (S, S, 3)):
        raise ValueError("out must be a contiguous cuda uint8 tensor [>=n, S, S, 3]")
    xform = torch.zeros(n, 3, dtype=torch.float32)
    if n == 0:
        return out[:0], xform
    offsets = np.zeros(n, np.int64)
    total = 0
    for i, im in enumerate(imgs):
        offsets[i] = total
        total += (im.size + 255) & ~255
    staging = torch.empty(total, dtype=torch.uint8, pin_memory=True)
    flat = staging.numpy()
    for off, im in zip(offsets, imgs):   # one memcpy per frame into pinned memory (a thread pool measured slower on the GPU box)
        np.copyto(flat[off:off + im.size], im.reshape(-1))
    src = staging.cuda(non_blocking=True)
    heights = np.array([im.shape[0] for im in imgs], np.int32)
    widths = np.array([im.shape[1] for im in imgs], np.int32)
    _lib.check(_lib.lib().vgh_letterbox(src.data_ptr(), offsets.ctypes.data, heights.ctypes.data, widths.ctypes.data, n, S,
                                        out.data_ptr(), xform.data_ptr(), _lib.stream_ptr()), "vgh_letterbox")
    # `staging` / `src` may be dropped here: torch's caching allocators hand their blocks out again only in
    # stream order (pinned block: after the copy's event; device block: same stream as the kernel)
    return out[:n], xform
