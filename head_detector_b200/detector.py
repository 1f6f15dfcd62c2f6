"""`HeadDetector` with the reference's constructor and call signature
(head_detector/detector.py:18-102), running on the B200 engine.

    HeadDetector(model="vgg_heads_l", image_size=640)(image, confidence_threshold=0.5)
        -> PredictionResult with .heads[i].bbox / .score / .flame_params / .vertices_3d / .head_pose

Differences that are extensions, not changes: `weights=` (a deploy-form weight dict or a path to a
torch-saved one; the HF download of detector.py:25-30 is impossible offline), `batch_size=` and
`detect_batch()` (batched semantics of yolo_heads_post_prediction_callback.py:55-97), and
`device_letterbox=` (default True: `_transform_image` runs as one CUDA kernel for the whole batch,
bit-exact with the cv2 calls of detector.py:47-50; False keeps the reference's host cv2 path) and
`sparse_heads=` (default True: the FLAME branch of the detection heads is evaluated after NMS on 8x8
windows around the surviving anchors instead of on the whole feature maps - same heads, same numbers)."""
import os
import warnings
from typing import Any, Dict, List, Optional, Tuple, Union

import numpy as np
import torch

from . import arch
from .detection_result import PredictionResult
from .engine import Engine
from .flame import FLAMELayer
from .head_info import FLAME_CONSTS, Bbox, FlameParams, HeadMetadata
from .preprocess import letterbox_batch, letterbox_geometry
from .utils import rpy_from_rotations


class HeadDetector:
    def __init__(self, model: str = "vgg_heads_l", image_size: int = 640, weights: Union[None, str, Dict[str, torch.Tensor]] = None,
                 batch_size: int = 1, keep_top_k: int = 100, device_letterbox: bool = True, sparse_heads: bool = True):
        if not torch.cuda.is_available():
            raise RuntimeError("head_detector_b200.HeadDetector needs a CUDA device (sm_100a); there is no CPU fallback")
        self._image_size = image_size
        self._device = torch.device("cuda")
        self._flame = FLAMELayer()
        self._batch = batch_size
        self._keep_top_k = keep_top_k
        self._device_letterbox = device_letterbox
        self._sparse_heads = sparse_heads   # FLAME branch of the heads on the NMS survivors only (same predictions)
        self.model = self._read_model(model, weights)

    def _read_model(self, model: str, weights=None) -> Engine:
        if model != "vgg_heads_l":
            raise ValueError(f"unknown model {model!r}; only 'vgg_heads_l' (YoloHeads_L) is built")
        if weights is None:
            weights = os.environ.get("VGGHEADS_B200_WEIGHTS")
        if isinstance(weights, str):
            weights = torch.load(weights, map_location="cpu")
        if weights is None:
            warnings.warn("no weights given and the released vgg_heads_l checkpoint is unreachable offline: "
                          "using seeded random-init weights (architecture and cost are exact, detections are not meaningful)")
            weights = arch.synthetic_weights(0)
        return Engine(weights, self._batch, self._image_size, self._keep_top_k, self._flame, sparse_heads=self._sparse_heads)

    # -- host-side pre-processing, same arithmetic as detector.py:32-56
    def _convert_image(self, image) -> np.ndarray:
        if isinstance(image, str):
            import cv2

            image = cv2.cvtColor(cv2.imread(image), cv2.COLOR_BGR2RGB)
        elif not isinstance(image, np.ndarray):
            image = np.array(image)
        return image

    def _transform_image(self, image: np.ndarray) -> Tuple[np.ndarray, Tuple[int, int], float]:
        import cv2

        S = self._image_size
        h, w = image.shape[:2]
        new_h, new_w = (S, int(w * S / h)) if h > w else (int(h * S / w), S)
        scale = S / max(h, w)
        if (new_h, new_w) != (h, w):
            image = cv2.resize(image, (new_w, new_h), interpolation=cv2.INTER_LANCZOS4)
        pad_w, pad_h = S - image.shape[1], S - image.shape[0]
        if pad_w or pad_h:
            image = cv2.copyMakeBorder(image, pad_h // 2, pad_h - pad_h // 2, pad_w // 2, pad_w - pad_w // 2,
                                       cv2.BORDER_CONSTANT, value=127)
        return np.ascontiguousarray(image[..., :3], dtype=np.uint8), (pad_w // 2, pad_h // 2), scale

    # -- batched device path
    def detect_batch(self, images_u8: torch.Tensor, confidence_threshold: float = 0.5, img_xform: Optional[torch.Tensor] = None):
        """uint8 [B,S,S,3] (cuda) -> dict of device tensors: keep_cnt, keep_boxes, keep_scores, offsets,
        params [N,413], vertices [N,5023,3], rotations [N,3,3] (N = total heads, image-major)."""
        eng = self.model
        eng.forward(images_u8.to(self._device))
        eng.postprocess(confidence_threshold, 0.5, 1000, img_xform)
        n = int(eng.head_offsets[-1])
        return {"keep_cnt": eng.keep_cnt, "keep_boxes": eng.keep_boxes, "keep_scores": eng.keep_scores,
                "offsets": eng.head_offsets, "params": eng.head_params(n), "vertices": eng.head_verts(n),
                "rotations": eng.head_rot(n)}

    def _parse_batch(self, out: Dict[str, torch.Tensor], caches: List[Dict[str, Any]]) -> List[List[HeadMetadata]]:
        """detector.py:61-90 for the first len(caches) images of a batch, vectorised over the batch: ONE
        device->host copy per tensor, box un-letterboxing / rounding, roll-pitch-yaw and the `scale /= scale`
        of detector.py:79 as whole-batch array ops, parameter fields as views of one [N,413] host tensor
        (`Tensor.split`), so that the per-head Python work is only the construction of the result objects."""
        S, n_img = self._image_size, len(caches)
        offsets = out["offsets"].cpu().numpy().astype(np.int64)
        total = int(offsets[n_img])
        cnt = np.diff(offsets[:n_img + 1])
        img_of = np.repeat(np.arange(n_img), cnt)                                   # image of every head
        slot = np.arange(total) - np.repeat(offsets[:n_img], cnt)                   # its slot in the image's keep list
        boxes = out["keep_boxes"][:n_img].cpu().numpy()[img_of, slot]               # [N,4]
        scores = out["keep_scores"][:n_img].cpu().numpy()[img_of, slot]
        verts = out["vertices"][:total].cpu().numpy()                               # un-padded / un-scaled on the device
        params = out["params"][:total].cpu()
        rots = out["rotations"][:total].cpu().numpy()
        pad = np.array([[c["padding"][0], c["padding"][1]] for c in caches], dtype=np.float32).reshape(n_img, 2)[img_of]
        scale64 = np.array([c["scale"] for c in caches], dtype=np.float64).reshape(n_img)[img_of]
        scale = scale64.astype(np.float32)
        boxes = boxes.clip(0, S)
        boxes[:, [0, 2]] -= pad[:, :1]
        boxes[:, [1, 3]] -= pad[:, 1:]
        boxes /= scale[:, None]
        boxes = np.rint(boxes).astype(int)
        wh = boxes[:, 2:] - boxes[:, :2]
        poses = rpy_from_rotations(rots)
        fields, at = {}, 0
        for key in ("shape", "expression", "jaw", "rotation", "eyeballs", "neck", "translation", "scale"):   # READ order, head_info.py:54-78
            width = FLAME_CONSTS[key]
            col = params[:, at:at + width]
            if key == "scale":
                col = col / torch.from_numpy(scale64).to(col.dtype)[:, None]        # detector.py:79
            fields[key] = col.split(1) if total else ()
            at += width
        heads: List[List[HeadMetadata]] = [[] for _ in range(n_img)]
        for i in range(total):
            fp = FlameParams(**{k: v[i] for k, v in fields.items()})
            b = boxes[i]
            heads[img_of[i]].append(HeadMetadata(bbox=Bbox(x=b[0], y=b[1], w=wh[i, 0], h=wh[i, 1]), score=scores[i],
                                                 flame_params=fp, vertices_3d=verts[i], head_pose=poses[i]))
        return heads

    def _parse_predictions(self, out: Dict[str, torch.Tensor], img: int, cache: Dict[str, Any]) -> List[HeadMetadata]:
        """detector.py:61-90 for image `img` of the batch."""
        lo, hi = int(out["offsets"][img]), int(out["offsets"][img + 1])
        one = {"offsets": torch.tensor([0, hi - lo]), "keep_boxes": out["keep_boxes"][img:img + 1], "keep_scores": out["keep_scores"][img:img + 1],
               "vertices": out["vertices"][lo:hi], "params": out["params"][lo:hi], "rotations": out["rotations"][lo:hi]}
        return self._parse_batch(one, [cache])[0]

    def predict_batch(self, images: List[Any], confidence_threshold: float = 0.5) -> List[PredictionResult]:
        """Batched `__call__` (extension; semantics of yolo_heads_post_prediction_callback.py:55-97: every image
        independently).  `len(images)` must not exceed the `batch_size` the detector was built with; the
        batch is padded with copies of the last image."""
        if not 0 < len(images) <= self._batch:
            raise ValueError(f"predict_batch takes 1..{self._batch} images, got {len(images)}")
        originals = [self._convert_image(im) for im in images]
        batch, xf, caches = self._prepare_batch(originals)
        out = self.detect_batch(batch, confidence_threshold, xf)
        heads = self._parse_batch(out, caches)
        return [PredictionResult(original_image=o, heads=h) for o, h in zip(originals, heads)]

    def _prepare_batch(self, originals: List[np.ndarray]):
        """`_preprocess` (detector.py:54-56) for up to batch_size images: letterboxed uint8 cuda batch
        [batch_size,S,S,3] (padded with copies of the last image), xform [batch_size,3], per-image caches."""
        S, n = self._image_size, len(originals)
        if self._device_letterbox:
            frames = torch.empty(self._batch, S, S, 3, dtype=torch.uint8, device=self._device)
            _, xf = letterbox_batch(originals, S, out=frames)
            if n < self._batch:
                frames[n:] = frames[n - 1]
                xf = torch.cat([xf, xf[-1:].expand(self._batch - n, 3)])
            geo = [letterbox_geometry(o.shape[0], o.shape[1], S) for o in originals]
            return frames, xf, [{"padding": g[1], "scale": g[2]} for g in geo]
        prepared = [self._transform_image(o) for o in originals]
        caches = [{"padding": p[1], "scale": p[2]} for p in prepared]
        while len(prepared) < self._batch:
            prepared.append(prepared[-1])
        batch = torch.from_numpy(np.stack([p[0] for p in prepared])).to(self._device)
        xf = torch.tensor([[p[1][0], p[1][1], p[2]] for p in prepared], dtype=torch.float32)
        return batch, xf, caches

    def __call__(self, image, confidence_threshold: float = 0.5) -> PredictionResult:
        original = self._convert_image(image)
        batch, xf, caches = self._prepare_batch([original])
        out = self.detect_batch(batch, confidence_threshold, xf)
        heads = self._parse_predictions(out, 0, caches[0])
        return PredictionResult(original_image=original, heads=heads)
