"""`HeadDetector` with the reference's constructor and call signature
(head_detector/detector.py:18-102), running on the B200 engine.

    HeadDetector(model="vgg_heads_l", image_size=640)(image, confidence_threshold=0.5)
        -> PredictionResult with .heads[i].bbox / .score / .flame_params / .vertices_3d / .head_pose

The reference's seam is kept name for name: `_read_model(model)`, `_convert_image`, `_transform_image`, `_preprocess`,
`_process`, `_postprocess(predictions, cache, confidence_threshold)`, `_parse_predictions(bboxes_xyxy, scores,
flame_params, cache)`, `__call__` (detector.py:25-102), so a subclass written against the reference still fits.

Extensions (keyword-only constructor arguments, extra methods): `weights=` - path to the TorchScript blob
`vgg_heads_l.trcd` the reference downloads (the HF download itself is impossible offline), a state_dict checkpoint, a
deploy-form dict, or "synthetic" (`weights.py`; nothing is substituted silently); `batch_size=`, `predict_batch()` and
`detect_batch()` (batched semantics of yolo_heads_post_prediction_callback.py:55-97); `device_letterbox=` (default True:
`_transform_image` runs as one CUDA kernel for the whole batch, bit-exact with the cv2 calls of detector.py:47-50; False
keeps the reference's host cv2 path); `sparse_heads=` (default True: the FLAME branch of the detection heads is
evaluated after NMS on 8x8 windows around the surviving anchors instead of on the whole feature maps - same heads, same
numbers); `parity=` (fp32-class conv arithmetic for end-to-end comparisons with the reference)."""
import os
from typing import Any, Dict, List, Optional, Tuple, Union

import numpy as np
import torch

from .detection_result import PredictionResult
from .engine import Engine
from .flame import FLAMELayer
from .head_info import FLAME_CONSTS, Bbox, FlameParams, HeadMetadata
from .preprocess import letterbox_batch, letterbox_geometry
from .utils import rpy_from_rotations
from .weights import resolve as resolve_weights


class HeadDetector:
    def __init__(self, model: str = "vgg_heads_l", image_size: int = 640, *, weights: Union[None, str, Dict[str, torch.Tensor]] = None,
                 batch_size: int = 1, keep_top_k: int = 100, device_letterbox: bool = True, sparse_heads: bool = True, parity: bool = False,
                 act_dtype: Optional[str] = None):
        """`model`, `image_size`: the reference's arguments (detector.py:19).  Keyword-only extensions: `weights` - path to
        `vgg_heads_l.trcd` (the TorchScript blob the reference downloads) / a state_dict checkpoint / a deploy-form dict /
        the literal "synthetic" (default: $VGGHEADS_B200_WEIGHTS; there is no silent fallback to random weights);
        `batch_size` for `predict_batch`; `parity=True` runs the conv network in the fp32-class split-bf16 mode
        (slow; dense heads) for end-to-end comparisons with the reference's fp32 path."""
        if not torch.cuda.is_available():
            raise RuntimeError("head_detector_b200.HeadDetector needs a CUDA device (sm_100a); there is no CPU fallback")
        self._image_size = image_size
        self._device = torch.device("cuda")
        self._flame = FLAMELayer()
        self._batch = batch_size
        self._keep_top_k = keep_top_k
        self._device_letterbox = device_letterbox
        self._sparse_heads = sparse_heads and not parity   # FLAME branch of the heads on the NMS survivors only (same predictions)
        self._parity = parity
        self._act_dtype = act_dtype   # "fp16" | "bf16" storage of the throughput mode (None: $VGGHEADS_B200_ACT or "fp16")
        self._weights = weights if weights is not None else os.environ.get("VGGHEADS_B200_WEIGHTS")
        self.model = self._read_model(model)

    def _read_model(self, model: str) -> Engine:
        """detector.py:25-30: where the reference `torch.jit.load`s the downloaded blob, this builds the engine from the
        same blob (weights.load_checkpoint: TorchScript -> state_dict -> re-parameterised, packed weights)."""
        if model != "vgg_heads_l":
            raise ValueError(f"unknown model {model!r}; only 'vgg_heads_l' (YoloHeads_L) is built")
        w = resolve_weights(self._weights, model)
        return Engine(w, self._batch, self._image_size, self._keep_top_k, self._flame, sparse_heads=self._sparse_heads, parity=self._parity,
                      act_dtype=self._act_dtype)

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

    # -- the reference's seam (detector.py:54-59,61-95), same names / arguments / return types
    def _preprocess(self, image: np.ndarray) -> Tuple[torch.Tensor, Dict[str, Any]]:
        """detector.py:54-56.  The model input here is the letterboxed uint8 frame [batch,S,S,3] on the device (the /255
        of detector.py:51 is folded into the stem weights)."""
        batch, xf, caches = self._prepare_batch([image])
        cache = dict(caches[0])
        cache["_xform"] = xf
        return batch, cache

    def _process(self, image: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]]:
        """detector.py:58-59 `self.model(image)` -> (boxes [B,A,4], scores [B,A,1], flame [B,A,413]).  With sparse heads the
        dense 413-wide tensor is never built (its rows are assembled for the NMS survivors only): flame is None."""
        boxes, scores = self.model.forward(image.to(self._device))
        return boxes, scores[..., None], (None if self.model.sparse_heads else self.model.dense_flame())

    def _postprocess(self, predictions, cache: Dict[str, Any], confidence_threshold: float) -> List[HeadMetadata]:
        """detector.py:92-95: `nms` (first image, utils.py:159-194) + `_parse_predictions`.  Select / top-k / NMS, the survivors'
        413-float rows and the FLAME decode all run on the device in `Engine.postprocess`; the decoded vertices ride along
        in the cache so that `_parse_predictions` does not decode a second time."""
        eng = self.model
        eng.postprocess(confidence_threshold, 0.5, 1000, cache.get("_xform"))
        n = int(eng.head_offsets[1])
        cache = dict(cache)
        cache["_decoded"] = (eng.head_verts(n), eng.head_rot(n))
        return self._parse_predictions(eng.keep_boxes[0, :n], eng.keep_scores[0, :n], eng.head_params(n), cache)

    def _parse_predictions(self, bboxes_xyxy: torch.Tensor, scores: torch.Tensor, flame_params: torch.Tensor, cache: Dict[str, Any]) -> List[HeadMetadata]:
        """detector.py:61-90, reference signature: kept boxes [n,4], scores [n], flame rows [n,413] + {"padding", "scale"}."""
        n = int(flame_params.shape[0])
        pad, scale = cache["padding"], cache["scale"]
        if "_decoded" in cache:
            verts, rots = cache["_decoded"]
        else:   # called on its own: decode here (vertices un-padded / un-scaled by the kernel, detector.py:67-69)
            xf = torch.tensor([[pad[0], pad[1], scale]], dtype=torch.float32).expand(max(n, 1), 3)[:n]
            _, rots, verts = self._flame.decode(flame_params, xf, live=(300, 100)) if n else (None, torch.zeros(0, 3, 3), torch.zeros(0, 5023, 3))
        out = {"offsets": torch.tensor([0, n]), "keep_boxes": bboxes_xyxy.reshape(1, n, 4), "keep_scores": scores.reshape(1, n),
               "vertices": verts, "params": flame_params, "rotations": rots}
        return self._parse_batch(out, [cache])[0]

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
        """detector.py:97-102, step for step."""
        original_image = self._convert_image(image)
        image, cache = self._preprocess(original_image)
        predictions = self._process(image)
        heads = self._postprocess(predictions, cache, confidence_threshold)
        return PredictionResult(original_image=original_image, heads=heads)
