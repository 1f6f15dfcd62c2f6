"""`PredictionResult` - the object `HeadDetector.__call__` returns (reference:
head_detector/detection_result.py:38-81).  The container and `.heads` are the contract of the
hot path; drawing / PNCC / aligned crops are host-side visualisation that SURVEY.md section 8
marks out of scope - `draw('bbox')` and `save_meshes` are provided, the rest raise."""
import os
from typing import List

import numpy as np

from .head_info import HeadMetadata

_FACES = None


def _faces():
    global _FACES
    if _FACES is None:
        z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets", "flame_generic.npz"))
        _FACES = z["faces"].astype(np.int64) + 1
    return _FACES


class PredictionResult:
    def __init__(self, original_image: np.ndarray, heads: List[HeadMetadata]):
        self.original_image = original_image
        self.heads = heads

    def draw(self, method: str = "bbox"):
        import cv2

        if method != "bbox":
            raise NotImplementedError("only draw('bbox') is provided; landmark/pose rendering is out of scope (SURVEY.md 8)")
        image = self.original_image.copy()
        for h in self.heads:
            x, y, w, hh = (int(v) for v in h.bbox)
            cv2.rectangle(image, (x, y), (x + w, y + hh), (0, 255, 0), 2)
        return image

    def get_pncc(self):
        raise NotImplementedError("PNCC rendering (CPU rasteriser Sim3DR) is out of scope of the B200 hot path (SURVEY.md 8 f4)")

    def get_aligned_heads(self):
        raise NotImplementedError("aligned head crops are host-side visualisation, out of scope (SURVEY.md 8 f4)")

    def save_meshes(self, save_folder: str):
        os.makedirs(save_folder, exist_ok=True)
        for i, head in enumerate(self.heads):
            with open(os.path.join(save_folder, f"head_{i}.obj"), "w") as f:
                for v in head.vertices_3d:
                    f.write("v %.8f %.8f %.8f\n" % tuple(v))
                for t in _faces():
                    f.write("f %d %d %d\n" % tuple(t))

    def __repr__(self):
        return f"PredictionResult(original_image={self.original_image.shape}, num heads={len(self.heads)})"
