"""`PredictionResult` - the object `HeadDetector.__call__` returns, with the reference's public surface
(head_detector/detection_result.py:38-81): `.heads`, `.original_image`, `draw(method='full')`, `get_pncc()`,
`get_aligned_heads()`, `save_meshes(folder)`.

What differs from the reference is cost, not behaviour: the reference builds a `PNCCProcessor` (a Python filter over the
9976 faces, ~90 ms) and a `MeshSaver` in EVERY constructor; here those tables are module-level assets, `get_pncc()` runs
the device rasteriser (`vgh_pncc_render`, bit-identical to Sim3DR) and the refined head boxes of `get_aligned_heads()`
are the reference's arithmetic (batched device form: `mesh.refined_head_bboxes` / `vgh_head_bbox`).  `get_pncc()` keeps the reference's quirk of negating `head.vertices_3d[:, 2]` in place
(pncc_processor.py:70), so calling it twice flips z back - user code that relies on it keeps working."""
import os
from typing import List

import numpy as np

from . import draw_utils, mesh
from .head_info import HeadMetadata
from .utils import extend_bbox, extend_to_rect, refined_head_bbox, vertically_align

MAX_YAW = 60


class PredictionResult:
    def __init__(self, original_image: np.ndarray, heads: List[HeadMetadata]):
        self.original_image = original_image
        self.heads = heads

    def draw(self, method: str = "full"):
        image = self.original_image.copy()
        for head in self.heads:
            for fn in draw_utils.DRAW_MAPPING[method]:
                image = fn(image, head)
        return image

    def get_pncc(self):
        h, w = self.original_image.shape[:2]
        img = mesh.pncc_image(h, w, [head.vertices_3d for head in self.heads])   # depth = -z, as the reference rasterises
        for head in self.heads:
            head.vertices_3d[:, 2] *= -1          # reference quirk (pncc_processor.py:70): z negated in place
        return img

    def get_aligned_heads(self):
        crops = []
        for head in self.heads:
            image, vertices = self.original_image.copy(), head.vertices_3d
            if np.abs(head.head_pose.yaw) < MAX_YAW:
                image, vertices = vertically_align(image, vertices, head.flame_params, head.head_pose.roll)
            box = refined_head_bbox(vertices)     # (batched device form for device-resident meshes: mesh.refined_head_bboxes)
            x, y, w, h = extend_to_rect(extend_bbox([box.x, box.y, box.w, box.h], offset=0.1))
            crops.append(image[y:y + h, x:x + w])
        return crops

    def save_meshes(self, save_folder: str):
        os.makedirs(save_folder, exist_ok=True)
        faces = mesh.tables()["full_faces"].astype(np.int64) + 1     # OBJ indices start at 1 (MeshSaver, detection_result.py:22-35)
        for i, head in enumerate(self.heads):
            with open(os.path.join(save_folder, f"head_{i}.obj"), "w") as f:
                for v in head.vertices_3d:
                    f.write("v %.8f %.8f %.8f\n" % tuple(v))
                for t in faces:
                    f.write("f %d %d %d\n" % tuple(t))

    def __repr__(self):
        return f"PredictionResult(original_image={self.original_image.shape}, num heads={len(self.heads)})"
