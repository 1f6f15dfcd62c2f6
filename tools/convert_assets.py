"""One-off: the reference's index assets (head_detector/assets/*, read-only data: SURVEY.md section 2 #8) -> one npz under
head_detector_b200/assets/.  Adds the two derived tables PNCCProcessor.__init__ builds on every PredictionResult in the
reference (pncc_processor.py:59-64: faces restricted to `head_w_ears`, NCC colour codes of the float64 template).
usage: python tools/convert_assets.py [/root/reference/head_detector/assets] [out.npz]"""
import os
import sys

import numpy as np


def convert(ref_assets: str, out: str):
    face = np.load(os.path.join(ref_assets, "flame_indices", "face.npy"), allow_pickle=True)[()]
    head = np.load(os.path.join(ref_assets, "flame_indices", "head_indices.npy"), allow_pickle=True)[()]
    ears = np.load(os.path.join(ref_assets, "flame_indices", "head_w_ears.npy"))
    tri = np.loadtxt(os.path.join(ref_assets, "triangles.txt"), delimiter=",").astype(np.int32)
    faces = np.load(os.path.join(ref_assets, "full_faces.npy"))
    vt = np.load(os.path.join(ref_assets, "v_template.npy"))
    pncc_tri = faces[np.isin(faces, ears).all(axis=1)].astype(np.int32)
    sub = vt[ears]
    lo, hi = sub.min(axis=0, keepdims=True, initial=0), sub.max(axis=0, keepdims=True, initial=0)
    np.savez_compressed(out, face=np.asarray(face, np.int32), head_indices=np.asarray(head, np.int32), head_w_ears=ears.astype(np.int32),
                        triangles=tri, full_faces=faces.astype(np.int32), pncc_triangles=pncc_tri,
                        ncc_colors=((vt - lo) / (hi - lo)).astype(np.float32))
    return out


if __name__ == "__main__":
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/head_detector/assets"
    dst = sys.argv[2] if len(sys.argv) > 2 else os.path.join(root, "head_detector_b200", "assets", "flame_indices.npz")
    print(convert(src, dst))
