"""Host-side mirror of the reference interface (no GPU needed)."""
import os

import numpy as np
import pytest
import torch

from head_detector_b200.head_info import FLAME_CONSTS, FlameParams
from head_detector_b200.utils import limit_angle, rot_mat_from_6dof, rpy_from_rotations
from oracle import flame_oracle as fo


def test_from_3dmm_layout_and_error():
    p = torch.arange(413.)[None]
    f = FlameParams.from_3dmm(p)
    assert f.shape.shape == (1, 300) and f.expression.shape == (1, 100)
    assert f.jaw[0].tolist() == [400, 401, 402] and f.rotation[0].tolist() == [403, 404, 405, 406, 407, 408]
    assert f.translation[0].tolist() == [409, 410, 411] and f.scale[0].tolist() == [412]
    assert f.eyeballs.shape == (1, 0) and f.neck.shape == (1, 0)
    with pytest.raises(ValueError):
        FlameParams.from_3dmm(torch.zeros(1, 412))
    assert sum(FLAME_CONSTS.values()) == 413


def test_roundtrip_is_not_identity():
    """to_3dmm_tensor(from_3dmm(x)) swaps the jaw/rot blocks (SURVEY section 7 gotcha 1)."""
    p = torch.arange(413.)[None]
    q = FlameParams.from_3dmm(p).to_3dmm_tensor()
    assert q[0, 400:409].tolist() == [403, 404, 405, 406, 407, 408, 400, 401, 402]


def test_rot6d_matches_oracle_and_rpy_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "parse_ref.npz"))
    p = torch.from_numpy(g["params"])
    R = rot_mat_from_6dof(p[:, 403:409])
    assert torch.allclose(R, fo.rot6d_to_matrix(p[:, 403:409]))
    rpy = rpy_from_rotations(R.numpy())
    assert np.allclose(np.array([[a.roll, a.pitch, a.yaw] for a in rpy]), g["rpy"], atol=1e-4)


def test_limit_angle():
    assert limit_angle(190.0) == -170.0 and limit_angle(-190.0) == 170.0 and limit_angle(10.0) == 10.0


def test_letterbox_matches_reference(golden_dir):
    """`_transform_image` (reference detector.py:40-52: Lanczos4 resize, centred pad 127) - same pixels,
    same padding / scale bookkeeping as the unmodified reference run in the build container."""
    import hashlib

    from head_detector_b200.detector import HeadDetector

    g = np.load(os.path.join(golden_dir, "letterbox_ref.npz"))
    det = object.__new__(HeadDetector)
    det._image_size = 640
    src = np.random.default_rng(int(g["seed"])).integers(0, 256, tuple(g["shape"]), dtype=np.uint8)
    img, pad, scale = det._transform_image(src)
    assert img.shape == (640, 640, 3) and img.dtype == np.uint8
    assert tuple(pad) == tuple(g["pad"]) and abs(scale - float(g["scale"])) < 1e-12
    assert np.array_equal(img[::37, ::41], g["probe"])
    assert np.array_equal(np.frombuffer(hashlib.sha1(img.tobytes()).digest(), dtype=np.uint8), g["sha1"])


def test_batch_parse_equals_per_image_reference_loop():
    """`HeadDetector._parse_batch` (vectorised over the batch) against a literal per-image transcription of the
    reference's `_parse_predictions` (detector.py:61-90): same integer boxes, scores, vertices, roll/pitch/yaw,
    parameter views and rescaled `scale`, including images without heads and per-image padding / scale."""
    import torch
    from scipy.spatial.transform import Rotation

    from head_detector_b200.detector import HeadDetector
    from head_detector_b200.head_info import Bbox, FlameParams, HeadMetadata
    from head_detector_b200.utils import rpy_from_rotations

    def reference_loop(out, img, cache, S=640):
        pad, scale = cache["padding"], cache["scale"]
        lo, hi = int(out["offsets"][img]), int(out["offsets"][img + 1])
        n = hi - lo
        boxes = out["keep_boxes"][img, :n].numpy().copy()
        scores = out["keep_scores"][img, :n].numpy()
        verts, params, rots = out["vertices"][lo:hi].numpy(), out["params"][lo:hi], out["rotations"][lo:hi].numpy()
        boxes = boxes.clip(0, S)                                  # detector.py:70
        boxes[:, [0, 2]] -= pad[0]                                # :71
        boxes[:, [1, 3]] -= pad[1]                                # :72
        boxes /= scale                                            # :73
        boxes = np.rint(boxes).astype(int)                        # :74
        poses = rpy_from_rotations(rots)
        heads = []
        for i in range(n):
            fp = FlameParams.from_3dmm(params[i:i + 1])           # :78
            fp.scale = fp.scale / scale                           # :79
            b = boxes[i]
            heads.append(HeadMetadata(bbox=Bbox(x=b[0], y=b[1], w=b[2] - b[0], h=b[3] - b[1]), score=scores[i],
                                      flame_params=fp, vertices_3d=verts[i], head_pose=poses[i]))
        return heads

    B = 9
    g = torch.Generator().manual_seed(0)
    cnt = torch.randint(0, 7, (B,), generator=g)
    cnt[2] = 0
    n = int(cnt.sum())
    out = {"offsets": torch.cat([torch.zeros(1, dtype=torch.int64), cnt.cumsum(0)]).int(),
           "keep_boxes": torch.rand(B, 100, 4, generator=g) * 700 - 30, "keep_scores": torch.rand(B, 100, generator=g),
           "vertices": torch.rand(n, 50, 3, generator=g), "params": torch.randn(n, 413, generator=g),
           "rotations": torch.from_numpy(Rotation.random(n, random_state=1).as_matrix()).float()}
    caches = [{"padding": ((i % 3) * 40, (i % 5) * 17), "scale": 640 / (700 + 13 * i)} for i in range(B)]
    det = object.__new__(HeadDetector)
    det._image_size = 640
    got = det._parse_batch(out, caches)
    assert [len(h) for h in got] == cnt.tolist()
    for i in range(B):
        want = reference_loop(out, i, caches[i])
        lo, hi = int(out["offsets"][i]), int(out["offsets"][i + 1])
        one = det._parse_predictions(out["keep_boxes"][i, :hi - lo], out["keep_scores"][i, :hi - lo], out["params"][lo:hi],   # reference signature
                                     dict(caches[i], _decoded=(out["vertices"][lo:hi], out["rotations"][lo:hi])))
        assert len(one) == len(want) and all(np.array_equal(a.vertices_3d, b.vertices_3d) for a, b in zip(one, want))
        for a, b in zip(want, got[i]):
            assert tuple(int(v) for v in a.bbox) == tuple(int(v) for v in b.bbox)
            assert a.score == b.score and np.array_equal(a.vertices_3d, b.vertices_3d) and a.head_pose == b.head_pose
            for k in ("shape", "expression", "jaw", "rotation", "translation", "scale", "eyeballs", "neck"):
                x, y = getattr(a.flame_params, k), getattr(b.flame_params, k)
                assert x.shape == y.shape and torch.equal(x, y), k
