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
