"""Device mesh consumers (f4): PNCC rasteriser and refined head boxes against the reference's own outputs
(tests/golden/pncc_ref.npz: PNCCProcessor + Sim3DR C++) - bit-exact - and against the numpy oracle on other inputs."""
import os

import numpy as np
import pytest
import torch

from oracle import pncc_oracle

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_pncc_image_is_bit_identical_to_the_reference():
    from head_detector_b200 import mesh

    z = np.load(os.path.join(GOLD, "pncc_ref.npz"))
    got = mesh.pncc_image(480, 640, list(z["vertices"]))
    assert got.dtype == np.uint8 and got.shape == (480, 640, 3)
    assert np.array_equal(got, z["pncc"]), int((got != z["pncc"]).any(2).sum())
    # order matters where heads overlap: reversed order paints the first head on top
    rev = mesh.pncc_image(480, 640, list(z["vertices"][::-1]))
    assert not np.array_equal(rev, got) and (rev.sum(2) != 0).sum() == (got.sum(2) != 0).sum()
    assert mesh.pncc_image(480, 640, []).sum() == 0


def test_pncc_partially_outside_and_tiny_frames_vs_oracle():
    from head_detector_b200 import mesh

    z = np.load(os.path.join(GOLD, "pncc_ref.npz"))
    t = mesh.tables()
    po = pncc_oracle.PNCCOracle.__new__(pncc_oracle.PNCCOracle)
    po.triangles, po.colors = t["pncc_triangles"], t["ncc_colors"]
    v = z["vertices"][1].copy()
    v[:, 0] -= v[:, 0].min() + 40.0          # a third of the head left of the frame
    v[:, 1] -= v[:, 1].min() - 5.0
    want = po((96, 128, 3), [v])
    got = mesh.pncc_image(96, 128, [v])
    assert (want.sum(2) != 0).sum() > 500 and np.array_equal(got, want)


def test_get_pncc_keeps_the_reference_in_place_flip():
    from test_oracle_pncc import _result_from_fixture

    z = np.load(os.path.join(GOLD, "pncc_ref.npz"))
    _, res = _result_from_fixture(z)
    img = res.get_pncc()
    assert np.array_equal(img, z["pncc"])
    assert np.array_equal(np.stack([h.vertices_3d[:, 2] for h in res.heads]), z["z_after"])    # z negated in place (pncc_processor.py:70)


def test_refined_head_bboxes_device():
    from head_detector_b200 import mesh

    z = np.load(os.path.join(GOLD, "pncc_ref.npz"))
    got = mesh.refined_head_bboxes(torch.from_numpy(z["vertices"])).cpu().numpy()
    assert got.tolist() == z["bbox"].tolist()
    neg = z["vertices"].copy()
    neg[..., :2] -= 400.5                      # negative coordinates: int() truncates towards zero
    want = [list(pncc_oracle.refined_head_bbox(v, mesh.tables()["head_indices"])) for v in neg]
    assert mesh.refined_head_bboxes(torch.from_numpy(neg)).cpu().numpy().tolist() == want
