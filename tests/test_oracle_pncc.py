"""PNCC / refined-bbox restatements (oracle/pncc_oracle.py) against the reference: its C++ rasteriser compiled in place
(oracle/_ref, build container only) and tests/golden/pncc_ref.npz (reference `PredictionResult.get_pncc()`); plus the
host-side pieces of the result object (draw methods, aligned crops) against the same fixture.  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import pncc_oracle, sim3dr_ref

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _tables():
    from head_detector_b200 import mesh

    return mesh.tables()


def test_precomputed_tables_equal_what_the_reference_builds():
    t = _tables()
    keep = np.isin(t["full_faces"], t["head_w_ears"]).all(axis=1)          # PNCCProcessor.__init__ (pncc_processor.py:59-64)
    assert np.array_equal(t["pncc_triangles"], t["full_faces"][keep]) and t["pncc_triangles"].shape == (6814, 3)
    z = np.load(os.path.join(os.path.dirname(GOLD), "..", "head_detector_b200", "assets", "flame_generic.npz"))
    assert np.array_equal(z["faces"], t["full_faces"])                      # MeshSaver's full_faces.npy == the FLAME faces
    assert np.abs(pncc_oracle.ncc_colors(z["v_template"].astype(np.float64), t["head_w_ears"]) - t["ncc_colors"]).max() < 1e-6
    assert t["ncc_colors"][t["head_w_ears"]].min() >= 0 and t["ncc_colors"][t["head_w_ears"]].max() <= 1


def test_oracle_rasteriser_reproduces_the_reference_fixture():
    z = np.load(os.path.join(GOLD, "pncc_ref.npz"))
    t = _tables()
    po = pncc_oracle.PNCCOracle.__new__(pncc_oracle.PNCCOracle)
    po.triangles, po.colors = t["pncc_triangles"], t["ncc_colors"]
    got = po((480, 640, 3), list(z["vertices"]))                           # four overlapping heads: a few seconds in numpy
    assert (got.sum(2) != 0).sum() > 10000 and np.array_equal(got, z["pncc"])
    for v, want in zip(z["vertices"], z["bbox"]):
        assert list(pncc_oracle.refined_head_bbox(v, t["head_indices"])) == want.tolist()


@pytest.mark.skipif(not sim3dr_ref.available(), reason="oracle/_ref/libsim3dr_ref.so is built in the build container only")
def test_oracle_rasteriser_equals_reference_cpp():
    z = np.load(os.path.join(GOLD, "pncc_ref.npz"))
    t = _tables()
    po = pncc_oracle.PNCCOracle.__new__(pncc_oracle.PNCCOracle)
    po.triangles, po.colors = t["pncc_triangles"], t["ncc_colors"]
    heads = [z["vertices"][0], z["vertices"][3]]                           # the overlapping pair: painter's order matters
    a = po((480, 640, 3), heads)
    b = po((480, 640, 3), heads, raster=sim3dr_ref.rasterize)
    assert np.array_equal(a, b)
    full = po((480, 640, 3), list(z["vertices"]), raster=sim3dr_ref.rasterize)
    assert np.array_equal(full, z["pncc"])


def _result_from_fixture(z):
    from head_detector_b200.detection_result import PredictionResult
    from head_detector_b200.head_info import Bbox, FlameParams, HeadMetadata
    from head_detector_b200.utils import calculate_rpy

    frame = np.random.default_rng(int(z["frame_seed"])).integers(0, 256, (480, 640, 3), dtype=np.uint8)
    heads = []
    for v, p, bb in zip(z["vertices"], torch.from_numpy(z["params"]), z["bbox"]):
        fp = FlameParams.from_3dmm(p[None])
        heads.append(HeadMetadata(bbox=Bbox(*[int(t) for t in bb]), score=0.9, flame_params=fp, vertices_3d=v.copy(), head_pose=calculate_rpy(fp)))
    return frame, PredictionResult(frame, heads)


def test_draw_methods_and_aligned_crops_match_the_reference():
    z = np.load(os.path.join(GOLD, "pncc_ref.npz"))
    frame, res = _result_from_fixture(z)
    import inspect

    assert inspect.signature(res.draw).parameters["method"].default == "full"      # detection_result.py:44
    for method in ("full", "bbox", "landmarks", "points", "pose"):
        img = res.draw(method)
        m = (img != frame).any(2)
        assert np.array_equal(np.argwhere(m).astype(np.int16), z[f"draw_{method}_yx"]), method
        assert np.array_equal(img[m], z[f"draw_{method}_rgb"]), method
    with pytest.raises(KeyError):
        res.draw("nope")
    crops = res.get_aligned_heads()
    assert len(crops) == 4
    for i, c in enumerate(crops):
        assert np.array_equal(c, z[f"crop{i}"]), i
