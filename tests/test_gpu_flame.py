"""CUDA FLAME decode (vgh_flame_decode through the python mirror) vs the oracle and golden vectors.
Tolerances: model-space vertices 2e-7 abs (fp64 accumulation on our side, fp32 on the reference's),
detector-space vertices 1e-4 abs (BASELINE.json north_star)."""
import os

import numpy as np
import pytest
import torch

from oracle import flame_oracle as fo
from oracle.make_golden import network_like_heads

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def flame():
    from head_detector_b200.flame import FLAMELayer

    return FLAMELayer()


@pytest.fixture(scope="module")
def consts():
    return fo.load_flame_constants()


def test_reference_fixture_1json(flame, golden_dir):
    g = np.load(os.path.join(golden_dir, "flame_1json.npz"))
    p = torch.tensor(g["params"], dtype=torch.float32)[None].clone()
    p[:, 409:412] = 0
    p[:, 412] = 1
    _, _, proj = flame.decode(p.cuda())
    assert np.abs(proj[0].cpu().numpy() - g["vertices_3d"]).max() < 3e-7


def test_golden_reference_heads(flame, golden_dir):
    from head_detector_b200.flame import reproject_spatial_vertices

    g = np.load(os.path.join(golden_dir, "flame_ref_heads.npz"))
    v, R, proj = reproject_spatial_vertices(flame, torch.from_numpy(g["params"]).cuda(), to_2d=False)
    assert np.abs(v.cpu().numpy() - g["vertices"]).max() < 2e-7
    assert np.abs(R.cpu().numpy() - g["rotation"]).max() < 1e-6
    err = np.abs(proj.cpu().numpy() - g["projected"]).max()
    assert err < 1e-4, err
    v2, _, p2 = reproject_spatial_vertices(flame, torch.from_numpy(g["params"]).cuda(), to_2d=True)
    assert p2.shape == (6, 5023, 2) and torch.equal(p2, proj[..., :2])


@pytest.mark.parametrize("n", [1, 7, 16, 17, 70])
def test_random_heads_vs_oracle(flame, consts, n):
    p = network_like_heads(n, seed=100 + n)
    v, R, proj = flame.decode(p.cuda())
    ov, oR, oproj = fo.reproject(p, consts)
    assert (v.cpu() - ov).abs().max() < 2e-7
    assert (R.cpu() - oR).abs().max() < 1e-6
    assert (proj.cpu() - oproj).abs().max() < 1e-4
    # only the first 128/64 coefficients are live for network rows: same result, less work
    _, _, proj_live = flame.decode(p.cuda(), live=(128, 64))
    assert torch.equal(proj_live, proj)


def test_dense_coefficients_and_letterbox(flame, consts):
    g = torch.Generator().manual_seed(9)
    p = network_like_heads(5, seed=9)
    p[:, :300] = torch.randn(5, 300, generator=g)
    p[:, 300:400] = torch.randn(5, 100, generator=g)
    xf = torch.tensor([[0., 80., 0.5], [12., 0., 0.75], [0., 0., 1.0], [3., 4., 1.25], [100., 0., 0.3]])
    _, _, proj = flame.decode(p.cuda(), xform=xf.cuda())
    for i in range(5):
        ref = fo.detector_vertices(p[i:i + 1], consts, (xf[i, 0].item(), xf[i, 1].item()), xf[i, 2].item())
        tol = 1e-4 / min(1.0, xf[i, 2].item())
        assert (proj[i].cpu() - ref[0]).abs().max() < tol


def test_empty_batch(flame):
    from head_detector_b200.flame import reproject_spatial_vertices

    v, R, p = reproject_spatial_vertices(flame, torch.zeros(0, 413, device="cuda"), to_2d=False)
    assert v.shape == (0, 5023, 3) and R.shape == (0, 3, 3) and p.shape == (0, 5023, 3)
    with pytest.raises(ValueError):
        reproject_spatial_vertices(flame, torch.zeros(2, 412, device="cuda"))


def test_large_batch_properties(flame, consts):
    """2000 heads (BASELINE config-3 scale): translation equivariance, scale linearity, spot check."""
    n = 2000
    p = network_like_heads(n, seed=77)
    _, _, a = flame.decode(p.cuda(), live=(128, 64))
    q = p.clone()
    q[:, 409:412] += torch.tensor([16.0, -32.0, 8.0])
    _, _, b = flame.decode(q.cuda(), live=(128, 64))
    d = (b - a).cpu()
    assert (d - torch.tensor([16.0, -32.0, 8.0])).abs().max() < 2e-4
    idx = torch.tensor([0, 999, 1999])
    _, _, ref = fo.reproject(p[idx], consts)
    assert (a[idx.cuda()].cpu() - ref).abs().max() < 1e-4


def test_layer_forward_api(flame, consts):
    from head_detector_b200.head_info import FlameParams

    p = network_like_heads(3, seed=4)
    fp = FlameParams.from_3dmm(p.cuda())
    v = flame.forward(fp, zero_rot=True)
    assert (v.cpu() - fo.flame_vertices(p, consts)).abs().max() < 2e-7
