"""The as-trained network restatement (oracle/sg_net.py), the checkpoint loader (head_detector_b200/weights.py) and the
golden vectors produced by the reference's OWN head / decode / top-k / detector code (oracle/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from head_detector_b200 import arch, weights
from oracle import net_oracle, nms_oracle, ref_heads, sg_net

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def net():
    return sg_net.build(0)


@pytest.fixture(scope="module")
def deploy(net):
    return weights.deploy_from_state_dict(net.state_dict())


def test_parameter_count_is_the_surveys(net):
    # 54.54 M unfused parameters, 191 convs
    assert abs(sum(p.numel() for p in net.parameters()) / 1e6 - 54.54) < 0.01
    assert len(weights.layer_map()) == 191 == len(arch.conv_names())
    assert {n for n, _, _ in weights.layer_map()} == {n for n, *_ in arch.conv_names()}


def test_loader_fold_equals_unfused_network(net, deploy):
    x = torch.rand(2, 3, 128, 128, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        b, s, f = net(x)
        ob, os_, of = net_oracle.DeployNet(deploy).forward(x)
    assert (ob - b).abs().max() < 2e-3 and (os_ - s).abs().max() < 1e-5
    assert ((of - f).abs() / (f.abs() + 1)).max() < 2e-4
    # the same in fp64: the fold algebra itself is exact (fp32 rounding of the folded weights remains)
    net64 = sg_net.build(0).double()
    d64 = {k: v.double() for k, v in weights.deploy_from_state_dict(net64.state_dict()).items()}
    with torch.no_grad():
        b64, _, f64 = net64(x.double())
        ob, _, of = net_oracle.DeployNet(d64).forward(x.double())
    assert (ob - b64).abs().max() < 2e-4 and ((of - f64).abs() / (f64.abs() + 1)).max() < 2e-5


def test_loader_is_strict_and_prefix_agnostic(net, deploy):
    sd = net.state_dict()
    wrapped = {"model." + k: v for k, v in sd.items()}       # ConvertableCompletePipelineModel keeps the detector under .model
    again = weights.deploy_from_state_dict(wrapped)
    assert all(torch.equal(again[k], v) for k, v in deploy.items())
    broken = dict(sd)
    del broken["backbone.stage2.blocks.bottlenecks.1.cv1.branch_1x1.weight"]
    with pytest.raises(KeyError, match="lacks"):
        weights.deploy_from_state_dict(broken)
    extra = dict(sd)
    extra["neck.neck5.conv.conv.weight"] = torch.zeros(1)
    with pytest.raises(KeyError, match="does not use"):
        weights.deploy_from_state_dict(extra)
    assert "stem.w" in weights.deploy_from_state_dict(extra, strict=False)
    bad = dict(sd)
    bad["heads.head1.reg_pred.weight"] = torch.zeros(60, 128, 1, 1)
    with pytest.raises(ValueError, match="shape"):
        weights.deploy_from_state_dict(bad)
    with pytest.raises(KeyError, match="not a YoloHeads checkpoint"):
        weights.deploy_from_state_dict({"foo.weight": torch.zeros(1)})
    with pytest.raises(FileNotFoundError, match="synthetic"):
        weights.resolve(None)
    assert "stem.w" in weights.resolve("synthetic")


def test_fused_reparam_checkpoints_load(net, deploy):
    """A checkpoint exported with fused blocks carries one `rbr_reparam` conv (+ post_bn when partially fused)."""
    sd = dict(net.state_dict())
    p = "backbone.stage1.blocks.bottlenecks.0.cv1"
    w, b = weights._fold_qarep(weights._Used({k: v for k, v in sd.items()}), p, True, arch.BN_EPS)
    for k in [k for k in sd if k.startswith(p + ".")]:
        del sd[k]
    sd[p + ".rbr_reparam.weight"], sd[p + ".rbr_reparam.bias"] = w, b
    again = weights.deploy_from_state_dict(sd)
    assert torch.allclose(again["stage1.csp.b0.cv1.w"], deploy["stage1.csp.b0.cv1.w"], atol=1e-6)


def test_torchscript_blob_roundtrip(tmp_path, net, deploy):
    """`vgg_heads_l.trcd` is a TorchScript trace (detector.py:25-30): trace -> jit.load -> state_dict -> deploy dict."""
    path = sg_net.trace_to(str(tmp_path / "vgg_heads_l.trcd"), net, image_size=64)
    again = weights.load_checkpoint(path)
    assert set(again) == set(deploy) and all(torch.equal(again[k], v) for k, v in deploy.items())
    torch.save({"ema_net": net.state_dict()}, str(tmp_path / "ckpt.pth"))
    again = weights.load_checkpoint(str(tmp_path / "ckpt.pth"))
    assert all(torch.equal(again[k], v) for k, v in deploy.items())


@pytest.mark.skipif(not ref_heads.available(), reason="needs /root/reference (build container)")
def test_restated_heads_equal_the_reference_classes(net):
    ns = ref_heads.load()
    ref = ref_heads.build_heads(ns)
    ref.load_state_dict({k[len("heads."):]: v for k, v in net.state_dict().items() if k.startswith("heads.")}, strict=True)
    g = torch.Generator().manual_seed(3)
    feats = [torch.randn(2, c, 128 // s, 128 // s, generator=g) for c, s in ((96, 8), (192, 16), (384, 32))]
    with torch.no_grad():
        dec, _ = ref(feats)
        b, s, f = net.heads(feats)
    assert torch.equal(dec.boxes_xyxy, b) and torch.equal(dec.scores, s) and torch.equal(dec.flame_params, f)


def _gold_heads():
    from oracle.make_golden import HEADS_SEED, heads_feats

    z = np.load(os.path.join(GOLD, "heads_ref.npz"))
    return z, sg_net.build(HEADS_SEED), heads_feats()


def test_heads_fixture_pins_oracle_decode_and_deploy_heads():
    """heads_ref.npz comes from the reference's YoloHeadsNDFLHeads.  (1) The deploy-form oracle heads fed with the loader's
    folded weights reproduce the reference's raw outputs; (2) `net_oracle.decode_heads` applied to the REFERENCE's raw
    outputs reproduces the reference's decoded boxes / scores / flame - pins a4, a5."""
    z, net, feats = _gold_heads()
    dw = weights.deploy_from_state_dict(net.state_dict())
    dn = net_oracle.DeployNet(dw)
    with torch.no_grad():
        raw = dn.raw_heads(feats)
    for l, (reg, cls, t) in enumerate(raw, start=1):
        assert np.abs(reg.numpy() - z[f"reg{l}"]).max() < 2e-4 and np.abs(cls.numpy() - z[f"cls{l}"]).max() < 2e-4
        for tw in ("shape", "expr", "rot", "jaw", "transl", "scale"):
            assert np.abs(t[tw].numpy() - z[f"{tw}{l}"]).max() < 2e-4, (l, tw)
    ref_raw = [(torch.from_numpy(z[f"reg{l}"]), torch.from_numpy(z[f"cls{l}"]),
                {tw: torch.from_numpy(z[f"{tw}{l}"]) for tw in ("shape", "expr", "rot", "jaw", "transl", "scale")}) for l in (1, 2, 3)]
    b, s, f = net_oracle.decode_heads(ref_raw)
    assert np.abs(b.numpy() - z["boxes"]).max() < 1e-4 and np.abs(s.numpy() - z["scores"]).max() < 1e-6
    assert (np.abs(f.numpy() - z["flame"]) / (np.abs(z["flame"]) + 1)).max() < 1e-6
    # the channel rotation of 400..408 is really there: head order [rot6|jaw3] -> output [..jaw-slot <- rot[3:6]..]
    fl1 = z["flame1"]     # level 1, head order, [B,413,H,W]
    assert np.array_equal(z["flame"][0, 0, 400:403], fl1[0, 403:406, 0, 0]) and np.array_equal(z["flame"][0, 0, 406:409], fl1[0, 400:403, 0, 0])


def test_topk_fixture_pins_select_nms():
    """Reference VGGHeadDecodingModule (top-k 1000) + utils.nms == the oracle's select_nms on the undecimated anchors (a6)."""
    z = np.load(os.path.join(GOLD, "topk_ref.npz"))
    for b in range(2):
        keep = nms_oracle.select_nms(z["boxes"][b], z["scores"][b], 0.5, 0.5, 1000, 100)
        assert keep.tolist() == z[f"keep{b}"].tolist()
        order = np.argsort(-z["scores"][b], kind="stable")[:1000]
        assert order.tolist() == z["topk_ids"][b].tolist()


def test_detector_fixture_vs_port():
    """detector_ref.npz = the UNMODIFIED HeadDetector on the synthetic blob.  The oracle port (deploy-form net + utils.nms
    restatement + FLAME restatement) on the same frame finds the same heads - so the CPU arm of bench.py measures the
    reference's algorithm."""
    import cv2

    from oracle import flame_oracle
    from oracle.make_golden import detector_image, detector_net

    z = np.load(os.path.join(GOLD, "detector_ref.npz"))
    img = detector_image()
    assert img.shape == (480, 640, 3)
    lb = cv2.copyMakeBorder(img, 80, 80, 0, 0, cv2.BORDER_CONSTANT, value=127)   # 480x640 -> pad (0, 80), scale 1 (no resize)
    x = torch.from_numpy(lb).permute(2, 0, 1)[None].float() / 255.0
    dw = weights.deploy_from_state_dict(detector_net().state_dict())
    with torch.no_grad():
        boxes, scores, flame = net_oracle.DeployNet(dw).forward(x)
    keep = nms_oracle.select_nms(boxes[0].numpy(), scores[0, :, 0].numpy(), 0.5, 0.5, 1000, 100)
    n = len(z["scores"])
    assert n > 10
    # fp32 fold vs unfused rounding may flip a borderline candidate: compare the overlapping top of the two lists
    m = min(n, len(keep))
    got_scores = scores[0, keep, 0].numpy()
    assert abs(n - len(keep)) <= 2 and np.abs(got_scores[:m - 2] - z["scores"][:m - 2]).max() < 1e-4
    kb = boxes[0, keep].numpy().clip(0, 640)
    kb[:, [1, 3]] -= 80
    xywh = np.stack([kb[:, 0], kb[:, 1], kb[:, 2] - kb[:, 0], kb[:, 3] - kb[:, 1]], 1)
    assert np.abs(np.rint(xywh[:m - 2]) - z["bbox_xywh"][:m - 2]).max() <= 1
    rows = flame[0][torch.from_numpy(keep[:4])]
    verts = flame_oracle.detector_vertices(rows, flame_oracle.load_flame_constants(), pad_xy=(0, 80), img_scale=1.0)
    assert np.abs(verts.numpy() - z["vertices_3d"][:4]).max() < 0.05   # pixels; fold rounding amplified by scale ~1e3
