"""Self-consistency of the conv-network oracle (parity unpinned at the super_gradients boundary)
and of the product's execution plan / weight packing against it - all on CPU."""
import numpy as np
import pytest
import torch

import plan_emulator as pe
from head_detector_b200 import arch
from oracle import net_oracle as no


def test_layer_enumeration_matches_survey():
    assert len(no.conv_layer_list()) == 191
    assert abs(no.total_macs(640) / 1e9 - 83.341) < 0.001
    assert abs(no.total_macs(1280) / no.total_macs(640) - 4.0) < 1e-9
    assert arch.total_macs(640) == no.total_macs(640)
    assert len(arch.conv_names()) == 191
    assert {n for n, *_ in arch.conv_names()} == {n for n, *_ in no.conv_layer_list()}


def test_qarepvgg_fold_exact_in_fp64():
    g = torch.Generator().manual_seed(0)
    for cin, cout, stride, residual, alpha in [(16, 16, 1, True, False), (8, 24, 2, False, False), (12, 12, 1, False, True)]:
        m = no.QARepVGGUnfused(cin, cout, stride, residual, alpha).double().eval()
        for bn in (m.bn3, m.post_bn):
            no.randomize_bn_(bn, g)
        if alpha:
            m.alpha.data.fill_(0.7)
        x = torch.randn(2, cin, 10, 10, generator=g, dtype=torch.float64)
        w, b = no.fold_qarepvgg(m)
        with torch.no_grad():
            ref = m(x)
        got = torch.relu(torch.nn.functional.conv2d(x, w, b, stride=stride, padding=1))
        assert (ref - got).abs().max() < 1e-12


def test_anchor_count_and_order():
    pts, strides = no.anchor_points([(80, 80), (40, 40), (20, 20)])
    assert pts.shape == (8400, 2) and strides[0] == 8 and strides[-1] == 32
    assert pts[0].tolist() == [0.5, 0.5] and pts[1].tolist() == [1.5, 0.5] and pts[80].tolist() == [0.5, 1.5]


@pytest.mark.parametrize("act_dtype", ["bf16", "fp16"])
def test_plan_and_packing_reproduce_oracle(act_dtype):
    """Interpret the product's plan with its packed 16-bit weights (either storage format) on CPU; compare with the oracle."""
    S = 128
    w = no.synthetic_weights(3)
    pk = arch.pack(arch.build_plan(S, act_dtype=act_dtype), w)
    dt = arch.ACT_DTYPES[act_dtype]
    torch.manual_seed(0)
    img = torch.randint(0, 256, (2, S, S, 3), dtype=torch.uint8)
    wq = {k: (v.to(dt).float() if k.endswith(".w") else v) for k, v in w.items()}
    wq["stem.w"] = (w["stem.w"] / 255.0).to(dt).float() * 255.0  # the packed stem weights carry the /255
    taps = {}
    with torch.no_grad():
        no.DeployNet(wq).forward(img.permute(0, 3, 1, 2).float() / 255, taps)
        bufs = pe.run_plan(pk, img, emulate_bf16=False)
    for l in range(3):
        reg, cls, t = pe.raw_to_oracle_layout(pk, bufs, l)
        oreg, ocls, ot = taps["raw"][l]
        assert (reg - oreg).abs().max() < 1e-4 and (cls - ocls).abs().max() < 1e-4
        for k in t:
            assert (t[k] - ot[k]).abs().max() < 1e-4, k
    for name in ("c2", "c3", "c4", "c5", "p3", "p4", "p5"):
        assert (bufs[pk.plan.buf_names[name]].permute(0, 3, 1, 2) - taps[name]).abs().max() < 1e-4


def test_decode_channel_rotation():
    """o[400:409] = [c[403:409], c[400:403]] (bug-compatible layout)."""
    H = [(2, 2), (1, 1), (1, 1)]
    raw = []
    for h, w in H:
        t = {"shape": torch.zeros(1, 128, h, w), "expr": torch.zeros(1, 64, h, w), "rot": torch.arange(6.).view(1, 6, 1, 1).expand(1, 6, h, w) + 10,
             "jaw": torch.arange(3.).view(1, 3, 1, 1).expand(1, 3, h, w) + 20, "scale": torch.zeros(1, 1, h, w), "transl": torch.zeros(1, 3, h, w)}
        raw.append((torch.zeros(1, 68, h, w), torch.zeros(1, 1, h, w), t))
    boxes, scores, fl = no.decode_heads(raw)
    assert fl.shape == (1, 6, 413)
    assert fl[0, 0, 400:409].tolist() == [13, 14, 15, 20, 21, 22, 10, 11, 12]
    assert fl[0, 0, 412].item() == 8 / 0.05 and fl[0, 4, 412].item() == 16 / 0.05
    assert fl[0, 1, 409].item() == 1.5 * 8 and fl[0, 2, 410].item() == 1.5 * 8
    assert abs(scores[0, 0, 0].item() - 0.5) < 1e-7
    assert torch.allclose(boxes[0, 0], torch.tensor([0.5 - 8, 0.5 - 8, 0.5 + 8, 0.5 + 8]) * 8)


def test_product_fold_matches_oracle_fold_and_unfused_forward():
    """arch.fold_qarepvgg / fold_conv_bn / deploy_from_unfused (what a real checkpoint goes through)
    against the oracle's unfused blocks."""
    g = torch.Generator().manual_seed(4)
    for cin, cout, stride, residual, alpha in [(16, 16, 1, True, False), (8, 24, 2, False, False), (12, 12, 1, False, True)]:
        m = no.QARepVGGUnfused(cin, cout, stride, residual, alpha).eval()
        for bn in (m.bn3, m.post_bn):
            no.randomize_bn_(bn, g)
        sd = {"blk." + k: v.detach() for k, v in m.state_dict().items()}
        sd["blk.residual"] = torch.tensor(m.residual)
        w, b = arch.fold_qarepvgg(sd, "blk")
        ow, ob = no.fold_qarepvgg(m)
        assert torch.allclose(w, ow, atol=1e-6) and torch.allclose(b, ob, atol=1e-6)
        x = torch.randn(1, cin, 9, 9, generator=g)
        with torch.no_grad():
            assert (m(x) - torch.relu(torch.nn.functional.conv2d(x, w, b, stride=stride, padding=1))).abs().max() < 1e-4
    bn = torch.nn.BatchNorm2d(6, eps=1e-6).eval()
    no.randomize_bn_(bn, g)
    conv_w = torch.randn(6, 4, 1, 1, generator=g)
    sd = {"c.conv.weight": conv_w, **{"c.bn." + k: v for k, v in bn.state_dict().items()}}
    w, b = arch.fold_conv_bn(sd, "c")
    ow, ob = no.fold_conv_bn(conv_w, bn)
    assert torch.allclose(w, ow.detach(), atol=1e-6) and torch.allclose(b, ob.detach(), atol=1e-6)
    # whole-network plumbing: plain tensors pass through, shapes are checked
    dw = no.synthetic_weights(1)
    out = arch.deploy_from_unfused(dw)
    assert set(out) == set(dw) and all(torch.equal(out[k], dw[k].float()) for k in dw if k.endswith(".w"))
    bad = dict(dw)
    bad["stem.w"] = torch.zeros(48, 3, 5, 5)
    import pytest
    with pytest.raises(ValueError):
        arch.deploy_from_unfused(bad)


@pytest.mark.parametrize("act_dtype", ["fp16", "bf16"])
def test_sparse_heads_plan_equals_dense_plan_at_survivors(act_dtype):
    """Two-phase plan (FLAME branch on 8x8 survivor patches after NMS) against the dense plan, both interpreted
    on the CPU: identical raw FLAME rows at the survivors - interior anchors, anchors on the image border / in
    the corners (the patch mask reproduces the dense graph's zero padding layer by layer) and vertically stacked
    neighbours - and an untouched box branch."""
    import plan_emulator as pe
    from head_detector_b200 import arch

    S, B, K = 128, 2, 6
    w = no.synthetic_weights(3)
    dense = arch.pack(arch.build_plan(S, act_dtype=act_dtype), w)
    sparse = arch.pack(arch.build_plan(S, sparse_heads=(B, K), act_dtype=act_dtype), w)
    assert sparse.plan.n_dense_ops is not None and all(op.level > 0 for op in sparse.plan.ops[sparse.plan.n_dense_ops:])
    assert all(op.level == 0 for op in sparse.plan.ops[:sparse.plan.n_dense_ops])
    torch.manual_seed(0)
    img = torch.randint(0, 256, (B, S, S, 3), dtype=torch.uint8)
    patches = {1: [(0, 0, 0), (1, 15, 7), (0, 8, 8), (1, 3, 15), (0, 15, 15)], 2: [(0, 0, 3), (1, 7, 7), (1, 4, 4)], 3: [(1, 0, 0), (0, 3, 2), (0, 1, 3)]}
    with torch.no_grad():
        ref = pe.run_plan(dense, img, emulate_bf16=True)
        got = pe.run_plan(sparse, img, emulate_bf16=True, patches=patches)
    for l in (1, 2, 3):
        assert torch.equal(ref[dense.plan.reg_buf[l - 1]], got[sparse.plan.reg_buf[l - 1]])
        fd, fs = ref[dense.plan.flame_buf[l - 1]], got[sparse.plan.flame_buf[l - 1]]
        for pi, (b, y, x) in enumerate(patches[l]):
            want, have = fd[b, y, x], fs[0, pi * arch.PATCH + arch.PATCH_C, arch.PATCH_C]
            assert (want - have).abs().max().item() <= 1e-5 * max(1.0, want.abs().max().item()), (l, b, y, x)


def test_act_dtype_selection_and_fp16_range_check(monkeypatch):
    """Storage format of the throughput mode: fp16 by default, $VGGHEADS_B200_ACT selects, anything else is
    refused; packing refuses weights that do not fit fp16 instead of writing infinities; the fp16 pack is the closer one."""
    monkeypatch.delenv("VGGHEADS_B200_ACT", raising=False)
    assert arch.default_act_dtype() == "fp16"
    monkeypatch.setenv("VGGHEADS_B200_ACT", "bf16")
    assert arch.default_act_dtype() == "bf16"
    monkeypatch.setenv("VGGHEADS_B200_ACT", "fp8")
    with pytest.raises(ValueError):
        arch.default_act_dtype()
    w = no.synthetic_weights(3)
    packs = {dt: arch.pack(arch.build_plan(64, act_dtype=dt), w) for dt in ("bf16", "fp16")}
    assert packs["bf16"].weights.shape == packs["fp16"].weights.shape and (packs["bf16"].bias == packs["fp16"].bias).all()
    m = packs["fp16"].op_meta[1]
    sl = slice(m["w_off"], m["w_off"] + m["n_pad"] * m["k_total"])
    ref = torch.from_numpy(packs["bf16"].weights[sl].view(np.int16).copy()).view(torch.bfloat16).double()
    got = torch.from_numpy(packs["fp16"].weights[sl].view(np.int16).copy()).view(torch.float16).double()
    assert (ref - got).abs().max() <= 2.0 ** -8 * ref.abs().max()          # the same matrix, each within its own rounding
    big = dict(w)
    big["stage1.down.w"] = w["stage1.down.w"] * 1e6
    arch.pack(arch.build_plan(64, act_dtype="bf16"), big)                  # fits bf16's range
    with pytest.raises(ValueError, match="fp16"):
        arch.pack(arch.build_plan(64, act_dtype="fp16"), big)
    with pytest.raises(AssertionError):
        arch.split_plan(arch.build_plan(64, act_dtype="fp16", fused_stem=False))   # the parity mode splits into bf16 terms
