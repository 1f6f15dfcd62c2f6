"""Conv network on the GPU (tcgen05 implicit GEMM + aux kernels) vs the CPU interpretation of the
same plan and vs the oracle network.  16-bit storage => comparisons allow rounding flips of the storage format:
|err| <= 2^-6 * max|ref| per buffer for bf16 (2^-9 for fp16, the default) and a mean error two orders below that."""
import numpy as np
import pytest
import torch

import plan_emulator as pe
from head_detector_b200 import arch
from oracle import nms_oracle, flame_oracle, net_oracle as no

pytestmark = pytest.mark.gpu


# max flip of one stored value relative to the buffer's max: bf16 keeps 8 significant bits, fp16 11
FLIP_TOL = {"bf16": (2 ** -6, 2e-3), "fp16": (2 ** -9, 2.5e-4)}


@pytest.fixture(scope="module", params=["fp16", "bf16"])
def small(request):
    from head_detector_b200.engine import Engine

    S, B = 128, 2
    w = no.synthetic_weights(3)
    eng = Engine(w, B, S, act_dtype=request.param)
    assert eng.act_dtype == request.param
    torch.manual_seed(0)
    img = torch.randint(0, 256, (B, S, S, 3), dtype=torch.uint8)
    eng.forward(img.cuda())
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = pe.run_plan(eng.packed, img, emulate_bf16=True)
    return eng, img, ref, w


def test_every_buffer_matches_cpu_interpretation(small):
    eng, img, ref, _ = small
    bad = []
    for name, i in eng.plan.buf_names.items():
        got, want = eng.read_buffer(name), ref[i]
        scale = want.abs().max().item() + 1e-6
        err = (got - want).abs()
        tol_max, tol_mean = FLIP_TOL[eng.act_dtype]
        if not (err.max().item() <= tol_max * scale + 1e-5 and err.mean().item() <= tol_mean * scale):
            bad.append((name, err.max().item(), err.mean().item(), scale))
    assert not bad, bad


def test_decode_kernels_vs_oracle_decode(small):
    """boxes / scores / dense flame from the GPU's own raw head outputs (stage-wise, same inputs)."""
    eng, _, _, _ = small
    raw = []
    for l in range(3):
        reg = eng.read_buffer(f"head{l + 1}.reg_raw").permute(0, 3, 1, 2)
        fl = eng.read_buffer(f"head{l + 1}.flame_raw").permute(0, 3, 1, 2)
        from head_detector_b200 import arch
        raw.append((reg[:, :68], reg[:, 68:69], {tw: fl[:, arch.RAW_ROW_OFF[tw]:arch.RAW_ROW_OFF[tw] + oc] for tw, _, oc in arch.TOWERS}))
    boxes, scores, flame = no.decode_heads(raw)
    assert (eng.boxes.cpu() - boxes).abs().max() < 1e-3
    assert (eng.scores.cpu() - scores[..., 0]).abs().max() < 1e-6
    dense = eng.dense_flame().cpu()
    assert dense.shape == flame.shape
    rel = (dense - flame).abs() / (flame.abs() + 1.0)
    assert rel.max() < 1e-5


@pytest.mark.parametrize("act_dtype", ["fp16", "bf16"])
def test_end_to_end_vs_oracle_network_640(act_dtype):
    """Whole network at the reference resolution vs the oracle (a) with weights / activations rounded to the same 16-bit
    storage format and (b) in plain fp32 - the precision cost of the throughput mode, per format.  Thresholds: ~3x the
    measured errors (fp16 is 8-12x closer to fp32 than bf16: 11 vs 8 significant bits)."""
    from head_detector_b200.engine import Engine

    w = no.synthetic_weights(0)
    eng = Engine(w, 1, 640, act_dtype=act_dtype)
    torch.manual_seed(1)
    img = torch.randint(0, 256, (1, 640, 640, 3), dtype=torch.uint8)
    boxes, scores = eng.forward(img.cuda())
    dense = eng.dense_flame().cpu()
    dt = arch.ACT_DTYPES[eng.act_dtype]
    wq = {k: (v.to(dt).float() if k.endswith(".w") else v) for k, v in w.items()}
    wq["stem.w"] = (w["stem.w"] / 255.0).to(dt).float() * 255.0  # the packed stem weights carry the /255
    x = img.permute(0, 3, 1, 2).float() / 255
    with torch.no_grad():
        same = no.DeployNet(wq, act_round=pe.act_round(eng.act_dtype)).forward(x)
        fp32 = no.DeployNet(w).forward(x)
    assert eng.A == 8400 and boxes.shape == (1, 8400, 4)
    #        boxes max px, boxes mean px, scores, 3*tanh coefficients, scale (relative)
    tol = {"bf16": (1.0, 0.05, 2e-3, 0.1, 0.05), "fp16": (0.25, 0.015, 3e-4, 0.03, 0.01)}[act_dtype]
    for tag, (ob, os_, of) in (("same-format oracle", same), ("fp32 oracle", fp32)):
        k = 1.0 if tag == "same-format oracle" else 2.0   # vs fp32 both sides' roundings add up
        err = (boxes.cpu() - ob).abs()                     # pixels, boxes span ~[-300, 900]
        rel_scale = ((dense[..., 412] - of[..., 412]).abs() / of[..., 412]).max().item()
        got = (err.max().item(), err.mean().item(), (scores.cpu() - os_[..., 0]).abs().max().item(),
               (dense[..., :400] - of[..., :400]).abs().max().item(), rel_scale)
        print(f"{act_dtype} vs {tag}: boxes max {got[0]:.4f} mean {got[1]:.5f} px, scores {got[2]:.2e}, coeffs {got[3]:.4f}, scale {got[4]:.5f}")
        assert all(g < k * t for g, t in zip(got, tol)), (act_dtype, tag, got, tol)


def test_pipeline_stagewise_parity_and_host_path():
    """forward -> (engineered scores) -> select/NMS -> survivor rows -> FLAME: NMS ids bit-exact vs
    the oracle on the same boxes/scores, vertices within 1e-4 of the oracle on the same rows, and
    the host-buffer graph path returns the same numbers as the staged path."""
    from head_detector_b200.engine import Engine
    from oracle.make_golden import clustered_anchors

    B, S = 3, 640
    eng = Engine(no.synthetic_weights(0), B, S)
    torch.manual_seed(2)
    img = torch.randint(0, 256, (B, S, S, 3), dtype=torch.uint8)
    ob, osc = zip(*[clustered_anchors(8400, 8, 12, seed=50 + i) for i in range(B)])
    ob, osc = torch.stack(ob).cuda(), torch.stack(osc).cuda()
    eng.set_override(ob, osc)
    xf = torch.tensor([[0., 0., 1.0], [0., 80., 0.5], [16., 0., 0.8]])
    eng.forward(img.cuda())
    eng.boxes.copy_(ob); eng.scores.copy_(osc)     # staged path: apply the override by hand
    eng.postprocess(0.5, 0.5, 1000, xf.cuda())
    torch.cuda.synchronize()
    cnt = eng.keep_cnt.cpu().numpy()
    idx = eng.keep_idx.cpu().numpy()
    off = eng.head_offsets.cpu().numpy()
    dense = eng.dense_flame().cpu()
    consts = flame_oracle.load_flame_constants()
    total = int(off[-1])
    assert total == cnt.sum() and total > 0
    params, verts = eng.head_params(total).cpu(), eng.head_verts(total).cpu()
    for b in range(B):
        want = nms_oracle.select_nms(ob[b].cpu().numpy(), osc[b].cpu().numpy())
        assert idx[b, :cnt[b]].tolist() == want.tolist()
        rows = params[off[b]:off[b + 1]]
        assert torch.equal(rows, dense[b][torch.from_numpy(want)])
        ref = flame_oracle.detector_vertices(rows, consts, (xf[b, 0].item(), xf[b, 1].item()), xf[b, 2].item())
        assert (verts[off[b]:off[b + 1]] - ref).abs().max() < 1e-4 / xf[b, 2].item()
    # host-buffer end-to-end (CUDA graph) must reproduce the staged results
    out = eng.alloc_host_outputs(B * 100)
    n = eng.run_host(img.pin_memory(), out, 0.5, 0.5, 1000, xf.pin_memory())
    assert n == total and out["keep_cnt"].numpy().tolist() == cnt.tolist()
    assert torch.equal(out["params"][:n], params) and torch.equal(out["verts"][:n], verts)
    assert eng.launch_count > 100


def test_head_detector_call_api():
    """`HeadDetector()(image)` -> PredictionResult with the reference's fields (random weights: the
    engineered-score override provides heads)."""
    import head_detector_b200
    from oracle.make_golden import clustered_anchors

    det = head_detector_b200.HeadDetector(weights=no.synthetic_weights(0))
    ob, osc = clustered_anchors(8400, 4, 10, seed=5)
    det.model.set_override(ob[None].cuda(), osc[None].cuda())
    rng = np.random.default_rng(0)
    image = rng.integers(0, 256, (480, 640, 3), dtype=np.uint8)     # letterboxed: pad (0, 80), scale 1.0
    res = det(image, confidence_threshold=0.5)
    want = nms_oracle.select_nms(ob.numpy(), osc.numpy())
    assert len(res.heads) == len(want) > 0
    h = res.heads[0]
    assert h.vertices_3d.shape == (5023, 3) and h.vertices_3d.dtype == np.float32
    assert h.flame_params.rotation.shape == (1, 6) and len(h.bbox) == 4 and -180 <= h.head_pose.yaw <= 180
    assert "num heads" in repr(res)


def test_pipelined_host_path_matches_synchronous_calls():
    """submit_host/collect_host (two batches in flight, copies overlapped) == run_host, batch by batch."""
    from head_detector_b200 import synth
    from head_detector_b200.engine import Engine

    B, S = 2, 640
    eng = Engine(no.synthetic_weights(0), B, S)
    boxes, scores = synth.engineered_heads(B, eng.A, S, heads=5, per_cluster=8, seed=3)
    eng.set_override(boxes.cuda(), scores.cuda())
    imgs = [synth.synthetic_images(B, S, seed=40 + i).pin_memory() for i in range(4)]
    ref, out = [], eng.alloc_host_outputs(B * 100)
    for im in imgs:
        n = eng.run_host(im, out)
        ref.append((n, out["keep_cnt"].clone(), out["keep_boxes"].clone(), out["params"][:n].clone(), out["verts"][:n].clone()))
    got = []
    eng.submit_host(imgs[0])
    for i in range(1, 4):
        eng.submit_host(imgs[i])
        n = eng.collect_host(out)
        got.append((n, out["keep_cnt"].clone(), out["keep_boxes"].clone(), out["params"][:n].clone(), out["verts"][:n].clone()))
    n = eng.collect_host(out)
    got.append((n, out["keep_cnt"].clone(), out["keep_boxes"].clone(), out["params"][:n].clone(), out["verts"][:n].clone()))
    with pytest.raises(RuntimeError):
        eng.collect_host(out)          # nothing outstanding
    for r, g in zip(ref, got):
        assert r[0] == g[0] > 0
        for a, b in zip(r[1:], g[1:]):
            assert torch.equal(a, b)
    # different images must give different FLAME rows (the staging really follows the inputs)
    assert not torch.equal(ref[0][3], ref[1][3])


def test_highres_1280_many_heads():
    """BASELINE configs[4] shape: 1280x1280 (33600 anchors), ~30 heads/image -> top-k path + FLAME."""
    from head_detector_b200 import synth
    from head_detector_b200.engine import Engine

    B, S = 1, 1280
    eng = Engine(no.synthetic_weights(0), B, S)
    assert eng.A == 33600
    boxes, scores = synth.engineered_heads(B, eng.A, S, heads=30, per_cluster=40, seed=9)
    assert int((scores >= 0.5).sum()) > 1000       # exercises the radix-select top-k
    eng.set_override(boxes.cuda(), scores.cuda())
    eng.forward(synth.synthetic_images(B, S, seed=1).cuda())
    eng.postprocess(0.5, 0.5, 1000)
    torch.cuda.synchronize()
    want = nms_oracle.select_nms(boxes[0].numpy(), scores[0].numpy())
    n = int(eng.head_offsets[-1])
    assert eng.keep_idx.cpu()[0, :n].tolist() == want.tolist() and 20 <= n <= 100
    rows = eng.head_params(n).cpu()
    assert torch.isfinite(rows).all() and torch.isfinite(eng.boxes).all()
    ref = flame_oracle.detector_vertices(rows, flame_oracle.load_flame_constants())
    assert (eng.head_verts(n).cpu() - ref).abs().max() < 2e-4   # coordinates up to 1280: 1 ulp = 1.2e-4


def test_key_buffers_640_batch3_vs_cpu_interpretation():
    """Reference resolution, odd batch: exercises the 40x40 / 20x20 tilings (incl. tiles that pack the
    left-over rows of several images) against the CPU interpretation of the same plan."""
    from head_detector_b200 import synth
    from head_detector_b200.engine import Engine

    B, S = 3, 640
    eng = Engine(no.synthetic_weights(5), B, S)
    img = synth.synthetic_images(B, S, seed=11)
    eng.forward(img.cuda())
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = pe.run_plan(eng.packed, img, emulate_bf16=True)
    for name in ("c3", "c4", "c5", "p4", "p5", "head2.flame_raw", "head3.reg_raw", "head1.reg_raw"):
        got, want = eng.read_buffer(name), ref[eng.plan.buf_names[name]]
        scale = want.abs().max().item() + 1e-6
        err = (got - want).abs()
        assert err.max().item() <= 2 ** -5 * scale + 1e-4, (name, err.max().item(), scale)
        assert err.mean().item() <= 4e-3 * scale, (name, err.mean().item(), scale)


def test_two_detector_handles_run_concurrently():
    """Two handles on two streams (two batches in flight, as bench.py runs them) give exactly the
    results each gives alone."""
    from head_detector_b200 import synth
    from head_detector_b200.engine import Engine

    B, S = 2, 640
    w = no.synthetic_weights(0)
    engs = [Engine(w, B, S) for _ in range(2)]
    boxes, scores = synth.engineered_heads(B, engs[0].A, S, heads=6, per_cluster=10, seed=12)
    imgs = [synth.synthetic_images(B, S, seed=70 + i).cuda() for i in range(2)]
    alone = []
    for e, im in zip(engs, imgs):
        e.set_override(boxes.cuda(), scores.cuda())
        e.input.copy_(im)
        e.run_device()
        torch.cuda.synchronize()
        n = int(e.head_offsets[-1])
        alone.append((n, e.keep_idx.clone(), e.head_params(n).clone(), e.head_verts(n).clone(), e.boxes.clone()))
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for _ in range(3):
        for e, im, st in zip(engs, imgs, streams):
            with torch.cuda.stream(st):
                e.input.copy_(im, non_blocking=True)
                e.run_device()
    torch.cuda.synchronize()
    for e, ref in zip(engs, alone):
        n = int(e.head_offsets[-1])
        assert n == ref[0] > 0
        assert torch.equal(e.keep_idx, ref[1]) and torch.equal(e.head_params(n), ref[2])
        assert torch.equal(e.head_verts(n), ref[3]) and torch.equal(e.boxes, ref[4])


def test_parse_predictions_matches_reference_golden(golden_dir):
    """Host glue of `HeadDetector` (detector.py:61-90) on the reference's own output for the same kept
    heads: integer bboxes equal, vertices within 1e-4/scale, roll/pitch/yaw and rescaled `scale` equal."""
    import os

    from head_detector_b200.detector import HeadDetector
    from head_detector_b200.flame import FLAMELayer

    g = np.load(os.path.join(golden_dir, "parse_ref.npz"))
    flame = FLAMELayer()
    params = torch.from_numpy(g["params"]).cuda()
    n = params.shape[0]
    pad, scale = (0, 80), 0.5
    xf = torch.tensor([[pad[0], pad[1], scale]] * n).cuda()
    det = object.__new__(HeadDetector)
    det._image_size, det._flame = 640, flame
    # the reference signature (detector.py:61): kept boxes, scores, 413-float rows, cache - decodes the vertices itself
    heads = det._parse_predictions(torch.from_numpy(g["boxes"]).cuda(), torch.from_numpy(g["scores"]).cuda(), params, {"padding": pad, "scale": scale})
    # ... and with the vertices the engine already decoded riding along in the cache (what _postprocess does)
    _, rot, verts = flame.decode(params, xform=xf, live=(128, 64))
    again = det._parse_predictions(torch.from_numpy(g["boxes"]).cuda(), torch.from_numpy(g["scores"]).cuda(), params,
                                   {"padding": pad, "scale": scale, "_decoded": (verts, rot)})
    assert all(np.array_equal(a.vertices_3d, b.vertices_3d) and a.bbox == b.bbox for a, b in zip(heads, again))
    assert [[int(v) for v in h.bbox] for h in heads] == g["bbox_xywh"].tolist()
    assert np.abs(np.stack([h.vertices_3d for h in heads]) - g["vertices_3d"]).max() < 1e-4 / scale
    assert np.allclose([[h.head_pose.roll, h.head_pose.pitch, h.head_pose.yaw] for h in heads], g["rpy"], atol=1e-3)
    assert np.allclose([float(h.flame_params.scale) for h in heads], g["out_scale"], rtol=1e-6)
    assert np.allclose([float(h.score) for h in heads], g["scores"])


def test_predict_batch_api():
    """Batched extension of the public call: one PredictionResult per image, per-image letterbox."""
    import head_detector_b200
    from head_detector_b200 import synth

    det = head_detector_b200.HeadDetector(weights=no.synthetic_weights(0), batch_size=3)
    boxes, scores = synth.engineered_heads(3, 8400, 640, heads=4, per_cluster=8, seed=2)
    det.model.set_override(boxes.cuda(), scores.cuda())
    rng = np.random.default_rng(1)
    images = [rng.integers(0, 256, (640, 640, 3), dtype=np.uint8), rng.integers(0, 256, (320, 480, 3), dtype=np.uint8)]
    res = det.predict_batch(images)
    assert len(res) == 2
    for b, r in enumerate(res):
        want = nms_oracle.select_nms(boxes[b].numpy(), scores[b].numpy())
        assert len(r.heads) == len(want) > 0 and r.original_image.shape == images[b].shape
        assert all(h.vertices_3d.shape == (5023, 3) for h in r.heads)
    with pytest.raises(ValueError):
        det.predict_batch([images[0]] * 4)


@pytest.mark.parametrize("swap,xr,cluster,pair", [("1", "1", "1", "0"), ("1", "1", "0", "1"), ("1", "1", "0", "0"), ("1", "0", "1", "0"), ("1", "0", "0", "0"),
                                                  ("0", "0", "0", "0")])
def test_kernel_variants_every_buffer_small(monkeypatch, swap, xr, cluster, pair):
    """The un-tuned heuristic only uses the operand-swapped kernel (and its 3x3 tap-reuse variant) on large
    maps; force it onto every eligible layer of the small case - and switch the tap reuse off - so that each
    variant is checked buffer by buffer against the CPU interpretation of the plan."""
    from head_detector_b200.engine import Engine

    monkeypatch.setenv("VGGHEADS_B200_SWAP", swap)
    monkeypatch.setenv("VGGHEADS_B200_XR", xr)
    monkeypatch.setenv("VGGHEADS_B200_CLUSTER", cluster)   # CTA pairs sharing the weight stream (odd tile counts: filler tiles)
    monkeypatch.setenv("VGGHEADS_B200_PAIR", pair)         # cta_group::2 MMA pairs on the layers with >= 256 output channels
    S, B = 128, 3
    eng = Engine(no.synthetic_weights(4), B, S)
    used = [eng.op_config(i) for i, op in enumerate(eng.plan.ops) if op.kind == 1]
    if swap == "1":
        assert any(c["mt"] < 0 for c in used)
        assert any(c["stages"] >= 100 for c in used) == (xr == "1")   # op_config reports tap reuse as 100*pixel slots + weight slots
        assert any(-20 < c["mt"] <= -11 for c in used) == (cluster == "1")  # ... weight-multicast pairs as mt = -(10 + k-blocks per stage)
        assert any(c["mt"] <= -21 for c in used) == (pair == "1")           # ... and cta_group::2 MMA pairs as mt = -(20 + ...)
    torch.manual_seed(1)
    img = torch.randint(0, 256, (B, S, S, 3), dtype=torch.uint8)
    eng.forward(img.cuda())
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = pe.run_plan(eng.packed, img, emulate_bf16=True)
    bad = []
    for name, i in eng.plan.buf_names.items():
        got, want = eng.read_buffer(name), ref[i]
        scale = want.abs().max().item() + 1e-6
        err = (got - want).abs()
        tol_max, tol_mean = FLIP_TOL[eng.act_dtype]
        if not (err.max().item() <= tol_max * scale + 1e-5 and err.mean().item() <= tol_mean * scale):
            bad.append((name, err.max().item(), err.mean().item(), scale))
    assert not bad, bad


@pytest.mark.parametrize("wgroup,stg2,rings", [("0", "0", "0"), ("1", "0", "1"), ("0", "1", "1")])
def test_ring_variants_every_buffer_small(monkeypatch, wgroup, stg2, rings):
    """The 32-channel K blocks of the tap-reuse kernel (96- and 32-channel layers): grouped weight stages (three row taps
    per stage), the second staging tile with the residual prefetch, deep rings - each switched off in turn (the defaults,
    all on, are what every other test runs), buffer by buffer against the CPU interpretation."""
    from head_detector_b200.engine import Engine

    monkeypatch.setenv("VGGHEADS_B200_SWAP", "1")
    monkeypatch.setenv("VGGHEADS_B200_WGROUP", wgroup)
    monkeypatch.setenv("VGGHEADS_B200_STG2", stg2)
    monkeypatch.setenv("VGGHEADS_B200_DEEP_RINGS", rings)
    S, B = 128, 3
    eng = Engine(no.synthetic_weights(4), B, S)
    used = [eng.op_config(i) for i, op in enumerate(eng.plan.ops) if op.kind == 1]
    assert any(c["mt"] == -3 and c["stages"] >= 100 for c in used) == (wgroup == "1")   # tap reuse with three k-blocks per weight stage
    torch.manual_seed(1)
    img = torch.randint(0, 256, (B, S, S, 3), dtype=torch.uint8)
    eng.forward(img.cuda())
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = pe.run_plan(eng.packed, img, emulate_bf16=True)
    bad = []
    tol_max, tol_mean = FLIP_TOL[eng.act_dtype]
    for name, i in eng.plan.buf_names.items():
        got, want = eng.read_buffer(name), ref[i]
        scale = want.abs().max().item() + 1e-6
        err = (got - want).abs()
        if not (err.max().item() <= tol_max * scale + 1e-5 and err.mean().item() <= tol_mean * scale):
            bad.append((name, err.max().item(), err.mean().item(), scale))
    assert not bad, bad


@pytest.mark.parametrize("xr,cluster", [("1", "1"), ("0", "1"), ("1", "0"), ("0", "0")])
def test_kernel_variants_key_buffers_640(monkeypatch, xr, cluster):
    """Reference resolution with the swapped kernel forced everywhere: 160/80/40/20-pixel maps, tiles that
    overhang the 20- and 40-pixel maps, channel groups, residual tiles - with and without tap reuse."""
    from head_detector_b200 import synth
    from head_detector_b200.engine import Engine

    monkeypatch.setenv("VGGHEADS_B200_SWAP", "1")
    monkeypatch.setenv("VGGHEADS_B200_XR", xr)
    monkeypatch.setenv("VGGHEADS_B200_CLUSTER", cluster)
    B, S = 2, 640
    eng = Engine(no.synthetic_weights(6), B, S)
    img = synth.synthetic_images(B, S, seed=12)
    eng.forward(img.cuda())
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = pe.run_plan(eng.packed, img, emulate_bf16=True)
    for name in ("c2", "c3", "c4", "c5", "p3", "p4", "p5", "head1.flame_raw", "head2.flame_raw", "head3.reg_raw", "head1.reg_raw"):
        got, want = eng.read_buffer(name), ref[eng.plan.buf_names[name]]
        scale = want.abs().max().item() + 1e-6
        err = (got - want).abs()
        assert err.max().item() <= 2 ** -5 * scale + 1e-4, (name, err.max().item(), scale)
        assert err.mean().item() <= 4e-3 * scale, (name, err.mean().item(), scale)


def test_autotuned_configuration_keeps_parity():
    """bench.py runs the per-layer configurations `vgh_detector_autotune` picks (operand-swapped / tap-reuse /
    multi-tile variants, tile shapes, pipeline depths).  Whatever it picks must compute the same network:
    key buffers of an autotuned engine against the CPU interpretation, and the tuned engine against itself
    before tuning (different kernels, same math: bf16-rounding-level agreement)."""
    from head_detector_b200 import synth
    from head_detector_b200.engine import Engine

    B, S = 4, 640
    eng = Engine(no.synthetic_weights(7), B, S)
    img = synth.synthetic_images(B, S, seed=13)
    eng.forward(img.cuda())
    torch.cuda.synchronize()
    names = ("c2", "c3", "c4", "c5", "p3", "p4", "p5", "head1.flame_raw", "head2.flame_raw", "head3.flame_raw", "head1.reg_raw", "head3.reg_raw")
    before = {n: eng.read_buffer(n) for n in names}
    cfg0 = [eng.op_config(i) for i, op in enumerate(eng.plan.ops) if op.kind == 1]
    eng.autotune(2)
    cfg1 = [eng.op_config(i) for i, op in enumerate(eng.plan.ops) if op.kind == 1]
    assert cfg0 != cfg1, "autotune changed nothing - the test would not exercise the tuned kernels"
    eng.forward(img.cuda())
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = pe.run_plan(eng.packed, img, emulate_bf16=True)
    for name in names:
        got, want = eng.read_buffer(name), ref[eng.plan.buf_names[name]]
        scale = want.abs().max().item() + 1e-6
        err = (got - want).abs()
        assert err.max().item() <= 2 ** -5 * scale + 1e-4, (name, err.max().item(), scale)
        assert err.mean().item() <= 4e-3 * scale, (name, err.mean().item(), scale)
        drift = (got - before[name]).abs()
        assert drift.max().item() <= 2 ** -5 * scale + 1e-4 and drift.mean().item() <= 4e-3 * scale, (name, drift.max().item())


def _border_and_cluster_heads(B, A, S, seed, heads=5):
    """Engineered detections (clusters of overlapping boxes) plus isolated high-score anchors in the corners / on
    the edges of every head level, so that survivor patches hang over the image border."""
    from head_detector_b200 import synth

    boxes, scores = synth.engineered_heads(B, A, S, heads=heads, per_cluster=8, seed=seed)
    a_off, extra = 0, []
    for stride in (8, 16, 32):
        W = S // stride
        for (y, x) in ((0, 0), (0, W - 1), (W - 1, 0), (W - 1, W - 1), (0, W // 2), (W // 2, W - 1), (1, 1)):
            extra.append((a_off + y * W + x, (x + 0.5) * stride, (y + 0.5) * stride, stride))
        a_off += W * W
    for k, (a, cx, cy, stride) in enumerate(extra):
        b = k % B
        boxes[b, a] = torch.tensor([cx - 1.5, cy - 1.5, cx + 1.5, cy + 1.5])     # 3 px boxes: no overlap with anything
        scores[b, a] = 0.97 - 1e-3 * k
    return boxes, scores


@pytest.mark.parametrize("S,B,heads", [(128, 3, 5), (640, 2, 5), (1280, 1, 30)])
def test_sparse_heads_match_dense_heads(S, B, heads):
    """FLAME branch on survivor patches (after NMS) vs on the whole maps: same survivors, same 413-float rows,
    same vertices - including survivors in the corners / on the borders of every level (patch masks)."""
    from head_detector_b200 import synth
    from head_detector_b200.engine import Engine

    w = no.synthetic_weights(8)
    dense, sparse = Engine(w, B, S, sparse_heads=False), Engine(w, B, S, sparse_heads=True)
    assert sparse.launch_count == 0 and len(sparse.plan.ops) > len(dense.plan.ops)
    img = synth.synthetic_images(B, S, seed=21).cuda()
    boxes, scores = _border_and_cluster_heads(B, dense.A, S, seed=5, heads=heads)
    out = []
    for eng in (dense, sparse):
        eng.set_override(boxes.cuda(), scores.cuda())
        eng.forward(img)
        eng.postprocess(0.5, 0.5, 1000)
        torch.cuda.synchronize()
        n = int(eng.head_offsets[-1])
        out.append((n, eng.keep_idx.cpu(), eng.head_params(n).cpu(), eng.head_verts(n).cpu()))
    (n0, idx0, p0, v0), (n1, idx1, p1, v1) = out
    assert n0 == n1 >= 21 + min(B * heads, 20) // 2 and torch.equal(idx0, idx1)
    scale = p0.abs().amax(dim=0).clamp_min(1.0)
    assert ((p0 - p1).abs() / scale).max().item() < 1e-5, ((p0 - p1).abs() / scale).max().item()
    assert (v0 - v1).abs().max().item() < 1e-3 * max(1.0, v0.abs().max().item() / 640)
    # the graph path replays the same two-phase pipeline
    sparse.input.copy_(img)
    sparse.run_device(0.5, 0.5, 1000)
    torch.cuda.synchronize()
    assert int(sparse.head_offsets[-1]) == n1 and torch.equal(sparse.head_params(n1).cpu(), p1)
    with pytest.raises(RuntimeError):
        sparse.dense_flame()
    # no survivor at all: every patch-phase launch is empty
    sparse.run_device(0.9999, 0.5, 1000)
    torch.cuda.synchronize()
    assert int(sparse.head_offsets[-1]) == 0 and int(sparse.keep_cnt.sum()) == 0
    sparse.run_device(0.5, 0.5, 1000)
    torch.cuda.synchronize()
    assert int(sparse.head_offsets[-1]) == n1 and torch.equal(sparse.head_params(n1).cpu(), p1)


def test_full_batch_is_batch_composition_invariant():
    """BASELINE batch (32 x 640 x 640), size-independent property: what the network computes for an image does not
    depend on which batch it sits in or where - image k of the 32-batch equals the same image inside a 4-batch
    (other tilings, other work-item packing, same arithmetic per pixel), and a repeated image gives repeated rows."""
    from head_detector_b200 import synth
    from head_detector_b200.engine import Engine

    w = no.synthetic_weights(9)
    big, small = Engine(w, 32, 640), Engine(w, 4, 640)
    imgs = synth.synthetic_images(32, 640, seed=31)
    imgs[9] = imgs[20]                                   # a repeated image inside the big batch
    pick = [5, 17, 20, 31]
    big.forward(imgs.cuda())
    small.forward(imgs[pick].cuda())
    torch.cuda.synchronize()
    for name in ("c3", "c5", "p3", "p4", "p5", "head1.reg_raw", "head2.flame_raw", "head3.flame_raw"):
        a, b = big.read_buffer(name), small.read_buffer(name)
        scale = a.abs().max().item() + 1e-6
        assert torch.equal(a[9], a[20]), name
        err = (a[pick] - b).abs()
        assert err.max().item() <= 2 ** -7 * scale and err.mean().item() <= 1e-4 * scale, (name, err.max().item(), err.mean().item(), scale)
    assert torch.equal(big.boxes[9], big.boxes[20]) and torch.equal(big.scores[9], big.scores[20])
    assert (big.boxes[pick] - small.boxes).abs().max().item() <= 0.05
