"""Generate tests/golden/*.npz by running the UNMODIFIED reference python.  TEST INFRASTRUCTURE ONLY.

Runs only in the build container (needs /root/reference).  The reference package is imported
from where it lies with three stand-ins on sys.path (oracle/shims): `chumpy` (unpickle stub),
`smplx` (its `lbs` is routed to oracle.flame_oracle.lbs_torch - smplx==0.1.26 is not
installed here) and `Sim3DR_Cython` (import stub, rasteriser is off the hot path).

    python oracle/make_golden.py

Fixtures written (all small, committed):
  flame_1json.npz        the reference's own known-answer vector yolo_head_training/tests/1.json
  flame_ref_heads.npz    reference `reproject_spatial_vertices` (flame.py:179-208) on seeded heads
  nms_ref_cases.npz      reference `utils.nms` (utils.py:159-194) on seeded anchors, with the
                         surviving ORIGINAL anchor ids recovered through an index column
  parse_ref.npz          reference `HeadDetector._parse_predictions` (detector.py:61-90)
  letterbox_ref.npz      reference `HeadDetector._transform_image` (detector.py:40-52), one shape
  letterbox_ref_cases.npz  the same on seven more shapes (up/down-scaling, tall, wide, identity);
                         `python oracle/make_golden.py --only-letterbox` rewrites just this file
  heads_ref.npz          the reference's own `YoloHeadsNDFLHeads` / `YoloHeadsDFLHead` (yolo_head_ndfl_heads.py:117-175,
                         yolo_head_dfl_head.py:141-186; super_gradients stand-ins in oracle/shims) on seeded feature
                         maps and seeded weights: raw per-level outputs + decoded boxes / scores / flame[B,A,413]
  topk_ref.npz           reference `VGGHeadDecodingModule` (yolo_heads.py:44-86) followed by reference `utils.nms`
  detector_ref.npz       the UNMODIFIED `HeadDetector.__call__` (detector.py:97-102) on a synthetic `vgg_heads_l.trcd`
                         (oracle/sg_net.py traced; hf_hub_download patched to the local file): boxes, scores, vertices,
                         plus the stage timing of BASELINE configs[0] -> profiles/r2_reference_cpu_config0.json
                         (`python oracle/make_golden.py --only-net` rewrites these three)
  pncc_ref.npz           reference `PredictionResult.get_pncc()` / `get_aligned_heads()` / `draw(...)` on four heads
                         (PNCCProcessor + the Sim3DR C++ rasteriser compiled in place: oracle/Makefile -> oracle/_ref/)
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.path[:0] = [os.path.join(HERE, "shims"), REF, ROOT]
OUT = os.path.join(ROOT, "tests", "golden")


def network_like_heads(n: int, seed: int, size: float = 640.0) -> torch.Tensor:
    """413-float rows shaped like what the network emits."""
    g = torch.Generator().manual_seed(seed)
    p = torch.zeros(n, 413)
    p[:, 0:128] = 3 * torch.tanh(torch.randn(n, 128, generator=g))
    p[:, 300:364] = 3 * torch.tanh(torch.randn(n, 64, generator=g))
    p[:, 400:403] = 0.1 * torch.randn(n, 3, generator=g)
    p[:, 403:409] = torch.randn(n, 6, generator=g)
    p[:, 409:411] = torch.rand(n, 2, generator=g) * size
    p[:, 411] = 10 * torch.randn(n, generator=g)
    p[:, 412] = 200 + 400 * torch.rand(n, generator=g)
    return p


def clustered_anchors(n_anchor: int, n_clusters: int, per_cluster: int, seed: int, size: float = 640.0, bg_hi: float = 0.3):
    """Seeded boxes/scores with engineered clusters (tie-free scores)."""
    g = torch.Generator().manual_seed(seed)
    boxes = torch.rand(n_anchor, 4, generator=g) * size
    boxes = torch.stack(
        [torch.minimum(boxes[:, 0], boxes[:, 2]), torch.minimum(boxes[:, 1], boxes[:, 3]),
         torch.maximum(boxes[:, 0], boxes[:, 2]) + 1, torch.maximum(boxes[:, 1], boxes[:, 3]) + 1], dim=1)
    scores = torch.rand(n_anchor, generator=g) * bg_hi
    perm = torch.randperm(n_anchor, generator=g)
    k = 0
    for c in range(n_clusters):
        cx, cy = (torch.rand(2, generator=g) * (size - 200) + 100).tolist()
        half = float(torch.rand(1, generator=g) * 50 + 40)
        for _ in range(per_cluster):
            i = int(perm[k]); k += 1
            j = (torch.rand(4, generator=g) - 0.5) * 12
            boxes[i] = torch.tensor([cx - half, cy - half, cx + half, cy + half]) + j
            scores[i] = 0.55 + 0.4 * float(torch.rand(1, generator=g))
    scores = scores + torch.arange(n_anchor) * 1e-7  # no exact ties
    return boxes.float(), scores.float()


LETTERBOX_SHAPES = ((720, 1280), (1280, 720), (640, 640), (200, 150), (1080, 1920), (333, 1000), (641, 640))


def letterbox_cases():
    """Unmodified reference `_transform_image` (detector.py:40-52) on seeded images of several shapes:
    sha1 of the letterboxed uint8 frame, a strided probe of it, padding and scale."""
    import hashlib

    from head_detector.detector import HeadDetector

    det = object.__new__(HeadDetector)
    det._image_size, det._device = 640, torch.device("cpu")
    out = {"shapes": np.array(LETTERBOX_SHAPES), "seed0": np.array(100)}
    for i, (h, w) in enumerate(LETTERBOX_SHAPES):
        src = np.random.default_rng(100 + i).integers(0, 256, (h, w, 3), dtype=np.uint8)
        t, pad, scale = det._transform_image(src)
        u8 = (t[0].permute(1, 2, 0) * 255.0).round().to(torch.uint8).numpy()
        out[f"sha1_{i}"] = np.frombuffer(hashlib.sha1(u8.tobytes()).digest(), dtype=np.uint8)
        out[f"probe_{i}"] = u8[::53, ::47].copy()
        out[f"pad_{i}"] = np.array(pad)
        out[f"scale_{i}"] = np.array(scale)
        print("letterbox", (h, w), "pad", pad, "scale", scale)
    np.savez_compressed(os.path.join(OUT, "letterbox_ref_cases.npz"), **out)


HEADS_SEED, HEADS_S, HEADS_B = 5, 128, 2
DETECTOR_SEED, DETECTOR_CLS_BIAS = 1, 1.5


def heads_feats(seed=HEADS_SEED, S=HEADS_S, B=HEADS_B):
    """Seeded neck outputs p3/p4/p5, bf16-representable (the CUDA path stores activations in bf16)."""
    g = torch.Generator().manual_seed(1000 + seed)
    return [(torch.randn(B, c, S // st, S // st, generator=g) * 1.5).clamp_min(0).to(torch.bfloat16).float() for c, st in ((96, 8), (192, 16), (384, 32))]


def heads_cases():
    """The reference's head + decode classes on seeded weights (oracle/sg_net.py names == the reference's names)."""
    from oracle import ref_heads, sg_net

    ns = ref_heads.load()
    ref = ref_heads.build_heads(ns)
    sd = {k[len("heads."):]: v for k, v in sg_net.build(HEADS_SEED).state_dict().items() if k.startswith("heads.")}
    ref.load_state_dict(sd, strict=True)
    feats = heads_feats()
    out = {"seed": np.array(HEADS_SEED), "S": np.array(HEADS_S)}
    with torch.no_grad():
        dec, _ = ref(feats)
        for l, f in enumerate(feats, start=1):
            h = getattr(ref, f"head{l}")
            reg, cls, flame = h(f)                                   # yolo_head_dfl_head.py:141-186
            pose = h.pose_stem(h.stem(f))
            out[f"reg{l}"], out[f"cls{l}"], out[f"flame{l}"] = reg.numpy(), cls.numpy(), flame.numpy()
            for tw, attr in (("shape", "flame_shape_pred"), ("expr", "flame_expression_pred"), ("rot", "flame_rotation_pred"),
                             ("jaw", "flame_jaw_pred"), ("transl", "flame_translation_pred"), ("scale", "flame_scale_pred")):
                out[f"{tw}{l}"] = getattr(h, attr)(pose).numpy()     # pre-activation tower outputs
    out["boxes"], out["scores"], out["flame"] = dec.boxes_xyxy.numpy(), dec.scores.numpy(), dec.flame_params.numpy()
    np.savez_compressed(os.path.join(OUT, "heads_ref.npz"), **out)
    print("heads_ref: anchors", out["boxes"].shape, "scale range", out["flame"][..., 412].min(), out["flame"][..., 412].max())

    # top-k decoding module + utils.nms, per image (utils.nms handles the first image only)
    from head_detector.utils import nms as ref_nms

    dm = ns.VGGHeadDecodingModule(1000)
    B, A = 2, 2100
    boxes = torch.stack([clustered_anchors(A, 20, 14, 40 + b, size=320.0, bg_hi=0.8)[0] for b in range(B)])
    scores = torch.stack([clustered_anchors(A, 20, 14, 40 + b, size=320.0, bg_hi=0.8)[1] for b in range(B)])
    tag = torch.zeros(B, A, 413)
    tag[..., 0] = torch.arange(A)
    preds = type("P", (), {"boxes_xyxy": boxes, "scores": scores[..., None], "flame_params": tag})()
    tb, ts, tf = dm((preds, None))
    cases = {"boxes": boxes.numpy(), "scores": scores.numpy(), "topk_ids": tf[..., 0].long().numpy()}
    for b in range(B):
        kb, ks, kf = ref_nms(tb[b:b + 1], ts[b:b + 1], tf[b:b + 1], confidence_threshold=0.5)
        cases[f"keep{b}"] = kf[:, 0].long().numpy()
        print("topk_ref image", b, "candidates", int((scores[b] >= 0.5).sum()), "kept", len(ks))
    np.savez_compressed(os.path.join(OUT, "topk_ref.npz"), **cases)


def detector_image(seed=DETECTOR_SEED):
    """Non-square seeded frame with smooth structure (letterboxed by the detector: 480x640 -> pad (0, 80))."""
    rng = np.random.default_rng(seed)
    low = rng.integers(0, 256, (15, 20, 3), dtype=np.uint8)
    import cv2
    return cv2.resize(low, (640, 480), interpolation=cv2.INTER_CUBIC)


def detector_net(seed=DETECTOR_SEED):
    """Synthetic detector whose class logits are shifted so that a few hundred anchors pass the 0.5 threshold."""
    from oracle import sg_net

    net = sg_net.build(seed)
    with torch.no_grad():
        for l in (1, 2, 3):
            getattr(net.heads, f"head{l}").cls_pred.bias.fill_(DETECTOR_CLS_BIAS)
    return net


def detector_case():
    """BASELINE configs[0]: the unmodified reference HeadDetector on torch-CPU with a synthetic vgg_heads_l.trcd."""
    import tempfile
    import time

    import head_detector.detector as refdet
    from oracle import sg_net

    tmp = tempfile.mkdtemp()
    path = sg_net.trace_to(os.path.join(tmp, "vgg_heads_l.trcd"), detector_net(), 640)
    refdet.hf_hub_download = lambda repo, name: path          # the one patch: no network
    det = refdet.HeadDetector()                                # reference constructor, torch.jit.load of the blob
    img = detector_image()
    res = det(img, confidence_threshold=0.5)                   # reference __call__
    n = len(res.heads)
    print("detector_ref: heads", n, "first bbox", tuple(res.heads[0].bbox) if n else None)
    keep_v = min(n, 12)
    np.savez_compressed(
        os.path.join(OUT, "detector_ref.npz"), seed=np.array(DETECTOR_SEED), cls_bias=np.array(DETECTOR_CLS_BIAS),
        bbox_xywh=np.array([[int(v) for v in h.bbox] for h in res.heads]).reshape(n, 4), scores=np.array([float(h.score) for h in res.heads]),
        vertices_3d=np.stack([h.vertices_3d for h in res.heads[:keep_v]]) if n else np.zeros((0, 5023, 3), np.float32),
        vertex_mean=np.array([h.vertices_3d.mean(0) for h in res.heads]).reshape(n, 3),
        rpy=np.array([[h.head_pose.roll, h.head_pose.pitch, h.head_pose.yaw] for h in res.heads]).reshape(n, 3),
        params=np.stack([torch.cat([h.flame_params.shape, h.flame_params.expression, h.flame_params.jaw, h.flame_params.rotation,
                                    h.flame_params.translation, h.flame_params.scale], dim=1)[0].numpy() for h in res.heads]) if n else np.zeros((0, 413), np.float32))
    # stage split of config 0: model / nms / flame+parse / result object, torch-CPU, this container's cores
    x, cache = det._preprocess(img)
    reps, t = 3, {}
    with torch.no_grad():
        det._process(x)
        t0 = time.perf_counter()
        for _ in range(reps):
            out = det._process(x)
        t["model_ms"] = 1e3 * (time.perf_counter() - t0) / reps
        t0 = time.perf_counter()
        for _ in range(reps):
            kept = refdet.nms(*out, confidence_threshold=0.5)
        t["nms_ms"] = 1e3 * (time.perf_counter() - t0) / reps
        t0 = time.perf_counter()
        for _ in range(reps):
            det._parse_predictions(*kept, cache)
        t["flame_parse_ms"] = 1e3 * (time.perf_counter() - t0) / reps
        t0 = time.perf_counter()
        for _ in range(reps):
            det(img)
        t["call_ms"] = 1e3 * (time.perf_counter() - t0) / reps
    t.update(heads=n, threads=torch.get_num_threads(), cores=os.cpu_count(),
             what="unmodified reference HeadDetector.__call__ (detector.py:97-102), torch-CPU fp32, synthetic vgg_heads_l.trcd traced from oracle/sg_net.py, 480x640 frame")
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    json.dump(t, open(os.path.join(ROOT, "profiles", "r2_reference_cpu_config0.json"), "w"), indent=1)
    print("config 0 timing", t)


def pncc_case():
    """The reference's `PredictionResult.get_pncc()` / `get_aligned_heads()` / `draw()` on reference-decoded heads (three
    of them overlapping): PNCCProcessor + Sim3DR C++ (compiled in place, oracle/Makefile)."""
    import subprocess

    subprocess.run(["make", "-C", HERE, "_ref/libsim3dr_ref.so"], check=True, capture_output=True)
    from head_detector.detection_result import PredictionResult
    from head_detector.head_info import FlameParams, HeadMetadata
    from head_detector.utils import calculate_rpy, refined_head_bbox

    g = np.load(os.path.join(OUT, "flame_ref_heads.npz"))
    verts = [g["projected"][0].copy(), g["projected"][1].copy(), g["projected"][2].copy(), g["projected"][0].copy() + np.float32([18.5, -11.25, 40.0])]
    verts = [v + np.float32([150.0, -60.0, 0.0]) for v in verts]            # inside a 640x480 frame, the last one overlapping the first
    params = torch.from_numpy(g["params"][[0, 1, 2, 0]])
    heads = []
    for v, p in zip(verts, params):
        fp = FlameParams.from_3dmm(p[None])
        bb = refined_head_bbox(v)
        heads.append(HeadMetadata(bbox=bb, score=0.9, flame_params=fp, vertices_3d=v.copy(), head_pose=calculate_rpy(fp)))
    rng = np.random.default_rng(3)
    frame = rng.integers(0, 256, (480, 640, 3), dtype=np.uint8)
    res = PredictionResult(frame, heads)
    drawn = {m: res.draw(m) for m in ("full", "bbox", "landmarks", "points", "pose")}
    crops = res.get_aligned_heads()
    pncc = res.get_pncc()                                                    # flips z of every head in place (reference quirk)
    print("pncc_ref: painted px", int((pncc.sum(2) != 0).sum()), "crops", [c.shape for c in crops])
    np.savez_compressed(os.path.join(OUT, "pncc_ref.npz"), vertices=np.stack(verts), params=params.numpy(), frame_seed=np.array(3), pncc=pncc,
                        bbox=np.array([[h.bbox.x, h.bbox.y, h.bbox.w, h.bbox.h] for h in heads]),
                        **{f"draw_{k}_{part}": arr for k, v in drawn.items()      # only the pixels a draw method touched
                           for part, arr in zip(("yx", "rgb"), (lambda m: (np.argwhere(m).astype(np.int16), v[m]))((v != frame).any(2)))},
                        **{f"crop{i}": c for i, c in enumerate(crops)},
                        z_after=np.stack([h.vertices_3d[:, 2] for h in heads]))


def main():
    if "--only-pncc" in sys.argv:
        pncc_case()
        return
    if "--only-letterbox" in sys.argv:
        letterbox_cases()
        return
    if "--only-net" in sys.argv:
        heads_cases()
        detector_case()
        return
    os.makedirs(OUT, exist_ok=True)
    from head_detector.flame import FLAMELayer, reproject_spatial_vertices  # the reference, unmodified
    from head_detector.head_info import FlameParams
    from head_detector.utils import nms as ref_nms, calculate_rpy
    from head_detector.detector import HeadDetector

    # --- 1. the reference's own fixture -------------------------------------------------------
    rec = json.load(open(os.path.join(REF, "yolo_head_training", "tests", "1.json")))[0]
    np.savez_compressed(
        os.path.join(OUT, "flame_1json.npz"),
        params=np.asarray(rec["3dmm_params"], dtype=np.float64),
        vertices_3d=np.asarray(rec["3d_vertices"], dtype=np.float64),
        projected_vertices=np.asarray(rec["projected_vertices"], dtype=np.float64),
    )
    flame = FLAMELayer()
    p1 = torch.tensor(rec["3dmm_params"], dtype=torch.float32)[None]
    v1 = flame.forward(FlameParams.from_3dmm(p1), zero_rot=False)[0].numpy()
    print("reference FLAMELayer vs 1.json: max abs", np.abs(v1 - np.asarray(rec["3d_vertices"])).max())

    # --- 2. reference reproject on seeded heads -----------------------------------------------
    heads = torch.cat([network_like_heads(5, seed=11), p1 * 1.0])  # 5 network-like + the dense 400-coef fixture row
    heads[5, 409:412] = torch.tensor([311.5, 207.25, 3.0]); heads[5, 412] = 350.0
    verts, rot, proj = reproject_spatial_vertices(flame, heads, to_2d=False)
    np.savez_compressed(os.path.join(OUT, "flame_ref_heads.npz"), params=heads.numpy(), vertices=verts.numpy(),
                        rotation=rot.numpy(), projected=proj.numpy())

    # --- 3. reference nms, anchor ids recovered through column 0 of the flame tensor -----------
    cases = {}
    for name, (na, ncl, per, seed, bg) in {
        "few": (8400, 8, 12, 3, 0.3),          # ~96 candidates
        "many": (8400, 40, 12, 4, 0.9),        # >1000 candidates -> top-k path
        "none": (8400, 0, 0, 5, 0.3),          # nothing above threshold
        "hires": (33600, 30, 14, 6, 0.6),      # S=1280 anchor count
    }.items():
        boxes, scores = clustered_anchors(na, ncl, per, seed, size=640.0 if na == 8400 else 1280.0, bg_hi=bg)
        tag = torch.zeros(na, 413); tag[:, 0] = torch.arange(na)
        b, s, f = ref_nms(boxes[None], scores[None, :, None], tag[None], confidence_threshold=0.5)
        cases[f"{name}_boxes"] = boxes.numpy(); cases[f"{name}_scores"] = scores.numpy()
        cases[f"{name}_keep"] = f[:, 0].long().numpy(); cases[f"{name}_keep_scores"] = s.numpy()
        print(name, "candidates", int((scores >= 0.5).sum()), "kept", len(s))
    np.savez_compressed(os.path.join(OUT, "nms_ref_cases.npz"), **cases)

    # --- 4. reference _parse_predictions (letterboxed: pad (0,80), scale 0.5) -------------------
    det = object.__new__(HeadDetector)
    det._image_size, det._device, det._flame = 640, torch.device("cpu"), flame
    hp = network_like_heads(3, seed=21)
    hb = torch.tensor([[10.4, 90.6, 200.5, 300.5], [-5.0, 100.0, 650.0, 500.49], [300.5, 301.5, 420.5, 480.5]])
    hs = torch.tensor([0.9, 0.8, 0.7])
    res = det._parse_predictions(hb.clone(), hs.clone(), hp.clone(), {"padding": (0, 80), "scale": 0.5})
    np.savez_compressed(
        os.path.join(OUT, "parse_ref.npz"), params=hp.numpy(), boxes=hb.numpy(), scores=hs.numpy(),
        bbox_xywh=np.array([[int(v) for v in h.bbox] for h in res]),
        vertices_3d=np.stack([h.vertices_3d for h in res]),
        rpy=np.array([[h.head_pose.roll, h.head_pose.pitch, h.head_pose.yaw] for h in res], dtype=np.float64),
        out_scale=np.array([float(h.flame_params.scale) for h in res]),
    )
    # --- 5. reference letterbox (detector.py:40-52) on a non-square image: checksum + geometry ------
    import hashlib
    rng = np.random.default_rng(17)
    src = rng.integers(0, 256, (375, 500, 3), dtype=np.uint8)
    t, pad, scale = det._transform_image(src)          # float [1,3,640,640] in [0,1]
    u8 = (t[0].permute(1, 2, 0) * 255.0).round().to(torch.uint8).numpy()
    np.savez_compressed(os.path.join(OUT, "letterbox_ref.npz"), sha1=np.frombuffer(hashlib.sha1(u8.tobytes()).digest(), dtype=np.uint8),
                        pad=np.array(pad), scale=np.array(scale), probe=u8[::37, ::41].copy(), seed=np.array(17), shape=np.array(src.shape))
    letterbox_cases()
    heads_cases()
    detector_case()
    pncc_case()
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
