#!/usr/bin/env python
"""bench.py - images/s of the VGGHeads hot path on N B200s (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this build (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

A "step" = one pass of the whole path over one batch of synthetic input on every rank:
uint8 images [B,640,640,3] per GPU -> conv backbone/neck/heads (tcgen05 implicit GEMM) -> box decode
-> engineered scores (~8 heads/image) -> select+NMS -> survivor FLAME rows ->
fused FLAME decode to 5023-vertex meshes [-> N > 1: the step's packed prediction record is stored straight
into rank 0's receive ring over NVLink by the snapshot kernel (or sent with NCCL, --gather nccl)].
Workload: BASELINE configs[2] (batch 64, full path, ~8 heads/image) on every GPU - weak scaling, so N = 8 processes a
global batch of 512 per step (configs[3]'s batch-256 shard of 32/GPU: `--per-gpu-batch 32`).

`value`  : device-timed (CUDA events, inputs resident in HBM, one CUDA-graph replay per step).
`e2e`    : the same step through the host-buffer C-ABI calls (vgh_detector_submit_host/collect_host): pinned host
           images -> H2D -> graph -> D2H of counts/boxes/scores/params/vertices, every step.
`roofline`: conv_igemm kernel launches of one step timed live with CUDA-event pairs (eager pass),
           executed FLOPs / that time (`frac`, `frac_serial`) and / the timed step (`frac_step`).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PER_GPU_BATCH = 64
IMAGE_SIZE = 640
HEADS_PER_IMAGE = 8
CONF, IOU, TOPK = 0.5, 0.5, 1000


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tflops": d.get("bf16_tflops_sustained", d.get("bf16_tflops")), "hbm": d.get("hbm_gbs"), "src": "measured (MEASURED_PEAKS.json, sustained bf16)"}
    return {"tflops": 1400.0, "hbm": 6650.0, "src": "fallback (B200_PROFILING.md)"}


def ncu_conv_traffic(dense_heads=False, batch=PER_GPU_BATCH):
    """DRAM bytes moved by the conv_igemm launches of ONE step (dram__bytes_read+write summed over the
    launches) from the committed ncu launch list of the same workload (profiles/r2b_ncu_launches.csv: batch 64, sparse
    heads); None when no list of this configuration is committed."""
    import csv

    path = os.path.join(ROOT, "profiles", "r2b_ncu_launches.csv")
    if dense_heads or batch != 64 or not os.path.exists(path):
        return None
    tot = 0.0
    with open(path) as f:
        rows = csv.DictReader(l for l in f if l.startswith('"'))
        for r in rows:
            if "conv_igemm" in r["Kernel Name"] and r["Metric Name"] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                v = float(r["Metric Value"].replace(",", ""))
                unit = r["Metric Unit"].lower()
                tot += v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)
    return tot or None


class ClockSampler:
    """nvidia-smi clock / throttle sampling DURING the timed region (profiling recipe's clocks line)."""

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index
        self.windows = []  # [t0, t1] host-clock intervals of the timed regions; only samples inside them count

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def begin(self):
        self.windows.append([time.time(), None])

    def end(self):
        self.windows[-1][1] = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        rows = [r for t, r in self.rows if any(a <= t <= (b or t) for a, b in self.windows)]
        sm = [float(r[0]) for r in rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "sampled": "nvidia-smi every 20 ms, samples inside the two timed regions (device-timed and e2e)"}


# ------------------------------------------------------------------------------------------ CPU arms
def cpu_reference_step(n_images: int, seed: int, state: dict):
    """One bounded sample of the workload on the host cores through the oracle port of the
    reference path (torch-CPU fp32 network in deploy form, utils.nms, FLAME decode)."""
    import torch

    from head_detector_b200 import synth
    from oracle import flame_oracle, net_oracle, nms_oracle

    if "net" not in state:
        from head_detector_b200 import arch

        state["net"] = net_oracle.DeployNet(arch.synthetic_weights(0))
        state["consts"] = flame_oracle.load_flame_constants()
    img = synth.synthetic_images(n_images, IMAGE_SIZE, seed)
    with torch.no_grad():
        _, _, flame = state["net"].forward(img.permute(0, 3, 1, 2).float() / 255.0)
        boxes, scores = synth.engineered_heads(n_images, flame.shape[1], IMAGE_SIZE, HEADS_PER_IMAGE, seed=seed)
        heads = 0
        for b in range(n_images):
            keep = nms_oracle.select_nms(boxes[b].numpy(), scores[b].numpy(), CONF, IOU, TOPK, 100)
            rows = flame[b][torch.from_numpy(keep)]
            flame_oracle.detector_vertices(rows, state["consts"])
            heads += len(keep)
    return heads


def pick_cpu_threads(state):
    """All the host threads the CPU path can USE: torch-CPU convs stop scaling (and collapse when
    oversubscribed) well below the core count of the GPU boxes, so time one image per candidate."""
    import torch

    cores = os.cpu_count() or 1
    best, best_t = None, 1
    for t in (8, 16, 32, 64, cores):
        if t > cores or (best is not None and t == best_t):
            continue
        torch.set_num_threads(t)
        if best is None:
            cpu_reference_step(1, 900, state)  # first-call warm-up
        t0 = time.perf_counter()
        cpu_reference_step(1, 901, state)
        dt = time.perf_counter() - t0
        if best is None or dt < best:
            best, best_t = dt, t
        if dt > 4 * best:
            break
    torch.set_num_threads(best_t)
    return best_t, cores


def run_reference(args, rank):
    import torch

    if rank != 0:
        return
    per_step = 1
    state = {}
    threads, cores = pick_cpu_threads(state)
    for i in range(args.warmup):
        cpu_reference_step(per_step, 1000 + i, state)
    t0 = time.perf_counter()
    for i in range(args.steps):
        cpu_reference_step(per_step, i, state)
    dt = time.perf_counter() - t0
    v = per_step * args.steps / dt
    line = {
        "impl": "reference", "metric": "images/sec (640x640)", "value": v, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": _config(args.gpus, args.per_gpu_batch),
        "cpu_baseline": {"value": v, "unit": "images/s", "cores": threads, "kind": "port", "host_cores": cores,
                         "sample": f"{per_step} image/step x {args.steps} steps, full path (deploy-form torch-CPU net + NMS + FLAME), {threads} threads (fastest of 8/16/32/64/{cores})"},
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def _config(n, batch=PER_GPU_BATCH, dense_heads=False, gather=None):
    return {"heads": ("dense: FLAME branch of the heads computed on the whole feature maps" if dense_heads else
                      "sparse: FLAME branch of the heads (pose stem, towers, output convs) computed after NMS on 8x8 windows around the "
                      "survivors only - identical boxes / 413-float rows / vertices (tests/test_gpu_net.py::test_sparse_heads_match_dense_heads); "
                      "the `dense_heads` key of this line is the same step with the reference's dense graph"),
            "workload": f"BASELINE configs[2] per GPU: batch {batch}/GPU x {n} GPU(s), {IMAGE_SIZE}x{IMAGE_SIZE} uint8 RGB, "
                        f"VGGHeads_L full path (backbone+neck+heads, box decode, select+NMS, FLAME decode to 5023 verts), ~{HEADS_PER_IMAGE} heads/image; "
                        "superset of configs[1]; configs[3] = the same with --per-gpu-batch 32 on 8 GPUs",
            "global_batch": batch * n, "image_size": IMAGE_SIZE, "heads_per_image": HEADS_PER_IMAGE,
            "parallelism": (f"dp{n} (batch-sharded; predictions gathered to rank 0 every step: {gather})" if n > 1 else "single GPU"),
            "weights": "seeded random-init, deploy (re-parameterised) form",
            "precision": "activations and weights stored in the 16-bit format named by `dtype` (fp16 by default: 11 significant bits, the class of "
                         "the TF32 convs the reference runs on a GPU; --act-dtype bf16 for the other), fp32 accumulation; parity.end_to_end "
                         "holds the measured distance to the fp32 oracle for both formats and for the fp32-class parity mode",
            "in_flight": "2 batches per GPU (two detector handles on two streams; --engines)",
            "l2": f"per-step working set ~{0.157 * batch:.0f} GB >> 126 MB L2; 4 rotating input batches"}


def parity_check(eng, boxes, scores, batch):
    """The other half of the BASELINE metric, on the step the engine has just run: kept anchor ids against the
    utils.nms restatement (bit-exact) and vertices against the FLAME restatement applied to the device's own
    413-float rows (bar: 1e-4 px).  oracle/ is the checker here, never the thing measured."""
    try:
        import torch

        from oracle import flame_oracle, nms_oracle

        torch.cuda.synchronize()
        off, cnt, idx = eng.head_offsets.cpu().numpy(), eng.keep_cnt.cpu().numpy(), eng.keep_idx.cpu().numpy()
        n_chk = int(off[-1])
        ids_ok = all(idx[b, :cnt[b]].tolist() == nms_oracle.select_nms(boxes[b].numpy(), scores[b].numpy(), CONF, IOU, TOPK, 100).tolist()
                     for b in range(batch))
        err = 0.0
        if n_chk:
            p_chk, v_chk = eng.head_params(n_chk).cpu(), eng.head_verts(n_chk).cpu()
            err = float((v_chk - flame_oracle.detector_vertices(p_chk, flame_oracle.load_flame_constants())).abs().max())
        return {"vertices_3d_max_abs_err_px": err, "nms_ids_bit_exact": bool(ids_ok), "heads_checked": n_chk,
                "checker": "oracle/ on this step's own data: kept anchor ids vs the utils.nms restatement, vertices vs the FLAME "
                           "restatement applied to the device's 413-float rows (tolerance of the metric: 1e-4 px)"}
    except Exception as ex:  # never lose the measurement line to the checker
        return {"error": repr(ex)}


def parity_end_to_end(weights, batch=2):
    """End-to-end precision of the conv path on `batch` images of the workload, with oracle/ as the checker: the fp32
    oracle network (torch-CPU) against (a) the throughput mode in the measured 16-bit storage format (`fast`; fp16 by
    default) and in the other one (`fast_bf16` / `fast_fp16`) and (b) the parity mode (three-term split bf16: fp32-class
    arithmetic on the same tensor-core kernels).  Per stage max-abs-error relative to
    the stage's max |value|; then, on REAL network outputs (threshold at the 99.5th score percentile, as random weights
    never reach 0.5): are the kept anchor ids those of `utils.nms` on the oracle's boxes / scores, and how far are the
    decoded vertices from the oracle's (bar of the metric: 1e-4 px)."""
    try:
        import torch

        from head_detector_b200 import synth
        from head_detector_b200.engine import Engine
        from oracle import flame_oracle, net_oracle, nms_oracle

        img = synth.synthetic_images(batch, IMAGE_SIZE, seed=4242)
        taps = {}
        with torch.no_grad():
            ob, os_, of = net_oracle.DeployNet(weights).forward(img.permute(0, 3, 1, 2).float() / 255.0, taps)
        thr = float(torch.quantile(os_.flatten(), 0.995))
        consts = flame_oracle.load_flame_constants()
        want_ids = [nms_oracle.select_nms(ob[b].numpy(), os_[b, :, 0].numpy(), thr, IOU, TOPK, 100) for b in range(batch)]
        out = {"batch": batch, "conf_threshold": thr, "reference_heads": [int(len(k)) for k in want_ids],
               "checker": "oracle/net_oracle.DeployNet (torch-CPU fp32) + utils.nms / FLAME restatements on the same images"}
        from head_detector_b200 import arch

        fast_dt = arch.default_act_dtype()                       # the format of the measured step
        other_dt = "bf16" if fast_dt == "fp16" else "fp16"
        out["fast_act_dtype"] = fast_dt
        for mode in ("parity", "fast", "fast_" + other_dt):
            eng = Engine(weights, batch, IMAGE_SIZE, sparse_heads=False, parity=(mode == "parity"),
                         act_dtype=None if mode == "parity" else (fast_dt if mode == "fast" else other_dt))
            boxes, scores = eng.forward(img.cuda())
            eng.postprocess(thr, IOU, TOPK)
            torch.cuda.synchronize()
            stage = {}
            for name in ("c2", "c3", "c4", "c5", "p3", "p4", "p5"):
                want = taps[name].permute(0, 2, 3, 1)
                stage[name] = float((eng.read_buffer(name) - want).abs().max() / (want.abs().max() + 1e-6))
            r = {"stage_rel_err": stage, "boxes_max_abs_err_px": float((boxes.cpu() - ob).abs().max()),
                 "scores_max_abs_err": float((scores.cpu() - os_[..., 0]).abs().max())}
            off, cnt, idx = eng.head_offsets.cpu().numpy(), eng.keep_cnt.cpu().numpy(), eng.keep_idx.cpu().numpy()
            r["nms_ids_equal_reference"] = bool(all(idx[b, :cnt[b]].tolist() == want_ids[b].tolist() for b in range(batch)))
            r["heads"] = [int(c) for c in cnt]
            verr, matched = 0.0, 0
            n = int(off[-1])
            if n:
                verts = eng.head_verts(n).cpu()
                for b in range(batch):
                    ids = idx[b, :cnt[b]].tolist()
                    ref_rows = {a: i for i, a in enumerate(want_ids[b].tolist())}
                    common = [a for a in ids if a in ref_rows]
                    if not common:
                        continue
                    ref_v = flame_oracle.detector_vertices(of[b][torch.tensor(common)], consts)
                    got_v = verts[off[b]:off[b + 1]][[ids.index(a) for a in common]]
                    verr = max(verr, float((got_v - ref_v).abs().max()))
                    matched += len(common)
            r["vertices_3d_max_abs_err_px"], r["heads_compared"] = verr, matched
            out[mode] = r
            del eng
            torch.cuda.empty_cache()
        return out
    except Exception as ex:  # never lose the measurement line to the checker
        import traceback

        return {"error": repr(ex), "trace": traceback.format_exc()[-600:]}


# ------------------------------------------------------------------------------------------ GPU arm
class Pipeline:
    """Two engines per GPU = two batches in flight: the select/NMS/FLAME tail of batch i (few, small kernels) overlaps
    the stem / stage-1 kernels of batch i+1.  Every step is still one full pass over one batch.  N > 1: every step's
    prediction record goes to rank 0 (parallel.PeerGather over NVLink peer memory, or parallel.RecordGather over NCCL)."""

    def __init__(self, engs, dev_imgs, host_imgs, world, rank, gather_mode):
        import torch

        from head_detector_b200 import parallel

        self.torch, self.engs, self.dev_imgs, self.host_imgs, self.world, self.rank = torch, engs, dev_imgs, host_imgs, world, rank
        self.stream = torch.cuda.current_stream()
        self.lanes = [torch.cuda.Stream() for _ in engs]
        self.t = 0                       # this rank's step counter (identical on every rank)
        self.peer = self.rg = None
        self.gather = None
        if world > 1:
            layout = engs[0].record_layout()
            if gather_mode in ("auto", "peer"):
                try:
                    self.peer = parallel.PeerGather(layout, depth=4)
                    self.gather = "peer: pack kernel stores the record into rank 0's ring over NVLink (CUDA IPC), device-side flags, no host sync"
                except parallel.PeerUnavailable as ex:
                    if gather_mode == "peer":
                        raise
                    if rank == 0:
                        print(f"bench.py: peer-memory gather unavailable ({ex}); using NCCL send/recv", file=sys.stderr, flush=True)
            if self.peer is None:
                self.rg = parallel.RecordGather(layout, lag=2)
                self.gather = "nccl: local pack, counts all-gathered and read 2 steps late, exactly-sized send/recv"

    # -- device-resident step (inputs in HBM)
    def device_step(self, i):
        torch = self.torch
        k = i % len(self.engs)
        e, lane = self.engs[k], self.lanes[k]
        with torch.cuda.stream(lane):
            e.input.copy_(self.dev_imgs[i % len(self.dev_imgs)], non_blocking=True)
            if self.world == 1:
                e.run_device(CONF, IOU, TOPK)
            elif self.peer is not None:
                self.peer.arm(e, self.t)
                e.submit_device(CONF, IOU, TOPK)
                if self.rank == 0:
                    self.peer.consume(self.t)
            else:
                slot, rec = self.rg.acquire()
                ev = self.rg.wait_event(slot)
                if ev is not None:
                    lane.wait_event(ev)
                e.arm_push(rec.data_ptr())
                e.submit_device(CONF, IOU, TOPK)
                ready = torch.cuda.Event()
                ready.record(lane)
                self.rg.submit(slot, ready)
        self.t += 1

    def drain(self):
        if self.rg is not None:
            self.rg.flush(self.stream)
        for l in self.lanes:
            self.stream.wait_stream(l)
        if self.peer is not None and self.rank == 0:
            self.stream.wait_stream(self.peer.side)

    # -- host-buffer steps (H2D + D2H inside), up to 3 in flight over the two engines
    def host_steps(self, steps, out):
        inflight = []

        def collect():
            e, slot = inflight.pop(0)
            e.collect_host(out)
            if slot is not None:
                self.rg.submit(slot, None)   # the record was packed before the results were downloaded

        for i in range(steps):
            e = self.engs[i % len(self.engs)]
            if len(inflight) >= 2 * len(self.engs) - 1:
                collect()
            slot = None
            if self.peer is not None:
                self.peer.arm(e, self.t)
            elif self.rg is not None:
                slot, rec = self.rg.acquire()
                ev = self.rg.wait_event(slot)
                if ev is not None:
                    ev.synchronize()   # the send of 4 steps ago; never pending in practice
                e.arm_push(rec.data_ptr())
            e.submit_host(self.host_imgs[i % len(self.host_imgs)], CONF, IOU, TOPK)
            if self.peer is not None and self.rank == 0:
                self.peer.consume(self.t)
            self.t += 1
            inflight.append((e, slot))
        while inflight:
            collect()
        self.drain()

    def status(self):
        st = 0
        for e in self.engs:
            st |= e.push_status()
        if self.peer is not None and self.rank == 0:
            st |= int(self.peer.status[0])
        return st

    def close(self):
        if self.peer is not None:
            self.peer.close()


def run_ours(args, rank, world, local_rank):
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # many concurrent streams: keep their hardware queues apart
    import torch
    import torch.distributed as dist

    from head_detector_b200 import arch, synth
    from head_detector_b200.engine import Engine

    torch.cuda.set_device(local_rank)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    B = args.per_gpu_batch
    weights = arch.synthetic_weights(0)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # started during set-up so that nvidia-smi is already streaming when the timed regions begin
    engs = [Engine(weights, B, IMAGE_SIZE, sparse_heads=not args.dense_heads) for _ in range(args.engines)]
    eng = engs[0]
    n_rot = 4
    host_imgs = [synth.synthetic_images(B, IMAGE_SIZE, seed=100 * rank + i).pin_memory() for i in range(n_rot)]
    dev_imgs = [h.cuda() for h in host_imgs]
    boxes, scores = synth.engineered_heads(B, eng.A, IMAGE_SIZE, HEADS_PER_IMAGE, seed=7 + rank)
    ovr = (boxes.cuda(), scores.cuda())
    for e in engs:
        e.set_override(*ovr)
        if not args.no_autotune:
            e.autotune(5)  # one-off per-layer kernel configuration search (setup, not timed)
    out = eng.alloc_host_outputs(B * 100)
    pipe = Pipeline(engs, dev_imgs, host_imgs, world, rank, args.gather)
    stream = pipe.stream

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t[0])
        return x

    def timed_device(steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for l in pipe.lanes:
            l.wait_event(e0)
        for i in range(steps):
            pipe.device_step(i)
        pipe.drain()
        e1.record(stream)
        sync_all()
        return max_over_ranks(e0.elapsed_time(e1))

    for i in range(max(args.warmup, 3) + 1):
        pipe.device_step(i)
    pipe.drain()
    sync_all()
    sampler.begin()
    ms_dev = timed_device(args.steps)
    sampler.end()
    heads_total = int(eng.head_offsets[-1])

    # end to end through the host-buffer C-ABI: every step uploads its images from pinned host memory and downloads
    # counts / boxes / scores / params / vertices; N > 1: plus the gather of the step's record to rank 0
    pipe.host_steps(4, out)
    sync_all()
    sampler.begin()
    t0 = time.perf_counter()
    pipe.host_steps(args.steps, out)
    sync_all()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    sampler.end()
    clocks = sampler.stop() if rank == 0 else None
    n_heads = int(out["total"][0])
    h2d = B * IMAGE_SIZE * IMAGE_SIZE * 3
    d2h = B * 4 + 4 + B * 100 * 16 + B * 100 * 4 + n_heads * (413 * 4 + 5023 * 12)
    gather_status = pipe.status()
    gathered_heads = None
    if pipe.peer is not None and rank == 0:
        gathered_heads = int(pipe.peer.totals[(pipe.t - 1) % pipe.peer.depth])
    elif pipe.rg is not None and rank == 0 and pipe.rg.last_counts is not None:
        gathered_heads = int(sum(pipe.rg.last_counts))
    if gather_status:
        raise RuntimeError(f"rank {rank}: prediction gather timed out (status bits {gather_status}: 1 = producer waited for a slot, 2 = consumer waited for a record)")

    line = None
    if rank == 0:
        # live roofline of the dominant kernel (conv_igemm): event pairs around every launch of one step (eager pass)
        rows = eng.profile(iters=3, conf=CONF, iou=IOU, top_k=TOPK)
        conv_ms = sum(t for _, t, f in rows if f > 0)
        conv_flops = sum(f for _, t, f in rows if f > 0)
        all_ms = sum(t for _, t, _ in rows)
        pk = _peaks()
        ach = conv_flops / (conv_ms * 1e-3) / 1e12
        step_ms = ms_dev / args.steps
        dense_flops = 2 * arch.total_macs(IMAGE_SIZE) * B
        roof = {"bound": "tensor", "kernel": f"conv_igemm_swap_kernel / conv_igemm_kernel ({sum(1 for _, t, f in rows if f > 0)} dense launches/step)",
                "achieved": ach, "peak": pk["tflops"], "unit": "TFLOP/s",
                "flops_note": "EXECUTED conv FLOPs of the dense launches / their summed durations" + ("" if args.dense_heads else
                              f"; the reference graph's {dense_flops / 1e12:.2f} TFLOP/step include {100 * (1 - conv_flops / dense_flops):.0f} % "
                              "of FLAME-branch work at anchors NMS discards, which this build does not execute (and does not count)"),
                "frac": ach / pk["tflops"], "frac_serial": ach / pk["tflops"],
                "frac_step": conv_flops / (step_ms * 1e-3) / 1e12 / pk["tflops"],
                "frac_note": "frac = frac_serial: executed FLOPs / summed event durations of the conv launches in a serial eager pass; "
                             "frac_step: the same FLOPs / ms_per_step of the timed CUDA-graph region (all kernels of the step, two batches overlapping)",
                "peak_source": pk["src"], "traffic": ncu_conv_traffic(args.dense_heads, B),
                "traffic_note": "DRAM bytes of all conv_igemm launches of one step from the committed ncu launch list (profiles/r2b_ncu_launches.csv, batch 64); layer-minimal activation bytes are 0.415 GB/image (SURVEY 8d): activations that fit L2 are consumed from L2",
                "conv_ms_per_step": conv_ms, "conv_share_of_step": conv_ms / all_ms,
                "algorithmic_flops_per_step": conv_flops}
        parity = parity_check(eng, boxes, scores, B)
        cpu_base = None
        if world == 1 and not args.no_cpu_baseline:
            st = {}
            threads, cores = pick_cpu_threads(st)
            t0 = time.perf_counter()
            n_img = 0
            while n_img < 8 or time.perf_counter() - t0 < 10.0:
                cpu_reference_step(1, 5001 + n_img, st)
                n_img += 1
                if time.perf_counter() - t0 > 30.0:
                    break
            dt = time.perf_counter() - t0
            cpu_base = {"value": n_img / dt, "unit": "images/s", "cores": threads, "kind": "port", "host_cores": cores,
                        "sample": f"{n_img} images of the same workload (deploy-form torch-CPU fp32 network + utils.nms + FLAME decode restatements), {threads} threads (fastest of 8/16/32/64/{cores})"}
        imgs = B * world * args.steps
        line = {
            "metric": "images/sec (640x640)", "value": imgs / (ms_dev * 1e-3), "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": eng.act_dtype, "data": "synthetic", "config": _config(world, B, args.dense_heads, pipe.gather),
            "clocks": clocks,
            "e2e": {"value": imgs / e2e_s, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "mode": "vgh_detector_submit_host/collect_host over 2 detector handles (3 batches in flight)" + (" + gather of every step's record to rank 0" if world > 1 else "")},
            "gpu_launches": (eng.launch_count + (0 if world == 1 else 1)) * args.steps,
            "roofline": roof, "cpu_baseline": cpu_base, "parity": parity,
            "heads_per_step_per_gpu": heads_total,
            "conv_gflop_per_image": 2 * arch.total_macs(IMAGE_SIZE) / 1e9,
        }
        if world > 1:
            line["gather"] = {"transport": pipe.gather, "heads_gathered_last_step": gathered_heads, "status": gather_status}
    pipe.close()
    del pipe
    if rank == 0 and world == 1 and not args.dense_heads and not args.no_extras:
        for e in engs:
            e.__del__()
        engs.clear()
        torch.cuda.empty_cache()
        try:
            line["dense_heads"] = extra_dense(args, weights, B)
        except Exception as ex:  # never lose the headline to an extra
            line["dense_heads"] = {"error": repr(ex)}
        torch.cuda.empty_cache()
        line["parity"]["end_to_end"] = parity_end_to_end(weights)
        torch.cuda.empty_cache()
        try:
            line["config4_shard"] = extra_config4(args, weights)
        except Exception as ex:
            line["config4_shard"] = {"error": repr(ex)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        if not args.keep_process_group:
            dist.destroy_process_group()
    return line


def extra_dense(args, weights, B):
    """The same step with the reference's dense graph (FLAME branch on the whole maps): the apples-to-apples figure
    against the reference's own network, device-timed the same way (fewer steps)."""
    import torch

    from head_detector_b200 import synth
    from head_detector_b200.engine import Engine

    engs = [Engine(weights, B, IMAGE_SIZE, sparse_heads=False) for _ in range(2)]
    boxes, scores = synth.engineered_heads(B, engs[0].A, IMAGE_SIZE, HEADS_PER_IMAGE, seed=7)
    ovr = (boxes.cuda(), scores.cuda())
    imgs = [synth.synthetic_images(B, IMAGE_SIZE, seed=i).cuda() for i in range(4)]
    for e in engs:
        e.set_override(*ovr)
        if not args.no_autotune:
            e.autotune(3)
    pipe = Pipeline(engs, imgs, None, 1, 0, "auto")
    steps = max(6, min(args.steps, 12))
    for i in range(4):
        pipe.device_step(i)
    pipe.drain()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(pipe.stream)
    for l in pipe.lanes:
        l.wait_event(e0)
    for i in range(steps):
        pipe.device_step(i)
    pipe.drain()
    e1.record(pipe.stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return {"value": B * steps / (ms * 1e-3), "unit": "images/s", "ms_per_step": ms / steps, "steps": steps,
            "note": "--dense-heads: full reference graph (166.68 GFLOP/image), same inputs, device-timed"}


def extra_config4(args, weights):
    """BASELINE configs[4], per-GPU shard: batch 32 at 1280x1280 (33 600 anchors), ~30 heads/image - the top-k > 1000
    path of select/NMS and ~1000 FLAME decodes per step.  Device-timed graph replays, one detector handle."""
    import torch

    from head_detector_b200 import synth
    from head_detector_b200.engine import Engine

    B, S, heads = 32, 1280, 30
    eng = Engine(weights, B, S, sparse_heads=True)
    eng.input.copy_(synth.synthetic_images(B, S, 3).cuda())
    boxes, scores = synth.engineered_heads(B, eng.A, S, heads, per_cluster=40, seed=7)
    eng.set_override(boxes.cuda(), scores.cuda())
    if not args.no_autotune:
        eng.autotune(3)
    for _ in range(3):
        eng.run_device(CONF, IOU, TOPK)
    torch.cuda.synchronize()
    steps = 6
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        eng.run_device(CONF, IOU, TOPK)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"value": B / ms * 1e3, "unit": "images/s", "ms_per_step": ms, "steps": steps, "batch": B, "image_size": S, "anchors": eng.A,
            "heads_per_step": int(eng.head_offsets[-1]),
            "note": "BASELINE configs[4] per-GPU shard (batch 128 over 4 GPUs), full path, device-timed, one handle"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--per-gpu-batch", type=int, default=PER_GPU_BATCH, help="images per GPU per step (64 = BASELINE configs[2]; 32 = the configs[3] shard)")
    ap.add_argument("--gather", default="auto", choices=["auto", "peer", "nccl"],
                    help="N > 1: transport of the per-step prediction gather to rank 0 (auto = NVLink peer memory, NCCL if IPC is unavailable)")
    ap.add_argument("--dense-heads", action="store_true",
                    help="run the FLAME branch of the heads on the whole feature maps (as the reference graph does) instead of on the "
                         "8x8 windows around the NMS survivors; same predictions, ~20 %% more work")
    ap.add_argument("--act-dtype", default=None, choices=["fp16", "bf16"],
                    help="16-bit storage format of activations and weights (default: $VGGHEADS_B200_ACT or fp16; fp32 accumulation either way)")
    ap.add_argument("--engines", type=int, default=2, help="detector handles per GPU = batches in flight on the device-timed path")
    ap.add_argument("--no-autotune", action="store_true", help="skip the per-layer configuration search (tests)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU leg of the N=1 line (tests)")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra dense-heads measurement of the N=1 line")
    ap.add_argument("--keep-process-group", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.act_dtype:
        os.environ["VGGHEADS_B200_ACT"] = args.act_dtype   # every Engine of this process (extras included) follows
    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    if world == 1 and args.gpus > 1:
        raise SystemExit("launch multi-GPU runs with torch.distributed.run (one process per GPU)")
    try:
        run_ours(args, rank, world, local_rank)
    except BaseException:
        import traceback

        sys.stderr.write(f"rank {rank}: bench.py failed\n" + "".join(f"rank {rank}: {l}" for l in traceback.format_exc().splitlines(True)))
        sys.stderr.flush()
        raise


if __name__ == "__main__":
    main()
