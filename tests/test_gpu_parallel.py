"""Multi-GPU path on the device: packed result records (pack kernel vs its torch twin), the host pipeline after an odd
number of device steps (the slot desync that broke the round-1 scaling run), flow-control flags, and - when the box has
two GPUs - both gather transports and bench.run_ours at world size 2."""
import os
import socket
import sys
import traceback

import pytest
import torch

from oracle import net_oracle as no

pytestmark = pytest.mark.gpu

S, B = 128, 2
CONF, IOU, TOPK = 0.5, 0.5, 1000


def _engine(seed=0, sparse=True, weights=None):
    from head_detector_b200 import synth
    from head_detector_b200.engine import Engine

    eng = Engine(weights if weights is not None else no.synthetic_weights(3), B, S, sparse_heads=sparse)
    boxes, scores = synth.engineered_heads(B, eng.A, S, heads=4, per_cluster=6, seed=11 + seed)
    eng.set_override(boxes.cuda(), scores.cuda())
    eng.input.copy_(synth.synthetic_images(B, S, seed=seed).cuda())
    return eng


def _live(eng):
    n = int(eng.head_offsets[-1])
    return {"keep_cnt": eng.keep_cnt.clone(), "boxes": eng.keep_boxes.clone(), "scores": eng.keep_scores.clone(),
            "params": eng.head_params(n).clone(), "verts": eng.head_verts(n).clone()}


def test_record_layout_matches_library():
    from head_detector_b200 import parallel

    eng = _engine()
    assert eng.record_layout() == parallel.record_layout(B, eng.keep_k)


def test_pack_kernel_matches_torch_twin():
    from head_detector_b200 import parallel

    eng = _engine()
    lay = eng.record_layout()
    rec = torch.zeros(lay["capacity_words"], device="cuda")   # (pad words are never written)
    flags = torch.zeros(2, dtype=torch.int64, device="cuda")   # [0] wait flag (already satisfied), [1] done flag
    flags[0] = 3
    eng.arm_push(rec.data_ptr(), flags[0:].data_ptr(), 3, flags[1:].data_ptr(), 41)
    eng.submit_device(CONF, IOU, TOPK)
    torch.cuda.synchronize()
    assert int(flags[1]) == 41 and eng.push_status() == 0
    live = _live(eng)
    n = live["params"].shape[0]
    assert n > 0
    want = parallel.pack_record(lay, live["keep_cnt"], live["boxes"], live["scores"], live["params"], live["verts"], seq=0)
    words = parallel.record_words(lay, n)
    assert torch.equal(rec[:words].view(torch.int32), want.view(torch.int32))
    u = parallel.unpack_record(lay, rec)
    assert u["n_heads"] == n and torch.equal(u["verts"], live["verts"]) and torch.equal(u["keep_cnt"], live["keep_cnt"])
    # a second step without arming leaves the record alone; sequence numbers count pushes
    rec2 = rec.clone()
    eng.submit_device(CONF, IOU, TOPK)
    eng.arm_push(rec.data_ptr())
    eng.submit_device(CONF, IOU, TOPK)
    torch.cuda.synchronize()
    assert int(rec[3:4].view(torch.int32)[0]) == 1
    rec2[3:4].view(torch.int32)[0] = 1
    assert torch.equal(rec[:words].view(torch.int32), rec2[:words].view(torch.int32))


def test_wait_flag_time_out_is_reported_not_hung(monkeypatch):
    monkeypatch.setenv("VGGHEADS_B200_PUSH_TIMEOUT_MS", "50")
    eng = _engine()
    lay = eng.record_layout()
    rec = torch.zeros(lay["capacity_words"], device="cuda")
    flag = torch.zeros(1, dtype=torch.int64, device="cuda")
    eng.arm_push(rec.data_ptr(), flag.data_ptr(), 1)   # nobody will ever raise the flag
    eng.submit_device(CONF, IOU, TOPK)
    torch.cuda.synchronize()
    assert eng.push_status() == 1
    assert int(rec[:1].view(torch.int32)[0]) == int(eng.head_offsets[-1])   # the step still completed


@pytest.mark.parametrize("n_device", [1, 3, 13])
def test_host_pipeline_after_odd_number_of_device_steps(n_device):
    """Round 1: submit_device advanced the host pipeline's slot counter, so an odd number of device steps made the third
    submit_host fail with 'pipeline full'.  The two paths are independent now."""
    from head_detector_b200 import synth

    eng = _engine()
    lay = eng.record_layout()
    rec = torch.zeros(lay["capacity_words"], device="cuda")
    for _ in range(n_device):
        eng.arm_push(rec.data_ptr())
        eng.submit_device(CONF, IOU, TOPK)
    torch.cuda.synchronize()
    want = _live(eng)
    n = want["params"].shape[0]
    host = synth.synthetic_images(B, S, seed=0).pin_memory()
    out = eng.alloc_host_outputs(B * eng.keep_k)
    eng.submit_host(host, CONF, IOU, TOPK)
    eng.submit_host(host, CONF, IOU, TOPK)
    with pytest.raises(RuntimeError, match="pipeline full"):
        eng.submit_host(host, CONF, IOU, TOPK)
    for i in range(7):
        assert eng.collect_host(out) == n
        assert torch.equal(out["verts"][:n], want["verts"].cpu()) and torch.equal(out["keep_cnt"], want["keep_cnt"].cpu())
        eng.submit_host(host, CONF, IOU, TOPK)
    assert eng.collect_host(out) == n and eng.collect_host(out) == n
    with pytest.raises(RuntimeError, match="nothing to collect"):
        eng.collect_host(out)
    # and device steps interleaved with host steps
    eng.submit_host(host, CONF, IOU, TOPK)
    assert eng.collect_host(out) == n
    eng.submit_device(CONF, IOU, TOPK)
    torch.cuda.synchronize()
    eng.submit_host(host, CONF, IOU, TOPK)
    assert eng.collect_host(out) == n


# ------------------------------------------------------------------------------------------ two ranks (NCCL + NVLink)
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _spawn(worker, world, *args, timeout=600):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_guard, args=(worker, r, world, port, q) + args) for r in range(world)]
    for p in procs:
        p.start()
    try:
        out = q.get(timeout=timeout)
    finally:
        for p in procs:
            p.join(timeout)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    return out


def _guard(worker, rank, world, port, q, *args):
    import torch.distributed as dist

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    try:
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        worker(rank, world, q, *args)
        dist.barrier()
        dist.destroy_process_group()
    except BaseException:
        sys.stderr.write(f"rank {rank}:\n{traceback.format_exc()}")
        sys.stderr.flush()
        if rank == 0:
            q.put({"error": traceback.format_exc()})
        raise


def _gather_worker(rank, world, q, transport, steps):
    from head_detector_b200 import parallel

    w = no.synthetic_weights(3)
    engs = [_engine(seed=10 * rank + k, weights=w) for k in range(2)]
    lay = engs[0].record_layout()
    lanes = [torch.cuda.Stream() for _ in engs]
    # what rank 0 must receive from rank r: recomputed locally from the same seeds (same kernels, same GPU type -> bit-equal)
    expect = None
    if rank == 0:
        expect = []
        for r in range(world):
            per_eng = []
            for k in range(2):
                e = _engine(seed=10 * r + k, weights=w)
                e.run_device(CONF, IOU, TOPK)
                torch.cuda.synchronize()
                per_eng.append({kk: v.cpu() for kk, v in _live(e).items()})
            expect.append(per_eng)
    bad = []

    def check(t, recs, counts=None):
        for r, rec in enumerate(recs):
            u = parallel.unpack_record(lay, rec, None if counts is None else counts[r])
            want = expect[r][t % 2]
            if u["n_heads"] != want["params"].shape[0] or not all(torch.equal(u[k].cpu(), want[k]) for k in ("keep_cnt", "boxes", "scores", "params", "verts")):
                bad.append((t, r, u["n_heads"]))

    if transport == "peer":
        pg = parallel.PeerGather(lay, depth=3)
        for t in range(steps):
            e, lane = engs[t % 2], lanes[t % 2]
            with torch.cuda.stream(lane):
                pg.arm(e, t)
                e.submit_device(CONF, IOU, TOPK)
            if rank == 0:
                if t % 4 == 3 or t == steps - 1:   # a consumer that reads the records: wait, copy out on `side`, then hand the slots back
                    pg.consume(t, ack=False)
                    with torch.cuda.stream(pg.side):
                        recs = [pg.record(r, t).clone() for r in range(world)]
                    pg.ack(t)
                    pg.side.synchronize()          # (host sync only in the test)
                    assert int(pg.totals[t % pg.depth]) == sum(expect[r][t % 2]["params"].shape[0] for r in range(world))
                    check(t, recs)
                else:
                    pg.consume(t)
        torch.cuda.synchronize()
        st = sum(e.push_status() for e in engs) + (int(pg.status[0]) if rank == 0 else 0)
        pg.close()
    else:
        rg = parallel.RecordGather(lay, lag=2)
        done = 0
        for t in range(steps):
            e, lane = engs[t % 2], lanes[t % 2]
            slot, rec = rg.acquire()
            with torch.cuda.stream(lane):
                ev = rg.wait_event(slot)
                if ev is not None:
                    lane.wait_event(ev)
                e.arm_push(rec.data_ptr())
                e.submit_device(CONF, IOU, TOPK)
                ready = torch.cuda.Event()
                ready.record(lane)
            rg.submit(slot, ready)
            if rank == 0 and rg._done > done:
                rg.side.synchronize()
                check(done, rg.last, rg.last_counts)
                done = rg._done
        while rg.pending:
            rg._finish(*rg.pending.pop(0))
            if rank == 0:
                rg.side.synchronize()
                check(done, rg.last, rg.last_counts)
                done = rg._done
        rg.flush()
        torch.cuda.synchronize()
        st = sum(e.push_status() for e in engs)
        assert rank != 0 or done == steps
    if rank == 0:
        q.put({"bad": bad, "status": st})
    else:
        assert st == 0


needs2 = pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")


@needs2
@pytest.mark.parametrize("transport,steps", [("peer", 9), ("peer", 2), ("nccl", 9), ("nccl", 1)])
def test_two_rank_gather(transport, steps):
    out = _spawn(_gather_worker, 2, transport, steps)
    assert "error" not in out, out.get("error")
    assert out["bad"] == [] and out["status"] == 0


def _bench_worker(rank, world, q, steps, gather):
    import argparse

    import bench

    lines = []
    for k in steps:
        args = argparse.Namespace(gpus=world, steps=k, warmup=5, impl="ours", per_gpu_batch=2, gather=gather, dense_heads=False,
                                  no_autotune=True, no_cpu_baseline=True, no_extras=True, keep_process_group=True, engines=2)
        line = bench.run_ours(args, rank, world, rank)
        if rank == 0:
            lines.append({k2: line[k2] for k2 in ("value", "n_gpus", "steps", "gather", "heads_per_step_per_gpu")} | {"e2e": line["e2e"]["value"], "parity": line["parity"]})
    if rank == 0:
        q.put(lines)


@needs2
@pytest.mark.parametrize("gather", ["peer", "nccl"])
def test_bench_run_ours_two_ranks(gather):
    """The driver's scaling run (--steps 20 --warmup 5) and the step counts around it, at world size 2."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    lines = _spawn(_bench_worker, 2, (1, 20, 30), gather, timeout=900)
    assert not isinstance(lines, dict), lines
    for k, line in zip((1, 20, 30), lines):
        assert line["n_gpus"] == 2 and line["steps"] == k and line["value"] > 0 and line["e2e"] > 0
        assert line["gather"]["status"] == 0 and line["gather"]["heads_gathered_last_step"] == 2 * line["heads_per_step_per_gpu"]
        assert line["parity"]["nms_ids_bit_exact"] and line["parity"]["vertices_3d_max_abs_err_px"] <= 1e-4
