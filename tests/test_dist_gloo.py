"""world_size-2 gloo tests of the batch sharding and the prediction gather (the N>1 path, host-side logic):
the dict-level ragged gather and the packed-record transport that bench.py / multi-GPU callers use."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from head_detector_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(worker, world=2, timeout=120):
    """Spawns `world` ranks; rank 0 reports plain python / numpy data through an mp.Queue (no shared torch storages,
    get() with a time-out, exit codes checked)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        out = q.get(timeout=timeout)
    finally:
        for p in procs:
            p.join(timeout)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    return out


def _init(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)


def _local(rank):
    cnt = torch.tensor([2, 0, 1], dtype=torch.int32) if rank == 0 else torch.tensor([0, 4, 0], dtype=torch.int32)
    n = int(cnt.sum())
    g = torch.Generator().manual_seed(rank)
    return {"keep_cnt": cnt, "boxes": torch.rand(3 * 5, 4, generator=g), "scores": torch.rand(3 * 5, generator=g),
            "rot": torch.rand(n, 9, generator=g), "params": torch.rand(n, 413, generator=g),
            "verts": torch.rand(n, 7, 3, generator=g), "head_img": torch.arange(n, dtype=torch.int32) + 100 * rank}


def _worker(rank, world, port, q):
    _init(rank, world, port)
    out = parallel.gather_predictions(_local(rank), dst=0)
    if rank == 0:
        q.put({k: v.numpy().copy() for k, v in out.items()})
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_covers_batch():
    for total, world in ((256, 8), (10, 4), (3, 8)):
        spans = [parallel.shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def test_ragged_gather_two_ranks():
    out = _run(_worker)
    a, b = _local(0), _local(1)
    assert out["keep_cnt"].tolist() == [2, 0, 1, 0, 4, 0] and out["keep_cnt"].dtype == np.int32
    for k in ("boxes", "scores", "rot", "params", "verts", "head_img"):
        want = torch.cat([a[k], b[k]]).numpy()
        assert out[k].dtype == want.dtype and np.array_equal(out[k], want), k


def _worker_empty(rank, world, port, q):
    _init(rank, world, port)
    local = _local(rank)
    local = {k: (torch.zeros_like(v) if k == "keep_cnt" else v if k in parallel.FIXED_KEYS else v[:0]) for k, v in local.items()}
    out = parallel.gather_predictions(local, dst=0, n_heads=0)
    if rank == 0:
        q.put({k: tuple(v.shape) for k, v in out.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_gather_without_any_head():
    shapes = _run(_worker_empty)
    assert shapes["keep_cnt"] == (6,) and shapes["params"] == (0, 413) and shapes["verts"] == (0, 7, 3) and shapes["boxes"] == (30, 4)


# ------------------------------------------------------------------------------------------ packed records
B, K = 3, 5
LAYOUT = parallel.record_layout(B, K)


def _step_data(rank, step):
    """Ragged per-step predictions of one rank (0 .. B*K heads, including the empty and the full step)."""
    g = torch.Generator().manual_seed(1000 * rank + step)
    n = [0, 4, B * K, 1, 7][(step + 2 * rank) % 5]
    cnt = torch.zeros(B, dtype=torch.int32)
    left = n
    for b in range(B):
        cnt[b] = min(K, left)
        left -= int(cnt[b])
    return {"keep_cnt": cnt, "boxes": torch.rand(B, K, 4, generator=g), "scores": torch.rand(B, K, generator=g),
            "params": torch.rand(n, 413, generator=g), "verts": torch.rand(n, 5023, 3, generator=g)}


def test_record_pack_unpack_roundtrip():
    for step in range(5):
        d = _step_data(0, step)
        rec = parallel.pack_record(LAYOUT, d["keep_cnt"], d["boxes"], d["scores"], d["params"], d["verts"], seq=step)
        assert rec.numel() == parallel.record_words(LAYOUT, d["params"].shape[0]) <= LAYOUT["capacity_words"]
        u = parallel.unpack_record(LAYOUT, rec)
        assert u["n_heads"] == d["params"].shape[0] and u["seq"] == step
        for k in ("keep_cnt", "boxes", "scores", "params", "verts"):
            assert torch.equal(u[k], d[k]), k
    full = parallel.record_layout(64, 100)   # the figures quoted in DESIGN.md / the C header
    assert full["fixed_words"] % 4 == 0 and full["capacity_words"] == full["fixed_words"] + 64 * 100 * 413 + 64 * 100 * 15069


def _worker_records(rank, world, port, q):
    _init(rank, world, port)
    steps, lag = 7, 2
    rg = parallel.RecordGather(LAYOUT, lag=lag, device=torch.device("cpu"))
    got = []
    seen = 0

    def harvest():
        nonlocal seen
        # rank 0: a gather completes `lag` submissions late; copy it out before the ring slot is reused
        if rank == 0 and rg.last is not None and rg.last_counts is not None and rg._done > seen:
            seen = rg._done
            got.append([{k: (v.numpy().copy() if torch.is_tensor(v) else v) for k, v in parallel.unpack_record(LAYOUT, rec, n).items()}
                        for rec, n in zip(rg.last, rg.last_counts)])

    for t in range(steps):
        slot, rec = rg.acquire()
        d = _step_data(rank, t)
        parallel.pack_record(LAYOUT, d["keep_cnt"], d["boxes"], d["scores"], d["params"], d["verts"], seq=t, out=rec)
        rg.submit(slot)
        harvest()
    while rg.pending:
        rg._finish(*rg.pending.pop(0))
        harvest()
    if rank == 0:
        q.put(got)
    dist.barrier()
    dist.destroy_process_group()


def test_record_gather_two_ranks_lagged():
    got = _run(_worker_records)
    assert len(got) == 7
    for t, per_rank in enumerate(got):
        for r, u in enumerate(per_rank):
            d = _step_data(r, t)
            assert u["n_heads"] == d["params"].shape[0], (t, r)
            for k in ("keep_cnt", "boxes", "scores", "params", "verts"):
                assert np.array_equal(u[k], d[k].numpy()), (t, r, k)
