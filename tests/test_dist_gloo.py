"""world_size-2 gloo test of the batch sharding and the ragged prediction gather (the N>1 path)."""
import os
import socket

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


def _local(rank):
    cnt = torch.tensor([2, 0, 1], dtype=torch.int32) if rank == 0 else torch.tensor([0, 4, 0], dtype=torch.int32)
    n = int(cnt.sum())
    g = torch.Generator().manual_seed(rank)
    return {"keep_cnt": cnt, "boxes": torch.rand(3 * 5, 4, generator=g), "scores": torch.rand(3 * 5, generator=g),
            "rot": torch.rand(n, 9, generator=g), "params": torch.rand(n, 413, generator=g),
            "verts": torch.rand(n, 7, 3, generator=g), "head_img": torch.arange(n, dtype=torch.int32) + 100 * rank}


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out = parallel.gather_predictions(_local(rank), dst=0)
    if rank == 0:
        q.put({k: v.clone() for k, v in out.items()})
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
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    a, b = _local(0), _local(1)
    assert out["keep_cnt"].tolist() == [2, 0, 1, 0, 4, 0] and out["keep_cnt"].dtype == torch.int32
    for k in ("boxes", "scores", "rot", "params", "verts", "head_img"):
        assert out[k].dtype == a[k].dtype and torch.equal(out[k], torch.cat([a[k], b[k]])), k


def _worker_empty(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    local = _local(rank)
    local = {k: (torch.zeros_like(v) if k == "keep_cnt" else v if k in parallel.FIXED_KEYS else v[:0]) for k, v in local.items()}
    out = parallel.gather_predictions(local, dst=0, n_heads=0)
    if rank == 0:
        q.put({k: tuple(v.shape) for k, v in out.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_gather_without_any_head():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_empty, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    shapes = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert shapes["keep_cnt"] == (6,) and shapes["params"] == (0, 413) and shapes["verts"] == (0, 7, 3) and shapes["boxes"] == (30, 4)
