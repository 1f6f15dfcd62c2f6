"""Batch-sharded multi-GPU inference: one process per GPU, images sharded by rank, no collective on
the data path except the final gather of predictions to rank 0 (BASELINE.json north_star; SURVEY.md
8e).  torch.distributed (NCCL on GPUs, gloo in the CPU tests) is plumbing only."""
from typing import Dict, Optional

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int):
    """Contiguous slice of the global batch owned by `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


FIXED_KEYS = ("keep_cnt", "boxes", "scores")  # per-image, fixed-capacity tensors: gathered as they are


def gather_predictions(local: Dict[str, torch.Tensor], dst: int = 0, group=None, n_heads: Optional[int] = None) -> Optional[Dict[str, torch.Tensor]]:
    """Ragged gather of one step's predictions to rank `dst`.

    `local`: keep_cnt [b] int32 plus tensors keyed by name.  Keys in FIXED_KEYS have the same shape on
    every rank (per-image capacity buffers) and are gathered directly; all other tensors have the
    number of local heads as first dim (image-major) and are padded to the largest rank.  One tiny
    all-gather of the head totals, then one `dist.gather` per tensor.  Returns the concatenated dict
    on `dst`, None elsewhere."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    cnt = local["keep_cnt"]
    dev = cnt.device
    if n_heads is None:
        n_heads = int(cnt.sum())
    mine = torch.tensor([n_heads], device=dev, dtype=torch.int64)
    sizes = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(sizes, mine, group=group)
    heads = [int(s[0]) for s in sizes]
    max_heads = max(heads + [1])
    out = {} if rank == dst else None
    for key, t in local.items():
        if key in FIXED_KEYS:
            src = t.contiguous()
            bufs = [torch.empty_like(src) for _ in range(world)] if rank == dst else None
            dist.gather(src, bufs, dst=dst, group=group)
            if rank == dst:
                out[key] = torch.cat(bufs)
            continue
        pad = torch.zeros((max_heads,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
        pad[:n_heads] = t[:n_heads]
        bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
        dist.gather(pad, bufs, dst=dst, group=group)
        if rank == dst:
            out[key] = torch.cat([b[:c] for b, c in zip(bufs, heads)])
    return out
