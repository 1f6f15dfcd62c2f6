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


def gather_predictions(local: Dict[str, torch.Tensor], dst: int = 0, group=None) -> Optional[Dict[str, torch.Tensor]]:
    """Ragged gather: `local` holds keep_cnt [b] int32 and per-head tensors (first dim = number of
    local heads, image-major).  Step 1 all-gathers the per-rank head totals and image counts,
    step 2 gathers each per-head tensor padded to the largest rank.  Returns the concatenated
    dict on `dst`, None elsewhere."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    cnt = local["keep_cnt"]
    dev = cnt.device
    n_local = torch.tensor([int(cnt.sum()), cnt.numel()], device=dev, dtype=torch.int64)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    heads = [int(s[0]) for s in sizes]
    imgs = [int(s[1]) for s in sizes]
    max_heads, max_imgs = max(heads + [1]), max(imgs)
    out = {} if rank == dst else None

    def gather_padded(t, n_rows, max_rows, counts):
        pad = torch.zeros((max_rows,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
        pad[:n_rows] = t[:n_rows]
        bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
        dist.gather(pad, bufs, dst=dst, group=group)
        return torch.cat([b[:c] for b, c in zip(bufs, counts)]) if rank == dst else None

    g = gather_padded(cnt, cnt.numel(), max_imgs, imgs)
    if rank == dst:
        out["keep_cnt"] = g
    for key, t in local.items():
        if key == "keep_cnt":
            continue
        g = gather_padded(t, heads[rank], max_heads, heads)
        if rank == dst:
            out[key] = g
    return out
