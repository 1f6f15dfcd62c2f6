"""Batch-sharded multi-GPU inference: one process per GPU, images sharded by rank, no collective on
the data path except the final gather of predictions to rank 0 (BASELINE.json north_star; SURVEY.md
8e).  torch.distributed (NCCL on GPUs, gloo in the CPU tests) is plumbing only."""
import math
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
    """Ragged gather of one step's predictions to rank `dst` in TWO collectives.

    `local`: keep_cnt [b] int32 plus tensors keyed by name.  Keys in FIXED_KEYS have the same shape on every
    rank (per-image capacity buffers); all other tensors have the number of local heads as first dim
    (image-major).  One tiny all-gather of the head totals; then every rank packs
    `[fixed tensors | ragged tensors (its own n heads) | padding to the largest rank]` into ONE float32 record
    buffer (a single `torch.cat`) and ONE `dist.gather` moves it; `dst` slices the records apart again
    (integer tensors travel as float32 - counts and ids below 2^24 are exact - and get their dtype back).
    Returns the concatenated dict on `dst`, None elsewhere."""
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
    keys = list(local.keys())
    fixed = [k for k in keys if k in FIXED_KEYS]
    ragged = [k for k in keys if k not in FIXED_KEYS]
    row_w = {k: math.prod(local[k].shape[1:]) for k in ragged}
    fixed_len = sum(local[k].numel() for k in fixed)
    max_len = fixed_len + max(heads + [0]) * sum(row_w.values())
    parts = [local[k].reshape(-1).to(torch.float32) for k in fixed] + [local[k][:n_heads].reshape(-1).to(torch.float32) for k in ragged]
    used = fixed_len + n_heads * sum(row_w.values())
    buf = torch.zeros(max(max_len, 1), dtype=torch.float32, device=dev)
    if used:
        torch.cat(parts, out=buf[:used])
    bufs = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    out: Dict[str, torch.Tensor] = {}
    pieces: Dict[str, list] = {k: [] for k in keys}
    for r, rb in enumerate(bufs):
        at = 0
        for k in fixed:
            n = local[k].numel()
            pieces[k].append(rb[at:at + n].reshape(local[k].shape))
            at += n
        for k in ragged:
            n = heads[r] * row_w[k]
            pieces[k].append(rb[at:at + n].reshape((heads[r],) + tuple(local[k].shape[1:])))
            at += n
    for k in keys:
        t = torch.cat(pieces[k])
        out[k] = t if local[k].dtype == torch.float32 else t.round().to(local[k].dtype)
    return out
