"""Synthetic workload of BASELINE.json's configs (no dataset or checkpoint is reachable offline).

Random-init weights never produce detections (scores ~0.01), so - as SURVEY.md 8d prescribes - the
head population is engineered: per image `heads` clusters of `per_cluster` overlapping anchors
(IoU > 0.5 inside a cluster, < 0.1 across) with scores U(0.55, 0.95); every other anchor scores
U(0, 0.3); a 1e-7 * anchor_id ramp makes scores tie-free.  NMS must reduce each cluster to one head."""
import torch


def synthetic_images(batch: int, size: int, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (batch, size, size, 3), generator=g, dtype=torch.uint8)


def engineered_heads(batch: int, num_anchors: int, size: int, heads: int = 8, per_cluster: int = 12, seed: int = 0):
    """-> boxes [B,A,4] xyxy fp32, scores [B,A] fp32."""
    g = torch.Generator().manual_seed(seed)
    grid = 1
    while grid * grid < heads:
        grid += 1
    cell = size / grid
    half_lo, half_hi = 0.125 * cell * 0.5 * 3, 0.28 * cell * 0.5 * 3   # 40..90 px at 640 / 3x3
    boxes = torch.rand(batch, num_anchors, 4, generator=g) * size
    boxes = torch.stack([torch.minimum(boxes[..., 0], boxes[..., 2]), torch.minimum(boxes[..., 1], boxes[..., 3]),
                         torch.maximum(boxes[..., 0], boxes[..., 2]) + 1, torch.maximum(boxes[..., 1], boxes[..., 3]) + 1], dim=-1)
    scores = torch.rand(batch, num_anchors, generator=g) * 0.3
    for b in range(batch):
        cells = torch.randperm(grid * grid, generator=g)[:heads]
        anchors = torch.randperm(num_anchors, generator=g)[: heads * per_cluster].view(heads, per_cluster)
        for c in range(heads):
            cy, cx = divmod(int(cells[c]), grid)
            jitter = (torch.rand(2, generator=g) - 0.5) * 0.15 * cell
            ctr = torch.tensor([(cx + 0.5) * cell, (cy + 0.5) * cell]) + jitter
            half = half_lo + (half_hi - half_lo) * torch.rand(1, generator=g)
            base = torch.cat([ctr - half, ctr + half])
            boxes[b, anchors[c]] = base + (torch.rand(per_cluster, 4, generator=g) - 0.5) * 12.0 * (size / 640.0)
            scores[b, anchors[c]] = 0.55 + 0.4 * torch.rand(per_cluster, generator=g)
    scores = scores + torch.arange(num_anchors) * 1e-7
    return boxes.float().contiguous(), scores.float().contiguous()
