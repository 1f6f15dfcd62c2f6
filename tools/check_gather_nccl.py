"""NCCL smoke of parallel.gather_predictions on CUDA tensors (world size = number of ranks launched; 1 is enough to
exercise the CUDA code path of the packing).  usage: python tools/check_gather_nccl.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from head_detector_b200 import parallel  # noqa: E402

rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29533")
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", torch.cuda.current_device()))
B, K = 32, 100
for n in (0, 7, 256):
    g = torch.Generator(device="cuda").manual_seed(rank * 10 + n)
    cnt = torch.zeros(B, dtype=torch.int32, device="cuda")
    cnt[: min(n, B)] = 1
    local = {"keep_cnt": cnt, "boxes": torch.rand(B * K, 4, device="cuda", generator=g), "scores": torch.rand(B * K, device="cuda", generator=g),
             "params": torch.rand(max(n, 1), 413, device="cuda", generator=g)[:n], "verts": torch.rand(max(n, 1), 5023, 3, device="cuda", generator=g)[:n]}
    out = parallel.gather_predictions(local, n_heads=n)
    torch.cuda.synchronize()
    if rank == 0 and world == 1:
        for k, v in local.items():
            assert out[k].dtype == v.dtype and torch.equal(out[k], v), k
print("nccl gather ok", rank, world)
dist.barrier()
dist.destroy_process_group()
