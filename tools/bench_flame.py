"""Stand-alone timing of the fused FLAME decode (vgh_flame_decode) -> gpurun_out/flame_<tag>.json.
Per head count: device time (CUDA events, L2 flushed between launches), FP64 FMA rate, algorithmic HBM GB/s
(61 940 B per head + the live basis once per launch)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from head_detector_b200 import synth  # noqa: E402
from head_detector_b200.flame import FLAMELayer  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "r2b"
fl = FLAMELayer()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = {}
for n, live in ((8, (128, 64)), (64, (128, 64)), (512, (128, 64)), (984, (128, 64)), (4096, (128, 64)), (512, (300, 100))):
    g = torch.Generator().manual_seed(n)
    p = torch.zeros(n, 413)
    p[:, :live[0]] = 3 * torch.tanh(torch.randn(n, live[0], generator=g))
    p[:, 300:300 + live[1]] = 3 * torch.tanh(torch.randn(n, live[1], generator=g))
    p[:, 400:403] = 0.1 * torch.randn(n, 3, generator=g)
    p[:, 403:409] = torch.randn(n, 6, generator=g)
    p[:, 409:412] = 300 * torch.rand(n, 3, generator=g)
    p[:, 412] = 200 + 400 * torch.rand(n, generator=g)
    p = p.cuda()
    for _ in range(3):
        fl.decode(p, live=live)
    ts = []
    for _ in range(10):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fl.decode(p, live=live)   # includes the three output allocations of the python wrapper (cached allocator: no cudaMalloc)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    us = sorted(ts)[len(ts) // 2]
    k = live[0] + live[1] + 9
    fma = n * 5120 * 3 * ((k + 7) // 8 * 8)   # executed (padded vertices and coefficient rows included)
    alg_bytes = n * 61940 + k * 5023 * 3 * 4
    out[f"{n}x{live[0]}+{live[1]}"] = {"heads": n, "live": live, "us": us, "fp64_tfma_s": fma / us / 1e6, "fma_per_clk_sm_at_1.9GHz": fma / (us * 1e-6) / 148 / 1.9e9,
                                      "algorithmic_GBps": alg_bytes / us / 1e3}
    print(n, live, f"{us:.1f} us  {fma / us / 1e6:.2f} TFMA/s  {alg_bytes / us / 1e3:.0f} GB/s algorithmic")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"flame_{tag}.json"), "w"), indent=1)
