"""Device-timed letterbox (vgh_letterbox) on a batch of 1080p frames: us/batch and algorithmic GB/s
(source bytes read once + S*S*3 written) against the measured HBM peak.  `ncu` mode: one launch between
cudaProfilerStart/Stop.

    python tools/bench_letterbox.py [batch] [h] [w] [ncu]"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from head_detector_b200 import _lib  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
H = int(sys.argv[2]) if len(sys.argv) > 2 else 1080
W = int(sys.argv[3]) if len(sys.argv) > 3 else 1920
ncu = len(sys.argv) > 4 and sys.argv[4] == "ncu"
S = 640
src = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device="cuda")
out = torch.empty(B, S, S, 3, dtype=torch.uint8, device="cuda")
offs = (np.arange(B, dtype=np.int64) * H * W * 3)
hs, ws = np.full(B, H, np.int32), np.full(B, W, np.int32)
xf = np.zeros((B, 3), np.float32)
stream = torch.cuda.current_stream()


def launch():
    _lib.check(_lib.lib().vgh_letterbox(src.data_ptr(), offs.ctypes.data, hs.ctypes.data, ws.ctypes.data, B, S, out.data_ptr(),
                                        xf.ctypes.data, C.c_void_p(stream.cuda_stream)), "vgh_letterbox")


for _ in range(3):
    launch()
torch.cuda.synchronize()
if ncu:
    torch.cuda.profiler.start()
    launch()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    sys.exit(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
times = []
for _ in range(10):
    flush.fill_(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    launch()
    e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
ms = sorted(times)[len(times) // 2]
alg = B * (H * W * 3 + S * S * 3)
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
gbs = alg / (ms * 1e-3) / 1e9
print(json.dumps({"kernel": "letterbox_kernel", "batch": B, "src": [H, W], "ms_per_launch_incl_table_upload": ms, "algorithmic_bytes": alg,
                  "achieved_gbs": gbs, "hbm_peak_gbs": peaks["hbm_gbs"], "frac": gbs / peaks["hbm_gbs"], "images_per_s": B / (ms * 1e-3)}))
