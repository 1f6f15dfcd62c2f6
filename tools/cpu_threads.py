"""Host-thread scaling of the CPU reference arm (oracle port) - picks the fastest thread count."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from head_detector_b200 import arch, synth
from oracle import net_oracle
net = net_oracle.DeployNet(arch.synthetic_weights(0))
img = synth.synthetic_images(1, 640, 0).permute(0, 3, 1, 2).float() / 255
for t in (8, 16, 32, 64, 128):
    if t > (os.cpu_count() or 1):
        break
    torch.set_num_threads(t)
    with torch.no_grad():
        net.forward(img)
        t0 = time.perf_counter(); net.forward(img); dt = time.perf_counter() - t0
    print(f"threads {t}: {dt:.2f} s/img", flush=True)
