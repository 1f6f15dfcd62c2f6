"""Runs bench.parity_check on one small sparse-heads step (sanity of the checker block of bench.py)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from head_detector_b200 import arch, synth
from head_detector_b200.engine import Engine
B = 4
eng = Engine(arch.synthetic_weights(0), B, 640, sparse_heads=True)
boxes, scores = synth.engineered_heads(B, eng.A, 640, 8, seed=7)
eng.set_override(boxes.cuda(), scores.cuda())
eng.input.copy_(synth.synthetic_images(B, 640, 0).cuda())
eng.run_device(bench.CONF, bench.IOU, bench.TOPK)
print(bench.parity_check(eng, boxes, scores, B))
