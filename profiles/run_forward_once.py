#!/usr/bin/env python3
"""Profiling driver: N forwards of the full YOLOv2-416 COCO INT16 net at a given batch (for ncu)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "yolo-fpga-accelerator_b200"))
import numpy as np  # noqa: E402
from yolo2_b200 import cfg as ycfg, weights as yw  # noqa: E402
from yolo2_b200.model import Yolo2Net  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
net = ycfg.parse_network_cfg(ycfg.yolov2_cfg_text(416, 416, 80))
pack = yw.synth_pack(net, "int16", seed=0)
y = Yolo2Net(net, pack, max_batch=batch)
frames = np.tile(yw.synth_frames(net, 4), (batch // 4 + 1, 1, 1, 1))[:batch]
for _ in range(reps):
    r = y.forward(frames)
print("ok", r.shape, y.launches_per_forward)
