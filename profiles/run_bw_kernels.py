#!/usr/bin/env python3
"""Profiling driver (for ncu): every bandwidth-bound kernel of the path once at a realistic size, plus the fp32 conv kernel.
  letterbox_kernel                64 frames 640x480 -> 416x416
  frames_to_c4 / max-pools / reorg_driver_c4 / region_kernel   one INT16 forward of YOLOv2-416 COCO at 128 frames
  conv_f32_c4_kernel              one FP32 forward at 32 frames
Usage: python profiles/run_bw_kernels.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "yolo-fpga-accelerator_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from yolo2_b200 import cfg as ycfg, weights as yw  # noqa: E402
from yolo2_b200.accel import Accelerator, letterbox_image  # noqa: E402
from yolo2_b200.model import Yolo2Net  # noqa: E402

net = ycfg.parse_network_cfg(ycfg.yolov2_cfg_text(416, 416, 80))
acc = Accelerator(0, "int16")
imgs = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (64, 480, 640, 3), dtype=np.uint8)).cuda()
for _ in range(2):
    out = letterbox_image(acc, imgs, 416, 416)
acc.synchronize()
y = Yolo2Net(net, yw.synth_pack(net, "int16", seed=0), max_batch=128, accel=acc)
frames = np.tile(yw.synth_frames(net, 4), (32, 1, 1, 1))
for _ in range(2):
    r = y.forward(frames)
y.close()
acc.close()
y32 = Yolo2Net(net, yw.synth_pack(net, "fp32", seed=0), max_batch=32)
for _ in range(2):
    r32 = y32.forward(frames[:32])
y32.close()
print("ok", r.shape, r32.shape)
