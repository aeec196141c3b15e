#!/usr/bin/env python3
"""Runs one conv layer through the per-layer entry with the -DY2_TC32_PROFILE build of csrc/conv_i16_tc32.cu
(profiles/build_variant_tc32.sh prof32 -DY2_TC32_PROFILE): the launcher prints the per-tile timeline of one CTA.
Usage: YOLO2CUDA_LIB=.../libyolo2cuda_prof32.so Y2_TN=32 python profiles/tc32_timeline.py"""
import os, sys
os.environ.setdefault("YOLO2CUDA_TC", "2")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "yolo-fpga-accelerator_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import make_conv_case, accel_call
from yolo2_b200.accel import Accelerator
tn = int(os.environ.get("Y2_TN", "32"))
acc = Accelerator(0, "int16")
acc.set_tile_params(tn, 32)
shapes = [tuple(int(v) for v in s.split(",")) for s in os.environ.get("Y2_SHAPES", "").split(";") if s] or [(512, 1024, 3, 52, 52)]
for (c, n, k, w, h) in shapes:
    a, x, wr, b, _ = make_conv_case(1, c, n, k, 1, w, h, 1, amp=600, xamp=2000, tn=tn)
    for _ in range(int(os.environ.get("Y2_REPS", "2"))):
        accel_call(acc, a, x, wr, b, (14, 10, 10, 10))
    print(c, n, k, w, h, acc.last_kernel, flush=True)
