#!/usr/bin/env python3
"""Runs one wide conv layer through the per-layer entry with the -DY2_TC2_PROFILE build of the library
(lib/libyolo2cuda_prof.so): the kernel prints, per warp role of one CTA, the clock64 cycles spent waiting on
each pipeline barrier.  Usage: YOLO2CUDA_LIB=.../libyolo2cuda_prof.so YOLO2CUDA_TC=2 python profiles/tc2_role_profile.py"""
import os, sys
os.environ.setdefault("YOLO2CUDA_TC", "2")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "yolo-fpga-accelerator_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import make_conv_case, accel_call
from yolo2_b200.accel import Accelerator
acc = Accelerator(0, "int16")
shapes = [tuple(int(v) for v in s.split(",")) for s in os.environ.get("Y2_SHAPES", "").split(";") if s] or [(512, 256, 3, 13, 13), (64, 128, 3, 104, 104), (1024, 512, 1, 13, 13)]
for (c, n, k, w, h) in shapes:
    a, x, wr, b, _ = make_conv_case(1, c, n, k, 1, w, h, 1, amp=600, xamp=2000)
    for _ in range(int(os.environ.get("Y2_REPS", "2"))):
        accel_call(acc, a, x, wr, b, (14, 10, 10, 10))
    print(c, n, k, w, h, acc.last_kernel, flush=True)
