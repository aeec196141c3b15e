#!/usr/bin/env python3
"""Profiling driver (for ncu): one wide 3x3 conv layer through the per-layer entry with the tcgen05 path
(YOLO2CUDA_TC=2, csrc/conv_i16_tc2.cu).  Usage: python profiles/run_tc2_layer.py [c n k w h]"""
import os
import sys

os.environ.setdefault("YOLO2CUDA_TC", "2")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "yolo-fpga-accelerator_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import make_conv_case, accel_call  # noqa: E402
from yolo2_b200.accel import Accelerator  # noqa: E402

c, n, k, w, h = (int(v) for v in sys.argv[1:6]) if len(sys.argv) >= 6 else (512, 1024, 3, 52, 52)
acc = Accelerator(0, "int16")
a, x, wr, b, _ = make_conv_case(1, c, n, k, 1, w, h, 1, amp=600, xamp=2000)
for _ in range(2):
    accel_call(acc, a, x, wr, b, (14, 10, 10, 10))
print("ok", acc.last_kernel)
