#!/usr/bin/env python3
"""Runs every distinct wide conv shape of YOLOv2-416 through the per-layer entry with the tcgen05 path, one
subprocess per shape with a timeout, and checks each against the CUDA-core kernel (bit-exact).  Development aid."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHAPES = [tuple(int(v) for v in s.split(",")) for s in os.environ.get("Y2_SHAPES", "").split(";") if s] or [(64, 128, 3, 104, 104), (128, 256, 3, 52, 52), (256, 128, 1, 52, 52), (256, 512, 3, 26, 26), (512, 256, 1, 26, 26),
          (512, 1024, 3, 13, 13), (1024, 512, 1, 13, 13), (1024, 1024, 3, 13, 13), (1280, 1024, 3, 13, 13), (1024, 425, 1, 13, 13)]
CHILD = r'''
import os, sys, numpy as np
sys.path.insert(0, "{root}"); sys.path.insert(0, "{root}/yolo-fpga-accelerator_b200"); sys.path.insert(0, "{root}/tests")
from helpers import make_conv_case, accel_call
from yolo2_b200.accel import Accelerator
c, n, k, w, h = {shape}
a, x, wr, b, _ = make_conv_case(c + n, c, n, k, 1, w, h, 1, amp=600, xamp=2000)
os.environ["YOLO2CUDA_TC"] = "{tc}"
acc = Accelerator(0, "int16")
got = accel_call(acc, a, x, wr, b, (14, 10, 10, 10))
name = acc.last_kernel
os.environ["YOLO2CUDA_TC"] = "0"
ref = Accelerator(0, "int16")
want = accel_call(ref, a, x, wr, b, (14, 10, 10, 10))
print({shape}, name, "vs", ref.last_kernel, "mismatch", int((got != want).sum()))
'''
tc = os.environ.get("YOLO2CUDA_TC", "2")
for s in SHAPES:
    try:
        r = subprocess.run([sys.executable, "-c", CHILD.format(root=ROOT, shape=s, tc=tc)], capture_output=True, text=True, timeout=int(os.environ.get("Y2_TIMEOUT", "40")))
        print(r.stdout.strip() or r.stderr.strip()[-300:], flush=True)
    except subprocess.TimeoutExpired:
        print(s, "TIMEOUT", flush=True)
