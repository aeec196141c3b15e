#!/usr/bin/env python3
"""One conv layer through the per-layer entry on the tcgen05 kernel, compared with the oracle (debug driver).
Usage: [YOLO2CUDA_LIB=...] python profiles/tc2_one_case.py c,n,k,w,h [...]"""
import os, sys
os.environ.setdefault("YOLO2CUDA_TC", "2")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "yolo-fpga-accelerator_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import make_conv_case, accel_call, oracle_conv, valid
from yolo2_b200.accel import Accelerator
from oracle.oracle import Oracle
orc = Oracle()
acc = Accelerator(0, "int16")
for s in sys.argv[1:]:
    c, n, k, w, h = (int(v) for v in s.split(","))
    a, x, wr, b, _ = make_conv_case(1, c, n, k, 1, w, h, 1, amp=600, xamp=2000)
    q = (14, 10, 10, 10)
    try:
        got = accel_call(acc, a, x, wr, b, q)
        want = oracle_conv(orc, a, x, wr, b, q)
        bad = int((valid(got, w) != valid(want, w)).sum())
        print(s, acc.last_kernel, "mismatches", bad, "of", valid(want, w).size, flush=True)
    except Exception as e:
        print(s, "ERROR", str(e)[:200], flush=True)
        break
