import os, sys, numpy as np
os.environ.setdefault("YOLO2CUDA_TC", "1")
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/yolo-fpga-accelerator_b200"); sys.path.insert(0, "/root/repo/tests")
from helpers import make_conv_case, oracle_conv, accel_call, valid
from oracle.oracle import Oracle
from yolo2_b200.accel import Accelerator
o = Oracle(); acc = Accelerator(0, "int16")
cases = [(64, 128, 3, 26, 26, (14,10,10,10)), (4, 128, 1, 8, 8, (14,10,10,10)), (28, 128, 1, 8, 8, (14,10,10,10)), (8, 128, 3, 13, 13, (14,10,10,10)), (64, 128, 3, 13, 13, (14,10,10,10)),
         (64, 200, 3, 26, 26, (13, 9, 12, 7)), (256, 256, 3, 13, 13, (15, 12, 8, 10)), (96, 40, 1, 19, 19, (12,12,7,8)), (17, 33, 3, 20, 11, (13,9,12,7))]
for (c, n, k, w, h, q) in cases:
    a, x, wr, b, _ = make_conv_case(c*n+k, c, n, k, 1, w, h, 1, amp=32767 if c==17 else 600, xamp=32767 if c==17 else 2000)
    want = oracle_conv(o, a, x, wr, b, q)
    got = accel_call(acc, a, x, wr, b, q)
    d = valid(got, w).astype(int) - valid(want, w).astype(int)
    print(c, n, k, w, h, q, acc.last_kernel, "mismatch", int((d != 0).sum()), "of", d.size, "maxabs", int(np.abs(d).max()))
    if (d != 0).any():
        idx = np.argwhere(d != 0)[:5]; print("  first bad (m,y,x):", idx.tolist(), "got", [int(got[tuple(i)]) for i in idx], "want", [int(want[tuple(i)]) for i in idx])
