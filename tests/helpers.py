"""Shared builders for seeded test cases (reference layouts)."""
import numpy as np

from yolo2_b200.accel import conv_call_args, pool_call_args
from yolo2_b200.weights import weight_reorg


def align8(w):
    return (w + 7) & ~7


def make_conv_case(seed, c, n, size, stride, w, h, leaky, dtype=np.int16, amp=600, xamp=2000, pad=None, poison=12345, tn=4, tm=32):
    rng = np.random.default_rng(seed)
    pad = size // 2 if pad is None else pad
    a = conv_call_args(c, n, size, stride, w, h, pad, leaky, tn=tn, tm=tm)
    x = np.full((c, h, align8(w)), poison, dtype)           # poisoned pad columns must not matter
    if dtype == np.int16:
        x[:, :, :w] = rng.integers(-xamp, xamp + 1, (c, h, w))
        wd = rng.integers(-amp, amp + 1, (n, c, size * size)).astype(np.int16)
        b = rng.integers(-2000, 2001, n).astype(np.int16)
    else:
        x[:, :, :w] = rng.normal(0, 1, (c, h, w))
        wd = rng.normal(0, 0.05, (n, c, size * size)).astype(np.float32)
        b = rng.normal(0, 0.5, n).astype(np.float32)
    wr = weight_reorg(wd, c, n, size, a["TM"], a["TN"])
    return a, x, wr, b, wd


def oracle_conv(o, a, x, wr, b, q=(0, 0, 0, 0)):
    return o.conv(x, wr, b, a["IFM_num"], a["OFM_num"], a["Ksize"], a["Kstride"], a["Input_w"], a["Input_h"],
                  a["Output_w"], a["Output_h"], a["Padding"], a["IsNL"], a["TM"], a["TN"], *q)


def accel_call(acc, a, x, w, b, q=(0, 0, 0, 0), out=None):
    if out is None:
        out = np.zeros((a["OFM_num"], a["Output_h"], align8(a["Output_w"])), acc.dtype)
    acc.YOLO2_FPGA(x, out, w, b, *[a[k] for k in (
        "IFM_num", "OFM_num", "Ksize", "Kstride", "Input_w", "Input_h", "Output_w", "Output_h", "Padding", "IsNL",
        "IsBN", "TM", "TN", "TR", "TC", "OFM_num_bound", "mLoopsxTM", "mLoops_a1xTM", "LayerType")], *q)
    return out


def valid(t, w):
    return t[..., :w]
