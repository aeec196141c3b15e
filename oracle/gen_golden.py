#!/usr/bin/env python3
"""TEST INFRASTRUCTURE - freezes outputs of the REAL reference into tests/golden/.

The reference ships no golden vectors, so the fixtures are produced by running
the unmodified reference (oracle/_ref/libref_*.so, built by oracle/Makefile from /root/reference)
on seeded inputs.  Run here (where /root/reference exists):   python oracle/gen_golden.py
The .npz files carry inputs AND reference outputs, so they stay valid even if numpy's RNG streams
change; tests/test_golden.py replays them through the oracle on any machine, tests/test_gpu_parity.py
through the CUDA path.
"""
import hashlib
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path = [p for p in sys.path if os.path.abspath(p or '.') != os.path.dirname(os.path.abspath(__file__))]
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "yolo-fpga-accelerator_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import make_conv_case  # noqa: E402
from oracle.oracle import Ref, align8  # noqa: E402
from yolo2_b200 import cfg as ycfg, weights as yw  # noqa: E402
from yolo2_b200.accel import pool_call_args  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
ARG_KEYS = ("IFM_num", "OFM_num", "Ksize", "Kstride", "Input_w", "Input_h", "Output_w", "Output_h", "Padding", "IsNL",
            "IsBN", "TM", "TN", "TR", "TC", "OFM_num_bound", "mLoopsxTM", "mLoops_a1xTM", "LayerType")

CONV_I16 = [  # c, n, size, stride, w, h, leaky, q, amp, xamp
    (3, 32, 3, 1, 26, 26, 1, (14, 10, 10, 10), 600, 2000),
    (64, 32, 3, 1, 13, 13, 1, (14, 10, 10, 10), 600, 2000),
    (32, 70, 3, 2, 27, 19, 0, (14, 10, 9, 12), 600, 2000),
    (16, 425, 1, 1, 13, 13, 0, (12, 12, 7, 8), 600, 2000),
    (17, 33, 3, 1, 20, 11, 1, (13, 9, 12, 7), 32767, 32767),
    (8, 8, 3, 1, 9, 9, 1, (3, 3, 10, 12), 300, 300),
    (8, 8, 1, 1, 9, 9, 1, (15, 15, 0, 31), 32767, 32767),
    (24, 32, 3, 1, 26, 13, 1, (15, 12, 2, 10), 3000, 8000),
    (20, 40, 3, 1, 19, 19, 1, (13, 9, 12, 11), 600, 2000),
]
CONV_F32 = [(3, 32, 3, 1, 26, 26, 1), (64, 48, 3, 1, 13, 13, 1), (96, 40, 1, 1, 19, 19, 0), (16, 16, 3, 2, 21, 9, 1)]
POOL = [(32, 26, 26, 2), (5, 13, 13, 2), (8, 14, 10, 1)]


def args_array(a):
    return np.array([a[k] for k in ARG_KEYS], np.int32)


def main():
    os.makedirs(OUT, exist_ok=True)
    r16, r32 = Ref("int16"), Ref("fp32")
    blob = {}
    for i, (c, n, size, stride, w, h, leaky, q, amp, xamp) in enumerate(CONV_I16):
        a, x, wr, b, _ = make_conv_case(100 + i, c, n, size, stride, w, h, leaky, amp=amp, xamp=xamp)
        out = r16.run_layer(x, wr, b, a, q)
        blob.update({f"ci16_{i}_x": x, f"ci16_{i}_w": wr, f"ci16_{i}_b": b, f"ci16_{i}_args": args_array(a),
                     f"ci16_{i}_q": np.array(q, np.int32), f"ci16_{i}_out": out})
    for i, (c, n, size, stride, w, h, leaky) in enumerate(CONV_F32):
        a, x, wr, b, _ = make_conv_case(200 + i, c, n, size, stride, w, h, leaky, dtype=np.float32, poison=1e30)
        out = r32.run_layer(x, wr, b, a)
        blob.update({f"cf32_{i}_x": x, f"cf32_{i}_w": wr, f"cf32_{i}_b": b, f"cf32_{i}_args": args_array(a), f"cf32_{i}_out": out})
    rng = np.random.default_rng(7)
    for i, (c, w, h, stride) in enumerate(POOL):
        ow, oh = (w + 1 - 2) // stride + 1, (h + 1 - 2) // stride + 1
        a = pool_call_args(c, 2, stride, w, h, ow, oh, 1)
        x = np.full((c, h, align8(w)), 32000, np.int16)
        x[:, :, :w] = rng.integers(-32768, 32768, (c, h, w))
        out = r16.run_layer(x, None, None, a)
        blob.update({f"pool_{i}_x": x, f"pool_{i}_args": args_array(a), f"pool_{i}_out": out})
        xf = x.astype(np.float32) * 40.0   # reaches below the reference's -1024*1024 pool floor
        outf = r32.run_layer(xf, None, None, a)
        blob.update({f"poolf_{i}_x": xf, f"poolf_{i}_out": outf})
    # region head + boxes/NMS from the reference's own host code
    rf = rng.normal(0, 2.0, (5 * 25, 7, 7)).astype(np.float32)
    reg = r16.region_forward(rf, 7, 7, 5, 20)
    anchors = np.array([1.3221, 1.73145, 3.19275, 4.00944, 5.05587, 8.09892, 9.47112, 4.84053, 11.2364, 10.0071], np.float32)
    boxes, probs, obj = r16.region_boxes_nms(reg, 7, 7, 5, 20, anchors, 640, 480, 224, 224, 0.3, 0.45)
    blob.update({"region_in": rf, "region_out": reg, "region_anchors": anchors, "det_boxes": boxes, "det_probs": probs, "det_obj": obj})
    np.savez_compressed(os.path.join(OUT, "layer_cases.npz"), **blob)

    # whole-network golden: the UNMODIFIED yolov2_hls_ps + get_network_boxes + do_nms_sort on COCO 416
    net = ycfg.parse_network_cfg(ycfg.yolov2_cfg_text(416, 416, 80))
    full = {}
    for tag, table, seed in (("default", "default", 1), ("stress", "stress", 2)):
        pack = yw.synth_pack(net, "int16", seed=seed, table=table)
        frame = yw.synth_frames(net, 1, seed=1000)[0]
        with tempfile.TemporaryDirectory() as d:
            yw.save_reference_files(pack, net, os.path.join(d, "weights"))
            region, boxes, probs, obj, k = r16.full_forward("/root/reference/config/yolov2.cfg", frame, d, 768, 576,
                                                            thresh=0.0, nms=0.0)
        # thresh 0 keeps every candidate (the synthetic scores are not detections); NMS is covered by layer_cases
        full.update({f"{tag}_region": region.astype(np.float32), f"{tag}_act_q": pack.act_q, f"{tag}_weight_q": pack.weight_q,
                     f"{tag}_bias_q": pack.bias_q,
                     f"{tag}_weights_sha": np.frombuffer(hashlib.sha256(pack.weights.tobytes()).digest(), np.uint8),
                     f"{tag}_frame_sha": np.frombuffer(hashlib.sha256(frame.tobytes()).digest(), np.uint8),
                     f"{tag}_boxes": boxes[:k], f"{tag}_obj": obj[:k]})
        print(tag, "region sha", hashlib.sha256(region.tobytes()).hexdigest()[:16], "candidates", k)
    np.savez_compressed(os.path.join(OUT, "yolov2_coco416_full.npz"), **full)
    for f in os.listdir(OUT):
        print(f, os.path.getsize(os.path.join(OUT, f)) >> 10, "KiB")


LETTERBOX = [  # (image w, h, c, net_w, net_h): landscape, portrait, non-square net, upscale, one-pixel-wide, no-op size
    (97, 61, 3, 64, 64), (61, 97, 3, 64, 64), (50, 50, 3, 64, 48), (20, 15, 3, 48, 48), (1, 5, 3, 8, 8), (32, 32, 1, 32, 32),
    (160, 120, 3, 104, 104)]


def gen_letterbox():
    """letterbox_image of the unmodified reference (src/core/yolo_image.cpp:146-165) on seeded stb-layout u8 images."""
    os.makedirs(OUT, exist_ok=True)
    r16 = Ref("int16")
    blob = {}
    for i, (w, h, c, nw, nh) in enumerate(LETTERBOX):
        img = np.random.default_rng(7000 + i).integers(0, 256, (h, w, c), dtype=np.uint8)
        blob[f"lb_{i}_img"] = img
        blob[f"lb_{i}_net"] = np.array([nw, nh], np.int32)
        blob[f"lb_{i}_out"] = r16.letterbox_u8(img, nw, nh)
    np.savez_compressed(os.path.join(OUT, "letterbox.npz"), **blob)
    print("letterbox.npz", os.path.getsize(os.path.join(OUT, "letterbox.npz")) >> 10, "KiB")


VARIANTS = [  # (tn, c, n, size, w, h, q, amp): YOLO2_FPGA of the reference BUILT with --tn <tn> (oracle/Makefile ref-variants)
    (8, 24, 40, 3, 13, 13, (14, 10, 10, 10), 600), (8, 37, 33, 1, 20, 11, (13, 9, 12, 7), 32767),
    (32, 64, 128, 3, 13, 13, (14, 10, 10, 10), 600), (32, 37, 130, 3, 20, 11, (13, 9, 12, 7), 32767), (32, 96, 40, 1, 19, 19, (10, 10, 7, 8), 3000)]


def gen_variants():
    os.makedirs(OUT, exist_ok=True)
    blob = {}
    for i, (tn, c, n, size, w, h, q, amp) in enumerate(VARIANTS):
        a, x, wr, b, _ = make_conv_case(500 + i, c, n, size, 1, w, h, 1, amp=amp, xamp=32767 if amp > 600 else 2000, tn=tn)
        out = Ref("int16", tn).run_layer(x, wr, b, a, q)
        blob.update({f"v_{i}_x": x, f"v_{i}_w": wr, f"v_{i}_b": b, f"v_{i}_args": args_array(a), f"v_{i}_q": np.array(q, np.int32),
                     f"v_{i}_tn": np.array([tn], np.int32), f"v_{i}_out": out})
    np.savez_compressed(os.path.join(OUT, "layer_cases_tn_variants.npz"), **blob)
    print("layer_cases_tn_variants.npz", os.path.getsize(os.path.join(OUT, "layer_cases_tn_variants.npz")) >> 10, "KiB")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "letterbox":
        gen_letterbox()
    elif len(sys.argv) > 1 and sys.argv[1] == "variants":
        gen_variants()
    else:
        main()
        gen_letterbox()
        gen_variants()
