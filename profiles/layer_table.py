#!/usr/bin/env python3
"""Per-layer device times (CUDA events inside the library) of one full YOLOv2 pass (default 416 COCO INT16; Y2_SIZE / Y2_CLASSES / Y2_TN),
with each kernel's roofline: exact round-and-saturate steps/s (and the int8-OP equivalent) for conv,
algorithmic GB/s for the bandwidth kernels.  Usage: python profiles/layer_table.py [batch] [precision]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "yolo-fpga-accelerator_b200"))
import numpy as np  # noqa: E402
from yolo2_b200 import cfg as ycfg, weights as yw  # noqa: E402
from yolo2_b200.model import Yolo2Net  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
precision = sys.argv[2] if len(sys.argv) > 2 else "int16"
eb = 2 if precision == "int16" else 4
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}
size = int(os.environ.get("Y2_SIZE", "416"))        # BASELINE configs: 416 (COCO / VOC), 608 (COCO)
classes = int(os.environ.get("Y2_CLASSES", "80"))   # 80 = COCO, 20 = VOC
net = ycfg.parse_network_cfg(ycfg.yolov2_cfg_text(size, size, classes))
tn = int(os.environ.get("Y2_TN", "4"))          # emulate a reference built with --tn <Y2_TN> (yolo2cuda_set_tile_params)
pack = yw.synth_pack(net, precision, seed=0, tn=tn)
from yolo2_b200.accel import Accelerator  # noqa: E402
acc = Accelerator(0, precision)
acc.set_tile_params(tn, 32)
y = Yolo2Net(net, pack, max_batch=batch, accel=acc)
frames = np.tile(yw.synth_frames(net, 4), (batch // 4 + 1, 1, 1, 1))[:batch]
for _ in range(2):
    y.forward(frames)
y.layer_times()
y.forward(frames)
ms = y.layer_times()
names = {0: "conv", 1: "maxpool", 2: "reorg", 3: "route", 4: "region"}
rows, tot = [], float(ms.sum())
for i, (l, t) in enumerate(zip(net.layers, ms)):
    r = {"layer": i, "type": names[l.type], "shape": f"{l.c}x{l.h}x{l.w}->{l.out_c}x{l.out_h}x{l.out_w}", "k": l.size, "ms": round(float(t), 4),
         "share": round(float(t) / tot, 4)}
    if l.type == ycfg.CONV:
        steps = ((l.c + 3) // 4) * l.size * l.size * l.n * l.out_h * l.out_w * batch
        macs = l.c * l.size * l.size * l.n * l.out_h * l.out_w * batch
        t = float(t); r["steps_per_s"] = steps / (t * 1e-3)
        r["int8_equiv_TOPs"] = macs * 8 / (t * 1e-3) / 1e12
        r["frac_of_int8_peak"] = r["int8_equiv_TOPs"] / (2 * peaks["bf16_tflops_sustained"])
    elif l.type in (ycfg.MAXPOOL, ycfg.REORG):
        t = float(t); by = (l.c * l.h * l.w + l.out_c * l.out_h * l.out_w) * eb * batch
        r["GBps"] = by / (t * 1e-3) / 1e9
        r["frac_of_hbm_peak"] = r["GBps"] / peaks["hbm_gbs"]
    elif l.type == ycfg.REGION:
        t = float(t); by = (l.c * l.h * l.w * eb + l.c * l.h * l.w * 4) * batch
        r["GBps"] = by / (t * 1e-3) / 1e9
        r["frac_of_hbm_peak"] = r["GBps"] / peaks["hbm_gbs"]
    r["kernel"] = y.layer_kernel(i)
    rows.append(r)
fast_tiles, exact_tiles = acc.tc_path_counts() if precision == "int16" else (0, 0)
print(json.dumps({"tc_fast_tiles": fast_tiles, "tc_exact_tiles": exact_tiles, "activation_bytes": y.activation_bytes, "batch": batch, "precision": precision, "tn": tn, "size": size, "classes": classes, "total_ms": tot, "fps": batch / (tot * 1e-3), "layers": rows}, indent=1))
