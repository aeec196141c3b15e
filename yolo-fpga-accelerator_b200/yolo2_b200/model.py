"""Whole-network path: the `--backend cuda` replacement for yolov2_hls_ps.

Reference interface: void yolov2_hls_ps(network*, const float* input, Precision)
(hls/models/yolov2/yolo2_accel.hpp:21-23, defined yolo2_model.cpp:229-449): run every layer on
the accelerator and leave the region tensor in net->layers[n-1].output.  Yolo2Net does the same
for a batch of frames; yolov2_cuda_ps keeps the single-frame signature.
"""
import ctypes as C

import numpy as np

from . import _capi, cfg as _cfg
from .accel import Accelerator
from .weights import WeightsPack


class Yolo2Net:
    def __init__(self, net: _cfg.Network, pack: WeightsPack, device: int = 0, max_batch: int = 32, accel: Accelerator = None):
        self.net = net
        self.precision = "int16" if pack.is_int16 else "fp32"
        self.accel = accel or Accelerator(device, self.precision)
        if self.accel.precision != self.precision:
            raise _capi.Yolo2CudaError(_capi.ERROR, "context precision does not match the weight pack")
        self.lib = self.accel.lib
        self.max_batch = max_batch
        self.handle = C.c_void_p()
        self._descs = _cfg.to_desc_array(net)
        _capi.check(self.accel.ctx, self.lib.yolo2cuda_net_create(self.accel.ctx, self._descs, len(net.layers), max_batch,
                                                                  C.byref(self.handle)))
        self.load_weights(pack)
        last = net.layers[-1]
        self.region_outputs = last.c * last.h * last.w

    def load_weights(self, pack: WeightsPack):
        w = np.ascontiguousarray(pack.weights)
        b = np.ascontiguousarray(pack.bias)
        vp = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
        wq = np.ascontiguousarray(pack.weight_q, np.int32) if pack.weight_q is not None else None
        bq = np.ascontiguousarray(pack.bias_q, np.int32) if pack.bias_q is not None else None
        aq = np.ascontiguousarray(pack.act_q, np.int32) if pack.act_q is not None else None
        nq = min(len(wq), len(bq)) if wq is not None else 0
        rc = self.lib.yolo2cuda_net_load_weights(self.handle, vp(w), w.size, vp(b), b.size, vp(wq), vp(bq), nq,
                                                 vp(aq), len(aq) if aq is not None else 0)
        _capi.check(self.accel.ctx, rc)

    def set_ramp_frames(self, frames: int):
        """first (short) pass of a multi-pass host forward: yolo2cuda_net_set_ramp_frames (-1 default, 0 none)"""
        _capi.check(self.accel.ctx, self.lib.yolo2cuda_net_set_ramp_frames(self.handle, int(frames)))

    def close(self):
        if getattr(self, "handle", None):
            self.lib.yolo2cuda_net_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- forward ---------------------------------------------------------------------------------
    def forward(self, frames: np.ndarray, out: np.ndarray = None) -> np.ndarray:
        """frames: float32 [B][c][h][w] on the HOST (pinned or not). Returns the region tensors
        float32 [B][n][coords+1+classes][h][w] (layers[n-1].output per frame)."""
        frames = np.ascontiguousarray(frames, np.float32)
        B = frames.shape[0]
        if out is None:
            out = np.empty((B, self.region_outputs), np.float32)
        rc = self.lib.yolo2cuda_net_forward_host(self.handle, frames.ctypes.data_as(C.c_void_p), B,
                                                 out.ctypes.data_as(C.c_void_p))
        _capi.check(self.accel.ctx, rc)
        return self._shape_region(out, B)

    def forward_images(self, images: np.ndarray) -> np.ndarray:
        """Raw stb-layout images uint8 [batch][h][w][c] (all one size) -> letterbox on the GPU -> network.
        Reference call sequence: load_image_stb, letterbox_image, yolov2_hls_ps (yolov2_main.cpp:255-292)."""
        images = np.ascontiguousarray(images, dtype=np.uint8)
        B, ih, iw, ic = images.shape
        assert ic == self.net.c
        out = np.empty((B, self.region_outputs), np.float32)
        rc = self.lib.yolo2cuda_net_forward_images_host(self.handle, images.ctypes.data_as(C.c_void_p), B, iw, ih,
                                                        out.ctypes.data_as(C.c_void_p))
        _capi.check(self.accel.ctx, rc)
        return self._shape_region(out, B)

    def forward_ptr(self, frames_ptr: int, batch: int, out_ptr: int, device: bool):
        """Raw-pointer form (torch tensors: .data_ptr()). device=True is asynchronous."""
        fn = self.lib.yolo2cuda_net_forward_dev if device else self.lib.yolo2cuda_net_forward_host
        _capi.check(self.accel.ctx, fn(self.handle, C.c_void_p(frames_ptr), batch, C.c_void_p(out_ptr)))

    def _shape_region(self, flat, B):
        l = self.net.layers[-1]
        return flat.reshape(B, l.n, l.coords + 1 + l.classes, l.h, l.w)

    def set_debug_keep(self, keep: bool = True):
        """Give every layer its own output buffer so that layer_output works for every layer (yolo2cuda_net_set_debug_keep);
        the default is one liveness-packed arena in which buffers are recycled down the network."""
        _capi.check(self.accel.ctx, self.lib.yolo2cuda_net_set_debug_keep(self.handle, int(bool(keep))))

    @property
    def activation_bytes(self) -> int:
        return int(self.lib.yolo2cuda_net_activation_bytes(self.handle))

    def layer_output(self, layer: int, frame: int = 0) -> np.ndarray:
        """ofm of `layer` for `frame` of the last forward, reference layout [out_c][out_h][ceil8 out_w]."""
        l = self.net.layers[layer]
        wa = (l.out_w + 7) & ~7
        dst = np.zeros((l.out_c, l.out_h, wa), self.accel.dtype)
        rc = self.lib.yolo2cuda_net_get_layer_output(self.handle, layer, frame, dst.ctypes.data_as(C.c_void_p), dst.size)
        _capi.check(self.accel.ctx, rc)
        return dst

    @property
    def region_q(self):
        return int(self.lib.yolo2cuda_net_region_q(self.handle))

    @property
    def launches_per_forward(self):
        return int(self.lib.yolo2cuda_net_launches_per_forward(self.handle))

    def layer_kernel(self, layer: int) -> str:
        """kernel variant the last forward launched for `layer` ("" for route layers)"""
        return self.lib.yolo2cuda_net_layer_kernel(self.handle, layer).decode()

    def layer_times(self):
        ms = np.zeros(len(self.net.layers), np.float32)
        _capi.check(self.accel.ctx, self.lib.yolo2cuda_net_layer_times(self.handle, ms.ctypes.data_as(C.c_void_p), ms.size))
        return ms

    # -- detections (get_network_boxes + do_nms_sort, yolov2_main.cpp:311-320) --------------------
    def detections(self, region: np.ndarray, im_w: int, im_h: int, thresh: float = 0.25, nms: float = 0.45):
        return region_detections(self.net, region, im_w, im_h, thresh, nms)


def region_detections(net: _cfg.Network, region: np.ndarray, im_w: int, im_h: int, thresh=0.25, nms=0.45):
    """One frame's region tensor -> (boxes[k][4] x,y,w,h relative, probs[k][classes], objectness[k])."""
    lib = _capi.load_library()
    l = net.layers[-1]
    total = l.w * l.h * l.n
    region = np.ascontiguousarray(region, np.float32).reshape(-1)
    boxes = np.zeros((total, 4), np.float32)
    probs = np.zeros((total, l.classes), np.float32)
    obj = np.zeros(total, np.float32)
    anchors = np.asarray(l.anchors, np.float32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    k = lib.yolo2cuda_region_detections(vp(region), l.w, l.h, l.n, l.classes, vp(anchors), im_w, im_h, net.w, net.h,
                                        C.c_float(thresh), C.c_float(nms), vp(boxes), vp(probs), vp(obj))
    if k < 0:
        raise _capi.Yolo2CudaError(k, "yolo2cuda_region_detections failed")
    return boxes[:k], probs[:k], obj[:k]


def region_detections_gpu(acc: Accelerator, net: _cfg.Network, region, im_w: int, im_h: int, thresh=0.25, nms=0.45):
    """Boxes + per-class NMS on the GPU for a batch (yolo2cuda_region_detections_dev).  region: torch.float32 CUDA tensor
    [batch][...] (what forward_ptr(..., device=True) leaves on the device).  Returns CUDA tensors boxes [batch][total][4],
    probs [batch][total][classes], objectness [batch][total]; POSITIONAL (entry cell*n + anchor), zero rows for rejected cells."""
    import torch
    l = net.layers[-1]
    total = l.w * l.h * l.n
    region = region.contiguous().view(region.shape[0], -1)
    B = region.shape[0]
    boxes = torch.empty((B, total, 4), dtype=torch.float32, device=region.device)
    probs = torch.empty((B, total, l.classes), dtype=torch.float32, device=region.device)
    obj = torch.empty((B, total), dtype=torch.float32, device=region.device)
    anchors = np.asarray(l.anchors, np.float32)
    acc.use_torch_stream(region.device)
    p = lambda t: C.c_void_p(t.data_ptr())
    _capi.check(acc.ctx, acc.lib.yolo2cuda_region_detections_dev(acc.ctx, p(region), B, l.w, l.h, l.n, l.classes,
                                                                 anchors.ctypes.data_as(C.c_void_p), im_w, im_h, net.w, net.h,
                                                                 C.c_float(thresh), C.c_float(nms), p(boxes), p(probs), p(obj)))
    return boxes, probs, obj


def compact_detections_gpu(acc: Accelerator, boxes, probs, obj, cap: int = 256):
    """Positional detect output (CUDA tensors of region_detections_gpu) -> (records int32 [batch][cap][8], counts int32 [batch]) on the
    device: {entry, class, prob bits, x, y, w, h, objectness bits} per surviving (entry, class), ordered; see yolo2cuda.h."""
    import torch
    B, total, classes = probs.shape
    records = torch.zeros((B, cap, 8), dtype=torch.int32, device=probs.device)
    counts = torch.zeros((B,), dtype=torch.int32, device=probs.device)
    acc.use_torch_stream(probs.device)
    p = lambda t: C.c_void_p(t.data_ptr())
    _capi.check(acc.ctx, acc.lib.yolo2cuda_compact_detections_dev(acc.ctx, p(boxes), p(probs), p(obj), B, total, classes, cap, p(records), p(counts)))
    return records, counts


def detections_jsonl(boxes: np.ndarray, probs: np.ndarray, width: int, height: int, labels=None, thresh: float = 0.25,
                     source: str = "", frame_index: int = 0, mode: str = "image") -> str:
    """One JSONL record per inference in the format of the reference's board application
    (linux_app/src/main.c:1028-1075): best class per box, normalised box and pixel corners."""
    import json
    dets = []
    for b, pr in zip(np.asarray(boxes), np.asarray(probs)):
        c = int(np.argmax(pr))
        best = float(pr[c])
        if not best > thresh:
            continue
        x, y, w, h = (np.float32(v) for v in b)            # the pixel corners are float32 arithmetic in the reference
        half, fw, fh = np.float32(0.5), np.float32(width), np.float32(height)
        dets.append({"class_id": c, "label": (labels[c] if labels is not None and c < len(labels) else "unknown"),
                     "prob": round(best, 6),
                     "bbox_norm": {"x": round(float(x), 6), "y": round(float(y), 6), "w": round(float(w), 6), "h": round(float(h), 6)},
                     "bbox_px": {"x0": int((x - w * half) * fw), "y0": int((y - h * half) * fh),
                                 "x1": int((x + w * half) * fw), "y1": int((y + h * half) * fh)}})
    return json.dumps({"mode": mode, "source": source, "frame_index": frame_index, "inference_index": frame_index,
                       "width": width, "height": height, "detections": dets}, separators=(",", ":"))


def yolov2_cuda_ps(net: _cfg.Network, input: np.ndarray, pack: WeightsPack, device: int = 0) -> np.ndarray:
    """Single-frame mirror of yolov2_hls_ps(net, input, precision): returns layers[n-1].output."""
    y = Yolo2Net(net, pack, device=device, max_batch=1)
    try:
        return y.forward(np.asarray(input, np.float32).reshape(1, net.c, net.h, net.w))[0]
    finally:
        y.close()


def _tensor_core_tile(l):
    """csrc/capi.cu auto policy: (channels per work item, pixels per work item) of the tcgen05 kernel, or None for the CUDA-core
    kernel.  128-channel tiles at least 80 % full, or 64-channel tiles x two pixel sets when those fill better (half_mode)."""
    up = lambda n, m: -(-n // m) * m
    fill128, fill64 = l.n * 5 >= up(l.n, 128) * 4, l.n * 5 >= up(l.n, 64) * 4
    deep = (l.size == 3 and l.c >= 32) or (l.size == 1 and l.c >= 128)
    if not deep or not (fill128 or fill64):
        return None
    half = l.n * 5 < up(l.n, 128) * 4 and up(l.n, 64) < up(l.n, 128)
    return (64, 96) if half else (128, 48)


def pass_efficiency(net: _cfg.Network, n: int, sms: int = 148) -> float:
    """Step-weighted fill of the last round / wave of every conv launch for a pass of n frames.

    The persistent tcgen05 kernel (csrc/conv_i16_tc2.cu) walks ceil(n*H*W/48) * ceil(OFM/128) equal work items on `sms` CTAs;
    the CUDA-core kernel (csrc/conv_i16.cu) launches ceil(n*H / (64 // (W/13))) bands x ceil(OFM/16) CTAs, two per SM."""
    import math
    ideal = real = 0.0
    for l in net.layers:
        if l.type != _cfg.CONV:
            continue
        steps = math.ceil(l.c / 4) * l.size * l.size
        tile = _tensor_core_tile(l)
        if tile:
            units, slots, cost = math.ceil(n * l.h * l.w / tile[1]) * math.ceil(l.n / tile[0]), sms, steps * 48 * 128
        else:
            tp = 13 if l.w % 13 == 0 else 7
            sw = -(-l.w // tp)
            rb = max(1, 64 // sw)
            units, slots, cost = math.ceil(n * l.h / rb) * math.ceil(l.n / 16), 2 * sms, steps * rb * l.w * 16 * 1.6
        ideal += units / slots * cost
        real += math.ceil(units / slots) * cost
    return ideal / real


def best_pass_size(net: _cfg.Network, lo: int = 128, hi: int = 400, sms: int = 148) -> int:
    """Frames per device pass in [lo, hi] with the best pass_efficiency (the largest among equals).  For YOLOv2-416 on 148 SMs
    these are the multiples of 21 frames: 21 * 169 pixels = 74 tiles of 48, x 8 channel tiles = 4 * 148 items."""
    return max(range(lo, hi + 1), key=lambda n: (round(pass_efficiency(net, n, sms), 2), n))


def best_ramp_size(net: _cfg.Network, max_batch: int, sms: int = 148) -> int:
    """First (short) pass of a multi-pass host forward: the best-filling size between 32 frames and a quarter of a pass."""
    hi = max(33, max_batch // 4)
    return max(range(32, hi + 1), key=lambda n: (round(pass_efficiency(net, n, sms), 2), -n))
