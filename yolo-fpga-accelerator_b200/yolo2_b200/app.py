"""The whole detection application on the GPU: raw images in, detections out.

Mirrors the per-image flow of the reference's hosts - `yolov2_detect` (load_image_stb, letterbox_image, yolov2_hls_ps,
get_network_boxes, do_nms_sort: src/models/yolov2/yolov2_main.cpp:255-325) and the board application's capture loop with its JSONL
output (linux_app/src/main.c:940-1075) - for BATCHES of same-sized frames:

    u8 frames (pinned host)  --H2D-->  letterbox_kernel  -->  network (tcgen05 / CUDA-core conv, pool, reorg, region)
                                         -->  detect_kernel (boxes + per-class NMS)  --D2H-->  boxes / probabilities  -->  JSONL

Every stage is the bit-exact kernel the parity tests cover; nothing is computed on the host except the JSON text.
"""
import numpy as np

from . import cfg as _cfg
from .accel import letterbox_image
from .model import Yolo2Net, detections_jsonl, region_detections, region_detections_gpu
from .weights import WeightsPack


class DetectionStream:
    """net / pack: as for Yolo2Net.  batch: frames per call (every call may pass fewer)."""

    def __init__(self, net: _cfg.Network, pack: WeightsPack, batch: int, device: int = 0, thresh: float = 0.25, nms: float = 0.45):
        import torch
        self.torch = torch
        self.net, self.batch, self.thresh, self.nms = net, batch, thresh, nms
        self.y = Yolo2Net(net, pack, device=device, max_batch=batch)
        self.dev = torch.device("cuda", device)
        self.region = torch.empty((batch, self.y.region_outputs), dtype=torch.float32, device=self.dev)

    def close(self):
        self.y.close()

    def detect(self, images):
        """images: uint8 [b][h][w][c] (stb layout; numpy, ideally pinned, or a torch CPU/CUDA tensor), b <= batch.
        Returns host numpy arrays boxes [b][total][4], probs [b][total][classes], objectness [b][total] (positional, see
        yolo2cuda_region_detections_dev)."""
        torch = self.torch
        if isinstance(images, np.ndarray):
            images = torch.from_numpy(np.ascontiguousarray(images, dtype=np.uint8))
        b, ih, iw, _ = images.shape
        assert b <= self.batch
        self.y.accel.use_torch_stream(self.dev)
        dev_images = images.to(self.dev, non_blocking=True)
        frames = letterbox_image(self.y.accel, dev_images, self.net.w, self.net.h)
        self.y.forward_ptr(frames.data_ptr(), b, self.region.data_ptr(), device=True)
        l = self.net.layers[-1]
        total = l.w * l.h * l.n
        if total <= 1024:       # detect_kernel sorts a frame's candidates in one CTA's shared memory: at most 1024 (416x416: 845)
            boxes, probs, obj = region_detections_gpu(self.y.accel, self.net, self.region[:b], iw, ih, self.thresh, self.nms)
            return boxes.cpu().numpy(), probs.cpu().numpy(), obj.cpu().numpy()
        # larger grids (608x608: 19*19*5 = 1805 candidates): the region tensors come back and the library's host tail
        # (yolo2cuda_region_detections, the bit-exact default of the C ABI) runs per frame; its compact scan-order list is
        # scattered to the same positional layout the GPU kernel produces (entry = cell * n + anchor)
        self.y.accel.synchronize()
        reg = self.region[:b].cpu().numpy()
        boxes = np.zeros((b, total, 4), np.float32)
        probs = np.zeros((b, total, l.classes), np.float32)
        obj = np.zeros((b, total), np.float32)
        for f in range(b):
            bb, pp, oo = region_detections(self.net, reg[f], iw, ih, self.thresh, self.nms)
            o_map = reg[f].reshape(l.n, l.coords + 1 + l.classes, l.h * l.w)[:, l.coords, :]      # [anchor][cell]
            pos = np.nonzero((o_map.T > np.float32(self.thresh)).reshape(-1))[0]                   # cell-major, then anchor
            assert len(pos) == len(bb)
            boxes[f, pos], probs[f, pos], obj[f, pos] = bb, pp, oo
        return boxes, probs, obj

    def detect_jsonl(self, images, labels=None, source="", first_index=0):
        boxes, probs, _ = self.detect(images)
        h, w = images.shape[1], images.shape[2]
        return [detections_jsonl(boxes[f], probs[f], w, h, labels, self.thresh, source, first_index + f, mode="stream")
                for f in range(boxes.shape[0])]
