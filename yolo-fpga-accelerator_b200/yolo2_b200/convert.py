"""darknet `.weights` -> the accelerator's weight files, in-repo.

The reference delegates this step to an external, un-vendored tool (weights/README.md:37-57:
`solomontesema/nn-weight-extractor`, cloned at run time, no pinned version), so there is nothing in
/root/reference to pin bit-parity against: **parity unpinned** for the quantiser's choices (rounding
mode, Q selection).  What IS contract - the file formats (hls/models/yolov2/yolo2_model.cpp:158-227),
the reorganised order (src/models/yolov2/yolov2_weight_gen.cpp:34-68), darknet's file layout and its
inference-time batch-norm (`(x - mean) / (sqrt(var) + .000001f) * scale + bias`) - is followed exactly,
and the tests check the pipeline against an explicit numpy batch-norm network and against the fp32
CUDA path.

Pipeline:  load_darknet_weights -> fold_batchnorm -> make_fp32_pack  (weights_reorg.bin / bias.bin)
           -> calibrate_activation_q (fp32 CUDA path on calibration frames) -> quantize_pack
           (weights_reorg_int16.bin, bias_int16.bin, weight_int16_Q.bin, bias_int16_Q.bin, iofm_Q.bin)
           -> weights.save_reference_files.
"""
import struct
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from . import cfg as _cfg
from .weights import Tm, Tn, WeightsPack, weight_reorg


@dataclass
class DarknetConv:
    """One convolutional layer's parameters as darknet stores them (weights in [n][c][kh][kw] order)."""
    weights: np.ndarray
    biases: np.ndarray
    scales: Optional[np.ndarray] = None
    rolling_mean: Optional[np.ndarray] = None
    rolling_variance: Optional[np.ndarray] = None


def load_darknet_weights(net: _cfg.Network, path: str) -> List[DarknetConv]:
    """Parses a darknet weights file for `net`'s convolutional layers, in cfg order.
    Header: int32 major, minor, revision; `seen` is 8 bytes when major*10+minor >= 2, else 4."""
    with open(path, "rb") as f:
        major, minor, _rev = struct.unpack("<iii", f.read(12))
        f.read(8 if major * 10 + minor >= 2 else 4)
        out = []
        for l in net.conv_layers:
            def rd(n):
                a = np.fromfile(f, dtype="<f4", count=n)
                if a.size != n:
                    raise RuntimeError(f"{path}: truncated (layer with {l.n} filters)")
                return a
            d = DarknetConv(weights=None, biases=rd(l.n))
            if l.batch_normalize:
                d.scales, d.rolling_mean, d.rolling_variance = rd(l.n), rd(l.n), rd(l.n)
            d.weights = rd(l.n * l.c * l.size * l.size).reshape(l.n, l.c, l.size, l.size)
            out.append(d)
    return out


def save_darknet_weights(layers: List[DarknetConv], path: str, major: int = 0, minor: int = 2, seen: int = 0):
    """Inverse of load_darknet_weights (used by the tests to fabricate a darknet file)."""
    with open(path, "wb") as f:
        f.write(struct.pack("<iii", major, minor, 0))
        f.write(struct.pack("<q" if major * 10 + minor >= 2 else "<i", seen))
        for d in layers:
            d.biases.astype("<f4").tofile(f)
            if d.scales is not None:
                d.scales.astype("<f4").tofile(f)
                d.rolling_mean.astype("<f4").tofile(f)
                d.rolling_variance.astype("<f4").tofile(f)
            d.weights.astype("<f4").tofile(f)


def fold_batchnorm(layers: List[DarknetConv]):
    """Folds darknet's inference batch-norm into each convolution: returns [(w', b')] in float32 with
    w' = w * scale / (sqrt(var) + 1e-6), b' = bias - mean * scale / (sqrt(var) + 1e-6)."""
    out = []
    for d in layers:
        w = d.weights.astype(np.float32)
        b = d.biases.astype(np.float32)
        if d.scales is not None:
            g = (d.scales.astype(np.float64) / (np.sqrt(d.rolling_variance.astype(np.float64)) + 1e-6))
            w = (w.astype(np.float64) * g[:, None, None, None]).astype(np.float32)
            b = (b.astype(np.float64) - d.rolling_mean.astype(np.float64) * g).astype(np.float32)
        out.append((w, b))
    return out


def make_fp32_pack(net: _cfg.Network, folded) -> WeightsPack:
    """fp32 pack in the accelerator's reorganised order (what yolov2_weight_gen writes to weights_reorg.bin)."""
    ws, bs = [], []
    for l, (w, b) in zip(net.conv_layers, folded):
        ws.append(weight_reorg(w.reshape(l.n, l.c, l.size * l.size), l.c, l.n, l.size, min(l.n, Tm), min(l.c, Tn)))
        bs.append(b)
    return WeightsPack(np.concatenate(ws).astype(np.float32), np.concatenate(bs).astype(np.float32))


def best_q(max_abs: float, bits: int = 16, q_max: int = 15) -> int:
    """Largest Q in [0, q_max] with max_abs * 2^Q <= 2^(bits-1) - 1 (all-zero tensors get q_max)."""
    if not np.isfinite(max_abs) or max_abs <= 0:
        return q_max
    q = int(np.floor(np.log2((2 ** (bits - 1) - 1) / max_abs)))
    return max(0, min(q_max, q))


def harmonise_route_q(net: _cfg.Network, act_q: np.ndarray) -> np.ndarray:
    """Makes an iofm_Q table consistent with the way the driver loop reads it (yolo2_model.cpp:311-336,379-399; the same
    bookkeeping is in csrc/capi.cu net_load_weights).  The loop takes a conv's INPUT Q from the table entry of the conv that
    precedes it IN THE FILE, not from the tensor it really reads, and it rescales only the reorg branch of the concat:

      1. single-input route (`route -9`): the conv after it reads the route SOURCE, but is given the output Q of the conv just
         before the route.  Both entries must therefore be equal -> both become their minimum.
      2. concat of (reorg branch, skip conv): the reorg branch is shifted DOWN to min(Q_skip, Q_branch); the skip half is never
         touched, so Q_skip <= Q_branch must hold -> Q_skip is lowered if needed (and rule 1 re-applied, the skip conv being the
         one before the route in YOLOv2).

    Entry k+1 of the table is the output Q of conv k.  Lowering a Q only costs precision, never range."""
    q = np.array(act_q, np.int32, copy=True)
    conv_no = {}                       # layer index -> conv index
    for i, l in enumerate(net.layers):
        if l.type == _cfg.CONV:
            conv_no[i] = len(conv_no)

    def producer_conv(i):              # the conv whose output Q a tensor carries (pool / reorg / single routes keep the Q)
        while net.layers[i].type != _cfg.CONV:
            l = net.layers[i]
            i = l.inputs[0] if l.type == _cfg.ROUTE else i - 1
            if i < 0:
                return None
        return conv_no[i]

    def next_conv(i):
        for j in range(i + 1, len(net.layers)):
            if net.layers[j].type == _cfg.CONV:
                return conv_no[j]
        return None

    ties, clamps = [], []
    for i, l in enumerate(net.layers):
        if l.type != _cfg.ROUTE:
            continue
        if len(l.inputs) == 1:
            src, nxt = producer_conv(l.inputs[0]), next_conv(i)
            if src is not None and nxt is not None and nxt > 0:
                ties.append((src + 1, nxt))          # entry nxt = output Q of conv nxt-1 = what the driver hands conv nxt as Qa_in
        else:
            reorgs = [a for a in l.inputs if net.layers[a].type == _cfg.REORG]
            for a in reorgs:
                for b in l.inputs:
                    if b != a:
                        skip, branch = producer_conv(b), producer_conv(a)
                        if skip is not None and branch is not None:
                            clamps.append((skip + 1, branch + 1))
    for _ in range(len(ties) + len(clamps) + 1):     # to a fixed point (two rules, a handful of entries)
        before = q.copy()
        for a, b in ties:
            q[a] = q[b] = min(q[a], q[b])
        for skip, branch in clamps:
            q[skip] = min(q[skip], q[branch])
        if np.array_equal(before, q):
            break
    return q


def calibrate_activation_q(net: _cfg.Network, fp32_pack: WeightsPack, frames: np.ndarray, device: int = 0, headroom: int = 1,
                           harmonise: bool = True) -> np.ndarray:
    """iofm_Q table (n_conv + 1 entries: network input, then every conv layer's output) from the largest
    magnitude the fp32 CUDA path produces on the calibration frames (float32 [B][c][h][w] in [0,1]).
    `headroom` bits are kept free on the conv outputs: the datapath's accumulator is 16 bits wide in the OUTPUT's
    Q format and saturates after every 4-MAC step (core_compute.cpp:108-118), so partial sums need room too."""
    from .model import Yolo2Net
    y = Yolo2Net(net, fp32_pack, device=device, max_batch=1)
    try:
        y.set_debug_keep(True)
        conv_idx = [i for i, l in enumerate(net.layers) if l.type == _cfg.CONV]
        amax = np.zeros(len(conv_idx) + 1)
        for f in range(frames.shape[0]):
            amax[0] = max(amax[0], float(np.abs(frames[f]).max()))
            y.forward(frames[f:f + 1])
            for j, i in enumerate(conv_idx):
                l = net.layers[i]
                amax[j + 1] = max(amax[j + 1], float(np.abs(y.layer_output(i)[:, :, :l.out_w]).max()))
    finally:
        y.close()
    q = np.array([best_q(a) for a in amax], np.int32)
    q[1:] = np.maximum(q[1:] - headroom, 0)
    return harmonise_route_q(net, q) if harmonise else q


def quantize_pack(net: _cfg.Network, folded, act_q: np.ndarray) -> WeightsPack:
    """int16 pack: per-layer Qw / Qb = the largest Q that does not overflow, values rounded to nearest
    (ties to even, numpy rint) and clipped to int16."""
    convs = net.conv_layers
    ws, bs, wq, bq = [], [], [], []
    for l, (w, b) in zip(convs, folded):
        qw, qb = best_q(float(np.abs(w).max())), best_q(float(np.abs(b).max()))
        wi = np.clip(np.rint(w.astype(np.float64) * 2.0 ** qw), -32768, 32767).astype(np.int16)
        bi = np.clip(np.rint(b.astype(np.float64) * 2.0 ** qb), -32768, 32767).astype(np.int16)
        ws.append(weight_reorg(wi.reshape(l.n, l.c, l.size * l.size), l.c, l.n, l.size, min(l.n, Tm), min(l.c, Tn)))
        bs.append(bi)
        wq.append(qw)
        bq.append(qb)
    return WeightsPack(np.concatenate(ws), np.concatenate(bs), np.array(wq, np.int32), np.array(bq, np.int32),
                       np.asarray(act_q, np.int32))


def convert_darknet(net: _cfg.Network, darknet_path: str, calibration_frames: np.ndarray, out_dir: str, device: int = 0):
    """The whole pipeline; writes both the fp32 and the int16 file sets the reference's loader expects."""
    from .weights import save_reference_files
    folded = fold_batchnorm(load_darknet_weights(net, darknet_path))
    fp32 = make_fp32_pack(net, folded)
    save_reference_files(fp32, net, out_dir)
    act_q = calibrate_activation_q(net, fp32, calibration_frames, device)
    i16 = quantize_pack(net, folded, act_q)
    save_reference_files(i16, net, out_dir)
    return fp32, i16
