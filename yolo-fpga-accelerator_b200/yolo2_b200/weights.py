"""Weight / bias / Q-table files of the accelerator path, and seeded synthetic fixtures.

File contracts follow the reference loader load_weights (hls/models/yolov2/yolo2_model.cpp:158-227)
and the offline reorganiser WeightReorg (src/models/yolov2/yolov2_weight_gen.cpp:34-68).
The reference ships no weights (weights/.gitignore there), so tests and benchmarks synthesise
them from fixed seeds.
"""
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import cfg as _cfg

Tm, Tn = 32, 4  # hls/core/params.hpp defaults (scripts/hw_params_gen.py:16-23)


@dataclass
class WeightsPack:
    """WeightsPack of yolo2_model.cpp:150-156: blobs hold conv layers back to back, no pad element."""
    weights: np.ndarray            # reorganised order, int16 or float32
    bias: np.ndarray
    weight_q: Optional[np.ndarray] = None   # int32 per conv layer
    bias_q: Optional[np.ndarray] = None
    act_q: Optional[np.ndarray] = None      # int32, entry 0 = input Q, entry i+1 = output Q of conv i

    @property
    def is_int16(self):
        return self.weights.dtype == np.int16


def weight_reorg(w: np.ndarray, ifm: int, ofm: int, ksize: int, tm: int = Tm, tn: int = Tn) -> np.ndarray:
    """darknet order [ofm][ifm][kh][kw] -> accelerator order: for each (32-row m-tile, 4-channel
    n-tile) a contiguous [tap][tm][tn] block (yolov2_weight_gen.cpp:43-66)."""
    k2 = ksize * ksize
    w = np.asarray(w).reshape(ofm, ifm, k2)
    out = np.empty(ofm * ifm * k2, dtype=w.dtype)
    off = 0
    for m in range(0, ofm, tm):
        tmm = min(tm, ofm - m)
        for n in range(0, ifm, tn):
            tnn = min(tn, ifm - n)
            blk = w[m:m + tmm, n:n + tnn, :].transpose(2, 0, 1)  # [tap][tm][tn]
            out[off:off + blk.size] = blk.reshape(-1)
            off += blk.size
    return out


def load_reference_files(net: _cfg.Network, precision: str, directory: str = "weights") -> WeightsPack:
    """Reads the reference's weight files with its size checks and its odd-length pad rule
    (yolo2_model.cpp:171-225)."""
    counts = net.weight_counts()
    exp_w = sum(c[0] for c in counts)
    exp_b = sum(c[1] for c in counts)
    if precision == "fp32":
        w = np.fromfile(os.path.join(directory, "weights_reorg.bin"), dtype=np.float32)
        b = np.fromfile(os.path.join(directory, "bias.bin"), dtype=np.float32)
        if w.size < exp_w:
            raise RuntimeError("weights file too small")
        if b.size < exp_b:
            raise RuntimeError("bias file too small")
        return WeightsPack(w[:exp_w].copy(), b[:exp_b].copy())
    w = np.fromfile(os.path.join(directory, "weights_reorg_int16.bin"), dtype=np.int16)
    b = np.fromfile(os.path.join(directory, "bias_int16.bin"), dtype=np.int16)
    if w.size < exp_w:
        raise RuntimeError("weights file too small")
    if b.size < exp_b:
        raise RuntimeError("bias file too small")
    wq = np.fromfile(os.path.join(directory, "weight_int16_Q.bin"), dtype=np.int32)
    bq = np.fromfile(os.path.join(directory, "bias_int16_Q.bin"), dtype=np.int32)
    if wq.size < len(counts) or bq.size < len(counts):
        raise RuntimeError("Q tables too small for conv layers")
    aq_path = os.path.join(directory, "iofm_Q.bin")
    aq = np.fromfile(aq_path, dtype=np.int32) if os.path.exists(aq_path) else np.zeros(0, np.int32)
    wbuf = np.empty(exp_w, np.int16)
    bbuf = np.empty(exp_b, np.int16)
    wf = wo = bf = bo = 0
    for li, (wl, bl) in enumerate(counts):
        if wf + wl > w.size:
            raise RuntimeError(f"int16 weight truncated at layer {li}")
        if bf + bl > b.size:
            raise RuntimeError(f"int16 bias truncated at layer {li}")
        wbuf[wo:wo + wl] = w[wf:wf + wl]
        bbuf[bo:bo + bl] = b[bf:bf + bl]
        wf += wl + (wl & 1)   # one pad element after odd-length layers (:216-223)
        wo += wl
        bf += bl + (bl & 1)
        bo += bl
    return WeightsPack(wbuf, bbuf, wq, bq, aq)


def save_reference_files(pack: WeightsPack, net: _cfg.Network, directory: str):
    """Writes a pack in the reference's on-disk format (inverse of load_reference_files)."""
    os.makedirs(directory, exist_ok=True)
    if not pack.is_int16:
        pack.weights.astype(np.float32).tofile(os.path.join(directory, "weights_reorg.bin"))
        pack.bias.astype(np.float32).tofile(os.path.join(directory, "bias.bin"))
        return
    ws, bs, wo, bo = [], [], 0, 0
    for wl, bl in net.weight_counts():
        ws.append(pack.weights[wo:wo + wl]); wo += wl
        if wl & 1:
            ws.append(np.zeros(1, np.int16))
        bs.append(pack.bias[bo:bo + bl]); bo += bl
        if bl & 1:
            bs.append(np.zeros(1, np.int16))
    np.concatenate(ws).astype(np.int16).tofile(os.path.join(directory, "weights_reorg_int16.bin"))
    np.concatenate(bs).astype(np.int16).tofile(os.path.join(directory, "bias_int16.bin"))
    pack.weight_q.astype(np.int32).tofile(os.path.join(directory, "weight_int16_Q.bin"))
    pack.bias_q.astype(np.int32).tofile(os.path.join(directory, "bias_int16_Q.bin"))
    pack.act_q.astype(np.int32).tofile(os.path.join(directory, "iofm_Q.bin"))


def synth_pack(net: _cfg.Network, precision: str = "int16", seed: int = 0, table: str = "default",
               w_amp: int = 600, b_amp: int = 2000, tn: int = Tn, tm: int = Tm) -> WeightsPack:
    """Seeded synthetic weights in darknet order, reorganised like yolov2_weight_gen does.
    table: "default" (Qw=14,Qb=10,Qa=10 everywhere), "stress" (per-layer random Qa in [7,12],
    Qw in [12,15], Qb in [8,12]: both shift signs and the route Q-align), "saturate" (full-range
    +-32767 weights with a small shift: heavy saturation, int32-overflow guard)."""
    rng = np.random.default_rng(seed)
    ws, bs = [], []
    convs = net.conv_layers
    for l in convs:
        k2 = l.size * l.size
        if precision == "fp32":
            w = rng.normal(0.0, 0.02, size=(l.n, l.c, k2)).astype(np.float32)
            b = rng.normal(0.0, 0.1, size=l.n).astype(np.float32)
        else:
            amp = 32767 if table == "saturate" else w_amp
            w = rng.integers(-amp, amp + 1, size=(l.n, l.c, k2)).astype(np.int16)
            b = rng.integers(-b_amp, b_amp + 1, size=l.n).astype(np.int16)
        ws.append(weight_reorg(w, l.c, l.n, l.size, min(l.n, tm), min(l.c, tn)))
        bs.append(b)
    W = np.concatenate(ws)
    Bv = np.concatenate(bs)
    if precision == "fp32":
        return WeightsPack(W, Bv)
    n = len(convs)
    if table == "default":
        wq = np.full(n, 14, np.int32); bq = np.full(n, 10, np.int32); aq = np.full(n + 1, 10, np.int32)
    elif table == "stress":
        wq = rng.integers(12, 16, size=n).astype(np.int32)
        bq = rng.integers(8, 13, size=n).astype(np.int32)
        aq = rng.integers(7, 13, size=n + 1).astype(np.int32)
    elif table == "saturate":
        wq = np.full(n, 15, np.int32); bq = np.full(n, 10, np.int32); aq = np.full(n + 1, 10, np.int32)
    else:
        raise ValueError(table)
    return WeightsPack(W, Bv, wq, bq, aq)


def synth_frames(net: _cfg.Network, batch: int, seed: int = 1000) -> np.ndarray:
    """float32 [batch][c][h][w] uniform [0,1), one seed per frame (seed + frame index)."""
    out = np.empty((batch, net.c, net.h, net.w), np.float32)
    for f in range(batch):
        out[f] = np.random.default_rng(seed + f).random((net.c, net.h, net.w), dtype=np.float32)
    return out
