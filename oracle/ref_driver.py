"""TEST INFRASTRUCTURE ONLY - the reference's layer loop around the UNMODIFIED YOLO2_FPGA.

yolov2_hls_ps (hls/models/yolov2/yolo2_model.cpp:229-449) is hard-wired to COCO 416x416
(SURVEY.md finding 2), so for VOC / 608 / batched configs the real per-layer entry
(oracle/_ref/libref_*.so -> YOLO2_FPGA) is driven by this generalised restatement of the loop:
dims from the cfg, weights pre-loaded (no file I/O in the timed region), per-layer ofm kept.
Used (a) to pin oracle/yolo2_oracle.c's net forward against the real reference at sizes the
reference driver cannot run, and (b) as bench.py's CPU baseline / `--impl reference` arm.
"""
import time

import numpy as np

from .oracle import CONV, MAXPOOL, REGION, REORG, ROUTE, Oracle, align8

_ARG_ORDER = ("IFM_num", "OFM_num", "Ksize", "Kstride", "Input_w", "Input_h", "Output_w", "Output_h", "Padding", "IsNL",
              "IsBN", "TM", "TN", "TR", "TC", "OFM_num_bound", "mLoopsxTM", "mLoops_a1xTM", "LayerType")


def _skip_layer(net):
    for l in net.layers:
        if l.type == ROUTE and len(l.inputs) >= 2:
            for a in l.inputs:
                if net.layers[a].type == REORG:
                    return [b for b in l.inputs if b != a][0]
    return -1


def ref_net_forward(ref, net, frame, pack, keep_layers=False, helper: Oracle = None):
    """One frame through the real YOLO2_FPGA per conv/pool layer. Returns (region, dumps, seconds in
    YOLO2_FPGA + host ops, excluding setup)."""
    from yolo2_b200.accel import conv_call_args, pool_call_args  # argument recipe of yolo2_model.cpp:299-355
    helper = helper or Oracle()
    tn, tm = getattr(ref, "tn", 4), getattr(ref, "tm", 32)        # the tile parameters this reference build was compiled with
    i16 = pack.is_int16
    dt = np.int16 if i16 else np.float32
    slack = 8192
    outs, dumps = {}, {}
    woff = boff = ci = 0
    if i16:
        aq = pack.act_q
        current_qa, route_q, pending = int(aq[0]), 0, -1
        x0 = helper.quantize_input(frame.reshape(-1), int(aq[0])).reshape(net.c, net.h, net.w)
    else:
        x0 = np.asarray(frame, np.float32).reshape(net.c, net.h, net.w)
    cur = np.zeros((net.c, net.h, align8(net.w)), dt)
    cur[:, :, :net.w] = x0
    skip = _skip_layer(net)
    region = None
    t0 = time.perf_counter()
    for i, l in enumerate(net.layers):
        if l.type == CONV:
            a = conv_call_args(l.c, l.n, l.size, l.stride, l.w, l.h, l.pad, l.leaky, l.batch_normalize, tn=tn, tm=tm)
            q = (0, 0, 0, 0)
            if i16:
                qa_in = int(aq[ci]) if ci < len(aq) else current_qa
                qa_out = int(aq[ci + 1]) if ci + 1 < len(aq) else qa_in
                if pending >= 0:
                    qa_in = pending
                q = (int(pack.weight_q[ci]), qa_in, qa_out, int(pack.bias_q[ci]))
                current_qa = qa_out
                if i == skip:
                    route_q = current_qa
                pending = -1
            nw = l.c * l.n * l.size * l.size
            buf = np.zeros(cur.size + 2 * slack, dt)
            buf[slack:slack + cur.size] = cur.reshape(-1)
            wbuf = np.zeros(nw + 4096, dt)
            wbuf[:nw] = pack.weights[woff:woff + nw]
            out = np.zeros((l.out_c, l.out_h, align8(l.out_w)), dt)
            ref.yolo2_fpga(buf[slack:], out, wbuf, np.ascontiguousarray(pack.bias[boff:boff + l.n]),
                           *[a[k] for k in _ARG_ORDER], *q)
            woff += nw
            boff += l.n
            ci += 1
        elif l.type == MAXPOOL:
            a = pool_call_args(l.c, l.size, l.stride, l.w, l.h, l.out_w, l.out_h, l.pad, tn=tn, tm=tm)
            buf = np.zeros(cur.size + 2 * slack, dt)
            buf[slack:slack + cur.size] = cur.reshape(-1)
            out = np.zeros((l.out_c, l.out_h, align8(l.out_w)), dt)
            ref.yolo2_fpga(buf[slack:], out, None, None, *[a[k] for k in _ARG_ORDER], 0, 0, 0, 0)
        elif l.type == REORG:
            shift = 0
            if i16 and route_q > 0:
                target = min(route_q, current_qa)
                shift = current_qa - target
                if shift:
                    current_qa = target
                pending = current_qa
            out = helper.reorg_driver(cur, l.c, l.h, l.w, shift)
        elif l.type == ROUTE:
            out = np.concatenate([outs[s] for s in l.inputs], axis=0)
        elif l.type == REGION:
            region = helper.region_from_ofm(cur, l.w, l.h, l.n, l.classes, l.coords, l.softmax, l.background,
                                            current_qa if i16 else 0)
            out = None
        outs[i] = out
        if keep_layers and out is not None and l.type in (CONV, MAXPOOL, REORG):
            dumps[i] = out
        if out is not None:
            cur = out
    return region, dumps, time.perf_counter() - t0


def _worker(args):
    precision, cfg_text, seed, table, frame_seed, nframes = args
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "yolo-fpga-accelerator_b200"))
    from yolo2_b200 import cfg as ycfg, weights as yw
    from .oracle import Ref
    net = ycfg.parse_network_cfg(cfg_text)
    pack = yw.synth_pack(net, precision, seed=seed, table=table)
    ref = Ref(precision)
    helper = Oracle()
    frames = yw.synth_frames(net, nframes, seed=frame_seed)
    secs = 0.0
    for f in range(nframes):
        _, _, s = ref_net_forward(ref, net, frames[f], pack, helper=helper)
        secs += s
    return secs


def time_reference_cpu(cfg_text, precision="int16", procs=1, frames_per_proc=1, seed=0, table="default"):
    """Runs `procs` independent processes (the reference is single-threaded and non-reentrant:
    function-local statics, yolo2_accel.cpp:103-113), each pushing `frames_per_proc` frames through
    the real YOLO2_FPGA. Returns (frames/s aggregate, seconds per frame per core, wall seconds)."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    jobs = [(precision, cfg_text, seed, table, 1000 + 17 * p, frames_per_proc) for p in range(procs)]
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        secs = pool.map(_worker, jobs)
    wall = time.perf_counter() - t0
    worst = max(secs)
    return procs * frames_per_proc / worst, sum(secs) / (procs * frames_per_proc), wall


def _frame_worker(job):
    """one frame (per-layer ofm kept) through the real YOLO2_FPGA, or through the pinned C restatement when oracle/_ref was not built"""
    cfg_text, precision, pack_seed, table, frame_seed, index, use_ref, keep = job
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "yolo-fpga-accelerator_b200"))
    from yolo2_b200 import cfg as ycfg, weights as yw
    from .oracle import Ref
    net = ycfg.parse_network_cfg(cfg_text)
    pack = yw.synth_pack(net, precision, seed=pack_seed, table=table)
    frame = np.random.default_rng(frame_seed + index).random((net.c, net.h, net.w), dtype=np.float32)   # == synth_frames(...)[index]
    if use_ref:
        region, dumps, _ = ref_net_forward(Ref(precision), net, frame, pack, keep_layers=keep, helper=Oracle())
    else:
        region, dumps = Oracle().net_forward(net, frame, pack, dump_layers=keep)
    return index, np.asarray(region, np.float32).reshape(-1), dumps


def reference_frames(cfg_text, precision, pack_seed, table, frame_seed, indices, keep_layers=True, procs=0):
    """Frames `indices` of yolo2_b200.weights.synth_frames(net, *, frame_seed) through the checker, one process per frame.
    Returns ({index: (region flat float32, {layer: ofm})}, "reference" | "port")."""
    import multiprocessing as mp
    import os
    from .oracle import have_ref
    use_ref = have_ref(precision)
    jobs = [(cfg_text, precision, pack_seed, table, frame_seed, int(i), use_ref, keep_layers) for i in indices]
    n = max(1, min(len(jobs), procs or (os.cpu_count() or 1)))
    if not use_ref:
        os.environ.setdefault("OMP_NUM_THREADS", str(max(1, (os.cpu_count() or 1) // n)))
    with mp.get_context("spawn").Pool(n) as pool:
        res = pool.map(_frame_worker, jobs)
    return {i: (region, dumps) for i, region, dumps in res}, ("reference" if use_ref else "port")
