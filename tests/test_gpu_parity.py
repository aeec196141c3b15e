"""GPU parity tests: the CUDA path, called through the C ABI, against the oracle on the same
seeded inputs.  Bar: bit-exact for int16; 1e-4 relative (of the layer's max |value|) for fp32."""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import accel_call, align8, make_conv_case, oracle_conv, valid
from yolo2_b200 import _capi, cfg as ycfg, weights as yw
from yolo2_b200.accel import pool_call_args
from yolo2_b200.model import Yolo2Net

pytestmark = pytest.mark.gpu

CONV_CASES = [
    # c, n, size, stride, w, h, leaky, (Qw, Qa_in, Qa_out, Qb), expected kernel prefix
    (3, 32, 3, 1, 26, 26, 1, (14, 10, 10, 10), "conv_i16_c4<13,3,scaled"),
    (64, 32, 3, 1, 13, 13, 1, (14, 10, 10, 10), "conv_i16_c4<13,3,scaled"),
    (32, 64, 3, 1, 52, 39, 1, (14, 10, 10, 10), "conv_i16_c4<13,3,scaled"),
    (128, 64, 1, 1, 26, 26, 1, (14, 10, 10, 10), "conv_i16_c4<13,1,scaled"),
    (16, 425, 1, 1, 13, 13, 0, (12, 12, 7, 8), "conv_i16_c4<13,1,scaled"),
    (20, 40, 3, 1, 19, 19, 1, (13, 9, 12, 11), "conv_i16_c4<7,3,scaled"),
    (7, 9, 3, 1, 33, 5, 0, (14, 10, 10, 10), "conv_i16_c4"),
    (40, 24, 1, 1, 19, 7, 1, (15, 10, 10, 2), "conv_i16_c4<7,1,scaled"),
    (24, 32, 3, 1, 26, 13, 1, (15, 12, 2, 10), "conv_i16_c4<13,3,unscaled"),   # shift_out = 25 > 22
    (24, 20, 1, 1, 14, 9, 0, (15, 15, 0, 3), "conv_i16_c4<7,1,unscaled"),      # shift_out = 30
    (16, 16, 3, 1, 13, 13, 1, (15, 15, 8, 0), "conv_i16_c4<13,3,scaled"),      # shift_out = 22 (largest scaled)
    (16, 16, 3, 1, 13, 13, 1, (4, 10, 6, 12), "conv_i16_c4<13,3,scaled"),      # shift_out = 8 (smallest), bias shift left
    (32, 70, 3, 2, 27, 19, 0, (14, 10, 9, 12), "conv_i16_generic"),      # stride 2
    (8, 8, 3, 1, 9, 9, 1, (3, 3, 10, 12), "conv_i16_generic"),           # negative shift_out
    (8, 16, 3, 1, 13, 13, 1, (10, 5, 10, 10), "conv_i16_generic"),       # shift_out = 5 < 8
    (12, 16, 2, 1, 14, 14, 1, (14, 10, 10, 10), "conv_i16_generic"),     # Ksize 2
]


@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: "c%d_n%d_k%d_s%d_%dx%d" % c[:6])
def test_conv_int16_bit_exact(case, accel16, oracle):
    c, n, size, stride, w, h, leaky, q, kern = case
    pad = size // 2 if size != 2 else 0
    a, x, wr, b, _ = make_conv_case(hash(case[:6]) & 0xffff, c, n, size, stride, w, h, leaky, pad=pad)
    want = oracle_conv(oracle, a, x, wr, b, q)
    got = accel_call(accel16, a, x, wr, b, q)
    assert accel16.last_kernel.startswith(kern), accel16.last_kernel
    ow = a["Output_w"]
    assert np.array_equal(valid(got, ow), valid(want, ow))


@pytest.mark.parametrize("amp,q", [(32767, (13, 9, 12, 7)), (32767, (15, 15, 0, 31)), (32767, (8, 0, 0, 0)),
                                   (20000, (14, 10, 2, 0))])
def test_conv_int16_saturation_and_extremes(amp, q, accel16, oracle):
    """Full-range operands: |P| reaches 2^32 in the reference's int64; heavy saturation."""
    a, x, wr, b, _ = make_conv_case(7, 17, 33, 3, 1, 20, 11, 1, amp=amp, xamp=32767)
    x[:, :, :20][::2] = 32767
    x[:, :, :20][1::2] = -32768
    want = oracle_conv(oracle, a, x, wr, b, q)
    got = accel_call(accel16, a, x, wr, b, q)
    assert np.array_equal(valid(got, 20), valid(want, 20))
    if q[2] == 12:
        assert (np.abs(valid(want, 20).astype(int)) >= 32767).mean() > 0.05      # the case really saturates


def test_conv_pad_columns_untouched(accel16):
    a, x, wr, b, _ = make_conv_case(3, 8, 8, 3, 1, 13, 13, 1)
    out = np.full((8, 13, 16), -777, np.int16)
    accel_call(accel16, a, x, wr, b, (14, 10, 10, 10), out=out)
    assert (out[:, :, 13:] == -777).all()          # core_compute.cpp:218-219 writes TC_MIN valid columns only


@pytest.mark.parametrize("c,w,h,stride", [(32, 26, 26, 2), (5, 13, 13, 2), (8, 14, 10, 1), (64, 52, 52, 2)])
def test_maxpool_int16(c, w, h, stride, accel16, oracle):
    rng = np.random.default_rng(c * w)
    pad = 1
    ow, oh = (w + pad - 2) // stride + 1, (h + pad - 2) // stride + 1
    a = pool_call_args(c, 2, stride, w, h, ow, oh, pad)
    x = np.full((c, h, align8(w)), 32000, np.int16)
    x[:, :, :w] = rng.integers(-32768, 32768, (c, h, w))
    want = oracle.maxpool(x, c, 2, stride, w, h, ow, oh)
    got = accel_call(accel16, a, x, None, None)
    assert np.array_equal(valid(got, ow), valid(want, ow))


def test_bad_arguments_return_error(accel16):
    from yolo2_b200 import Yolo2CudaError
    a, x, wr, b, _ = make_conv_case(3, 8, 8, 3, 1, 13, 13, 1)
    for key, bad in [("IFM_num", 4096), ("Ksize", 5), ("Kstride", 3), ("TM", 64), ("TN", 9), ("Padding", 7),
                     ("OFM_num_bound", 8), ("LayerType", 3)]:
        a2 = dict(a)
        a2[key] = bad
        with pytest.raises(Yolo2CudaError) as e:
            accel_call(accel16, a2, x, wr, b)
        assert e.value.code == -1


@pytest.mark.parametrize("c,n,size,w,h", [(3, 32, 3, 26, 26), (64, 48, 3, 13, 13), (96, 40, 1, 19, 19), (16, 16, 3, 21, 9)])
def test_conv_fp32_tolerance(c, n, size, w, h, accel32, oracle):
    a, x, wr, b, _ = make_conv_case(c + n, c, n, size, 1, w, h, 1, dtype=np.float32, poison=1e30)
    want = oracle_conv(oracle, a, x, wr, b)
    got = accel_call(accel32, a, x, wr, b)
    ref_max = np.abs(valid(want, w)).max()
    assert np.abs(valid(got, w) - valid(want, w)).max() <= 1e-4 * ref_max   # BASELINE.json: 1e-4 relative


def _net_case(width, height, classes, channel_div, table, seed, precision="int16"):
    net = ycfg.parse_network_cfg(ycfg.yolov2_cfg_text(width, height, classes, channel_div=channel_div))
    pack = yw.synth_pack(net, precision, seed=seed, table=table)
    return net, pack


def _check_net(net, pack, frames, oracle, max_batch, tol=None):
    y = Yolo2Net(net, pack, max_batch=max_batch)
    try:
        region_compact = y.forward(frames)              # default: compact arena, buffers recycled down the network
        compact_bytes = y.activation_bytes
        y.set_debug_keep(True)                          # per-layer dumps need every tensor kept
        region = y.forward(frames)
        assert np.array_equal(region.view(np.uint32), region_compact.view(np.uint32)), "compact arena and keep-all mode differ"
        assert compact_bytes < y.activation_bytes
        B = frames.shape[0]
        for f in sorted({0, B - 1}):
            if B > max_batch and f == 0:
                continue  # per-layer dumps are only kept for the last chunk
            want_region, dumps = oracle.net_forward(net, frames[f], pack, dump_layers=True)
            fl = f % max_batch if B > max_batch else f
            for i, want in dumps.items():
                got = y.layer_output(i, fl)
                ow = net.layers[i].out_w
                if tol is None:
                    assert np.array_equal(valid(got, ow), valid(want, ow)), f"frame {f} layer {i} differs"
                else:
                    m = np.abs(valid(want, ow)).max()
                    assert np.abs(valid(got, ow) - valid(want, ow)).max() <= tol * max(m, 1e-6), f"frame {f} layer {i}"
            if tol is None:
                assert np.array_equal(region[f].view(np.uint32), want_region.view(np.uint32)), f"frame {f} region differs"
            else:
                assert np.abs(region[f] - want_region).max() <= 1e-4
        return region, y.launches_per_forward
    finally:
        y.close()


@pytest.mark.parametrize("table", ["default", "stress", "saturate"])
def test_thin_net_416_every_layer_bit_exact(table, oracle):
    """The 32-section YOLOv2 topology at 416x416 (hidden channels / 8): every conv/pool/reorg ofm and
    the region tensor bit-exact, for the default, stress (mixed shift signs, route Q-align) and
    saturating Q tables."""
    net, pack = _net_case(416, 416, 3, 8, table, seed=11)
    frames = yw.synth_frames(net, 3, seed=2000)
    _check_net(net, pack, frames, oracle, max_batch=2)      # 3 frames through chunks of 2


def test_thin_net_608_voc_head(oracle):
    """608x608 (widths 19..608: the 7-pixel segment kernels) with a 20-class head."""
    net, pack = _net_case(608, 608, 20, 8, "stress", seed=5)
    frames = yw.synth_frames(net, 2, seed=3000)
    _check_net(net, pack, frames, oracle, max_batch=2)


def test_thin_net_fp32(oracle):
    net, pack = _net_case(416, 416, 3, 8, "default", seed=3, precision="fp32")
    frames = yw.synth_frames(net, 2, seed=4000)
    _check_net(net, pack, frames, oracle, max_batch=2, tol=1e-4)


@pytest.mark.slow
def test_full_yolov2_416_coco_bit_exact(oracle):
    """BASELINE config #1 shape: full-width YOLOv2 COCO 416, one frame, every layer + region + boxes."""
    net, pack = _net_case(416, 416, 80, 1, "default", seed=1)
    frames = yw.synth_frames(net, 1, seed=1000)
    region, launches = _check_net(net, pack, frames, oracle, max_batch=1)
    assert launches >= 30
    from yolo2_b200.model import region_detections
    l = net.layers[-1]
    thresh = float(np.sort(region[0][:, 4].reshape(-1))[-40])      # ~40 candidates whatever the synthetic scores
    b, p, o = region_detections(net, region[0], 768, 576, thresh, 0.45)
    wb, wp, wo = oracle.region_boxes_nms(region[0], l.w, l.h, l.n, l.classes, l.anchors, 768, 576, net.w, net.h, thresh, 0.45)
    live = wo > 0
    key = lambda bb, pp: sorted((tuple(x.tolist()), tuple(np.nonzero(q)[0].tolist()), tuple(q[q > 0].tolist())) for x, q in zip(bb, pp))
    assert key(b, p) == key(wb[live], wp[live])


def test_batch_linearity_and_idempotence(oracle):
    """Size-independent properties at a batch the oracle cannot afford: frames are independent, so
    the region tensor of frame i does not depend on its neighbours or on the chunking, and a second
    forward returns identical bits."""
    net, pack = _net_case(416, 416, 3, 8, "default", seed=21)
    frames = yw.synth_frames(net, 16, seed=5000)
    y = Yolo2Net(net, pack, max_batch=8)
    try:
        r1 = y.forward(frames)
        r2 = y.forward(frames)
        assert np.array_equal(r1.view(np.uint32), r2.view(np.uint32))
        perm = np.arange(16)[::-1].copy()
        r3 = y.forward(frames[perm])
        assert np.array_equal(r3[perm].view(np.uint32), r1.view(np.uint32))
        dup = np.repeat(frames[:1], 5, axis=0)
        r4 = y.forward(dup)
        assert all(np.array_equal(r4[i].view(np.uint32), r1[0].view(np.uint32)) for i in range(5))
    finally:
        y.close()
    want, _ = oracle.net_forward(net, frames[7], pack)
    assert np.array_equal(r1[7].view(np.uint32), want.view(np.uint32))


# ---- golden fixtures (outputs of the real reference) through the CUDA path -------------------------

def _golden():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "layer_cases.npz"))


def _gargs(arr):
    from oracle.gen_golden import ARG_KEYS
    return dict(zip(ARG_KEYS, [int(v) for v in arr]))


@pytest.mark.parametrize("i", range(9))
def test_cuda_conv_int16_vs_reference_golden(i, accel16):
    g = _golden()
    a = _gargs(g[f"ci16_{i}_args"])
    got = accel_call(accel16, a, g[f"ci16_{i}_x"], g[f"ci16_{i}_w"], g[f"ci16_{i}_b"], [int(v) for v in g[f"ci16_{i}_q"]])
    ow = a["Output_w"]
    assert np.array_equal(valid(got, ow), valid(g[f"ci16_{i}_out"], ow))


@pytest.mark.parametrize("i", range(4))
def test_cuda_conv_fp32_vs_reference_golden(i, accel32):
    g = _golden()
    a = _gargs(g[f"cf32_{i}_args"])
    got = accel_call(accel32, a, g[f"cf32_{i}_x"], g[f"cf32_{i}_w"], g[f"cf32_{i}_b"])
    ow = a["Output_w"]
    want = valid(g[f"cf32_{i}_out"], ow)
    assert np.abs(valid(got, ow) - want).max() <= 1e-4 * np.abs(want).max()


@pytest.mark.parametrize("i", range(3))
def test_cuda_maxpool_vs_reference_golden(i, accel16, accel32):
    g = _golden()
    a = _gargs(g[f"pool_{i}_args"])
    ow = a["Output_w"]
    assert np.array_equal(valid(accel_call(accel16, a, g[f"pool_{i}_x"], None, None), ow), valid(g[f"pool_{i}_out"], ow))
    assert np.array_equal(valid(accel_call(accel32, a, g[f"poolf_{i}_x"], None, None), ow), valid(g[f"poolf_{i}_out"], ow))


@pytest.mark.parametrize("tag,table,seed", [("default", "default", 1), ("stress", "stress", 2)])
def test_cuda_full_coco416_vs_reference_golden(tag, table, seed):
    """--backend cuda vs the UNMODIFIED yolov2_hls_ps: region tensor of layers[31].output, bit for bit."""
    import hashlib
    import os
    full = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "yolov2_coco416_full.npz"))
    net = ycfg.parse_network_cfg(ycfg.yolov2_cfg_text(416, 416, 80))
    pack = yw.synth_pack(net, "int16", seed=seed, table=table)
    frame = yw.synth_frames(net, 1, seed=1000)
    if hashlib.sha256(pack.weights.tobytes()).digest() != full[f"{tag}_weights_sha"].tobytes():
        pytest.skip("numpy RNG stream differs from the one the golden was generated with")
    from yolo2_b200.model import yolov2_cuda_ps
    region = yolov2_cuda_ps(net, frame[0], pack)
    assert np.array_equal(region.reshape(-1).view(np.uint32), full[f"{tag}_region"].view(np.uint32))


# ---- LayerType 2 and the driver-side operators on device memory -------------------------------------

@pytest.mark.parametrize("ch,ow,oh,TM", [(4, 13, 13, 4), (8, 26, 20, 4), (4, 5, 5, 2), (6, 7, 9, 4)])
def test_reorg_layertype2(ch, ow, oh, TM, accel16, oracle):
    rng = np.random.default_rng(ch * ow)
    iw, ih = 2 * ow, 2 * oh
    mLoops = -(-ch // TM)
    a = dict(IFM_num=ch, OFM_num=ch, Ksize=2, Kstride=2, Input_w=iw, Input_h=ih, Output_w=ow, Output_h=oh, Padding=0,
             IsNL=0, IsBN=0, TM=TM, TN=0, TR=min(13, oh), TC=min(13, ow), OFM_num_bound=(mLoops + 2) * TM,
             mLoopsxTM=mLoops * TM, mLoops_a1xTM=(mLoops + 1) * TM, LayerType=2)
    x = np.zeros((ch, ih, align8(iw)), np.int16)
    x[:, :, :iw] = rng.integers(-30000, 30000, (ch, ih, iw))
    want = oracle.reorg_hls(x, ch, TM, iw, ih, ow, oh)
    got = accel_call(accel16, a, x, None, None)
    # tile channels beyond the four phases are never written by the reference: compare written ones
    written = [m + q for m in range(0, ch, TM) for q in range(min(TM, ch - m, 4))]
    assert np.array_equal(valid(got, ow)[written], valid(want, ow)[written])


def test_driver_side_operators_on_device(accel16, oracle):
    import ctypes as C
    import torch
    lib, ctx = accel16.lib, accel16.ctx
    rng = np.random.default_rng(0)
    # input quantiser
    x = (rng.random(3 * 32 * 32, dtype=np.float32) * 80 - 40).astype(np.float32)
    x[:6] = [0.5, -0.5, 1.5 / 1024, 2.5 / 1024, -2.5 / 1024, 0.0]
    dx = torch.from_numpy(x).cuda()
    dq = torch.empty(x.size, dtype=torch.int16, device="cuda")
    assert lib.yolo2cuda_quantize_input_dev(ctx, C.c_void_p(dx.data_ptr()), C.c_void_p(dq.data_ptr()), x.size, 10) == 0
    accel16.synchronize()
    assert np.array_equal(dq.cpu().numpy(), oracle.quantize_input(x, 10))
    # flat-memory reorg + Q-align shift (26x26x64 is the reference's own case; 38x38x16 a 608-net one)
    for c, h, w, shift in [(64, 26, 26, 0), (64, 26, 26, 3), (16, 38, 38, 1)]:
        t = np.zeros((c, h, align8(w)), np.int16)
        t[:, :, :w] = rng.integers(-32768, 32768, (c, h, w))
        want = oracle.reorg_driver(t, c, h, w, shift)
        din = torch.from_numpy(t).cuda()
        dout = torch.full(want.shape, 99, dtype=torch.int16, device="cuda")
        assert lib.yolo2cuda_reorg_dev(ctx, C.c_void_p(din.data_ptr()), C.c_void_p(dout.data_ptr()), c, h, w, shift) == 0
        accel16.synchronize()
        assert np.array_equal(dout.cpu().numpy(), want)        # including the zeroed pad columns (yolo2_model.cpp:375)
    # region head
    for classes, n, w in [(80, 5, 13), (20, 5, 19)]:
        ch = n * (5 + classes)
        t = np.zeros((ch, w, align8(w)), np.int16)
        t[:, :, :w] = rng.integers(-6000, 6000, (ch, w, w))
        want = oracle.region_from_ofm(t, w, w, n, classes, 4, 1, 0, q=9)
        din = torch.from_numpy(t).cuda()
        dout = torch.empty(want.size, dtype=torch.float32, device="cuda")
        assert lib.yolo2cuda_region_dev(ctx, C.c_void_p(din.data_ptr()), C.c_void_p(dout.data_ptr()), w, w, n, classes, 4, 1, 0, 9) == 0
        accel16.synchronize()
        assert np.array_equal(dout.cpu().numpy().view(np.uint32), want.reshape(-1).view(np.uint32))
    assert lib.yolo2cuda_reorg_dev(ctx, C.c_void_p(din.data_ptr()), C.c_void_p(dout.data_ptr()), 3, 5, 5, 0) == -1


def test_layer_dev_entry_matches_host_entry(accel16, oracle):
    """yolo2cuda_layer_dev on torch device tensors == the host entry == the oracle."""
    import torch
    a, x, wr, b, _ = make_conv_case(5, 32, 48, 3, 1, 26, 26, 1)
    q = (14, 10, 10, 10)
    want = oracle_conv(oracle, a, x, wr, b, q)
    dx, dw, db = (torch.from_numpy(t).cuda() for t in (x, wr, b))
    dout = torch.zeros(want.shape, dtype=torch.int16, device="cuda")
    accel16.YOLO2_FPGA_dev(dx.data_ptr(), dout.data_ptr(), dw.data_ptr(), db.data_ptr(), *[a[k] for k in (
        "IFM_num", "OFM_num", "Ksize", "Kstride", "Input_w", "Input_h", "Output_w", "Output_h", "Padding", "IsNL",
        "IsBN", "TM", "TN", "TR", "TC", "OFM_num_bound", "mLoopsxTM", "mLoops_a1xTM", "LayerType")], *q)
    accel16.synchronize()
    assert np.array_equal(valid(dout.cpu().numpy(), 26), valid(want, 26))


# ---- tensor-core (tcgen05 + TMEM) conv path (csrc/conv_i16_tc2.cu): YOLO2CUDA_TC=2 forces it wherever the shape is eligible ------

@pytest.fixture()
def accel16_tc(monkeypatch):
    from yolo2_b200.accel import Accelerator
    monkeypatch.setenv("YOLO2CUDA_TC", "2")
    a = Accelerator(0, "int16")
    yield a
    a.close()


@pytest.mark.parametrize("c,n,k,w,h,q,amp", [
    (4, 128, 1, 8, 8, (14, 10, 10, 10), 600), (64, 128, 3, 13, 13, (14, 10, 10, 10), 600),
    (64, 200, 3, 26, 26, (13, 9, 12, 7), 600), (256, 256, 3, 13, 13, (15, 12, 8, 10), 600),
    (96, 40, 1, 19, 19, (12, 12, 7, 8), 600), (17, 33, 3, 20, 11, (13, 9, 12, 7), 32767),
    (20, 130, 3, 21, 9, (4, 10, 6, 12), 32767), (36, 64, 3, 13, 13, (15, 15, 8, 0), 3000)])
def test_tensor_core_conv_bit_exact(c, n, k, w, h, q, amp, accel16_tc, oracle):
    """tcgen05.mma kind::i8 on hi/lo byte planes with a block-diagonal activation operand (7 chain steps per
    K=32 slice), recombined per step on the CUDA cores: same bits as the reference, incl. saturation."""
    a, x, wr, b, _ = make_conv_case(c * n + k, c, n, k, 1, w, h, 1, amp=amp, xamp=32767 if amp > 600 else 2000)
    want = oracle_conv(oracle, a, x, wr, b, q)
    got = accel_call(accel16_tc, a, x, wr, b, q)
    assert accel16_tc.last_kernel.startswith("conv_i16_tc2<")
    assert np.array_equal(valid(got, w), valid(want, w))


@pytest.mark.parametrize("so", list(range(8, 23)))
def test_tensor_core_v2_every_shift(so, monkeypatch, oracle):
    """csrc/conv_i16_tc2.cu instantiates one kernel per accumulator shift (the step's LEA.HI takes an immediate) and switches
    formula at so = 17 and the rounding-constant plane at so = 16: every instantiation, full-range operands (saturation)."""
    from yolo2_b200.accel import Accelerator
    monkeypatch.setenv("YOLO2CUDA_TC", "2")
    acc = Accelerator(0, "int16")
    try:
        q = (so - 2, 10, 8, 9)                      # Qw, Qa_in, Qa_out, Qb -> shift_out = Qa_in + Qw - Qa_out = so
        for (c, n, k, w, h) in ((24, 130, 3, 13, 13), (60, 128, 1, 7, 5)):
            a, x, wr, b, _ = make_conv_case(so * 100 + c, c, n, k, 1, w, h, 1, amp=32767, xamp=32767)
            want = oracle_conv(oracle, a, x, wr, b, q)
            got = accel_call(acc, a, x, wr, b, q)
            assert acc.last_kernel.startswith("conv_i16_tc2<")
            assert np.array_equal(valid(got, w), valid(want, w))
    finally:
        acc.close()


@pytest.mark.parametrize("so", list(range(8, 17)))
def test_tensor_core_fast_path_every_shift(so, monkeypatch, oracle):
    """moderate operands: most (warp, tile) units pass the no-saturation test and take the two-instruction step (HH from the dense
    column, 65536 HH a multiple of 2^so), some fail it and take the exact step in the same launch; every shift instantiation"""
    from yolo2_b200.accel import Accelerator
    monkeypatch.setenv("YOLO2CUDA_TC", "2")
    acc = Accelerator(0, "int16")
    try:
        q = (so - 2, 10, 8, 9)
        # the single-call entry knows nothing about the caller's tensor, so the bound assumes |x| <= 32768: Dsum = sum|w| * 2^(15-so).
        # Small weights (sum over 28 of them < 2^so / 4) keep it below the int16 range for every shift; the activations stay large.
        amp = max(2, (1 << so) // 512)
        for (c, n, k, w, h) in ((96, 130, 3, 13, 13), (200, 128, 1, 26, 5)):
            a, x, wr, b, _ = make_conv_case(so * 10 + c, c, n, k, 1, w, h, 1, amp=amp, xamp=12000)
            want = oracle_conv(oracle, a, x, wr, b, q)
            acc.tc_path_counts(reset=True)
            got = accel_call(acc, a, x, wr, b, q)
            fast, exact = acc.tc_path_counts()
            assert acc.last_kernel.startswith("conv_i16_tc2<")
            assert np.array_equal(valid(got, w), valid(want, w))
            assert fast > 0, (so, fast, exact)
    finally:
        acc.close()


@pytest.mark.parametrize("xamp,expect_mixed", [(300, False), (9000, True), (20000, True)])
def test_tensor_core_fast_and_exact_paths_mix(xamp, expect_mixed, monkeypatch, oracle):
    """deep chains (2304 steps) whose accumulators wander towards the int16 limits: the per-K-block range test sends the risky
    tiles to the exact step and the rest to the fast one; the result is the reference's bits whatever the mix.  With
    YOLO2CUDA_TC_EXACT=1 the fast path is off and the bits are the same."""
    from yolo2_b200.accel import Accelerator
    a, x, wr, b, _ = make_conv_case(4242 + xamp, 1024, 128, 3, 1, 13, 13, 1, amp=600, xamp=xamp)
    q = (14, 10, 10, 10)
    want = oracle_conv(oracle, a, x, wr, b, q)
    for force_exact in ("0", "1"):
        monkeypatch.setenv("YOLO2CUDA_TC", "2")
        monkeypatch.setenv("YOLO2CUDA_TC_EXACT", force_exact)
        acc = Accelerator(0, "int16")
        try:
            got = accel_call(acc, a, x, wr, b, q)
            fast, exact = acc.tc_path_counts()
            assert np.array_equal(valid(got, 13), valid(want, 13)), (xamp, force_exact)
            if force_exact == "1":
                assert fast == 0 and exact > 0
            else:
                assert fast > 0 and (exact > 0) == expect_mixed, (xamp, fast, exact)
        finally:
            acc.close()
    if xamp == 20000:
        assert (np.abs(valid(want, 13).astype(int)) >= 32767).any()        # the case really saturates somewhere


def test_tensor_core_net_fast_path_statistics(monkeypatch, oracle):
    """inside the network executor the bound uses the producing layer's tracked maximum |activation|: on the default table nearly
    every tile of a thin net takes the fast path, on the saturating table the exact path runs; both bit-exact (checked against the
    oracle in test_tensor_core_net_bit_exact) and identical to a run with the fast path disabled"""
    from yolo2_b200.accel import Accelerator
    regions = {}
    for table in ("default", "saturate"):
        for force_exact in ("0", "1"):
            monkeypatch.setenv("YOLO2CUDA_TC", "2")
            monkeypatch.setenv("YOLO2CUDA_TC_MIN_OFM", "8")
            monkeypatch.setenv("YOLO2CUDA_TC_EXACT", force_exact)
            net, pack = _net_case(416, 416, 3, 8, table, seed=11)
            frames = yw.synth_frames(net, 3, seed=2000)
            y = Yolo2Net(net, pack, max_batch=3)
            try:
                regions[(table, force_exact)] = y.forward(frames).copy()
                fast, exact = y.accel.tc_path_counts()
            finally:
                y.close()
            if force_exact == "1":
                assert fast == 0 and exact > 0
            elif table == "default":
                assert fast > 20 * max(exact, 1), (fast, exact)
            else:
                assert exact > 0
        assert np.array_equal(regions[(table, "0")].view(np.uint32), regions[(table, "1")].view(np.uint32))


@pytest.mark.parametrize("table", ["stress", "saturate", "default"])
def test_tensor_core_net_bit_exact(table, monkeypatch, oracle):
    monkeypatch.setenv("YOLO2CUDA_TC", "2")
    monkeypatch.setenv("YOLO2CUDA_TC_MIN_OFM", "8")
    net, pack = _net_case(416, 416, 3, 8, table, seed=11)
    frames = yw.synth_frames(net, 3, seed=2000)
    _check_net(net, pack, frames, oracle, max_batch=2)


# ---- image front-end on the GPU ---------------------------------------------------------------------

@pytest.mark.parametrize("i", range(7))
def test_letterbox_golden_gpu(i, accel16):
    """letterbox_kernel (csrc/bw_ops.cu) against the reference's letterbox_image output frozen in tests/golden/letterbox.npz."""
    import torch
    from yolo2_b200.accel import letterbox_image
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "letterbox.npz"))
    nw, nh = (int(v) for v in g[f"lb_{i}_net"])
    img = torch.from_numpy(g[f"lb_{i}_img"][None].copy()).cuda()
    got = letterbox_image(accel16, img, nw, nh).cpu().numpy()[0]
    assert np.array_equal(got.view(np.uint32), g[f"lb_{i}_out"].view(np.uint32))


@pytest.mark.parametrize("w,h", [(640, 480), (333, 500), (1280, 720)])
def test_letterbox_batch_vs_oracle(w, h, accel16, oracle):
    """a batch of video-sized frames: every float of the 416x416 network input identical to the oracle's"""
    import torch
    from yolo2_b200.accel import letterbox_image
    imgs = np.random.default_rng(w + h).integers(0, 256, (3, h, w, 3), dtype=np.uint8)
    got = letterbox_image(accel16, torch.from_numpy(imgs).cuda(), 416, 416).cpu().numpy()
    for f in range(3):
        assert np.array_equal(got[f].view(np.uint32), oracle.letterbox_u8(imgs[f], 416, 416).view(np.uint32))


def test_forward_images_equals_forward_of_letterboxed(oracle):
    """yolo2cuda_net_forward_images_host == letterbox (oracle) + yolo2cuda_net_forward_host, bit for bit, across passes"""
    net, pack = _net_case(416, 416, 3, 8, "default", seed=5)
    imgs = np.random.default_rng(99).integers(0, 256, (3, 120, 160, 3), dtype=np.uint8)
    y = Yolo2Net(net, pack, max_batch=2)
    try:
        frames = np.stack([oracle.letterbox_u8(im, 416, 416) for im in imgs])
        want = y.forward(frames)
        got = y.forward_images(imgs)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    finally:
        y.close()


# ---- rounding-group variants: yolo2cuda_set_tile_params emulates a reference built with another Tn ----

@pytest.mark.parametrize("tn", [8, 16, 32])
def test_tile_param_variants_bit_exact(tn, oracle):
    from yolo2_b200.accel import Accelerator
    acc = Accelerator(0, "int16")
    try:
        acc.set_tile_params(tn, 32)
        for i, (c, n, k, w, h, q, amp) in enumerate([(64, 40, 3, 13, 13, (14, 10, 10, 10), 600), (37, 33, 3, 20, 11, (13, 9, 12, 7), 32767),
                                                     (96, 64, 1, 19, 19, (12, 12, 7, 8), 3000), (3, 16, 3, 26, 26, (14, 10, 10, 10), 600)]):
            a, x, wr, b, _ = make_conv_case(900 + 10 * tn + i, c, n, k, 1, w, h, 1, amp=amp, xamp=32767 if amp > 600 else 2000, tn=tn)
            want = oracle_conv(oracle, a, x, wr, b, q)
            got = accel_call(acc, a, x, wr, b, q)
            assert np.array_equal(valid(got, w), valid(want, w)), (tn, c, n, k, acc.last_kernel)
        # a TN the emulated build does not have is the reference's assert (yolo2_accel.cpp:75-87)
        a, x, wr, b, _ = make_conv_case(1, 64, 32, 3, 1, 13, 13, 1, tn=2 * tn)
        with pytest.raises(Exception):
            accel_call(acc, a, x, wr, b, (14, 10, 10, 10))
    finally:
        acc.close()


@pytest.mark.parametrize("c,n,k,w,h,q,amp", [
    (64, 128, 3, 13, 13, (14, 10, 10, 10), 600), (32, 64, 3, 52, 39, (14, 10, 10, 10), 600), (256, 200, 3, 26, 26, (13, 9, 12, 7), 600),
    (96, 40, 1, 19, 19, (10, 10, 7, 8), 3000), (37, 33, 3, 20, 11, (13, 9, 12, 7), 32767), (16, 130, 3, 21, 9, (4, 10, 6, 12), 32767),
    (512, 256, 1, 26, 26, (15, 10, 9, 11), 600), (1024, 128, 3, 13, 13, (14, 10, 10, 10), 600)])
def test_tn32_tensor_core_conv_bit_exact(c, n, k, w, h, q, amp, oracle):
    """csrc/conv_i16_tc32.cu: with the reference built as Tn = 32 one int8 MMA K slice is one rounding group; same bits as the
    oracle with TN = 32 (itself pinned against that reference build in tests/test_oracle_vs_ref.py), incl. saturation, ragged
    channel counts (c % 32 != 0, c < 32), partial pixel tiles and output-channel tiles."""
    from yolo2_b200.accel import Accelerator
    acc = Accelerator(0, "int16")
    try:
        acc.set_tile_params(32, 32)
        a, x, wr, b, _ = make_conv_case(c * n + k, c, n, k, 1, w, h, 1, amp=amp, xamp=32767 if amp > 600 else 2000, tn=32)
        want = oracle_conv(oracle, a, x, wr, b, q)
        got = accel_call(acc, a, x, wr, b, q)
        assert acc.last_kernel.startswith("conv_i16_tc32<"), acc.last_kernel
        assert np.array_equal(valid(got, w), valid(want, w))
    finally:
        acc.close()


@pytest.mark.parametrize("tn", [8, 16])
@pytest.mark.parametrize("c,n,k,w,h,q,amp", [
    (64, 128, 3, 13, 13, (14, 10, 10, 10), 600), (32, 64, 3, 52, 39, (14, 10, 10, 10), 600), (256, 200, 3, 26, 26, (13, 9, 12, 7), 600),
    (96, 40, 1, 19, 19, (10, 10, 7, 8), 3000), (37, 33, 3, 20, 11, (13, 9, 12, 7), 32767), (16, 130, 3, 21, 9, (4, 10, 6, 12), 32767),
    (512, 256, 1, 26, 26, (15, 10, 9, 11), 600), (1024, 128, 3, 13, 13, (14, 10, 10, 10), 600), (72, 24, 1, 13, 13, (14, 10, 10, 10), 32767),
    (200, 136, 3, 104, 7, (14, 10, 10, 10), 600)])
def test_tn8_tn16_tensor_core_conv_bit_exact(tn, c, n, k, w, h, q, amp, oracle, monkeypatch):
    """csrc/conv_i16_tc32.cu with TNW = 16 / 8: one MMA K slice holds two / four consecutive rounding steps as a block-diagonal
    activation operand.  Same bits as the oracle with TN = tn (pinned against that reference build in tests/test_oracle_vs_ref.py):
    saturation, ragged channel counts (last group short, slices that straddle two channel groups and two staging chunks, step slots
    past the last step), partial pixel tiles and output-channel tiles, both kernel sizes."""
    monkeypatch.setenv("YOLO2CUDA_TC", "2")
    from yolo2_b200.accel import Accelerator
    acc = Accelerator(0, "int16")
    try:
        acc.set_tile_params(tn, 32)
        a, x, wr, b, _ = make_conv_case(c * n + k + tn, c, n, k, 1, w, h, 1, amp=amp, xamp=32767 if amp > 600 else 2000, tn=tn)
        want = oracle_conv(oracle, a, x, wr, b, q)
        got = accel_call(acc, a, x, wr, b, q)
        if 8 <= q[0] + q[1] - q[2] <= 16:
            assert acc.last_kernel.startswith("conv_i16_tc32<") and acc.last_kernel.endswith(f",tn{tn}>"), acc.last_kernel
        assert np.array_equal(valid(got, w), valid(want, w))
    finally:
        acc.close()


@pytest.mark.parametrize("tn,tc", [(8, ""), (16, ""), (32, ""), (8, "2"), (16, "2")])
def test_whole_net_rounding_group_variant(tn, tc, oracle, monkeypatch):
    """a whole (thin) YOLOv2 through the network executor emulating a reference built with --tn 8 / 16 (CUDA-core kernel with 2 / 4
    C4 words per rounding step, or - forced here with YOLO2CUDA_TC=2, by policy on deep layers - the tensor-core kernel with 4 / 2
    steps per K slice) and --tn 32 (tensor-core kernel), every layer's ofm and the region tensor bit-exact to the oracle with the
    same tile parameters"""
    if tc:
        monkeypatch.setenv("YOLO2CUDA_TC", tc)
    from yolo2_b200.accel import Accelerator
    net = ycfg.parse_network_cfg(ycfg.yolov2_cfg_text(416, 416, 3, channel_div=8))
    pack = yw.synth_pack(net, "int16", seed=21, table="default", tn=tn)
    frames = yw.synth_frames(net, 2, seed=4000)
    acc = Accelerator(0, "int16")
    acc.set_tile_params(tn, 32)
    oracle.set_tile_params(tn, 32)
    y = Yolo2Net(net, pack, max_batch=2, accel=acc)
    y.set_debug_keep(True)
    try:
        region = y.forward(frames)
        for f in (0, 1):
            want_region, dumps = oracle.net_forward(net, frames[f], pack, dump_layers=True)
            for i, want in dumps.items():
                got = y.layer_output(i, f)
                assert np.array_equal(valid(got, net.layers[i].out_w), valid(want, net.layers[i].out_w)), (tn, i)
            assert np.array_equal(region[f].view(np.uint32), want_region.view(np.uint32))
        kernels = {y.layer_kernel(i) for i in range(len(net.layers))}
        assert any(k.endswith({8: "tn8>", 16: "tn16>", 32: ">"}[tn]) and
                   k.startswith("conv_i16_tc32<" if (tc or tn == 32) else "conv_i16_c4<") for k in kernels), kernels
    finally:
        oracle.set_tile_params(4, 32)
        y.close()
        acc.close()


@pytest.mark.parametrize("i", range(5))
def test_tn_variant_golden_gpu(i):
    """the CUDA path (grouped CUDA-core kernel for Tn = 8 / 16 in this single-frame entry, tcgen05 kernel for Tn = 32) against outputs of
    the reference built with that Tn"""
    from oracle.gen_golden import ARG_KEYS
    from yolo2_b200.accel import Accelerator
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "layer_cases_tn_variants.npz"))
    a = dict(zip(ARG_KEYS, [int(v) for v in g[f"v_{i}_args"]]))
    tn = int(g[f"v_{i}_tn"][0])
    acc = Accelerator(0, "int16")
    try:
        acc.set_tile_params(tn, 32)
        got = accel_call(acc, a, g[f"v_{i}_x"], g[f"v_{i}_w"], g[f"v_{i}_b"], [int(v) for v in g[f"v_{i}_q"]])
        assert np.array_equal(valid(got, a["Output_w"]), valid(g[f"v_{i}_out"], a["Output_w"])), acc.last_kernel
        if tn == 32:
            assert acc.last_kernel.startswith("conv_i16_tc32<")
    finally:
        acc.close()


# ---- boxes + per-class NMS on the GPU --------------------------------------------------------------------------

def _det_set(boxes, probs):
    """order-free view of a detection result: sorted rows (class, prob bits, box bits) of every non-zero probability"""
    rows = []
    for b, p in zip(boxes, probs):
        for k in np.nonzero(p)[0]:
            rows.append((int(k), int(np.float32(p[k]).view(np.uint32)), *[int(v) for v in np.asarray(b, np.float32).view(np.uint32)]))
    return sorted(rows)


@pytest.mark.parametrize("im_w,im_h,thresh", [(640, 480, 0.25), (333, 500, 0.1), (416, 416, 0.24)])
def test_gpu_detections_equal_reference_set(im_w, im_h, thresh, accel16, oracle):
    """detect_kernel (csrc/bw_ops.cu) against the oracle's get_region_detections + correct_region_boxes + do_nms_sort: the same
    surviving (class, probability, box) tuples, bit for bit (incl. the box w/h through the restated glibc expf), for a batch."""
    import torch
    from yolo2_b200.model import region_detections_gpu
    net = ycfg.parse_network_cfg(ycfg.yolov2_cfg_text(416, 416, 80))
    l = net.layers[-1]
    rng = np.random.default_rng(im_w + im_h)
    B = 3
    raw = rng.normal(0, 1.5, (B, l.n, 5 + l.classes, l.h, l.w)).astype(np.float32)
    raw[:, :, 4] = rng.normal(-2.0, 2.0, (B, l.n, l.h, l.w))          # a few dozen cells above the objectness threshold
    raw[:, :, 5] += 3.0                                                  # one popular class: real suppression work in its NMS
    region = np.stack([oracle.region_forward(r.reshape(-1), l.w, l.h, l.n, l.classes) for r in raw])
    gb, gp, go = region_detections_gpu(accel16, net, torch.from_numpy(region).cuda(), im_w, im_h, thresh, 0.45)
    gb, gp, go = gb.cpu().numpy(), gp.cpu().numpy(), go.cpu().numpy()
    kept = 0
    for f in range(B):
        wb, wp, wo = oracle.region_boxes_nms(region[f], l.w, l.h, l.n, l.classes, l.anchors, im_w, im_h, net.w, net.h, thresh, 0.45)
        want, got = _det_set(wb, wp), _det_set(gb[f], gp[f])
        assert got == want
        assert sorted(np.float32(wo[wo != 0]).view(np.uint32).tolist()) == sorted(go[f][go[f] != 0].view(np.uint32).tolist())
        kept += len(want)
    assert kept > 20        # the case is not vacuous


def test_detection_stream_end_to_end(oracle):
    """yolo2_b200.app.DetectionStream (u8 frames -> letterbox -> network -> boxes + NMS, all on the GPU) against the same chain through
    the oracle: letterbox_u8, net_forward, region_boxes_nms - identical detection sets, and well-formed JSONL records."""
    import json
    from yolo2_b200.app import DetectionStream
    net, pack = _net_case(416, 416, 3, 8, "default", seed=5)
    imgs = np.random.default_rng(123).integers(0, 256, (3, 120, 160, 3), dtype=np.uint8)
    ds = DetectionStream(net, pack, batch=4, thresh=0.05, nms=0.45)
    try:
        boxes, probs, obj = ds.detect(imgs)
        lines = ds.detect_jsonl(imgs, labels=["a", "b", "c"], source="test")
    finally:
        ds.close()
    l = net.layers[-1]
    for f in range(3):
        region, _ = oracle.net_forward(net, oracle.letterbox_u8(imgs[f], 416, 416), pack)
        wb, wp, wo = oracle.region_boxes_nms(region, l.w, l.h, l.n, l.classes, l.anchors, 160, 120, net.w, net.h, 0.05, 0.45)
        assert _det_set(boxes[f], probs[f]) == _det_set(wb, wp)
        rec = json.loads(lines[f])
        assert rec["width"] == 160 and rec["height"] == 120 and rec["frame_index"] == f


def test_detection_stream_608_uses_host_tail(oracle):
    """a 608x608 net has 19*19*5 = 1805 candidates per frame, more than detect_kernel sorts in one CTA: DetectionStream then
    runs the library's host tail per frame and returns the same positional layout; detection set == the checker's"""
    from yolo2_b200.app import DetectionStream
    net, pack = _net_case(608, 608, 3, 8, "default", seed=6)
    imgs = np.random.default_rng(124).integers(0, 256, (2, 90, 160, 3), dtype=np.uint8)
    ds = DetectionStream(net, pack, batch=2, thresh=0.05, nms=0.45)
    try:
        boxes, probs, obj = ds.detect(imgs)
    finally:
        ds.close()
    l = net.layers[-1]
    assert boxes.shape == (2, 19 * 19 * 5, 4)
    for f in range(2):
        region, _ = oracle.net_forward(net, oracle.letterbox_u8(imgs[f], 608, 608), pack)
        wb, wp, wo = oracle.region_boxes_nms(region, l.w, l.h, l.n, l.classes, l.anchors, 160, 90, net.w, net.h, 0.05, 0.45)
        assert _det_set(boxes[f], probs[f]) == _det_set(wb, wp)


@pytest.mark.parametrize("q", [0, 7, 10, 15])
def test_quantizer_boundaries_bit_exact(q, accel16, oracle):
    """the quantiser without the 64-bit llroundf (trunc + exact tie test): every half-way value k + 0.5 and its two binary32
    neighbours over the whole int16 range, the values just below 0.5, saturation, infinities, NaN, denormals, random floats"""
    import ctypes as C
    import torch
    k = np.arange(-33000, 33001, dtype=np.float64) + 0.5
    cand = np.concatenate([k, k - 0.5, np.array([0.49999997, -0.49999997, 0.5, -0.5, 1e30, -1e30, np.inf, -np.inf, np.nan, 1e-45, -1e-45, 0.0, -0.0])])
    v = (cand * 2.0 ** -q).astype(np.float32)
    v = np.concatenate([v, np.nextafter(v, np.float32(np.inf)), np.nextafter(v, np.float32(-np.inf)),
                        np.random.default_rng(q).normal(0, 3000 * 2.0 ** -q, 200000).astype(np.float32)])
    v = np.ascontiguousarray(np.resize(v, (v.size + 3) // 4 * 4))
    want = oracle.quantize_input(v, q)
    dx = torch.from_numpy(v).cuda()
    dq = torch.empty(v.size, dtype=torch.int16, device="cuda")
    assert accel16.lib.yolo2cuda_quantize_input_dev(accel16.ctx, C.c_void_p(dx.data_ptr()), C.c_void_p(dq.data_ptr()), v.size, q) == 0
    accel16.synchronize()
    got = dq.cpu().numpy()
    bad = np.nonzero(got != want)[0]
    assert bad.size == 0, (v[bad[:5]], got[bad[:5]], want[bad[:5]])


def test_gpu_detections_tie_order(accel16, oracle):
    """equal class probabilities + overlapping boxes: detect_kernel must resolve the ties like do_nms_sort's carried-over stable
    sort (src/core/yolo_post.cpp:70-74), i.e. by the probabilities of the previously processed classes, then scan order"""
    import torch
    from test_host_logic import _tied_region
    from yolo2_b200.model import region_detections_gpu
    net = ycfg.parse_network_cfg(ycfg.yolov2_cfg_text(416, 416, 6))
    l = net.layers[-1]
    region = np.stack([_tied_region(s) for s in range(4)])
    gb, gp, go = region_detections_gpu(accel16, net, torch.from_numpy(region).cuda(), 640, 480, 0.1, 0.45)
    gb, gp = gb.cpu().numpy(), gp.cpu().numpy()
    for f in range(4):
        wb, wp, wo = oracle.region_boxes_nms(region[f], l.w, l.h, l.n, l.classes, l.anchors, 640, 480, net.w, net.h, 0.1, 0.45)
        assert _det_set(gb[f], gp[f]) == _det_set(wb, wp), f


def test_compact_detections_records(accel16, oracle):
    """compact_detections_kernel: the surviving (entry, class) pairs of detect_kernel's positional output as ordered fixed-size
    records (what the multi-GPU bench gathers instead of region tensors), against numpy"""
    import torch
    from test_host_logic import _tied_region
    from yolo2_b200.model import compact_detections_gpu, region_detections_gpu
    net = ycfg.parse_network_cfg(ycfg.yolov2_cfg_text(416, 416, 6))
    region = np.stack([_tied_region(s) for s in range(3)])
    gb, gp, go = region_detections_gpu(accel16, net, torch.from_numpy(region).cuda(), 640, 480, 0.1, 0.45)
    for cap in (4096, 50):
        rec, cnt = compact_detections_gpu(accel16, gb, gp, go, cap=cap)
        rec, cnt = rec.cpu().numpy(), cnt.cpu().numpy()
        b, p, o = gb.cpu().numpy(), gp.cpu().numpy(), go.cpu().numpy()
        for f in range(3):
            e, c = np.nonzero(p[f] > 0)
            assert cnt[f] == len(e) > 100
            k = min(cap, len(e))
            want = np.zeros((k, 8), np.uint32)
            want[:, 0], want[:, 1] = e[:k], c[:k]
            want[:, 2] = p[f][e[:k], c[:k]].view(np.uint32)
            want[:, 3:7] = b[f][e[:k]].view(np.uint32)
            want[:, 7] = o[f][e[:k]].view(np.uint32)
            assert np.array_equal(rec[f, :k].view(np.uint32), want)


def test_host_forward_pass_schedule_is_invisible(oracle):
    """yolo2cuda_net_forward_host splits a batch larger than max_batch into a short ramp pass + full passes
    (yolo2cuda_net_set_ramp_frames); the region tensors must not depend on the schedule."""
    net, pack = _net_case(416, 416, 3, 8, "default", seed=5)
    frames = yw.synth_frames(net, 11, seed=77)
    y = Yolo2Net(net, pack, max_batch=4)
    try:
        outs = []
        for ramp in (-1, 0, 1, 3, 4, 7):
            y.set_ramp_frames(ramp)
            outs.append(y.forward(frames).copy())
        for o in outs[1:]:
            assert np.array_equal(o.view(np.uint32), outs[0].view(np.uint32))
        want = oracle.net_forward(net, frames[10], pack)[0]
        assert np.array_equal(np.ascontiguousarray(outs[0][10]).reshape(-1).view(np.uint32), np.asarray(want, np.float32).reshape(-1).view(np.uint32))
    finally:
        y.close()


def test_device_exp_equals_host_libm(accel16, oracle):
    """glibc_exp in csrc/bw_ops.cu (the exponential of the region head's logistic and softmax) against the host libm's exp on 2^21
    inputs - the region head's range, float-valued arguments, |x| up to 1100 (scaled special cases, over- and underflow) and
    arbitrary bit patterns - bit for bit (NaN payloads aside)."""
    import torch
    rng = np.random.default_rng(7)
    n = 1 << 19
    x = np.concatenate([rng.uniform(-40, 40, n), rng.uniform(-90, 90, n).astype(np.float32).astype(np.float64),
                        rng.uniform(-1100, 1100, n), rng.integers(0, 2 ** 64, n, dtype=np.uint64).view(np.float64),
                        np.array([0.0, -0.0, np.inf, -np.inf, 709.78, -745.2, 1e-300, -1e-300, 512.0, -512.0, 1024.0, -1075.0])])
    xd = torch.from_numpy(x).cuda()
    yd = torch.empty_like(xd)
    _capi.check(accel16.ctx, accel16.lib.yolo2cuda_selftest_exp_dev(accel16.ctx, C.c_void_p(xd.data_ptr()), C.c_void_p(yd.data_ptr()), x.size))
    accel16.synchronize()
    got, want = yd.cpu().numpy(), oracle.libm_exp(x)
    both_nan = np.isnan(got) & np.isnan(want)
    assert np.array_equal(got.view(np.uint64)[~both_nan], want.view(np.uint64)[~both_nan])
