"""GPU parity tests: the CUDA path, called through the C ABI, against the oracle on the same
seeded inputs.  Bar: bit-exact for int16; 1e-4 relative (of the layer's max |value|) for fp32."""
import numpy as np
import pytest

from helpers import accel_call, align8, make_conv_case, oracle_conv, valid
from yolo2_b200 import cfg as ycfg, weights as yw
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
        region = y.forward(frames)
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
