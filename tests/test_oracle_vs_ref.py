"""Pins the C restatement (oracle/yolo2_oracle.c) against the UNMODIFIED reference compiled by
oracle/Makefile (oracle/_ref/libref_*.so).  Skipped where the reference library was not built."""
import numpy as np
import pytest

from helpers import make_conv_case, oracle_conv, valid
from oracle.oracle import align8
from oracle.ref_driver import ref_net_forward
from yolo2_b200 import cfg as ycfg, weights as yw
from yolo2_b200.accel import pool_call_args

RNG = np.random.default_rng(1234)


def _rand_conv_cases(n):
    out = []
    for k in range(n):
        size = int(RNG.choice([1, 2, 3]))
        stride = int(RNG.choice([1, 1, 2]))
        c = int(RNG.integers(1, 40))
        m = int(RNG.integers(1, 70))
        w = int(RNG.integers(size, 40))
        h = int(RNG.integers(size, 30))
        pad = int(RNG.integers(0, 2)) if size > 1 else 0
        qw, qi, qo, qb = int(RNG.integers(8, 16)), int(RNG.integers(4, 14)), int(RNG.integers(4, 14)), int(RNG.integers(4, 14))
        out.append((k, c, m, size, stride, w, h, int(RNG.integers(0, 2)), pad, (qw, qi, qo, qb)))
    return out


@pytest.mark.parametrize("case", _rand_conv_cases(24), ids=lambda c: "r%d_c%d_n%d_k%d_s%d_%dx%d_p%d" % (c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[8]))
def test_conv_int16_random_shapes(case, oracle, ref16):
    k, c, n, size, stride, w, h, leaky, pad, q = case
    amp = 32767 if k % 3 == 0 else 600
    a, x, wr, b, _ = make_conv_case(k, c, n, size, stride, w, h, leaky, amp=amp, xamp=20000 if k % 3 == 0 else 2000, pad=pad)
    if a["Output_w"] <= 0 or a["Output_h"] <= 0:
        pytest.skip("degenerate")
    want = ref16.run_layer(x, wr, b, a, q)
    got = oracle_conv(oracle, a, x, wr, b, q)
    ow = a["Output_w"]
    assert np.array_equal(valid(got, ow), valid(want, ow))


@pytest.mark.parametrize("c,n,size,stride,w,h", [(3, 32, 3, 1, 26, 26), (30, 17, 3, 2, 21, 9), (64, 48, 1, 1, 13, 13), (5, 9, 2, 1, 12, 12)])
def test_conv_fp32_same_bits_as_reference(c, n, size, stride, w, h, oracle, ref32):
    a, x, wr, b, _ = make_conv_case(c * n, c, n, size, stride, w, h, 1, dtype=np.float32, poison=1e30, pad=size // 2 if size != 2 else 0)
    want = ref32.run_layer(x, wr, b, a)
    got = oracle_conv(oracle, a, x, wr, b)
    ow = a["Output_w"]
    assert np.abs(valid(got, ow) - valid(want, ow)).max() <= 1e-6 * np.abs(valid(want, ow)).max()


@pytest.mark.parametrize("c,w,h,stride", [(4, 26, 26, 2), (7, 13, 13, 2), (9, 15, 11, 1), (3, 27, 8, 2)])
def test_maxpool_matches_reference(c, w, h, stride, oracle, ref16, ref32):
    ow, oh = (w + 1 - 2) // stride + 1, (h + 1 - 2) // stride + 1
    a = pool_call_args(c, 2, stride, w, h, ow, oh, 1)
    x = np.full((c, h, align8(w)), 31000, np.int16)
    x[:, :, :w] = RNG.integers(-32768, 32768, (c, h, w))
    assert np.array_equal(valid(oracle.maxpool(x, c, 2, stride, w, h, ow, oh), ow), valid(ref16.run_layer(x, None, None, a), ow))
    xf = x.astype(np.float32) * 50
    assert np.array_equal(valid(oracle.maxpool(xf, c, 2, stride, w, h, ow, oh), ow), valid(ref32.run_layer(xf, None, None, a), ow))


@pytest.mark.parametrize("ch,ow,oh,TM", [(4, 13, 13, 4), (8, 26, 20, 4), (4, 5, 5, 2), (6, 7, 9, 4)])
def test_reorg_layertype2_matches_reference(ch, ow, oh, TM, oracle, ref16):
    iw, ih = 2 * ow, 2 * oh
    mLoops = -(-ch // TM)
    a = dict(IFM_num=ch, OFM_num=ch, Ksize=2, Kstride=2, Input_w=iw, Input_h=ih, Output_w=ow, Output_h=oh, Padding=0,
             IsNL=0, IsBN=0, TM=TM, TN=0, TR=min(13, oh), TC=min(13, ow), OFM_num_bound=(mLoops + 2) * TM,
             mLoopsxTM=mLoops * TM, mLoops_a1xTM=(mLoops + 1) * TM, LayerType=2)
    x = np.zeros((ch, ih, align8(iw)), np.int16)
    x[:, :, :iw] = RNG.integers(-30000, 30000, (ch, ih, iw))
    assert np.array_equal(valid(oracle.reorg_hls(x, ch, TM, iw, ih, ow, oh), ow), valid(ref16.run_layer(x, None, None, a), ow))


def test_region_forward_and_boxes_match_reference(oracle, ref16):
    rf = RNG.normal(0, 2.5, (5 * 85, 13, 13)).astype(np.float32)
    want = ref16.region_forward(rf, 13, 13, 5, 80)
    got = oracle.region_forward(rf, 13, 13, 5, 80)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    anchors = np.array([0.57273, 0.677385, 1.87446, 2.06253, 3.33843, 5.47434, 7.88282, 3.52778, 9.77052, 9.16828], np.float32)
    for im_w, im_h, thresh in [(768, 576, 0.2), (300, 500, 0.05), (416, 416, 0.5)]:
        wb, wp, wo = ref16.region_boxes_nms(want, 13, 13, 5, 80, anchors, im_w, im_h, 416, 416, thresh, 0.45)
        gb, gp, go = oracle.region_boxes_nms(want, 13, 13, 5, 80, anchors, im_w, im_h, 416, 416, thresh, 0.45)
        assert np.array_equal(gb.view(np.uint32), wb.view(np.uint32))
        assert np.array_equal(gp.view(np.uint32), wp.view(np.uint32))
        assert np.array_equal(go.view(np.uint32), wo.view(np.uint32))


def test_nms_tie_order_matches_reference(oracle, ref16):
    """equal probabilities and overlapping boxes: do_nms_sort's qsort starts each class from the order the previous class left
    (src/core/yolo_post.cpp:70-74); the restatement must resolve the ties the same way (same list, same order, same bits)"""
    from test_host_logic import _tied_region
    anchors = np.array([0.57273, 0.677385, 1.87446, 2.06253, 3.33843, 5.47434, 7.88282, 3.52778, 9.77052, 9.16828], np.float32)
    for seed in range(4):
        region = _tied_region(seed)
        wb, wp, wo = ref16.region_boxes_nms(region, 13, 13, 5, 6, anchors, 640, 480, 416, 416, 0.1, 0.45)
        gb, gp, go = oracle.region_boxes_nms(region, 13, 13, 5, 6, anchors, 640, 480, 416, 416, 0.1, 0.45)
        assert np.array_equal(gb.view(np.uint32), wb.view(np.uint32))
        assert np.array_equal(gp.view(np.uint32), wp.view(np.uint32))
        assert np.array_equal(go.view(np.uint32), wo.view(np.uint32))


@pytest.mark.slow
@pytest.mark.parametrize("width,classes,table,precision", [(416, 3, "stress", "int16"), (608, 20, "stress", "int16"),
                                                            (416, 20, "saturate", "int16"), (416, 3, "default", "fp32")])
def test_generalised_driver_vs_real_yolo2_fpga(width, classes, table, precision, oracle, ref16, ref32):
    """VOC / 608 cannot run through the reference's hard-wired yolov2_hls_ps; the real YOLO2_FPGA driven
    layer by layer (oracle/ref_driver.py) must agree with the oracle's own net forward on every ofm."""
    net = ycfg.parse_network_cfg(ycfg.yolov2_cfg_text(width, width, classes, channel_div=8))
    pack = yw.synth_pack(net, precision, seed=9, table=table)
    frame = yw.synth_frames(net, 1, seed=77)[0]
    ref = ref16 if precision == "int16" else ref32
    want_region, want_dumps, _ = ref_net_forward(ref, net, frame, pack, keep_layers=True, helper=oracle)
    got_region, got_dumps = oracle.net_forward(net, frame, pack, dump_layers=True)
    for i, w in want_dumps.items():
        ow = net.layers[i].out_w
        if precision == "int16":
            assert np.array_equal(valid(got_dumps[i], ow), valid(w, ow)), f"layer {i}"
        else:
            assert np.abs(valid(got_dumps[i], ow) - valid(w, ow)).max() <= 1e-5 * max(np.abs(valid(w, ow)).max(), 1e-9), f"layer {i}"
    if precision == "int16":
        assert np.array_equal(got_region.view(np.uint32), want_region.view(np.uint32))
    else:
        assert np.abs(got_region - want_region).max() <= 1e-5


@pytest.mark.parametrize("w,h,c,nw,nh", [(640, 480, 3, 416, 416), (333, 500, 3, 416, 416), (416, 416, 3, 416, 416), (77, 31, 1, 40, 56)])
def test_letterbox_oracle_vs_reference(w, h, c, nw, nh, oracle, ref16):
    """orc_letterbox_u8 against the reference's own letterbox_image (yolo_image.cpp:146-165), bit for bit."""
    img = np.random.default_rng(w * 1000 + h).integers(0, 256, (h, w, c), dtype=np.uint8)
    assert np.array_equal(oracle.letterbox_u8(img, nw, nh).view(np.uint32), ref16.letterbox_u8(img, nw, nh).view(np.uint32))


# ---- rounding-group variants: the reference BUILT with --tn 8/16/32 (oracle/Makefile ref-variants) ----

@pytest.mark.parametrize("tn", [8, 16, 32])
def test_conv_int16_variant_builds(tn, oracle):
    """orc_conv_i16 with TN = tn against the unmodified YOLO2_FPGA of a reference compiled with Tn = tn, bit for bit"""
    from oracle.oracle import Ref, have_ref
    if not have_ref("int16", tn):
        pytest.skip(f"oracle/_ref/libref_int16_tn{tn}.so not built (make -C oracle ref-variants)")
    ref = Ref("int16", tn)
    for i, (c, n, k, w, h, q, amp) in enumerate([(64, 40, 3, 13, 13, (14, 10, 10, 10), 600), (37, 33, 3, 20, 11, (13, 9, 12, 7), 32767),
                                                 (96, 64, 1, 19, 19, (12, 12, 7, 8), 3000), (3, 16, 3, 26, 26, (14, 10, 10, 10), 600)]):
        a, x, wr, b, _ = make_conv_case(900 + 10 * tn + i, c, n, k, 1, w, h, 1, amp=amp, xamp=32767 if amp > 600 else 2000, tn=tn)
        assert a["TN"] == min(c, tn)
        want = ref.run_layer(x, wr, b, a, q)
        got = oracle_conv(oracle, a, x, wr, b, q)
        assert np.array_equal(valid(got, w), valid(want, w))


@pytest.mark.parametrize("tn,table", [(8, "default"), (16, "stress"), (32, "default")])
def test_net_forward_variant_builds(tn, table, oracle):
    """a whole (thin) YOLOv2 through the unmodified YOLO2_FPGA of a reference COMPILED with Tn = tn, layer by layer
    (oracle/ref_driver.py), against the oracle's net forward with the same tile parameters: every ofm and the region tensor bit for
    bit.  This pins the checker the GPU tests of the rounding-group variants use (test_whole_net_rounding_group_variant,
    test_full_width_rounding_group_variant)."""
    from oracle.oracle import Ref, have_ref
    if not have_ref("int16", tn):
        pytest.skip(f"oracle/_ref/libref_int16_tn{tn}.so not built (make -C oracle ref-variants)")
    ref = Ref("int16", tn)
    net = ycfg.parse_network_cfg(ycfg.yolov2_cfg_text(416, 416, 3, channel_div=8))
    pack = yw.synth_pack(net, "int16", seed=31 + tn, table=table, tn=tn)
    frame = yw.synth_frames(net, 1, seed=500 + tn)[0]
    oracle.set_tile_params(tn, 32)
    try:
        want_region, want_dumps, _ = ref_net_forward(ref, net, frame, pack, keep_layers=True, helper=oracle)
        got_region, got_dumps = oracle.net_forward(net, frame, pack, dump_layers=True)
    finally:
        oracle.set_tile_params(4, 32)
    for i, w in want_dumps.items():
        ow = net.layers[i].out_w
        assert np.array_equal(valid(got_dumps[i], ow), valid(w, ow)), f"tn {tn} layer {i}"
    assert np.array_equal(got_region.view(np.uint32), want_region.view(np.uint32))
