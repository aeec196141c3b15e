"""GPU parity at the shapes BASELINE.json states (configs[1..4]), FULL WIDTH and multi-frame: the CUDA network executor
against the UNMODIFIED reference accelerator entry YOLO2_FPGA (hls/models/yolov2/yolo2_accel.cpp:25-171, compiled into
oracle/_ref/libref_*.so) driven layer by layer through oracle/ref_driver.py - every conv / pool / reorg ofm of the first,
an interior and the last frame of the batch plus the region tensor, bit for bit (int16) or within 1e-4 relative (fp32).
When oracle/_ref was not built (no /root/reference at build time) the pinned C restatement (oracle/yolo2_oracle.c) checks
instead.  Each test also asserts WHICH kernels produced the bits (the tcgen05 kernel must have run on the deep layers)."""
import numpy as np
import pytest

from helpers import valid
from oracle.ref_driver import reference_frames
from yolo2_b200 import cfg as ycfg, weights as yw
from yolo2_b200.model import Yolo2Net

pytestmark = pytest.mark.gpu


def _run_case(width, classes, batch, max_batch, precision, table, check_frames, pack_seed=1, frame_seed=1000, tol=None):
    cfg_text = ycfg.yolov2_cfg_text(width, width, classes)
    net = ycfg.parse_network_cfg(cfg_text)
    pack = yw.synth_pack(net, precision, seed=pack_seed, table=table)
    frames = yw.synth_frames(net, batch, seed=frame_seed)
    want, _ = reference_frames(cfg_text, precision, pack_seed, table, frame_seed, check_frames)
    y = Yolo2Net(net, pack, max_batch=max_batch)
    try:
        region_compact = y.forward(frames)              # default: compact arena (what bench.py runs)
        y.set_debug_keep(True)                          # per-layer dumps need every tensor kept
        region = y.forward(frames)
        assert np.array_equal(region.view(np.uint32), region_compact.view(np.uint32)), "compact arena and keep-all mode differ"
        kernels = {i: y.layer_kernel(i) for i in range(len(net.layers))}
        last_chunk0 = ((batch - 1) // max_batch) * max_batch           # per-layer dumps exist for the frames of the LAST device pass
        for f in check_frames:
            wreg, dumps = want[f]
            if tol is None:
                assert np.array_equal(region[f].view(np.uint32), wreg.reshape(region[f].shape).view(np.uint32)), f"frame {f}: region tensor differs"
            else:
                assert np.abs(region[f] - wreg.reshape(region[f].shape)).max() <= tol
            if f < last_chunk0:
                continue
            for i, w in dumps.items():
                got = y.layer_output(i, f - last_chunk0)
                ow = net.layers[i].out_w
                if tol is None:
                    assert np.array_equal(valid(got, ow), valid(w, ow)), f"frame {f} layer {i} ({kernels[i]}) differs"
                else:
                    m = np.abs(valid(w, ow)).max()
                    assert np.abs(valid(got, ow) - valid(w, ow)).max() <= tol * max(m, 1e-6), f"frame {f} layer {i} ({kernels[i]})"
        return net, kernels
    finally:
        y.close()


def _assert_tensor_core_ran(net, kernels, min_layers):
    tc = [i for i, k in kernels.items() if k.startswith("conv_i16_tc2<")]
    assert len(tc) >= min_layers, f"tcgen05 kernel ran on layers {tc} only: {kernels}"
    assert all(net.layers[i].type == ycfg.CONV for i in tc)


def test_full_width_voc416_batch64():
    """BASELINE configs[2]: YOLOv2-VOC 416x416 INT16, batch 64 in one device pass (48-pixel tiles straddle frame boundaries)."""
    net, kernels = _run_case(416, 20, 64, 64, "int16", "default", [0, 29, 63])
    _assert_tensor_core_ran(net, kernels, 10)


def test_full_width_coco608_passes_of_2():
    """BASELINE configs[3] shape (608x608, route/reorg stress Q table), three frames through device passes of two: every layer of
    the last pass, the region tensor of every frame."""
    net, kernels = _run_case(608, 80, 3, 2, "int16", "stress", [0, 1, 2], pack_seed=2, frame_seed=3000)
    _assert_tensor_core_ran(net, kernels, 4)


def test_full_width_coco416_batch40():
    """BASELINE configs[4] shape: full-width COCO 416, 40 frames in one pass (the benched network, multi-frame)."""
    net, kernels = _run_case(416, 80, 40, 40, "int16", "default", [0, 17, 39], pack_seed=0)
    _assert_tensor_core_ran(net, kernels, 10)


def test_full_width_coco416_saturating_table():
    """full-range weights: the chain saturates on most steps, so the tcgen05 kernel's exact (slow) path produces the bits"""
    net, kernels = _run_case(416, 80, 2, 2, "int16", "saturate", [0, 1], pack_seed=3)
    _assert_tensor_core_ran(net, kernels, 10)


def test_full_width_fp32_per_layer():
    """BASELINE configs[1]: full-width COCO 416 fp32, per-layer ofm diff against the reference's float build <= 1e-4 relative."""
    net, kernels = _run_case(416, 80, 2, 2, "fp32", "default", [0, 1], pack_seed=4, tol=1e-4)
    assert any(k.startswith("conv_f32_c4") for k in kernels.values()), kernels


@pytest.mark.parametrize("tn,min_tc", [(16, 18), (8, 16)])
def test_full_width_rounding_group_variant(tn, min_tc, oracle):
    """SURVEY.md 8f-4 at full width: the network executor emulating a reference built with --tn 16 / 8, default policy (the deep
    layers on the tensor-core kernel with 2 / 4 rounding steps per MMA K slice, csrc/conv_i16_tc32.cu; the shallow ones on the
    grouped CUDA-core kernel), three frames in one pass: every layer of the first and last frame and all region tensors bit-exact
    to the oracle with the same tile parameters (itself pinned against the reference built with that Tn)."""
    from yolo2_b200.accel import Accelerator
    net = ycfg.parse_network_cfg(ycfg.yolov2_cfg_text(416, 416, 80))
    pack = yw.synth_pack(net, "int16", seed=5, table="default", tn=tn)
    frames = yw.synth_frames(net, 3, seed=7000)
    acc = Accelerator(0, "int16")
    acc.set_tile_params(tn, 32)
    oracle.set_tile_params(tn, 32)
    y = Yolo2Net(net, pack, max_batch=3, accel=acc)
    y.set_debug_keep(True)
    try:
        region = y.forward(frames)
        kernels = {i: y.layer_kernel(i) for i in range(len(net.layers))}
        for f in (0, 1, 2):
            want_region, dumps = oracle.net_forward(net, frames[f], pack, dump_layers=f != 1)
            assert np.array_equal(region[f].view(np.uint32), want_region.reshape(region[f].shape).view(np.uint32)), (tn, f)
            for i, want in (dumps or {}).items():
                ow = net.layers[i].out_w
                assert np.array_equal(valid(y.layer_output(i, f), ow), valid(want, ow)), (tn, f, i, kernels[i])
        tc = [i for i, k in kernels.items() if k.startswith("conv_i16_tc32<") and k.endswith(f",tn{tn}>")]
        assert len(tc) >= min_tc, kernels
    finally:
        oracle.set_tile_params(4, 32)
        y.close()
        acc.close()
