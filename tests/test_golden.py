"""Oracle vs the committed golden fixtures (outputs of the REAL reference, frozen by
oracle/gen_golden.py).  Runs anywhere: needs neither /root/reference nor a GPU."""
import hashlib
import os

import numpy as np
import pytest

from oracle.gen_golden import ARG_KEYS
from yolo2_b200 import cfg as ycfg, weights as yw

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def cases():
    return np.load(os.path.join(GOLD, "layer_cases.npz"))


@pytest.fixture(scope="module")
def full():
    return np.load(os.path.join(GOLD, "yolov2_coco416_full.npz"))


def _args(arr):
    return dict(zip(ARG_KEYS, [int(v) for v in arr]))


@pytest.mark.parametrize("i", range(9))
def test_conv_int16_golden(i, cases, oracle):
    a = _args(cases[f"ci16_{i}_args"])
    q = [int(v) for v in cases[f"ci16_{i}_q"]]
    got = oracle.conv(cases[f"ci16_{i}_x"], cases[f"ci16_{i}_w"], cases[f"ci16_{i}_b"], a["IFM_num"], a["OFM_num"],
                      a["Ksize"], a["Kstride"], a["Input_w"], a["Input_h"], a["Output_w"], a["Output_h"], a["Padding"],
                      a["IsNL"], a["TM"], a["TN"], *q)
    ow = a["Output_w"]
    assert np.array_equal(got[..., :ow], cases[f"ci16_{i}_out"][..., :ow])


@pytest.mark.parametrize("i", range(4))
def test_conv_fp32_golden(i, cases, oracle):
    a = _args(cases[f"cf32_{i}_args"])
    got = oracle.conv(cases[f"cf32_{i}_x"], cases[f"cf32_{i}_w"], cases[f"cf32_{i}_b"], a["IFM_num"], a["OFM_num"],
                      a["Ksize"], a["Kstride"], a["Input_w"], a["Input_h"], a["Output_w"], a["Output_h"], a["Padding"],
                      a["IsNL"], a["TM"], a["TN"])
    ow = a["Output_w"]
    want = cases[f"cf32_{i}_out"][..., :ow]
    # same operation order as the reference: the restatement reproduces it to the last bit on this compiler,
    # but only the 1e-4 relative bar of BASELINE.json is claimed
    assert np.abs(got[..., :ow] - want).max() <= 1e-4 * np.abs(want).max()


@pytest.mark.parametrize("i", range(3))
def test_maxpool_golden(i, cases, oracle):
    a = _args(cases[f"pool_{i}_args"])
    ow = a["Output_w"]
    got = oracle.maxpool(cases[f"pool_{i}_x"], a["IFM_num"], a["Ksize"], a["Kstride"], a["Input_w"], a["Input_h"], ow, a["Output_h"])
    assert np.array_equal(got[..., :ow], cases[f"pool_{i}_out"][..., :ow])
    gotf = oracle.maxpool(cases[f"poolf_{i}_x"], a["IFM_num"], a["Ksize"], a["Kstride"], a["Input_w"], a["Input_h"], ow, a["Output_h"])
    assert np.array_equal(gotf[..., :ow], cases[f"poolf_{i}_out"][..., :ow])


def test_region_and_boxes_golden(cases, oracle):
    reg = oracle.region_forward(cases["region_in"], 7, 7, 5, 20)
    assert np.array_equal(reg.view(np.uint32), cases["region_out"].view(np.uint32))
    b, p, o = oracle.region_boxes_nms(cases["region_out"], 7, 7, 5, 20, cases["region_anchors"], 640, 480, 224, 224, 0.3, 0.45)
    assert np.array_equal(b.view(np.uint32), cases["det_boxes"].view(np.uint32))
    assert np.array_equal(p.view(np.uint32), cases["det_probs"].view(np.uint32))
    assert np.array_equal(o.view(np.uint32), cases["det_obj"].view(np.uint32))


@pytest.mark.slow
@pytest.mark.parametrize("tag,table,seed", [("default", "default", 1), ("stress", "stress", 2)])
def test_full_coco416_region_golden(tag, table, seed, full, oracle):
    """The oracle's generalised driver reproduces the UNMODIFIED yolov2_hls_ps bit for bit."""
    net = ycfg.parse_network_cfg(ycfg.yolov2_cfg_text(416, 416, 80))
    pack = yw.synth_pack(net, "int16", seed=seed, table=table)
    frame = yw.synth_frames(net, 1, seed=1000)[0]
    if hashlib.sha256(pack.weights.tobytes()).digest() != full[f"{tag}_weights_sha"].tobytes() or \
            hashlib.sha256(frame.tobytes()).digest() != full[f"{tag}_frame_sha"].tobytes():
        pytest.skip("numpy RNG stream differs from the one the golden was generated with")
    region, _ = oracle.net_forward(net, frame, pack)
    assert np.array_equal(region.reshape(-1).view(np.uint32), full[f"{tag}_region"].view(np.uint32))


# ---- image front-end: oracle vs the reference's letterbox_image frozen in letterbox.npz ----

@pytest.mark.parametrize("i", range(7))
def test_letterbox_golden(i, oracle):
    g = np.load(os.path.join(GOLD, "letterbox.npz"))
    nw, nh = (int(v) for v in g[f"lb_{i}_net"])
    got = oracle.letterbox_u8(g[f"lb_{i}_img"], nw, nh)
    assert got.dtype == np.float32 and np.array_equal(got.view(np.uint32), g[f"lb_{i}_out"].view(np.uint32))


# ---- rounding-group variants: outputs of the reference BUILT with --tn 8 / 32, frozen in layer_cases_tn_variants.npz ----

@pytest.mark.parametrize("i", range(5))
def test_conv_int16_tn_variant_golden(i, oracle):
    g = np.load(os.path.join(GOLD, "layer_cases_tn_variants.npz"))
    a = _args(g[f"v_{i}_args"])
    q = [int(v) for v in g[f"v_{i}_q"]]
    assert a["TN"] == min(int(g[f"v_{i}_tn"][0]), a["IFM_num"])
    got = oracle.conv(g[f"v_{i}_x"], g[f"v_{i}_w"], g[f"v_{i}_b"], a["IFM_num"], a["OFM_num"], a["Ksize"], a["Kstride"], a["Input_w"],
                      a["Input_h"], a["Output_w"], a["Output_h"], a["Padding"], a["IsNL"], a["TM"], a["TN"], *q)
    ow = a["Output_w"]
    assert np.array_equal(got[..., :ow], g[f"v_{i}_out"][..., :ow])
