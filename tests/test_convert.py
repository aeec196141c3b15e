"""In-repo weight pipeline (yolo2_b200/convert.py): darknet file parsing, batch-norm folding and quantisation.
The reference delegates this to an un-vendored tool (weights/README.md:37-57), so these tests pin the
pipeline against first principles: an explicit numpy batch-norm convolution and the fp32 CUDA path."""
import os

import numpy as np
import pytest

from yolo2_b200 import cfg as ycfg, convert, weights as yw


def _fake_darknet(net, seed=0):
    rng = np.random.default_rng(seed)
    out = []
    for l in net.conv_layers:
        d = convert.DarknetConv(weights=rng.normal(0, 0.05, (l.n, l.c, l.size, l.size)).astype(np.float32),
                                biases=rng.normal(0, 0.2, l.n).astype(np.float32))
        if l.batch_normalize:
            d.scales = rng.uniform(0.5, 1.5, l.n).astype(np.float32)
            d.rolling_mean = rng.normal(0, 0.1, l.n).astype(np.float32)
            d.rolling_variance = rng.uniform(0.05, 1.0, l.n).astype(np.float32)
        out.append(d)
    return out


@pytest.fixture(scope="module")
def thin_net():
    return ycfg.parse_network_cfg(ycfg.yolov2_cfg_text(416, 416, 3, channel_div=16))


@pytest.mark.parametrize("major,minor", [(0, 1), (0, 2)])
def test_darknet_file_round_trip(major, minor, thin_net, tmp_path):
    layers = _fake_darknet(thin_net, 1)
    path = os.path.join(tmp_path, "net.weights")
    convert.save_darknet_weights(layers, path, major, minor, seen=12345)
    back = convert.load_darknet_weights(thin_net, path)
    assert len(back) == len(layers)
    for a, b, l in zip(layers, back, thin_net.conv_layers):
        assert np.array_equal(a.weights, b.weights) and np.array_equal(a.biases, b.biases)
        assert (b.scales is not None) == bool(l.batch_normalize)
        if l.batch_normalize:
            assert np.array_equal(a.rolling_variance, b.rolling_variance)
    with open(path, "r+b") as f:          # a truncated file is an error, not silently short weights
        f.truncate(os.path.getsize(path) - 8)
    with pytest.raises(RuntimeError):
        convert.load_darknet_weights(thin_net, path)


def test_fold_batchnorm_equals_explicit_batchnorm(thin_net):
    """conv(x; w', b') == scale * (conv(x; w) - mean) / (sqrt(var) + 1e-6) + bias on random patches (float64 check)"""
    layers = _fake_darknet(thin_net, 2)
    folded = convert.fold_batchnorm(layers)
    rng = np.random.default_rng(3)
    for d, (w, b), l in zip(layers, folded, thin_net.conv_layers):
        x = rng.normal(0, 1, (l.c, l.size, l.size))
        raw = np.tensordot(d.weights.astype(np.float64), x, axes=([1, 2, 3], [0, 1, 2]))
        if l.batch_normalize:
            want = (raw - d.rolling_mean) / (np.sqrt(d.rolling_variance.astype(np.float64)) + 1e-6) * d.scales + d.biases
        else:
            want = raw + d.biases
        got = np.tensordot(w.astype(np.float64), x, axes=([1, 2, 3], [0, 1, 2])) + b
        assert np.allclose(got, want, rtol=1e-5, atol=1e-5)


def test_quantize_pack_q_selection_and_layout(thin_net):
    folded = convert.fold_batchnorm(_fake_darknet(thin_net, 4))
    n = len(thin_net.conv_layers)
    pack = convert.quantize_pack(thin_net, folded, np.full(n + 1, 10, np.int32))
    assert pack.is_int16 and pack.weight_q.shape == (n,) and pack.act_q.shape == (n + 1,)
    off = 0
    for l, (w, b), qw in zip(thin_net.conv_layers, folded, pack.weight_q):
        assert np.abs(w).max() * 2.0 ** qw <= 32767 and (qw == 15 or np.abs(w).max() * 2.0 ** (qw + 1) > 32767)
        cnt = w.size
        want = yw.weight_reorg(np.rint(w.astype(np.float64) * 2.0 ** qw).astype(np.int16).reshape(l.n, l.c, -1), l.c, l.n, l.size,
                               min(l.n, yw.Tm), min(l.c, yw.Tn))
        assert np.array_equal(pack.weights[off:off + cnt], want)
        off += cnt
    assert convert.best_q(0.0) == 15 and convert.best_q(1.0) == 14 and convert.best_q(40000.0) == 0


def _route_entries(net):
    """table entries tied by `route -9` (source conv output, conv before the route) and ordered by the concat (skip, branch)"""
    convs = [i for i, l in enumerate(net.layers) if l.type == ycfg.CONV]
    no = {i: k for k, i in enumerate(convs)}
    single = next(i for i, l in enumerate(net.layers) if l.type == ycfg.ROUTE and len(l.inputs) == 1)
    multi = next(i for i, l in enumerate(net.layers) if l.type == ycfg.ROUTE and len(l.inputs) == 2)
    src = no[net.layers[single].inputs[0]] + 1
    before_route = no[max(c for c in convs if c < single)] + 1
    reorg = next(a for a in net.layers[multi].inputs if net.layers[a].type == ycfg.REORG)
    skip = no[next(b for b in net.layers[multi].inputs if b != reorg)] + 1
    branch = no[max(c for c in convs if c < reorg)] + 1
    return src, before_route, skip, branch


def test_harmonise_route_q_rules(thin_net):
    """the driver loop reads Qa_in by table position and only rescales the reorg branch (yolo2_model.cpp:311-336,379-399): the
    calibrated table must tie the `route -9` entries and keep Q_skip <= Q_branch; everything else stays as calibrated"""
    src, before_route, skip, branch = _route_entries(thin_net)
    assert (src, before_route, skip, branch) == (13, 20, 20, 21)       # YOLOv2: conv 12 = layer 16, conv 19 = layer 24, conv 20 = layer 26
    n = len(thin_net.conv_layers)
    rng = np.random.default_rng(0)
    for _ in range(50):
        q = rng.integers(3, 14, n + 1).astype(np.int32)
        h = convert.harmonise_route_q(thin_net, q)
        assert h[src] == h[before_route] and h[skip] <= h[branch]
        assert (h <= q).all()
        untouched = [k for k in range(n + 1) if k not in (src, before_route, skip)]
        assert np.array_equal(h[untouched], q[untouched])
        assert h[src] == min(q[src], q[before_route], q[branch]) or h[src] == min(q[src], q[before_route])
        assert np.array_equal(convert.harmonise_route_q(thin_net, h), h)     # idempotent


@pytest.mark.gpu
def test_route_q_mismatch_is_harmonised(thin_net, tmp_path):
    """a model whose skip conv (layer 24) produces much smaller values than the route source (layer 16): the raw calibration
    gives them different Qs, which the driver loop would mis-read; with the harmonised table the int16 net tracks the fp32 net,
    with the raw table it does not"""
    from yolo2_b200.model import Yolo2Net
    layers = _fake_darknet(thin_net, 9)
    def scale_output(d, g):                                # folded conv output x g (through the batch-norm gain when there is one)
        if d.scales is not None:
            d.scales = d.scales * np.float32(g)
        else:
            d.weights = d.weights * np.float32(g)
        d.biases = d.biases * np.float32(g)
    scale_output(layers[12], 64.0)                        # conv 12 = layer 16 (the route source): large outputs, low Q
    scale_output(layers[13], 1.0 / 64)                    #   ... the main branch is brought back to its usual range by the next conv
    scale_output(layers[19], 1.0 / 16)                    # conv 19 = layer 24 (the skip conv): small outputs, Q at the cap
    folded = convert.fold_batchnorm(layers)
    fp32 = convert.make_fp32_pack(thin_net, folded)
    frames = yw.synth_frames(thin_net, 2, seed=78)
    raw = convert.calibrate_activation_q(thin_net, fp32, frames, harmonise=False)
    src, before_route, skip, branch = _route_entries(thin_net)
    assert raw[src] != raw[before_route]                  # the scenario really has the inconsistency
    errs = {}
    for name, table in (("raw", raw), ("harmonised", convert.harmonise_route_q(thin_net, raw))):
        ya, yb = Yolo2Net(thin_net, fp32, max_batch=2), Yolo2Net(thin_net, convert.quantize_pack(thin_net, folded, table), max_batch=2)
        try:
            errs[name] = float(np.abs(ya.forward(frames) - yb.forward(frames)).mean())
        finally:
            ya.close(); yb.close()
    assert errs["harmonised"] < 0.04 and errs["harmonised"] < 0.5 * errs["raw"], errs


@pytest.mark.gpu
def test_darknet_to_int16_end_to_end(thin_net, tmp_path):
    """darknet file -> fold -> fp32 pack -> activation calibration on the CUDA fp32 path -> int16 pack -> files -> reload;
    the int16 network's region output tracks the fp32 network's."""
    from yolo2_b200.model import Yolo2Net
    path = os.path.join(tmp_path, "net.weights")
    convert.save_darknet_weights(_fake_darknet(thin_net, 5), path)
    frames = yw.synth_frames(thin_net, 2, seed=77)
    fp32, i16 = convert.convert_darknet(thin_net, path, frames, str(tmp_path))
    assert i16.act_q.min() >= 0 and i16.act_q.max() <= 15
    re16 = yw.load_reference_files(thin_net, "int16", str(tmp_path))
    assert np.array_equal(re16.weights, i16.weights) and np.array_equal(re16.act_q, i16.act_q)
    ya, yb = Yolo2Net(thin_net, fp32, max_batch=2), Yolo2Net(thin_net, re16, max_batch=2)
    try:
        ra, rb = ya.forward(frames), yb.forward(frames)
    finally:
        ya.close(); yb.close()
    err = np.abs(ra - rb)                      # sigmoid/softmax outputs in [0,1], box offsets O(1)
    # the datapath rounds to the OUTPUT Q after every 4-MAC step (16-bit accumulator), so the int16 net carries accumulated
    # rounding noise of a few per cent by construction; a wrong Q table, layout or fold is off by O(1)
    assert err.mean() < 0.04 and err.max() < 0.5, (err.mean(), err.max())
