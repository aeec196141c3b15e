"""Host-side logic and the C-ABI surface, without a GPU."""
import ctypes as C
import os
import re
import sys

import numpy as np
import pytest

from yolo2_b200 import _capi, cfg as ycfg, weights as yw

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = _capi.load_library()
    header = open(os.path.join(ROOT, "include", "yolo2cuda.h")).read()
    declared = set(re.findall(r"\b(yolo2cuda_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    assert declared == set(_capi.SYMBOLS), declared ^ set(_capi.SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None


def test_no_gpu_means_init_error_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _capi.load_library()
    ctx = C.c_void_p()
    assert lib.yolo2cuda_create(C.byref(ctx), 0, 16) == _capi.INIT_ERROR
    from yolo2_b200.accel import Accelerator
    with pytest.raises(_capi.Yolo2CudaError):
        Accelerator(0, "int16")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "yolo-fpga-accelerator_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(d, f), errors="ignore").read()
                for pat in ("import oracle", "from oracle", "oracle/", "liboracle", "orc_", "_ref/", "libref"):
                    assert pat not in txt, f"{f} references the checker ({pat})"


def test_cfg_matches_reference_model_tables():
    net = ycfg.parse_network_cfg(ycfg.yolov2_cfg_text(416, 416, 80))
    assert len(net.layers) == 32
    counts = net.weight_counts()
    # hls/models/yolov2/model_config.cpp:4-10
    assert [c[0] for c in counts] == [864, 18432, 73728, 8192, 73728, 294912, 32768, 294912, 1179648, 131072, 1179648,
                                      131072, 1179648, 4718592, 524288, 4718592, 524288, 4718592, 9437184, 9437184,
                                      32768, 11796480, 435200]
    assert [c[1] for c in counts] == [32, 64, 128, 64, 128, 256, 128, 256, 512, 256, 512, 256, 512, 1024, 512, 1024, 512,
                                      1024, 1024, 1024, 64, 1024, 425]
    assert sum(c[0] for c in counts) == 50941792 and sum(c[1] for c in counts) == 10761
    l = net.layers
    assert (l[25].type, l[25].inputs) == (ycfg.ROUTE, [16]) and (l[28].type, l[28].inputs) == (ycfg.ROUTE, [27, 24])
    assert (l[27].out_c, l[27].out_h, l[27].out_w) == (256, 13, 13) and l[29].c == 1280
    assert (l[31].n, l[31].classes, l[31].w) == (5, 80, 13)
    big = ycfg.parse_network_cfg(ycfg.yolov2_cfg_text(608, 608, 80))
    assert [x.out_w for x in big.layers if x.type == ycfg.MAXPOOL] == [304, 152, 76, 38, 19]


def test_cfg_rejects_sections_outside_the_path():
    with pytest.raises(ValueError):
        ycfg.parse_network_cfg("[net]\nwidth=32\nheight=32\nchannels=3\n[upsample]\nstride=2\n")
    with pytest.raises(ValueError):
        ycfg.parse_network_cfg("[convolutional]\nfilters=3\n\n")


def test_weight_reorg_formula_and_oracle_agree(oracle):
    rng = np.random.default_rng(0)
    for ifm, ofm, k in [(3, 32, 3), (64, 70, 3), (17, 425, 1), (5, 5, 2)]:
        w = rng.integers(-999, 999, (ofm, ifm, k * k)).astype(np.int16)
        tm, tn = min(ofm, 32), min(ifm, 4)
        a = yw.weight_reorg(w, ifm, ofm, k, tm, tn)
        assert np.array_equal(a, oracle.weight_reorg(w, ifm, ofm, k, tm, tn))
        # SURVEY.md appendix A address formula
        for _ in range(50):
            m, c, t = int(rng.integers(ofm)), int(rng.integers(ifm)), int(rng.integers(k * k))
            m0, n0 = (m // tm) * tm, (c // tn) * tn
            tmm, tnn = min(tm, ofm - m0), min(tn, ifm - n0)
            off = m0 * ifm * k * k + tmm * n0 * k * k + (t * tmm + (m - m0)) * tnn + (c - n0)
            assert a[off] == w[m, c, t]


def test_weight_files_roundtrip_with_odd_length_padding(tmp_path):
    net = ycfg.parse_network_cfg(ycfg.yolov2_cfg_text(416, 416, 80, channel_div=16))
    pack = yw.synth_pack(net, "int16", seed=3, table="stress")
    yw.save_reference_files(pack, net, str(tmp_path))
    # the 425-entry bias layer is odd: one pad element follows it in the file (yolo2_model.cpp:216-223)
    nb = os.path.getsize(tmp_path / "bias_int16.bin") // 2
    assert nb == pack.bias.size + sum(c[1] & 1 for c in net.weight_counts())
    back = yw.load_reference_files(net, "int16", str(tmp_path))
    assert np.array_equal(back.weights, pack.weights) and np.array_equal(back.bias, pack.bias)
    assert np.array_equal(back.act_q, pack.act_q)
    (tmp_path / "weights_reorg_int16.bin").write_bytes(b"\0" * 10)
    with pytest.raises(RuntimeError, match="weights file too small"):
        yw.load_reference_files(net, "int16", str(tmp_path))


def test_quantize_input_rounding(oracle):
    x = np.array([0.0, 0.5, -0.5, 1.5 / 1024, 2.5 / 1024, -2.5 / 1024, 40.0, -40.0, 0.99999], np.float32)
    q = oracle.quantize_input(x, 10)
    assert q.tolist() == [0, 512, -512, 2, 3, -3, 32767, -32768, 1024]   # half away from zero, saturating


def test_round_shift_semantics(oracle):
    assert oracle.round_shift(5, 1) == 3 and oracle.round_shift(-5, 1) == -2        # half up toward +inf
    assert oracle.round_shift(-1, 4) == 0 and oracle.round_shift(-9, 4) == -1
    assert oracle.round_shift(3, -2) == 12 and oracle.round_shift(7, 0) == 7
    assert oracle.round_shift(1 << 40, 35) == oracle.round_shift(1 << 40, 30)       # |shift| clamped to 30


def test_host_detections_match_oracle(oracle):
    from yolo2_b200.model import region_detections
    rng = np.random.default_rng(5)
    net = ycfg.parse_network_cfg(ycfg.yolov2_cfg_text(416, 416, 80))
    l = net.layers[-1]
    region = oracle.region_forward(rng.normal(0, 2.5, (425, 13, 13)).astype(np.float32), 13, 13, 5, 80)
    for im_w, im_h, thresh in [(768, 576, 0.2), (300, 500, 0.05)]:
        b, p, o = region_detections(net, region, im_w, im_h, thresh, 0.45)
        wb, wp, wo = oracle.region_boxes_nms(region, 13, 13, 5, 80, l.anchors, im_w, im_h, 416, 416, thresh, 0.45)
        live = wo > 0
        key = lambda bb, pp, oo: sorted((tuple(x.tolist()), float(z), tuple(np.nonzero(q)[0].tolist()), tuple(q[q > 0].tolist()))
                                        for x, q, z in zip(bb, pp, oo))
        assert len(b) == live.sum() > 10
        assert key(b, p, o) == key(wb[live], wp[live], wo[live])


def _tied_region(seed, classes=6, n=5, w=13, h=13):
    """a post-activation region tensor with MANY equal class probabilities and heavily overlapping boxes: which of two tied,
    overlapping candidates survives NMS then depends on the order the previous class's qsort left (yolo_post.cpp:70-74)"""
    rng = np.random.default_rng(seed)
    r = np.zeros((n, 5 + classes, h, w), np.float32)
    r[:, 0:2] = 0.5
    r[:, 2:4] = rng.choice([1.0, 1.25], (n, 2, h, w))                  # exp(.) * anchor / 13: boxes several cells wide
    r[:, 4] = rng.choice([0.0, 0.5, 0.5, 0.75], (n, h, w))            # objectness: few distinct values
    r[:, 5:] = rng.choice([0.125, 0.25, 0.5], (n, classes, h, w))     # class probabilities: few distinct values
    return r


def test_host_detections_tie_order_matches_oracle(oracle):
    """equal scores + overlapping boxes: the carried-over qsort order decides the survivor; host path == checker"""
    from yolo2_b200.model import region_detections
    net = ycfg.parse_network_cfg(ycfg.yolov2_cfg_text(416, 416, 6))
    l = net.layers[-1]
    key = lambda bb, pp, oo: sorted((tuple(x.tolist()), float(z), tuple(np.nonzero(q)[0].tolist()), tuple(q[q > 0].tolist()))
                                    for x, q, z in zip(bb, pp, oo))
    suppressed = 0
    for seed in range(4):
        region = _tied_region(seed)
        b, p, o = region_detections(net, region, 640, 480, 0.1, 0.45)
        wb, wp, wo = oracle.region_boxes_nms(region, 13, 13, 5, 6, l.anchors, 640, 480, 416, 416, 0.1, 0.45)
        live = wo > 0
        assert len(b) == live.sum() > 100
        assert key(b, p, o) == key(wb[live], wp[live], wo[live])
        suppressed += int((p == 0).sum())
    assert suppressed > 1000          # NMS really decided between tied candidates


def test_glibc_expf_restatement_matches_libm(tmp_path):
    """oracle/expf_check.c: the double-arithmetic restatement of glibc's expf that the CUDA detection kernel uses for the box
    width / height equals the host libm's expf (what the reference's std::exp(float) calls) on 2^22 inputs, bit for bit."""
    import os, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(tmp_path, "expf_check")
    subprocess.check_call(["gcc", "-O2", "-fno-builtin", "-ffp-contract=off", "-o", exe, os.path.join(root, "oracle", "expf_check.c"), "-lm"])
    assert subprocess.check_output([exe]).decode().strip() == "0"


def test_detections_jsonl_format():
    """the board application's JSONL record (linux_app/src/main.c:1028-1075): best class per box, normalised box, pixel corners"""
    import json
    import numpy as np
    from yolo2_b200.model import detections_jsonl
    boxes = np.array([[0.5, 0.5, 0.2, 0.4], [0.1, 0.1, 0.05, 0.05]], np.float32)
    probs = np.array([[0.0, 0.9, 0.3], [0.1, 0.0, 0.0]], np.float32)
    rec = json.loads(detections_jsonl(boxes, probs, 640, 480, labels=["a", "b", "c"], thresh=0.25, source="cam0", frame_index=7))
    assert rec["width"] == 640 and rec["frame_index"] == 7 and len(rec["detections"]) == 1
    d = rec["detections"][0]
    assert d["class_id"] == 1 and d["label"] == "b" and abs(d["prob"] - 0.9) < 1e-5
    f = np.float32      # (int)((b.x - b.w * 0.5f) * (float)frame_w) in float arithmetic, linux_app/src/main.c:1054-1057
    assert d["bbox_px"] == {"x0": int((f(0.5) - f(0.2) * f(0.5)) * f(640)), "y0": int((f(0.5) - f(0.4) * f(0.5)) * f(480)),
                            "x1": int((f(0.5) + f(0.2) * f(0.5)) * f(640)), "y1": int((f(0.5) + f(0.4) * f(0.5)) * f(480))}
    assert d["bbox_px"]["x0"] == 256


def test_bench_reference_arm_prints_one_json_line():
    """bench.py contract: stdout is exactly ONE JSON line (native banners such as NCCL's go to stderr through the fd redirection);
    checked on the CPU-only reference arm with a tiny sample."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--ref-procs", "1"], capture_output=True, text=True, timeout=900, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port")


def test_glibc_exp_restatement_matches_libm(tmp_path):
    """oracle/exp_check.c: the restatement of glibc's double exp (FMA form, data from glibc_exp_data.h) that the CUDA region kernel
    uses for the logistic and the softmax equals the host libm's exp (what the reference's yolo_math.cpp calls) on 2^22 inputs -
    the region head's range, float-valued arguments, the scaled special cases, arbitrary bit patterns - bit for bit."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(tmp_path, "exp_check")
    subprocess.check_call(["gcc", "-O2", "-fno-builtin", "-ffp-contract=off", "-o", exe, os.path.join(root, "oracle", "exp_check.c"), "-lm"])
    out = subprocess.run([exe, str(1 << 22)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "0", (out.stdout, out.stderr)


def test_glibc_exp_data_header_matches_this_libm(tmp_path):
    """the committed csrc/glibc_exp_data.h is what csrc/gen_glibc_exp_data.py reads out of the libm the tests run with"""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libm = "/lib/x86_64-linux-gnu/libm.so.6"
    if not os.path.exists(libm):
        pytest.skip("no x86-64 glibc libm at the usual path")
    out = os.path.join(tmp_path, "h.h")
    subprocess.check_call([sys.executable, os.path.join(root, "yolo-fpga-accelerator_b200", "csrc", "gen_glibc_exp_data.py"), libm, out])
    assert open(out).read() == open(os.path.join(root, "yolo-fpga-accelerator_b200", "csrc", "glibc_exp_data.h")).read()


def test_pass_size_model_follows_the_persistent_kernel():
    """model.pass_efficiency mirrors the work-item arithmetic of the persistent tcgen05 kernel: for YOLOv2-416 on 148 SMs a multiple
    of 21 frames is exactly four rounds of (48-pixel x 128-channel) items on the 13x13x1024 layers, so best_pass_size / best_ramp_size
    pick multiples of 21 and the 364-frame passes of round 1 rate lower; the policy mirror names the HALF-mode and CUDA-core layers."""
    from yolo2_b200.model import _tensor_core_tile, best_pass_size, best_ramp_size, pass_efficiency
    net = ycfg.parse_network_cfg(ycfg.yolov2_cfg_text(416, 416, 80))
    assert (21 * 169 + 47) // 48 * 8 == 4 * 148
    b = best_pass_size(net, 128, 400)
    assert b % 21 == 0 and b >= 357
    assert best_ramp_size(net, b) % 21 == 0 and 32 <= best_ramp_size(net, b) <= b // 4
    assert pass_efficiency(net, b) > 0.999 > pass_efficiency(net, 364) > 0.99
    tiles = {i: _tensor_core_tile(l) for i, l in enumerate(net.layers) if l.type == ycfg.CONV}
    assert tiles[0] is None                                        # 3 input channels: conv_i16_g1_kernel
    assert [i for i, t in tiles.items() if t == (64, 96)] == [2, 5, 26]      # 64 output channels: HALF mode
    assert all(t == (128, 48) for i, t in tiles.items() if i not in (0, 2, 5, 26))


def test_variant_tensor_core_kernel_register_pool_balances():
    """csrc/conv_i16_tc32.cu re-splits registers between its warp roles with setmaxnreg, which moves registers inside the CTA's OWN
    pool only: what the helper warps hand back below the launch allocation must cover what the epilogue warps take above it, or
    the epilogue spins in USETMAXREG.TRY_ALLOC for ever (the hang of the first six-builder build).  Checked here against the
    launch allocation ptxas reports for every instantiation in the build log, and: no instantiation spills."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "yolo-fpga-accelerator_b200")
    src = open(os.path.join(pkg, "csrc", "conv_i16_tc32.cu")).read()
    log_path = os.path.join(pkg, "lib", "conv_i16_tc32.ptxas.log")
    if not os.path.exists(log_path):
        pytest.skip("library not built here (make writes lib/*.ptxas.log)")
    log = open(log_path).read()
    m = re.search(r"kLaunchRegs = (\d+), kEpiRegs = (\d+), kHelperRegs = (\d+)", src)
    launch, epi, helper = (int(v) for v in m.groups())
    threads = 32 * int(re.search(r"kThreads = (\d+) \* 32", src).group(1))
    epi_warps = int(re.search(r"kEpiWarps = (\d+)", src).group(1))
    helper_warps = threads // 32 - epi_warps
    assert launch * threads <= 65536
    assert epi_warps * (epi - launch) <= helper_warps * (launch - helper)
    used = [int(v) for v in re.findall(r"Used (\d+) registers, used 1 barriers", log)]
    assert len(used) == 54 and set(used) == {launch}, sorted(set(used))          # 2 kernel sizes x 9 shifts x 3 rounding groups
    assert not re.search(r"[1-9]\d* bytes spill", log)


@pytest.mark.parametrize("tn", [8, 16, 32])
@pytest.mark.parametrize("c,n,k,w,h,q,amp", [(37, 5, 3, 6, 5, (13, 9, 12, 7), 32767), (20, 3, 1, 7, 4, (14, 10, 10, 10), 600),
                                             (64, 4, 3, 5, 5, (12, 10, 8, 9), 4000)])
def test_variant_tensor_core_decomposition_mirror(tn, c, n, k, w, h, q, amp, oracle):
    """numpy mirror of what csrc/conv_i16_tc32.cu asks of the tensor cores and of the CUDA cores behind them, against the oracle
    with TN = tn: per K = 32 slice the weight tile A (K byte kk = channel kk % tn of the rounding group of chain step
    slice * 32/tn + kk / tn), the block-diagonal activation tile B (column (step slot, pixel) non-zero only in that slot's K rows),
    signed-hi / unsigned-lo byte planes, HH / M / LL as three int32 products, then per column in chain order
    t = 256 M + LL + half, d = HH 2^(16-so) + (t >> so), acc = clamp(acc + d) on the offset accumulator, leaky / 10."""
    from helpers import make_conv_case, oracle_conv, valid
    a, x, wr, b, wd = make_conv_case(7 * tn + c + k, c, n, k, 1, w, h, 1, amp=amp, xamp=32767 if amp > 600 else 2000, tn=tn)
    Qw, Qa_in, Qa_out, Qb = q
    so, sb = Qa_in + Qw - Qa_out, Qb - Qa_out
    assert 8 <= so <= 16
    want = valid(oracle_conv(oracle, a, x, wr, b, q), w)
    K2, pad, S = k * k, k // 2, 32 // tn
    nsteps = -(-c // tn) * K2
    nslices = -(-nsteps // S)
    xp = np.zeros((c, h + 2 * pad, w + 2 * pad), np.int64)
    xp[:, pad:pad + h, pad:pad + w] = x[:, :, :w]
    npix = h * w
    rs = lambda v, s: (v + (1 << (s - 1))) >> s if s > 0 else v << -s          # round-half-up arithmetic shift
    U = np.empty((n, npix), np.int64)
    rb = (1 << (38 - so)) + 2
    U[:] = np.clip(rs(b.astype(np.int64), sb) + 32768, -rb, 65535 + rb)[:, None]
    for sl in range(nslices):
        A = np.zeros((n, 32), np.int64)                                     # weights of the slice: row = output channel
        B = np.zeros((32, S, npix), np.int64)                               # activations: column = (step slot, pixel)
        for kk in range(32):
            sp, t = kk // tn, kk % tn
            step = sl * S + sp
            g, tap = step // K2, step % K2
            ch = g * tn + t
            if step >= nsteps or ch >= c:
                continue
            A[:, kk] = wd[:, ch, tap]
            ti, tj = tap // k, tap % k
            B[kk, sp, :] = xp[ch, ti:ti + h, tj:tj + w].reshape(-1)          # only in the K rows of its own step slot
        Ah, Al = A >> 8, A & 255                                            # signed hi, unsigned lo
        Bh, Bl = B >> 8, B & 255
        HH = np.einsum("mk,ksp->msp", Ah, Bh)
        M = np.einsum("mk,ksp->msp", Ah, Bl) + np.einsum("mk,ksp->msp", Al, Bh)
        LL = np.einsum("mk,ksp->msp", Al, Bl)
        assert np.abs(M).max() < 2 ** 31 and np.abs(HH).max() < 2 ** 31      # int32 accumulators of the MMA suffice
        for sp in range(S):                                                  # chain order; slots past the last step add d = 0
            t = 256 * M[:, sp] + LL[:, sp] + (1 << (so - 1))
            assert np.abs(t).max() < 2 ** 31
            d = HH[:, sp] * (1 << (16 - so)) + (t >> so)
            U = np.clip(U + d, 0, 65535)
    out = U - 32768
    out = np.where(out < 0, -((-out) // 10), out)                            # leaky: C division truncates toward zero
    assert np.array_equal(out.reshape(n, h, w).astype(np.int16), want)


@pytest.mark.parametrize("c,n,k,w,h,q,amp,xamp,known_xmax", [
    (40, 6, 3, 6, 5, (14, 10, 10, 10), 600, 2000, True),       # the bench's default table: every K-block on the fast path
    (40, 6, 3, 6, 5, (14, 10, 10, 10), 600, 2000, False),      # producer unknown: xmax = 32768 assumed, fewer fast blocks, same bits
    (37, 5, 3, 5, 4, (13, 9, 12, 7), 32767, 32767, True),      # full range: saturating chain, mostly the exact step
    (64, 4, 1, 7, 3, (12, 10, 6, 9), 3000, 9000, True),        # so = 16, 1x1
    (21, 3, 3, 4, 4, (9, 8, 8, 8), 300, 500, True),            # so = 9
    (48, 12, 3, 6, 6, (14, 10, 10, 6), 2500, 2000, True)])     # biases shifted left by 4: accumulators near both rails, fast and exact blocks mix
def test_headline_kernel_fast_path_is_exact_mirror(c, n, k, w, h, q, amp, xamp, known_xmax, oracle):
    """numpy mirror of the no-saturation fast path of csrc/conv_i16_tc2.cu (the reference's default Tn = 4): per K-block of seven chain
    steps the kernel tests Dsum <= acc + 32768 <= 65535 - Dsum with Dsum = ((sum of the block's 28 |w|) * xmax >> so) + 8 and, when
    it holds, replaces seven round-and-saturate steps by acc += 2^(16-so) * sum HH_s + sum ((256 M_s + LL_s + half) >> so).  Here
    both forms run side by side on every (channel, pixel, K-block): wherever the test holds they must agree, the mixed result must be
    the oracle's bits, and the test must not be vacuous (blocks pass in the moderate cases, the full-range case runs on the exact step)."""
    from helpers import make_conv_case, oracle_conv, valid
    a, x, wr, b, wd = make_conv_case(c * 31 + n + k, c, n, k, 1, w, h, 1, amp=amp, xamp=xamp)
    Qw, Qa_in, Qa_out, Qb = q
    so, sb = Qa_in + Qw - Qa_out, Qb - Qa_out
    assert 8 <= so <= 16
    want = valid(oracle_conv(oracle, a, x, wr, b, q), w)
    K2, pad, G = k * k, k // 2, -(-c // 4)
    nsteps = G * K2
    nkb = -(-nsteps // 7)
    xp = np.zeros((4 * G, h + 2 * pad, w + 2 * pad), np.int64)
    xp[:c, pad:pad + h, pad:pad + w] = x[:, :, :w]
    wz = np.zeros((n, 4 * G, K2), np.int64)
    wz[:, :c] = wd
    xmax = int(np.abs(x[:, :, :w].astype(np.int64)).max()) if known_xmax else 32768
    npix = h * w
    rs = lambda v, s: (v + (1 << (s - 1))) >> s if s > 0 else v << -s
    rb = (1 << (33 - so)) + 2
    U = np.empty((n, npix), np.int64)
    U[:] = np.clip(rs(b.astype(np.int64), sb) + 32768, -rb, 65535 + rb)[:, None]
    n_fast = n_exact = 0
    for kb in range(nkb):
        HH, T, wsum = [], [], np.zeros(n, np.int64)
        for s in range(7):
            sigma = kb * 7 + s
            hh = np.zeros((n, npix), np.int64)
            t = np.full((n, npix), 1 << (so - 1), np.int64)                   # the rounding constant rides in K row 28 of every step slot
            if sigma < nsteps:
                g, tap = sigma // K2, sigma % K2
                ti, tj = tap // k, tap % k
                W4 = wz[:, 4 * g:4 * g + 4, tap]                             # [n][4]
                X4 = xp[4 * g:4 * g + 4, ti:ti + h, tj:tj + w].reshape(4, -1)  # [4][pix]
                wh, wl, xh, xl = W4 >> 8, W4 & 255, X4 >> 8, X4 & 255
                hh = wh @ xh
                t = t + 256 * (wh @ xl + wl @ xh) + wl @ xl
                assert np.array_equal(65536 * hh + t - (1 << (so - 1)), W4 @ X4)   # the four byte-plane products recombine to P
                wsum += np.abs(W4).sum(axis=1)
            HH.append(hh); T.append(t)
        dsum = ((wsum * xmax) >> so) + 8
        dsum = np.where(dsum > 40000, 1 << 30, dsum)[:, None]
        ok = (U >= dsum) & (U <= 65535 - dsum)
        exact = U.copy()
        for s in range(7):
            exact = np.clip(exact + HH[s] * (1 << (16 - so)) + (T[s] >> so), 0, 65535)
        fast = U + (1 << (16 - so)) * sum(HH) + sum(ts >> so for ts in T)
        assert np.array_equal(fast[ok], exact[ok]), f"K-block {kb}: the bound let a saturating block through"
        n_fast += int(ok.sum()); n_exact += int((~ok).sum())
        U = np.where(ok, fast, exact)
    out = U - 32768
    out = np.where(out < 0, -((-out) // 10), out)
    assert np.array_equal(out.reshape(n, h, w).astype(np.int16), want)
    if amp == 2500:
        assert n_fast > 100 and n_exact > 100, (n_fast, n_exact)
    if amp == 32767:
        assert n_exact > n_fast          # (full-range operands: the bound exceeds the accumulator range, every block takes the exact step)
    else:
        assert n_fast > 0
    if amp == 600 and known_xmax:
        assert n_exact == 0
