"""TEST INFRASTRUCTURE ONLY - ctypes access to the checker libraries.

  Oracle()  -> oracle/_build/liboracle.so   our plain-C restatement (yolo2_oracle.c)
  Ref(prec) -> oracle/_ref/libref_{int16,fp32}.so  the unmodified reference, compiled by oracle/Makefile

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; nothing under yolo-fpga-accelerator_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "liboracle.so")
REF_SO = {"int16": os.path.join(HERE, "_ref", "libref_int16.so"), "fp32": os.path.join(HERE, "_ref", "libref_fp32.so")}

CONV, MAXPOOL, REORG, ROUTE, REGION = 0, 1, 2, 3, 4


def align8(w):
    return (w + 7) & ~7


class OrcLayer(C.Structure):
    _fields_ = [("type", C.c_int), ("c", C.c_int), ("h", C.c_int), ("w", C.c_int), ("out_c", C.c_int),
                ("out_h", C.c_int), ("out_w", C.c_int), ("n", C.c_int), ("size", C.c_int), ("stride", C.c_int),
                ("pad", C.c_int), ("leaky", C.c_int), ("batch_normalize", C.c_int), ("n_inputs", C.c_int),
                ("inputs", C.c_int * 4), ("classes", C.c_int), ("coords", C.c_int), ("softmax", C.c_int),
                ("background", C.c_int), ("anchors", C.c_float * 32)]


def build(ref=True, quiet=True):
    """Compiles the checker (and, when /root/reference exists, oracle/_ref). Building is not using."""
    out = subprocess.DEVNULL if quiet else None
    subprocess.check_call(["make", "-C", HERE, "oracle"], stdout=out)
    if ref and os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-C", HERE, "ref"], stdout=out)
        subprocess.check_call(["make", "-C", HERE, "ref-variants"], stdout=out)   # the reference built with --tn 8 / 16 / 32
        if os.path.exists(os.path.join(HERE, "..", "yolo-fpga-accelerator_b200", "lib", "libyolo2cuda.so")):
            # the reference's own CLI with the `--backend cuda` patch of INTEGRATION.md (tests/test_cli_backend_cuda.py)
            subprocess.check_call(["make", "-C", HERE, "ref-detect"], stdout=out)


def ref_so(precision="int16", tn=4):
    """library of the reference BUILT with rounding group tn (oracle/Makefile: `make ref TN=<tn>` / `make ref-variants`)"""
    return REF_SO[precision] if tn == 4 else REF_SO[precision].replace(".so", f"_tn{tn}.so")


def have_ref(precision="int16", tn=4):
    return os.path.exists(ref_so(precision, tn))


def _vp(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def layer_table(net):
    """yolo2_b200.cfg.Network -> orc_layer array."""
    arr = (OrcLayer * len(net.layers))()
    for d, l in zip(arr, net.layers):
        d.type = l.type
        d.c, d.h, d.w = l.c, l.h, l.w
        d.out_c, d.out_h, d.out_w = l.out_c, l.out_h, l.out_w
        d.n, d.size, d.stride, d.pad = l.n, l.size, l.stride, l.pad
        d.leaky, d.batch_normalize = l.leaky, l.batch_normalize
        d.n_inputs = len(l.inputs)
        for i, s in enumerate(l.inputs[:4]):
            d.inputs[i] = s
        d.classes, d.coords, d.softmax, d.background = l.classes, l.coords, l.softmax, l.background
        for i, a in enumerate(l.anchors[:32]):
            d.anchors[i] = a
    return arr


class Oracle:
    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        self.lib = C.CDLL(ORACLE_SO)
        self.lib.orc_round_shift.restype = C.c_int64
        self.lib.orc_round_shift.argtypes = [C.c_int64, C.c_int]

    def libm_exp(self, x):
        """the host libm's double exp, element-wise (NOT numpy's own SIMD exp)"""
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty_like(x)
        self.lib.orc_libm_exp(_vp(x), _vp(y), C.c_long(x.size))
        return y

    def letterbox_u8(self, img, net_w, net_h):
        """img: uint8 [h][w][c] (stb layout) -> float32 [c][net_h][net_w]"""
        img = np.ascontiguousarray(img, dtype=np.uint8)
        ih, iw, ic = img.shape
        out = np.empty((ic, net_h, net_w), np.float32)
        rc = self.lib.orc_letterbox_u8(_vp(img), iw, ih, ic, _vp(out), net_w, net_h)
        assert rc == 0, rc
        return out

    def set_tile_params(self, tn=4, tm=32):
        """reference build parameters used by net_forward (conv() takes TM/TN per call)"""
        self.lib.orc_set_tile_params(int(tn), int(tm))

    def round_shift(self, v, s):
        return int(self.lib.orc_round_shift(int(v), int(s)))

    def conv(self, x, w_reorg, bias, ifm, ofm, ksize, kstride, iw, ih, ow, oh, pad, is_nl, TM, TN,
             qw=0, qa_in=0, qa_out=0, qb=0, out=None):
        """x: [ifm][ih][align8 iw]; returns [ofm][oh][align8 ow] (pad columns keep `out`'s values)."""
        i16 = x.dtype == np.int16
        if out is None:
            out = np.zeros((ofm, oh, align8(ow)), x.dtype)
        if i16:
            rc = self.lib.orc_conv_i16(_vp(x), _vp(out), _vp(w_reorg), _vp(bias), ifm, ofm, ksize, kstride, iw, ih, ow,
                                       oh, pad, int(is_nl), TM, TN, qw, qa_in, qa_out, qb)
        else:
            rc = self.lib.orc_conv_f32(_vp(x), _vp(out), _vp(w_reorg), _vp(bias), ifm, ofm, ksize, kstride, iw, ih, ow,
                                       oh, pad, int(is_nl), TM, TN)
        assert rc == 0, rc
        return out

    def maxpool(self, x, ch, ksize, kstride, iw, ih, ow, oh, out=None):
        if out is None:
            out = np.zeros((ch, oh, align8(ow)), x.dtype)
        fn = self.lib.orc_maxpool_i16 if x.dtype == np.int16 else self.lib.orc_maxpool_f32
        rc = fn(_vp(x), _vp(out), ch, ksize, kstride, iw, ih, ow, oh)
        assert rc == 0, rc
        return out

    def reorg_hls(self, x, ch, TM, iw, ih, ow, oh, out=None):
        if out is None:
            out = np.zeros((ch, oh, align8(ow)), x.dtype)
        rc = self.lib.orc_reorg_hls_i16(_vp(x), _vp(out), ch, TM, iw, ih, ow, oh)
        assert rc == 0, rc
        return out

    def weight_reorg(self, w, ifm, ofm, ksize, Tm=32, Tn=4):
        w = np.ascontiguousarray(w)
        out = np.empty(ifm * ofm * ksize * ksize, w.dtype)
        self.lib.orc_weight_reorg(_vp(w), _vp(out), ifm, ofm, ksize, Tm, Tn, w.dtype.itemsize)
        return out

    def quantize_input(self, x, q):
        x = np.ascontiguousarray(x, np.float32)
        out = np.empty(x.size, np.int16)
        self.lib.orc_quantize_input(_vp(x), _vp(out), C.c_size_t(x.size), q)
        return out.reshape(x.shape)

    def reorg_driver(self, x, c, h, w, shift=0):
        out = np.zeros((4 * c, h // 2, align8(w // 2)), x.dtype)
        if x.dtype == np.int16:
            self.lib.orc_reorg_driver_i16(_vp(x), _vp(out), c, h, w, shift)
        else:
            self.lib.orc_reorg_driver_f32(_vp(x), _vp(out), c, h, w)
        return out

    def region_from_ofm(self, ofm, w, h, n, classes, coords=4, softmax=1, background=0, q=0):
        """ofm: [ch][h][align8 w] int16 or float -> region tensor [n][coords+1+classes][h][w]."""
        ch = n * (coords + 1 + classes)
        rf = np.empty(ch * h * w, np.float32)
        if ofm.dtype == np.int16:
            self.lib.orc_region_strip_dequant_i16(_vp(ofm), _vp(rf), ch, h, w, q)
        else:
            self.lib.orc_region_strip_f32(_vp(ofm), _vp(rf), ch, h, w)
        return self.region_forward(rf, w, h, n, classes, coords, softmax, background)

    def region_forward(self, rf, w, h, n, classes, coords=4, softmax=1, background=0):
        rf = np.ascontiguousarray(rf, np.float32)
        out = np.empty(rf.size, np.float32)
        self.lib.orc_region_forward(_vp(rf), _vp(out), w, h, n, classes, coords, softmax, background)
        return out.reshape(n, coords + 1 + classes, h, w)

    def region_boxes_nms(self, region, lw, lh, n, classes, anchors, im_w, im_h, net_w, net_h, thresh, nms):
        region = np.ascontiguousarray(region, np.float32)
        total = lw * lh * n
        boxes = np.zeros((total, 4), np.float32)
        probs = np.zeros((total, classes), np.float32)
        obj = np.zeros(total, np.float32)
        anchors = np.asarray(anchors, np.float32)
        self.lib.orc_region_boxes_nms(_vp(region), lw, lh, n, classes, _vp(anchors), im_w, im_h, net_w, net_h,
                                      C.c_float(thresh), C.c_float(nms), _vp(boxes), _vp(probs), _vp(obj))
        return boxes, probs, obj

    def net_forward(self, net, frame, pack, dump_layers=False):
        """Generalised yolov2_hls_ps. Returns (region [n][..][h][w], {layer index: ofm}) for one frame."""
        table = layer_table(net)
        L = len(net.layers)
        frame = np.ascontiguousarray(frame, np.float32)
        last = net.layers[-1]
        region = np.zeros(last.c * last.h * last.w, np.float32)
        dt = np.int16 if pack.is_int16 else np.float32
        dumps, ptrs = {}, (C.c_void_p * L)()
        if dump_layers:
            for i, l in enumerate(net.layers):
                if l.type in (CONV, MAXPOOL, REORG):
                    dumps[i] = np.zeros((l.out_c, l.out_h, align8(l.out_w)), dt)
                    ptrs[i] = dumps[i].ctypes.data
        dp = C.cast(ptrs, C.c_void_p) if dump_layers else None
        if pack.is_int16:
            rc = self.lib.orc_net_forward_i16(table, L, _vp(frame), _vp(pack.weights), _vp(pack.bias),
                                              _vp(np.ascontiguousarray(pack.weight_q, np.int32)),
                                              _vp(np.ascontiguousarray(pack.bias_q, np.int32)),
                                              _vp(np.ascontiguousarray(pack.act_q, np.int32)), len(pack.act_q), dp,
                                              _vp(region))
        else:
            rc = self.lib.orc_net_forward_f32(table, L, _vp(frame), _vp(pack.weights), _vp(pack.bias), dp, _vp(region))
        assert rc == 0, rc
        return region.reshape(last.n, last.coords + 1 + last.classes, last.h, last.w), dumps


class Ref:
    """The real reference (compiled unmodified). precision: "int16" | "fp32"."""

    def __init__(self, precision="int16", tn=4):
        self.precision = precision
        self.dtype = np.int16 if precision == "int16" else np.float32
        if not have_ref(precision, tn):
            raise FileNotFoundError(ref_so(precision, tn) + " (run `make -C oracle ref [ref-variants]` where /root/reference exists)")
        self.lib = C.CDLL(ref_so(precision, tn))
        assert self.lib.ref_precision_bits() == (16 if precision == "int16" else 32)
        t = (C.c_int * 5)()
        self.lib.ref_tile_params(C.byref(t, 0), C.byref(t, 4), C.byref(t, 8), C.byref(t, 12), C.byref(t, 16))
        assert t[0] == tn, (t[0], tn)
        self.tn, self.tm = t[0], t[1]

    def letterbox_u8(self, img, net_w, net_h):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        ih, iw, ic = img.shape
        out = np.empty((ic, net_h, net_w), np.float32)
        assert self.lib.ref_letterbox_u8(_vp(img), iw, ih, ic, _vp(out), net_w, net_h) == 0
        return out

    def yolo2_fpga(self, Input, Output, Weight, Beta, IFM_num, OFM_num, Ksize, Kstride, Input_w, Input_h, Output_w,
                   Output_h, Padding, IsNL, IsBN, TM, TN, TR, TC, OFM_num_bound, mLoopsxTM, mLoops_a1xTM, LayerType,
                   Qw=0, Qa_in=0, Qa_out=0, Qb=0):
        """Unmodified YOLO2_FPGA. Input needs >= 4096 elements of slack on both sides: the loader reads
        whole 8-element beats around the tile (yolo2_model.cpp:243-244 keeps 512 each side)."""
        self.lib.ref_yolo2_fpga(_vp(Input), _vp(Output), _vp(Weight), _vp(Beta), IFM_num, OFM_num, Ksize, Kstride,
                                Input_w, Input_h, Output_w, Output_h, Padding, int(IsNL), int(IsBN), TM, TN, TR, TC,
                                OFM_num_bound, mLoopsxTM, mLoops_a1xTM, LayerType, Qw, Qa_in, Qa_out, Qb)

    def run_layer(self, x, w, b, args, q=(0, 0, 0, 0)):
        """Convenience: pads the input with slack, calls YOLO2_FPGA with `args` (dict from
        yolo2_b200.accel.conv_call_args / pool_call_args) and returns [OFM][OH][align8 OW]."""
        slack = 8192
        buf = np.zeros(x.size + 2 * slack, self.dtype)
        buf[slack:slack + x.size] = x.reshape(-1)
        out = np.zeros((args["OFM_num"], args["Output_h"], align8(args["Output_w"])), self.dtype)
        wpad = None
        if w is not None:
            wpad = np.zeros(w.size + 4096, self.dtype)
            wpad[:w.size] = w
        inp = buf[slack:]
        self.yolo2_fpga(inp, out, wpad, b, *[args[k] for k in (
            "IFM_num", "OFM_num", "Ksize", "Kstride", "Input_w", "Input_h", "Output_w", "Output_h", "Padding", "IsNL",
            "IsBN", "TM", "TN", "TR", "TC", "OFM_num_bound", "mLoopsxTM", "mLoops_a1xTM", "LayerType")], *q)
        return out

    def region_forward(self, rf, w, h, n, classes, coords=4, softmax=1, background=0):
        rf = np.ascontiguousarray(rf, np.float32)
        out = np.zeros(rf.size, np.float32)
        self.lib.ref_region_forward(_vp(rf), _vp(out), w, h, n, classes, coords, softmax, background)
        return out.reshape(n, coords + 1 + classes, h, w)

    def region_boxes_nms(self, region, lw, lh, n, classes, anchors, im_w, im_h, net_w, net_h, thresh, nms):
        region = np.ascontiguousarray(region, np.float32)
        total = lw * lh * n
        boxes = np.zeros((total, 4), np.float32)
        probs = np.zeros((total, classes), np.float32)
        obj = np.zeros(total, np.float32)
        anchors = np.ascontiguousarray(anchors, np.float32)
        k = self.lib.ref_region_boxes_nms(_vp(region), lw, lh, n, classes, _vp(anchors), im_w, im_h, net_w, net_h,
                                          C.c_float(thresh), C.c_float(nms), _vp(boxes), _vp(probs), _vp(obj), total)
        assert k == total, k
        return boxes, probs, obj

    def full_forward(self, cfg_path, frame, workdir, im_w, im_h, thresh=0.25, nms=0.45, classes=80, total=845,
                     region_len=71825):
        """load_network + yolov2_hls_ps + get_network_boxes + do_nms_sort, run with cwd=workdir
        (which must hold weights/*.bin). COCO 416 only (SURVEY.md finding 2)."""
        frame = np.ascontiguousarray(frame, np.float32)
        region = np.zeros(region_len, np.float32)
        boxes = np.zeros((total, 4), np.float32)
        probs = np.zeros((total, classes), np.float32)
        obj = np.zeros(total, np.float32)
        cfgp = os.path.abspath(cfg_path).encode()
        cwd = os.getcwd()
        os.chdir(workdir)
        try:
            k = self.lib.ref_full_forward(cfgp, _vp(frame), im_w, im_h, C.c_float(thresh), C.c_float(0.5), C.c_float(nms),
                                          _vp(region), region_len, _vp(boxes), _vp(probs), _vp(obj), total)
        finally:
            os.chdir(cwd)
        assert k >= 0, k
        return region, boxes, probs, obj, k
