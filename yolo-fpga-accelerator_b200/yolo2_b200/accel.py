"""YOLO2_FPGA drop-in: one accelerator call with the reference's argument contract.

Reference interface: hls/models/yolov2/yolo2_accel.hpp:10-17 (defined yolo2_accel.cpp:25-171).
Same names, order and meaning; bad arguments raise Yolo2CudaError(YOLO2CUDA_ERROR) where the
reference asserts (yolo2_accel.cpp:75-87).
"""
import ctypes as C
import math

import numpy as np

from . import _capi


class Accelerator:
    """Owns one yolo2cuda context (device, stream, scratch)."""

    def __init__(self, device: int = 0, precision: str = "int16"):
        self.lib = _capi.load_library()
        self.precision = precision
        self.dtype = np.int16 if precision == "int16" else np.float32
        self.ctx = C.c_void_p()
        rc = self.lib.yolo2cuda_create(C.byref(self.ctx), device, 16 if precision == "int16" else 32)
        if rc != _capi.SUCCESS:
            raise _capi.Yolo2CudaError(rc, "no usable B200 (sm_100) CUDA device - there is no CPU fallback")

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.yolo2cuda_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_tile_params(self, tn: int = 4, tm: int = 32):
        """Emulate a reference BUILT with another rounding group Tn / weight block Tm (scripts/hw_params_gen.py --tn/--tm)."""
        _capi.check(self.ctx, self.lib.yolo2cuda_set_tile_params(self.ctx, tn, tm))
        self.tn, self.tm = tn, tm

    def use_torch_stream(self, device=None):
        """Launch on torch's current stream.  torch's default stream has handle 0, which yolo2cuda_set_stream reads as "back to
        the context's own (non-blocking) stream"; the legacy default stream is addressed by CUDA's cudaStreamLegacy handle (1)."""
        import torch
        self.set_stream(torch.cuda.current_stream(device).cuda_stream or 1)

    def set_stream(self, cuda_stream_ptr):
        _capi.check(self.ctx, self.lib.yolo2cuda_set_stream(self.ctx, C.c_void_p(cuda_stream_ptr)))

    def synchronize(self):
        _capi.check(self.ctx, self.lib.yolo2cuda_synchronize(self.ctx))

    @property
    def launch_count(self):
        return int(self.lib.yolo2cuda_launch_count(self.ctx))

    def tc_path_counts(self, reset: bool = False):
        """(fast, exact): (warp, tile) units of the tcgen05 conv kernel through its no-saturation fast path / its exact step"""
        f, x = C.c_uint64(0), C.c_uint64(0)
        _capi.check(self.ctx, self.lib.yolo2cuda_tc_path_counts(self.ctx, C.byref(f), C.byref(x), int(reset)))
        return int(f.value), int(x.value)

    @property
    def last_kernel(self):
        return self.lib.yolo2cuda_last_kernel(self.ctx).decode()

    # -- the accelerator call ---------------------------------------------------------------
    def YOLO2_FPGA(self, Input, Output, Weight, Beta, IFM_num, OFM_num, Ksize, Kstride, Input_w, Input_h,
                   Output_w, Output_h, Padding, IsNL, IsBN, TM, TN, TR, TC, OFM_num_bound, mLoopsxTM,
                   mLoops_a1xTM, LayerType, Qw=0, Qa_in=0, Qa_out=0, Qb=0):
        """Input/Output/Weight/Beta: contiguous numpy arrays of the context dtype in the reference
        layouts (Weight/Beta None for pool). Output is written in place, pad columns untouched."""
        def ptr(a, need):
            if a is None:
                return None
            if a.dtype != self.dtype or not a.flags["C_CONTIGUOUS"]:
                raise _capi.Yolo2CudaError(_capi.ERROR, "buffers must be C-contiguous arrays of the context dtype")
            if a.size < need:
                raise _capi.Yolo2CudaError(_capi.ERROR, f"buffer too small: {a.size} < {need}")
            return a.ctypes.data_as(C.c_void_p)
        al = lambda w: (w + 7) & ~7
        safe = all(isinstance(v, (int, np.integer)) and 0 < v <= 2048 for v in (IFM_num, OFM_num, Input_w, Input_h, Output_w, Output_h, Ksize))
        need_in = IFM_num * Input_h * al(Input_w) if safe else 0
        need_out = OFM_num * Output_h * al(Output_w) if safe else 0
        need_w = IFM_num * OFM_num * Ksize * Ksize if safe else 0
        rc = self.lib.yolo2cuda_layer_host(self.ctx, ptr(Input, need_in), ptr(Output, need_out), ptr(Weight, need_w),
                                           ptr(Beta, OFM_num if safe else 0), IFM_num, OFM_num, Ksize, Kstride, Input_w,
                                           Input_h, Output_w, Output_h, Padding, int(IsNL), int(IsBN), TM, TN, TR, TC,
                                           OFM_num_bound, mLoopsxTM, mLoops_a1xTM, LayerType, Qw, Qa_in, Qa_out, Qb)
        _capi.check(self.ctx, rc)

    def YOLO2_FPGA_dev(self, Input, Output, Weight, Beta, *scalars):
        """Same call on device memory: arguments are raw device pointers (ints), e.g. torch
        tensor.data_ptr(); asynchronous on the context stream."""
        p = lambda v: C.c_void_p(v) if v else None
        rc = self.lib.yolo2cuda_layer_dev(self.ctx, p(Input), p(Output), p(Weight), p(Beta), *[int(s) for s in scalars])
        _capi.check(self.ctx, rc)


# -- how the reference driver fills the tile/pipeline arguments (yolo2_model.cpp:299-355) ---------
Tn, Tm, Tr, Tc, OnChipIB = 4, 32, 13, 13, 27


def letterbox_image(acc: "Accelerator", images, net_w: int, net_h: int):
    """GPU image front-end (reference: load_image_stb + letterbox_image, src/core/yolo_image.cpp:84-187).
    images: torch.uint8 CUDA tensor [batch][h][w][c] (stb layout) -> torch.float32 CUDA tensor [batch][c][net_h][net_w]."""
    import torch
    assert images.is_cuda and images.dtype == torch.uint8 and images.dim() == 4 and images.is_contiguous()
    b, ih, iw, ic = images.shape
    out = torch.empty((b, ic, net_h, net_w), dtype=torch.float32, device=images.device)
    acc.use_torch_stream(images.device)
    _capi.check(acc.ctx, acc.lib.yolo2cuda_letterbox_dev(acc.ctx, C.c_void_p(images.data_ptr()), b, iw, ih, ic, C.c_void_p(out.data_ptr()), net_w, net_h))
    return out


def conv_call_args(c, n, size, stride, w, h, pad, leaky, bn=0, tn=Tn, tm=Tm):
    """tn / tm: the reference build's tile parameters (scripts/hw_params_gen.py --tn/--tm); defaults are its defaults"""
    ow = (w - size + 2 * pad) // stride + 1
    oh = (h - size + 2 * pad) // stride + 1
    TR = min((OnChipIB - size) // stride + 1, Tr, oh)
    TC = min((OnChipIB - size) // stride + 1, Tc, ow)
    TM, TN = min(n, tm), min(c, tn)
    mLoops = math.ceil(n / TM)
    return dict(IFM_num=c, OFM_num=n, Ksize=size, Kstride=stride, Input_w=w, Input_h=h, Output_w=ow, Output_h=oh,
                Padding=pad, IsNL=leaky, IsBN=bn, TM=TM, TN=TN, TR=TR, TC=TC, OFM_num_bound=(mLoops + 1) * TM,
                mLoopsxTM=mLoops * TM, mLoops_a1xTM=(mLoops + 1) * TM, LayerType=0)


def pool_call_args(c, size, stride, w, h, out_w, out_h, pad, tn=Tn, tm=Tm):
    TR = min((OnChipIB - size) // stride + 1, Tr, out_h)
    TC = min((OnChipIB - size) // stride + 1, Tc, out_w)
    TM = min(tm, tn, c)
    mLoops = math.ceil(c / TM)
    return dict(IFM_num=c, OFM_num=c, Ksize=size, Kstride=stride, Input_w=w, Input_h=h, Output_w=out_w, Output_h=out_h,
                Padding=pad, IsNL=0, IsBN=0, TM=TM, TN=0, TR=TR, TC=TC, OFM_num_bound=(mLoops + 2) * TM,
                mLoopsxTM=mLoops * TM, mLoops_a1xTM=(mLoops + 1) * TM, LayerType=1)
