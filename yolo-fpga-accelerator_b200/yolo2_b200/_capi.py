"""ctypes binding of include/yolo2cuda.h (the C ABI a cgo/JNI/C++ host would bind the same way)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SUCCESS, ERROR, TIMEOUT, INIT_ERROR, MEMORY_ERROR, LAUNCH_ERROR = 0, -1, -2, -3, -4, -5
PRECISION_INT16, PRECISION_FP32 = 16, 32
CONV, MAXPOOL, REORG, ROUTE, REGION = 0, 1, 2, 3, 4


class Yolo2CudaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"yolo2cuda error {code}: {msg}")
        self.code = code


class LayerDesc(C.Structure):
    """struct yolo2cuda_layer_desc"""
    _fields_ = [("type", C.c_int32), ("c", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
                ("out_c", C.c_int32), ("out_h", C.c_int32), ("out_w", C.c_int32), ("n", C.c_int32),
                ("size", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32), ("leaky", C.c_int32),
                ("batch_normalize", C.c_int32), ("n_inputs", C.c_int32), ("inputs", C.c_int32 * 4),
                ("classes", C.c_int32), ("coords", C.c_int32), ("softmax", C.c_int32),
                ("background", C.c_int32), ("anchors", C.c_float * 32)]


def lib_path():
    # YOLO2CUDA_LIB: development override (e.g. a build with -DY2_TC2_PROFILE); still a CUDA library, never a fallback
    return os.environ.get("YOLO2CUDA_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libyolo2cuda.so")


_LAYER_ARGS = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 23

# every symbol include/yolo2cuda.h declares: (restype, argtypes)
SYMBOLS = {
    "yolo2cuda_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int]),
    "yolo2cuda_destroy": (C.c_int, [C.c_void_p]),
    "yolo2cuda_set_tile_params": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "yolo2cuda_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "yolo2cuda_synchronize": (C.c_int, [C.c_void_p]),
    "yolo2cuda_last_error": (C.c_char_p, [C.c_void_p]),
    "yolo2cuda_launch_count": (C.c_uint64, [C.c_void_p]),
    "yolo2cuda_last_kernel": (C.c_char_p, [C.c_void_p]),
    "yolo2cuda_tc_path_counts": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_int]),
    "yolo2cuda_layer_host": (C.c_int, _LAYER_ARGS),
    "yolo2cuda_layer_dev": (C.c_int, _LAYER_ARGS),
    "yolo2cuda_quantize_input_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]),
    "yolo2cuda_reorg_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    "yolo2cuda_region_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 8),
    "yolo2cuda_net_create": (C.c_int, [C.c_void_p, C.POINTER(LayerDesc), C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "yolo2cuda_net_destroy": (C.c_int, [C.c_void_p]),
    "yolo2cuda_net_load_weights": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                             C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "yolo2cuda_net_forward_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "yolo2cuda_net_forward_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "yolo2cuda_selftest_exp_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "yolo2cuda_letterbox_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]),
    "yolo2cuda_net_forward_images_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "yolo2cuda_net_get_layer_output": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]),
    "yolo2cuda_net_region_q": (C.c_int, [C.c_void_p]),
    "yolo2cuda_net_launches_per_forward": (C.c_uint64, [C.c_void_p]),
    "yolo2cuda_net_set_debug_keep": (C.c_int, [C.c_void_p, C.c_int]),
    "yolo2cuda_net_set_ramp_frames": (C.c_int, [C.c_void_p, C.c_int]),
    "yolo2cuda_net_activation_bytes": (C.c_size_t, [C.c_void_p]),
    "yolo2cuda_net_layer_times": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "yolo2cuda_net_layer_kernel": (C.c_char_p, [C.c_void_p, C.c_int]),
    "yolo2cuda_region_detections": (C.c_int, [C.c_void_p] + [C.c_int] * 4 + [C.c_void_p] + [C.c_int] * 4 +
                                    [C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "yolo2cuda_region_detections_dev": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 5 + [C.c_void_p] + [C.c_int] * 4 +
                                        [C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "yolo2cuda_compact_detections_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                                   C.c_void_p, C.c_void_p]),
}


def load_library():
    """Loads lib/libyolo2cuda.so; fails loudly when it was not built (no fallback exists)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise Yolo2CudaError(INIT_ERROR, f"{path} is missing - run `make -C yolo-fpga-accelerator_b200/csrc` "
                                         "(or __graft_entry__.build()); there is no CPU fallback")
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(ctx, rc):
    if rc != SUCCESS:
        msg = load_library().yolo2cuda_last_error(ctx).decode() if ctx else "context creation failed"
        raise Yolo2CudaError(rc, msg)
