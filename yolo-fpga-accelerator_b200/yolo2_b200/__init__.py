"""B200-native YOLOv2 accelerator datapath (drop-in for the reference's YOLO2_FPGA path).

Host-side mirror of the reference interface for this path:
  accel.YOLO2_FPGA        <-> hls/models/yolov2/yolo2_accel.hpp:10-17   (one accelerator call)
  model.yolov2_cuda_ps    <-> hls/models/yolov2/yolo2_accel.hpp:21-23   (yolov2_hls_ps)
  cfg.parse_network_cfg   <-> src/core/yolo_net.cpp:218-291             (layer table only)
  weights.*               <-> yolo2_model.cpp:158-227, yolov2_weight_gen.cpp:34-68
  accel.letterbox_image   <-> src/core/yolo_image.cpp:84-187            (load_image_stb conversion + letterbox_image, on the GPU)
  model.region_detections[_gpu] <-> src/core/yolo_region.cpp:169-236, yolo_post.cpp:54-85 (boxes + NMS, host / GPU)
  accel.Accelerator.set_tile_params <-> scripts/hw_params_gen.py --tn/--tm (the reference's build parameters)
  convert.*               <-> weights/README.md:37-57                   (darknet .weights -> the accelerator's files; replaces the
                                                                         un-vendored nn-weight-extractor step)

Everything computes through lib/libyolo2cuda.so (hand-written sm_100a CUDA behind the C ABI of
include/yolo2cuda.h).  There is no CPU fallback: importing works anywhere, computing without
the library or without a B200 raises.
"""
from . import cfg, weights  # noqa: F401
from ._capi import Yolo2CudaError, lib_path, load_library  # noqa: F401
