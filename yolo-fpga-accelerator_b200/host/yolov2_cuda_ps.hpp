// Model-level drop-in for the reference host: same signature and side effect as
//   void yolov2_hls_ps(network *net, const float *input, Precision precision);
// (hls/models/yolov2/yolo2_accel.hpp:21-23, defined hls/models/yolov2/yolo2_model.cpp:229-449):
// runs every layer of `net` on the B200 through the C ABI of include/yolo2cuda.h and leaves the
// region tensor in net->layers[net->n-1].output.  Reads the same weights/*.bin files from the cwd.
// Compiled against the reference's own headers (include/core/yolo.h) - see INTEGRATION.md.
#pragma once

struct network;
enum class Precision;

void yolov2_cuda_ps(network *net, const float *input, Precision precision);
// Batched form used by `--batch N`: frames = N letterboxed images back to back; out = N region tensors.
void yolov2_cuda_ps_batch(network *net, const float *frames, int batch, float *region_out, Precision precision);
