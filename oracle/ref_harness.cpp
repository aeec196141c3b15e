// TEST INFRASTRUCTURE ONLY - never linked into, or called by, the product path.
//
// Thin extern "C" wrapper around the UNMODIFIED reference sources, compiled where they
// lie under /root/reference by oracle/Makefile into oracle/_ref/libref_{int16,fp32}.so.
// Nothing here restates reference logic: every entry point forwards to a reference
// symbol so that tests can call the real thing through ctypes.
//
//   ref_yolo2_fpga      -> YOLO2_FPGA            hls/models/yolov2/yolo2_accel.cpp:25-171
//   ref_region_forward  -> forward_region_layer  src/core/yolo_region.cpp:123-141
//   ref_letterbox_u8    -> letterbox_image       src/core/yolo_image.cpp:146-165 (+ the u8->float loop of :178-187)
//   ref_full_forward    -> load_network + yolov2_hls_ps + get_network_boxes + do_nms_sort
//                          (the call sequence of src/models/yolov2/yolov2_main.cpp:255-325)
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <vector>

#include <core/yolo.h>
#include <core/precision.hpp>
#include <api.hpp>

extern "C" {

// 16 for the -DINT16_MODE build, 32 for the float build (IO_Dtype, hls/core/types.hpp:8-14).
int ref_precision_bits(void) { return (int)(8 * sizeof(IO_Dtype)); }

// Tile constants the reference was generated with (hls/core/params.hpp via scripts/hw_params_gen.py).
void ref_tile_params(int *tn, int *tm, int *tr, int *tc, int *ib)
{
    *tn = Tn; *tm = Tm; *tr = Tr; *tc = Tc; *ib = OnChipIB_Width;
}

// One accelerator call; arguments in the order of hls/models/yolov2/yolo2_accel.hpp:10-17.
void ref_yolo2_fpga(void *Input, void *Output, void *Weight, void *Beta,
                    int IFM_num, int OFM_num, int Ksize, int Kstride,
                    int Input_w, int Input_h, int Output_w, int Output_h,
                    int Padding, int IsNL, int IsBN,
                    int TM, int TN, int TR, int TC,
                    int OFM_num_bound, int mLoopsxTM, int mLoops_a1xTM, int LayerType,
                    int Qw, int Qa_in, int Qa_out, int Qb)
{
    YOLO2_FPGA((IO_Dtype *)Input, (IO_Dtype *)Output, (IO_Dtype *)Weight, (IO_Dtype *)Beta,
               IFM_num, OFM_num, Ksize, Kstride, Input_w, Input_h, Output_w, Output_h,
               Padding, IsNL != 0, IsBN != 0, TM, TN, TR, TC,
               OFM_num_bound, mLoopsxTM, mLoops_a1xTM, LayerType, Qw, Qa_in, Qa_out, Qb);
}

// Region layer forward on a caller-supplied [n*(coords+1+classes)][h][w] float tensor.
void ref_region_forward(const float *in, float *out, int w, int h, int n, int classes,
                        int coords, int softmax, int background)
{
    layer l;
    std::memset(&l, 0, sizeof(l));
    l.type = REGION;
    l.n = n; l.batch = 1; l.w = w; l.h = h;
    l.c = n * (classes + coords + 1);
    l.out_w = w; l.out_h = h; l.out_c = l.c;
    l.classes = classes; l.coords = coords;
    l.outputs = h * w * l.c; l.inputs = l.outputs;
    l.softmax = softmax; l.background = background;
    l.output = out;
    std::vector<float> tmp(in, in + l.outputs);
    forward_region_layer(l, tmp.data());
}

// Whole reference pipeline for one letterboxed frame. The process cwd must hold weights/*.bin
// (yolo2_model.cpp:171-193 uses relative paths). Returns the number of candidate boxes
// (w*h*n) or <0 on error. `region_out` receives layers[n-1].output; boxes/probs/objectness
// receive the post-NMS detection table in reference order.
int ref_full_forward(const char *cfg_path, const float *input, int im_w, int im_h,
                     float thresh, float hier, float nms,
                     float *region_out, int region_cap,
                     float *boxes, float *probs, float *objectness, int box_cap)
{
    try {
        network *net = load_network(const_cast<char *>(cfg_path));
        if (!net) return -1;
        set_batch_network(net, 1);
#ifdef INT16_MODE
        yolov2_hls_ps(net, input, Precision::INT16);
#else
        yolov2_hls_ps(net, input, Precision::FP32);
#endif
        layer last = net->layers[net->n - 1];
        if (region_out) {
            int cnt = last.outputs < region_cap ? last.outputs : region_cap;
            std::memcpy(region_out, last.output, sizeof(float) * cnt);
        }
        int nboxes = 0;
        detection *dets = get_network_boxes(net, im_w, im_h, thresh, hier, 0, 1, &nboxes);
        if (!dets) return -2;
        if (nms > 0.0f) do_nms_sort(dets, nboxes, last.classes, nms);
        for (int i = 0; i < nboxes && i < box_cap; ++i) {
            boxes[4 * i + 0] = dets[i].bbox.x;
            boxes[4 * i + 1] = dets[i].bbox.y;
            boxes[4 * i + 2] = dets[i].bbox.w;
            boxes[4 * i + 3] = dets[i].bbox.h;
            objectness[i] = dets[i].objectness;
            for (int j = 0; j < last.classes; ++j) probs[(size_t)i * last.classes + j] = dets[i].prob[j];
        }
        free_detections(dets, nboxes);
        return nboxes;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "ref_full_forward: %s\n", e.what());
        return -3;
    }
}

// Box decode + NMS only, on a region tensor the caller already has (the tail of
// yolov2_main.cpp:311-320 with a single REGION layer network).
int ref_region_boxes_nms(const float *region, int lw, int lh, int n, int classes,
                         const float *anchors, int im_w, int im_h, int net_w, int net_h,
                         float thresh, float nms,
                         float *boxes, float *probs, float *objectness, int box_cap)
{
    layer l;
    std::memset(&l, 0, sizeof(l));
    l.type = REGION;
    l.n = n; l.batch = 1; l.w = lw; l.h = lh;
    l.classes = classes; l.coords = 4;
    l.c = n * (classes + 5);
    l.outputs = lw * lh * l.c; l.inputs = l.outputs;
    l.output = const_cast<float *>(region);
    l.biases = const_cast<float *>(anchors);
    network net;
    std::memset(&net, 0, sizeof(net));
    net.n = 1; net.layers = &l; net.w = net_w; net.h = net_h;
    int nboxes = 0;
    detection *dets = get_network_boxes(&net, im_w, im_h, thresh, 0.5f, 0, 1, &nboxes);
    if (!dets) return -1;
    if (nms > 0.0f) do_nms_sort(dets, nboxes, classes, nms);
    for (int i = 0; i < nboxes && i < box_cap; ++i) {
        boxes[4 * i + 0] = dets[i].bbox.x;
        boxes[4 * i + 1] = dets[i].bbox.y;
        boxes[4 * i + 2] = dets[i].bbox.w;
        boxes[4 * i + 3] = dets[i].bbox.h;
        objectness[i] = dets[i].objectness;
        for (int j = 0; j < classes; ++j) probs[(size_t)i * classes + j] = dets[i].prob[j];
    }
    free_detections(dets, nboxes);
    return nboxes;
}

// The reference's own letterbox_image / resize_image (src/core/yolo_image.cpp:84-165) on an in-memory stb-layout image.
// Only the u8 -> float loop of load_image_stb (:178-187, it reads a file) is repeated here to build the `image`.
int ref_letterbox_u8(const unsigned char *hwc, int iw, int ih, int ic, float *out, int net_w, int net_h)
{
    image im = make_image(iw, ih, ic);
    for (int k = 0; k < ic; ++k)
        for (int j = 0; j < ih; ++j)
            for (int i = 0; i < iw; ++i) im.data[i + iw * j + iw * ih * k] = (float)hwc[k + ic * i + ic * iw * j] / 255.;
    image boxed = letterbox_image(im, net_w, net_h);
    memcpy(out, boxed.data, sizeof(float) * (size_t)net_w * net_h * ic);
    free_image(im);
    free_image(boxed);
    return 0;
}

} // extern "C"
