/*
 * TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference's YOLOv2 accelerator datapath, used as the parity
 * checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs.  Nothing under yolo-fpga-accelerator_b200/ may include, link or call it.
 *
 * Parity status: PINNED.  The reference ships no golden vectors, so the
 * restatement is pinned against the reference ITSELF: oracle/Makefile compiles the unmodified
 * reference sources into oracle/_ref/libref_{int16,fp32}.so, tests/test_oracle_vs_ref.py
 * checks every function below against it bit for bit, and oracle/gen_golden.py freezes
 * reference outputs into tests/golden/ for machines where /root/reference is absent.
 *
 * All feature maps use the reference layout: planar [C][H][ceil8(W)]   (yolo2_accel.cpp:89-99).
 * All citations are relative to /root/reference.
 */
#ifndef YOLO2_ORACLE_H
#define YOLO2_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_CONV = 0, ORC_MAXPOOL = 1, ORC_REORG = 2, ORC_ROUTE = 3, ORC_REGION = 4 };

/* One cfg section, already shape-propagated (the fields of `layer` the reference driver reads,
 * include/core/yolo.h + yolo2_model.cpp:294-446). */
typedef struct orc_layer {
    int type;
    int c, h, w;              /* input dims  */
    int out_c, out_h, out_w;  /* output dims */
    int n;                    /* conv filters / region anchors */
    int size, stride, pad;
    int leaky;                /* activation==LEAKY */
    int batch_normalize;
    int n_inputs;             /* route */
    int inputs[4];            /* absolute layer indices */
    int classes, coords, softmax, background;
    float anchors[32];
} orc_layer;

int orc_align8(int w);

/* rs(v,s) of SURVEY §2.3 (core_compute.cpp:49-62,86-94,108-113). */
int64_t orc_round_shift(int64_t v, int shift);

/* LayerType 0, INT16_MODE (core_compute.cpp:32-120 + :175-210, scheduler core_scheduler.cpp:33-62,
 * loaders core_io.cpp:82-199).  Weights in reorganised order for tile (TM, TN). */
int orc_conv_i16(const int16_t *in, int16_t *out, const int16_t *w_reorg, const int16_t *bias,
                 int ifm, int ofm, int ksize, int kstride, int iw, int ih, int ow, int oh,
                 int pad, int is_nl, int TM, int TN, int qw, int qa_in, int qa_out, int qb);

/* LayerType 0, float build (core_compute.cpp:121-172, :200-204). */
int orc_conv_f32(const float *in, float *out, const float *w_reorg, const float *bias,
                 int ifm, int ofm, int ksize, int kstride, int iw, int ih, int ow, int oh,
                 int pad, int is_nl, int TM, int TN);

/* LayerType 1 (core_compute.cpp:266-305, loader pad value core_io.cpp:96-103, Padding forced to 0
 * core_scheduler.cpp:72-73). */
int orc_maxpool_i16(const int16_t *in, int16_t *out, int ch, int ksize, int kstride,
                    int iw, int ih, int ow, int oh);
int orc_maxpool_f32(const float *in, float *out, int ch, int ksize, int kstride,
                    int iw, int ih, int ow, int oh);

/* LayerType 2 (core_compute.cpp:354-379): out[m+2ky+kx][y][x] = in[m][2y+ky][2x+kx], m stepping by TM. */
int orc_reorg_hls_i16(const int16_t *in, int16_t *out, int ch, int TM, int iw, int ih, int ow, int oh);

/* Offline weight reorder, darknet [ofm][ifm][kh][kw] -> accelerator order
 * (src/models/yolov2/yolov2_weight_gen.cpp:34-68). elem = bytes per element (2 or 4). */
void orc_weight_reorg(const void *w, void *w_reorg, int ifm, int ofm, int ksize, int Tm, int Tn, int elem);

/* Input quantiser (yolo2_model.cpp:257-273). */
void orc_quantize_input(const float *in, int16_t *out, size_t count, int q_in);

/* Driver-side reorg + Q alignment (yolo2_model.cpp:112-129, :358-401), generalised from the
 * hard-wired 26x26x64.  in: [c][h][ceil8 w]; out: [4c][h/2][ceil8(w/2)] with zeroed pad columns.
 * shift >= 0 is the arithmetic right shift applied afterwards (0 = none). */
void orc_reorg_driver_i16(const int16_t *in, int16_t *out, int c, int h, int w, int shift);
void orc_reorg_driver_f32(const float *in, float *out, int c, int h, int w);

/* Region head: strip padding + dequantise (yolo2_model.cpp:406-425) then forward_region_layer
 * (src/core/yolo_region.cpp:123-141, src/core/yolo_math.cpp:19,226-250).
 * in_f: compact [n*(coords+1+classes)][h][w] floats. */
void orc_region_strip_dequant_i16(const int16_t *in, float *out, int ch, int h, int w, int q);
void orc_region_strip_f32(const float *in, float *out, int ch, int h, int w);
void orc_region_forward(const float *in_f, float *out, int w, int h, int n, int classes,
                        int coords, int softmax, int background);

/* get_region_detections + correct_region_boxes + do_nms_sort
 * (src/core/yolo_region.cpp:18-53,169-195, src/core/yolo_post.cpp:7-85).
 * Outputs: boxes[w*h*n][4] (x,y,w,h relative), probs[w*h*n][classes], objectness[w*h*n];
 * returns w*h*n. */
int orc_region_boxes_nms(const float *region, int lw, int lh, int n, int classes,
                         const float *anchors, int im_w, int im_h, int net_w, int net_h,
                         float thresh, float nms, float *boxes, float *probs, float *objectness);

/* Generalised restatement of yolov2_hls_ps (yolo2_model.cpp:229-446): dims come from the layer
 * table, weights/bias are the reorganised blobs with layers back to back (no per-layer pad
 * element), Q tables as in the int16 files.  dump[i], when non-NULL, receives layer i's ofm in
 * the reference layout.  region_out receives layers[n-1].output. */
int orc_net_forward_i16(const orc_layer *layers, int n_layers, const float *frame,
                        const int16_t *w_reorg, const int16_t *bias,
                        const int32_t *weight_q, const int32_t *bias_q, const int32_t *act_q, int n_act_q,
                        int16_t **dump, float *region_out);
int orc_net_forward_f32(const orc_layer *layers, int n_layers, const float *frame,
                        const float *w_reorg, const float *bias, float **dump, float *region_out);

/* stb u8 [ih][iw][ic] image -> float [ic][net_h][net_w] letterboxed network input (yolo_image.cpp:84-165,178-187) */
void orc_libm_exp(const double *x, double *y, long n);
int orc_letterbox_u8(const unsigned char *hwc, int iw, int ih, int ic, float *out, int net_w, int net_h);

/* reference build parameters Tn (rounding group) / Tm used by orc_net_forward_* (defaults 4 / 32) */
void orc_set_tile_params(int tn, int tm);

#ifdef __cplusplus
}
#endif
#endif
