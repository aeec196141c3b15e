/*
 * yolo2cuda.h - C ABI of the B200 (sm_100a) YOLOv2 accelerator datapath.
 *
 * Drop-in boundary for the reference's accelerator entry YOLO2_FPGA
 * (reference: hls/api.hpp:3 -> hls/models/yolov2/yolo2_accel.hpp:10-17, defined in
 * hls/models/yolov2/yolo2_accel.cpp:25-171) and for the model-level driver yolov2_hls_ps
 * (hls/models/yolov2/yolo2_accel.hpp:21-23, defined in hls/models/yolov2/yolo2_model.cpp:229-449).
 *
 * Plain C: pointers, sizes and ints only.  Every entry point returns one of the
 * YOLO2CUDA_* codes (styled after the board driver's YOLO2_SUCCESS.. codes,
 * linux_app/include/yolo2_config.h:146-151); nothing asserts or aborts.  There is no CPU
 * fallback: without a CUDA device every compute entry returns YOLO2CUDA_INIT_ERROR.
 *
 * Thread safety: a yolo2cuda_ctx and the networks created from it are single-threaded objects (one host thread at a
 * time per context); different contexts - e.g. one per GPU - may be used from different threads concurrently.  The
 * library keeps no mutable process-global state (YOLO2_FPGA itself is not re-entrant: function-local statics,
 * yolo2_accel.cpp:103-113).
 *
 * Data contracts shared with the reference:
 *   feature maps   planar [C][H][ceil8(W)], int16 or float   (yolo2_accel.cpp:89-99)
 *   weights        reorganised order produced by yolov2_weight_gen (yolov2_weight_gen.cpp:34-68)
 *   Q values       fractional-bit counts, real = int * 2^-Q      (yolo2_model.cpp:261-262)
 */
#ifndef YOLO2CUDA_H
#define YOLO2CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YOLO2CUDA_SUCCESS         0
#define YOLO2CUDA_ERROR          -1   /* invalid argument (what the reference asserts on) */
#define YOLO2CUDA_TIMEOUT        -2   /* kept for code compatibility; never produced */
#define YOLO2CUDA_INIT_ERROR     -3   /* no CUDA device / context creation failed */
#define YOLO2CUDA_MEMORY_ERROR   -4   /* device or host allocation failed */
#define YOLO2CUDA_LAUNCH_ERROR   -5   /* a CUDA call or kernel launch failed */

#define YOLO2CUDA_PRECISION_INT16 16  /* reference build with -DINT16_MODE (hls/core/types.hpp:8-11) */
#define YOLO2CUDA_PRECISION_FP32  32  /* reference float build              (hls/core/types.hpp:12-14) */

/* LayerType values of YOLO2_FPGA (yolo2_accel.cpp:140-146, core_scheduler.cpp:33,63,88). */
#define YOLO2CUDA_LAYER_CONV    0
#define YOLO2CUDA_LAYER_MAXPOOL 1
#define YOLO2CUDA_LAYER_REORG   2

/* Hardware tile limits the reference asserts against (generated hls/core/params.hpp,
 * scripts/hw_params_gen.py:16-23): used only for argument validation. */
#define YOLO2CUDA_Tn 4
#define YOLO2CUDA_Tm 32
#define YOLO2CUDA_Tr 13
#define YOLO2CUDA_Tc 13

typedef struct yolo2cuda_ctx yolo2cuda_ctx;   /* owns the device, stream and scratch memory   */
typedef struct yolo2cuda_net yolo2cuda_net;   /* one loaded network: plan, weights, arena     */

/* ---- context ----------------------------------------------------------------------------- */

/* Creates a context on CUDA device `device` for `precision` (16 or 32). */
int yolo2cuda_create(yolo2cuda_ctx **ctx, int device, int precision);
int yolo2cuda_destroy(yolo2cuda_ctx *ctx);
/* Tile parameters of the reference BUILD this context emulates (scripts/hw_params_gen.py --tn/--tm ->
 * hls/core/params.hpp): Tn = input-channel tile = the ROUNDING GROUP of the int16 accumulator
 * (core_scheduler.cpp:45, core_compute.cpp:65-120), Tm = output-channel weight block.  Defaults 4 / 32 (the
 * reference's); results are bit-exact to a reference compiled with the same values.  Set before net_create. */
int yolo2cuda_set_tile_params(yolo2cuda_ctx *ctx, int Tn, int Tm);
/* Launch all work of this context on `cuda_stream` (a cudaStream_t).  NULL = back to the context's own non-blocking stream;
 * the legacy default stream (handle 0 in most frameworks) is addressed as cudaStreamLegacy, i.e. (void *)1. */
int yolo2cuda_set_stream(yolo2cuda_ctx *ctx, void *cuda_stream);
int yolo2cuda_synchronize(yolo2cuda_ctx *ctx);
/* Text of the last error on this context (never NULL). */
const char *yolo2cuda_last_error(const yolo2cuda_ctx *ctx);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
uint64_t yolo2cuda_launch_count(const yolo2cuda_ctx *ctx);
/* Name of the kernel variant chosen by the last conv launch (diagnostics / tests). */
const char *yolo2cuda_last_kernel(const yolo2cuda_ctx *ctx);
/* Diagnostics of the tcgen05 conv kernel (csrc/conv_i16_tc2.cu): how many (warp, tile) units went through its no-saturation fast
 * path and how many through the exact step, since context creation or the last reset.  Both paths produce the reference's
 * bits; the environment variable YOLO2CUDA_TC_EXACT=1 (read at yolo2cuda_create) disables the fast path.  Synchronises. */
int yolo2cuda_tc_path_counts(yolo2cuda_ctx *ctx, uint64_t *fast_tiles, uint64_t *exact_tiles, int reset);

/* ---- one accelerator call: replaces YOLO2_FPGA --------------------------------------------
 * Argument order, meaning and limits are those of yolo2_accel.hpp:10-17 / yolo2_accel.cpp:75-87.
 * IO element type is int16_t or float according to the context precision.
 *   - TM/TR/TC and the three pipeline bounds do not change results; they are validated
 *     (bounds against ceil(OFM/TM)) and otherwise ignored.
 *   - TN is the number of input channels summed before each round-and-saturate step
 *     (core_scheduler.cpp:45, core_compute.cpp:96-118) and the weight block width.
 *   - LayerType 1 ignores Weight/Beta/Q and Padding (core_scheduler.cpp:72-73).
 * The _host form takes HOST pointers exactly like the reference call (copies in, runs,
 * copies out, synchronises).  The _dev form takes DEVICE pointers in the same layouts and
 * is asynchronous on the context stream. */
int yolo2cuda_layer_host(yolo2cuda_ctx *ctx, const void *Input, void *Output, const void *Weight,
                         const void *Beta, int IFM_num, int OFM_num, int Ksize, int Kstride,
                         int Input_w, int Input_h, int Output_w, int Output_h, int Padding,
                         int IsNL, int IsBN, int TM, int TN, int TR, int TC, int OFM_num_bound,
                         int mLoopsxTM, int mLoops_a1xTM, int LayerType,
                         int Qw, int Qa_in, int Qa_out, int Qb);
int yolo2cuda_layer_dev(yolo2cuda_ctx *ctx, const void *Input, void *Output, const void *Weight,
                        const void *Beta, int IFM_num, int OFM_num, int Ksize, int Kstride,
                        int Input_w, int Input_h, int Output_w, int Output_h, int Padding,
                        int IsNL, int IsBN, int TM, int TN, int TR, int TC, int OFM_num_bound,
                        int mLoopsxTM, int mLoops_a1xTM, int LayerType,
                        int Qw, int Qa_in, int Qa_out, int Qb);

/* ---- driver-side operators the reference runs on the host CPU (device pointers) -----------
 * quantize : yolo2_model.cpp:257-273      float[count] -> int16[count]
 * reorg    : yolo2_model.cpp:112-129,358-401  [c][h][ceil8 w] -> [4c][h/2][ceil8(w/2)], then >> shift
 * region   : yolo2_model.cpp:406-425 + src/core/yolo_region.cpp:123-141
 *            [n*(coords+1+classes)][h][ceil8 w] -> float [n][coords+1+classes][h][w] */
int yolo2cuda_quantize_input_dev(yolo2cuda_ctx *ctx, const float *in, int16_t *out, size_t count, int q_in);
int yolo2cuda_reorg_dev(yolo2cuda_ctx *ctx, const void *in, void *out, int c, int h, int w, int shift);
int yolo2cuda_region_dev(yolo2cuda_ctx *ctx, const void *in, float *out, int w, int h, int n,
                         int classes, int coords, int softmax, int background, int q);

/* ---- whole network: replaces yolov2_hls_ps ------------------------------------------------ */

enum { YOLO2CUDA_CONV = 0, YOLO2CUDA_MAXPOOL = 1, YOLO2CUDA_REORG = 2, YOLO2CUDA_ROUTE = 3, YOLO2CUDA_REGION = 4 };

/* One cfg section after shape propagation: the fields of the reference's `layer`
 * (include/core/yolo.h) that yolov2_hls_ps reads (yolo2_model.cpp:294-446). */
typedef struct yolo2cuda_layer_desc {
    int32_t type;
    int32_t c, h, w;               /* input dims  */
    int32_t out_c, out_h, out_w;   /* output dims */
    int32_t n;                     /* conv filters / region anchors */
    int32_t size, stride, pad;
    int32_t leaky;                 /* activation == LEAKY */
    int32_t batch_normalize;
    int32_t n_inputs;              /* route */
    int32_t inputs[4];             /* absolute layer indices */
    int32_t classes, coords, softmax, background;
    float   anchors[32];
} yolo2cuda_layer_desc;

/* Builds the execution plan and the device arena for up to `max_batch` frames per call. */
int yolo2cuda_net_create(yolo2cuda_ctx *ctx, const yolo2cuda_layer_desc *layers, int n_layers,
                         int max_batch, yolo2cuda_net **net);
int yolo2cuda_net_destroy(yolo2cuda_net *net);

/* Uploads the reference weight files' contents (HOST pointers):
 *   weights/bias : reorganised blobs, conv layers back to back, NO per-layer pad element
 *                  (the loader strips it, yolo2_model.cpp:216-223); int16_t or float
 *   weight_q/bias_q : >= n_conv int32 entries; act_q : >= n_conv+1 entries (int16 only) */
int yolo2cuda_net_load_weights(yolo2cuda_net *net, const void *weights, size_t n_weights,
                               const void *bias, size_t n_bias,
                               const int32_t *weight_q, const int32_t *bias_q, int n_q,
                               const int32_t *act_q, int n_act_q);

/* frames: float [batch][c][h][w] letterboxed images in [0,1] (what yolov2_hls_ps receives);
 * region_out: float [batch][outputs of the last layer] = net->layers[n-1].output per frame.
 * _host: HOST pointers, copies + sync inside.  _dev: DEVICE pointers, async on the stream. */
int yolo2cuda_net_forward_host(yolo2cuda_net *net, const float *frames, int batch, float *region_out);
int yolo2cuda_net_forward_dev(yolo2cuda_net *net, const float *frames, int batch, float *region_out);

/* Image front-end on the GPU: what load_image_stb + letterbox_image do on the host in the reference
 * (src/core/yolo_image.cpp:178-187 u8 -> float/255, :84-127 two-pass bilinear resize, :146-165 letterbox with 0.5 fill),
 * bit-exact.  src: u8 [batch][ih][iw][ic] interleaved (stbi_load's layout), every image the same size;
 * dst: float [batch][ic][net_h][net_w].  DEVICE pointers, async on the stream. */
int yolo2cuda_letterbox_dev(yolo2cuda_ctx *ctx, const unsigned char *src, int batch, int iw, int ih, int ic,
                            float *dst, int net_w, int net_h);
/* yolo2cuda_net_forward_host fed with raw images: HOST u8 [batch][ih][iw][in_c] -> letterbox on the GPU -> network. */
int yolo2cuda_net_forward_images_host(yolo2cuda_net *net, const unsigned char *images, int batch, int iw, int ih,
                                      float *region_out);

/* Copies layer `layer`'s output feature map of frame `frame` (from the last forward) to the
 * HOST buffer `dst` in the reference layout [out_c][out_h][ceil8(out_w)] (int16_t or float).
 * Needs yolo2cuda_net_set_debug_keep(net, 1) BEFORE the forward. */
int yolo2cuda_net_get_layer_output(yolo2cuda_net *net, int layer, int frame, void *dst, size_t dst_elems);
/* Activation Q of the tensor entering the region layer after the last forward (int16). */
int yolo2cuda_net_region_q(const yolo2cuda_net *net);
/* Kernels launched by one forward of `batch` frames (0 before the first forward). */
uint64_t yolo2cuda_net_launches_per_forward(const yolo2cuda_net *net);
/* Pass schedule of yolo2cuda_net_forward_host for batches larger than max_batch (forward_images_host runs its passes back to back): the upload of the first
 * pass is the only one that does not overlap compute, so such a batch starts with a short RAMP pass of `frames` frames followed by
 * passes of max_batch (-1 = default max(32, max_batch / 6); 0 = no ramp pass).  The results do not depend on the schedule. */
int yolo2cuda_net_set_ramp_frames(yolo2cuda_net *net, int frames);
/* Activation memory.  Default (keep = 0): ONE device arena in which a tensor occupies its bytes only between its first
 * writer and its last reader, so buffers are recycled down the network like the reference's ping-pong scratch arena
 * (yolo2_model.cpp:56-110); the route source (layer 16) and the concat buffer stay alive across the layers between their
 * writers and readers.  keep = 1: every tensor owns its memory, so yolo2cuda_net_get_layer_output works for EVERY layer of
 * the last forward (tests, per-layer dumps); get_layer_output fails in the default mode.  Switching re-places the tensors
 * (allowed at any time between forwards; the weights stay loaded). */
int yolo2cuda_net_set_debug_keep(yolo2cuda_net *net, int keep);
/* Bytes of device memory the activation tensors occupy in the current mode (arena size, or the sum over layers). */
size_t yolo2cuda_net_activation_bytes(const yolo2cuda_net *net);
/* Per-layer device time in ms of the last forward (CUDA events; enables them on first use). */
int yolo2cuda_net_layer_times(yolo2cuda_net *net, float *ms, int n_layers);
/* Name of the kernel variant the last forward launched for `layer` ("" for layers that launch nothing, e.g. route;
 * never NULL).  Diagnostics / tests: lets a parity test assert WHICH conv kernel produced the bits it compared. */
const char *yolo2cuda_net_layer_kernel(const yolo2cuda_net *net, int layer);

/* Diagnostics: y[i] = exp(x[i]) exactly as the region kernel computes the logistic / softmax exponentials of
 * src/core/yolo_math.cpp:19,234 - glibc's double exp algorithm restated on the device (sysdeps/ieee754/dbl-64/e_exp.c, FMA form),
 * bit-identical to the libm of an x86-64 host with FMA.  DEVICE pointers, asynchronous on the context stream. */
int yolo2cuda_selftest_exp_dev(yolo2cuda_ctx *ctx, const double *x, double *y, size_t n);

/* ---- detections: get_network_boxes + do_nms_sort on the host (src/core/yolo_region.cpp:169-236,
 * src/core/yolo_post.cpp:54-85).  region: one frame's region tensor (HOST).  The output arrays must hold
 * lw*lh*n entries (the worst case); the function fills the first K of them, K = the number of candidates with
 * objectness > thresh, in the reference's candidate scan order (cell-major, then anchor - the order of its list before
 * do_nms_sort's qsort re-orders it), and RETURNS K (or a negative YOLO2CUDA_* code).  boxes [K][4] = x,y,w,h relative to
 * the original image; probs [K][classes] after per-class NMS (suppressed entries 0); objectness [K].  The set of surviving
 * (box, class, probability) equals the reference's, including ties (same libc qsort on an order carried across classes). */
int yolo2cuda_region_detections(const float *region, int lw, int lh, int n, int classes,
                                const float *anchors, int im_w, int im_h, int net_w, int net_h,
                                float thresh, float nms, float *boxes, float *probs, float *objectness);

/* The same on the GPU for a batch of region tensors that are still on the device: DEVICE pointers,
 * async on the stream.  region: float [batch][n*(5+classes)*lw*lh]; outputs per frame boxes [lw*lh*n][4], probs
 * [lw*lh*n][classes], objectness [lw*lh*n].  POSITIONAL: entry cell*n + anchor (the order in which the reference fills its
 * candidate list); entries at or below the objectness threshold are zero.  The reference's list is compacted and re-ordered by
 * qsort; the SET of surviving (box, class, probability) is bit-identical.  lw*lh*n <= 1024. */
int yolo2cuda_region_detections_dev(yolo2cuda_ctx *ctx, const float *region, int batch, int lw, int lh, int n, int classes,
                                    const float *anchors_host, int im_w, int im_h, int net_w, int net_h,
                                    float thresh, float nms, float *boxes, float *probs, float *objectness);

/* Compacts the positional output of yolo2cuda_region_detections_dev into fixed-size records per frame (DEVICE pointers, async):
 * records [batch][cap][8] 32-bit words = {entry (cell * n + anchor), class, probability, x, y, w, h, objectness} (the floats as
 * their bit patterns), ordered by (entry, class); counts [batch] = surviving (entry, class) pairs of the frame (the first `cap`
 * are stored).  This is what a multi-GPU host gathers instead of the 287 KB region tensors. */
int yolo2cuda_compact_detections_dev(yolo2cuda_ctx *ctx, const float *boxes, const float *probs, const float *objectness, int batch,
                                     int total, int classes, int cap, uint32_t *records, int32_t *counts);

#ifdef __cplusplus
}
#endif
#endif /* YOLO2CUDA_H */
