// Shared declarations of the B200 YOLOv2 datapath kernels (sm_100a only).
//
// Device-side data layouts
//   planar : the reference's [C][H][ceil8(W)]                      (yolo2_accel.cpp:89-99)
//   C4     : [frame][G=ceil(C/4)][H][W][4] - the four channels of one reference rounding group
//            (Tn=4, core_scheduler.cpp:45) sit in one 8-byte (int16) / 16-byte (float) word, so
//            one load feeds exactly one round-and-saturate step.  Channels >= C are zero.
//   weights: [ceil(OFM/16)][G][K*K][16] of {lo-bytes x4, hi-bytes x4} (int16) or float4 (fp32),
//            zero padded; built from the reference's reorganised blob by wprep_* kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace y2 {

__host__ __device__ inline int align8(int w) { return (w + 7) & ~7; }
__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

constexpr int kCM = 16;  // output channels per CTA / per weight block

// Element offset of w[m][c][tap] inside one layer of the reference's reorganised blob
// (producer yolov2_weight_gen.cpp:43-66, consumer core_io.cpp:154-198).
__host__ __device__ inline size_t reorg_woff(int m, int c, int tap, int ifm, int ofm, int k2, int TM, int TN)
{
    int m0 = (m / TM) * TM, n0 = (c / TN) * TN;
    int tmm = min(TM, ofm - m0), tnn = min(TN, ifm - n0);
    return (size_t)m0 * ifm * k2 + (size_t)tmm * n0 * k2 + ((size_t)tap * tmm + (m - m0)) * tnn + (c - n0);
}

// rs(v,s) of core_compute.cpp:49-62,86-94,108-113 in 64-bit (generic path and bias init).
__host__ __device__ inline long long round_shift64(long long v, int shift)
{
    if (shift > 0) {
        int mag = shift > 30 ? 30 : shift;
        return (v + (1LL << (mag - 1))) >> mag;
    }
    if (shift < 0) {
        int mag = -shift > 30 ? 30 : -shift;
        return (long long)((unsigned long long)v << mag);
    }
    return v;
}

struct ConvFastParams {
    const void *in;        // C4 input, frame 0 / group 0
    void *out;             // C4 output, already offset to the first output group
    const void *w;         // device weight layout
    const void *bias;      // [OFM] int16 / float
    int B, H, W;           // frames and spatial dims (stride-1 "same" conv: out dims == in dims)
    int G;                 // input groups = ceil(IFM/4)
    int OFM;
    long long in_frame_stride;   // elements between frames
    long long out_frame_stride;  // elements between frames
    int TP;                // pixels per thread segment (13 or 7)
    int SW;                // segments per row
    int RB;                // image rows per CTA band
    int PW;                // smem row pitch in pixels
    int GS;                // groups per pipeline stage
    int so, sb;            // effective shift_out in [8,30]; raw shift_bias
    int leaky;
    // largest |value| bookkeeping for the tcgen05 kernel's no-saturation fast path (int16 only; both may be NULL):
    const int *xmax_in;    // device scalar: upper bound of |x| over the input tensor (NULL = unknown, 32768 is assumed)
    int *xmax_out;         // device scalar the kernel atomicMax-es the largest |output| into
    unsigned long long *tc_stats;   // device [2]: warp-tiles through the fast / the exact path of the tcgen05 kernel (NULL = not counted)
    int tc_force_exact;    // tests: the tcgen05 kernel never takes the fast path
    int group_words;       // C4 words per rounding group of the CUDA-core int16 kernel: 0 / 1 = Tn 4 (default), 2 = Tn 8, 4 = Tn 16
};

// ---- launchers (defined in the .cu files; all asynchronous on `st`) ---------------------------
// Returns the number of kernels launched, or -1 when the shape is not eligible.
int launch_conv_i16_fast(const ConvFastParams &p, int ksize, cudaStream_t st, const char **variant);
int launch_conv_f32_fast(const ConvFastParams &p, int ksize, cudaStream_t st, const char **variant);
// 3x3 conv of a layer with one input channel group (IFM <= 4), optionally with the following 2x2 / stride-2 max-pool fused (csrc/conv_i16_g1.cu)
int launch_conv_i16_g1(const ConvFastParams &p, int ksize, int pool, cudaStream_t st, const char **variant);
size_t conv_fast_plan(ConvFastParams &p, int ksize, int elem_bytes);  // fills SW/RB/PW/GS, returns smem bytes (0 = not eligible)

void launch_wprep_i16(const int16_t *blob, void *dst, int ifm, int ofm, int ksize, int TM, int TN, cudaStream_t st);
void launch_wprep_f32(const float *blob, void *dst, int ifm, int ofm, int ksize, int TM, int TN, cudaStream_t st);
size_t wprep_bytes(int ifm, int ofm, int ksize, int elem_bytes);

void launch_conv_i16_generic(const int16_t *in, int16_t *out, const int16_t *w, const int16_t *bias,
                             int ifm, int ofm, int ksize, int kstride, int iw, int ih, int ow, int oh,
                             int pad, int is_nl, int TM, int TN, int so, int sb, cudaStream_t st);
void launch_conv_f32_generic(const float *in, float *out, const float *w, const float *bias,
                             int ifm, int ofm, int ksize, int kstride, int iw, int ih, int ow, int oh,
                             int pad, int is_nl, int TM, int TN, cudaStream_t st);

// planar <-> C4 (elem_bytes 2 or 4); frames laid out back to back with the given strides (elements)
void launch_planar_to_c4(const void *src, void *dst, int B, int C, int H, int W, long long src_frame_stride,
                         long long dst_frame_stride, int elem_bytes, cudaStream_t st);
void launch_c4_to_planar(const void *src, void *dst, int B, int C, int H, int W, long long src_frame_stride,
                         long long dst_frame_stride, int elem_bytes, cudaStream_t st);

void launch_maxpool_planar(const void *in, void *out, int ch, int ksize, int kstride, int iw, int ih,
                           int ow, int oh, int elem_bytes, cudaStream_t st);
void launch_maxpool_c4(const void *in, void *out, int B, int G, int kstride, int iw, int ih, int ow, int oh,
                       long long in_frame_stride, long long out_frame_stride, int elem_bytes, cudaStream_t st);
void launch_reorg_hls_planar(const void *in, void *out, int ch, int TM, int iw, int ih, int ow, int oh,
                             int elem_bytes, cudaStream_t st);

void launch_quantize(const float *in, int16_t *out, size_t count, int q_in, cudaStream_t st);
// frames float [B][C][H][W] -> C4 (quantised when elem_bytes==2)
void launch_frames_to_c4(const float *frames, void *dst, int B, int C, int H, int W, long long dst_frame_stride,
                         int q_in, int elem_bytes, cudaStream_t st);
// boxes + per-class NMS on the GPU (yolo_region.cpp:15-53,169-195, yolo_post.cpp:22-85); positional output, see bw_ops.cu
int launch_detect(const float *region, float *boxes, float *probs, float *objectness, int B, int lw, int lh, int n, int classes,
                  const float *anchors, int im_w, int im_h, int net_w, int net_h, float thresh, float nms, const double *expf_tab,
                  cudaStream_t st);
// positional detect output -> [B][cap][8] records + [B] counts (see bw_ops.cu)
void launch_compact_detections(const float *boxes, const float *probs, const float *objectness, int B, int total, int classes, int cap,
                               unsigned *records, int *counts, cudaStream_t st);
// stb u8 [B][ih][iw][ic] -> float [B][ic][net_h][net_w], darknet letterbox (yolo_image.cpp:84-165,178-187), bit-exact
// glibc's double exp as the region kernel computes it, element-wise on device arrays (diagnostics / tests)
void launch_glibc_exp(const double *x, double *y, long long n, cudaStream_t st);
void launch_letterbox(const unsigned char *src, float *dst, int B, int iw, int ih, int ic, int net_w, int net_h, cudaStream_t st);
void launch_reorg_driver_planar(const void *in, void *out, int c, int h, int w, int shift, int elem_bytes, cudaStream_t st);
void launch_reorg_driver_c4(const void *in, void *out, int B, int c, int h, int w, int shift,
                            long long in_frame_stride, long long out_frame_stride, int elem_bytes, cudaStream_t st);
// layout: 0 = planar [ch][h][ceil8 w] single frame, 1 = C4 batch
void launch_region(const void *in, float *out, int B, int w, int h, int n, int classes, int coords,
                   int softmax, int background, int q, int layout, long long in_frame_stride,
                   int elem_bytes, cudaStream_t st);

}  // namespace y2

namespace y2 {
// tensor-core (tcgen05 + TMEM) int16 conv, csrc/conv_i16_tc2.cu
size_t wprep_tc2_bytes(int ifm, int ofm, int ksize);
void launch_wprep_tc2(const int16_t *blob, void *dst, int ifm, int ofm, int ksize, int TM, int TN, int so, cudaStream_t st);
int launch_conv_i16_tc2(const ConvFastParams &p, int ksize, cudaStream_t st, const char **variant);
bool conv_i16_tc2_eligible(const ConvFastParams &p, int ksize, int frames);   // shape / alignment rules of the launcher, without launching
// tensor-core conv for a reference built with rounding group tn = 32, 16 or 8 (one MMA K slice = 32 / tn chain steps), csrc/conv_i16_tc32.cu
size_t wprep_tc32_bytes(int ifm, int ofm, int ksize, int tn);
void launch_wprep_tc32(const int16_t *blob, void *dst, int ifm, int ofm, int ksize, int TM, int TN, int tn, cudaStream_t st);
int launch_conv_i16_tc32(const ConvFastParams &p, int ksize, int ifm, int tn, cudaStream_t st, const char **variant);
int conv_i16_tc32_slices(int ifm, int ksize, int tn);
bool conv_i16_tc32_eligible(int W, int ksize, int so, int tn);   // shape rules of the launcher, without launching
}  // namespace y2
