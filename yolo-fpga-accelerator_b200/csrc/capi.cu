// C ABI of the B200 YOLOv2 datapath (include/yolo2cuda.h): context, the YOLO2_FPGA drop-in
// (yolo2cuda_layer_*), the driver-side operators and the whole-network executor that replaces
// yolov2_hls_ps (hls/models/yolov2/yolo2_model.cpp:229-449).  No CPU fallback anywhere: every
// compute entry needs a CUDA device and fails with a YOLO2CUDA_* code otherwise.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/yolo2cuda.h"
#include "common.cuh"

using namespace y2;

struct yolo2cuda_ctx {
    int device = 0;
    int precision = 16;
    int elem = 2;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    std::string err = "ok";
    uint64_t launches = 0;
    const char *last_kernel = "";
    int force_generic = 0;
    int Tn = YOLO2CUDA_Tn, Tm = YOLO2CUDA_Tm;   // tile parameters of the emulated reference build (yolo2cuda_set_tile_params)
    int use_tc = -1;  // YOLO2CUDA_TC: unset = auto (the tcgen05 kernel csrc/conv_i16_tc2.cu on the layers where it measured faster),
                      // 0 = CUDA-core kernels only, anything else = the tcgen05 kernel wherever the shape is eligible (tests / profiling)
    int tc_min_ofm = 96;
    int use_g1 = 1;                  // YOLO2CUDA_G1=0: the generic C4 kernel (and a separate pool launch) for one-group 3x3 layers (A/B measurements)
    int tc_force_exact = 0;          // YOLO2CUDA_TC_EXACT=1: the tcgen05 kernel never takes its no-saturation fast path (tests)
    unsigned long long *d_tc_stats = nullptr;   // device [2]: warp-tiles of the tcgen05 kernel through the fast / the exact path
    // growable device scratch for the per-layer entry points
    struct Scratch { void *p = nullptr; size_t bytes = 0; } s_in, s_out, s_w, s_b, s_c4in, s_c4out, s_wprep;
};

namespace {

int fail(yolo2cuda_ctx *ctx, int code, const char *fmt, ...)
{
    if (ctx) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof(buf), fmt, ap);
        va_end(ap);
        ctx->err = buf;
    }
    return code;
}

#define CUDA_OK(ctx, call)                                                                            \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? YOLO2CUDA_MEMORY_ERROR : YOLO2CUDA_LAUNCH_ERROR, \
                        "%s failed: %s", #call, cudaGetErrorString(e_));                              \
    } while (0)

int ensure(yolo2cuda_ctx *ctx, yolo2cuda_ctx::Scratch &s, size_t bytes)
{
    if (s.bytes >= bytes) return YOLO2CUDA_SUCCESS;
    if (s.p) {
        cudaStreamSynchronize(ctx->stream);
        cudaFree(s.p);
        s.p = nullptr;
        s.bytes = 0;
    }
    size_t want = bytes + (bytes >> 2) + 4096;
    CUDA_OK(ctx, cudaMalloc(&s.p, want));
    s.bytes = want;
    return YOLO2CUDA_SUCCESS;
}

size_t planar_elems(int c, int h, int w) { return (size_t)c * h * align8(w); }
size_t c4_elems(int c, int h, int w) { return (size_t)ceil_div(c, 4) * h * w * 4; }

// The checks of yolo2_accel.cpp:75-87 (and the board driver's validate_conv_params,
// linux_app/src/yolo2_accel_linux.c:383-414), returned as an error instead of assert().
const char *validate_layer_args(int IFM, int OFM, int K, int S, int Iw, int Ih, int Ow, int Oh, int Pad, int TM, int TN,
                                int TR, int TC, int bound, int mLxTM, int mLa1xTM, int type, int Tn_build = YOLO2CUDA_Tn,
                                int Tm_build = YOLO2CUDA_Tm)
{
    if (OFM <= 0 || OFM > 2048) return "OFM_num out of (0,2048]";
    if (IFM <= 0 || IFM > 2048) return "IFM_num out of (0,2048]";
    if (S <= 0 || S > 2) return "Kstride out of (0,2]";
    if (K <= 0 || K > 3) return "Ksize out of (0,3]";
    if (Iw <= 0 || Iw > 1024 || Ih <= 0 || Ih > 1024) return "input dims out of (0,1024]";
    if (Ow <= 0 || Ow > 1024 || Oh <= 0 || Oh > 1024) return "output dims out of (0,1024]";
    if (Pad < 0 || Pad > 4) return "Padding out of [0,4]";
    if (TM <= 0 || TM > Tm_build) return "TM out of (0,Tm]";
    if (TN < 0 || TN > Tn_build) return "TN out of [0,Tn]";
    if (TR <= 0 || TR > YOLO2CUDA_Tr) return "TR out of (0,Tr]";
    if (TC <= 0 || TC > YOLO2CUDA_Tc) return "TC out of (0,Tc]";
    if (type < 0 || type > 2) return "LayerType must be 0, 1 or 2";
    // software-pipeline drain bounds (yolo2_accel.cpp:136-146): the only values for which the
    // reference writes every output tile exactly once
    const int mLoops = ceil_div(OFM, TM);
    if (mLxTM != mLoops * TM) return "mLoopsxTM != ceil(OFM/TM)*TM";
    if (type == 0) {
        if (bound != (mLoops + 1) * TM) return "OFM_num_bound != (mLoops+1)*TM for conv";
        if (TN <= 0) return "TN must be > 0 for conv";
    } else {
        if (bound != (mLoops + 2) * TM) return "OFM_num_bound != (mLoops+2)*TM for pool/reorg";
        if (mLa1xTM != (mLoops + 1) * TM) return "mLoops_a1xTM != (mLoops+1)*TM for pool/reorg";
    }
    return nullptr;
}

bool fast_shift_ok(int so) { return so >= 8; }  // effective shift = min(so,30), see conv_i16.cu

int run_conv_layer_dev(yolo2cuda_ctx *ctx, const void *Input, void *Output, const void *Weight, const void *Beta,
                       int IFM, int OFM, int K, int S, int Iw, int Ih, int Ow, int Oh, int Pad, int IsNL, int TM,
                       int TN, int Qw, int Qa_in, int Qa_out, int Qb)
{
    cudaStream_t st = ctx->stream;
    const int so = Qa_in + Qw - Qa_out, sb = Qb - Qa_out;
    // the C4 kernels' rounding group is 4 channels (or all of them when IFM <= 4); a context emulating a reference built with
    // Tn = 8 / 16 (yolo2cuda_set_tile_params) groups 2 / 4 C4 words per step in the CUDA-core int16 kernel
    const int gwords = (ctx->elem == 2 && (ctx->Tn == 8 || ctx->Tn == 16) && IFM > 4 && TN == (IFM < ctx->Tn ? IFM : ctx->Tn) && so <= 22) ? ctx->Tn / 4 : 1;
    bool fast = !ctx->force_generic && (K == 1 || K == 3) && S == 1 && Pad == K / 2 && Ow == Iw && Oh == Ih &&
                (TN == 4 || (IFM <= TN && IFM <= 4) || gwords > 1);
    if (ctx->elem == 2) fast = fast && fast_shift_ok(so);
    ConvFastParams p{};
    p.group_words = gwords;
    // a reference built with Tn = 32: one MMA K slice is one rounding group (csrc/conv_i16_tc32.cu); Tn = 16 / 8: two / four
    // rounding groups per K slice (block-diagonal operand) - in this single-frame entry only when YOLO2CUDA_TC forces it, like the
    // Tn = 4 tensor-core kernel (the network executor decides per layer)
    const bool tcn_shape = !ctx->force_generic && ctx->elem == 2 && IFM > 4 && TN == (IFM < ctx->Tn ? IFM : ctx->Tn) && (K == 1 || K == 3) &&
                           S == 1 && Pad == K / 2 && Ow == Iw && Oh == Ih && so >= 8 && so <= 16;
    const bool tc32 = tcn_shape && ((ctx->Tn == 32 && !fast) || ((ctx->Tn == 16 || ctx->Tn == 8) && ctx->use_tc > 0));
    if (tc32) {
        int rc;
        if ((rc = ensure(ctx, ctx->s_c4in, c4_elems(IFM, Ih, Iw) * ctx->elem))) return rc;
        if ((rc = ensure(ctx, ctx->s_c4out, c4_elems(OFM, Oh, Ow) * ctx->elem))) return rc;
        if ((rc = ensure(ctx, ctx->s_wprep, wprep_tc32_bytes(IFM, OFM, K, ctx->Tn)))) return rc;
        launch_planar_to_c4(Input, ctx->s_c4in.p, 1, IFM, Ih, Iw, 0, 0, ctx->elem, st);
        launch_wprep_tc32((const int16_t *)Weight, ctx->s_wprep.p, IFM, OFM, K, TM, TN, ctx->Tn, st);
        p.B = 1; p.H = Ih; p.W = Iw; p.G = ceil_div(IFM, 4); p.OFM = OFM;
        p.in = ctx->s_c4in.p; p.out = ctx->s_c4out.p; p.w = ctx->s_wprep.p; p.bias = Beta;
        p.in_frame_stride = 0; p.out_frame_stride = 0;
        p.so = so; p.sb = sb; p.leaky = IsNL;
        if (launch_conv_i16_tc32(p, K, IFM, ctx->Tn, st, &ctx->last_kernel) > 0) {
            launch_c4_to_planar(ctx->s_c4out.p, Output, 1, OFM, Oh, Ow, 0, 0, ctx->elem, st);
            ctx->launches += 4;
            CUDA_OK(ctx, cudaGetLastError());
            return YOLO2CUDA_SUCCESS;
        }
    }
    if (fast) {
        p.B = 1; p.H = Ih; p.W = Iw; p.G = ceil_div(IFM, 4); p.OFM = OFM;
        if (conv_fast_plan(p, K, ctx->elem) == 0) fast = false;
    }
    if (!fast) {
        ctx->last_kernel = ctx->elem == 2 ? "conv_i16_generic" : "conv_f32_generic";
        if (ctx->elem == 2)
            launch_conv_i16_generic((const int16_t *)Input, (int16_t *)Output, (const int16_t *)Weight,
                                    (const int16_t *)Beta, IFM, OFM, K, S, Iw, Ih, Ow, Oh, Pad, IsNL, TM, TN, so, sb, st);
        else
            launch_conv_f32_generic((const float *)Input, (float *)Output, (const float *)Weight, (const float *)Beta,
                                    IFM, OFM, K, S, Iw, Ih, Ow, Oh, Pad, IsNL, TM, TN, st);
        ctx->launches += 1;
        CUDA_OK(ctx, cudaGetLastError());
        return YOLO2CUDA_SUCCESS;
    }
    int rc;
    if ((rc = ensure(ctx, ctx->s_c4in, c4_elems(IFM, Ih, Iw) * ctx->elem))) return rc;
    if ((rc = ensure(ctx, ctx->s_c4out, c4_elems(OFM, Oh, Ow) * ctx->elem))) return rc;
    const bool tc = ctx->use_tc > 0 && ctx->elem == 2 && so >= 8 && so <= 22 && gwords == 1;   // (auto mode: single-frame calls stay on the CUDA cores)
    if ((rc = ensure(ctx, ctx->s_wprep, tc ? wprep_tc2_bytes(IFM, OFM, K) : wprep_bytes(IFM, OFM, K, ctx->elem)))) return rc;
    launch_planar_to_c4(Input, ctx->s_c4in.p, 1, IFM, Ih, Iw, 0, 0, ctx->elem, st);
    if (tc) {
        launch_wprep_tc2((const int16_t *)Weight, ctx->s_wprep.p, IFM, OFM, K, TM, TN, so, st);
        p.in = ctx->s_c4in.p; p.out = ctx->s_c4out.p; p.w = ctx->s_wprep.p; p.bias = Beta;
        p.in_frame_stride = 0; p.out_frame_stride = 0;
        p.so = so; p.sb = sb; p.leaky = IsNL;
        p.tc_stats = ctx->d_tc_stats; p.tc_force_exact = ctx->tc_force_exact;     // (xmax_in = NULL: the caller's tensor is unknown, 32768 is assumed)
        if (launch_conv_i16_tc2(p, K, st, &ctx->last_kernel) < 0) return fail(ctx, YOLO2CUDA_LAUNCH_ERROR, "tc conv not eligible");
        launch_c4_to_planar(ctx->s_c4out.p, Output, 1, OFM, Oh, Ow, 0, 0, ctx->elem, st);
        ctx->launches += 4;
        CUDA_OK(ctx, cudaGetLastError());
        return YOLO2CUDA_SUCCESS;
    }
    if (ctx->elem == 2) launch_wprep_i16((const int16_t *)Weight, ctx->s_wprep.p, IFM, OFM, K, TM, TN, st);
    else launch_wprep_f32((const float *)Weight, ctx->s_wprep.p, IFM, OFM, K, TM, TN, st);
    p.in = ctx->s_c4in.p; p.out = ctx->s_c4out.p; p.w = ctx->s_wprep.p; p.bias = Beta;
    p.in_frame_stride = 0; p.out_frame_stride = 0;
    p.so = so > 30 ? 30 : so; p.sb = sb; p.leaky = IsNL;
    int n = ctx->elem == 2 ? launch_conv_i16_fast(p, K, st, &ctx->last_kernel) : launch_conv_f32_fast(p, K, st, &ctx->last_kernel);
    if (n < 0) return fail(ctx, YOLO2CUDA_LAUNCH_ERROR, "no fast conv variant for TP=%d K=%d", p.TP, K);
    launch_c4_to_planar(ctx->s_c4out.p, Output, 1, OFM, Oh, Ow, 0, 0, ctx->elem, st);
    ctx->launches += 3 + n;
    CUDA_OK(ctx, cudaGetLastError());
    return YOLO2CUDA_SUCCESS;
}

}  // namespace

extern "C" {

int yolo2cuda_create(yolo2cuda_ctx **out, int device, int precision)
{
    if (!out) return YOLO2CUDA_ERROR;
    *out = nullptr;
    if (precision != YOLO2CUDA_PRECISION_INT16 && precision != YOLO2CUDA_PRECISION_FP32) return YOLO2CUDA_ERROR;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) {
        cudaGetLastError();
        return YOLO2CUDA_INIT_ERROR;
    }
    if (cudaSetDevice(device) != cudaSuccess) return YOLO2CUDA_INIT_ERROR;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return YOLO2CUDA_INIT_ERROR;
    if (prop.major != 10) {
        fprintf(stderr, "yolo2cuda: device %d is sm_%d%d; this library contains sm_100a code only\n", device, prop.major, prop.minor);
        return YOLO2CUDA_INIT_ERROR;
    }
    yolo2cuda_ctx *ctx = new yolo2cuda_ctx();
    ctx->device = device;
    ctx->precision = precision;
    ctx->elem = precision == YOLO2CUDA_PRECISION_INT16 ? 2 : 4;
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return YOLO2CUDA_INIT_ERROR;
    }
    ctx->stream = ctx->own_stream;
    const char *fg = getenv("YOLO2CUDA_FORCE_GENERIC");
    ctx->force_generic = (fg && fg[0] && fg[0] != '0') ? 1 : 0;
    const char *tc = getenv("YOLO2CUDA_TC");
    ctx->use_tc = (tc && tc[0]) ? (tc[0] == '0' ? 0 : 2) : -1;
    if (const char *mo = getenv("YOLO2CUDA_TC_MIN_OFM")) ctx->tc_min_ofm = atoi(mo);
    if (const char *g1 = getenv("YOLO2CUDA_G1")) ctx->use_g1 = g1[0] != '0';
    const char *ex = getenv("YOLO2CUDA_TC_EXACT");
    ctx->tc_force_exact = (ex && ex[0] && ex[0] != '0') ? 1 : 0;
    if (cudaMalloc(&ctx->d_tc_stats, 2 * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMemset(ctx->d_tc_stats, 0, 2 * sizeof(unsigned long long)) != cudaSuccess) {
        cudaStreamDestroy(ctx->own_stream);
        delete ctx;
        return YOLO2CUDA_MEMORY_ERROR;
    }
    *out = ctx;
    return YOLO2CUDA_SUCCESS;
}

int yolo2cuda_destroy(yolo2cuda_ctx *ctx)
{
    if (!ctx) return YOLO2CUDA_ERROR;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    yolo2cuda_ctx::Scratch *all[] = {&ctx->s_in, &ctx->s_out, &ctx->s_w, &ctx->s_b, &ctx->s_c4in, &ctx->s_c4out, &ctx->s_wprep};
    for (auto *s : all)
        if (s->p) cudaFree(s->p);
    if (ctx->d_tc_stats) cudaFree(ctx->d_tc_stats);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return YOLO2CUDA_SUCCESS;
}

int yolo2cuda_tc_path_counts(yolo2cuda_ctx *ctx, uint64_t *fast_tiles, uint64_t *exact_tiles, int reset)
{
    if (!ctx) return YOLO2CUDA_ERROR;
    CUDA_OK(ctx, cudaSetDevice(ctx->device));
    CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    unsigned long long h[2] = {0, 0};
    CUDA_OK(ctx, cudaMemcpy(h, ctx->d_tc_stats, sizeof(h), cudaMemcpyDeviceToHost));
    if (fast_tiles) *fast_tiles = h[0];
    if (exact_tiles) *exact_tiles = h[1];
    if (reset) CUDA_OK(ctx, cudaMemset(ctx->d_tc_stats, 0, sizeof(h)));
    return YOLO2CUDA_SUCCESS;
}

int yolo2cuda_set_tile_params(yolo2cuda_ctx *ctx, int Tn, int Tm)
{
    if (!ctx) return YOLO2CUDA_ERROR;
    if (Tn <= 0 || Tn > 64 || Tm <= 0 || Tm > 2048) return fail(ctx, YOLO2CUDA_ERROR, "tile parameters out of range (0 < Tn <= 64, 0 < Tm <= 2048)");
    ctx->Tn = Tn;
    ctx->Tm = Tm;
    return YOLO2CUDA_SUCCESS;
}

int yolo2cuda_set_stream(yolo2cuda_ctx *ctx, void *cuda_stream)
{
    if (!ctx) return YOLO2CUDA_ERROR;
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return YOLO2CUDA_SUCCESS;
}

int yolo2cuda_synchronize(yolo2cuda_ctx *ctx)
{
    if (!ctx) return YOLO2CUDA_ERROR;
    CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return YOLO2CUDA_SUCCESS;
}

const char *yolo2cuda_last_error(const yolo2cuda_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }
uint64_t yolo2cuda_launch_count(const yolo2cuda_ctx *ctx) { return ctx ? ctx->launches : 0; }
const char *yolo2cuda_last_kernel(const yolo2cuda_ctx *ctx) { return ctx ? ctx->last_kernel : ""; }

int yolo2cuda_layer_dev(yolo2cuda_ctx *ctx, const void *Input, void *Output, const void *Weight, const void *Beta,
                        int IFM_num, int OFM_num, int Ksize, int Kstride, int Input_w, int Input_h, int Output_w,
                        int Output_h, int Padding, int IsNL, int IsBN, int TM, int TN, int TR, int TC,
                        int OFM_num_bound, int mLoopsxTM, int mLoops_a1xTM, int LayerType, int Qw, int Qa_in,
                        int Qa_out, int Qb)
{
    (void)IsBN;  // unused by the reference too (BN is pre-folded, yolo2_accel.cpp:27)
    if (!ctx) return YOLO2CUDA_ERROR;
    if (!Input || !Output) return fail(ctx, YOLO2CUDA_ERROR, "Input/Output must not be NULL");
    const char *why = validate_layer_args(IFM_num, OFM_num, Ksize, Kstride, Input_w, Input_h, Output_w, Output_h, Padding,
                                          TM, TN, TR, TC, OFM_num_bound, mLoopsxTM, mLoops_a1xTM, LayerType, ctx->Tn, ctx->Tm);
    if (why) return fail(ctx, YOLO2CUDA_ERROR, "invalid layer arguments: %s", why);
    CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    if (LayerType == YOLO2CUDA_LAYER_CONV) {
        if (!Weight || !Beta) return fail(ctx, YOLO2CUDA_ERROR, "conv needs Weight and Beta");
        // every output pixel must read inside the zero-padded input (core_io.cpp:53-70 pads, never wraps)
        return run_conv_layer_dev(ctx, Input, Output, Weight, Beta, IFM_num, OFM_num, Ksize, Kstride, Input_w, Input_h,
                                  Output_w, Output_h, Padding, IsNL, TM, TN, Qw, Qa_in, Qa_out, Qb);
    }
    if (LayerType == YOLO2CUDA_LAYER_MAXPOOL) {
        // the reference store is hard-wired to window position (1,1): only 2x2 windows are defined (core_compute.cpp:299-300)
        if (Ksize != 2) return fail(ctx, YOLO2CUDA_ERROR, "maxpool supports Ksize==2 only (reference store is hard-wired)");
        launch_maxpool_planar(Input, Output, OFM_num, Ksize, Kstride, Input_w, Input_h, Output_w, Output_h, ctx->elem, st);
        ctx->last_kernel = "maxpool_planar";
    } else {
        launch_reorg_hls_planar(Input, Output, OFM_num, TM, Input_w, Input_h, Output_w, Output_h, ctx->elem, st);
        ctx->last_kernel = "reorg_hls_planar";
    }
    ctx->launches += 1;
    CUDA_OK(ctx, cudaGetLastError());
    return YOLO2CUDA_SUCCESS;
}

int yolo2cuda_layer_host(yolo2cuda_ctx *ctx, const void *Input, void *Output, const void *Weight, const void *Beta,
                         int IFM_num, int OFM_num, int Ksize, int Kstride, int Input_w, int Input_h, int Output_w,
                         int Output_h, int Padding, int IsNL, int IsBN, int TM, int TN, int TR, int TC,
                         int OFM_num_bound, int mLoopsxTM, int mLoops_a1xTM, int LayerType, int Qw, int Qa_in,
                         int Qa_out, int Qb)
{
    if (!ctx) return YOLO2CUDA_ERROR;
    if (!Input || !Output) return fail(ctx, YOLO2CUDA_ERROR, "Input/Output must not be NULL");
    const char *why = validate_layer_args(IFM_num, OFM_num, Ksize, Kstride, Input_w, Input_h, Output_w, Output_h, Padding,
                                          TM, TN, TR, TC, OFM_num_bound, mLoopsxTM, mLoops_a1xTM, LayerType, ctx->Tn, ctx->Tm);
    if (why) return fail(ctx, YOLO2CUDA_ERROR, "invalid layer arguments: %s", why);
    CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t e = ctx->elem;
    const size_t in_b = planar_elems(IFM_num, Input_h, Input_w) * e, out_b = planar_elems(OFM_num, Output_h, Output_w) * e;
    int rc;
    if ((rc = ensure(ctx, ctx->s_in, in_b))) return rc;
    if ((rc = ensure(ctx, ctx->s_out, out_b))) return rc;
    CUDA_OK(ctx, cudaMemcpyAsync(ctx->s_in.p, Input, in_b, cudaMemcpyHostToDevice, st));
    // pad columns W..ceil8(W)-1 of Output are never written by the reference (core_compute.cpp:218-219):
    // round-trip the caller's buffer so they keep their values
    CUDA_OK(ctx, cudaMemcpyAsync(ctx->s_out.p, Output, out_b, cudaMemcpyHostToDevice, st));
    const void *dW = nullptr, *dB = nullptr;
    if (LayerType == YOLO2CUDA_LAYER_CONV) {
        if (!Weight || !Beta) return fail(ctx, YOLO2CUDA_ERROR, "conv needs Weight and Beta");
        const size_t w_b = (size_t)IFM_num * OFM_num * Ksize * Ksize * e, b_b = (size_t)OFM_num * e;
        if ((rc = ensure(ctx, ctx->s_w, w_b))) return rc;
        if ((rc = ensure(ctx, ctx->s_b, b_b))) return rc;
        CUDA_OK(ctx, cudaMemcpyAsync(ctx->s_w.p, Weight, w_b, cudaMemcpyHostToDevice, st));
        CUDA_OK(ctx, cudaMemcpyAsync(ctx->s_b.p, Beta, b_b, cudaMemcpyHostToDevice, st));
        dW = ctx->s_w.p;
        dB = ctx->s_b.p;
    }
    rc = yolo2cuda_layer_dev(ctx, ctx->s_in.p, ctx->s_out.p, dW, dB, IFM_num, OFM_num, Ksize, Kstride, Input_w, Input_h,
                             Output_w, Output_h, Padding, IsNL, IsBN, TM, TN, TR, TC, OFM_num_bound, mLoopsxTM,
                             mLoops_a1xTM, LayerType, Qw, Qa_in, Qa_out, Qb);
    if (rc) return rc;
    CUDA_OK(ctx, cudaMemcpyAsync(Output, ctx->s_out.p, out_b, cudaMemcpyDeviceToHost, st));
    CUDA_OK(ctx, cudaStreamSynchronize(st));
    return YOLO2CUDA_SUCCESS;
}

int yolo2cuda_quantize_input_dev(yolo2cuda_ctx *ctx, const float *in, int16_t *out, size_t count, int q_in)
{
    if (!ctx || !in || !out) return YOLO2CUDA_ERROR;
    CUDA_OK(ctx, cudaSetDevice(ctx->device));
    launch_quantize(in, out, count, q_in, ctx->stream);
    ctx->launches += 1;
    CUDA_OK(ctx, cudaGetLastError());
    return YOLO2CUDA_SUCCESS;
}

int yolo2cuda_reorg_dev(yolo2cuda_ctx *ctx, const void *in, void *out, int c, int h, int w, int shift)
{
    if (!ctx || !in || !out) return YOLO2CUDA_ERROR;
    if (c <= 0 || h <= 0 || w <= 0 || (h & 1) || (w & 1) || (((long long)h * c) & 3) || shift < 0 || shift > 31)
        return fail(ctx, YOLO2CUDA_ERROR, "reorg needs even h,w, h*c divisible by 4 and 0<=shift<=31");
    CUDA_OK(ctx, cudaSetDevice(ctx->device));
    launch_reorg_driver_planar(in, out, c, h, w, shift, ctx->elem, ctx->stream);
    ctx->launches += 1;
    CUDA_OK(ctx, cudaGetLastError());
    return YOLO2CUDA_SUCCESS;
}

int yolo2cuda_region_dev(yolo2cuda_ctx *ctx, const void *in, float *out, int w, int h, int n, int classes, int coords,
                         int softmax, int background, int q)
{
    if (!ctx || !in || !out) return YOLO2CUDA_ERROR;
    if (w <= 0 || h <= 0 || n <= 0 || classes <= 0 || coords < 4) return fail(ctx, YOLO2CUDA_ERROR, "bad region dims");
    CUDA_OK(ctx, cudaSetDevice(ctx->device));
    launch_region(in, out, 1, w, h, n, classes, coords, softmax, background, q, 0, 0, ctx->elem, ctx->stream);
    ctx->launches += 1;
    CUDA_OK(ctx, cudaGetLastError());
    return YOLO2CUDA_SUCCESS;
}

int yolo2cuda_region_detections_dev(yolo2cuda_ctx *ctx, const float *region, int batch, int lw, int lh, int n, int classes,
                                    const float *anchors_host, int im_w, int im_h, int net_w, int net_h, float thresh, float nms,
                                    float *boxes, float *probs, float *objectness)
{
    if (!ctx || !region || !anchors_host || !boxes || !probs || !objectness) return YOLO2CUDA_ERROR;
    if (batch <= 0 || lw <= 0 || lh <= 0 || n <= 0 || classes <= 0 || im_w <= 0 || im_h <= 0 || net_w <= 0 || net_h <= 0)
        return fail(ctx, YOLO2CUDA_ERROR, "region_detections_dev: bad dimensions");
    CUDA_OK(ctx, cudaSetDevice(ctx->device));
    // 2^(i/32) correctly rounded: the table of glibc's expf (long double exp2, then one rounding to double).  A function-local
    // static initialised by a lambda: C++11 guarantees one thread builds it and every other caller sees it complete.
    struct ExpfTable { double v[32]; };
    static const ExpfTable tab = [] {
        ExpfTable t;
        for (int i = 0; i < 32; ++i) t.v[i] = (double)exp2l((long double)i / 32);
        return t;
    }();
    if (launch_detect(region, boxes, probs, objectness, batch, lw, lh, n, classes, anchors_host, im_w, im_h, net_w, net_h, thresh, nms,
                      tab.v, ctx->stream) < 0)
        return fail(ctx, YOLO2CUDA_ERROR, "region_detections_dev: more than 1024 candidates per frame (use yolo2cuda_region_detections)");
    ctx->launches += 1;
    CUDA_OK(ctx, cudaGetLastError());
    return YOLO2CUDA_SUCCESS;
}

int yolo2cuda_compact_detections_dev(yolo2cuda_ctx *ctx, const float *boxes, const float *probs, const float *objectness, int batch,
                                     int total, int classes, int cap, uint32_t *records, int32_t *counts)
{
    if (!ctx || !boxes || !probs || !objectness || !records || !counts) return YOLO2CUDA_ERROR;
    if (batch <= 0 || total <= 0 || classes <= 0 || cap <= 0) return fail(ctx, YOLO2CUDA_ERROR, "compact_detections_dev: bad dimensions");
    CUDA_OK(ctx, cudaSetDevice(ctx->device));
    launch_compact_detections(boxes, probs, objectness, batch, total, classes, cap, records, counts, ctx->stream);
    ctx->launches += 1;
    CUDA_OK(ctx, cudaGetLastError());
    return YOLO2CUDA_SUCCESS;
}

}  // extern "C"

// ================================ whole-network executor ========================================

struct TensorView {
    void *base = nullptr;          // frame 0, first group of this tensor
    long long frame_stride = 0;    // elements between frames
    int C = 0, H = 0, W = 0;
};

struct LayerPlan {
    yolo2cuda_layer_desc d;
    TensorView in, out;
    bool fast = false;
    ConvFastParams cp{};
    int conv_index = -1;
    size_t w_off = 0, b_off = 0;   // element offsets into the reference blobs
    void *w_dev = nullptr;         // device weight layout (fast path)
    void *w_tc = nullptr;          // tcgen05 operand tiles (tensor-core path)
    bool tc = false;
    bool tc32 = false;             // Tn = 32 / 16 / 8 build on the tensor cores: csrc/conv_i16_tc32.cu
    bool g1 = false;               // one input channel group (IFM <= 4), 3x3: csrc/conv_i16_g1.cu
    bool pool_fusable = false;     //   ... and the next layer is a 2x2 / stride-2 max-pool of even dims that only reads this layer
    bool fused_away = false;       // max-pool layer whose work the previous conv launch did (set per forward)
    int Qw = 0, Qa_in = 0, Qa_out = 0, Qb = 0;
    int reorg_shift = 0;
    int region_q = 0;
    const char *variant = "";
};

struct yolo2cuda_net {
    yolo2cuda_ctx *ctx = nullptr;
    std::vector<LayerPlan> L;
    int max_batch = 0;
    int ramp_frames = -1;          // first pass of a multi-pass host forward; -1 = max(32, max_batch / 6), 0 = no ramp pass
    int in_c = 0, in_h = 0, in_w = 0;
    size_t region_outputs = 0;
    std::vector<void *> owned;     // device allocations that live as long as the net (weights, staging, scratch)
    // activation tensors: [0] = the quantised C4 input, [1 + i] = output of layer i (conv / pool / reorg outside a concat) or the
    // concat buffer of a multi-input route i.  place_tensors() maps them to device memory.
    struct Tensor { size_t bytes = 0; int first_def = 1 << 30, last_use = -1; size_t offset = 0; void *own = nullptr; };
    std::vector<Tensor> T;
    std::vector<int> concat_of, concat_goff;   // per layer: the route whose concat buffer this layer writes into (-1), and its group offset
    std::vector<int> tensor_in, tensor_out;    // per layer: tensor read / written (-1: none)
    // largest-|activation| bookkeeping for the tcgen05 kernel's fast path: one device int per tensor CLASS (tensors joined by
    // max-pool / reorg / concat share a slot: those operators never increase the largest magnitude)
    int *d_xmax = nullptr;
    std::vector<int> xslot;        // per tensor: slot index
    std::vector<char> xknown;      // per slot: every writer of the class leaves its maximum behind
    void *arena = nullptr;         // compact mode: one allocation, tensors at liveness-packed offsets
    size_t arena_bytes = 0;
    bool keep_all = false;         // debug mode: every tensor owns its memory, so every layer's ofm survives the forward
    void *d_input_c4 = nullptr;
    void *d_frames2[2] = {nullptr, nullptr}, *d_region2[2] = {nullptr, nullptr};   // double-buffered staging
    void *d_lb_frames = nullptr, *d_lb_region = nullptr, *d_lb_img = nullptr;       // forward_images_host: letterboxed frames, region, raw u8 images
    size_t lb_img_bytes = 0;
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_d2h[2] = {nullptr, nullptr};
    void *d_wblob = nullptr, *d_bblob = nullptr;
    void *d_tmp_planar_in = nullptr, *d_tmp_planar_out = nullptr;  // generic-fallback / dump scratch
    size_t tmp_planar_elems = 0;
    bool weights_loaded = false;
    int region_q = 0;
    int input_q = 0;
    uint64_t launches_per_forward = 0;
    int last_batch = 0;
    bool timing = false;
    std::vector<cudaEvent_t> ev;
    std::vector<float> layer_ms;
};

namespace {

int net_alloc(yolo2cuda_net *net, void **p, size_t bytes)
{
    yolo2cuda_ctx *ctx = net->ctx;
    CUDA_OK(ctx, cudaMalloc(p, bytes ? bytes : 16));
    CUDA_OK(ctx, cudaMemsetAsync(*p, 0, bytes ? bytes : 16, ctx->stream));
    net->owned.push_back(*p);
    return YOLO2CUDA_SUCCESS;
}

// Index of the conv whose output Q is remembered for the concat (the reference hard-codes
// `i == 24`, yolo2_model.cpp:332-334): the non-reorg input of the first multi-input route that
// has a reorg input.
int find_skip_layer(const std::vector<LayerPlan> &L)
{
    for (size_t i = 0; i < L.size(); ++i)
        if (L[i].d.type == YOLO2CUDA_ROUTE && L[i].d.n_inputs >= 2)
            for (int a = 0; a < L[i].d.n_inputs; ++a)
                if (L[L[i].d.inputs[a]].d.type == YOLO2CUDA_REORG)
                    for (int b = 0; b < L[i].d.n_inputs; ++b)
                        if (b != a) return L[i].d.inputs[b];
    return -1;
}

// Maps the activation tensors to device memory and (re)builds every layer's in/out view.
//   compact (default): ONE arena; a tensor occupies [offset, offset + bytes) only between its first writer and its last reader,
//     so buffers are recycled down the network like the reference's ping-pong scratch arena (yolo2_model.cpp:56-110) - about a
//     third of the keep-all footprint for YOLOv2 (the two live tensors of the widest layers instead of all 31).
//   keep_all: every tensor owns its memory, so yolo2cuda_net_get_layer_output works for every layer after a forward.
int place_tensors(yolo2cuda_net *net, bool keep_all)
{
    yolo2cuda_ctx *ctx = net->ctx;
    const int e = ctx->elem;
    cudaStreamSynchronize(ctx->stream);
    if (net->arena) { cudaFree(net->arena); net->arena = nullptr; net->arena_bytes = 0; }
    for (auto &t : net->T)
        if (t.own) { cudaFree(t.own); t.own = nullptr; }
    net->keep_all = keep_all;
    auto base_of = [&](int t) -> char * { return keep_all ? (char *)net->T[t].own : (char *)net->arena + net->T[t].offset; };
    if (keep_all) {
        for (auto &t : net->T)
            if (t.bytes) {
                CUDA_OK(ctx, cudaMalloc(&t.own, t.bytes));
                CUDA_OK(ctx, cudaMemsetAsync(t.own, 0, t.bytes, ctx->stream));
            }
    } else {
        // first-fit by address over the tensors in order of their first writer; two tensors may share bytes only when their
        // [first_def, last_use] layer intervals are disjoint (closed intervals: a layer's input and output never overlap)
        std::vector<int> order;
        for (int t = 0; t < (int)net->T.size(); ++t)
            if (net->T[t].bytes) order.push_back(t);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return net->T[a].first_def < net->T[b].first_def; });
        std::vector<int> placed;
        size_t top = 0;
        for (int t : order) {
            auto &x = net->T[t];
            const size_t need = (x.bytes + 1023) & ~(size_t)1023;
            std::vector<std::pair<size_t, size_t>> busy;       // [begin, end) of the placed tensors alive at the same time
            for (int u : placed) {
                const auto &y = net->T[u];
                if (y.first_def <= x.last_use && x.first_def <= y.last_use) busy.emplace_back(y.offset, y.offset + ((y.bytes + 1023) & ~(size_t)1023));
            }
            std::sort(busy.begin(), busy.end());
            size_t at = 0;
            for (auto &b : busy) {
                if (at + need <= b.first) break;
                at = std::max(at, b.second);
            }
            x.offset = at;
            top = std::max(top, at + need);
            placed.push_back(t);
        }
        net->arena_bytes = top;
        CUDA_OK(ctx, cudaMalloc(&net->arena, top ? top : 16));
        CUDA_OK(ctx, cudaMemsetAsync(net->arena, 0, top ? top : 16, ctx->stream));
    }
    // ---- views ------------------------------------------------------------------------------------
    const int n_layers = (int)net->L.size();
    net->d_input_c4 = base_of(0);
    for (int i = 0; i < n_layers; ++i) {
        LayerPlan &l = net->L[i];
        const yolo2cuda_layer_desc &d = l.d;
        if (i == 0) {
            l.in.base = net->d_input_c4;
            l.in.frame_stride = (long long)c4_elems(net->in_c, net->in_h, net->in_w);
            l.in.C = net->in_c; l.in.H = net->in_h; l.in.W = net->in_w;
        } else if (d.type != YOLO2CUDA_ROUTE) {
            l.in = net->L[i - 1].out;
        }
        if (d.type == YOLO2CUDA_ROUTE) {
            if (d.n_inputs == 1) {
                l.out = net->L[d.inputs[0]].out;
            } else {
                l.out.base = base_of(1 + i);
                l.out.frame_stride = (long long)c4_elems(d.out_c, d.out_h, d.out_w);
                l.out.C = d.out_c; l.out.H = d.out_h; l.out.W = d.out_w;
            }
        } else if (d.type == YOLO2CUDA_REGION) {
            l.out = TensorView{};
        } else {
            l.out.C = d.out_c; l.out.H = d.out_h; l.out.W = d.out_w;
            const int co = net->concat_of[i];
            if (co >= 0) {
                const yolo2cuda_layer_desc &cd = net->L[co].d;
                l.out.frame_stride = (long long)c4_elems(cd.out_c, cd.out_h, cd.out_w);
                l.out.base = base_of(1 + co) + (size_t)net->concat_goff[i] * d.out_h * d.out_w * 4 * e;
            } else {
                l.out.base = base_of(1 + i);
                l.out.frame_stride = (long long)c4_elems(d.out_c, d.out_h, d.out_w);
            }
        }
        // launch parameters prepared by net_load_weights point into the tensors
        l.cp.in = l.in.base; l.cp.out = l.out.base;
        l.cp.in_frame_stride = l.in.frame_stride; l.cp.out_frame_stride = l.out.frame_stride;
    }
    CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return YOLO2CUDA_SUCCESS;
}

int forward_chunk(yolo2cuda_net *net, const float *frames_dev, int B, float *region_dev)
{
    yolo2cuda_ctx *ctx = net->ctx;
    cudaStream_t st = ctx->stream;
    const int e = ctx->elem;
    uint64_t launches = 0;
    const long long in_stride = (long long)c4_elems(net->in_c, net->in_h, net->in_w);
    if (net->d_xmax) cudaMemsetAsync(net->d_xmax, 0, net->T.size() * sizeof(int), st);   // this pass's largest |activation| per tensor class
    launch_frames_to_c4(frames_dev, net->d_input_c4, B, net->in_c, net->in_h, net->in_w, in_stride,
                        net->input_q, e, st);
    ++launches;
    for (size_t i = 0; i < net->L.size(); ++i) {
        LayerPlan &l = net->L[i];
        if (net->timing) cudaEventRecord(net->ev[i], st);
        switch (l.d.type) {
        case YOLO2CUDA_CONV: {
            if (l.tc32) {
                ConvFastParams p = l.cp;
                p.B = B;
                int n = launch_conv_i16_tc32(p, l.d.size, l.d.c, ctx->Tn, st, &l.variant);
                if (n < 0) return fail(ctx, YOLO2CUDA_LAUNCH_ERROR, "layer %zu: Tn=%d tensor-core conv not eligible", i, ctx->Tn);
                launches += n;
                ctx->last_kernel = l.variant;
            } else if (l.fast) {
                ConvFastParams p = l.cp;
                p.B = B;
                int n;
                if (l.tc) {
                    p.w = l.w_tc;
                    n = launch_conv_i16_tc2(p, l.d.size, st, &l.variant);
                } else if (l.g1) {
                    // fused with the max-pool that follows unless every layer's output has to survive (debug keep-all mode)
                    const bool pool = l.pool_fusable && !net->keep_all;
                    if (pool) {
                        LayerPlan &nx = net->L[i + 1];
                        p.out = nx.out.base;
                        p.out_frame_stride = nx.out.frame_stride;
                        nx.fused_away = true;
                    }
                    n = launch_conv_i16_g1(p, l.d.size, pool, st, &l.variant);
                } else {
                    n = e == 2 ? launch_conv_i16_fast(p, l.d.size, st, &l.variant) : launch_conv_f32_fast(p, l.d.size, st, &l.variant);
                }
                if (n < 0) return fail(ctx, YOLO2CUDA_LAUNCH_ERROR, "layer %zu: no fast conv variant", i);
                launches += n;
                ctx->last_kernel = l.variant;
            } else {
                // contract-complete fallback, one frame at a time through the planar kernels
                const int TM = l.d.n < ctx->Tm ? l.d.n : ctx->Tm, TN = l.d.c < ctx->Tn ? l.d.c : ctx->Tn;
                for (int f = 0; f < B; ++f) {
                    launch_c4_to_planar((char *)l.in.base + (size_t)f * l.in.frame_stride * e, net->d_tmp_planar_in, 1, l.d.c,
                                        l.d.h, l.d.w, 0, 0, e, st);
                    if (e == 2)
                        launch_conv_i16_generic((const int16_t *)net->d_tmp_planar_in, (int16_t *)net->d_tmp_planar_out,
                                                (const int16_t *)net->d_wblob + l.w_off, (const int16_t *)net->d_bblob + l.b_off,
                                                l.d.c, l.d.n, l.d.size, l.d.stride, l.d.w, l.d.h, l.d.out_w, l.d.out_h, l.d.pad,
                                                l.d.leaky, TM, TN, l.Qa_in + l.Qw - l.Qa_out, l.Qb - l.Qa_out, st);
                    else
                        launch_conv_f32_generic((const float *)net->d_tmp_planar_in, (float *)net->d_tmp_planar_out,
                                                (const float *)net->d_wblob + l.w_off, (const float *)net->d_bblob + l.b_off,
                                                l.d.c, l.d.n, l.d.size, l.d.stride, l.d.w, l.d.h, l.d.out_w, l.d.out_h, l.d.pad,
                                                l.d.leaky, TM, TN, st);
                    launch_planar_to_c4(net->d_tmp_planar_out, (char *)l.out.base + (size_t)f * l.out.frame_stride * e, 1,
                                        l.d.out_c, l.d.out_h, l.d.out_w, 0, 0, e, st);
                    launches += 3;
                }
                ctx->last_kernel = l.variant = e == 2 ? "conv_i16_generic" : "conv_f32_generic";
            }
            break;
        }
        case YOLO2CUDA_MAXPOOL:
            if (l.fused_away) {       // the previous conv launch pooled in its store (csrc/conv_i16_g1.cu)
                l.fused_away = false;
                l.variant = "(fused into the conv store)";
                break;
            }
            launch_maxpool_c4(l.in.base, l.out.base, B, ceil_div(l.d.c, 4), l.d.stride, l.d.w, l.d.h, l.d.out_w, l.d.out_h,
                              l.in.frame_stride, l.out.frame_stride, e, st);
            l.variant = "maxpool_c4";
            ++launches;
            break;
        case YOLO2CUDA_REORG:
            launch_reorg_driver_c4(l.in.base, l.out.base, B, l.d.c, l.d.h, l.d.w, l.reorg_shift, l.in.frame_stride,
                                   l.out.frame_stride, e, st);
            l.variant = "reorg_driver_c4";
            ++launches;
            break;
        case YOLO2CUDA_ROUTE:
            break;  // no-op by placement, like the reference arena (yolo2_model.cpp:97-104,404-405)
        case YOLO2CUDA_REGION:
            launch_region(l.in.base, region_dev, B, l.d.w, l.d.h, l.d.n, l.d.classes, l.d.coords, l.d.softmax,
                          l.d.background, l.region_q, 1, l.in.frame_stride, e, st);
            l.variant = "region";
            ++launches;
            break;
        default:
            return fail(ctx, YOLO2CUDA_ERROR, "layer %zu: unsupported type %d", i, l.d.type);
        }
    }
    if (net->timing) cudaEventRecord(net->ev[net->L.size()], st);
    CUDA_OK(ctx, cudaGetLastError());
    ctx->launches += launches;
    net->launches_per_forward = launches;
    net->last_batch = B;
    return YOLO2CUDA_SUCCESS;
}

}  // namespace

extern "C" {

int yolo2cuda_net_create(yolo2cuda_ctx *ctx, const yolo2cuda_layer_desc *layers, int n_layers, int max_batch,
                         yolo2cuda_net **out)
{
    if (!ctx || !layers || n_layers <= 0 || max_batch <= 0 || !out) return YOLO2CUDA_ERROR;
    *out = nullptr;
    CUDA_OK(ctx, cudaSetDevice(ctx->device));
    const int e = ctx->elem;
    yolo2cuda_net *net = new yolo2cuda_net();
    net->ctx = ctx;
    net->max_batch = max_batch;
    net->L.resize(n_layers);
    for (int i = 0; i < n_layers; ++i) net->L[i].d = layers[i];
    net->in_c = layers[0].c; net->in_h = layers[0].h; net->in_w = layers[0].w;

#define NET_FAIL(...)                                    \
    do {                                                 \
        int rc_ = fail(ctx, YOLO2CUDA_ERROR, __VA_ARGS__); \
        yolo2cuda_net_destroy(net);                      \
        return rc_;                                      \
    } while (0)

    // ---- validate the table -------------------------------------------------------------------
    for (int i = 0; i < n_layers; ++i) {
        const yolo2cuda_layer_desc &d = layers[i];
        if (d.type == YOLO2CUDA_ROUTE) {
            if (d.n_inputs < 1 || d.n_inputs > 4) NET_FAIL("layer %d: route needs 1..4 inputs", i);
            for (int a = 0; a < d.n_inputs; ++a)
                if (d.inputs[a] < 0 || d.inputs[a] >= i) NET_FAIL("layer %d: route input %d out of range", i, d.inputs[a]);
        } else if (d.type == YOLO2CUDA_CONV) {
            if (d.c <= 0 || d.n <= 0 || d.size <= 0 || d.size > 3 || d.stride <= 0 || d.stride > 2 || d.pad < 0 || d.pad > 4)
                NET_FAIL("layer %d: conv parameters outside the accelerator contract", i);
            if (d.out_c != d.n) NET_FAIL("layer %d: out_c != filters", i);
        } else if (d.type == YOLO2CUDA_MAXPOOL) {
            if (d.size != 2 || d.stride <= 0 || d.stride > 2) NET_FAIL("layer %d: maxpool must be 2x2, stride 1 or 2", i);
        } else if (d.type == YOLO2CUDA_REORG) {
            if (d.stride != 2 || (d.h & 1) || (d.w & 1) || (d.c & 3)) NET_FAIL("layer %d: reorg needs stride 2, even h,w and c%%4==0", i);
        } else if (d.type == YOLO2CUDA_REGION) {
            if (i != n_layers - 1) NET_FAIL("layer %d: region must be the last layer", i);
            if (d.n * (d.coords + 1 + d.classes) != d.c) NET_FAIL("layer %d: region channel count mismatch", i);
        } else {
            NET_FAIL("layer %d: unknown type %d", i, d.type);
        }
    }

    // ---- place tensors: concat inputs write straight into the concat buffer ---------------------
    std::vector<int> concat_of(n_layers, -1), concat_goff(n_layers, 0);
    for (int i = 0; i < n_layers; ++i) {
        const yolo2cuda_layer_desc &d = layers[i];
        if (d.type != YOLO2CUDA_ROUTE || d.n_inputs < 2) continue;
        int goff = 0;
        for (int a = 0; a < d.n_inputs; ++a) {
            int s = d.inputs[a];
            // resolve aliases (single-input routes) to the producing layer
            while (layers[s].type == YOLO2CUDA_ROUTE && layers[s].n_inputs == 1) s = layers[s].inputs[0];
            if (layers[s].type == YOLO2CUDA_ROUTE) NET_FAIL("layer %d: nested concat is not supported", i);
            if (concat_of[s] >= 0) NET_FAIL("layer %d: tensor %d feeds two concats", i, s);
            if (layers[s].out_c % 4) NET_FAIL("layer %d: concat input %d has out_c %% 4 != 0", i, s);
            if (layers[s].out_h != d.out_h || layers[s].out_w != d.out_w) NET_FAIL("layer %d: concat input dims differ", i);
            concat_of[s] = i;
            concat_goff[s] = goff;
            goff += layers[s].out_c / 4;
        }
        if (goff * 4 != d.out_c) NET_FAIL("layer %d: concat out_c mismatch", i);
    }

    int rc;
    // ---- tensor table + lifetimes (first writer .. last reader, in layer order) ------------------
    net->concat_of = concat_of;
    net->concat_goff = concat_goff;
    net->T.assign(1 + n_layers, yolo2cuda_net::Tensor{});
    net->T[0].bytes = c4_elems(net->in_c, net->in_h, net->in_w) * (size_t)max_batch * e;
    net->T[0].first_def = -1;
    size_t max_planar = planar_elems(net->in_c, net->in_h, net->in_w);
    std::vector<int> tensor_of(n_layers, -1);      // tensor a layer's OUTPUT view lives in
    net->tensor_in.assign(n_layers, -1);
    int prev_tensor = 0;                           // tensor the next non-route layer reads
    for (int i = 0; i < n_layers; ++i) {
        const yolo2cuda_layer_desc &d = layers[i];
        auto touch_read = [&](int t) { if (t >= 0) net->T[t].last_use = std::max(net->T[t].last_use, i); };
        if (d.type == YOLO2CUDA_ROUTE) {
            if (d.n_inputs == 1) {
                tensor_of[i] = tensor_of[d.inputs[0]];
            } else {
                tensor_of[i] = 1 + i;
                net->T[1 + i].bytes = c4_elems(d.out_c, d.out_h, d.out_w) * (size_t)max_batch * e;
            }
            max_planar = std::max(max_planar, planar_elems(d.out_c, d.out_h, d.out_w));
        } else {
            if (i > 0 && (layers[i - 1].out_c != d.c || layers[i - 1].out_h != d.h || layers[i - 1].out_w != d.w))
                NET_FAIL("layer %d: input dims %dx%dx%d do not match the previous output %dx%dx%d", i, d.c, d.h, d.w,
                         layers[i - 1].out_c, layers[i - 1].out_h, layers[i - 1].out_w);
            touch_read(prev_tensor);
            net->tensor_in[i] = prev_tensor;
            if (d.type == YOLO2CUDA_REGION) {
                net->region_outputs = (size_t)d.c * d.h * d.w;
            } else {
                const int t = concat_of[i] >= 0 ? 1 + concat_of[i] : 1 + i;
                tensor_of[i] = t;
                if (concat_of[i] < 0) net->T[t].bytes = c4_elems(d.out_c, d.out_h, d.out_w) * (size_t)max_batch * e;
                net->T[t].first_def = std::min(net->T[t].first_def, i);
                max_planar = std::max(max_planar, planar_elems(d.out_c, d.out_h, d.out_w));
            }
        }
        prev_tensor = tensor_of[i];
    }
    // a concat buffer is also "read" by the route layer itself (keeps it alive between its writers and its consumer)
    for (int i = 0; i < n_layers; ++i)
        if (layers[i].type == YOLO2CUDA_ROUTE && tensor_of[i] >= 0)
            net->T[tensor_of[i]].last_use = std::max(net->T[tensor_of[i]].last_use, i);
    // A max-pool that may be fused into the store of the one-group 3x3 conv before it (csrc/conv_i16_g1.cu) is WRITTEN while that
    // conv runs: its output becomes live one layer early, so that the arena never packs it over the conv's own input.
    for (int i = 0; i + 1 < n_layers; ++i)
        if (layers[i].type == YOLO2CUDA_CONV && layers[i].size == 3 && layers[i].c <= 4 && layers[i + 1].type == YOLO2CUDA_MAXPOOL &&
            tensor_of[i + 1] >= 0)
            net->T[tensor_of[i + 1]].first_def = std::min(net->T[tensor_of[i + 1]].first_def, i);
    for (auto &t : net->T)
        if (t.bytes && t.last_use < t.first_def) t.last_use = t.first_def;       // written, never read: still needs a home
    net->tensor_out = tensor_of;
    if (e == 2 && (rc = net_alloc(net, (void **)&net->d_xmax, (1 + n_layers) * sizeof(int)))) {
        yolo2cuda_net_destroy(net);
        return rc;
    }
    if ((rc = place_tensors(net, false))) {
        yolo2cuda_net_destroy(net);
        return rc;
    }
    net->tmp_planar_elems = max_planar;
    if ((rc = net_alloc(net, &net->d_tmp_planar_in, max_planar * e)) || (rc = net_alloc(net, &net->d_tmp_planar_out, max_planar * e))) {
        yolo2cuda_net_destroy(net);
        return rc;
    }
    // weight blob offsets
    size_t woff = 0, boff = 0;
    int ci = 0;
    for (int i = 0; i < n_layers; ++i) {
        LayerPlan &l = net->L[i];
        if (l.d.type != YOLO2CUDA_CONV) continue;
        l.conv_index = ci++;
        l.w_off = woff; l.b_off = boff;
        woff += (size_t)l.d.c * l.d.n * l.d.size * l.d.size;
        boff += (size_t)l.d.n;
    }
#undef NET_FAIL
    CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    *out = net;
    return YOLO2CUDA_SUCCESS;
}

int yolo2cuda_net_destroy(yolo2cuda_net *net)
{
    if (!net) return YOLO2CUDA_ERROR;
    cudaSetDevice(net->ctx->device);
    cudaStreamSynchronize(net->ctx->stream);
    for (void *p : net->owned) cudaFree(p);
    if (net->arena) cudaFree(net->arena);
    for (auto &t : net->T)
        if (t.own) cudaFree(t.own);
    if (net->d_lb_img) cudaFree(net->d_lb_img);
    for (auto ev : net->ev) cudaEventDestroy(ev);
    for (int i = 0; i < 2; ++i) {
        if (net->ev_h2d[i]) cudaEventDestroy(net->ev_h2d[i]);
        if (net->ev_comp[i]) cudaEventDestroy(net->ev_comp[i]);
        if (net->ev_d2h[i]) cudaEventDestroy(net->ev_d2h[i]);
    }
    if (net->s_h2d) cudaStreamDestroy(net->s_h2d);
    if (net->s_d2h) cudaStreamDestroy(net->s_d2h);
    delete net;
    return YOLO2CUDA_SUCCESS;
}

int yolo2cuda_net_load_weights(yolo2cuda_net *net, const void *weights, size_t n_weights, const void *bias, size_t n_bias,
                               const int32_t *weight_q, const int32_t *bias_q, int n_q, const int32_t *act_q, int n_act_q)
{
    if (!net || !weights || !bias) return YOLO2CUDA_ERROR;
    yolo2cuda_ctx *ctx = net->ctx;
    net->weights_loaded = false;   // a failed (re)load leaves the layer plans half-updated: nothing may run until one succeeds
    CUDA_OK(ctx, cudaSetDevice(ctx->device));
    const int e = ctx->elem;
    cudaStream_t st = ctx->stream;
    size_t need_w = 0, need_b = 0;
    int n_conv = 0;
    for (auto &l : net->L)
        if (l.d.type == YOLO2CUDA_CONV) {
            need_w += (size_t)l.d.c * l.d.n * l.d.size * l.d.size;
            need_b += (size_t)l.d.n;
            ++n_conv;
        }
    if (n_weights < need_w) return fail(ctx, YOLO2CUDA_ERROR, "weights file too small");  // yolo2_model.cpp:173,181
    if (n_bias < need_b) return fail(ctx, YOLO2CUDA_ERROR, "bias file too small");
    if (e == 2) {
        if (!weight_q || !bias_q || n_q < n_conv) return fail(ctx, YOLO2CUDA_ERROR, "Q tables too small for conv layers");  // :186-188
        if (!act_q || n_act_q <= 0) return fail(ctx, YOLO2CUDA_ERROR, "Activation Q table (iofm_Q.bin) is required for int16 inference.");  // :258-260
    }
    int rc;
    if (!net->d_wblob) {
        if ((rc = net_alloc(net, &net->d_wblob, need_w * e))) return rc;
        if ((rc = net_alloc(net, &net->d_bblob, need_b * e))) return rc;
    }
    CUDA_OK(ctx, cudaMemcpyAsync(net->d_wblob, weights, need_w * e, cudaMemcpyHostToDevice, st));
    CUDA_OK(ctx, cudaMemcpyAsync(net->d_bblob, bias, need_b * e, cudaMemcpyHostToDevice, st));

    // ---- Q bookkeeping of the driver loop (yolo2_model.cpp:290-292, 311-336, 379-399, 416) ------
    int current_qa = (e == 2) ? act_q[0] : 0, route_q = 0, pending_route_q = -1;
    net->input_q = current_qa;  // yolo2_model.cpp:261
    const int skip_layer = find_skip_layer(net->L);
    for (size_t i = 0; i < net->L.size(); ++i) {
        LayerPlan &l = net->L[i];
        if (l.d.type == YOLO2CUDA_CONV) {
            const int ci = l.conv_index;
            if (e == 2) {
                l.Qa_in = (ci < n_act_q) ? act_q[ci] : current_qa;
                l.Qa_out = (ci + 1 < n_act_q) ? act_q[ci + 1] : l.Qa_in;
                l.Qw = weight_q[ci];
                l.Qb = bias_q[ci];
                if (pending_route_q >= 0) l.Qa_in = pending_route_q;
                current_qa = l.Qa_out;
                if ((int)i == skip_layer) route_q = current_qa;
                pending_route_q = -1;
            }
            const int so = l.Qa_in + l.Qw - l.Qa_out;
            const int TM = l.d.n < ctx->Tm ? l.d.n : ctx->Tm, TN = l.d.c < ctx->Tn ? l.d.c : ctx->Tn;  // yolo2_model.cpp:307-308
            l.fast = !ctx->force_generic && (l.d.size == 1 || l.d.size == 3) && l.d.stride == 1 && l.d.pad == l.d.size / 2 &&
                     l.d.out_w == l.d.w && l.d.out_h == l.d.h && (e == 4 || fast_shift_ok(so)) &&
                     (e == 4 || TN == 4 || l.d.c <= 4 ||   // int16: the C4 kernels' rounding group is 4 channels, or 2 / 4 C4 words (Tn 8 / 16)
                      ((ctx->Tn == 8 || ctx->Tn == 16) && so <= 22));
            l.tc32 = false;
            const bool tcn_shape = !ctx->force_generic && e == 2 && l.d.c > 4 && (l.d.size == 1 || l.d.size == 3) && l.d.stride == 1 &&
                                   l.d.pad == l.d.size / 2 && l.d.out_w == l.d.w && l.d.out_h == l.d.h &&
                                   conv_i16_tc32_eligible(l.d.w, l.d.size, so, ctx->Tn);
            // Tn = 16 / 8 builds have a CUDA-core kernel too (2 / 4 C4 words per step, 3.0-3.4 T step equivalents/s on every layer);
            // the tensor-core kernel is one work item per CTA (prologue + pipeline fill), so it takes the layers whose CTAs run
            // enough K slices on full enough 128-channel tiles.  Measured per layer with the kernel forced on / off
            // (profiles/r2_layer_table_int16_b32_tn{16,8}_{tc2,tc0}.json): it wins from (K slices x tile fill) = 8 for Tn = 16
            // (256->128 1x1: 0.21 against 0.27 ms) and from 16 for Tn = 8 (512->256 1x1: 0.25 against 0.29 ms), and loses below
            // (32->64 3x3 @208, Tn = 16: 3.1 against 2.0 ms).  YOLO2CUDA_TC=2: every eligible layer, =0: none.
            bool tcn = false;
            if (tcn_shape && (ctx->Tn == 16 || ctx->Tn == 8) && l.fast && ctx->use_tc != 0) {
                const int slices = conv_i16_tc32_slices(l.d.c, l.d.size, ctx->Tn);
                const bool fill80 = l.d.n * 5 >= ceil_div(l.d.n, 128) * 128 * 4;      // 1x1 layers: the staging warps bound the kernel, a half-empty
                                                                                      // tile loses (512->64 1x1 @26, Tn = 16: 0.119 against 0.097 ms)
                tcn = ctx->use_tc > 0 || ((long long)slices * l.d.n >= (long long)(ctx->Tn == 16 ? 8 : 16) * ceil_div(l.d.n, 128) * 128 &&
                                          (l.d.size == 3 || fill80));
            }
            if (tcn_shape && ((!l.fast && ctx->Tn == 32) || tcn)) {
                // reference built with Tn = 32 / 16 / 8: tensor-core kernel, one MMA K slice per 1 / 2 / 4 rounding groups
                l.fast = false;
                ConvFastParams p{};
                p.B = net->max_batch; p.H = l.d.h; p.W = l.d.w; p.G = ceil_div(l.d.c, 4); p.OFM = l.d.n;
                if (!l.w_tc && (rc = net_alloc(net, &l.w_tc, wprep_tc32_bytes(l.d.c, l.d.n, l.d.size, ctx->Tn)))) return rc;
                launch_wprep_tc32((const int16_t *)net->d_wblob + l.w_off, l.w_tc, l.d.c, l.d.n, l.d.size, TM, TN, ctx->Tn, st);
                ctx->launches += 1;
                p.in = l.in.base; p.out = l.out.base; p.w = l.w_tc;
                p.bias = (char *)net->d_bblob + l.b_off * e;
                p.in_frame_stride = l.in.frame_stride; p.out_frame_stride = l.out.frame_stride;
                p.so = so; p.sb = l.Qb - l.Qa_out; p.leaky = l.d.leaky;
                l.cp = p;
                l.tc32 = true;
            }
            if (l.fast) {
                ConvFastParams p{};
                p.B = net->max_batch; p.H = l.d.h; p.W = l.d.w; p.G = ceil_div(l.d.c, 4); p.OFM = l.d.n;
                // rounding group of 4 channels (reference default, or a layer with <= 4 input channels under any Tn): every int16 kernel;
                // Tn = 8 / 16 builds: 2 / 4 C4 words per step in the CUDA-core kernel (the layers the variant tensor-core kernel did not take above)
                const bool group4 = e == 4 || TN == 4 || l.d.c <= 4;
                p.group_words = group4 ? 1 : ctx->Tn / 4;
                if (conv_fast_plan(p, l.d.size, e) == 0) l.fast = false;
                else {
                    if (!l.w_dev && (rc = net_alloc(net, &l.w_dev, wprep_bytes(l.d.c, l.d.n, l.d.size, e)))) return rc;
                    if (e == 2) launch_wprep_i16((const int16_t *)net->d_wblob + l.w_off, l.w_dev, l.d.c, l.d.n, l.d.size, TM, TN, st);
                    else launch_wprep_f32((const float *)net->d_wblob + l.w_off, l.w_dev, l.d.c, l.d.n, l.d.size, TM, TN, st);
                    ctx->launches += 1;
                    p.in = l.in.base; p.out = l.out.base; p.w = l.w_dev;
                    p.bias = (char *)net->d_bblob + l.b_off * e;
                    p.in_frame_stride = l.in.frame_stride; p.out_frame_stride = l.out.frame_stride;
                    p.so = so > 30 ? 30 : so; p.sb = l.Qb - l.Qa_out; p.leaky = l.d.leaky;
                    l.cp = p;
                    // tensor-core path: wide layers only (a CTA covers 128 output channels)
                    l.tc = ctx->use_tc > 0 && e == 2 && group4 && so >= 8 && so <= 22 && l.d.n >= ctx->tc_min_ofm;
                    // auto: the persistent tcgen05 kernel wherever its 128-channel tiles are at least 80 % full and the chain is deep
                    // enough to feed it (measured, profiles/r2_layer_table_int16_b128_persistent*.json: 5.1-5.6 T steps/s on every 3x3
                    // layer from 104 to 13 wide and 4.1-4.6 T on the 1x1 layers, against 2.6-3.8 T on the CUDA cores); layer 0 (3 input
                    // channels, 32 output channels) stays on the CUDA-core kernel.  Both paths are bit-exact, so mixing them is safe.
                    // (128-channel tiles, or 64-channel tiles x two pixel sets when those fill better: conv_i16_tc2.cu half_mode)
                    const bool fill128 = l.d.n * 5 >= ceil_div(l.d.n, 128) * 128 * 4, fill64 = l.d.n * 5 >= ceil_div(l.d.n, 64) * 64 * 4;
                    if (ctx->use_tc < 0 && e == 2 && group4 && so >= 8 && so <= 22 && (fill128 || fill64) &&
                        ((l.d.size == 3 && l.d.c >= 32) || (l.d.size == 1 && l.d.c >= 128)) &&
                        conv_i16_tc2_eligible(p, l.d.size, net->max_batch))
                        l.tc = true;
                    if (l.tc) {
                        if (!l.w_tc && (rc = net_alloc(net, &l.w_tc, wprep_tc2_bytes(l.d.c, l.d.n, l.d.size)))) return rc;
                        launch_wprep_tc2((const int16_t *)net->d_wblob + l.w_off, l.w_tc, l.d.c, l.d.n, l.d.size, TM, TN, so, st);
                        ctx->launches += 1;
                    }
                    // a 3x3 layer with one input channel group (YOLOv2's first layer): the specialised kernel, which can also do the
                    // 2x2 / stride-2 max-pool that follows in its store (only when nothing else reads this layer's output)
                    l.g1 = !l.tc && e == 2 && ctx->use_g1 && l.d.size == 3 && l.d.c <= 4;
                    l.pool_fusable = false;
                    if (l.g1 && i + 1 < net->L.size()) {
                        const yolo2cuda_layer_desc &nx = net->L[i + 1].d;
                        l.pool_fusable = nx.type == YOLO2CUDA_MAXPOOL && nx.size == 2 && nx.stride == 2 && !(l.d.out_h & 1) && !(l.d.out_w & 1) &&
                                         nx.out_h == l.d.out_h / 2 && nx.out_w == l.d.out_w / 2 && net->tensor_in[i + 1] == net->tensor_out[i] &&
                                         net->tensor_out[i] >= 0 && net->T[net->tensor_out[i]].last_use == (int)i + 1 && net->concat_of[i] < 0;
                    }
                }
            }
        } else if (l.d.type == YOLO2CUDA_REORG) {
            l.reorg_shift = 0;
            if (e == 2 && route_q > 0) {
                int target = route_q < current_qa ? route_q : current_qa;
                l.reorg_shift = current_qa - target;
                if (l.reorg_shift != 0) current_qa = target;
                pending_route_q = current_qa;
            }
        } else if (l.d.type == YOLO2CUDA_REGION) {
            l.region_q = current_qa;
            net->region_q = current_qa;
        }
    }
    // ---- largest-|activation| slots (int16): tensors joined by max-pool / reorg / concat form one class ----
    if (e == 2) {
        const int nt = (int)net->T.size();
        std::vector<int> parent(nt);
        for (int t = 0; t < nt; ++t) parent[t] = t;
        auto find = [&](int t) { while (parent[t] != t) t = parent[t] = parent[parent[t]]; return t; };
        for (size_t i = 0; i < net->L.size(); ++i) {
            const int ty = net->L[i].d.type;
            if ((ty == YOLO2CUDA_MAXPOOL || ty == YOLO2CUDA_REORG) && net->tensor_in[i] >= 0 && net->tensor_out[i] >= 0)
                parent[find(net->tensor_out[i])] = find(net->tensor_in[i]);
        }
        net->xslot.assign(nt, 0);
        net->xknown.assign(nt, 1);
        for (int t = 0; t < nt; ++t) net->xslot[t] = find(t);
        net->xknown[find(0)] = 0;                          // the quantised input frames: no maximum is tracked
        for (size_t i = 0; i < net->L.size(); ++i) {
            const LayerPlan &l = net->L[i];
            if (l.d.type != YOLO2CUDA_CONV || net->tensor_out[i] < 0) continue;
            const bool tracks = l.fast && !l.tc32;          // conv_i16_c4_kernel and conv_i16_tc2_kernel leave their maximum behind
            if (!tracks) net->xknown[net->xslot[net->tensor_out[i]]] = 0;
        }
        for (size_t i = 0; i < net->L.size(); ++i) {
            LayerPlan &l = net->L[i];
            if (l.d.type != YOLO2CUDA_CONV) continue;
            const int si = net->xslot[net->tensor_in[i]], so_ = net->xslot[net->tensor_out[i]];
            l.cp.xmax_in = net->xknown[si] ? net->d_xmax + si : nullptr;
            l.cp.xmax_out = net->d_xmax + so_;
            l.cp.tc_stats = ctx->d_tc_stats;
            l.cp.tc_force_exact = ctx->tc_force_exact;
        }
    }
    CUDA_OK(ctx, cudaGetLastError());
    CUDA_OK(ctx, cudaStreamSynchronize(st));
    net->weights_loaded = true;
    return YOLO2CUDA_SUCCESS;
}

int yolo2cuda_net_forward_dev(yolo2cuda_net *net, const float *frames, int batch, float *region_out)
{
    if (!net || !frames || !region_out || batch <= 0) return YOLO2CUDA_ERROR;
    yolo2cuda_ctx *ctx = net->ctx;
    if (!net->weights_loaded) return fail(ctx, YOLO2CUDA_ERROR, "weights not loaded");
    CUDA_OK(ctx, cudaSetDevice(ctx->device));
    const size_t frame_elems = (size_t)net->in_c * net->in_h * net->in_w;
    for (int b0 = 0; b0 < batch; b0 += net->max_batch) {
        int B = batch - b0 < net->max_batch ? batch - b0 : net->max_batch;
        int rc = forward_chunk(net, frames + (size_t)b0 * frame_elems, B, region_out + (size_t)b0 * net->region_outputs);
        if (rc) return rc;
    }
    return YOLO2CUDA_SUCCESS;
}

int yolo2cuda_net_forward_host(yolo2cuda_net *net, const float *frames, int batch, float *region_out)
{
    if (!net || !frames || !region_out || batch <= 0) return YOLO2CUDA_ERROR;
    yolo2cuda_ctx *ctx = net->ctx;
    if (!net->weights_loaded) return fail(ctx, YOLO2CUDA_ERROR, "weights not loaded");
    CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t frame_elems = (size_t)net->in_c * net->in_h * net->in_w;
    int rc;
    // Double-buffered passes: the H2D copy of pass k+1 and the D2H copy of pass k-1 run on their own
    // streams (separate DMA engines) while pass k computes on the context stream.
    // lazily created staging, one guard per resource: a call that fails half-way neither leaks nor re-creates what exists
    for (int i = 0; i < 2; ++i) {
        if (!net->d_frames2[i] && (rc = net_alloc(net, &net->d_frames2[i], frame_elems * net->max_batch * sizeof(float)))) return rc;
        if (!net->d_region2[i] && (rc = net_alloc(net, &net->d_region2[i], net->region_outputs * net->max_batch * sizeof(float)))) return rc;
        if (!net->ev_h2d[i]) CUDA_OK(ctx, cudaEventCreateWithFlags(&net->ev_h2d[i], cudaEventDisableTiming));
        if (!net->ev_comp[i]) CUDA_OK(ctx, cudaEventCreateWithFlags(&net->ev_comp[i], cudaEventDisableTiming));
        if (!net->ev_d2h[i]) CUDA_OK(ctx, cudaEventCreateWithFlags(&net->ev_d2h[i], cudaEventDisableTiming));
    }
    if (!net->s_h2d) CUDA_OK(ctx, cudaStreamCreateWithFlags(&net->s_h2d, cudaStreamNonBlocking));
    if (!net->s_d2h) CUDA_OK(ctx, cudaStreamCreateWithFlags(&net->s_d2h, cudaStreamNonBlocking));
    // Pass schedule.  Only the first pass's upload and the last pass's download are exposed (everything else overlaps a pass's
    // compute), so a batch of more than one pass starts with a SHORT ramp pass: at 364-frame passes the exposed upload shrinks
    // from 14 ms to 2 ms per call for ~1 ms of extra tail effect in the ramp pass's kernels.
    std::vector<int> poff;            // first frame of pass k; poff[npass] = batch
    {
        int done = 0;
        poff.push_back(0);
        if (batch > net->max_batch) {
            const int ramp = net->ramp_frames >= 0 ? net->ramp_frames : std::max(32, net->max_batch / 6);
            if (ramp > 0 && ramp < net->max_batch) { done = ramp; poff.push_back(done); }
        }
        while (done < batch) { done = std::min(batch, done + net->max_batch); poff.push_back(done); }
    }
    const int npass = (int)poff.size() - 1;
    auto pass_frames = [&](int k) { return poff[k + 1] - poff[k]; };
    auto issue_h2d = [&](int k) -> int {
        const int buf = k & 1;
        if (k >= 2) CUDA_OK(ctx, cudaStreamWaitEvent(net->s_h2d, net->ev_comp[buf], 0));   // pass k-2 has consumed this buffer
        CUDA_OK(ctx, cudaMemcpyAsync(net->d_frames2[buf], frames + (size_t)poff[k] * frame_elems,
                                     frame_elems * pass_frames(k) * sizeof(float), cudaMemcpyHostToDevice, net->s_h2d));
        CUDA_OK(ctx, cudaEventRecord(net->ev_h2d[buf], net->s_h2d));
        return YOLO2CUDA_SUCCESS;
    };
    // order the side streams after whatever the caller queued on the context stream before this call
    CUDA_OK(ctx, cudaEventRecord(net->ev_comp[0], st));
    CUDA_OK(ctx, cudaStreamWaitEvent(net->s_h2d, net->ev_comp[0], 0));
    if ((rc = issue_h2d(0))) return rc;
    for (int k = 0; k < npass; ++k) {
        const int buf = k & 1, B = pass_frames(k);
        if (k + 1 < npass && (rc = issue_h2d(k + 1))) return rc;
        CUDA_OK(ctx, cudaStreamWaitEvent(st, net->ev_h2d[buf], 0));
        if (k >= 2) CUDA_OK(ctx, cudaStreamWaitEvent(st, net->ev_d2h[buf], 0));               // region buffer drained
        if ((rc = forward_chunk(net, (const float *)net->d_frames2[buf], B, (float *)net->d_region2[buf]))) return rc;
        CUDA_OK(ctx, cudaEventRecord(net->ev_comp[buf], st));
        CUDA_OK(ctx, cudaStreamWaitEvent(net->s_d2h, net->ev_comp[buf], 0));
        CUDA_OK(ctx, cudaMemcpyAsync(region_out + (size_t)poff[k] * net->region_outputs, net->d_region2[buf],
                                     net->region_outputs * B * sizeof(float), cudaMemcpyDeviceToHost, net->s_d2h));
        CUDA_OK(ctx, cudaEventRecord(net->ev_d2h[buf], net->s_d2h));
    }
    CUDA_OK(ctx, cudaStreamSynchronize(net->s_d2h));
    CUDA_OK(ctx, cudaStreamSynchronize(st));
    return YOLO2CUDA_SUCCESS;
}

int yolo2cuda_selftest_exp_dev(yolo2cuda_ctx *ctx, const double *x, double *y, size_t n)
{
    if (!ctx || !x || !y || n == 0) return YOLO2CUDA_ERROR;
    CUDA_OK(ctx, cudaSetDevice(ctx->device));
    launch_glibc_exp(x, y, (long long)n, ctx->stream);
    ctx->launches += 1;
    CUDA_OK(ctx, cudaGetLastError());
    return YOLO2CUDA_SUCCESS;
}

int yolo2cuda_letterbox_dev(yolo2cuda_ctx *ctx, const unsigned char *src, int batch, int iw, int ih, int ic, float *dst,
                            int net_w, int net_h)
{
    if (!ctx || !src || !dst) return YOLO2CUDA_ERROR;
    if (batch <= 0 || iw <= 0 || ih <= 0 || ic <= 0 || net_w <= 0 || net_h <= 0) return fail(ctx, YOLO2CUDA_ERROR, "letterbox: bad dimensions");
    CUDA_OK(ctx, cudaSetDevice(ctx->device));
    launch_letterbox(src, dst, batch, iw, ih, ic, net_w, net_h, ctx->stream);
    ctx->launches += 1;
    CUDA_OK(ctx, cudaGetLastError());
    return YOLO2CUDA_SUCCESS;
}

int yolo2cuda_net_forward_images_host(yolo2cuda_net *net, const unsigned char *images, int batch, int iw, int ih, float *region_out)
{
    if (!net || !images || !region_out || batch <= 0 || iw <= 0 || ih <= 0) return YOLO2CUDA_ERROR;
    yolo2cuda_ctx *ctx = net->ctx;
    if (!net->weights_loaded) return fail(ctx, YOLO2CUDA_ERROR, "weights not loaded");
    CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t frame_elems = (size_t)net->in_c * net->in_h * net->in_w, img_bytes = (size_t)iw * ih * net->in_c;
    int rc;
    if (!net->d_lb_frames) {
        if ((rc = net_alloc(net, &net->d_lb_frames, frame_elems * net->max_batch * sizeof(float)))) return rc;
        if ((rc = net_alloc(net, &net->d_lb_region, net->region_outputs * net->max_batch * sizeof(float)))) return rc;
    }
    if (net->lb_img_bytes < img_bytes * net->max_batch) {
        if (net->d_lb_img) cudaFree(net->d_lb_img);
        net->d_lb_img = nullptr;
        CUDA_OK(ctx, cudaMalloc(&net->d_lb_img, img_bytes * net->max_batch));
        net->lb_img_bytes = img_bytes * net->max_batch;
    }
    for (int b0 = 0; b0 < batch; b0 += net->max_batch) {
        const int B = batch - b0 < net->max_batch ? batch - b0 : net->max_batch;
        CUDA_OK(ctx, cudaMemcpyAsync(net->d_lb_img, images + (size_t)b0 * img_bytes, img_bytes * B, cudaMemcpyHostToDevice, st));
        launch_letterbox((const unsigned char *)net->d_lb_img, (float *)net->d_lb_frames, B, iw, ih, net->in_c, net->in_w, net->in_h, st);
        ctx->launches += 1;
        if ((rc = forward_chunk(net, (const float *)net->d_lb_frames, B, (float *)net->d_lb_region))) return rc;
        CUDA_OK(ctx, cudaMemcpyAsync(region_out + (size_t)b0 * net->region_outputs, net->d_lb_region, net->region_outputs * B * sizeof(float),
                                     cudaMemcpyDeviceToHost, st));
    }
    CUDA_OK(ctx, cudaStreamSynchronize(st));
    return YOLO2CUDA_SUCCESS;
}

int yolo2cuda_net_get_layer_output(yolo2cuda_net *net, int layer, int frame, void *dst, size_t dst_elems)
{
    if (!net || !dst || layer < 0 || layer >= (int)net->L.size()) return YOLO2CUDA_ERROR;
    yolo2cuda_ctx *ctx = net->ctx;
    const LayerPlan &l = net->L[layer];
    if (l.d.type == YOLO2CUDA_REGION) return fail(ctx, YOLO2CUDA_ERROR, "region output is returned by net_forward");
    if (!net->keep_all)
        return fail(ctx, YOLO2CUDA_ERROR, "per-layer outputs are recycled in the compact arena: call yolo2cuda_net_set_debug_keep(net, 1) before the forward");
    if (frame < 0 || frame >= net->last_batch) return fail(ctx, YOLO2CUDA_ERROR, "frame %d not in the last forward (batch %d)", frame, net->last_batch);
    const size_t need = planar_elems(l.out.C, l.out.H, l.out.W);
    if (dst_elems < need) return fail(ctx, YOLO2CUDA_ERROR, "dst too small: need %zu elements", need);
    CUDA_OK(ctx, cudaSetDevice(ctx->device));
    const int e = ctx->elem;
    cudaStream_t st = ctx->stream;
    CUDA_OK(ctx, cudaMemsetAsync(net->d_tmp_planar_out, 0, need * e, st));
    launch_c4_to_planar((char *)l.out.base + (size_t)frame * l.out.frame_stride * e, net->d_tmp_planar_out, 1, l.out.C, l.out.H,
                        l.out.W, 0, 0, e, st);
    ctx->launches += 1;
    CUDA_OK(ctx, cudaMemcpyAsync(dst, net->d_tmp_planar_out, need * e, cudaMemcpyDeviceToHost, st));
    CUDA_OK(ctx, cudaStreamSynchronize(st));
    return YOLO2CUDA_SUCCESS;
}

int yolo2cuda_net_region_q(const yolo2cuda_net *net) { return net ? net->region_q : 0; }
uint64_t yolo2cuda_net_launches_per_forward(const yolo2cuda_net *net) { return net ? net->launches_per_forward : 0; }
int yolo2cuda_net_set_ramp_frames(yolo2cuda_net *net, int frames)
{
    if (!net || frames < -1) return YOLO2CUDA_ERROR;
    net->ramp_frames = frames;
    return YOLO2CUDA_SUCCESS;
}

int yolo2cuda_net_set_debug_keep(yolo2cuda_net *net, int keep)
{
    if (!net) return YOLO2CUDA_ERROR;
    CUDA_OK(net->ctx, cudaSetDevice(net->ctx->device));
    if ((keep != 0) == net->keep_all) return YOLO2CUDA_SUCCESS;
    net->last_batch = 0;                       // the tensors move: nothing of the last forward can be read back any more
    return place_tensors(net, keep != 0);
}

size_t yolo2cuda_net_activation_bytes(const yolo2cuda_net *net)
{
    if (!net) return 0;
    if (!net->keep_all) return net->arena_bytes;
    size_t s = 0;
    for (const auto &t : net->T) s += t.bytes;
    return s;
}

const char *yolo2cuda_net_layer_kernel(const yolo2cuda_net *net, int layer)
{
    if (!net || layer < 0 || layer >= (int)net->L.size()) return "";
    const LayerPlan &l = net->L[layer];
    return l.variant ? l.variant : "";
}

int yolo2cuda_net_layer_times(yolo2cuda_net *net, float *ms, int n_layers)
{
    if (!net || !ms) return YOLO2CUDA_ERROR;
    yolo2cuda_ctx *ctx = net->ctx;
    CUDA_OK(ctx, cudaSetDevice(ctx->device));
    if (!net->timing) {
        net->ev.resize(net->L.size() + 1);
        for (auto &ev : net->ev) CUDA_OK(ctx, cudaEventCreate(&ev));
        net->timing = true;
        for (int i = 0; i < n_layers; ++i) ms[i] = -1.0f;
        return YOLO2CUDA_SUCCESS;  // armed; numbers are available after the next forward
    }
    CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < n_layers && i < (int)net->L.size(); ++i) {
        float t = 0;
        if (cudaEventElapsedTime(&t, net->ev[i], net->ev[i + 1]) != cudaSuccess) { cudaGetLastError(); t = -1.0f; }
        ms[i] = t;
    }
    return YOLO2CUDA_SUCCESS;
}

}  // extern "C"
