// Bandwidth-bound operators of the YOLOv2 datapath: layout converters, 2x2 max-pool
// (hls/core/core_compute.cpp:266-305), the HLS reorg tile (:354-379), and the three operators
// the reference driver runs on the host CPU - input quantiser (yolo2_model.cpp:257-273),
// flat-memory reorg + Q alignment (:112-129,358-401) and the region head (:406-425 +
// src/core/yolo_region.cpp:123-141, src/core/yolo_math.cpp:19,226-250).
// All are HBM-bound: every thread moves one 8/16-byte C4 pixel word or one coalesced planar row
// element; nothing is re-read.
#include <cfloat>

#include <math_constants.h>
#include "common.cuh"
#include "glibc_exp_data.h"

namespace y2 {

namespace {

template <typename T> struct Vec4;
template <> struct Vec4<int16_t> { using type = uint2; };
template <> struct Vec4<float> { using type = float4; };

template <typename T> __device__ __forceinline__ T pool_floor();
template <> __device__ __forceinline__ int16_t pool_floor<int16_t>() { return (int16_t)-32768; }   // core_io.cpp:96-99
template <> __device__ __forceinline__ float pool_floor<float>() { return (float)(-1024 * 1024); }  // core_io.cpp:100-102

template <typename T>
__global__ void planar_to_c4_kernel(const T *__restrict__ src, T *__restrict__ dst, int B, int C, int H, int W,
                                    long long sfs, long long dfs)
{
    const int G = ceil_div(C, 4), Wa = align8(W);
    const long long total = (long long)B * G * H * W;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    int x = idx % W;
    long long r = idx / W;
    int y = r % H; r /= H;
    int g = r % G;
    int f = r / G;
    T v[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        int c = g * 4 + t;
        v[t] = (c < C) ? src[f * sfs + ((long long)c * H + y) * Wa + x] : (T)0;
    }
    T *d = dst + f * dfs + (((long long)g * H + y) * W + x) * 4;
    *reinterpret_cast<typename Vec4<T>::type *>(d) = *reinterpret_cast<typename Vec4<T>::type *>(v);
}

template <typename T>
__global__ void c4_to_planar_kernel(const T *__restrict__ src, T *__restrict__ dst, int B, int C, int H, int W,
                                    long long sfs, long long dfs)
{
    const int G = ceil_div(C, 4), Wa = align8(W);
    const long long total = (long long)B * G * H * W;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    int x = idx % W;
    long long r = idx / W;
    int y = r % H; r /= H;
    int g = r % G;
    int f = r / G;
    T v[4];
    *reinterpret_cast<typename Vec4<T>::type *>(v) =
        *reinterpret_cast<const typename Vec4<T>::type *>(src + f * sfs + (((long long)g * H + y) * W + x) * 4);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        int c = g * 4 + t;
        if (c < C) dst[f * dfs + ((long long)c * H + y) * Wa + x] = v[t];
    }
}

template <typename T>
__global__ void maxpool_planar_kernel(const T *__restrict__ in, T *__restrict__ out, int ch, int ksize, int kstride,
                                      int iw, int ih, int ow, int oh)
{
    const int iwa = align8(iw), owa = align8(ow);
    const long long total = (long long)ch * oh * ow;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    int x = idx % ow;
    long long r = idx / ow;
    int y = r % oh;
    int c = r / oh;
    T best = pool_floor<T>();
    for (int i = 0; i < ksize; ++i)
        for (int j = 0; j < ksize; ++j) {
            int iy = y * kstride + i, ix = x * kstride + j;  // loader Padding forced to 0, core_scheduler.cpp:72-73
            T v = (iy < ih && ix < iw) ? in[((long long)c * ih + iy) * iwa + ix] : pool_floor<T>();
            if (v > best) best = v;
        }
    out[((long long)c * oh + y) * owa + x] = best;
}

template <typename T>
__global__ void maxpool_c4_kernel(const T *__restrict__ in, T *__restrict__ out, int B, int G, int kstride, int iw,
                                  int ih, int ow, int oh, long long ifs, long long ofs)
{
    const long long total = (long long)B * G * oh * ow;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    int x = idx % ow;
    long long r = idx / ow;
    int y = r % oh; r /= oh;
    int g = r % G;
    int f = r / G;
    T best[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) best[t] = pool_floor<T>();
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            int iy = y * kstride + i, ix = x * kstride + j;
            if (iy < ih && ix < iw) {
                T v[4];
                *reinterpret_cast<typename Vec4<T>::type *>(v) = *reinterpret_cast<const typename Vec4<T>::type *>(
                    in + f * ifs + (((long long)g * ih + iy) * iw + ix) * 4);
#pragma unroll
                for (int t = 0; t < 4; ++t)
                    if (v[t] > best[t]) best[t] = v[t];
            }
        }
    *reinterpret_cast<typename Vec4<T>::type *>(out + f * ofs + (((long long)g * oh + y) * ow + x) * 4) =
        *reinterpret_cast<typename Vec4<T>::type *>(best);
}

// The network's five pools (2x2, stride 2, even dims, int16): the two window pixels of a row are one aligned 16-byte word, so a
// thread does two LDG.128 and one STG.64 and a warp reads 512 contiguous bytes per row.
__device__ __forceinline__ unsigned vmax_s16x2(unsigned a, unsigned b) { return __vmaxs2(a, b); }
__global__ void maxpool_c4_i16_s2_kernel(const int16_t *__restrict__ in, int16_t *__restrict__ out, long long total, int G, int iw, int ih,
                                         int ow, int oh, long long ifs, long long ofs)
{
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    int x = idx % ow;
    long long r = idx / ow;
    int y = r % oh; r /= oh;
    int g = r % G;
    long long f = r / G;
    const int16_t *src = in + f * ifs + (((long long)g * ih + 2 * y) * iw + 2 * x) * 4;
    const uint4 a = *reinterpret_cast<const uint4 *>(src), b = *reinterpret_cast<const uint4 *>(src + (long long)iw * 4);
    uint2 m;
    m.x = vmax_s16x2(vmax_s16x2(a.x, a.z), vmax_s16x2(b.x, b.z));
    m.y = vmax_s16x2(vmax_s16x2(a.y, a.w), vmax_s16x2(b.y, b.w));
    *reinterpret_cast<uint2 *>(out + f * ofs + (((long long)g * oh + y) * ow + x) * 4) = m;
}

// LayerType 2: out[m+2ky+kx][y][x] = in[m][2y+ky][2x+kx] for m stepping by TM (core_compute.cpp:354-379,
// core_scheduler.cpp:88-112, yolo2_accel.cpp:127-169).
template <typename T>
__global__ void reorg_hls_planar_kernel(const T *__restrict__ in, T *__restrict__ out, int ch, int TM, int iw, int ih,
                                        int ow, int oh)
{
    const int iwa = align8(iw), owa = align8(ow);
    const long long total = (long long)ch * oh * ow;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    int x = idx % ow;
    long long r = idx / ow;
    int y = r % oh;
    int oc = r / oh;
    int m = (oc / TM) * TM, q = oc - m;
    if (q >= 4) return;  // only four phases exist; further tile channels are never written
    int ky = q >> 1, kx = q & 1;
    int iy = 2 * y + ky, ix = 2 * x + kx;
    T v = (iy < ih && ix < iw) ? in[((long long)m * ih + iy) * iwa + ix] : (T)0;
    out[((long long)oc * oh + y) * owa + x] = v;
}

// sat16(llround(in * 2^Q)) of the driver's input quantiser (yolo2_model.cpp:265-271: clamp in float, llround = round half AWAY
// from zero, clamp again) without the 64-bit software llroundf: v - trunc(v) is exact in binary32, so the tie test is exact.
// NaN: the reference's float clamps pass it through and llround(NaN) is LLONG_MIN on the hosts it runs on -> -32768; fmaxf drops
// the NaN operand, which gives the same -32768.
__device__ __forceinline__ int16_t quantize_one(float in, float scale)
{
    float v = in * scale;
    v = fminf(fmaxf(v, -32768.f), 32767.f);
    const float t = truncf(v);
    const float r = t + ((fabsf(v - t) >= 0.5f) ? copysignf(1.0f, v) : 0.0f);
    return (int16_t)__float2int_rz(r);
}

__global__ void quantize_kernel(const float *__restrict__ in, int16_t *__restrict__ out, size_t count, float scale)
{
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < count) out[idx] = quantize_one(in[idx], scale);
}

template <typename T>
__global__ void frames_to_c4_kernel(const float *__restrict__ frames, T *__restrict__ dst, int B, int C, int H, int W,
                                    long long dfs, float scale)
{
    const int G = ceil_div(C, 4);
    const long long total = (long long)B * G * H * W;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    int x = idx % W;
    long long r = idx / W;
    int y = r % H; r /= H;
    int g = r % G;
    int f = r / G;
    T v[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        int c = g * 4 + t;
        float s = (c < C) ? frames[(((long long)f * C + c) * H + y) * W + x] : 0.0f;
        if constexpr (sizeof(T) == 2) v[t] = (c < C) ? quantize_one(s, scale) : (int16_t)0;
        else v[t] = s;
    }
    *reinterpret_cast<typename Vec4<T>::type *>(dst + f * dfs + (((long long)g * H + y) * W + x) * 4) =
        *reinterpret_cast<typename Vec4<T>::type *>(v);
}

// int16, C <= 4, W % 4 == 0, 16-byte aligned frames: one thread = four consecutive pixels of one row - a 16-byte load per channel
// plane (a warp reads 512 contiguous bytes of each plane) and two 16-byte stores (the warp writes 1 KB contiguous)
__global__ void frames_to_c4_i16_v4_kernel(const float *__restrict__ frames, int16_t *__restrict__ dst, int B, int C, int H, int W,
                                           long long dfs, float scale)
{
    const int W4 = W >> 2;
    const long long total = (long long)B * H * W4;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int x4 = idx % W4;
    long long r = idx / W4;
    const int y = r % H;
    const int f = r / H;
    float4 in[4];
#pragma unroll
    for (int c = 0; c < 4; ++c)
        in[c] = (c < C) ? __ldg(reinterpret_cast<const float4 *>(frames + (((long long)f * C + c) * H + y) * W) + x4) : make_float4(0.f, 0.f, 0.f, 0.f);
    uint2 px[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        unsigned q[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float sv = k == 0 ? in[c].x : k == 1 ? in[c].y : k == 2 ? in[c].z : in[c].w;
            q[c] = (c < C) ? (unsigned)(unsigned short)quantize_one(sv, scale) : 0u;
        }
        px[k] = make_uint2(q[0] | (q[1] << 16), q[2] | (q[3] << 16));
    }
    uint4 *o = reinterpret_cast<uint4 *>(dst + f * dfs + ((long long)y * W + 4 * x4) * 4);
    o[0] = make_uint4(px[0].x, px[0].y, px[1].x, px[1].y);
    o[1] = make_uint4(px[2].x, px[2].y, px[3].x, px[3].y);
}

// Source coordinates of the flat-memory reorg (yolo2_model.cpp:112-129 called as
// reorg_cpu(x, W, H*C/4, 4, 2, out), :373): compact output flat index f -> compact source index.
__device__ __forceinline__ void reorg_src(long long f, int c, int h, int w, int &cs, int &ys, int &xs)
{
    const long long hc = (long long)h * c / 4;
    long long i = f % w, jj = f / w, j = jj % hc, k = jj / hc;
    long long s = (2 * i + (k & 1)) + 2LL * w * (2 * j + (k >> 1));
    xs = (int)(s % w);
    long long row = s / w;
    ys = (int)(row % h);
    cs = (int)(row / h);
}

template <typename T> __device__ __forceinline__ T reorg_shift(T v, int shift);
template <> __device__ __forceinline__ int16_t reorg_shift<int16_t>(int16_t v, int shift)
{
    int t = (int)v;
    if (shift > 0) t >>= shift;  // arithmetic, no rounding (yolo2_model.cpp:386-391)
    return (int16_t)t;
}
template <> __device__ __forceinline__ float reorg_shift<float>(float v, int) { return v; }

template <typename T>
__global__ void reorg_driver_planar_kernel(const T *__restrict__ in, T *__restrict__ out, int c, int h, int w, int shift)
{
    const int ow = w / 2, oh = h / 2, owa = align8(ow), wa = align8(w);
    const long long total = (long long)4 * c * oh * owa;  // includes pad columns, which are zeroed (:375)
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    int x = idx % owa;
    long long row = idx / owa;
    if (x >= ow) { out[idx] = (T)0; return; }
    long long f = row * ow + x;
    int cs, ys, xs;
    reorg_src(f, c, h, w, cs, ys, xs);
    out[idx] = reorg_shift<T>(in[((long long)cs * h + ys) * wa + xs], shift);
}

template <typename T>
__global__ void reorg_driver_c4_kernel(const T *__restrict__ in, T *__restrict__ out, int B, int c, int h, int w,
                                       int shift, long long ifs, long long ofs)
{
    const int ow = w / 2, oh = h / 2, OG = c;  // 4c output channels = c output groups
    const long long total = (long long)B * OG * oh * ow;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    int x = idx % ow;
    long long r = idx / ow;
    int y = r % oh; r /= oh;
    int g = r % OG;
    int f = r / OG;
    T v[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        int oc = g * 4 + t;
        long long fl = ((long long)oc * oh + y) * ow + x;
        int cs, ys, xs;
        reorg_src(fl, c, h, w, cs, ys, xs);
        v[t] = reorg_shift<T>(in[f * ifs + (((long long)(cs >> 2) * h + ys) * w + xs) * 4 + (cs & 3)], shift);
    }
    *reinterpret_cast<typename Vec4<T>::type *>(out + f * ofs + (((long long)g * oh + y) * ow + x) * 4) =
        *reinterpret_cast<typename Vec4<T>::type *>(v);
}

// glibc's double exp restated (sysdeps/ieee754/dbl-64/e_exp.c of glibc >= 2.28, N = 128 table + degree-5 polynomial, in the FMA form
// the x86-64 ifunc selects on every CPU with FMA; data in glibc_exp_data.h, generated from the host's libm by
// gen_glibc_exp_data.py).  The reference's logistic_activate and softmax call exp(double) (yolo_math.cpp:19,234): with this
// the region head reproduces the host's bits by construction (the checker's exp_check.c: 0 differences against libm on 10^8 inputs
// incl. arbitrary bit patterns); CUDA's own exp may differ by 1 ulp.  Every operation is one explicitly rounded instruction.
__device__ const unsigned long long g_glibc_exp_tab[256] = {Y2_GLIBC_EXP_TAB};
__device__ __forceinline__ double glibc_exp(double x)
{
    unsigned abstop = (unsigned)((unsigned long long)__double_as_longlong(x) >> 52) & 0x7ffu;
    if (abstop - 0x3c9u >= 0x3fu) {                        // |x| < 2^-54, |x| >= 512, inf or NaN
        if (abstop - 0x3c9u >= 0x80000000u) return __dadd_rn(1.0, x);
        if (abstop >= 0x409u) {                            // |x| >= 1024
            if (__double_as_longlong(x) == __double_as_longlong(-CUDART_INF)) return 0.0;
            if (abstop >= 0x7ffu) return __dadd_rn(1.0, x);
            return (__double_as_longlong(x) < 0) ? 0.0 : CUDART_INF;       // __math_uflow / __math_oflow
        }
        abstop = 0;
    }
    double kd = __fma_rn(Y2_GLIBC_EXP_INVLN2N, x, Y2_GLIBC_EXP_SHIFT);
    const unsigned long long ki = (unsigned long long)__double_as_longlong(kd);
    kd = __dsub_rn(kd, Y2_GLIBC_EXP_SHIFT);
    const double r = __fma_rn(kd, Y2_GLIBC_EXP_NEGLN2LON, __fma_rn(kd, Y2_GLIBC_EXP_NEGLN2HIN, x));
    const unsigned idx = 2u * (unsigned)(ki & 127ull);
    const double tail = __longlong_as_double((long long)g_glibc_exp_tab[idx]);
    unsigned long long sbits = g_glibc_exp_tab[idx + 1] + (ki << 45);
    const double r2 = __dmul_rn(r, r);
    const double tmp = __fma_rn(__dmul_rn(r2, r2), __fma_rn(r, Y2_GLIBC_EXP_C5, Y2_GLIBC_EXP_C4),
                                __fma_rn(__fma_rn(r, Y2_GLIBC_EXP_C3, Y2_GLIBC_EXP_C2), r2, __dadd_rn(tail, r)));
    if (abstop == 0) {                                     // 512 <= |x| < 1024: the scale would over- / underflow (specialcase())
        if ((ki & 0x80000000ull) == 0) {
            sbits -= 1009ull << 52;
            const double scale = __longlong_as_double((long long)sbits);
            return __dmul_rn(0x1p1009, __fma_rn(scale, tmp, scale));
        }
        sbits += 1022ull << 52;
        const double scale = __longlong_as_double((long long)sbits), st = __dmul_rn(scale, tmp);
        double y = __dadd_rn(scale, st);
        if (y < 1.0) {
            double lo = __dadd_rn(__dsub_rn(scale, y), st);
            const double hi = __dadd_rn(1.0, y);
            lo = __dadd_rn(__dadd_rn(__dsub_rn(1.0, hi), y), lo);
            y = __dsub_rn(__dadd_rn(hi, lo), 1.0);
            if (y == 0.0) y = 0.0;
        }
        return __dmul_rn(0x1p-1022, y);
    }
    const double scale = __longlong_as_double((long long)sbits);
    return __fma_rn(scale, tmp, scale);
}

__device__ __forceinline__ float logistic_ref(float x) { return (float)__ddiv_rn(1., __dadd_rn(1., glibc_exp((double)(-x)))); }  // yolo_math.cpp:19

// One thread per (frame, anchor, cell): strip + dequantise + logistic(x,y,obj) + class softmax.
template <typename T, int LAYOUT>
__global__ void region_kernel(const T *__restrict__ in, float *__restrict__ out, int B, int w, int h, int n, int classes,
                              int coords, int softmax, int background, float scale, long long ifs)
{
    const int wh = w * h, per = coords + 1 + classes, wa = align8(w);
    const long long total = (long long)B * n * wh;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    int loc = idx % wh;
    long long r = idx / wh;
    int a = r % n;
    int f = r / n;
    int y = loc / w, x = loc - y * w;

    auto fetch = [&](int e) -> float {
        int ch = a * per + e;
        T v;
        if (LAYOUT == 0) v = in[f * ifs + ((long long)ch * h + y) * wa + x];
        else v = in[f * ifs + (((long long)(ch >> 2) * h + y) * w + x) * 4 + (ch & 3)];
        if (sizeof(T) == 2) return (float)v * scale;  // yolo2_model.cpp:416-420
        return (float)v;
    };
    float *o = out + ((long long)f * n + a) * per * wh + loc;
    for (int e = 0; e < coords; ++e) {
        float v = fetch(e);
        o[(long long)e * wh] = (e < 2) ? logistic_ref(v) : v;  // yolo_region.cpp:129-130
    }
    {
        float v = fetch(coords);
        o[(long long)coords * wh] = background ? v : logistic_ref(v);  // :131-132
    }
    const int first = coords + !background, nc = classes + background;
    if (softmax) {  // yolo_math.cpp:226-241 with temp = 1, stride = w*h
        float largest = -FLT_MAX;
        for (int i = 0; i < nc; ++i) {
            float v = fetch(first + i);
            if (v > largest) largest = v;
        }
        float sum = 0;
        for (int i = 0; i < nc; ++i) {
            float arg = __fsub_rn(__fdiv_rn(fetch(first + i), 1.0f), __fdiv_rn(largest, 1.0f));
            float e = (float)glibc_exp((double)arg);
            sum = __fadd_rn(sum, e);
            o[(long long)(first + i) * wh] = e;
        }
        for (int i = 0; i < nc; ++i) o[(long long)(first + i) * wh] = __fdiv_rn(o[(long long)(first + i) * wh], sum);
    } else {
        for (int i = 0; i < nc; ++i) o[(long long)(first + i) * wh] = fetch(first + i);  // raw copy (yolo_region.cpp:125)
    }
}

inline unsigned blocks_for(long long total, int threads) { return (unsigned)((total + threads - 1) / threads); }

}  // namespace

void launch_planar_to_c4(const void *src, void *dst, int B, int C, int H, int W, long long sfs, long long dfs,
                         int elem_bytes, cudaStream_t st)
{
    long long total = (long long)B * ceil_div(C, 4) * H * W;
    if (elem_bytes == 2)
        planar_to_c4_kernel<int16_t><<<blocks_for(total, 256), 256, 0, st>>>((const int16_t *)src, (int16_t *)dst, B, C, H, W, sfs, dfs);
    else
        planar_to_c4_kernel<float><<<blocks_for(total, 256), 256, 0, st>>>((const float *)src, (float *)dst, B, C, H, W, sfs, dfs);
}

void launch_c4_to_planar(const void *src, void *dst, int B, int C, int H, int W, long long sfs, long long dfs,
                         int elem_bytes, cudaStream_t st)
{
    long long total = (long long)B * ceil_div(C, 4) * H * W;
    if (elem_bytes == 2)
        c4_to_planar_kernel<int16_t><<<blocks_for(total, 256), 256, 0, st>>>((const int16_t *)src, (int16_t *)dst, B, C, H, W, sfs, dfs);
    else
        c4_to_planar_kernel<float><<<blocks_for(total, 256), 256, 0, st>>>((const float *)src, (float *)dst, B, C, H, W, sfs, dfs);
}

void launch_maxpool_planar(const void *in, void *out, int ch, int ksize, int kstride, int iw, int ih, int ow, int oh,
                           int elem_bytes, cudaStream_t st)
{
    long long total = (long long)ch * oh * ow;
    if (elem_bytes == 2)
        maxpool_planar_kernel<int16_t><<<blocks_for(total, 256), 256, 0, st>>>((const int16_t *)in, (int16_t *)out, ch, ksize, kstride, iw, ih, ow, oh);
    else
        maxpool_planar_kernel<float><<<blocks_for(total, 256), 256, 0, st>>>((const float *)in, (float *)out, ch, ksize, kstride, iw, ih, ow, oh);
}

void launch_maxpool_c4(const void *in, void *out, int B, int G, int kstride, int iw, int ih, int ow, int oh,
                       long long ifs, long long ofs, int elem_bytes, cudaStream_t st)
{
    long long total = (long long)B * G * oh * ow;
    if (elem_bytes == 2 && kstride == 2 && iw % 2 == 0 && ih % 2 == 0 && ow == iw / 2 && oh == ih / 2 && ifs % 8 == 0 &&
        ((uintptr_t)in & 15) == 0)
        maxpool_c4_i16_s2_kernel<<<blocks_for(total, 256), 256, 0, st>>>((const int16_t *)in, (int16_t *)out, total, G, iw, ih, ow, oh, ifs, ofs);
    else if (elem_bytes == 2)
        maxpool_c4_kernel<int16_t><<<blocks_for(total, 256), 256, 0, st>>>((const int16_t *)in, (int16_t *)out, B, G, kstride, iw, ih, ow, oh, ifs, ofs);
    else
        maxpool_c4_kernel<float><<<blocks_for(total, 256), 256, 0, st>>>((const float *)in, (float *)out, B, G, kstride, iw, ih, ow, oh, ifs, ofs);
}

void launch_reorg_hls_planar(const void *in, void *out, int ch, int TM, int iw, int ih, int ow, int oh, int elem_bytes,
                             cudaStream_t st)
{
    long long total = (long long)ch * oh * ow;
    if (elem_bytes == 2)
        reorg_hls_planar_kernel<int16_t><<<blocks_for(total, 256), 256, 0, st>>>((const int16_t *)in, (int16_t *)out, ch, TM, iw, ih, ow, oh);
    else
        reorg_hls_planar_kernel<float><<<blocks_for(total, 256), 256, 0, st>>>((const float *)in, (float *)out, ch, TM, iw, ih, ow, oh);
}

void launch_quantize(const float *in, int16_t *out, size_t count, int q_in, cudaStream_t st)
{
    quantize_kernel<<<blocks_for((long long)count, 256), 256, 0, st>>>(in, out, count, ldexpf(1.0f, q_in));
}

void launch_frames_to_c4(const float *frames, void *dst, int B, int C, int H, int W, long long dfs, int q_in,
                         int elem_bytes, cudaStream_t st)
{
    long long total = (long long)B * ceil_div(C, 4) * H * W;
    if (elem_bytes == 2 && C <= 4 && (W & 3) == 0 && ((uintptr_t)frames & 15) == 0 && ((uintptr_t)dst & 15) == 0 && (dfs & 7) == 0)
        frames_to_c4_i16_v4_kernel<<<blocks_for(total / 4, 256), 256, 0, st>>>(frames, (int16_t *)dst, B, C, H, W, dfs, ldexpf(1.0f, q_in));
    else if (elem_bytes == 2)
        frames_to_c4_kernel<int16_t><<<blocks_for(total, 256), 256, 0, st>>>(frames, (int16_t *)dst, B, C, H, W, dfs, ldexpf(1.0f, q_in));
    else
        frames_to_c4_kernel<float><<<blocks_for(total, 256), 256, 0, st>>>(frames, (float *)dst, B, C, H, W, dfs, 1.0f);
}

void launch_reorg_driver_planar(const void *in, void *out, int c, int h, int w, int shift, int elem_bytes, cudaStream_t st)
{
    long long total = (long long)4 * c * (h / 2) * align8(w / 2);
    if (elem_bytes == 2)
        reorg_driver_planar_kernel<int16_t><<<blocks_for(total, 256), 256, 0, st>>>((const int16_t *)in, (int16_t *)out, c, h, w, shift);
    else
        reorg_driver_planar_kernel<float><<<blocks_for(total, 256), 256, 0, st>>>((const float *)in, (float *)out, c, h, w, shift);
}

void launch_reorg_driver_c4(const void *in, void *out, int B, int c, int h, int w, int shift, long long ifs,
                            long long ofs, int elem_bytes, cudaStream_t st)
{
    long long total = (long long)B * c * (h / 2) * (w / 2);
    if (elem_bytes == 2)
        reorg_driver_c4_kernel<int16_t><<<blocks_for(total, 256), 256, 0, st>>>((const int16_t *)in, (int16_t *)out, B, c, h, w, shift, ifs, ofs);
    else
        reorg_driver_c4_kernel<float><<<blocks_for(total, 256), 256, 0, st>>>((const float *)in, (float *)out, B, c, h, w, shift, ifs, ofs);
}

void launch_region(const void *in, float *out, int B, int w, int h, int n, int classes, int coords, int softmax,
                   int background, int q, int layout, long long ifs, int elem_bytes, cudaStream_t st)
{
    long long total = (long long)B * n * w * h;
    const float scale = ldexpf(1.0f, -q);
    unsigned nb = blocks_for(total, 128);
    if (elem_bytes == 2) {
        if (layout == 0) region_kernel<int16_t, 0><<<nb, 128, 0, st>>>((const int16_t *)in, out, B, w, h, n, classes, coords, softmax, background, scale, ifs);
        else region_kernel<int16_t, 1><<<nb, 128, 0, st>>>((const int16_t *)in, out, B, w, h, n, classes, coords, softmax, background, scale, ifs);
    } else {
        if (layout == 0) region_kernel<float, 0><<<nb, 128, 0, st>>>((const float *)in, out, B, w, h, n, classes, coords, softmax, background, 1.0f, ifs);
        else region_kernel<float, 1><<<nb, 128, 0, st>>>((const float *)in, out, B, w, h, n, classes, coords, softmax, background, 1.0f, ifs);
    }
}

// ---- image front-end: stb u8 HWC image -> float CHW in [0,1] -> letterbox (bilinear resize + 0.5 fill) -------------------
// Bit-exact restatement of load_image_stb's conversion (src/core/yolo_image.cpp:178-187), resize_image (:84-127, the darknet
// two-pass bilinear: first along x into `part`, then along y) and letterbox_image / embed_image (:129-165).  Every float
// operation of the reference is one rounded operation here (__fmul_rn / __fadd_rn / __fsub_rn: no FMA contraction; the
// reference is built without -march, so g++ emits separate mulss/addss).  One thread = one pixel of the network input.
struct LetterboxParams {
    int iw, ih, ic, net_w, net_h, new_w, new_h, off_x, off_y;
    float w_scale, h_scale;
};

__device__ __forceinline__ float lb_src(const unsigned char *img, const LetterboxParams &p, int x, int y, int k)
{
    return __double2float_rn((double)img[k + p.ic * x + p.ic * p.iw * y] / 255.);   // (float)data[src_index]/255.
}
// one pixel of resize_image's first pass: part(cx, r, k)
__device__ __forceinline__ float lb_part(const unsigned char *img, const LetterboxParams &p, int cx, int r, int k)
{
    if (cx == p.new_w - 1 || p.iw == 1) return lb_src(img, p, p.iw - 1, r, k);
    const float sx = __fmul_rn((float)cx, p.w_scale);
    const int ix = (int)sx;
    const float dx = __fsub_rn(sx, (float)ix);
    return __fadd_rn(__fmul_rn(__fsub_rn(1.0f, dx), lb_src(img, p, ix, r, k)), __fmul_rn(dx, lb_src(img, p, ix + 1, r, k)));
}

__global__ void letterbox_kernel(const unsigned char *__restrict__ src, float *__restrict__ dst, int B, const LetterboxParams p)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long per = (long long)p.net_w * p.net_h;
    if (idx >= per * B) return;
    const int f = (int)(idx / per);
    const int rem = (int)(idx - (long long)f * per);
    const int y = rem / p.net_w, x = rem - y * p.net_w;
    const unsigned char *img = src + (size_t)f * p.iw * p.ih * p.ic;
    float *out = dst + (size_t)f * p.ic * per;
    const int cx = x - p.off_x, r = y - p.off_y;
    const bool inside = cx >= 0 && cx < p.new_w && r >= 0 && r < p.new_h;
    float sy = 0.f, dy = 0.f;
    int iy = 0;
    if (inside) {
        sy = __fmul_rn((float)r, p.h_scale);
        iy = (int)sy;
        dy = __fsub_rn(sy, (float)iy);
    }
    for (int k = 0; k < p.ic; ++k) {
        float v = 0.5f;                                                       // fill_image(boxed, .5)
        if (inside) {
            v = __fmul_rn(__fsub_rn(1.0f, dy), lb_part(img, p, cx, iy, k));   // set_pixel(resized, ..., (1-dy) * part(iy))
            if (!(r == p.new_h - 1 || p.ih == 1)) v = __fadd_rn(v, __fmul_rn(dy, lb_part(img, p, cx, iy + 1, k)));   // add_pixel
        }
        out[(size_t)k * per + rem] = v;
    }
}

// diagnostics: the restated exp on an array (the tests compare it with the host libm bit for bit)
__global__ void glibc_exp_kernel(const double *__restrict__ x, double *__restrict__ y, long long n)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = glibc_exp(x[i]);
}
void launch_glibc_exp(const double *x, double *y, long long n, cudaStream_t st) { glibc_exp_kernel<<<blocks_for(n, 256), 256, 0, st>>>(x, y, n); }

void launch_letterbox(const unsigned char *src, float *dst, int B, int iw, int ih, int ic, int net_w, int net_h, cudaStream_t st)
{
    LetterboxParams p;
    p.iw = iw; p.ih = ih; p.ic = ic; p.net_w = net_w; p.net_h = net_h;
    if (((float)net_w / iw) < ((float)net_h / ih)) { p.new_w = net_w; p.new_h = (ih * net_w) / iw; }   // letterbox_image :150-156
    else { p.new_h = net_h; p.new_w = (iw * net_h) / ih; }
    p.w_scale = (float)(iw - 1) / (p.new_w - 1);     // resize_image :89-90 (host float division, same expression)
    p.h_scale = (float)(ih - 1) / (p.new_h - 1);
    p.off_x = (net_w - p.new_w) / 2;
    p.off_y = (net_h - p.new_h) / 2;
    const long long total = (long long)B * net_w * net_h;
    letterbox_kernel<<<blocks_for(total, 256), 256, 0, st>>>(src, dst, B, p);
}

// ---- detections on the GPU: get_region_detections + correct_region_boxes + do_nms_sort -------------------------------------
// (src/core/yolo_region.cpp:15-53,169-195, src/core/yolo_post.cpp:22-85).  One CTA = one (class, frame).  Every float operation of the
// reference is one rounded operation here; the box w/h use glibc's expf ALGORITHM restated in double arithmetic (table of 2^(i/32) +
// cubic, sysdeps/ieee754/flt-32/e_expf.c of glibc >= 2.27): it reproduces the host's expf bit for bit (checked against libm on 2^25
// inputs by the checker program expf_check.c of the test tree), which CUDA's own expf (2 ulp) does not.  Output is POSITIONAL (entry e = cell*n + anchor, the order in
// which the reference fills its candidate list; entries at or below the objectness threshold are all-zero) instead of the
// reference's compacted list whose final order depends on qsort; the set of surviving (box, class, probability) is identical.  Equal
// probabilities inside one class are ordered by entry index (the stable order glibc's merge-sort qsort produces).
struct DetParams {
    int lw, lh, n, classes, im_w, im_h, net_w, net_h, new_w, new_h;
    float thresh, nms;
    float anchors[32];
    double expf_tab[32];       // 2^(i/32), correctly rounded
};

__device__ __forceinline__ float glibc_expf(float x, const double *tab)
{
    const double InvLn2N = 0x1.71547652b82fep+0 * 32, SHIFT = 0x1.8p+52;
    const double C0 = 0x1.c6af84b912394p-5 / 32 / 32 / 32, C1 = 0x1.ebfce50fac4f3p-3 / 32 / 32, C2 = 0x1.62e42ff0c52d6p-1 / 32;
    const double xd = (double)x;
    const double z = __dmul_rn(InvLn2N, xd);
    double kd = __dadd_rn(z, SHIFT);
    const unsigned long long ki = (unsigned long long)__double_as_longlong(kd);
    kd = __dadd_rn(kd, -SHIFT);
    const double r = __dadd_rn(z, -kd);
    unsigned long long t = (unsigned long long)__double_as_longlong(tab[ki % 32]) - ((ki % 32) << 47);
    t += ki << 47;
    const double s = __longlong_as_double((long long)t);
    const double zz = __dadd_rn(__dmul_rn(C0, r), C1);
    const double r2 = __dmul_rn(r, r);
    double y = __dadd_rn(__dmul_rn(C2, r), 1.0);
    y = __dadd_rn(__dmul_rn(zz, r2), y);
    y = __dmul_rn(y, s);
    return __double2float_rn(y);
}

__device__ __forceinline__ float det_overlap(float x1, float w1, float x2, float w2)
{
    const float l1 = __fsub_rn(x1, __fdiv_rn(w1, 2.f)), l2 = __fsub_rn(x2, __fdiv_rn(w2, 2.f));
    const float left = l1 > l2 ? l1 : l2;
    const float r1 = __fadd_rn(x1, __fdiv_rn(w1, 2.f)), r2 = __fadd_rn(x2, __fdiv_rn(w2, 2.f));
    const float right = r1 < r2 ? r1 : r2;
    return __fsub_rn(right, left);
}
__device__ __forceinline__ float det_iou(float4 a, float4 b)   // box = (x, y, w, h)
{
    const float w = det_overlap(a.x, a.z, b.x, b.z), h = det_overlap(a.y, a.w, b.y, b.w);
    const float inter = (w < 0 || h < 0) ? 0.f : __fmul_rn(w, h);
    const float uni = __fsub_rn(__fadd_rn(__fmul_rn(a.z, a.w), __fmul_rn(b.z, b.w)), inter);
    return __fdiv_rn(inter, uni);
}

constexpr int kDetMax = 1024;   // candidates per frame the kernel can hold (13 x 13 x 5 = 845 for YOLOv2-416, 19 x 19 x 5 > 1024 -> host path)

__global__ void __launch_bounds__(256) detect_kernel(const float *__restrict__ region, float *__restrict__ boxes, float *__restrict__ probs,
                                                      float *__restrict__ objectness, const DetParams p)
{
    __shared__ float4 s_box[kDetMax];
    __shared__ float s_p[kDetMax];
    __shared__ short s_src[kDetMax];     // sorted position -> entry
    __shared__ float4 s_sbox[kDetMax];
    __shared__ float s_sp[kDetMax];
    __shared__ float s_obj[kDetMax];
    __shared__ int s_m;
    const int k = blockIdx.x, f = blockIdx.y, tid = threadIdx.x;
    const int wh = p.lw * p.lh, total = wh * p.n, per = 5 + p.classes;
    const float *reg = region + (size_t)f * p.n * per * wh;
    float *bo = boxes + (size_t)f * total * 4, *po = probs + (size_t)f * total * p.classes, *oo = objectness + (size_t)f * total;
    if (tid == 0) s_m = 0;
    // decode (get_region_detections + correct_region_boxes), entry e = cell * n + anchor
    for (int e = tid; e < total; e += blockDim.x) {
        const int cell = e / p.n, a = e - cell * p.n;
        const int row = cell / p.lw, col = cell - row * p.lw;
        const float *x = reg + (size_t)a * per * wh + cell;
        const float obj = x[(size_t)4 * wh];
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        float pr = 0.f, o = 0.f;
        if (obj > p.thresh) {
            o = obj;
            b.x = __fdiv_rn(__fadd_rn((float)col, x[0]), (float)p.lw);
            b.y = __fdiv_rn(__fadd_rn((float)row, x[(size_t)wh]), (float)p.lh);
            b.z = __fdiv_rn(__fmul_rn(glibc_expf(x[(size_t)2 * wh], p.expf_tab), p.anchors[2 * a]), (float)p.lw);
            b.w = __fdiv_rn(__fmul_rn(glibc_expf(x[(size_t)3 * wh], p.expf_tab), p.anchors[2 * a + 1]), (float)p.lh);
            // correct_region_boxes, relative = 1 (double where the reference's expression is double)
            b.x = __double2float_rn(__ddiv_rn(__dsub_rn((double)b.x, __ddiv_rn(__ddiv_rn((double)(p.net_w - p.new_w), 2.), (double)p.net_w)),
                                              (double)__fdiv_rn((float)p.new_w, (float)p.net_w)));
            b.y = __double2float_rn(__ddiv_rn(__dsub_rn((double)b.y, __ddiv_rn(__ddiv_rn((double)(p.net_h - p.new_h), 2.), (double)p.net_h)),
                                              (double)__fdiv_rn((float)p.new_h, (float)p.net_h)));
            b.z = __fmul_rn(b.z, __fdiv_rn((float)p.net_w, (float)p.new_w));
            b.w = __fmul_rn(b.w, __fdiv_rn((float)p.net_h, (float)p.new_h));
            const float q = __fmul_rn(obj, x[(size_t)(5 + k) * wh]);
            pr = q > p.thresh ? q : 0.f;
        }
        s_box[e] = b;
        s_p[e] = pr;
        s_obj[e] = o;
        if (k == 0) {
            reinterpret_cast<float4 *>(bo)[e] = b;
            oo[e] = o;
        }
    }
    __syncthreads();
    if (p.nms > 0.f) {
        // do_nms_sort for class k: order the entries with a non-zero probability by probability, descending.  Ties: the reference
        // sorts ONE list in place class after class with a stable sort (glibc's qsort is a merge sort), so equal probabilities of
        // class k keep the order the sorts of classes k-1, k-2, ... 0 left: descending probability of class k-1, then k-2, ...,
        // finally the scan order (entry index).  Those probabilities are the values before any suppression (a class's values are
        // only zeroed after its own sort), i.e. recomputable from the region tensor; the loop only runs on exact ties.
        auto prob_of = [&](int e, int c) -> float {
            const float o = s_obj[e];
            if (o == 0.f) return 0.f;
            const int cell = e / p.n, a = e - cell * p.n;
            const float q = __fmul_rn(o, reg[(size_t)a * per * wh + cell + (size_t)(5 + c) * wh]);
            return q > p.thresh ? q : 0.f;
        };
        auto tie_before = [&](int j, int e) -> bool {
            for (int c = k - 1; c >= 0; --c) {
                const float pj = prob_of(j, c), pe = prob_of(e, c);
                if (pj != pe) return pj > pe;
            }
            return j < e;
        };
        for (int e = tid; e < total; e += blockDim.x) {
            const float pe = s_p[e];
            if (pe == 0.f) continue;
            int rank = 0;
            for (int j = 0; j < total; ++j) {
                const float pj = s_p[j];
                rank += (pj > pe) || (pj == pe && j != e && tie_before(j, e));
            }
            s_src[rank] = (short)e;
            s_sbox[rank] = s_box[e];
            s_sp[rank] = pe;
            atomicAdd(&s_m, 1);
        }
        __syncthreads();
        const int m = s_m;
        // ... then every surviving entry suppresses the later ones it overlaps
        for (int i = 0; i < m; ++i) {
            if (s_sp[i] != 0.f) {
                const float4 a = s_sbox[i];
                for (int j = i + 1 + tid; j < m; j += blockDim.x)
                    if (det_iou(a, s_sbox[j]) > p.nms) s_sp[j] = 0.f;
            }
            __syncthreads();
        }
        for (int i = tid; i < m; i += blockDim.x) s_p[s_src[i]] = s_sp[i];
        __syncthreads();
    }
    for (int e = tid; e < total; e += blockDim.x) po[(size_t)e * p.classes + k] = s_p[e];
}

// Positional detect_kernel output -> compact fixed-size records per frame (what leaves the GPU / crosses NVLink instead of
// 287 KB region tensors): record = {entry (cell * n + anchor), class, probability, x, y, w, h, objectness} as eight 32-bit words,
// ordered by (entry, class); counts[f] = number of (entry, class) pairs with a surviving probability (may exceed cap: the
// first cap are stored).  One CTA per frame, ordered block-wide compaction (ballot + prefix), deterministic.
__global__ void __launch_bounds__(256) compact_detections_kernel(const float *__restrict__ boxes, const float *__restrict__ probs,
                                                                  const float *__restrict__ objectness, int total, int classes, int cap,
                                                                  unsigned *__restrict__ records, int *__restrict__ counts)
{
    __shared__ int s_warp[8];
    __shared__ int s_base;
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *po = probs + (size_t)f * total * classes, *bo = boxes + (size_t)f * total * 4, *oo = objectness + (size_t)f * total;
    unsigned *ro = records + (size_t)f * cap * 8;
    if (tid == 0) s_base = 0;
    __syncthreads();
    const int n = total * classes;
    for (int i0 = 0; i0 < n; i0 += 256) {
        const int i = i0 + tid;
        const float pr = i < n ? po[i] : 0.f;
        const bool keep = pr > 0.f;
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        int before = s_base;
        for (int w = 0; w < warp; ++w) before += s_warp[w];
        const int pos = before + __popc(m & ((1u << lane) - 1u));
        if (keep && pos < cap) {
            const int e = i / classes, c = i - e * classes;
            const float4 b = reinterpret_cast<const float4 *>(bo)[e];
            unsigned *r = ro + (size_t)pos * 8;
            r[0] = (unsigned)e; r[1] = (unsigned)c; r[2] = __float_as_uint(pr);
            r[3] = __float_as_uint(b.x); r[4] = __float_as_uint(b.y); r[5] = __float_as_uint(b.z); r[6] = __float_as_uint(b.w);
            r[7] = __float_as_uint(oo[e]);
        }
        __syncthreads();
        if (tid == 0) {
            int t = 0;
            for (int w = 0; w < 8; ++w) t += s_warp[w];
            s_base += t;
        }
        __syncthreads();
    }
    if (tid == 0) counts[f] = s_base;
}

void launch_compact_detections(const float *boxes, const float *probs, const float *objectness, int B, int total, int classes, int cap,
                               unsigned *records, int *counts, cudaStream_t st)
{
    compact_detections_kernel<<<B, 256, 0, st>>>(boxes, probs, objectness, total, classes, cap, records, counts);
}

// returns 0 when launched, -1 when the frame has more candidates than the kernel holds
int launch_detect(const float *region, float *boxes, float *probs, float *objectness, int B, int lw, int lh, int n, int classes,
                  const float *anchors, int im_w, int im_h, int net_w, int net_h, float thresh, float nms, const double *expf_tab,
                  cudaStream_t st)
{
    if (lw * lh * n > kDetMax || n > 16) return -1;
    DetParams p;
    p.lw = lw; p.lh = lh; p.n = n; p.classes = classes; p.im_w = im_w; p.im_h = im_h; p.net_w = net_w; p.net_h = net_h;
    if (((float)net_w / im_w) < ((float)net_h / im_h)) { p.new_w = net_w; p.new_h = (im_h * net_w) / im_w; }   // correct_region_boxes :31-37
    else { p.new_h = net_h; p.new_w = (im_w * net_h) / im_h; }
    p.thresh = thresh; p.nms = nms;
    for (int i = 0; i < 2 * n; ++i) p.anchors[i] = anchors[i];
    for (int i = 0; i < 32; ++i) p.expf_tab[i] = expf_tab[i];
    detect_kernel<<<dim3(classes, B), 256, 0, st>>>(region, boxes, probs, objectness, p);
    return 0;
}

}  // namespace y2
