// FP32 convolution of the YOLOv2 accelerator datapath (reference float build:
// hls/core/core_compute.cpp:121-172, leaky :200-204).  acc = bias; for each 4-channel group and
// tap: acc += sum_{t<4} w*x.  Runs on FFMA; parity tolerance 1e-4 relative (BASELINE.json).
// Same tiling as the int16 kernel (conv_i16.cu) with a 16-byte C4 pixel word and 7-pixel segments.
#include "common.cuh"

namespace y2 {

namespace {

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc)
{
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

constexpr int kWM = 4, kWS = 2, kTMC = 4;
constexpr int kThreads = 32 * kWM * kWS;

template <int TP, int KS>
__global__ void __launch_bounds__(kThreads, 2) conv_f32_c4_kernel(const ConvFastParams p)
{
    constexpr int K2 = KS * KS;
    constexpr int PAD = KS / 2;
    constexpr int XW = TP + KS - 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int xrows = p.RB + KS - 1 + 1;
    const int zero_slot = xrows - 1;
    const int x_stage_px = p.GS * xrows * p.PW;
    const int w_stage_px = p.GS * K2 * kCM;
    const int stage_px = x_stage_px + w_stage_px;
    float4 *sm = reinterpret_cast<float4 *>(smem_raw);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int wm = warp % kWM, ws = warp / kWM;
    const int seg = ws * 32 + lane;
    const int rows_total = p.B * p.H;
    const int R0 = blockIdx.x * p.RB;
    const int mb = blockIdx.y;

    const int row_local = seg / p.SW;
    const int sx = seg - row_local * p.SW;
    const int R = R0 + row_local;
    const bool active = (row_local < p.RB) && (R < rows_total);
    const int f = active ? R / p.H : 0;
    const int y = active ? R - f * p.H : 0;

    int xoff[KS];
#pragma unroll
    for (int i = 0; i < KS; ++i) {
        int yin = y + i - PAD;
        int slot = (active && yin >= 0 && yin < p.H) ? row_local + i : zero_slot;
        xoff[i] = slot * p.PW + sx * TP;
    }

    for (int i = tid; i < 2 * stage_px; i += kThreads) sm[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    int2 *tbl = reinterpret_cast<int2 *>(sm + 2 * stage_px);   // per-CTA copy table, see conv_i16.cu
    const int n_tbl = (p.RB + KS - 1) * p.W;
    for (int idx = tid; idx < n_tbl; idx += kThreads) {
        int s = idx / p.W, x = idx - s * p.W;
        int Rr = R0 - PAD + s;
        int2 e = make_int2(-1, 0);
        if (Rr >= 0 && Rr < rows_total) {
            int ff = Rr / p.H, yy = Rr - ff * p.H;
            e.x = (int)(ff * (p.in_frame_stride >> 2)) + yy * p.W + x;
            e.y = s * p.PW + PAD + x;
        }
        tbl[idx] = e;
    }
    __syncthreads();

    const float4 *in_px = static_cast<const float4 *>(p.in);
    const float4 *wsrc = static_cast<const float4 *>(p.w) + (size_t)mb * p.G * K2 * kCM;

    auto load_stage = [&](int st, int buf) {
        const int g0 = st * p.GS;
        const int ng = min(p.GS, p.G - g0);
        float4 *xs = sm + buf * stage_px;
        float4 *wsm = xs + x_stage_px;
        const int plane = p.H * p.W;
        for (int idx = tid; idx < n_tbl; idx += kThreads) {
            const int2 e = tbl[idx];
            if (e.x < 0) continue;
            const float4 *src = in_px + e.x + (size_t)g0 * plane;
            float4 *dst = xs + e.y;
            for (int gg = 0; gg < ng; ++gg) cp_async16(dst + gg * xrows * p.PW, src + (size_t)gg * plane);
        }
        const float4 *wg = wsrc + (size_t)g0 * K2 * kCM;
        for (int idx = tid; idx < ng * K2 * kCM; idx += kThreads) cp_async16(wsm + idx, wg + idx);
    };

    float acc[kTMC][TP];
    {
        const float *bias = static_cast<const float *>(p.bias);
#pragma unroll
        for (int c = 0; c < kTMC; ++c) {
            int m = mb * kCM + wm * kTMC + c;
            float b = (m < p.OFM) ? bias[m] : 0.0f;
#pragma unroll
            for (int q = 0; q < TP; ++q) acc[c][q] = b;
        }
    }

    const int nstages = (p.G + p.GS - 1) / p.GS;
    load_stage(0, 0);
    cp_async_commit();
    for (int st = 0; st < nstages; ++st) {
        if (st + 1 < nstages) {
            load_stage(st + 1, (st + 1) & 1);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const float4 *xs = sm + (st & 1) * stage_px;
        const float4 *wsm = xs + x_stage_px + wm * kTMC;
        const int ng = min(p.GS, p.G - st * p.GS);
        for (int gg = 0; gg < ng; ++gg) {
            const float4 *xg = xs + gg * xrows * p.PW;
            const float4 *wg = wsm + gg * K2 * kCM;
#pragma unroll 1
            for (int i = 0; i < KS; ++i) {
                int xo = xoff[0];
#pragma unroll
                for (int t = 1; t < KS; ++t) xo = (i == t) ? xoff[t] : xo;
                const float4 *xr = xg + xo;
                float4 xv[XW];
#pragma unroll
                for (int q = 0; q < XW; ++q) xv[q] = xr[q];
#pragma unroll
                for (int j = 0; j < KS; ++j) {
                    float4 wv[kTMC];
#pragma unroll
                    for (int c = 0; c < kTMC; ++c) wv[c] = wg[(i * KS + j) * kCM + c];
#pragma unroll
                    for (int c = 0; c < kTMC; ++c)
#pragma unroll
                        for (int q = 0; q < TP; ++q) {
                            const float4 x = xv[q + j];
                            float ps = wv[c].x * x.x;  // core_compute.cpp:159-168
                            ps = fmaf(wv[c].y, x.y, ps);
                            ps = fmaf(wv[c].z, x.z, ps);
                            ps = fmaf(wv[c].w, x.w, ps);
                            acc[c][q] += ps;
                        }
                }
            }
        }
        __syncthreads();
    }

    if (!active) return;
    if (mb * kCM + wm * kTMC >= p.OFM) return;
    float *out = static_cast<float *>(p.out) + (size_t)f * p.out_frame_stride +
                 (((size_t)(mb * (kCM / 4) + wm) * p.H + y) * p.W + sx * TP) * 4;
#pragma unroll
    for (int q = 0; q < TP; ++q) {
        if (sx * TP + q >= p.W) break;
        float v[kTMC];
#pragma unroll
        for (int c = 0; c < kTMC; ++c) {
            float a = acc[c][q];
            if (a < 0.0f && p.leaky) a = a * 0.1f;  // core_compute.cpp:200-204
            v[c] = a;
        }
        *reinterpret_cast<float4 *>(out + q * 4) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

__global__ void wprep_f32_kernel(const float *__restrict__ blob, float4 *__restrict__ dst, int ifm, int ofm, int ksize,
                                 int TM, int TN, int G, int total)
{
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int k2 = ksize * ksize;
    int ml = idx % kCM;
    int r = idx / kCM;
    int tap = r % k2;
    r /= k2;
    int g = r % G;
    int mb = r / G;
    int m = mb * kCM + ml;
    float w[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        int c = g * 4 + t;
        w[t] = (m < ofm && c < ifm) ? blob[reorg_woff(m, c, tap, ifm, ofm, k2, TM, TN)] : 0.0f;
    }
    dst[idx] = make_float4(w[0], w[1], w[2], w[3]);
}

__global__ void conv_f32_generic_kernel(const float *__restrict__ in, float *__restrict__ out, const float *__restrict__ w,
                                        const float *__restrict__ bias, int ifm, int ofm, int ksize, int kstride, int iw,
                                        int ih, int ow, int oh, int pad, int is_nl, int TM, int TN)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int m = blockIdx.z;
    if (x >= ow) return;
    const int iwa = align8(iw), owa = align8(ow), k2 = ksize * ksize;
    float acc = bias[m];
    const int groups = (ifm + TN - 1) / TN;
    for (int g = 0; g < groups; ++g) {
        const int n0 = g * TN, tnn = min(TN, ifm - n0);
        for (int i = 0; i < ksize; ++i)
            for (int j = 0; j < ksize; ++j) {
                const int iy = y * kstride + i - pad, ix = x * kstride + j - pad;
                if (iy < 0 || iy >= ih || ix < 0 || ix >= iw) continue;
                float ps = 0.0f;
                for (int t = 0; t < tnn; ++t)
                    ps = __fadd_rn(ps, __fmul_rn(w[reorg_woff(m, n0 + t, i * ksize + j, ifm, ofm, k2, TM, TN)],
                                                 in[((size_t)(n0 + t) * ih + iy) * iwa + ix]));
                acc = __fadd_rn(acc, ps);
            }
    }
    if (acc < 0.0f && is_nl) acc = acc * 0.1f;
    out[((size_t)m * oh + y) * owa + x] = acc;
}

template <int TP, int KS>
int launch_f32_variant(const ConvFastParams &p, size_t smem, cudaStream_t st)
{
    // set on every launch (a microsecond): the attribute is per device and a host may drive several GPUs from several threads,
    // so a cached "already configured" flag would be a data race for nothing
    cudaFuncSetAttribute(conv_f32_c4_kernel<TP, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    dim3 grid(ceil_div(p.B * p.H, p.RB), ceil_div(p.OFM, kCM));
    conv_f32_c4_kernel<TP, KS><<<grid, kThreads, smem, st>>>(p);
    return 1;
}

}  // namespace

int launch_conv_f32_fast(const ConvFastParams &p, int ksize, cudaStream_t st, const char **variant)
{
    const int xrows = p.RB + ksize - 1 + 1;
    const size_t smem = 2 * (size_t)p.GS * ((size_t)xrows * p.PW + (size_t)ksize * ksize * kCM) * 16 +
                        (size_t)(p.RB + ksize - 1) * p.W * 8;  // + the copy table
    if (p.TP != 7) return -1;
    if (ksize == 3) { if (variant) *variant = "conv_f32_c4<7,3>"; return launch_f32_variant<7, 3>(p, smem, st); }
    if (ksize == 1) { if (variant) *variant = "conv_f32_c4<7,1>"; return launch_f32_variant<7, 1>(p, smem, st); }
    return -1;
}

void launch_wprep_f32(const float *blob, void *dst, int ifm, int ofm, int ksize, int TM, int TN, cudaStream_t st)
{
    const int G = ceil_div(ifm, 4);
    const int total = ceil_div(ofm, kCM) * G * ksize * ksize * kCM;
    wprep_f32_kernel<<<ceil_div(total, 256), 256, 0, st>>>(blob, static_cast<float4 *>(dst), ifm, ofm, ksize, TM, TN, G, total);
}

void launch_conv_f32_generic(const float *in, float *out, const float *w, const float *bias, int ifm, int ofm,
                             int ksize, int kstride, int iw, int ih, int ow, int oh, int pad, int is_nl, int TM, int TN,
                             cudaStream_t st)
{
    dim3 grid(ceil_div(ow, 128), oh, ofm);
    conv_f32_generic_kernel<<<grid, 128, 0, st>>>(in, out, w, bias, ifm, ofm, ksize, kstride, iw, ih, ow, oh, pad, is_nl, TM, TN);
}

}  // namespace y2
