// INT16 3x3 convolution of a layer with ONE input channel group (IFM <= 4: YOLOv2's first layer, 3 -> 32 channels at 416x416),
// optionally with the 2x2 / stride-2 max-pool that follows it FUSED into the store (north_star (b): "fused into the conv store where
// possible").  Same arithmetic as conv_i16_c4_kernel (csrc/conv_i16.cu: the 7-instruction exact step of
// hls/core/core_compute.cpp:65-120, bias init :49-62, leaky :193-198); what differs is everything around the nine steps such a layer
// has per output: no shared-memory pipeline (the input is 8 bytes per pixel: the window comes straight from global memory through L1),
// no per-CTA copy table, two output rows per thread so that the 2x2 pooling window lives in one thread's registers, and - fused - a
// quarter of the store and no pool launch.  max-pool commutes with the leaky activation (x -> x/10 truncating is monotonic), so the
// fused form pools the saturated accumulators and activates once (core_compute.cpp:193-198 then pool_yolo2 :231-310).
#include "common.cuh"

namespace y2 {

namespace {

__device__ __forceinline__ int dp2a_lo_su(int a, unsigned b, int c) { int d; asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ int dp2a_hi_su(int a, unsigned b, int c) { int d; asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ int dp2a_lo_ss(int a, int b, int c) { int d; asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ int dp2a_hi_ss(int a, int b, int c) { int d; asm("dp2a.hi.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }

constexpr int kTX = 8;                  // output columns per thread
constexpr int kTY = 2;                  // output rows per thread (= the pooling window's height)
constexpr int kWarps = 8;               // warp w = output channel quad w of the CTA's 32 channels; lanes = 32 row-pair segments
constexpr int kThreads = 32 * kWarps;
constexpr int kK2 = 9;

struct G1Params {
    const uint2 *in;                    // C4 input, one group: [frame][H][W] 8-byte pixels
    int16_t *out;                       // C4 output (conv resolution, or pooled resolution when POOL)
    const uint2 *w;                     // [ceil(OFM/16)][1][9][16] {lo bytes x4, hi bytes x4} (wprep_i16 layout)
    const int16_t *bias;
    int *xmax_out;
    int B, H, W, OFM;
    long long in_frame_stride, out_frame_stride;   // elements
    int so, sb, leaky;
    int segs_per_row;                   // ceil(W / kTX)
    long long nseg;                     // B * ceil(H / 2) * segs_per_row
};

template <bool SCALED, bool POOL>
__global__ void __launch_bounds__(kThreads, 2) conv_i16_g1_kernel(const G1Params p)
{
    __shared__ uint2 sWt[kWarps * 4 * kK2];          // [channel of the CTA's 32][tap]
    __shared__ int s_amax;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.y * 32 + warp * 4;       // first of this thread's four output channels
    for (int i = tid; i < kWarps * 4 * kK2; i += kThreads) {
        const int c = i / kK2, tap = i - c * kK2, m = blockIdx.y * 32 + c;
        sWt[i] = (m < p.OFM) ? p.w[((size_t)(m >> 4) * kK2 + tap) * kCM + (m & 15)] : make_uint2(0u, 0u);
    }
    if (tid == 0) s_amax = 0;
    __syncthreads();

    const long long seg = (long long)blockIdx.x * 32 + lane;
    const bool active = seg < p.nseg && m0 < p.OFM;
    const int rows2 = (p.H + 1) >> 1;
    const int sx = (int)(seg % p.segs_per_row);
    const long long rp = seg / p.segs_per_row;
    const int f = active ? (int)(rp / rows2) : 0;
    const int y0 = active ? (int)(rp - (long long)f * rows2) * 2 : 0;
    const int x0 = sx * kTX;

    const int half = 1 << (p.so - 1);
    const int k2 = p.so - 8;
    const int nmask = ~((1 << k2) - 1);
    const int ubound = 65535 << (SCALED ? k2 : 0);
    int acc[kTY][4][kTX];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int m = m0 + c;
        const long long b = (m < p.OFM) ? (long long)p.bias[m] : 0;
        const long long base = round_shift64(b, p.sb);
        const long long rb = (SCALED ? ((1LL << 25) >> k2) : (1LL << 26)) + 2;     // see conv_i16_c4_kernel
        long long boff = base + 32768;
        if (boff > 65535 + rb) boff = 65535 + rb;
        if (boff < -rb) boff = -rb;
        const int init = SCALED ? (int)(boff * (1LL << k2)) : (int)boff;
#pragma unroll
        for (int r = 0; r < kTY; ++r)
#pragma unroll
            for (int q = 0; q < kTX; ++q) acc[r][c][q] = init;
    }

    if (active) {
        const uint2 *inf = p.in + (size_t)f * (p.in_frame_stride >> 2);
        // input rows y0-1 .. y0+2: row A (y0) takes tap row ia = s, row B (y0+1) tap row ib = s-1; each chain sees its taps in order
#pragma unroll 1
        for (int s = 0; s < kTY + 2; ++s) {
            const int yin = y0 - 1 + s;
            uint2 xv[kTX + 2];
            const bool rowok = yin >= 0 && yin < p.H;
#pragma unroll
            for (int q = 0; q < kTX + 2; ++q) {
                const int xin = x0 - 1 + q;
                xv[q] = (rowok && xin >= 0 && xin < p.W) ? inf[(size_t)yin * p.W + xin] : make_uint2(0u, 0u);
            }
#pragma unroll
            for (int r = 0; r < kTY; ++r) {
                const int i = s - r;                      // tap row of output row r for this input row
                if (i < 0 || i > 2) continue;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    uint2 wv[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) wv[c] = sWt[(warp * 4 + c) * kK2 + i * 3 + j];
#pragma unroll
                    for (int c = 0; c < 4; ++c)
#pragma unroll
                        for (int q = 0; q < kTX; ++q) {
                            const uint2 x = xv[q + j];
                            int plo = dp2a_lo_su((int)x.x, wv[c].x, half);
                            plo = dp2a_hi_su((int)x.y, wv[c].x, plo);
                            if (SCALED) {
                                int phi = dp2a_lo_ss((int)x.x, (int)wv[c].y, acc[r][c][q] + (plo >> 8));
                                phi = dp2a_hi_ss((int)x.y, (int)wv[c].y, phi);
                                acc[r][c][q] = __vimin_s32_relu(phi & nmask, ubound);
                            } else {
                                int phi = dp2a_lo_ss((int)x.x, (int)wv[c].y, plo >> 8);
                                phi = dp2a_hi_ss((int)x.y, (int)wv[c].y, phi);
                                acc[r][c][q] = __viaddmin_s32_relu(acc[r][c][q], phi >> k2, 65535);
                            }
                        }
                }
            }
        }
    }

    int amax = 0;
    if (active) {
        auto finish = [&](int u) -> int {               // saturated accumulator (+32768, maybe scaled) -> activated output
            int a = (SCALED ? (u >> k2) : u) - 32768;
            if (p.leaky && a < 0) a = a / 10;           // C division, truncates toward zero
            amax = max(amax, a < 0 ? -a : a);
            return a & 0xffff;
        };
        if (POOL) {
            // (H and W are even here: the launcher checks) pooled pixel (y0/2, x0/2 + q): max over the 2x2 window, then the activation
            const int oh = p.H >> 1, ow = p.W >> 1;
            int16_t *out = p.out + (size_t)f * p.out_frame_stride + (((size_t)(m0 >> 2) * oh + (y0 >> 1)) * ow + (x0 >> 1)) * 4;
#pragma unroll
            for (int q = 0; q < kTX / 2; ++q) {
                if (x0 + 2 * q >= p.W) break;
                int v[4];
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    v[c] = finish(max(max(acc[0][c][2 * q], acc[0][c][2 * q + 1]), max(acc[1][c][2 * q], acc[1][c][2 * q + 1])));
                *reinterpret_cast<uint2 *>(out + q * 4) = make_uint2((unsigned)v[0] | ((unsigned)v[1] << 16), (unsigned)v[2] | ((unsigned)v[3] << 16));
            }
        } else {
#pragma unroll
            for (int r = 0; r < kTY; ++r) {
                if (y0 + r >= p.H) break;
                int16_t *out = p.out + (size_t)f * p.out_frame_stride + (((size_t)(m0 >> 2) * p.H + y0 + r) * p.W + x0) * 4;
#pragma unroll
                for (int q = 0; q < kTX; ++q) {
                    if (x0 + q >= p.W) break;
                    int v[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) v[c] = finish(acc[r][c][q]);
                    *reinterpret_cast<uint2 *>(out + q * 4) = make_uint2((unsigned)v[0] | ((unsigned)v[1] << 16), (unsigned)v[2] | ((unsigned)v[3] << 16));
                }
            }
        }
    }
    if (p.xmax_out) {
        amax = __reduce_max_sync(0xffffffffu, amax);
        if (lane == 0 && amax > 0) atomicMax(&s_amax, amax);
        __syncthreads();
        if (tid == 0 && s_amax > 0) atomicMax(p.xmax_out, s_amax);
    }
}

}  // namespace

// Returns 1 when launched, -1 when the shape is not this kernel's (G != 1, shift outside [8, 30], odd sizes with pool).
// cp.out / cp.out_frame_stride describe the tensor that is written: the conv output, or with pool != 0 the POOLED tensor.
int launch_conv_i16_g1(const ConvFastParams &cp, int ksize, int pool, cudaStream_t st, const char **variant)
{
    if (ksize != 3 || cp.G != 1 || cp.so < 8 || cp.so > 30) return -1;
    if (pool && ((cp.H | cp.W) & 1)) return -1;
    G1Params p{};
    p.in = (const uint2 *)cp.in; p.out = (int16_t *)cp.out; p.w = (const uint2 *)cp.w; p.bias = (const int16_t *)cp.bias;
    p.xmax_out = cp.xmax_out;
    p.B = cp.B; p.H = cp.H; p.W = cp.W; p.OFM = cp.OFM;
    p.in_frame_stride = cp.in_frame_stride; p.out_frame_stride = cp.out_frame_stride;
    p.so = cp.so; p.sb = cp.sb; p.leaky = cp.leaky;
    p.segs_per_row = ceil_div(cp.W, kTX);
    p.nseg = (long long)cp.B * ((cp.H + 1) / 2) * p.segs_per_row;
    dim3 grid((unsigned)((p.nseg + 31) / 32), ceil_div(cp.OFM, 32));
    const bool scaled = cp.so <= 22;
    if (pool) {
        if (scaled) conv_i16_g1_kernel<true, true><<<grid, kThreads, 0, st>>>(p); else conv_i16_g1_kernel<false, true><<<grid, kThreads, 0, st>>>(p);
        if (variant) *variant = "conv_i16_g1<3,pool2x2>";
    } else {
        if (scaled) conv_i16_g1_kernel<true, false><<<grid, kThreads, 0, st>>>(p); else conv_i16_g1_kernel<false, false><<<grid, kThreads, 0, st>>>(p);
        if (variant) *variant = "conv_i16_g1<3>";
    }
    return 1;
}

}  // namespace y2
