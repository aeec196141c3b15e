// INT16 convolution of the YOLOv2 accelerator datapath, bit-exact to the reference's
// compute() (hls/core/core_compute.cpp:32-120) + nonlinear_leaky_row (:175-210):
//
//   acc = rs(bias, Qb-Qa_out)                                   (not saturated)
//   for each 4-channel group g, tap (i,j), in that order:        core_scheduler.cpp:45, core_compute.cpp:65-67
//       P   = sum_{t<4} w[m][4g+t][i][j] * x[4g+t][..]           int16 x int16, exact
//       acc = clamp16(acc + rs(P, Qa_in+Qw-Qa_out))              :108-118
//   out = (leaky && acc<0) ? acc/10 : acc                        :193-198
//
// Because the accumulator is rounded and saturated after EVERY 4-MAC step, the work per step is
// fixed-function integer ALU work; a big-K tensor-core GEMM is not bit-exact.
// The fast kernel spends exactly 7 SASS instructions per step and output.  IDP.2A and VIADDMNMX
// issue at half rate on B200 (profiles/microbench), LEA/LOP3/VIMNMX/SHF at full rate, so the
// preferred "scaled" form keeps only the four IDP.2A on the half-rate pipe:
//   4x IDP.2A       int16 activations x weight bytes: P = 256*sum(x*w_hi) + sum(x*w_lo), no overflow
//   1x LEA.HI.SX32  U + ((Plo+half)>>8) as the hi chain's addend; U = (acc+32768) << (so-8) is the state
//   1x LOP3         drop the fraction bits (the per-step floor)
//   1x VIMNMX.RELU  clamp to [0, 65535<<(so-8)] = the 16-bit saturation
// (valid for 8 <= so <= 22).  For 23 <= so <= 30 the unscaled form is used:
//   4x IDP.2A, 2x SHF, 1x VIADDMNMX.RELU (accumulator kept as acc+32768 in [0,65535]).
#include "common.cuh"

namespace y2 {

namespace {

__device__ __forceinline__ int dp2a_lo_su(int a, unsigned b, int c)
{
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_hi_su(int a, unsigned b, int c)
{
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_lo_ss(int a, int b, int c)
{
    int d;
    asm("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_hi_ss(int a, int b, int c)
{
    int d;
    asm("dp2a.hi.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc)
{
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc));
}
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

constexpr int kWM = 4;       // warps along output channels
constexpr int kWS = 2;       // warps along segments
constexpr int kTMC = 4;      // output channels per thread (= one C4 output group)
constexpr int kNS = 32 * kWS;
constexpr int kThreads = 32 * kWM * kWS;

// One CTA: a band of RB image rows (flattened over frames) x 16 output channels.
// One thread: one row segment of TP pixels x 4 output channels = 4*TP saturating accumulators.
// NW = C4 words per ROUNDING GROUP: 1 for the reference's default Tn = 4; 2 / 4 emulate a reference built with
// scripts/hw_params_gen.py --tn 8 / 16 (hls/core/params.hpp), whose round-and-saturate step covers 8 / 16 input channels: the
// partial sums of the group's words are added up exactly (|sum x*w_lo| < 2^28, |sum x*w_hi| < 2^27) before the one step.
template <int TP, int KS, bool SCALED, int NW = 1>
__global__ void __launch_bounds__(kThreads, 2) conv_i16_c4_kernel(const ConvFastParams p)
{
    static_assert(NW == 1 || (SCALED && TP <= 7), "wider rounding groups: scaled form, 7-pixel segments (register budget)");
    constexpr int K2 = KS * KS;
    constexpr int PAD = KS / 2;
    constexpr int XW = TP + KS - 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int xrows = p.RB + KS - 1 + 1;  // + the shared all-zero row
    const int zero_slot = xrows - 1;
    const int x_stage_px = p.GS * xrows * p.PW;  // uint2 elements
    const int w_stage_px = p.GS * K2 * kCM;      // uint2 elements
    const int stage_px = (x_stage_px + w_stage_px + 1) & ~1;  // weights first: keeps their 16-byte cp.async aligned
    uint2 *sm = reinterpret_cast<uint2 *>(smem_raw);

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

    int xoff[KS];  // smem pixel offset of this thread's first input pixel for tap row i
#pragma unroll
    for (int i = 0; i < KS; ++i) {
        int yin = y + i - PAD;
        int slot = (active && yin >= 0 && yin < p.H) ? row_local + i : zero_slot;
        xoff[i] = slot * p.PW + sx * TP;
    }

    // zero both stage buffers once: halo columns, out-of-image rows and the zero row stay zero
    for (int i = tid; i < 2 * stage_px; i += kThreads) sm[i] = make_uint2(0u, 0u);
    // per-CTA copy table: (source pixel offset inside a group plane, smem pixel offset) of every band
    // pixel, computed once so that the per-stage loader is a table walk with no index arithmetic
    int2 *tbl = reinterpret_cast<int2 *>(sm + 2 * stage_px);
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

    const uint2 *in_px = static_cast<const uint2 *>(p.in);
    const uint2 *wsrc = static_cast<const uint2 *>(p.w) + (size_t)mb * p.G * K2 * kCM;

    auto load_stage = [&](int st, int buf) {
        const int g0 = st * p.GS;
        const int ng = min(p.GS, p.G - g0);
        uint2 *wsm = sm + buf * stage_px;
        uint2 *xs = wsm + w_stage_px;
        const int plane = p.H * p.W;
        for (int idx = tid; idx < n_tbl; idx += kThreads) {
            const int2 e = tbl[idx];
            if (e.x < 0) continue;
            const uint2 *src = in_px + e.x + (size_t)g0 * plane;
            uint2 *dst = xs + e.y;
            for (int gg = 0; gg < ng; ++gg) cp_async8(dst + gg * xrows * p.PW, src + (size_t)gg * plane);
        }
        const uint2 *wg = wsrc + (size_t)g0 * K2 * kCM;
        const int nw16 = ng * K2 * kCM / 2;  // 16-byte chunks
        for (int idx = tid; idx < nw16; idx += kThreads) cp_async16(wsm + idx * 2, wg + idx * 2);
    };

    // accumulators hold acc+32768 in [0,65535], pre-shifted left by k2 in the SCALED form
    const int half = 1 << (p.so - 1);
    const int k2 = p.so - 8;
    const int nmask = ~((1 << k2) - 1);
    const int ubound = 65535 << (SCALED ? k2 : 0);
    int acc[kTMC][TP];
    {
        const int16_t *bias = static_cast<const int16_t *>(p.bias);
#pragma unroll
        for (int c = 0; c < kTMC; ++c) {
            int m = mb * kCM + wm * kTMC + c;
            long long b = (m < p.OFM) ? (long long)bias[m] : 0;
            long long base = round_shift64(b, p.sb);
            // |rs(P,so)| <= rb on this path, so clamping base+32768 to [-rb, 65535+rb] cannot change
            // clamp(base+32768+r, 0, 65535); it keeps the (scaled) state inside int32
            const long long rb = (SCALED ? (((long long)NW << 25) >> k2) : (1LL << 26)) + 2;
            long long boff = base + 32768;
            if (boff > 65535 + rb) boff = 65535 + rb;
            if (boff < -rb) boff = -rb;
            const int init = SCALED ? (int)(boff * (1LL << k2)) : (int)boff;
#pragma unroll
            for (int q = 0; q < TP; ++q) acc[c][q] = init;
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
        const uint2 *wsm = sm + (st & 1) * stage_px + wm * kTMC;
        const uint2 *xs = sm + (st & 1) * stage_px + w_stage_px;
        const int ng = min(p.GS, p.G - st * p.GS);
        if constexpr (NW > 1) {
            // one step per (NW-word group, tap): partial sums over the group's words, then round + saturate once
            for (int gg = 0; gg < ng; gg += NW) {
#pragma unroll 1
                for (int i = 0; i < KS; ++i) {
                    int xo = xoff[0];
#pragma unroll
                    for (int t = 1; t < KS; ++t) xo = (i == t) ? xoff[t] : xo;
#pragma unroll 1
                    for (int j = 0; j < KS; ++j) {
                        int plo[kTMC][TP], phi[kTMC][TP];
#pragma unroll
                        for (int c = 0; c < kTMC; ++c)
#pragma unroll
                            for (int q = 0; q < TP; ++q) { plo[c][q] = half; phi[c][q] = 0; }
#pragma unroll
                        for (int w = 0; w < NW; ++w) {
                            if (gg + w >= ng) break;            // the layer's last group may be narrower (uniform)
                            const uint2 *xr = xs + (gg + w) * xrows * p.PW + xo + j;
                            const uint2 *wg = wsm + (gg + w) * K2 * kCM + (i * KS + j) * kCM;
                            uint2 xv[TP], wv[kTMC];
#pragma unroll
                            for (int q = 0; q < TP; ++q) xv[q] = xr[q];
#pragma unroll
                            for (int c = 0; c < kTMC; ++c) wv[c] = wg[c];
#pragma unroll
                            for (int c = 0; c < kTMC; ++c)
#pragma unroll
                                for (int q = 0; q < TP; ++q) {
                                    plo[c][q] = dp2a_hi_su((int)xv[q].y, wv[c].x, dp2a_lo_su((int)xv[q].x, wv[c].x, plo[c][q]));
                                    phi[c][q] = dp2a_hi_ss((int)xv[q].y, (int)wv[c].y, dp2a_lo_ss((int)xv[q].x, (int)wv[c].y, phi[c][q]));
                                }
                        }
#pragma unroll
                        for (int c = 0; c < kTMC; ++c)
#pragma unroll
                            for (int q = 0; q < TP; ++q)
                                acc[c][q] = __vimin_s32_relu((acc[c][q] + (plo[c][q] >> 8) + phi[c][q]) & nmask, ubound);
                    }
                }
            }
        } else
        for (int gg = 0; gg < ng; ++gg) {
            const uint2 *xg = xs + gg * xrows * p.PW;
            const uint2 *wg = wsm + gg * K2 * kCM;
#pragma unroll 1
            for (int i = 0; i < KS; ++i) {
                uint2 xv[XW];
                // select instead of a runtime-indexed array (keeps xoff in registers)
                int xo = xoff[0];
#pragma unroll
                for (int t = 1; t < KS; ++t) xo = (i == t) ? xoff[t] : xo;
                const uint2 *xr = xg + xo;
#pragma unroll
                for (int q = 0; q < XW; ++q) xv[q] = xr[q];
#pragma unroll
                for (int j = 0; j < KS; ++j) {
                    uint2 wv[kTMC];
#pragma unroll
                    for (int c = 0; c < kTMC; ++c) wv[c] = wg[(i * KS + j) * kCM + c];
#pragma unroll
                    for (int c = 0; c < kTMC; ++c)
#pragma unroll
                        for (int q = 0; q < TP; ++q) {
                            const uint2 x = xv[q + j];
                            int plo = dp2a_lo_su((int)x.x, wv[c].x, half);
                            plo = dp2a_hi_su((int)x.y, wv[c].x, plo);
                            if (SCALED) {
                                int phi = dp2a_lo_ss((int)x.x, (int)wv[c].y, acc[c][q] + (plo >> 8));
                                phi = dp2a_hi_ss((int)x.y, (int)wv[c].y, phi);
                                acc[c][q] = __vimin_s32_relu(phi & nmask, ubound);
                            } else {
                                int phi = dp2a_lo_ss((int)x.x, (int)wv[c].y, plo >> 8);
                                phi = dp2a_hi_ss((int)x.y, (int)wv[c].y, phi);
                                acc[c][q] = __viaddmin_s32_relu(acc[c][q], phi >> k2, 65535);
                            }
                        }
                }
            }
        }
        __syncthreads();
    }

    // largest |output| of the layer for the consumer's no-saturation bound (csrc/conv_i16_tc2.cu): warp -> CTA -> one global atomic
    __shared__ int s_amax;
    const bool track = p.xmax_out != nullptr;          // uniform
    if (track && tid == 0) s_amax = 0;
    int amax = 0;
    if (active && mb * kCM + wm * kTMC < p.OFM) {
        int16_t *out = static_cast<int16_t *>(p.out) + (size_t)f * p.out_frame_stride +
                       (((size_t)(mb * (kCM / 4) + wm) * p.H + y) * p.W + sx * TP) * 4;
#pragma unroll
        for (int q = 0; q < TP; ++q) {
            if (sx * TP + q >= p.W) break;
            int v[kTMC];
#pragma unroll
            for (int c = 0; c < kTMC; ++c) {
                int a = (SCALED ? (acc[c][q] >> k2) : acc[c][q]) - 32768;
                if (p.leaky && a < 0) a = a / 10;  // C division, truncates toward zero
                amax = max(amax, a < 0 ? -a : a);
                v[c] = a & 0xffff;
            }
            uint2 o = make_uint2((unsigned)v[0] | ((unsigned)v[1] << 16), (unsigned)v[2] | ((unsigned)v[3] << 16));
            *reinterpret_cast<uint2 *>(out + q * 4) = o;
        }
    }
    if (track) {
        __syncthreads();                               // s_amax = 0 is visible (every thread of the CTA gets here)
        amax = __reduce_max_sync(0xffffffffu, amax);
        if (lane == 0 && amax > 0) atomicMax(&s_amax, amax);
        __syncthreads();
        if (tid == 0 && s_amax > 0) atomicMax(p.xmax_out, s_amax);
    }
}

// Builds the device weight layout from one layer of the reference's reorganised blob.
__global__ void wprep_i16_kernel(const int16_t *__restrict__ blob, uint2 *__restrict__ dst, int ifm, int ofm,
                                 int ksize, int TM, int TN, int G, int total)
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
    unsigned lo = 0, hi = 0;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        int c = g * 4 + t;
        int w = 0;
        if (m < ofm && c < ifm) w = blob[reorg_woff(m, c, tap, ifm, ofm, k2, TM, TN)];
        lo |= (unsigned)(w & 0xff) << (8 * t);
        hi |= (unsigned)((w >> 8) & 0xff) << (8 * t);
    }
    dst[idx] = make_uint2(lo, hi);
}

// Contract-complete fallback: any Ksize<=3 / Kstride / Padding / TN / shift, planar layout,
// reorganised weights read in place, 64-bit arithmetic exactly as the reference writes it.
__global__ void conv_i16_generic_kernel(const int16_t *__restrict__ in, int16_t *__restrict__ out,
                                        const int16_t *__restrict__ w, const int16_t *__restrict__ bias,
                                        int ifm, int ofm, int ksize, int kstride, int iw, int ih, int ow, int oh,
                                        int pad, int is_nl, int TM, int TN, int so, int sb)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int m = blockIdx.z;
    if (x >= ow) return;
    const int iwa = align8(iw), owa = align8(ow), k2 = ksize * ksize;
    long long acc = round_shift64((long long)bias[m], sb);
    const int groups = (ifm + TN - 1) / TN;
    for (int g = 0; g < groups; ++g) {
        const int n0 = g * TN, tnn = min(TN, ifm - n0);
        for (int i = 0; i < ksize; ++i)
            for (int j = 0; j < ksize; ++j) {
                const int iy = y * kstride + i - pad, ix = x * kstride + j - pad;
                long long P = 0;
                if (iy >= 0 && iy < ih && ix >= 0 && ix < iw) {
                    for (int t = 0; t < tnn; ++t) {
                        int wv = w[reorg_woff(m, n0 + t, i * ksize + j, ifm, ofm, k2, TM, TN)];
                        int xv = in[((size_t)(n0 + t) * ih + iy) * iwa + ix];
                        P += (long long)(wv * xv);
                    }
                }
                acc += round_shift64(P, so);
                acc = acc > 32767 ? 32767 : (acc < -32768 ? -32768 : acc);
            }
    }
    int v = (int)acc;
    if (is_nl && v < 0) v = v / 10;
    out[((size_t)m * oh + y) * owa + x] = (int16_t)v;
}

// Smallest smem row pitch (pixels) >= min_pw for which the 32 lanes of a warp, reading one pixel
// word each at (row_local*PW + sx*TP), hit distinct banks within each LSU phase.
int pick_pitch(int min_pw, int TP, int SW, int px_words)
{
    const int lanes_per_phase = 32 / px_words;  // 64-bit loads: 16 lanes per phase; 128-bit: 8
    for (int pw = min_pw; pw < min_pw + 32; ++pw) {
        bool ok = true;
        for (int base = 0; base < 32 && ok; base += lanes_per_phase) {
            unsigned used = 0;
            for (int l = base; l < base + lanes_per_phase; ++l) {
                int row = l / SW, sx = l % SW;
                int slot = (row * pw + sx * TP) % lanes_per_phase;
                if (used & (1u << slot)) { ok = false; break; }
                used |= 1u << slot;
            }
        }
        if (ok) return pw;
    }
    return min_pw;
}

}  // namespace

size_t conv_fast_plan(ConvFastParams &p, int ksize, int elem_bytes)
{
    if (ksize != 1 && ksize != 3) return 0;
    if (p.W <= 0 || p.H <= 0 || p.B <= 0) return 0;
    // segment width: 13 when it tiles the row exactly (every YOLOv2-416 width) or nearly, else the better of 13 / 7;
    // the float kernel always uses 7 (a 16-byte pixel word doubles the register cost of a segment)
    int tp = 13;
    const int nw = (elem_bytes == 2 && p.group_words > 1) ? p.group_words : 1;
    if (elem_bytes == 4 || nw > 1) tp = 7;
    else if (p.W % 13 != 0) {
        // 13-pixel segments reuse each weight over more pixels and give taller row bands (measured 3.6-3.8 T steps/s against
        // 2.3-3.6 for 7): keep them while the ragged last segment wastes < 10 % (every width of the 608 net: 38 ... 608)
        double u13 = (double)p.W / (ceil_div(p.W, 13) * 13), u7 = (double)p.W / (ceil_div(p.W, 7) * 7);
        if (u13 < 0.9 && u7 > u13) tp = 7;
    }
    int sw = ceil_div(p.W, tp);
    if (sw > kNS && tp == 7 && elem_bytes != 4 && nw == 1 && ceil_div(p.W, 13) <= kNS) { tp = 13; sw = ceil_div(p.W, 13); }
    if (sw > kNS) return 0;
    p.TP = tp;
    p.SW = sw;
    p.RB = kNS / sw;
    const int px_bytes = 4 * elem_bytes;
    p.PW = pick_pitch(sw * tp + ksize - 1, tp, sw, px_bytes / 4);
    const int xrows = p.RB + ksize - 1 + 1;
    const size_t per_group = ((size_t)xrows * p.PW + (size_t)ksize * ksize * kCM) * px_bytes;
    const size_t tbl_bytes = (size_t)(p.RB + ksize - 1) * p.W * 8;  // per-CTA copy table
    // two CTAs per SM: keep 2 x (2 stages + table) under ~200 KB of the 227 KB
    size_t budget = 100 * 1024 > tbl_bytes + 2 * per_group ? (100 * 1024 - tbl_bytes) / 2 : per_group;
    int gs = (int)(budget / per_group);
    if (gs < 1) gs = 1;
    if (gs > 8) gs = 8;
    if (gs > p.G) gs = p.G;
    if (nw > 1) gs = gs < nw ? nw : gs / nw * nw;      // a rounding group never straddles two pipeline stages
    p.GS = gs;
    size_t smem = 2 * per_group * gs + 64 + tbl_bytes;
    if (smem > 200 * 1024) return 0;
    return smem;
}


template <int TP, int KS, bool SCALED>
static int launch_i16_variant(const ConvFastParams &p, size_t smem, cudaStream_t st)
{
    // set on every launch (a microsecond): the attribute is per device and a host may drive several GPUs from several threads,
    // so a cached "already configured" flag would be a data race for nothing
    cudaFuncSetAttribute(conv_i16_c4_kernel<TP, KS, SCALED>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    dim3 grid(ceil_div(p.B * p.H, p.RB), ceil_div(p.OFM, kCM));
    conv_i16_c4_kernel<TP, KS, SCALED><<<grid, kThreads, smem, st>>>(p);
    return 1;
}

template <int KS, int NW>
static int launch_i16_group_variant(const ConvFastParams &p, size_t smem, cudaStream_t st)
{
    cudaFuncSetAttribute(conv_i16_c4_kernel<7, KS, true, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    dim3 grid(ceil_div(p.B * p.H, p.RB), ceil_div(p.OFM, kCM));
    conv_i16_c4_kernel<7, KS, true, NW><<<grid, kThreads, smem, st>>>(p);
    return 1;
}

int launch_conv_i16_fast(const ConvFastParams &p, int ksize, cudaStream_t st, const char **variant)
{
    const int xrows = p.RB + ksize - 1 + 1;
    const size_t smem = 2 * (((size_t)p.GS * ((size_t)xrows * p.PW + (size_t)ksize * ksize * kCM) + 1) & ~(size_t)1) * 8 +
                        (size_t)(p.RB + ksize - 1) * p.W * 8;  // + the copy table
    const int tp = p.TP;
    const bool scaled = p.so <= 22;
    if (p.group_words > 1) {     // reference built with Tn = 8 / 16 (conv_fast_plan chose 7-pixel segments and GS % group_words == 0)
        if (tp != 7 || !scaled || (p.group_words != 2 && p.group_words != 4) || p.GS % p.group_words) return -1;
        if (variant) *variant = p.group_words == 2 ? (ksize == 3 ? "conv_i16_c4<7,3,tn8>" : "conv_i16_c4<7,1,tn8>") : (ksize == 3 ? "conv_i16_c4<7,3,tn16>" : "conv_i16_c4<7,1,tn16>");
        if (ksize == 3) return p.group_words == 2 ? launch_i16_group_variant<3, 2>(p, smem, st) : launch_i16_group_variant<3, 4>(p, smem, st);
        return p.group_words == 2 ? launch_i16_group_variant<1, 2>(p, smem, st) : launch_i16_group_variant<1, 4>(p, smem, st);
    }
#define Y2_VARIANT(TPV, KSV)                                                                                  \
    if (tp == TPV && ksize == KSV) {                                                                           \
        if (variant) *variant = scaled ? "conv_i16_c4<" #TPV "," #KSV ",scaled>" : "conv_i16_c4<" #TPV "," #KSV ",unscaled>"; \
        return scaled ? launch_i16_variant<TPV, KSV, true>(p, smem, st) : launch_i16_variant<TPV, KSV, false>(p, smem, st); \
    }
    Y2_VARIANT(13, 3)
    Y2_VARIANT(13, 1)
    Y2_VARIANT(7, 3)
    Y2_VARIANT(7, 1)
#undef Y2_VARIANT
    return -1;
}

size_t wprep_bytes(int ifm, int ofm, int ksize, int elem_bytes)
{
    return (size_t)ceil_div(ofm, kCM) * ceil_div(ifm, 4) * ksize * ksize * kCM * 4 * elem_bytes;
}

void launch_wprep_i16(const int16_t *blob, void *dst, int ifm, int ofm, int ksize, int TM, int TN, cudaStream_t st)
{
    const int G = ceil_div(ifm, 4);
    const int total = ceil_div(ofm, kCM) * G * ksize * ksize * kCM;
    wprep_i16_kernel<<<ceil_div(total, 256), 256, 0, st>>>(blob, static_cast<uint2 *>(dst), ifm, ofm, ksize, TM, TN, G, total);
}

void launch_conv_i16_generic(const int16_t *in, int16_t *out, const int16_t *w, const int16_t *bias,
                             int ifm, int ofm, int ksize, int kstride, int iw, int ih, int ow, int oh,
                             int pad, int is_nl, int TM, int TN, int so, int sb, cudaStream_t st)
{
    dim3 grid(ceil_div(ow, 128), oh, ofm);
    conv_i16_generic_kernel<<<grid, 128, 0, st>>>(in, out, w, bias, ifm, ofm, ksize, kstride, iw, ih, ow, oh, pad,
                                                  is_nl, TM, TN, so, sb);
}

}  // namespace y2
