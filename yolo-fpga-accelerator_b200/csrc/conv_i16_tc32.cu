// INT16 convolution on the tensor cores for a reference BUILT with rounding group Tn = 32, 16 or 8
// (scripts/hw_params_gen.py --tn 32 / 16 / 8), bit-exact to that build.  The text below describes Tn = 32; the
// template parameter TNW generalises it: a K = 32 slice of the MMA holds S = 32 / TNW consecutive chain steps as a block-diagonal
// activation operand (column (step s', pixel) is non-zero only in K rows TNW*s' .. TNW*s' + TNW - 1, the weights of step s' sit in
// the same K rows of the A tile), a tile is TNW pixels x S steps (N = 32 columns, column = s' * TNW + pixel), a CTA owns 3 * TNW
// consecutive pixels, and the epilogue applies a tile's S steps to its TNW accumulators in chain order.  Step slots past the
// layer's last step carry zero operands: d = (0 + half) >> so = 0, a no-op on the already saturated accumulator.
//
// With Tn = 32 one step of the reference's chain (hls/core/core_compute.cpp:65-120) is
//   acc = clamp16(acc + ((P + half) >> so)),  P = sum_{t<32} w[m][32g+t][tap] * x[32g+t][pixel+tap],
// i.e. exactly ONE K = 32 slice of an int8 MMA per byte-plane pair: no block-diagonal operand, no wasted K rows, and 8x fewer
// round-and-saturate steps per MAC than the default Tn = 4 build.  Same decomposition as csrc/conv_i16_tc2.cu:
//   HH = sum wh*xh,  M = sum (wh*xl + wl*xh),  LL = sum wl*xl,   P = 65536*HH + 256*M + LL   (four tcgen05.mma kind::i8),
// one TMEM column per (step, pixel).  Per step and output the CUDA cores do
//   t = 256*M + LL + half;  d = HH * 2^(16-so) + (t >> so);  acc = max(min(acc + d, 65535), 0)      (8 <= so <= 16).
//
// Tile = 128 output channels x 32 pixels x one step (N = 32 columns per plane, 96 TMEM columns, five buffers).  A CTA owns 96
// consecutive pixels = three tiles; tile r of every K slice belongs to epilogue group r (its accumulators stay in registers for
// the whole layer), to the builder warps r and r + 3 (even / odd slices: each lane gathers the channels of its column from the
// staged C4 patch and writes the hi/lo byte planes of its B-operand row) and to issuer warp r.  Two staging warps walk the copy
// table of the CTA's activation band (cp.async, chunks of 16 C4 planes, double buffered) - on warps of their own, because the
// builders' proxy fence (MEMBAR.ALL.CTA) waits for the issuing warp's outstanding copies.  Weights are the A operand from shared
// memory (one 8 KB canonical tile per K slice, streamed by cp.async.bulk through a 4-slot ring): with N = 32 the MMA is
// shared-memory bound at 40 cycles (profiles/microbench/umma_issue.cu), 160 cycles per 4096-step tile.  The barrier ring has
// 12 = 4 slices x 3 tiles slots: slot -> fixed tile index r.  What bounds it (per-tile timelines, profiles/r2_tc32_timeline_*.txt):
// the three-builder form of round 1 was builder bound (~1400 cycles per K slice); this form waits for the weight ring
// (~2.9 k cycles per 8 KB tile, four slices deep).  DESIGN.md section 4, "Rounding-group variants".
#include "common.cuh"
#ifdef Y2_TC32_PROFILE
#include <cstdio>
#include <cstring>
#endif

namespace y2 {

namespace {

constexpr int kM = 128;             // output channels per CTA = TMEM lanes
constexpr int kN = 32;              // pixels per tile = MMA N
constexpr int kR = 3;               // tiles per step -> 96 pixels per CTA; tile r <-> epilogue group r, builder r, issuer r
constexpr int kPT = kN * kR;        // pixels per CTA for Tn = 32 (the largest: table sizes); a TNW variant owns kR * TNW pixels
constexpr int kBufs = 5;            // TMEM accumulator buffers: HH | M | LL, 32 columns each
constexpr int kBufCols = 3 * kN;
constexpr int kRing = 12;           // activation tile ring (hi 1 KB | lo 1 KB) and barrier ring: 4 steps x 3 tiles
#ifndef Y2_TC32_WRING
#define Y2_TC32_WRING 4
#endif
constexpr int kWRing = Y2_TC32_WRING;   // weight ring: one 8 KB step tile per slot (hi 4 KB | lo 4 KB, canonical K-major)
// experiment builds (profiles/build_variant_tc32.sh; wrong results, timing only): -DY2_TC32_EXP=1 no step arithmetic, =2 no TMEM
// read-out (arithmetic on opaque registers), =3 no MMAs (the issuers only commit)
#ifndef Y2_TC32_EXP
#define Y2_TC32_EXP 0
#endif
// -DY2_TC32_PROFILE: per-tile timeline (clock64) of CTA (1,0), tiles Y2_TC32_TL0 .. +47 (tile = 3 * slice + r), printed by the launcher:
// 0 builder passed mma_done (slot free) | 1 built + arrived | 2 issuer passed w_full | 3 issuer passed go | 4 issued + committed |
// 5 epilogue (quadrant 0) passed mma_done | 6 first half read | 7 released | 8 computed | 9 loader passed w_empty | 10 loader issued
#ifdef Y2_TC32_PROFILE
#ifndef Y2_TC32_TL0
#define Y2_TC32_TL0 300
#endif
__device__ long long g_tc32_tl[48 * 16];
#define PROF_TL(tile, ev) do { if (blockIdx.x == 1 && blockIdx.y == 0 && lane == 0 && (tile) >= Y2_TC32_TL0 && (tile) < Y2_TC32_TL0 + 48) g_tc32_tl[((tile) - Y2_TC32_TL0) * 16 + (ev)] = clock64(); } while (0)
#else
#define PROF_TL(tile, ev)
#endif
constexpr int kEpiWarps = 12;       // warps 0-11: group kg = warp/4, TMEM lane quadrant = warp%4
constexpr int kBuilder = 12;        // warps 12-17: builder b builds tile b % 3 of the K slices with parity b / 3
constexpr int kNB = 6;
constexpr int kStager = 18;         // warp 18: activation staging (cp.async copy table walk), double-buffered chunks
constexpr int kLoader = 19;         // warp 19: weight ring
constexpr int kIssuer = 20;         // warps 20-22
constexpr int kStager2 = 23;        // warp 23: second half of the staging copies
constexpr int kThreads = 24 * 32;
// setmaxnreg moves registers inside the CTA's own pool only (USETMAXREG.TRY_ALLOC.CTAPOOL): what the 12 helper warps hand back
// (launch allocation 80 per thread at 768 threads) must cover what the 12 epilogue warps take: 80 - 48 >= 112 - 80.  (120 / 48 adds up
// against the 64 K register file but not against the pool: the epilogue warps spin in TRY_ALLOC for ever - the hang of the first
// six-builder build.)
constexpr int kLaunchRegs = 80, kEpiRegs = 112, kHelperRegs = 48;
static_assert(kEpiRegs - kLaunchRegs <= kLaunchRegs - kHelperRegs && kLaunchRegs * kThreads <= 65536, "setmaxnreg pool");
constexpr int kWBytes = 2 * kM * 32;
constexpr int kBBytes = kN * 32;    // one plane of one activation tile
static_assert(kRing % kR == 0 && kRing % kBufs != 0, "slot -> fixed tile index");

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(void *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(void *bar, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(void *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, void *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void umma_commit(void *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void umma_i8_ss(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned idesc, unsigned accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(unsigned taddr, int *r)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
// ties the registers of an asynchronous tcgen05.ld to the point after tcgen05.wait::ld
__device__ __forceinline__ void reg_fence16(int *r)
{
    asm volatile(""
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                   "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])::"memory");
}
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gsrc));
}

// K-major, no-swizzle canonical operand: core matrix = 8 rows x 16 bytes, contiguous (128 B);
// the two K chunks of a 32-byte row are LBO = 128 B apart, 8-row groups are SBO = 256 B apart.
__device__ __forceinline__ unsigned long long smem_desc(const void *p)
{
    unsigned long long d = (unsigned long long)((smem_u32(p) >> 4) & 0x3FFF);
    d |= (unsigned long long)(128 >> 4) << 16;
    d |= (unsigned long long)(256 >> 4) << 32;
    d |= 1ull << 46;  // descriptor version for sm_100
    return d;
}
__host__ __device__ constexpr unsigned idesc_i8(int a_signed, int b_signed)
{
    return (2u << 4) | ((unsigned)a_signed << 7) | ((unsigned)b_signed << 10) | ((unsigned)(kN >> 3) << 17) | ((unsigned)(kM >> 4) << 24);
}
__host__ __device__ inline int operand_off(int row, int k) { return (((row >> 3) * 2 + (k >> 4)) * 8 + (row & 7)) * 16 + (k & 15); }


struct Tc32Params {
    const uint2 *in;          // C4 input
    int16_t *out;             // C4 output (already offset to the first output group)
    const unsigned char *w;   // [mtile][step][hi 4 KB | lo 4 KB] canonical operand tiles
    const int16_t *bias;
    int B, H, W, G, OFM;      // G = C4 groups of the input
    long long in_frame_stride, out_frame_stride;  // elements
    int sb, leaky;
    int nsteps;               // chain steps per output: ceil(IFM/Tn) * K*K
    int nslices;              // MMA K slices: ceil(nsteps / (32/Tn))
    int ctab_cap;             // entries reserved for the activation copy table
    int PW, rows_max, gs_shift;  // staging: smem row pitch (pixels), band rows incl. halo + zero row, log2(C4 groups per chunk) >= 3
};

template <int SO>
__device__ __forceinline__ int tc32_step(int acc, int hh, int mm, int ll)
{
    const int t = mm * 256 + ll + (1 << (SO - 1));         // P + half without its 65536*HH part
    const int d = hh * (1 << (16 - SO)) + (t >> SO);       // 65536*HH is a multiple of 2^so
    return __viaddmin_s32_relu(acc, d, 65535);
}

template <int KS, int SO, int TNW>
__global__ void __launch_bounds__(kThreads, 1) conv_i16_tc32_kernel(const Tc32Params p)
{
    constexpr int K2 = KS * KS;
    constexpr int PAD = KS / 2;
    constexpr int S = 32 / TNW;         // chain steps per K slice
    constexpr int PPT = kN / S;         // pixels per tile (= TNW)
    constexpr int kPTv = PPT * kR;      // pixels per CTA
    constexpr int GW = TNW / 4;         // C4 words per rounding group
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *sW = smem;                                    // kWRing x 8 KB
    unsigned char *sB = sW + kWRing * kWBytes;                   // kRing x (hi 1 KB | lo 1 KB)
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(sB + kRing * 2 * kBBytes);
    unsigned long long *w_full = bars, *w_empty = w_full + kWRing, *go = w_empty + kWRing, *mma_done = go + kRing;
    // go[slot]: the tile in this ring slot may be issued = its activation tile is built (1 arrival, builder) AND its TMEM buffer
    // it % 5 has been read out by the epilogue of tile it-5 (4 arrivals; pre-arrived for the first five tiles)
    // x_full[b]: staging chunk buffer b holds its chunk (64 arrivals: the two staging warps' lanes, each after its own copies landed);
    // x_empty[b]: every builder warp is past the slices that read it (6 arrivals)
    unsigned long long *x_full = mma_done + kRing, *x_empty = x_full + 2;
    unsigned *tmem_slot = reinterpret_cast<unsigned *>(x_empty + 2 + 1);
    int *pxtab = reinterpret_cast<int *>(tmem_slot + 4);         // [96][4]: smem pixel offset for tap rows 0..2, valid flag
    int *rowinfo = pxtab + kPT * 4;                              // [32][2]: per staged row slot: first needed column, prefix of the copy table
    int2 *ctab = reinterpret_cast<int2 *>(rowinfo + 64);         // copy table of one C4 group plane: (global pixel offset, smem pixel offset)
    uint2 *sX = reinterpret_cast<uint2 *>(ctab + p.ctab_cap);    // 2 chunks x GS groups x rows_max x PW pixels

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long npix = (long long)p.B * p.H * p.W;
    const long long pix0 = (long long)blockIdx.x * kPTv;
    const int mtile = blockIdx.y;
    const int zero_slot = p.rows_max - 1;
    const int GS = 1 << p.gs_shift;
    const int chunk_px = GS * p.rows_max * p.PW;
    const long long row_first = pix0 / p.W;                      // global row (frame*H + y) of the first pixel

    if (tid == 0) {
        for (int i = 0; i < kWRing; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], kR); }
        for (int i = 0; i < kRing; ++i) { mbar_init(&go[i], 5); mbar_init(&mma_done[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&x_full[i], 64); mbar_init(&x_empty[i], kNB); }
        for (int i = 0; i < kBufs; ++i)
            for (int k = 0; k < 4; ++k) mbar_arrive(&go[i]);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kIssuer) {   // the first MMA warp owns the TMEM allocation
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int q = tid; q < kPTv; q += kThreads) {
        long long gp = pix0 + q;
        int valid = gp < npix;
        long long grow = valid ? gp / p.W : row_first;
        int x = valid ? (int)(gp - grow * p.W) : 0;
        int y = (int)(grow % p.H);
        int rl = (int)(grow - row_first);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            int yin = y + i - PAD;
            int slot = (valid && i < KS && yin >= 0 && yin < p.H) ? rl + i : zero_slot;
            pxtab[q * 4 + i] = slot * p.PW + x;
        }
        pxtab[q * 4 + 3] = valid;
    }
    // Activation copy table.  Staged row slot s holds global row row_first - PAD + s; only the columns some pixel of this CTA reads
    // are copied (a CTA's pixels are consecutive, so on wide images it touches a fraction of each row), and the loader walks a
    // precomputed (source, destination) list instead of doing index arithmetic per element.
    if (tid == 0) {
        const long long pix_last = (pix0 + kPTv < npix ? pix0 + kPTv : npix) - 1;
        const long long row_last = pix_last / p.W;
        const int nrow_cta = (int)(row_last - row_first) + 1;
        const int x_first = (int)(pix0 - row_first * p.W), x_last = (int)(pix_last - row_last * p.W);
        int total = 0;
        for (int s = 0; s < p.rows_max - 1; ++s) {
            int lo = p.W, hi = -1;
            const long long Rr = row_first - PAD + s;
            if (Rr >= 0 && Rr < (long long)p.B * p.H)
                for (int i = 0; i < KS; ++i) {
                    const int j = s - i;
                    if (j < 0 || j >= nrow_cta) continue;
                    const int xa = (j == 0 ? x_first : 0) - PAD, xb = (j == nrow_cta - 1 ? x_last : p.W - 1) + PAD;
                    lo = min(lo, max(xa, 0));
                    hi = max(hi, min(xb, p.W - 1));
                }
            rowinfo[2 * s] = lo;
            rowinfo[2 * s + 1] = total;
            total += hi >= lo ? hi - lo + 1 : 0;
        }
        rowinfo[2 * (p.rows_max - 1)] = 0;
        rowinfo[2 * (p.rows_max - 1) + 1] = total;      // = entries per C4 group plane
    }
    __syncthreads();
    for (int idx = tid; idx < (p.rows_max - 1) * p.W; idx += kThreads) {
        const int s = idx / p.W, x = idx - s * p.W;
        const int lo = rowinfo[2 * s], cnt = rowinfo[2 * s + 3] - rowinfo[2 * s + 1];
        if (x < lo || x >= lo + cnt) continue;
        const long long Rr = row_first - PAD + s;
        const long long ff = Rr / p.H;
        const int yy = (int)(Rr - ff * p.H);
        ctab[rowinfo[2 * s + 1] + x - lo] = make_int2((int)(ff * (p.in_frame_stride >> 2)) + yy * p.W + x, s * p.PW + PAD + x);
    }
    for (int i = tid; i < 2 * chunk_px; i += kThreads) sX[i] = make_uint2(0u, 0u);
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const unsigned tmem = *tmem_slot;

    // ===== activation staging (warps kStager and kStager2, 64 lanes): chunk c = GS consecutive C4 group planes of the CTA's band, copied
    // by walking the copy table (one cp.async per (group, needed pixel)).  The builders must never have copies of their own in flight:
    // their per-tile fence (MEMBAR) would wait for them (measured: a 5.8 k-cycle builder stall per chunk with the copies on the
    // builder warps, profiles/r2_tc32_timeline_3builders_tn32.txt) =====
    auto stage_loop = [&](int li) {
        const int nchunks = (p.G + GS - 1) >> p.gs_shift;
        const int per_group = rowinfo[2 * (p.rows_max - 1) + 1];
        const long long plane = (long long)p.H * p.W;
        for (int c = 0; c < nchunks; ++c) {
            if (c >= 2) mbar_wait(&x_empty[c & 1], ((c >> 1) - 1) & 1);
            uint2 *dst = sX + (c & 1) * chunk_px;
            const int g0 = c << p.gs_shift, ng = min(GS, p.G - g0);
            int gg = 0, rem = li;                            // entry idx = gg * per_group + rem, idx = li, li + 64, ...
            while (rem >= per_group) { rem -= per_group; ++gg; }
            while (gg < ng) {
                const int2 e = ctab[rem];
                cp_async8(dst + gg * p.rows_max * p.PW + e.y, p.in + (g0 + gg) * plane + e.x);
                rem += 64;
                while (rem >= per_group) { rem -= per_group; ++gg; }
            }
            asm volatile("cp.async.commit_group;");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            mbar_arrive(&x_full[c & 1]);
        }
    };

    if (warp >= kEpiWarps) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kHelperRegs));
        if (warp >= kIssuer) {
            // ===== three MMA issuer warps: warp iw issues tile iw of every step =====
            const int iw = warp - kIssuer;
            if (warp == kStager2) {
                stage_loop(32 + lane);
            } else if (iw < kR) {
                const unsigned long long dA0 = smem_desc(sW), dB0 = smem_desc(sB);
                constexpr unsigned long long kAStep = kWBytes >> 4, kAPlane = (kM * 32) >> 4;      // descriptor address units (16 B)
                constexpr unsigned long long kBStep = (2 * kBBytes) >> 4, kBPlane = kBBytes >> 4;
                unsigned elected;
                asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(elected));
                int slot = iw, sph = 0, tb = iw;                 // ring slot it % 12 + its phase parity, TMEM buffer it % 5 (it = 3 s + iw)
                for (int s = 0; s < p.nslices; ++s) {
                    const int ws = s % kWRing;
                    mbar_wait(&w_full[ws], (s / kWRing) & 1);
                    PROF_TL(s * kR + iw, 2);
                    mbar_wait(&go[slot], sph);
                    PROF_TL(s * kR + iw, 3);
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    if (elected) {
                        const unsigned long long dAh = dA0 + ws * kAStep, dAl = dAh + kAPlane;
                        const unsigned long long dBh = dB0 + slot * kBStep, dBl = dBh + kBPlane;
                        const unsigned d0 = tmem + tb * kBufCols;
#if Y2_TC32_EXP != 3
                        umma_i8_ss(d0, dAh, dBh, idesc_i8(1, 1), 0);            // HH
                        umma_i8_ss(d0 + kN, dAh, dBl, idesc_i8(1, 0), 0);       // M  = hi*lo
                        umma_i8_ss(d0 + kN, dAl, dBh, idesc_i8(0, 1), 1);       //    + lo*hi
                        umma_i8_ss(d0 + 2 * kN, dAl, dBl, idesc_i8(0, 0), 0);   // LL
#endif
                        umma_commit(&mma_done[slot]);                           // epilogue (tile ready) and builder (slot free)
                        umma_commit(&w_empty[ws]);                              // this warp's reads of the step's weights are done
                    }
                    __syncwarp();
                    PROF_TL(s * kR + iw, 4);
                    slot += kR;
                    if (slot >= kRing) { slot -= kRing; sph ^= 1; }
                    tb = tb >= kBufs - kR ? tb - (kBufs - kR) : tb + kR;
                }
            }
        } else if (warp == kLoader) {
            // ===== weight ring: one 8 KB bulk copy per step =====
            if (lane == 0) {
                const unsigned char *src = p.w + (size_t)mtile * p.nslices * kWBytes;
                for (int s = 0; s < p.nslices; ++s) {
                    const int ws = s % kWRing;
                    if (s >= kWRing) mbar_wait(&w_empty[ws], ((s / kWRing) - 1) & 1);
                    PROF_TL(s * kR, 9);
                    mbar_expect_tx(&w_full[ws], kWBytes);
                    bulk_g2s(sW + ws * kWBytes, src + (size_t)s * kWBytes, kWBytes, &w_full[ws]);
                    PROF_TL(s * kR, 10);
                }
            }
        } else if (warp == kStager) {
            stage_loop(lane);
        } else if (warp < kStager) {
            // ===== six builder warps; lane = one COLUMN of tile bw = (step slot sp, pixel): it gathers the TNW channels of its step's
            // rounding group (TNW/4 C4 words) at its pixel + tap and writes the hi / lo byte planes of its B-operand row (row = column,
            // K = 32 bytes: zero outside the K rows TNW*sp .. TNW*sp + TNW - 1).  Two warps per tile index, alternating K slices: one
            // warp's build of a slice is a ~1.3 k-cycle dependent chain (timeline of the three-builder form,
            // profiles/r2_tc32_timeline_3builders_tn32.txt), which was the period of the whole pipeline =====
            const int bidx = warp - kBuilder;
            const int bw = bidx % kR, par = bidx / kR;
            const int sp = lane / PPT;                  // step slot of this lane's column within a K slice
            const int bt = bw * PPT + (lane % PPT);     // the CTA pixel this lane owns
            const int r0 = operand_off(lane, 0), r1 = operand_off(lane, 16);   // this column's two 16-byte K chunks
            const int pxo0 = pxtab[bt * 4 + 0], pxo1 = pxtab[bt * 4 + 1], pxo2 = pxtab[bt * 4 + 2];
            int ready = -1, released = -1;              // highest chunk waited for / handed back
            for (int sl = par; sl < p.nslices; sl += 2) {
                const int slot = bw + kR * (sl & 3), sph = (sl >> 2) & 1;       // ring slot of tile (sl, bw) and its phase parity
                // the chunks holding the C4 words of this slice's first and last step (warp-uniform; a slice spans at most two)
                const int st_lo = sl * S, st_hi = min(st_lo + S, p.nsteps) - 1;
                const int c_lo = ((st_lo / K2) * GW) >> p.gs_shift, c_hi = ((st_hi / K2) * GW) >> p.gs_shift;
                while (released + 1 < c_lo) {           // this warp reads nothing below chunk c_lo any more
                    ++released;
                    if (lane == 0) mbar_arrive(&x_empty[released & 1]);
                }
                while (ready < c_hi) {
                    ++ready;
                    mbar_wait(&x_full[ready & 1], (ready >> 1) & 1);
                }
                if (sl >= kRing / kR) mbar_wait(&mma_done[slot], sph ^ 1);   // the MMAs of the previous tile in this slot have read it
                PROF_TL(sl * kR + bw, 0);
                const int step = st_lo + sp;
                unsigned hi[GW], lo[GW];
#pragma unroll
                for (int j = 0; j < GW; ++j) hi[j] = lo[j] = 0u;
                if (step < p.nsteps) {
                    const int g = step / K2, tap = step - g * K2;
                    const int ti = tap / KS, tj = tap - ti * KS;
                    const int w0 = g * GW;                           // first C4 word of the rounding group
                    const uint2 *xs = sX + ((w0 >> p.gs_shift) & 1) * chunk_px + (w0 & (GS - 1)) * p.rows_max * p.PW +
                                      (ti == 0 ? pxo0 : ti == 1 ? pxo1 : pxo2) + tj;
#pragma unroll
                    for (int j = 0; j < GW; ++j) {
                        uint2 x = make_uint2(0u, 0u);
                        if (w0 + j < p.G) x = xs[j * p.rows_max * p.PW];
                        hi[j] = __byte_perm(x.x, x.y, 0x7531);
                        lo[j] = __byte_perm(x.x, x.y, 0x6420);
                    }
                }
                uint4 h0, h1, l0, l1;                                // K bytes 0..15 and 16..31 of the row, per plane
                const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
                if constexpr (TNW == 32) {
                    h0 = make_uint4(hi[0], hi[1], hi[2], hi[3]); h1 = make_uint4(hi[4], hi[5], hi[6], hi[7]);
                    l0 = make_uint4(lo[0], lo[1], lo[2], lo[3]); l1 = make_uint4(lo[4], lo[5], lo[6], lo[7]);
                } else if constexpr (TNW == 16) {
                    const uint4 hv = make_uint4(hi[0], hi[1], hi[2], hi[3]), lv = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    h0 = sp == 0 ? hv : z4; h1 = sp == 0 ? z4 : hv;
                    l0 = sp == 0 ? lv : z4; l1 = sp == 0 ? z4 : lv;
                } else {
                    const uint4 hv = (sp & 1) ? make_uint4(0u, 0u, hi[0], hi[1]) : make_uint4(hi[0], hi[1], 0u, 0u);
                    const uint4 lv = (sp & 1) ? make_uint4(0u, 0u, lo[0], lo[1]) : make_uint4(lo[0], lo[1], 0u, 0u);
                    h0 = sp < 2 ? hv : z4; h1 = sp < 2 ? z4 : hv;
                    l0 = sp < 2 ? lv : z4; l1 = sp < 2 ? z4 : lv;
                }
                unsigned char *bh = sB + (slot * 2) * kBBytes;
                *reinterpret_cast<uint4 *>(bh + r0) = h0;
                *reinterpret_cast<uint4 *>(bh + r1) = h1;
                *reinterpret_cast<uint4 *>(bh + kBBytes + r0) = l0;
                *reinterpret_cast<uint4 *>(bh + kBBytes + r1) = l1;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&go[slot]);
                PROF_TL(sl * kR + bw, 1);
            }
            // (a warp whose last slice ends below the last chunks still owes their hand-back: the staging warp waits for all six)
            const int nchunks = (p.G + GS - 1) >> p.gs_shift;
            while (released + 1 < nchunks - 2) {
                ++released;
                if (lane == 0) mbar_arrive(&x_empty[released & 1]);
            }
        }
    } else {
        // ===== epilogue warps: thread = one output channel (TMEM lane); group kg = warp/4 owns tile kg (32 pixels) of every step =====
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kEpiRegs));
        const int q4 = warp & 3, kg = warp >> 2;
        const int m = mtile * kM + q4 * 32 + lane;
        const unsigned lane_base = tmem + ((unsigned)(q4 * 32) << 16);
        int U[PPT];                                 // acc + 32768 of the tile's PPT pixels; column n = step slot * PPT + pixel
        {
            long long bv = (m < p.OFM) ? (long long)p.bias[m] : 0;
            long long bs = round_shift64(bv, p.sb);
            const long long rb = (1LL << (38 - SO)) + 2;      // |(P + half) >> so| <= 2^(38-so) for <= 32 products: clamping the bias term there cannot change clamp16(bias + r)
            long long boff = bs + 32768;
            if (boff > 65535 + rb) boff = 65535 + rb;
            if (boff < -rb) boff = -rb;
#pragma unroll
            for (int j = 0; j < PPT; ++j) U[j] = (int)boff;
        }
        int slot = kg, sph = 0, tb = kg;
        for (int s = 0; s < p.nslices; ++s) {
            mbar_wait(&mma_done[slot], sph);
            if (q4 == 0) PROF_TL(s * kR + kg, 5);
            asm volatile("tcgen05.fence::after_thread_sync;");
            const unsigned base = lane_base + tb * kBufCols;
            int hh[16], mm[16], ll[16];
#if Y2_TC32_EXP == 2
#pragma unroll
            for (int n = 0; n < 16; ++n) asm volatile("" : "=r"(hh[n]), "=r"(mm[n]), "=r"(ll[n]));
#else
            tmem_ld16(base, hh); tmem_ld16(base + kN, mm); tmem_ld16(base + 2 * kN, ll);          // columns 0-15
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#endif
            reg_fence16(hh); reg_fence16(mm); reg_fence16(ll);
            if (q4 == 0) PROF_TL(s * kR + kg, 6);
#if Y2_TC32_EXP != 1
#pragma unroll
            for (int n = 0; n < 16; ++n) U[n % PPT] = tc32_step<SO>(U[n % PPT], hh[n], mm[n], ll[n]);     // (unrolled in column order = chain order per pixel)
#else
#pragma unroll
            for (int n = 0; n < 16; ++n) U[n % PPT] ^= hh[n] ^ mm[n] ^ ll[n];
#endif
#if Y2_TC32_EXP == 2
#pragma unroll
            for (int n = 0; n < 16; ++n) asm volatile("" : "=r"(hh[n]), "=r"(mm[n]), "=r"(ll[n]));
#else
            tmem_ld16(base + 16, hh); tmem_ld16(base + kN + 16, mm); tmem_ld16(base + 2 * kN + 16, ll);   // columns 16-31
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#endif
            reg_fence16(hh); reg_fence16(mm); reg_fence16(ll);
            // the whole tile has been read: hand the TMEM buffer to tile it+5
            asm volatile("tcgen05.fence::before_thread_sync;");
            __syncwarp();
            if (lane == 0) mbar_arrive(&go[slot + kBufs < kRing ? slot + kBufs : slot + kBufs - kRing]);
            if (q4 == 0) PROF_TL(s * kR + kg, 7);
#if Y2_TC32_EXP != 1
#pragma unroll
            for (int n = 0; n < 16; ++n) U[(16 + n) % PPT] = tc32_step<SO>(U[(16 + n) % PPT], hh[n], mm[n], ll[n]);
#else
#pragma unroll
            for (int n = 0; n < 16; ++n) U[(16 + n) % PPT] ^= hh[n] ^ mm[n] ^ ll[n];
#endif
            if (q4 == 0) PROF_TL(s * kR + kg, 8);
            slot += kR;
            if (slot >= kRing) { slot -= kRing; sph ^= 1; }
            tb = tb >= kBufs - kR ? tb - (kBufs - kR) : tb + kR;
        }
        if (m < p.OFM) {
#pragma unroll
            for (int j = 0; j < PPT; ++j) {
                const long long gp = pix0 + kg * PPT + j;
                if (gp >= npix) continue;
                const long long grow = gp / p.W;
                const int x = (int)(gp - grow * p.W);
                const long long f = grow / p.H;
                const int y = (int)(grow - f * p.H);
                int a = U[j] - 32768;
                if (p.leaky && a < 0) a = a / 10;
                p.out[f * p.out_frame_stride + (((long long)(m >> 2) * p.H + y) * p.W + x) * 4 + (m & 3)] = (int16_t)a;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == kIssuer) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

// Weight tiles from one layer of the reference's reorganised blob (addressed with the build's TM / TN):
// [mtile][K slice][plane hi | lo][canonical 128 rows x 32 K-bytes]; K byte k of slice sl is channel k % tn of the rounding group of
// chain step sl * (32 / tn) + k / tn (step = (group, tap)); step slots past the last step are zero.
__global__ void wprep_tc32_kernel(const int16_t *__restrict__ blob, unsigned char *__restrict__ dst, int ifm, int ofm, int ksize,
                                  int TM, int TN, int tn, int nsteps, int nslices, long long total)
{
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int k2 = ksize * ksize;
    int k = idx & 31;
    long long r = idx >> 5;
    int ml = r % kM; r /= kM;
    int sl = r % nslices;
    int mtile = r / nslices;
    int m = mtile * kM + ml;
    int step = sl * (32 / tn) + k / tn;
    int g = step / k2, tap = step - g * k2, c = g * tn + k % tn;
    int hi = 0, lo = 0;
    if (m < ofm && c < ifm && step < nsteps) {
        int w = blob[reorg_woff(m, c, tap, ifm, ofm, k2, TM, TN)];
        hi = (w >> 8) & 0xff;
        lo = w & 0xff;
    }
    unsigned char *tile = dst + ((size_t)mtile * nslices + sl) * kWBytes;
    tile[operand_off(ml, k)] = (unsigned char)hi;
    tile[kM * 32 + operand_off(ml, k)] = (unsigned char)lo;
}

template <int KS, int SO, int TNW>
void launch_one(const Tc32Params &p, dim3 grid, size_t smem, cudaStream_t st)
{
    // set on every launch (a microsecond): the attribute is per device and a host may drive several GPUs from several threads,
    // so a cached "already configured" flag would be a data race for nothing
    cudaFuncSetAttribute(conv_i16_tc32_kernel<KS, SO, TNW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    conv_i16_tc32_kernel<KS, SO, TNW><<<grid, kThreads, smem, st>>>(p);
}

// the setmaxnreg arithmetic above assumes the launch allocation ptxas chose; a build that allocates fewer registers per thread would
// leave the epilogue warps waiting for registers for ever, so the launcher refuses instead
template <int KS, int SO, int TNW>
bool pool_ok()
{
    cudaFuncAttributes fa{};
    if (cudaFuncGetAttributes(&fa, conv_i16_tc32_kernel<KS, SO, TNW>) != cudaSuccess) return false;
    return kEpiRegs - fa.numRegs <= fa.numRegs - kHelperRegs;
}

template <int KS, int TNW>
bool dispatch_so(const Tc32Params &p, int so, dim3 grid, size_t smem, cudaStream_t st)
{
    switch (so) {
#define Y2_TC32_CASE(S) case S: if (!pool_ok<KS, S, TNW>()) return false; launch_one<KS, S, TNW>(p, grid, smem, st); return true;
        Y2_TC32_CASE(8) Y2_TC32_CASE(9) Y2_TC32_CASE(10) Y2_TC32_CASE(11) Y2_TC32_CASE(12) Y2_TC32_CASE(13) Y2_TC32_CASE(14)
        Y2_TC32_CASE(15) Y2_TC32_CASE(16)
#undef Y2_TC32_CASE
    default: return false;
    }
}

template <int KS>
bool dispatch_tn(const Tc32Params &p, int so, int tn, dim3 grid, size_t smem, cudaStream_t st)
{
    switch (tn) {
    case 32: return dispatch_so<KS, 32>(p, so, grid, smem, st);
    case 16: return dispatch_so<KS, 16>(p, so, grid, smem, st);
    case 8: return dispatch_so<KS, 8>(p, so, grid, smem, st);
    default: return false;
    }
}

// staging geometry of one CTA (pt = 3 * tn consecutive pixels): false when the image is too wide for the shared-memory band
bool tc32_plan(int W, int ksize, int tn, Tc32Params &p, size_t &smem)
{
    const int pt = kR * tn;
    p.PW = W + ksize - 1;
    p.rows_max = (pt - 1) / W + 2 + (ksize - 1) + 1;
    if (p.rows_max > 30) return false;                  // rowinfo[] holds 32 staged rows
    p.ctab_cap = (p.rows_max - 1) * W;
    const size_t fixed = (size_t)kWRing * kWBytes + (size_t)kRing * 2 * kBBytes + 512 + kPT * 16 + 256 + (size_t)p.ctab_cap * 8;
    const size_t per_group = (size_t)p.rows_max * p.PW * 8;
    if (fixed + 2 * per_group * 8 > 200 * 1024) return false;   // a chunk holds whole rounding groups: eight C4 words for Tn = 32 (and
                                                                 // the chunk arithmetic of the narrower groups assumes at least that)
    int gs = (int)((200 * 1024 - fixed) / (2 * per_group));
    int sh = 3;
    while ((2 << sh) <= gs && (2 << sh) <= 16) ++sh;    // largest power of two <= min(gs, 16)
    p.gs_shift = sh;
    gs = 1 << sh;
    smem = fixed + 2 * per_group * gs + 1024;
    if (smem < 120 * 1024) smem = 120 * 1024;           // one CTA per SM: a CTA allocates all 512 TMEM columns
    return true;
}

inline int tc32_slices(int ifm, int ksize, int tn) { return ceil_div(ceil_div(ifm, tn) * ksize * ksize, 32 / tn); }

}  // namespace

size_t wprep_tc32_bytes(int ifm, int ofm, int ksize, int tn)
{
    return (size_t)ceil_div(ofm, kM) * tc32_slices(ifm, ksize, tn) * kWBytes;
}

void launch_wprep_tc32(const int16_t *blob, void *dst, int ifm, int ofm, int ksize, int TM, int TN, int tn, cudaStream_t st)
{
    const int nsteps = ceil_div(ifm, tn) * ksize * ksize, nslices = tc32_slices(ifm, ksize, tn);
    const long long total = (long long)ceil_div(ofm, kM) * nslices * kM * 32;
    wprep_tc32_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(blob, (unsigned char *)dst, ifm, ofm, ksize, TM, TN, tn, nsteps,
                                                                      nslices, total);
}

// K slices per CTA of a layer (the policy in capi.cu weighs them against the one-item-per-CTA prologue)
int conv_i16_tc32_slices(int ifm, int ksize, int tn) { return tc32_slices(ifm, ksize, tn); }

// the launcher's shape rules without launching (plan time)
bool conv_i16_tc32_eligible(int W, int ksize, int so, int tn)
{
    if ((ksize != 1 && ksize != 3) || so < 8 || so > 16 || (tn != 32 && tn != 16 && tn != 8)) return false;
    Tc32Params p{};
    size_t smem;
    return tc32_plan(W, ksize, tn, p, smem);
}

// Returns 1 when launched, -1 when the shape/shift is not eligible.  tn = the emulated build's rounding group (32, 16 or 8).
int launch_conv_i16_tc32(const ConvFastParams &cp, int ksize, int ifm, int tn, cudaStream_t st, const char **variant)
{
    if ((ksize != 1 && ksize != 3) || cp.so < 8 || cp.so > 16 || (tn != 32 && tn != 16 && tn != 8)) return -1;
    Tc32Params p{};
    p.in = (const uint2 *)cp.in; p.out = (int16_t *)cp.out; p.w = (const unsigned char *)cp.w; p.bias = (const int16_t *)cp.bias;
    p.B = cp.B; p.H = cp.H; p.W = cp.W; p.G = cp.G; p.OFM = cp.OFM;
    p.in_frame_stride = cp.in_frame_stride; p.out_frame_stride = cp.out_frame_stride;
    p.sb = cp.sb; p.leaky = cp.leaky;
    p.nsteps = ceil_div(ifm, tn) * ksize * ksize;
    p.nslices = tc32_slices(ifm, ksize, tn);
    const int pt = kR * tn;                             // pixels per CTA: three tiles of tn pixels x 32 / tn step slots
    size_t smem;
    if (!tc32_plan(cp.W, ksize, tn, p, smem)) return -1;
    dim3 grid((unsigned)(((long long)cp.B * cp.H * cp.W + pt - 1) / pt), ceil_div(cp.OFM, kM));
    const bool ok = ksize == 3 ? dispatch_tn<3>(p, cp.so, tn, grid, smem, st) : dispatch_tn<1>(p, cp.so, tn, grid, smem, st);
    if (!ok) return -1;
#ifdef Y2_TC32_PROFILE
    if (cudaDeviceSynchronize() == cudaSuccess && (long long)p.nslices * kR >= Y2_TC32_TL0 + 48 && grid.x > 1) {
        static long long tl[48 * 16];
        cudaMemcpyFromSymbol(tl, g_tc32_tl, sizeof(tl));
        fprintf(stderr, "tc32 timeline (CTA 1,0; tn %d, %d slices), cycles relative to the first event; per tile: slot free | built | w_full | go | issued | "
                        "mma_done | half read | released | computed | loader: w_empty | issued\n", tn, p.nslices);
        long long t0 = 0;
        for (int i = 0; i < 48 * 16; ++i) if (tl[i] && (!t0 || tl[i] < t0)) t0 = tl[i];
        for (int t = 0; t < 48; ++t) {
            fprintf(stderr, "  tile %3d (slice %3d r %d):", Y2_TC32_TL0 + t, (Y2_TC32_TL0 + t) / kR, (Y2_TC32_TL0 + t) % kR);
            for (int e = 0; e < 11; ++e) fprintf(stderr, " %6lld", tl[t * 16 + e] ? tl[t * 16 + e] - t0 : -1);
            fprintf(stderr, "\n");
        }
        memset(tl, 0, sizeof(tl));
        cudaMemcpyToSymbol(g_tc32_tl, tl, sizeof(tl));
    }
#endif
    if (variant)
        *variant = tn == 32 ? (ksize == 3 ? "conv_i16_tc32<3>" : "conv_i16_tc32<1>")
                 : tn == 16 ? (ksize == 3 ? "conv_i16_tc32<3,tn16>" : "conv_i16_tc32<1,tn16>")
                            : (ksize == 3 ? "conv_i16_tc32<3,tn8>" : "conv_i16_tc32<1,tn8>");
    return 1;
}

}  // namespace y2
