// INT16 convolution on the 5th-generation tensor cores (tcgen05 + TMEM), bit-exact.
//
// The north-star decomposition: each int16 operand is split into a signed-high and an
// unsigned-low int8 plane, four int8 products (hi*hi, hi*lo, lo*hi, lo*lo) run as tcgen05.mma
// kind::i8 and the partial sums are recombined on the CUDA cores with the reference's exact
// per-step rounding and 16-bit saturation (hls/core/core_compute.cpp:65-120).
//
// Because the reference rounds after EVERY (4-channel group x tap) step, one MMA K-slice (32) holds
// seven consecutive steps of the chain as a BLOCK-DIAGONAL activation operand:
//   A (smem, static)  : weights  [128 out-channels][K = 7 steps x 4 channels | rounding row | 0 0 0]
//   B (smem, built)   : for column (step s, pixel p) only rows 4s..4s+3 are non-zero (the pixel's 4
//                       channels at that step's tap), so D[m][(s,p)] = P_s[m][p] - one exact 4-MAC
//                       partial sum per column.  Row 28 carries the rounding constant `half`.
//   D (TMEM)          : three int32 accumulators per column: HH, M = HL+LH, LL.
// Epilogue thread = one TMEM lane = one output channel; it walks the steps of its pixels in order:
//   c = U + (LL >> 8); v = 256*HH + c + M; U = clamp(v & ~frac, 0, 65535 << (so-8))
// (5 CUDA-core instructions per step instead of 7, and only one of them on the half-rate FMA pipe).
// Valid for 8 <= so <= 22 like the scaled CUDA-core variant; everything else falls back.
#include "common.cuh"

namespace y2 {

namespace {

constexpr int kTcM = 128;        // output channels per CTA (TMEM lanes)
constexpr int kTcSteps = 7;      // chain steps per K-block
constexpr int kTcPx = 6;         // pixels per MMA (N = 8 step slots x 6 pixels = 48)
constexpr int kTcR = 9;          // pixel sub-tiles per CTA -> 54 pixels per CTA (3 per epilogue warp group)
constexpr int kTcPT = kTcPx * kTcR;
constexpr int kTcN = 48;
constexpr int kTcBufs = 3;       // TMEM accumulator buffers (3 x 3 x 48 = 432 of 512 columns)
constexpr int kTcBufCols = 3 * kTcN;
constexpr int kTcWRing = 4;      // weight K-block ring depth (8 KB each)
constexpr int kTcBRing = 6;      // B tile ring depth (2 x 1.5 KB each)
constexpr int kTcEpiWarps = 12;  // warps 0-11: epilogue (chain; group k = warp/4 owns TMEM buffer k), 12-13: B-tile builders, 14: MMA issuer, 15: weight loader
constexpr int kTcThreads = (kTcEpiWarps + 5) * 32;   // + 2 builders, 2 MMA issuers, 1 weight loader
constexpr int kTcABytes = 2 * kTcM * 32;   // hi + lo plane of one K-block
constexpr int kTcBBytes = kTcN * 32;       // one plane

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
__device__ __forceinline__ void umma_i8(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned idesc, unsigned accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld4(unsigned taddr, int (&r)[4])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr));
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
    return (2u << 4) | ((unsigned)a_signed << 7) | ((unsigned)b_signed << 10) | ((unsigned)(kTcN >> 3) << 17) |
           ((unsigned)(kTcM >> 4) << 24);
}
__host__ __device__ inline int operand_off(int row, int k) { return (((row >> 3) * 2 + (k >> 4)) * 8 + (row & 7)) * 16 + (k & 15); }

__device__ __forceinline__ void tmem_ld6(unsigned taddr, int (&r)[6])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r[4]), "=r"(r[5]) : "r"(taddr + 4));
}
__device__ __forceinline__ void reg_fence6(int (&r)[6])
{
    asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5])::"memory");
}
__device__ __forceinline__ void tmem_ld16(unsigned taddr, int *r)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8(unsigned taddr, int (&r)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
// ties the registers of an asynchronous tcgen05.ld to the point after tcgen05.wait::ld so that the
// compiler cannot schedule their first use above the wait
__device__ __forceinline__ void reg_fence8(int (&r)[8])
{
    asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])::"memory");
}
__device__ __forceinline__ void bar_sync_named(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }


struct TcParams {
    const uint2 *in;          // C4 input
    int16_t *out;             // C4 output (already offset to the first output group)
    const unsigned char *w;   // [mtile][kblock][hi 4 KB | lo 4 KB] canonical operand tiles
    const int16_t *bias;
    int B, H, W, G, OFM;
    long long in_frame_stride, out_frame_stride;  // elements
    int so, sb, leaky;
    int nkb;                  // K-blocks = ceil(G*K2/7)
    int PW, rows_max, gs_shift;  // staging: smem row pitch (pixels), band rows incl. halo + zero row, log2(groups per chunk)
};

// one exact step of the chain from the three int8-plane partial sums (2 FMA-pipe + 3 ALU-pipe instructions)
__device__ __forceinline__ int tc_step(int U, int hh, int mm, int ll, int nmask, int ubound)
{
    int t1 = mm * 256 + ll;                 // IMAD: 256*M + LL' (LL' already holds the rounding constant)
    int c = U + (t1 >> 8);                  // LEA.HI.SX32
    int v = hh * 256 + c;                   // IMAD
    return __vimin_s32_relu(v & nmask, ubound);   // LOP3 + VIMNMX.RELU
}

template <int KS>
__global__ void __launch_bounds__(kTcThreads, 1) conv_i16_tc_kernel(const TcParams p)
{
    constexpr int K2 = KS * KS;
    constexpr int PAD = KS / 2;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *sA = smem;                                    // kTcWRing x 8 KB
    unsigned char *sB = sA + kTcWRing * kTcABytes;               // kTcBRing x (hi 2 KB | lo 2 KB)
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(sB + kTcBRing * 2 * kTcBBytes);
    unsigned long long *w_full = bars, *w_empty = w_full + kTcWRing, *b_full = w_empty + kTcWRing, *b_empty = b_full + kTcBRing,
                       *t_full = b_empty + kTcBRing, *t_empty = t_full + kTcBufs;
    unsigned *tmem_slot = reinterpret_cast<unsigned *>(t_empty + kTcBufs);
    int *pxtab = reinterpret_cast<int *>(tmem_slot + 4);         // [64][4]: smem pixel offset for tap rows 0..2
    uint2 *sX = reinterpret_cast<uint2 *>(pxtab + kTcPT * 4);    // 2 chunks x GS groups x rows_max x PW pixels

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long npix = (long long)p.B * p.H * p.W;
    const long long pix0 = (long long)blockIdx.x * kTcPT;
    const int mtile = blockIdx.y;
    const int zero_slot = p.rows_max - 1;
    const int GS = 1 << p.gs_shift;
    const int chunk_px = GS * p.rows_max * p.PW;
    const long long row_first = pix0 / p.W;                      // global row (frame*H + y) of the first pixel

    if (tid == 0) {
        for (int i = 0; i < kTcWRing; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 2); }
        for (int i = 0; i < kTcBRing; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < kTcBufs; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 4); }   // t_full unused: b_empty[] doubles as "MMA of this slot done"
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kTcEpiWarps + 2) {   // the first MMA warp owns the TMEM allocation
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int q = tid; q < kTcPT; q += kTcThreads) {
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
    for (int i = tid; i < kTcBRing * 2 * kTcBBytes / 4; i += kTcThreads) reinterpret_cast<unsigned *>(sB)[i] = 0u;
    for (int i = tid; i < 2 * chunk_px; i += kTcThreads) sX[i] = make_uint2(0u, 0u);
    __syncthreads();
    {   // rounding row (k = 28) of every lo-plane B tile: b = 2^min(7, e) where a*b = 2^e is the constant to inject
        const int e = (p.so <= 15) ? p.so - 1 : p.so - 9;   // inject `half` into LL, or half/256 into M
        const int eb = e < 7 ? e : 7;
        for (int i = tid; i < kTcBRing * kTcN; i += kTcThreads) {
            int slot = i / kTcN, n = i - slot * kTcN;
            sB[(slot * 2 + 1) * kTcBBytes + operand_off(n, 28)] = (unsigned char)(1u << eb);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const unsigned tmem = *tmem_slot;
    const int nit = p.nkb * kTcR;

    if (warp == kTcEpiWarps + 4) {
        // ===== weight loader: one 8 KB bulk copy per K-block =====
        if (lane == 0) {
            const unsigned char *src = p.w + (size_t)mtile * p.nkb * kTcABytes;
            for (int b = 0; b < p.nkb; ++b) {
                const int s = b % kTcWRing;
                if (b >= kTcWRing) mbar_wait(&w_empty[s], ((b / kTcWRing) - 1) & 1);
                mbar_expect_tx(&w_full[s], kTcABytes);
                bulk_g2s(sA + s * kTcABytes, src + (size_t)b * kTcABytes, kTcABytes, &w_full[s]);
            }
        }
    } else if (warp == kTcEpiWarps + 2 || warp == kTcEpiWarps + 3) {
        // ===== two MMA issuer warps (even / odd iterations): the whole warp runs the (uniform) loop so that barrier addresses and operand
        // descriptors live in uniform registers; one elected lane issues the tcgen05 instructions =====
        const unsigned long long dA0 = smem_desc(sA), dB0 = smem_desc(sB);
        constexpr unsigned long long kAStep = kTcABytes >> 4, kAPlane = (kTcM * 32) >> 4;      // descriptor address units (16 B)
        constexpr unsigned long long kBStep = (2 * kTcBBytes) >> 4, kBPlane = kTcBBytes >> 4;
        unsigned elected;
        asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(elected));
        const int iw = warp - (kTcEpiWarps + 2);
        for (int it = iw; it < nit; it += 2) {
            const int b = it / kTcR, r = it - b * kTcR;
            const int slot = it % kTcBRing, tb = it % kTcBufs, wslot = b % kTcWRing;
            mbar_wait(&b_full[slot], (it / kTcBRing) & 1);
            if (it >= kTcBufs) mbar_wait(&t_empty[tb], ((it / kTcBufs) - 1) & 1);
            mbar_wait(&w_full[wslot], (b / kTcWRing) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;");
            if (elected) {
                const unsigned long long dAh = dA0 + wslot * kAStep, dAl = dAh + kAPlane;
                const unsigned long long dBh = dB0 + slot * kBStep, dBl = dBh + kBPlane;
                const unsigned d0 = tmem + tb * kTcBufCols;
                umma_i8(d0, dAh, dBh, idesc_i8(1, 1), 0);               // HH
                umma_i8(d0 + kTcN, dAh, dBl, idesc_i8(1, 0), 0);        // M  = hi*lo
                umma_i8(d0 + kTcN, dAl, dBh, idesc_i8(0, 1), 1);        //    + lo*hi
                umma_i8(d0 + 2 * kTcN, dAl, dBl, idesc_i8(0, 0), 0);    // LL
                umma_commit(&b_empty[slot]);                            // "MMA of this slot done": epilogue + builders wait on it
                if (r >= kTcR - 2) umma_commit(&w_empty[wslot]);        // last tile of each issuer in this K-block
            }
            __syncwarp();
        }
    } else if (warp >= kTcEpiWarps) {
        // ===== two builder warps: stage activations (cp.async) and write the block-diagonal B tiles =====
        const int bw = warp - kTcEpiWarps;          // builder 0 takes the even tiles of a K-block, builder 1 the odd ones
        const int bt = bw * 32 + lane;              // 0..63
        const int nrows = p.rows_max - 1;           // staged band rows (the last slot is the all-zero row)
        const int nchunks = (p.G + GS - 1) >> p.gs_shift;
        auto stage_chunk = [&](int c) {
            uint2 *dst = sX + (c & 1) * chunk_px;
            const int g0 = c << p.gs_shift, ng = min(GS, p.G - g0);
            const int per_group = nrows * p.W;
            for (int idx = bt; idx < ng * per_group; idx += 64) {
                int gg = idx / per_group, rem = idx - gg * per_group;
                int s = rem / p.W, x = rem - s * p.W;
                long long Rr = row_first - PAD + s;
                if (Rr >= 0 && Rr < (long long)p.B * p.H) {
                    long long ff = Rr / p.H;
                    int yy = (int)(Rr - ff * p.H);
                    const uint2 *src = p.in + ff * (p.in_frame_stride >> 2) + ((long long)(g0 + gg) * p.H + yy) * p.W + x;
                    cp_async8(dst + (gg * p.rows_max + s) * p.PW + PAD + x, src);
                }
            }
            asm volatile("cp.async.commit_group;");
        };
        // entries of one tile: (step slot s, pixel pp) for s < 7, pp < 6 -> 42 words per plane; this lane owns e0 and e1
        const int e0 = lane, e1 = 32 + lane;
        const int s0 = e0 / kTcPx, p0 = e0 - s0 * kTcPx;
        const int s1 = e1 / kTcPx, p1 = e1 - s1 * kTcPx;
        const bool has1 = e1 < kTcSteps * kTcPx;
        const int off0 = operand_off(s0 * kTcPx + p0, 4 * s0), off1 = has1 ? operand_off(s1 * kTcPx + p1, 4 * s1) : 0;
        stage_chunk(0);
        int staged = 0, ready = -1;
        for (int b = 0; b < p.nkb; ++b) {
            const int c_first = (min(p.G - 1, (b * kTcSteps) / K2)) >> p.gs_shift;
            const int c_need = (min(p.G - 1, (b * kTcSteps + kTcSteps - 1) / K2)) >> p.gs_shift;
            if (staged + 1 < nchunks && staged <= c_first) { stage_chunk(staged + 1); ++staged; }
            if (ready < c_need) {
                if (staged > c_need) asm volatile("cp.async.wait_group 1;" ::: "memory");
                else asm volatile("cp.async.wait_group 0;" ::: "memory");
                bar_sync_named(1, 64);   // both builder warps see each other's copies
                ready = c_need;
            }
            // per-K-block constants of this lane's two entries
            const int sg0 = b * kTcSteps + s0, sg1 = b * kTcSteps + s1;
            const bool live0 = sg0 < p.G * K2, live1 = has1 && sg1 < p.G * K2;
            const int g0 = live0 ? sg0 / K2 : 0, t0 = sg0 - g0 * K2, g1 = live1 ? sg1 / K2 : 0, t1 = sg1 - g1 * K2;
            const int ti0 = t0 / KS, tj0 = t0 - ti0 * KS, ti1 = t1 / KS, tj1 = t1 - ti1 * KS;
            const uint2 *xs0 = sX + ((g0 >> p.gs_shift) & 1) * chunk_px + (g0 & (GS - 1)) * p.rows_max * p.PW + tj0;
            const uint2 *xs1 = sX + ((g1 >> p.gs_shift) & 1) * chunk_px + (g1 & (GS - 1)) * p.rows_max * p.PW + tj1;
            for (int r = 0; r < kTcR; ++r) {
                const int it = b * kTcR + r;
                if ((it & 1) != bw) continue;        // the two builder warps alternate tiles
                const int slot = it % kTcBRing;
                if (it >= kTcBRing) mbar_wait(&b_empty[slot], ((it / kTcBRing) - 1) & 1);
                unsigned char *bh = sB + (slot * 2) * kTcBBytes;
                unsigned hi0 = 0, lo0 = 0, hi1 = 0, lo1 = 0;
                if (live0) {
                    const uint2 x = xs0[pxtab[(r * kTcPx + p0) * 4 + ti0]];
                    hi0 = __byte_perm(x.x, x.y, 0x7531);
                    lo0 = __byte_perm(x.x, x.y, 0x6420);
                }
                if (live1) {
                    const uint2 x = xs1[pxtab[(r * kTcPx + p1) * 4 + ti1]];
                    hi1 = __byte_perm(x.x, x.y, 0x7531);
                    lo1 = __byte_perm(x.x, x.y, 0x6420);
                }
                *reinterpret_cast<unsigned *>(bh + off0) = hi0;
                *reinterpret_cast<unsigned *>(bh + kTcBBytes + off0) = lo0;
                if (has1) {
                    *reinterpret_cast<unsigned *>(bh + off1) = hi1;
                    *reinterpret_cast<unsigned *>(bh + kTcBBytes + off1) = lo1;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&b_full[slot]);
            }
        }
    } else {
        // ===== epilogue warps: thread = one output channel (TMEM lane); warp group kg = warp/4 takes the
        // iterations that land in TMEM buffer kg; 6 pixels x 7 steps per iteration =====
        const int q4 = warp & 3, kg = warp >> 2;     // TMEM lane quadrant, warp group = TMEM buffer
        const int m = mtile * kTcM + q4 * 32 + lane;
        const int k2 = p.so - 8;
        const int nmask = ~((1 << k2) - 1);
        const int ubound = 65535 << k2;
        int U[kTcR / 3][kTcPx];
        {
            long long bv = (m < p.OFM) ? (long long)p.bias[m] : 0;
            long long base = round_shift64(bv, p.sb);
            const long long rb = ((1LL << 25) >> k2) + 2;
            long long boff = base + 32768;
            if (boff > 65535 + rb) boff = 65535 + rb;
            if (boff < -rb) boff = -rb;
            const int init = (int)(boff * (1LL << k2));
#pragma unroll
            for (int r = 0; r < kTcR / 3; ++r)
#pragma unroll
                for (int j = 0; j < kTcPx; ++j) U[r][j] = init;
        }
        const unsigned base = tmem + ((unsigned)(q4 * 32) << 16) + kg * kTcBufCols;
        for (int b = 0; b < p.nkb; ++b) {
#pragma unroll
            for (int rr = 0; rr < kTcR / 3; ++rr) {
                // iteration it = b*9 + 3*rr + kg uses TMEM buffer it % 3 = kg; its use count is b*3 + rr
                const int it = b * kTcR + 3 * rr + kg;
                mbar_wait(&b_empty[it % kTcBRing], (it / kTcBRing) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;");
                // column n = step*6 + pixel: three chunks of 16 columns per partial-sum plane
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    int hh[16], mm[16], ll[16];
                    tmem_ld16(base + c * 16, hh);
                    tmem_ld16(base + kTcN + c * 16, mm);
                    tmem_ld16(base + 2 * kTcN + c * 16, ll);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (c == 2) {   // everything is in registers: hand the TMEM buffer back to the MMA issuer
                        asm volatile("tcgen05.fence::before_thread_sync;");
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&t_empty[kg]);
                    }
#pragma unroll
                    for (int n = 0; n < 16; ++n) {
                        const int col = c * 16 + n;
                        if (col < kTcSteps * kTcPx) {
                            const int j = col % kTcPx;
                            U[rr][j] = tc_step(U[rr][j], hh[n], mm[n], ll[n], nmask, ubound);
                        }
                    }
                }
            }
        }
        if (m < p.OFM) {
#pragma unroll
            for (int rr = 0; rr < kTcR / 3; ++rr)
#pragma unroll
                for (int j = 0; j < kTcPx; ++j) {
                    const long long gp = pix0 + (3 * rr + kg) * kTcPx + j;
                    if (gp >= npix) continue;
                    const long long grow = gp / p.W;
                    const int x = (int)(gp - grow * p.W);
                    const long long f = grow / p.H;
                    const int y = (int)(grow - f * p.H);
                    int a = (U[rr][j] >> k2) - 32768;
                    if (p.leaky && a < 0) a = a / 10;
                    p.out[f * p.out_frame_stride + (((long long)(m >> 2) * p.H + y) * p.W + x) * 4 + (m & 3)] = (int16_t)a;
                }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == kTcEpiWarps + 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

// Weight tiles for the tensor-core path from one layer of the reference's reorganised blob.
// Output: [mtile][kblock][plane][canonical 128 x 32] bytes; row 28 = the rounding constant's weight-side factor.
__global__ void wprep_tc_kernel(const int16_t *__restrict__ blob, unsigned char *__restrict__ dst, int ifm, int ofm, int ksize,
                                int TM, int TN, int nkb, int so, long long total)
{
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int k2 = ksize * ksize;
    int k = idx & 31;
    long long r = idx >> 5;
    int ml = r % kTcM; r /= kTcM;
    int b = r % nkb;
    int mtile = r / nkb;
    int m = mtile * kTcM + ml;
    int hi = 0, lo = 0;
    if (k < 28) {
        int sigma = b * kTcSteps + (k >> 2), t = k & 3;
        int G = (ifm + 3) / 4;
        if (sigma < G * k2) {
            int g = sigma / k2, tap = sigma - g * k2, c = g * 4 + t;
            if (m < ofm && c < ifm) {
                int w = blob[reorg_woff(m, c, tap, ifm, ofm, k2, TM, TN)];
                hi = (w >> 8) & 0xff;
                lo = w & 0xff;
            }
        }
    } else if (k == 28) {
        const int e = (so <= 15) ? so - 1 : so - 9;
        const int eb = e < 7 ? e : 7, ea = e - eb;          // a * b = 2^e, b = 2^eb <= 128, a = 2^ea <= 64
        if (so <= 15) lo = 1 << ea; else hi = 1 << ea;
    }
    unsigned char *tile = dst + ((size_t)mtile * nkb + b) * kTcABytes;
    tile[operand_off(ml, k)] = (unsigned char)hi;
    tile[kTcM * 32 + operand_off(ml, k)] = (unsigned char)lo;
}

}  // namespace


size_t wprep_tc_bytes(int ifm, int ofm, int ksize)
{
    const int nkb = ceil_div(ceil_div(ifm, 4) * ksize * ksize, kTcSteps);
    return (size_t)ceil_div(ofm, kTcM) * nkb * kTcABytes;
}

void launch_wprep_tc(const int16_t *blob, void *dst, int ifm, int ofm, int ksize, int TM, int TN, int so, cudaStream_t st)
{
    const int nkb = ceil_div(ceil_div(ifm, 4) * ksize * ksize, kTcSteps);
    const long long total = (long long)ceil_div(ofm, kTcM) * nkb * kTcM * 32;
    wprep_tc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(blob, (unsigned char *)dst, ifm, ofm, ksize, TM, TN, nkb, so, total);
}

// Returns 1 when launched, -1 when the shape/shift is not eligible for the tensor-core path.
int launch_conv_i16_tc(const ConvFastParams &cp, int ksize, cudaStream_t st, const char **variant)
{
    if ((ksize != 1 && ksize != 3) || cp.so < 8 || cp.so > 22) return -1;
    TcParams p{};
    p.in = (const uint2 *)cp.in; p.out = (int16_t *)cp.out; p.w = (const unsigned char *)cp.w; p.bias = (const int16_t *)cp.bias;
    p.B = cp.B; p.H = cp.H; p.W = cp.W; p.G = cp.G; p.OFM = cp.OFM;
    p.in_frame_stride = cp.in_frame_stride; p.out_frame_stride = cp.out_frame_stride;
    p.so = cp.so; p.sb = cp.sb; p.leaky = cp.leaky;
    p.nkb = ceil_div(cp.G * ksize * ksize, kTcSteps);
    p.PW = cp.W + ksize - 1;
    // 64 consecutive pixels touch at most ceil(63/W)+1 rows; + halo rows + the zero row
    p.rows_max = (kTcPT - 1) / cp.W + 2 + (ksize - 1) + 1;
    const size_t fixed = (size_t)kTcWRing * kTcABytes + (size_t)kTcBRing * 2 * kTcBBytes + 512 + kTcPT * 16;
    const size_t per_group = (size_t)p.rows_max * p.PW * 8;
    int gs = (int)((200 * 1024 - fixed) / (2 * per_group));
    if (gs < 1) return -1;
    int sh = 0;
    while ((2 << sh) <= gs && (2 << sh) <= 16) ++sh;   // largest power of two <= min(gs, 16)
    p.gs_shift = sh;
    gs = 1 << sh;
    size_t smem = fixed + 2 * per_group * gs + 1024;
    if (smem < 120 * 1024) smem = 120 * 1024;          // one CTA per SM: a CTA allocates all 512 TMEM columns
    dim3 grid((unsigned)(((long long)cp.B * cp.H * cp.W + kTcPT - 1) / kTcPT), ceil_div(cp.OFM, kTcM));
    if (ksize == 3) {
        cudaFuncSetAttribute(conv_i16_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        conv_i16_tc_kernel<3><<<grid, kTcThreads, smem, st>>>(p);
        if (variant) *variant = "conv_i16_tc<3>";
    } else {
        cudaFuncSetAttribute(conv_i16_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        conv_i16_tc_kernel<1><<<grid, kTcThreads, smem, st>>>(p);
        if (variant) *variant = "conv_i16_tc<1>";
    }
    return 1;
}

}  // namespace y2
