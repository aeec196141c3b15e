// INT16 convolution on the 5th-generation tensor cores (tcgen05 + TMEM), bit-exact: a persistent, TMA-staged kernel.
//
// Arithmetic (hls/core/core_compute.cpp:65-120): every (4-channel group x tap) step of the reference's
// chain is  acc = clamp16(acc + ((P + half) >> so)),  P = sum_{t<4} w_t * x_t.  Each int16 operand is
// split into a signed-high and an unsigned-low int8 plane; four tcgen05.mma kind::i8 products give
//   HH = sum wh*xh,  M = sum (wh*xl + wl*xh),  LL = sum wl*xl (+ the rounding constant through K row 28)
// with P + half = 65536*HH + 256*M + LL.  One MMA K-slice (32) holds seven consecutive steps as a
// block-diagonal activation operand, so every TMEM column is one exact 4-MAC partial sum.
//
// Structure (DESIGN.md section 4 has the measurements behind every choice):
//  * PERSISTENT: one CTA per SM walks work items (48 consecutive pixels x 128 output channels; HALF mode: 96 pixels x 64
//    channels on the two halves of every lane quadrant) in a static round-robin; TMEM, barriers and rings are set up once and
//    every pipeline counter runs on across items.
//  * Roles, 24 warps: 16 epilogue warps in four groups (thread = one TMEM lane = one output channel, 96 registers), three builder
//    warps (block-diagonal hi / lo operand tiles from the staged activation runs, a tile pair per barrier), one TMA producer warp
//    (tensor-map boxes of the activation runs two chunks ahead + the per-item pixel tables), four MMA issuer warps (one elected
//    lane each; they also move each 8 KB weight K-block shared memory -> registers -> TMEM, where it is the A operand:
//    tcgen05.mma [d],[a],b-desc - an N = 32 MMA with A in shared memory re-reads 4 KB per instruction and is smem bound).
//  * N = 32 (8 step slots x 4 pixels) and FIVE 96-column accumulator buffers (HH | M | LL): an epilogue warp reads a whole tile
//    and releases the buffer before it computes, so the MMAs of a buffer's next tile overlap the epilogue work.
//  * The exact step for 8 <= so <= 16 (so is a template parameter; LEA.HI needs an immediate):
//       t = 256*M + LL            IMAD   (fits: |256 M| < 2^27, 0 <= LL < 2^19)
//       d = HH * 2^(16-so) + (t >> so)   IMAD / LEA.HI.SX32   (65536*HH is a multiple of 2^so: no rounding involved)
//       acc = max(min(acc + d, 65535), 0)   VIADDMNMX.RELU   (acc is kept as acc+32768)
//    and for 17 <= so <= 22:  c = 256*HH + M; c += LL >> 8; a = acc + (c >> (so-8)); clamp.
//  * NO-SATURATION FAST PATH.  65536*HH is a multiple of 2^so for so <= 16, so HH enters the chain linearly; if no step
//    of a K-block can saturate, its seven steps collapse to   acc += 2^(16-so) * sum_s HH_s + sum_s ((256 M_s + LL_s) >> so):
//    TWO instructions per step (IMAD, LEA.HI.SX32), and the HH plane is not read at all - the builders add a DENSE column per pixel
//    (all seven steps' hi bytes) in the four spare columns 28..31 of the activation tile, so the HH MMA delivers sum_s HH_s there.
//    "Cannot saturate" is decided per (output channel, pixel, K-block) from the accumulator itself:
//        Dsum <= acc+32768 <= 65535 - Dsum,   Dsum = ((sum_{28 weights} |w|) * xmax >> so) + 8  >=  sum_s |rs(P_s, so)|
//    with xmax = the largest |activation| of the layer input (an atomicMax the producing kernel leaves behind; 32768 when the
//    producer is unknown) and the weight norm from a table wprep writes next to the tiles.  One vote per warp and K-block (the
//    group's three tiles); a warp that fails takes the EXACT step (HH plane read in a second pass) for that K-block, so the
//    result is the reference's bits unconditionally.  profiles/microbench/tc_epilogue_rates.cu: 2.0 cycles per warp-step per
//    SMSP for the fast step against 4.5 for the exact one.
//  * The barrier ring (go / mma_done / rd_done, 12 slots = one K-block) is phase-safe: no barrier can complete a second phase
//    before every waiter has tested the first (comments at the declarations).  -DY2_TC2_PROFILE builds deadlock-detecting waits
//    and a per-role / per-tile cycle profile; -DY2_TC2_GRID=n runs n persistent CTAs (many items per CTA on small test shapes).
#include "common.cuh"
#include <cstdio>
#include <cuda.h>            // CUtensorMap + enums only: the encoder is fetched through cudaGetDriverEntryPoint (no libcuda link)
#include <cudaTypedefs.h>

namespace y2 {

namespace {

constexpr int kM = 128;             // output channels per CTA = TMEM lanes
constexpr int kSteps = 7;           // chain steps per K-block: K = 32 = 7 x 4 channels | rounding row | 3 zero rows
constexpr int kPx = 4;              // pixels per tile
constexpr int kN = 32;              // MMA N = 8 step slots x 4 pixels (slot 7 unused)
constexpr int kR = 12;              // tiles per K-block -> 48 pixels per pixel set
constexpr int kPT = kPx * kR;
#ifndef Y2_TC2_BUFS
#define Y2_TC2_BUFS 5
#endif
constexpr int kBufs = Y2_TC2_BUFS;  // TMEM accumulator buffers: HH | M | LL, 32 columns each (build option only for the in-flight scaling experiment)
constexpr int kPlaneCols = kN;
constexpr int kBufCols = 3 * kN;
constexpr int kACol = kBufs * kBufCols;   // 480: two weight slots of 16 columns (hi plane 8 | lo plane 8)
constexpr int kBRing = 12;          // activation tile ring (hi 1 KB | lo 1 KB) = the kR tiles of one K-block (see the go[] comment in the kernel)
#ifndef Y2_TC2_WRING
#define Y2_TC2_WRING 3
#endif
constexpr int kWRing = Y2_TC2_WRING;   // weight K-block ring in shared memory
constexpr int kGroups = 4;          // epilogue warpgroups; group kg takes the tiles r = kg (mod 4) of every K-block.  Measured (round 2,
                                    //   profiles/README.md): a generic tile loop (any group count, ring slot computed at run time) runs 17 %
                                    //   slower than this unrolled form with compile-time barrier addresses, and 5 / 6 groups then recover 6 %
constexpr int kTPG = kR / kGroups;  // tiles per group per K-block
constexpr int kEpiWarps = 4 * kGroups;   // warps 0..: group kg = warp/4, TMEM lane quadrant = warp%4
constexpr int kBuilders = 3;        // next warpgroup: builder bw takes the tile PAIRS (2j, 2j+1) with j = bw (mod 3); its fourth warp is the TMA producer
constexpr int kIssuer = kEpiWarps + 4;   // last warpgroup: issuer iw takes the tiles r = iw (mod 4).  One warp issues one MMA per ~29 cycles
constexpr int kIssuers = 4;         //   (profiles/microbench/umma_issue.cu) and pays ~100 cycles per mbarrier wait
constexpr int kThreads = (kEpiWarps + 8) * 32;   // 768 threads, 80 registers at launch; setmaxnreg then moves registers
constexpr int kEpiRegs = 96;        //   from the helper warps (48) to the epilogue warps (96).  (104 / 40 adds up on paper - 61440 registers at
constexpr int kHelperRegs = 48;     //   launch, 4096 free, 10240 released, 12288 requested - but setmaxnreg.inc never returns: measured, the kernel hangs)
static_assert(kBRing == kR && kR % kIssuers == 0 && kR % kGroups == 0 && (kR / 2) % kBuilders == 0, "ring = one K-block; see go[]");
constexpr int kWBytes = kM * 64;    // one K-block of weights: [row][hi 32 B | lo 32 B], 16-byte chunks XOR-swizzled by (row>>1)&3
constexpr int kBBytes = kN * 32;    // one plane of one activation tile

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(void *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
#ifdef Y2_TC2_PROFILE
__device__ int g_tc2_dbg[32 * 4];
__device__ int g_tc2_abort;
// debug build: a wait that does not complete within ~0.1 s records (line, barrier offset, parity) for its warp and gives up,
// so that a protocol deadlock ends the kernel and can be read back instead of hanging the GPU
__device__ __forceinline__ void mbar_wait_dbg(void *bar, unsigned parity, int line)
{
    const long long t0 = clock64();
    for (;;) {
        unsigned ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return;
        const long long waited = clock64() - t0;
        if (waited > 200000000LL || *(volatile int *)&g_tc2_abort) {
            // record only waits that were really stuck (not the unsatisfied waits a released warp runs into after the abort)
            if (waited > 20000000LL && blockIdx.x == 1 && blockIdx.y == 0 && g_tc2_dbg[(threadIdx.x >> 5) * 4] == 0) {
                int *d = g_tc2_dbg + (threadIdx.x >> 5) * 4;
                d[0] = line; d[1] = (int)(smem_u32(bar)); d[2] = (int)parity;
            }
            g_tc2_abort = 1;
            return;
        }
    }
}
#define mbar_wait(bar, parity) mbar_wait_dbg(bar, parity, __LINE__)
__device__ __forceinline__ void mbar_wait_fast(void *bar, unsigned parity)
#else
__device__ __forceinline__ void mbar_wait(void *bar, unsigned parity)
#endif
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
// wait flavours for the hand-off experiments: 0 = try_wait (hardware-suspended probe, default time limit), 1 = test_wait spin,
// 2 = try_wait with an explicit suspend-time hint of Y2_TC2_HINT_NS nanoseconds
#ifndef Y2_TC2_HINT_NS
#define Y2_TC2_HINT_NS 64
#endif
template <int MODE>
__device__ __forceinline__ void mbar_wait_m(void *bar, unsigned parity)
{
    if constexpr (MODE == 0) {
        mbar_wait(bar, parity);
    } else if constexpr (MODE == 1) {
#ifdef Y2_TC2_PROFILE
        mbar_wait(bar, parity);
#else
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "WAIT_LOOP:\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
            "@p bra DONE;\n\t"
            "bra WAIT_LOOP;\n\t"
            "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
            : "memory");
#endif
    } else {
#ifdef Y2_TC2_PROFILE
        mbar_wait(bar, parity);
#else
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "WAIT_LOOP:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
            "@p bra DONE;\n\t"
            "bra WAIT_LOOP;\n\t"
            "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity), "r"(Y2_TC2_HINT_NS)
            : "memory");
#endif
    }
}
#ifndef Y2_TC2_BLOCKCHECK
#define Y2_TC2_BLOCKCHECK 1
#endif
#ifndef Y2_TC2_LD32
#define Y2_TC2_LD32 1
#endif
#ifndef Y2_TC2_WAIT_HELPER
#define Y2_TC2_WAIT_HELPER 0
#endif
#ifndef Y2_TC2_WAIT_EPI
#define Y2_TC2_WAIT_EPI 0
#endif
__device__ __forceinline__ void mbar_arrive(void *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// arrive from the lanes where `on` is set, as a predicated instruction (no branch, no convergence barrier around it)
__device__ __forceinline__ void mbar_arrive_if(void *bar, bool on)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %1, 0;\n\t@p mbarrier.arrive.shared::cta.b64 _, [%0];\n\t}" ::"r"(smem_u32(bar)), "r"((int)on) : "memory");
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
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_i8_ts(unsigned tmem_d, unsigned tmem_a, unsigned long long db, unsigned idesc, unsigned accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld4(unsigned taddr, int *r)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
}
// the 28 live columns of one plane (x16 + x8 + x4) from ONE address operand: one R2UR per plane instead of one per load
__device__ __forceinline__ void tmem_ld28(unsigned taddr, int *r)
{
    asm volatile(
        "{\n\t.reg .b32 a1, a2;\n\t"
        "add.u32 a1, %28, 16;\n\t"
        "add.u32 a2, %28, 24;\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%28];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%16,%17,%18,%19,%20,%21,%22,%23}, [a1];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%24,%25,%26,%27}, [a2];\n\t}"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(unsigned taddr, int *r)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
                   "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
                   "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_st8(unsigned taddr, const uint4 &a, const uint4 &b)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(a.x), "r"(a.y), "r"(a.z),
                 "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
                 : "memory");
}
// ties the registers of an asynchronous tcgen05.ld to the point after tcgen05.wait::ld
__device__ __forceinline__ void reg_fence16(int *r)
{
    asm volatile(""
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                   "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])::"memory");
}
__device__ __forceinline__ void reg_fence12(int *r)
{
    asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                   "+r"(r[10]), "+r"(r[11])::"memory");
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
__host__ __device__ constexpr unsigned idesc_i8(int a_signed, int b_signed, unsigned m_rows = kM)
{
    return (2u << 4) | ((unsigned)a_signed << 7) | ((unsigned)b_signed << 10) | ((unsigned)(kN >> 3) << 17) | ((m_rows >> 4) << 24);
}
__host__ __device__ inline int operand_off(int row, int k) { return (((row >> 3) * 2 + (k >> 4)) * 8 + (row & 7)) * 16 + (k & 15); }

#ifdef Y2_TC2_PROFILE
__device__ long long g_tc2_prof[32 * 8];
#define PROF_DECL long long prof_[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long pt_ = clock64(); const long long pstart_ = pt_;
#define PROF_ADD(i) do { long long n_ = clock64(); prof_[i] += n_ - pt_; pt_ = n_; } while (0)
#define PROF_END do { prof_[7] = clock64() - pstart_; if (blockIdx.x == 1 && blockIdx.y == 0 && lane == 0) for (int i_ = 0; i_ < 8; ++i_) g_tc2_prof[warp * 8 + i_] = prof_[i_]; } while (0)
// per-tile timeline of 48 consecutive tiles of CTA (1,0): [tile - 600][event]: 0 builder starts (slot's previous tile read out), 1 built,
// 2 issuer passed go, 3 issued + committed, 4..7 epilogue warp q passed mma_done, 8..11 read out + released, 12..15 computed
__device__ long long g_tc2_tl[48 * 16];
#ifndef Y2_TC2_TL0
#define Y2_TC2_TL0 600      // first tile of the 48-tile timeline window (global tile counter of CTA 1: -DY2_TC2_TL0=<nkb*12 - 24> straddles the first item boundary)
#endif
#define PROF_TL(tile, ev) do { if (blockIdx.x == 1 && blockIdx.y == 0 && lane == 0 && (tile) >= Y2_TC2_TL0 && (tile) < Y2_TC2_TL0 + 48) g_tc2_tl[((tile) - Y2_TC2_TL0) * 16 + (ev)] = clock64(); } while (0)
#else
#define PROF_TL(tile, ev)
#define PROF_DECL
#define PROF_ADD(i)
#define PROF_END
#endif

struct Tc2Params {
    int16_t *out;             // C4 output (already offset to the first output group)
    const unsigned char *w;   // [mtile][kblock][128 rows][64 B]
    const unsigned *wnorm;    // [mtile][kblock][128 rows]: sum of |w| over the 28 weights of the row's K-block (fast-path bound)
    const int16_t *bias;
    const int *xmax_in;       // largest |activation| of the input tensor (device scalar), or NULL = unknown (32768)
    int *xmax_out;            // atomicMax target for the largest |output| of this layer, or NULL
    unsigned long long *stats;   // [0] += warp-tiles through the fast path, [1] += warp-tiles through the exact path (or NULL)
    int force_exact;          // tests: never take the fast path
    int B, H, W, G, OFM;
    long long out_frame_stride;  // elements
    int sb, leaky;
    int nkb;                  // K-blocks per work item = ceil(G*K2/7)
    int HW;                   // H * W
    int npt, mtiles, nitems;  // pixel tiles (48 consecutive pixels of the frame-flattened image), 128-channel tiles; work items = npt x mtiles,
                              // item = pixel tile * mtiles + channel tile: the CTAs running at one time share their activation runs in L2
    int gs_shift, nchunks;    // activation staging: log2(groups per chunk), chunks per work item
    int boxlen, nbox, seg_px; // TMA boxes per (group, frame segment): nbox boxes of boxlen pixels; seg_px = nbox * boxlen
};

// one 2-D tensor-map box (boxlen x 1 elements of 8 bytes) global -> shared, completion on an mbarrier
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *tmap, int c0, int c1, void *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
                 "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}

constexpr int kProducer = kEpiWarps + kBuilders;   // the fourth warp of the builder warpgroup: TMA producer of the activation runs
constexpr int kTabSlots = 4;                        // per-item tables live in a ring: the producer runs at most two items ahead of the epilogue
constexpr int kNumBars = 2 * kWRing + 4 + 3 * kBRing + 4;
// HALF mode (64 output channels per item): the 128 TMEM lanes hold 64 channels x TWO pixel sets.  A tcgen05.mma with M = 64 puts
// row r in lane (r/16)*32 + r%16, and its D (and TMEM A) address may carry a lane offset of 16 (profiles/microbench/umma_m64_layout.cu,
// profiles/r2_umma_m64_layout.jsonl), so two M = 64 products - same weights, the activation tiles of pixel set 0 / set 1 - fill all
// 128 lanes and every epilogue warp works with 32 live lanes.  An item is then 96 consecutive pixels x 64 channels; the 64-channel
// layers (32->64 3x3 @208, 128->64 and 512->64 1x1) would otherwise run half-empty 128-row tiles or stay on the CUDA cores.
template <bool HALF> struct Tc2Shape {
    static constexpr int kSets = HALF ? 2 : 1;              // pixel sets per item
    static constexpr int kPTI = kPT * kSets;                // pixels per item
    static constexpr int kRows = HALF ? 64 : 128;           // output channels per item
    static constexpr int kSlotBytes = kSets * 2 * kBBytes;  // one ring slot: [set][hi 1 KB | lo 1 KB]
    static constexpr int kFixedBytes = ((kWRing * kWBytes + kBRing * kSlotBytes + kNumBars * 8 + 16 + kTabSlots * kPTI * (16 + 8)) + 127) & ~127;
};

// PERSISTENT kernel (round 2): one CTA per SM walks the work items  item = blockIdx.x + n * gridDim.x  (static round-robin;
// item = (48-pixel tile, 128-channel tile)).  Tensor memory, the barriers, the operand ring and the rounding row are set up once;
// every pipeline counter (K-block parity, TMEM buffer, weight ring, activation chunk) runs on ACROSS items, so the producer,
// the builders and the issuers are already working on item n+1 while the epilogue warps finish item n - the ~14.6 k-cycle
// per-CTA prologue + pipeline fill of the one-item-per-CTA form is paid once per SM instead of once per item.
//
// Activation staging is TENSOR-MAP TMA (cp.async.bulk.tensor.2d, SASS UTMALDG): the C4 tensor is described as
// [frame][G*H*W] 8-byte pixels; for every (channel group, frame segment) ONE contiguous run of  48 + 2W + 2  pixels
// (frame-flat pixels pix0 - W - 1 ... pix0 + 47 + W + 1, everything the 3x3 taps of the item's 48 pixels touch) lands in
// shared memory, so tap (i, j) of pixel q is run[q + i*W + j].  Image borders are a 6-bit validity mask per pixel
// (3 tap rows | 3 tap columns) applied by the builders; rows of a neighbouring group that the run drags in are masked the same
// way, and coordinates outside the tensor are zero-filled by the TMA unit.  An item that straddles a frame boundary stages a
// second run for the next frame.  (A 4-D [frame][G][H][W] map with per-row boxes is not legal for the 13- and 19-wide layers:
// their 104- / 152-byte rows break the 16-byte global-stride rule, SURVEY.md appendix A.)
template <int KS, int SO, bool HALF>
__global__ void __launch_bounds__(kThreads, 1) conv_i16_tc2_kernel(const Tc2Params p, const __grid_constant__ CUtensorMap tmap)
{
    constexpr int K2 = KS * KS;
    constexpr int PAD = KS / 2;
    constexpr bool kFast = SO <= 16;   // 65536 * HH is a multiple of 2^so: HH enters the chain linearly
    constexpr bool kLd32 = SO <= 14 && Y2_TC2_LD32 != 0;   // (so = 15 spills with it)
    using Sh = Tc2Shape<HALF>;
    constexpr int kSets = Sh::kSets, kPTI = Sh::kPTI, kSlotBytes = Sh::kSlotBytes, kFixedBytes = Sh::kFixedBytes;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *sW = smem;                                    // kWRing x 8 KB
    unsigned char *sB = sW + kWRing * kWBytes;                   // kBRing x (hi 1 KB | lo 1 KB); a builder pair = two consecutive slots
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(sB + kBRing * kSlotBytes);
    unsigned long long *w_full = bars, *w_empty = w_full + kWRing, *a_full = w_empty + kWRing, *a_empty = a_full + 2,
                       *go = a_empty + 2, *mma_done = go + kBRing, *rd_done = mma_done + kBRing, *x_full = rd_done + kBRing,
                       *x_empty = x_full + 2;
    // go[r]: tile r of the current K-block may be issued = its activation tile is built (1 arrival, builder) AND its TMEM buffer
    // it % 5 has been read out by the epilogue of tile it-5 (4 arrivals, one per warp; pre-arrived for the first five tiles).
    // The ring has kR = 12 slots = one K-block, a multiple of the number of issuers (4), epilogue groups (4) and builders: every
    // barrier's consecutive phases are then awaited by the SAME warp in program order, which the parity wait needs (a warp that
    // could start waiting two phases ahead would see the previous phase's parity and fall through).
    // rd_done[r]: tile r read out by all four warps of its group; the builders wait for it before they rebuild slot r, which keeps
    // mma_done[r] from completing a second phase before a late epilogue warp has tested the first.
    // x_full[c & 1] / x_empty[c & 1]: activation chunk c (global chunk counter) has landed (TMA transaction bytes) / has been
    // left behind by all three builder warps.
    unsigned *tmem_slot = reinterpret_cast<unsigned *>(bars + kNumBars);
    int *s_amax = reinterpret_cast<int *>(tmem_slot + 1);        // largest |output| of this CTA
    unsigned *s_cnt = tmem_slot + 2;                             // [2]: warp-tiles through the fast / the exact path
    int4 *pxinfo = reinterpret_cast<int4 *>(tmem_slot + 4);      // [slot][48]: (run index of tap (0,0) for even / odd channel groups, tap-row mask | tap-column mask << 3, -)
    long long *outoff = reinterpret_cast<long long *>(pxinfo + kTabSlots * kPTI);   // [slot][48]: output element offset of the pixel's group-0 word, -1 = no such pixel
    uint2 *sX = reinterpret_cast<uint2 *>(smem + kFixedBytes);   // 2 chunks x GS groups x 2 segments x seg_px pixels

    // the warp index through a shuffle: the compiler then knows it (and every TMEM address / role branch derived from it) is warp-uniform
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    const int GS = 1 << p.gs_shift;
    const int nloc = (p.nitems - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // work items of this CTA
    const int total_kb = nloc * p.nkb;                          // K-blocks of this CTA, all items

    if (tid == 0) {
        for (int i = 0; i < kWRing; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 4); }
        for (int i = 0; i < kBRing; ++i) mbar_init(&rd_done[i], 4);
        for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], 4); mbar_init(&a_empty[i], kIssuers); mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], kBuilders); }
        for (int i = 0; i < kBRing; ++i) { mbar_init(&go[i], 5); mbar_init(&mma_done[i], 1); }
        for (int i = 0; i < kBufs; ++i)
            for (int k = 0; k < 4; ++k) mbar_arrive(&go[i]);
        *s_amax = 0;
        s_cnt[0] = s_cnt[1] = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kIssuer) {   // the first MMA warp owns the TMEM allocation
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int i = tid; i < kBRing * kSlotBytes / 4; i += kThreads) reinterpret_cast<unsigned *>(sB)[i] = 0u;
    __syncthreads();
    {   // rounding row (k = 28) of every lo-plane activation tile: b = 2^min(7, e) where a*b = 2^e is the constant to inject
        const int e = (SO <= 15) ? SO - 1 : SO - 9;   // `half` into LL, or half/256 into M
        const int eb = e < 7 ? e : 7;
        for (int i = tid; i < kBRing * kSets * kN; i += kThreads) {      // tile = (ring slot, pixel set); its lo plane is the second KB
            int tile = i / kN, n = i - tile * kN;
            sB[(tile * 2 + 1) * kBBytes + operand_off(n, 28)] = (unsigned char)(1u << eb);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const unsigned tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // warp-uniform for the compiler: TMEM addresses live in uniform registers

    if (warp >= kEpiWarps) {
        // the two helper warpgroups hand most of their registers to the epilogue warpgroups
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kHelperRegs));
        if (warp >= kIssuer) {
            // ===== four MMA issuer warps (tile r of every K-block with r = iw mod 4): the whole warp runs the uniform loop,
            // one elected lane issues.  Per tile ONE barrier (go[r]) gates the issue. =====
            const int iw = warp - kIssuer;
            const unsigned long long dB0 = smem_desc(sB);
            constexpr unsigned long long kBStep = kSlotBytes >> 4, kBPlane = kBBytes >> 4;   // descriptor address units (16 B)
            constexpr unsigned kMma = HALF ? 64u : 128u;
            unsigned elected;
            asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(elected));
            // weights of global K-block gbn (item gbn / nkb of this CTA, K-block gbn % nkb of that item's channel tile)
            auto wsrc_of = [&](int gbn) -> const unsigned char * {
                const int nn = gbn / p.nkb, bb = gbn - nn * p.nkb;
                const int mt = ((int)blockIdx.x + nn * (int)gridDim.x) % p.mtiles;
                return p.w + ((size_t)mt * p.nkb + bb) * kWBytes;
            };
            if (iw == 0 && elected) {
                for (int b = 0; b < kWRing && b < total_kb; ++b) {
                    mbar_expect_tx(&w_full[b], kWBytes);
                    bulk_g2s(sW + b * kWBytes, wsrc_of(b), kWBytes, &w_full[b]);
                }
            }
            // Weights of a K-block: shared memory -> registers -> tensor memory (A operand slot bn & 1), one lane quadrant per issuer
            // warp.  The issuer warps idle most of the time; doing this on the epilogue warps (round 1) put ~170 cycles per tile on the
            // kernel's critical resource - the serial read-out + compute time of an epilogue group (per-tile timeline of the
            // -DY2_TC2_PROFILE build, profiles/r2_tc2_timeline.txt).
            const int wrow = iw * 32 + lane;
            const unsigned wlane_base = tmem + ((unsigned)(iw * 32) << 16);
            auto stage_weights = [&](int bn) {
                const int s = bn % kWRing, a = bn & 1;
                mbar_wait(&w_full[s], (bn / kWRing) & 1);
                const uint4 *src = reinterpret_cast<const uint4 *>(sW + s * kWBytes + wrow * 64);
                const int sw = (wrow >> 1) & 3;
                const uint4 c0 = src[0 ^ sw], c1 = src[1 ^ sw], c2 = src[2 ^ sw], c3 = src[3 ^ sw];
                if (bn >= 2) mbar_wait(&a_empty[a], ((bn >> 1) - 1) & 1);   // the MMAs of block bn-2 have read this slot
                asm volatile("tcgen05.fence::after_thread_sync;");
                tmem_st8(wlane_base + kACol + a * 16, c0, c1);
                tmem_st8(wlane_base + kACol + a * 16 + 8, c2, c3);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;");
                __syncwarp();
                if (lane == 0) { mbar_arrive(&w_empty[s]); mbar_arrive(&a_full[a]); }
            };
            PROF_DECL
            stage_weights(0);
            int tb = iw % kBufs;                             // TMEM buffer it % 5 of this warp's next tile
            for (int b = 0; b < total_kb; ++b) {             // b = global K-block counter of this CTA
                PROF_ADD(4);
                if (iw == 0 && b >= 1 && b - 1 + kWRing < total_kb && elected) {   // the smem slot of block b-1 has been copied to TMEM: refill it with block b-1+kWRing
                    const int s = (b - 1) % kWRing;
                    mbar_wait(&w_empty[s], ((b - 1) / kWRing) & 1);
                    mbar_expect_tx(&w_full[s], kWBytes);
                    bulk_g2s(sW + s * kWBytes, wsrc_of(b - 1 + kWRing), kWBytes, &w_full[s]);
                }
                __syncwarp();
                mbar_wait(&a_full[b & 1], (b >> 1) & 1);
                PROF_ADD(0);
                const unsigned ah = tmem + kACol + (b & 1) * 16, al = ah + 8;
#pragma unroll
                for (int j = 0; j < kR / kIssuers; ++j) {
                    const int r = iw + j * kIssuers;
                    mbar_wait_m<Y2_TC2_WAIT_HELPER>(&go[r], b & 1);
                    PROF_ADD(1);
                    PROF_TL(b * kR + r, 2);
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    if (elected) {
#pragma unroll
                        for (int st = 0; st < kSets; ++st) {        // HALF: pixel set st -> lanes 16*st .. 16*st+15 of every quadrant
                            const unsigned long long dBh = dB0 + r * kBStep + st * (2 * kBPlane), dBl = dBh + kBPlane;
                            const unsigned lo16 = (unsigned)(16 * st) << 16;
                            const unsigned d0 = tmem + tb * kBufCols + lo16, ahs = ah + lo16, als = al + lo16;
                            umma_i8_ts(d0 + kPlaneCols, ahs, dBl, idesc_i8(1, 0, kMma), 0);       // M  = hi*lo
                            umma_i8_ts(d0 + kPlaneCols, als, dBh, idesc_i8(0, 1, kMma), 1);       //    + lo*hi
                            umma_i8_ts(d0 + 2 * kPlaneCols, als, dBl, idesc_i8(0, 0, kMma), 0);   // LL (+ the rounding constant through K row 28)
                            umma_i8_ts(d0, ahs, dBh, idesc_i8(1, 1, kMma), 0);                    // HH per step in columns 0..27, sum over the K-block's steps in 28..31
                        }
                        umma_commit(&mma_done[r]);                             // epilogue (tile ready) waits on it
                        if (j == kR / kIssuers - 1) umma_commit(&a_empty[b & 1]);   // this warp's reads of the weight slot are done
                    }
                    __syncwarp();
                    PROF_TL(b * kR + r, 3);
                    tb = (tb + kIssuers) % kBufs;
                    PROF_ADD(3);
                    // after this warp's first tile of block b: the MMAs of block b-1 (whose slot block b+1 takes) are done or about to be
                    if (j == 0 && b + 1 < total_kb) stage_weights(b + 1);
                }
            }
            PROF_END;
        } else if (warp == kProducer) {
            // ===== TMA producer: per work item the pixel tables, then the item's activation chunks (GS channel groups each) =====
            const int halo = (KS == 3) ? p.W + 1 : 0;
            int gc = 0;                                      // global chunk counter of this CTA
            for (int n = 0; n < nloc; ++n) {
                const int item = (int)blockIdx.x + n * (int)gridDim.x;
                const int pt = item / p.mtiles;
                const long long pix0 = (long long)pt * kPTI;
                const int f0 = (int)(pix0 / p.HW);
                const int pin0 = (int)(pix0 - (long long)f0 * p.HW);
                const int qsplit = min(kPTI, p.HW - pin0);   // pixels q >= qsplit lie in frame f0 + 1
                const int nseg = (qsplit < kPTI && f0 + 1 < p.B) ? 2 : 1;
                for (int c = 0; c < p.nchunks; ++c, ++gc) {
                    const int buf = gc & 1;
                    if (gc >= 2) mbar_wait(&x_empty[buf], ((gc >> 1) - 1) & 1);   // the builders have left the chunk that used this buffer
                    if (c == 0) {
                        // Tables of item n (ring slot n & 3), written AFTER the wait above: the builders have then finished item n-2 at least,
                        // so every epilogue warp has read a tile of item n-3 and is done with the tables of item n-4 (its output store).
                        // The builders / epilogue warps see them through the x_full -> go -> mma_done barrier chain.
                        int4 *pi = pxinfo + (n & (kTabSlots - 1)) * kPTI;
                        long long *oo = outoff + (n & (kTabSlots - 1)) * kPTI;
                        for (int q = lane; q < kPTI; q += 32) {
                            const int seg = q >= qsplit;
                            const int f = f0 + seg;
                            const int pin = seg ? q - qsplit : pin0 + q;
                            const bool valid = f < p.B && pin < p.HW;
                            const int y = pin / p.W, x = pin - y * p.W;
                            int mask = 0;
                            if (valid) {
#pragma unroll
                                for (int i = 0; i < KS; ++i) {
                                    if (y + i - PAD >= 0 && y + i - PAD < p.H) mask |= 1 << i;
                                    if (x + i - PAD >= 0 && x + i - PAD < p.W) mask |= 8 << i;
                                }
                            }
                            const int idx = seg * p.seg_px + (q - seg * qsplit);
                            const int start = (seg ? 0 : pin0) - halo;              // first pixel of the segment's run, frame-flat
                            pi[q] = make_int4(idx + (start & 1), idx + ((start + p.HW) & 1), mask, 0);   // even / odd channel group
                            oo[q] = valid ? (long long)f * p.out_frame_stride + (long long)pin * 4 : -1LL;
                        }
                        __syncwarp();
                    }
                    const int g0 = c << p.gs_shift, ng = min(GS, p.G - g0);
                    const int per_group = nseg * p.nbox, nops = ng * per_group;
                    if (lane == 0) mbar_expect_tx(&x_full[buf], (unsigned)(nops * p.boxlen * 8));
                    __syncwarp();
                    for (int o = lane; o < nops; o += 32) {
                        const int gg = o / per_group, rem = o - gg * per_group;
                        const int s = rem / p.nbox, bx = rem - s * p.nbox;
                        // a box must start on a 16-byte boundary of global memory (an unaligned start is an illegal-instruction fault,
                        // measured): the run starts at the even pixel at or below its first pixel, and the builders index it one
                        // pixel further when that rounded (the pxinfo entry carries both indices, by parity of the channel group)
                        const int c0 = ((g0 + gg) * p.HW + (s ? 0 : pin0) - halo) & ~1;
                        uint2 *dst = sX + (size_t)((((buf << p.gs_shift) + gg) << 1) + s) * p.seg_px + bx * p.boxlen;
                        tma_load_2d(dst, &tmap, c0 + bx * p.boxlen, f0 + s, &x_full[buf]);
                    }
                }
            }
        } else {
            // ===== three builder warps: write the block-diagonal activation tiles from the staged runs, a pair at a time =====
            const int bw = warp - kEpiWarps;
            // one tile = 28 entries (step slot s, pixel px) = column s*4+px; lane e < 28 owns entry e of BOTH tiles of a pair.
            // The same four hi bytes also go to rows 4s..4s+3 of the DENSE column 28+px (all seven steps of a pixel in one
            // column): the HH MMA then leaves sum_s HH_s there, which is all the fast path needs of that plane.
            const bool has = lane < kSteps * kPx;
            const int s0 = has ? lane / kPx : 0, p0 = lane - (lane / kPx) * kPx;
            const int off0 = operand_off(s0 * kPx + p0, 4 * s0);
            const int offd = operand_off(kSteps * kPx + p0, 4 * s0);
            int rel_gc = 0, have_gc = -1;               // next chunk to hand back to the producer / last chunk known to have landed
            int gb = 0;                                 // global K-block counter of this CTA
            PROF_DECL
            for (int n = 0; n < nloc; ++n) {
                const int4 *pi = pxinfo + (n & (kTabSlots - 1)) * kPTI;
                const int gc0 = n * p.nchunks;
                for (int b = 0; b < p.nkb; ++b, ++gb) {
                    PROF_ADD(4);
                    const int c_first = (min(p.G - 1, (b * kSteps) / K2)) >> p.gs_shift;
                    const int c_need = (min(p.G - 1, (b * kSteps + kSteps - 1) / K2)) >> p.gs_shift;
                    while (rel_gc < gc0 + c_first) {    // nobody in this warp reads chunk rel_gc any more (earlier items included)
                        __syncwarp();
                        mbar_arrive_if(&x_empty[rel_gc & 1], lane == 0);
                        ++rel_gc;
                    }
                    while (have_gc < gc0 + c_need) {
                        ++have_gc;
                        mbar_wait(&x_full[have_gc & 1], (have_gc >> 1) & 1);
                    }
                    PROF_ADD(0);
                    const int sg0 = b * kSteps + s0;
                    const bool live0 = has && sg0 < p.G * K2;
                    const int g0 = live0 ? sg0 / K2 : 0, t0 = sg0 - g0 * K2;
                    const int ti0 = live0 ? t0 / KS : 0, tj0 = live0 ? t0 - ti0 * KS : 0;
                    const int need = live0 ? ((1 << ti0) | (8 << tj0)) : 0x40;      // mask bits this lane's tap needs (0x40 is never set)
                    const int cbuf = (gc0 + (g0 >> p.gs_shift)) & 1;
                    const bool godd = g0 & 1;
                    const uint2 *xs0 = sX + (size_t)((((cbuf << p.gs_shift) + (g0 & (GS - 1))) << 1)) * p.seg_px + ti0 * p.W + tj0;
#pragma unroll
                    for (int jj = 0; jj < kR / 2 / kBuilders; ++jj) {
                        const int j = bw + jj * kBuilders;      // pair index in the K-block: ring slots 2j, 2j+1
                        PROF_ADD(4);
                        // every warp of the tiles' epilogue groups has READ OUT the previous K-block's tiles of these slots (which implies
                        // their MMAs have read the operand tiles): see rd_done[] above
                        if (gb >= 1) {
                            mbar_wait_m<Y2_TC2_WAIT_HELPER>(&rd_done[2 * j], (gb - 1) & 1);
                            mbar_wait_m<Y2_TC2_WAIT_HELPER>(&rd_done[2 * j + 1], (gb - 1) & 1);
                        }
                        PROF_ADD(1);
                        PROF_TL(gb * kR + 2 * j, 0); PROF_TL(gb * kR + 2 * j + 1, 0);
                        if (has) {
#pragma unroll
                            for (int st = 0; st < kSets; ++st) {    // pixel set st of the item: pixels st*48 + tile*4 + p0
                                unsigned char *bh0 = sB + (2 * j) * kSlotBytes + st * 2 * kBBytes, *bh1 = bh0 + kSlotBytes;
                                unsigned hi0 = 0, lo0 = 0, hi1 = 0, lo1 = 0;
                                const int4 ia = pi[st * kPT + 2 * j * kPx + p0], ib = pi[st * kPT + (2 * j + 1) * kPx + p0];
                                if ((ia.z & need) == need) {
                                    const uint2 xa = xs0[godd ? ia.y : ia.x];
                                    hi0 = __byte_perm(xa.x, xa.y, 0x7531);
                                    lo0 = __byte_perm(xa.x, xa.y, 0x6420);
                                }
                                if ((ib.z & need) == need) {
                                    const uint2 xb = xs0[godd ? ib.y : ib.x];
                                    hi1 = __byte_perm(xb.x, xb.y, 0x7531);
                                    lo1 = __byte_perm(xb.x, xb.y, 0x6420);
                                }
                                *reinterpret_cast<unsigned *>(bh0 + off0) = hi0;
                                *reinterpret_cast<unsigned *>(bh0 + kBBytes + off0) = lo0;
                                *reinterpret_cast<unsigned *>(bh1 + off0) = hi1;
                                *reinterpret_cast<unsigned *>(bh1 + kBBytes + off0) = lo1;
                                if constexpr (kFast) {
                                    *reinterpret_cast<unsigned *>(bh0 + offd) = hi0;
                                    *reinterpret_cast<unsigned *>(bh1 + offd) = hi1;
                                }
                            }
                        }
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        __syncwarp();
                        mbar_arrive_if(&go[2 * j], lane == 0);
                        mbar_arrive_if(&go[2 * j + 1], lane == 0);
                        PROF_TL(gb * kR + 2 * j, 1); PROF_TL(gb * kR + 2 * j + 1, 1);
                        PROF_ADD(2);
                    }
                }
            }
            PROF_END;
        }
    } else {
        // ===== epilogue warps: thread = one output channel (TMEM lane); warp group kg = warp/4 takes the tiles r = kg (mod kGroups) =====
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kEpiRegs));
        const int q4 = warp & 3, kg = warp >> 2;
        const int row = q4 * 32 + lane;
        const unsigned lane_base = tmem + ((unsigned)(q4 * 32) << 16);
        int UR[kTPG][kPx];                          // chain state (acc + 32768) of the group's three tiles x four pixels, in registers: 5.63 -> 5.77 T
                                                    // steps/s against keeping it in shared memory between tiles (profiles/README.md)
        const unsigned xmax = (kFast && !p.force_exact) ? (p.xmax_in ? (unsigned)min(max(*p.xmax_in, 0), 32768) : 32768u) : 0u;
        // fast-path bound of this channel for one K-block: Dsum >= sum over the block's steps of |rs(P_s, so)|, and the window
        // [Dsum, Dsum + lim] the offset accumulator must lie in for the block to be saturation-free (lim < 0: never)
        auto bound_of = [&](unsigned w1, int &dsum, int &lim) {
            const unsigned long long prod = (unsigned long long)w1 * xmax;
            const unsigned long long d = (prod >> SO) + 8;
            dsum = d > 40000ull ? 0x40000000 : (int)d;
            lim = 65535 - 2 * dsum;
            if (!kFast || p.force_exact || lim < 0) { dsum = 0x40000000; lim = 0; }    // U - dsum < 0 -> the unsigned compare fails
        };
        unsigned nfast = 0, nexact = 0;
        int amax = 0;
        PROF_DECL
        int tb = kg % kBufs;                        // TMEM buffer it % 5 of this group's next tile (stride kGroups)
        int gb = 0;                                 // global K-block counter of this CTA
        int mt = (int)blockIdx.x % p.mtiles;        // channel tile of the current item
        // loaded one item / one K-block ahead: a global load's latency (hundreds of cycles) must never sit between a tile becoming
        // ready and its read-out
        // HALF: lane L of quadrant q4 is channel 16*q4 + L%16 of pixel set L/16 (the weight tiles / norm tables are stored per LANE)
        const int crow = HALF ? 16 * q4 + (lane & 15) : row, pset = HALF ? lane >> 4 : 0;
        constexpr int kRows = Sh::kRows;
        int bias_next = (mt * kRows + crow < p.OFM) ? (int)p.bias[mt * kRows + crow] : 0;
        const unsigned *wn = p.wnorm + (size_t)mt * p.nkb * kM + row;
        unsigned w1_next = wn[0];
        for (int n = 0; n < nloc; ++n) {
            const int m = mt * kRows + crow;
            {
                const long long base = round_shift64((long long)bias_next, p.sb);
                const long long rb = (1LL << (33 - SO)) + 2;      // |(P + half) >> so| <= 2^(33-so): clamping the bias term there cannot change clamp16(bias + r)
                long long boff = base + 32768;
                if (boff > 65535 + rb) boff = 65535 + rb;
                if (boff < -rb) boff = -rb;
#pragma unroll
                for (int r = 0; r < kTPG; ++r) {
#pragma unroll
                    for (int j = 0; j < kPx; ++j) UR[r][j] = (int)boff;
                }
            }
            const int mt_next = (n + 1 < nloc) ? ((int)blockIdx.x + (n + 1) * (int)gridDim.x) % p.mtiles : mt;
            if (n + 1 < nloc) bias_next = (mt_next * kRows + crow < p.OFM) ? (int)p.bias[mt_next * kRows + crow] : 0;
            const unsigned *wn_next = p.wnorm + (size_t)mt_next * p.nkb * kM + row;
            for (int b = 0; b < p.nkb; ++b, ++gb) {
                PROF_ADD(4);
                int dsum, lim;
                bound_of(w1_next, dsum, lim);
                if (b + 1 < p.nkb) w1_next = wn[(size_t)(b + 1) * kM];
                else if (n + 1 < nloc) w1_next = wn_next[0];
#if Y2_TC2_BLOCKCHECK
                // one saturation-free test and one vote per K-block for the group's three tiles (the chain state is in registers):
                // 5.81 -> 5.91 T steps/s against a test per tile; a K-block in which any of the 12 pixels fails takes the exact step
                bool okb = kFast;
#pragma unroll
                for (int rr = 0; rr < kTPG; ++rr)
#pragma unroll
                    for (int j = 0; j < kPx; ++j) okb = okb && (unsigned)(UR[rr][j] - dsum) <= (unsigned)lim;
                const bool fastb = __all_sync(0xffffffffu, okb);
#endif
                PROF_ADD(0);
#pragma unroll
                for (int rr = 0; rr < kTPG; ++rr) {
                    const int r = kGroups * rr + kg;
                    PROF_ADD(4);
                    mbar_wait_m<Y2_TC2_WAIT_EPI>(&mma_done[r], gb & 1);
                    PROF_ADD(1);
                    PROF_TL(gb * kR + r, 4 + q4);
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    const unsigned base = lane_base + tb * kBufCols;
                    // column n = step*4 + pixel; hd = the dense columns 28..31 of the HH plane.  They and the 28 live M columns 32..59
                    // are ONE contiguous 32-column read (5.79 -> 5.85 T steps/s against x4 + x16/x8/x4)
                    int hm[32], ll[32];
                    int *hd = hm, *mm = hm + 4;
                    if constexpr (kLd32) {
                        tmem_ld32(base + kSteps * kPx, hm);
                    } else {
                        tmem_ld28(base + kN, mm);
                        if constexpr (kFast) tmem_ld4(base + kSteps * kPx, hd);
                    }
#if Y2_TC2_LD32 == 2
                    tmem_ld32(base + 2 * kN, ll);
#else
                    tmem_ld28(base + 2 * kN, ll);
#endif
                    // while the loads fly: chain state and the saturation-free test
                    int (&U)[kPx] = UR[rr];
#if Y2_TC2_BLOCKCHECK
                    const bool fast = fastb;
#else
                    bool ok = kFast;
#pragma unroll
                    for (int j = 0; j < kPx; ++j) ok = ok && (unsigned)(U[j] - dsum) <= (unsigned)lim;
                    const bool fast = __all_sync(0xffffffffu, ok);
#endif
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    reg_fence16(mm); reg_fence12(mm + 16); reg_fence16(ll); reg_fence12(ll + 16);
                    if constexpr (kFast) asm volatile("" : "+r"(hd[0]), "+r"(hd[1]), "+r"(hd[2]), "+r"(hd[3])::"memory");   // (on both paths: the
                                                                  // registers of an asynchronous load stay reserved until the wait above)
                    if (fast) {
                        if constexpr (kFast) {
                            // everything is in registers: tile it+5 (ring slot r+5 mod 12) may overwrite this TMEM buffer
                            asm volatile("tcgen05.fence::before_thread_sync;");
                            __syncwarp();
                            mbar_arrive_if(&go[r + kBufs < kR ? r + kBufs : r + kBufs - kR], lane == 0);
                            mbar_arrive_if(&rd_done[r], lane == 0);
                            PROF_ADD(2);
                            PROF_TL(gb * kR + r, 8 + q4);
#pragma unroll
                            for (int j = 0; j < kPx; ++j) U[j] += hd[j] * (1 << (16 - (SO <= 16 ? SO : 16)));
#pragma unroll
                            for (int nn = 0; nn < kSteps * kPx; ++nn) U[nn % kPx] += (mm[nn] * 256 + ll[nn]) >> SO;   // IMAD, LEA.HI.SX32
                            ++nfast;
                        }
                    } else {
                        // exact path: fold M and LL first, then read the per-step HH columns into the registers LL occupied
                        if constexpr (SO <= 16) {
#pragma unroll
                            for (int nn = 0; nn < kSteps * kPx; ++nn) mm[nn] = mm[nn] * 256 + ll[nn];              // t = 256 M + LL
                        } else {
#pragma unroll
                            for (int nn = 0; nn < kSteps * kPx; ++nn) mm[nn] = mm[nn] + (ll[nn] >> 8);             // M + (LL >> 8)
                        }
                        tmem_ld28(base, ll);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        reg_fence16(ll); reg_fence12(ll + 16);
                        asm volatile("tcgen05.fence::before_thread_sync;");
                        __syncwarp();
                        mbar_arrive_if(&go[r + kBufs < kR ? r + kBufs : r + kBufs - kR], lane == 0);
                        mbar_arrive_if(&rd_done[r], lane == 0);
                        PROF_ADD(2);
#pragma unroll
                        for (int nn = 0; nn < kSteps * kPx; ++nn) {
                            int d;
                            if constexpr (SO <= 16) d = ll[nn] * (1 << (16 - (SO <= 16 ? SO : 16))) + (mm[nn] >> SO);     // 65536 HH is a multiple of 2^so
                            else d = (ll[nn] * 256 + mm[nn]) >> (SO - 8);
                            U[nn % kPx] = __viaddmin_s32_relu(U[nn % kPx], d, 65535);   // VIADDMNMX.RELU: max(min(acc + d, 65535), 0)
                        }
                        ++nexact;
                    }
                    tb = (tb + kGroups) % kBufs;
                    PROF_ADD(3);
                    PROF_TL(gb * kR + r, 12 + q4);
                }
            }
            // ---- the item's 12 pixels of this thread's channel: leaky, largest |output|, store.  The four lanes of a channel quad
            // (channels 4k..4k+3 = one C4 word) transpose their 4 channels x 4 pixels through two shuffle stages, so that lane q of
            // the quad stores the complete 8-byte word of pixel q: one STG.64 per tile instead of four scattered STG.U16, and the
            // quad's four words are one full 32-byte sector (the 2-byte stores cost 1.6 % of a forward: -DY2_TC2_NOSTORE experiment)
            {
                PROF_ADD(4);
                const long long *oo = outoff + (n & (kTabSlots - 1)) * kPTI + pset * kPT;
                int16_t *om = p.out + (long long)(m >> 2) * p.HW * 4;          // the quad's C4 word plane (m >> 2 is the same for its four lanes)
                const int cq = lane & 3;
                const bool mvalid = m < p.OFM;
#pragma unroll
                for (int rr = 0; rr < kTPG; ++rr) {
                    int v[kPx];
#pragma unroll
                    for (int j = 0; j < kPx; ++j) {
                        int a = UR[rr][j] - 32768;
                        if (p.leaky && a < 0) a = a / 10;
                        if (!mvalid) a = 0;                                     // channels beyond OFM: the padding of the last C4 word stays zero
                        v[j] = a;
                    }
                    {
                        const int x0 = (cq & 2) ? v[0] : v[2], x1 = (cq & 2) ? v[1] : v[3];
                        const int r0 = __shfl_xor_sync(0xffffffffu, x0, 2), r1 = __shfl_xor_sync(0xffffffffu, x1, 2);
                        if (cq & 2) { v[0] = r0; v[1] = r1; } else { v[2] = r0; v[3] = r1; }
                        const int y0 = (cq & 1) ? v[0] : v[1], y1 = (cq & 1) ? v[2] : v[3];
                        const int s0 = __shfl_xor_sync(0xffffffffu, y0, 1), s1 = __shfl_xor_sync(0xffffffffu, y1, 1);
                        if (cq & 1) { v[0] = s0; v[2] = s1; } else { v[1] = s0; v[3] = s1; }
                    }
                    // v[c] = channel (m & ~3) + c of pixel cq
                    PROF_ADD(5);
                    const long long off = oo[(kGroups * rr + kg) * kPx + cq];
                    if (off >= 0 && (m & ~3) < p.OFM) {                        // (a quad entirely beyond OFM has no word in the output tensor)
#pragma unroll
                        for (int c = 0; c < 4; ++c) amax = max(amax, v[c] < 0 ? -v[c] : v[c]);
#ifndef Y2_TC2_NOSTORE
                        *reinterpret_cast<uint2 *>(om + off) = make_uint2(__byte_perm((unsigned)v[0], (unsigned)v[1], 0x5410), __byte_perm((unsigned)v[2], (unsigned)v[3], 0x5410));
#endif
                    }
                }
            }
            mt = mt_next;
            wn = wn_next;
            PROF_ADD(6);                            // (profile build: slot 5 = activation + quad transpose, slot 6 = table read + store)
        }
        PROF_END;
        if (p.xmax_out) {
            amax = __reduce_max_sync(0xffffffffu, amax);
            if (lane == 0) atomicMax(s_amax, amax);
        }
        if (p.stats && lane == 0) {
            if (nfast) atomicAdd(&s_cnt[0], nfast);
            if (nexact) atomicAdd(&s_cnt[1], nexact);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == kIssuer) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
    if (tid == 0) {     // one global atomic per CTA and quantity
        if (p.xmax_out && *s_amax > 0) atomicMax(p.xmax_out, *s_amax);
        if (p.stats) {
            if (s_cnt[0]) atomicAdd(p.stats, (unsigned long long)s_cnt[0]);
            if (s_cnt[1]) atomicAdd(p.stats + 1, (unsigned long long)s_cnt[1]);
        }
    }
}

// Weight tiles for the tensor-core path from one layer of the reference's reorganised blob.
// Output: [mtile][kblock][row 0..127][hi plane 32 B | lo plane 32 B], the four 16-byte chunks of a row stored at
// chunk ^ ((row>>1)&3) (conflict-free LDS.128 by 32 consecutive rows); K byte 28 = the rounding constant's weight-side factor.
__global__ void wprep_tc2_kernel(const int16_t *__restrict__ blob, unsigned char *__restrict__ dst, int ifm, int ofm, int ksize,
                                 int TM, int TN, int nkb, int so, long long total, int half)
{
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int k2 = ksize * ksize;
    int k = idx & 31;
    long long r = idx >> 5;
    int ml = r % kM; r /= kM;
    int b = r % nkb;
    int mtile = r / nkb;
    int m = half ? mtile * 64 + 16 * (ml >> 5) + (ml & 15) : mtile * kM + ml;    // HALF: tile row = TMEM lane, both pixel-set halves hold the channel
    int hi = 0, lo = 0;
    if (k < 28) {
        int sigma = b * kSteps + (k >> 2), t = k & 3;
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
        const int eb = e < 7 ? e : 7, ea = e - eb;          // a * b = 2^e, b = 2^eb <= 128, a = 2^ea <= 128
        if (so <= 15) lo = 1 << ea; else hi = 1 << ea;
    }
    unsigned char *tile = dst + ((size_t)mtile * nkb + b) * kWBytes + (size_t)ml * 64;
    const int sw = (ml >> 1) & 3;
    tile[(((k >> 4) ^ sw) << 4) + (k & 15)] = (unsigned char)hi;
    tile[(((2 + (k >> 4)) ^ sw) << 4) + (k & 15)] = (unsigned char)lo;
}

// Fast-path bound table: [mtile][kblock][row] = sum of |w| over the (up to) 28 weights of the row's K-block.
__global__ void wnorm_tc2_kernel(const int16_t *__restrict__ blob, unsigned *__restrict__ dst, int ifm, int ofm, int ksize, int TM, int TN,
                                 int nkb, long long total, int half)
{
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int k2 = ksize * ksize, G = (ifm + 3) / 4;
    const int ml = idx % kM;
    long long r = idx / kM;
    const int b = r % nkb, mtile = r / nkb;
    const int m = half ? mtile * 64 + 16 * (ml >> 5) + (ml & 15) : mtile * kM + ml;
    unsigned sum = 0;
    if (m < ofm)
        for (int k = 0; k < 28; ++k) {
            const int sigma = b * kSteps + (k >> 2), c = (sigma / k2) * 4 + (k & 3);
            if (sigma < G * k2 && c < ifm) {
                const int w = blob[reorg_woff(m, c, sigma % k2, ifm, ofm, k2, TM, TN)];
                sum += (unsigned)(w < 0 ? -w : w);
            }
        }
    dst[idx] = sum;
}

template <int KS, int SO, bool HALF>
void launch_one(const Tc2Params &p, const CUtensorMap &tmap, dim3 grid, size_t smem, cudaStream_t st)
{
    // set on every launch (a microsecond): the attribute is per device and a host may drive several GPUs from several threads,
    // so a cached "already configured" flag would be a data race for nothing
    cudaFuncSetAttribute(conv_i16_tc2_kernel<KS, SO, HALF>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    conv_i16_tc2_kernel<KS, SO, HALF><<<grid, kThreads, smem, st>>>(p, tmap);
}

template <int KS, bool HALF>
bool dispatch_so(const Tc2Params &p, const CUtensorMap &tmap, int so, dim3 grid, size_t smem, cudaStream_t st)
{
    switch (so) {
#define Y2_TC2_CASE(S) case S: launch_one<KS, S, HALF>(p, tmap, grid, smem, st); return true;
        Y2_TC2_CASE(8) Y2_TC2_CASE(9) Y2_TC2_CASE(10) Y2_TC2_CASE(11) Y2_TC2_CASE(12) Y2_TC2_CASE(13) Y2_TC2_CASE(14) Y2_TC2_CASE(15)
        Y2_TC2_CASE(16) Y2_TC2_CASE(17) Y2_TC2_CASE(18) Y2_TC2_CASE(19) Y2_TC2_CASE(20) Y2_TC2_CASE(21) Y2_TC2_CASE(22)
#undef Y2_TC2_CASE
    default: return false;
    }
}

// HALF mode (64 channels x two pixel sets per item) when the 128-row tiles would be less than 80 % full and 64-row tiles fill better
bool half_mode(int ofm)
{
    const int full128 = ceil_div(ofm, 128) * 128, full64 = ceil_div(ofm, 64) * 64;
    return ofm * 5 < full128 * 4 && full64 < full128;
}
int channel_tiles(int ofm) { return half_mode(ofm) ? ceil_div(ofm, 64) : ceil_div(ofm, kM); }

// one tile = 128 rows (TMEM lanes) x 64 B per K-block in both modes
size_t tiles_bytes(int ifm, int ofm, int ksize)
{
    const int nkb = ceil_div(ceil_div(ifm, 4) * ksize * ksize, kSteps);
    return (size_t)channel_tiles(ofm) * nkb * kWBytes;
}

}  // namespace

// operand tiles, then the fast-path bound table
size_t wprep_tc2_bytes(int ifm, int ofm, int ksize)
{
    const int nkb = ceil_div(ceil_div(ifm, 4) * ksize * ksize, kSteps);
    return tiles_bytes(ifm, ofm, ksize) + (size_t)channel_tiles(ofm) * nkb * kM * sizeof(unsigned);
}

void launch_wprep_tc2(const int16_t *blob, void *dst, int ifm, int ofm, int ksize, int TM, int TN, int so, cudaStream_t st)
{
    const int nkb = ceil_div(ceil_div(ifm, 4) * ksize * ksize, kSteps);
    const int half = half_mode(ofm);
    const long long total = (long long)channel_tiles(ofm) * nkb * kM * 32;
    wprep_tc2_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(blob, (unsigned char *)dst, ifm, ofm, ksize, TM, TN, nkb, so, total, half);
    const long long rows = total / 32;
    wnorm_tc2_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(blob, (unsigned *)((char *)dst + tiles_bytes(ifm, ofm, ksize)), ifm, ofm, ksize,
                                                                      TM, TN, nkb, rows, half);
}

// the driver's tensor-map encoder without linking libcuda: resolved once through the runtime (thread-safe static initialisation)
static PFN_cuTensorMapEncodeTiled tensor_map_encoder()
{
    static const PFN_cuTensorMapEncodeTiled fn = [] {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
        return (PFN_cuTensorMapEncodeTiled)f;
    }();
    return fn;
}

// Fills the launch plan; false = the shape / shift / alignment is not eligible for the tensor-core path.
static bool tc2_plan(const ConvFastParams &cp, int ksize, Tc2Params &p, size_t &smem, unsigned long long &frame_bytes, bool &half)
{
    if ((ksize != 1 && ksize != 3) || cp.so < 8 || cp.so > 22) return false;
    half = half_mode(cp.OFM);
    const int PTI = half ? Tc2Shape<true>::kPTI : Tc2Shape<false>::kPTI;                      // pixels per work item
    const size_t fixed = half ? Tc2Shape<true>::kFixedBytes : Tc2Shape<false>::kFixedBytes;
    p = Tc2Params{};
    p.out = (int16_t *)cp.out; p.w = (const unsigned char *)cp.w; p.bias = (const int16_t *)cp.bias;
    p.wnorm = (const unsigned *)((const char *)cp.w + tiles_bytes(cp.G * 4, cp.OFM, ksize));
    p.xmax_in = cp.xmax_in; p.xmax_out = cp.xmax_out; p.stats = cp.tc_stats; p.force_exact = cp.tc_force_exact;
    p.B = cp.B; p.H = cp.H; p.W = cp.W; p.G = cp.G; p.OFM = cp.OFM;
    p.out_frame_stride = cp.out_frame_stride;
    p.sb = cp.sb; p.leaky = cp.leaky;
    p.nkb = ceil_div(cp.G * ksize * ksize, kSteps);
    p.HW = cp.H * cp.W;
    const long long npix = (long long)cp.B * p.HW;
    const long long npt = (npix + PTI - 1) / PTI;
    const int mtiles = channel_tiles(cp.OFM);
    if (npt * mtiles > 0x3fffffffLL || (long long)cp.G * p.HW > 0x7fffffffLL - 4096) return false;
    p.npt = (int)npt;
    p.mtiles = mtiles;
    p.nitems = (int)(npt * mtiles);
    // a pixel tile may straddle ONE frame boundary (two staged runs); frames smaller than a tile only as a single frame
    if (p.HW < PTI && cp.B > 1) return false;
    // the staged run of one (group, frame segment): every pixel the taps of 48 consecutive pixels touch, in boxes of <= 256 pixels
    // whose shared-memory destinations stay 128-byte aligned (16 pixels)
    const int run = PTI + (ksize == 3 ? 2 * cp.W + 2 : 0) + 1;   // + 1: the run starts on an even pixel
    p.nbox = ceil_div(run, 256);
    p.boxlen = (ceil_div(run, p.nbox) + 15) & ~15;
    p.seg_px = p.nbox * p.boxlen;
    const size_t per_group = (size_t)2 * p.seg_px * 8;
    int gs = (int)((200 * 1024 - fixed) / (2 * per_group));
    if (gs < 1) return false;
    int sh = 0;
    while ((2 << sh) <= gs && (2 << sh) <= 16) ++sh;   // largest power of two <= min(gs, 16)
    p.gs_shift = sh;
    gs = 1 << sh;
    p.nchunks = ceil_div(cp.G, gs);
    smem = fixed + 2 * per_group * gs;
    if (smem < 120 * 1024) smem = 120 * 1024;          // one CTA per SM: a CTA allocates all 512 TMEM columns
    // tensor map over the C4 input as [frame][G*H*W] 8-byte pixels: 16-byte aligned base and frame stride
    frame_bytes = cp.B > 1 ? (unsigned long long)cp.in_frame_stride * 2 : (((unsigned long long)cp.G * p.HW * 8 + 15) & ~15ull);
    if ((frame_bytes & 15) || ((uintptr_t)cp.in & 15) || frame_bytes < (unsigned long long)cp.G * p.HW * 8) return false;
    return true;
}

bool conv_i16_tc2_eligible(const ConvFastParams &cp_in, int ksize, int frames)
{
    ConvFastParams cp = cp_in;
    cp.B = frames;
    Tc2Params p;
    size_t smem;
    unsigned long long fb;
    bool half;
    return tc2_plan(cp, ksize, p, smem, fb, half) && tensor_map_encoder() != nullptr;
}

// Returns 1 when launched, -1 when the shape/shift is not eligible for the tensor-core path, -2 when the tensor map cannot be built.
int launch_conv_i16_tc2(const ConvFastParams &cp, int ksize, cudaStream_t st, const char **variant)
{
    Tc2Params p;
    size_t smem;
    unsigned long long frame_bytes;
    bool half;
    if (!tc2_plan(cp, ksize, p, smem, frame_bytes, half)) return -1;
    const PFN_cuTensorMapEncodeTiled encode = tensor_map_encoder();
    if (!encode) return -2;
    alignas(64) CUtensorMap tmap;
    const cuuint64_t gdim[2] = {(cuuint64_t)cp.G * p.HW, (cuuint64_t)cp.B};
    const cuuint64_t gstride[1] = {(cuuint64_t)frame_bytes};
    const cuuint32_t box[2] = {(cuuint32_t)p.boxlen, 1u};
    const cuuint32_t estr[2] = {1u, 1u};
    if (encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, const_cast<void *>(cp.in), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return -2;
    int dev = 0, nsm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    if (nsm < 1) return -2;
#ifdef Y2_TC2_GRID
    nsm = Y2_TC2_GRID;                                  // experiments: fewer / more persistent CTAs than SMs
#endif
    dim3 grid((unsigned)(p.nitems < nsm ? p.nitems : nsm));
    const bool ok = ksize == 3 ? (half ? dispatch_so<3, true>(p, tmap, cp.so, grid, smem, st) : dispatch_so<3, false>(p, tmap, cp.so, grid, smem, st))
                               : (half ? dispatch_so<1, true>(p, tmap, cp.so, grid, smem, st) : dispatch_so<1, false>(p, tmap, cp.so, grid, smem, st));
    if (!ok) return -1;
#ifdef Y2_TC2_PROFILE
    {
        cudaStreamSynchronize(st);
        long long h[32 * 8];
        cudaMemcpyFromSymbol(h, g_tc2_prof, sizeof(h));
        fprintf(stderr, "tc2 profile (CTA 1,0) G=%d W=%d nkb=%d: per warp [w_stage|wait a, wait b_full|mma_done, wait t_empty|ld, issue|compute, other, -, -, total]\n", cp.G, cp.W, p.nkb);
        int hd[32 * 4], ab = 0;
        cudaMemcpyFromSymbol(hd, g_tc2_dbg, sizeof(hd));
        cudaMemcpyFromSymbol(&ab, g_tc2_abort, sizeof(ab));
        if (ab) {
            fprintf(stderr, "tc2 DEADLOCK (CTA 1,0): per warp [line, barrier smem addr, parity]\n");
            for (int w = 0; w < kThreads / 32; ++w) fprintf(stderr, "  warp %2d: line %d bar 0x%x parity %d\n", w, hd[w * 4], hd[w * 4 + 1], hd[w * 4 + 2]);
            int z[32 * 4] = {0}; ab = 0;
            cudaMemcpyToSymbol(g_tc2_dbg, z, sizeof(z)); cudaMemcpyToSymbol(g_tc2_abort, &ab, sizeof(ab));
        }
        for (int w = 0; w < kThreads / 32; ++w) {
            fprintf(stderr, "  warp %2d:", w);
            for (int i = 0; i < 8; ++i) fprintf(stderr, " %9lld", h[w * 8 + i]);
            fprintf(stderr, "\n");
        }
        if ((long long)p.nkb * kR * ((p.nitems + (int)grid.x - 1) / (int)grid.x) >= Y2_TC2_TL0 + 48) {
            long long tl[48 * 16];
            cudaMemcpyFromSymbol(tl, g_tc2_tl, sizeof(tl));
            fprintf(stderr, "tc2 timeline (CTA 1,0), cycles relative to the first event; per tile: build start | built | issuer awake | issued | epilogue awake x4 | read out x4 | computed x4\n");
            long long t0 = tl[0];
            for (int i = 0; i < 48 * 16; ++i) if (tl[i] && tl[i] < t0) t0 = tl[i];
            for (int t = 0; t < 48; ++t) {
                fprintf(stderr, "  tile %3d (slot %2d buf %d):", Y2_TC2_TL0 + t, (Y2_TC2_TL0 + t) % kR, (Y2_TC2_TL0 + t) % kBufs);
                for (int e = 0; e < 16; ++e) fprintf(stderr, " %6lld", tl[t * 16 + e] - t0);
                fprintf(stderr, "\n");
            }
        }
    }
#endif
    if (variant) *variant = ksize == 3 ? (half ? "conv_i16_tc2<3,half>" : "conv_i16_tc2<3>") : (half ? "conv_i16_tc2<1,half>" : "conv_i16_tc2<1>");
    return 1;
}

}  // namespace y2
