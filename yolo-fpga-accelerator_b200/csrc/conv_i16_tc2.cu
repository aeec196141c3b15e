// INT16 convolution on the 5th-generation tensor cores (tcgen05 + TMEM), bit-exact - pipeline v2.
//
// Arithmetic (hls/core/core_compute.cpp:65-120): every (4-channel group x tap) step of the reference's
// chain is  acc = clamp16(acc + ((P + half) >> so)),  P = sum_{t<4} w_t * x_t.  Each int16 operand is
// split into a signed-high and an unsigned-low int8 plane; four tcgen05.mma kind::i8 products give
//   HH = sum wh*xh,  M = sum (wh*xl + wl*xh),  LL = sum wl*xl (+ the rounding constant through K row 28)
// with P + half = 65536*HH + 256*M + LL.  One MMA K-slice (32) holds seven consecutive steps as a
// block-diagonal activation operand, so every TMEM column is one exact 4-MAC partial sum.
//
// What changed against csrc/conv_i16_tc.cu (the first tcgen05 version):
//  * weights are the A operand FROM TENSOR MEMORY (tcgen05.mma [d],[a],b-desc): an N=32..48 MMA with
//    A in shared memory re-reads 4 KB of weights per instruction and is shared-memory bound;
//    the epilogue warps copy each 8 KB K-block smem -> registers -> TMEM (tcgen05.st) once per K-block;
//  * N = 32 (8 step slots x 4 pixels) and FIVE 96-column accumulator buffers instead of three 144-column
//    ones: an epilogue warp loads a whole tile (tcgen05.ld x16 x 6) and releases the buffer before it
//    computes, so the MMA of a buffer's next tile overlaps three tiles of epilogue work;
//  * a 4-instruction step for 8 <= so <= 16 (so is a template parameter; LEA.HI needs an immediate):
//       t = 256*M + LL            IMAD   (fits: |256 M| < 2^27, 0 <= LL < 2^19)
//       a = HH * 2^(16-so) + acc  IMAD / LEA   (65536*HH is a multiple of 2^so: no rounding involved)
//       a = a + (t >> so)         LEA.HI.SX32
//       acc = max(min(a, 65535), 0)   VIMNMX.RELU   (acc is kept as acc+32768)
//    and for 17 <= so <= 22:  c = 256*HH + M; c += LL >> 8; a = acc + (c >> (so-8)); clamp.
//    Measured issue cost (profiles/microbench/tc_epilogue_rates.cu): 4.55 cycles per warp-step per SMSP
//    against 6.27 for the 5-instruction scaled step of v1 and ~9.9 for the CUDA-core kernel.
//  * (end of round 1) the kernel turned out to be bound by the ISSUE SLOTS of the SM sub-partitions (ncu: 80 % busy, 4 of 7 executed
//    instructions useful), so the instruction count was cut: warp-uniform role / TMEM addressing (shuffle-broadcast warp index),
//    one address operand per plane read-out, predicated barrier arrives, and for so = 14..16 a THREE-instruction step - the HH
//    product is issued 2^(16-so) times into its accumulator, TMEM delivers HH*2^(16-so), and  a = HHs + (t >> so)  is one
//    LEA.HI.SX32.  4.17 -> 5.03 T steps/s on the 13x13x1024 layers; DESIGN.md section 4, profiles/r1_tc2_handoff_experiments.md.
//  * four epilogue groups (chain state in shared memory between tiles) and a read-out handshake (rd_done[]) that makes the barrier
//    ring phase-safe: no barrier can complete a second phase before every waiter has tested the first.
#include "common.cuh"
#include <cstdio>

namespace y2 {

namespace {

constexpr int kM = 128;             // output channels per CTA = TMEM lanes
constexpr int kSteps = 7;           // chain steps per K-block: K = 32 = 7 x 4 channels | rounding row | 3 zero rows
constexpr int kPx = 4;              // pixels per tile
constexpr int kN = 32;              // MMA N = 8 step slots x 4 pixels (slot 7 unused)
constexpr int kR = 12;              // tiles per K-block -> 48 pixels per CTA
constexpr int kPT = kPx * kR;
#ifndef Y2_TC2_PLANES
#define Y2_TC2_PLANES 3
#endif
// Y2_TC2_PLANES = 2: the LL product leaves the tensor core (LL = dp4a(w_lo, x_lo) + half on the CUDA cores, one more FMA-pipe instruction per
// step), a tile is HH | M = 64 columns and SEVEN tiles are in flight in tensor memory instead of five (the kernel is bound by
// steps in flight / hand-off round trip), three MMAs and six loads per tile instead of four and nine.
constexpr int kPlanes = Y2_TC2_PLANES;
constexpr bool kLLCuda = kPlanes == 2;
constexpr int kBufs = kLLCuda ? 7 : 5;   // TMEM accumulator buffers: HH | M [| LL], 32 columns each
#ifndef Y2_TC2_GROUPS
#define Y2_TC2_GROUPS 4
#endif
constexpr int kPlaneCols = kN;
constexpr int kBufCols = kPlanes * kN;
constexpr int kACol = kBufs * kBufCols;   // 480 / 448: two weight slots of 16 columns (hi plane 8 | lo plane 8)
constexpr int kXlBytes = kLLCuda ? 2 * 12 * 128 : 0;   // compact copy of the lo activation bytes for the CUDA-core LL: [K-block parity][tile][step][pixel] words
constexpr int kBRing = 12;          // activation tile ring (hi 1 KB | lo 1 KB) = the kR tiles of one K-block (see the go[] comment in the kernel)
constexpr int kWRing = 3;           // weight K-block ring in shared memory
#ifndef Y2_TC2_GROUPS
#define Y2_TC2_GROUPS 4
#endif
constexpr int kGroups = Y2_TC2_GROUPS;   // epilogue warpgroups.  A group's tile costs wait (~250 cycles) + read-out (~250) + compute (~340) per warp
                                    //   (role profile, profiles/r1_tc2_role_profile.txt): with three warps per SM sub-partition the ALU pipe idles
                                    //   ~45 % of the time, so FOUR groups share it.  Registers: 4 groups cannot hold 96 + 16 each, so a tile is read
                                    //   as its 28 live columns (84 registers) and the chain state U lives in shared memory between tiles.
constexpr int kTPG = kR / kGroups;  // tiles per group per K-block
constexpr int kEpiWarps = 4 * kGroups;   // warps 0..: group kg = warp/4, TMEM lane quadrant = warp%4
constexpr int kBuilders = 3;        // next warpgroup: builder bw takes the tile PAIRS (2j, 2j+1) with j = bw (mod 3); its fourth warp idles
constexpr int kIssuer = kEpiWarps + 4;   // last warpgroup: issuer iw takes the tiles r = iw (mod 4).  One warp issues one MMA per ~29 cycles
constexpr int kIssuers = 4;         //   (profiles/microbench/umma_issue.cu) and pays ~100 cycles per mbarrier wait: four warps keep the
                                    //   per-tile issue cost (wait + 4 MMAs + commit, ~300 cycles) below the epilogue's ~140 cycles per tile
constexpr int kThreads = (kEpiWarps + 8) * 32;   // 3 groups: 640 threads, 96 registers at launch; 4 groups: 768 threads, 80 at launch.  Then
constexpr int kEpiRegs = kGroups == 3 ? 128 : 96;   // setmaxnreg moves 8 x 32 x (launch - 48) registers from the helper warps to the epilogue warps
constexpr int kHelperRegs = 48;     //   (+32 each for 3 groups, +16 for 4).  The CTA pool only holds what the helpers release.
#ifndef Y2_TC2_RD_HANDSHAKE
#define Y2_TC2_RD_HANDSHAKE 1
#endif
constexpr bool kRdHandshake = kLLCuda || Y2_TC2_RD_HANDSHAKE != 0;   // builders also wait for the READ-OUT of a slot's previous tile (see the builder loop)
#ifndef Y2_TC2_SLEEP_BUILDER
#define Y2_TC2_SLEEP_BUILDER 0
#endif
#ifndef Y2_TC2_SLEEP_ISSUER
#define Y2_TC2_SLEEP_ISSUER 0
#endif
#ifndef Y2_TC2_SLEEP_EPI
#define Y2_TC2_SLEEP_EPI 0
#endif
constexpr bool kUShared = kGroups != 3;
static_assert(!kLLCuda || kUShared, "the two-plane build uses the four-group epilogue");
constexpr int kUBytes = kUShared ? kEpiWarps * 32 * kTPG * 16 : 0;   // chain state [tile of the group][epilogue thread] x 4 pixels
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
__device__ __forceinline__ unsigned mbar_test(void *bar, unsigned parity)   // non-blocking probe
{
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok;
}
// wait with a back-off between probes: a failed try_wait returns after ~60 cycles, and the probe loops of the ~9 warps that are waiting at
// any time were 20 % of all issued instructions (ncu source counters) on sub-partitions whose issue slots are 80 % busy
template <int NS>
__device__ __forceinline__ void mbar_wait_sleep(void *bar, unsigned parity)
{
    if constexpr (NS == 0) {
        mbar_wait(bar, parity);
    } else {
#ifdef Y2_TC2_PROFILE
        mbar_wait(bar, parity);
#else
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "WAIT_LOOP:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
            "@p bra DONE;\n\t"
            "nanosleep.u32 %2;\n\t"
            "bra WAIT_LOOP;\n\t"
            "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity), "n"(NS)
            : "memory");
#endif
    }
}
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
__device__ __forceinline__ void tmem_ld16(unsigned taddr, int *r)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8(unsigned taddr, int *r)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
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
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gsrc));
}
__device__ __forceinline__ void bar_sync_named(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

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

#ifdef Y2_TC2_PROFILE
__device__ long long g_tc2_prof[32 * 8];
#define PROF_DECL long long prof_[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long pt_ = clock64(); const long long pstart_ = pt_;
#define PROF_ADD(i) do { long long n_ = clock64(); prof_[i] += n_ - pt_; pt_ = n_; } while (0)
#define PROF_END do { prof_[7] = clock64() - pstart_; if (blockIdx.x == 1 && blockIdx.y == 0 && lane == 0) for (int i_ = 0; i_ < 8; ++i_) g_tc2_prof[warp * 8 + i_] = prof_[i_]; } while (0)
#else
#define PROF_DECL
#define PROF_ADD(i)
#define PROF_END
#endif

struct Tc2Params {
    const uint2 *in;          // C4 input
    int16_t *out;             // C4 output (already offset to the first output group)
    const unsigned char *w;   // [mtile][kblock][128 rows][64 B]
    const int16_t *bias;
    int B, H, W, G, OFM;
    long long in_frame_stride, out_frame_stride;  // elements
    int sb, leaky;
    int nkb;                  // K-blocks = ceil(G*K2/7)
    int ctab_cap;             // entries reserved for the activation copy table
    int PW, rows_max, gs_shift;  // staging: smem row pitch (pixels), band rows incl. halo + zero row, log2(groups per chunk)
};

// one exact step of the chain from the three int8-plane partial sums.  Everything but the last instruction is independent
// of the accumulator, so the serial chain through a tile is one VIADDMNMX per step and the rest is free to overlap.
#ifndef Y2_TC2_HH_COPIES
#define Y2_TC2_HH_COPIES 1
#endif
// With Y2_TC2_HH_COPIES the HH product is issued 2^(16-so) times into the same accumulator (so = 14, 15: 4 or 2 MMAs), so the tensor core
// delivers HH * 2^(16-so) and the step is THREE instructions (IMAD, LEA.HI.SX32, VIADDMNMX) - the kernel is issue-slot bound.
template <int SO>
__host__ __device__ constexpr int hh_copies() { return (Y2_TC2_HH_COPIES != 0 && SO >= 14 && SO <= 16) ? 1 << (16 - SO) : 1; }

template <int SO>
__device__ __forceinline__ int tc2_step(int acc, int hh, int mm, int ll)
{
    int d;
    if constexpr (SO <= 16 && hh_copies<SO>() == (1 << (16 - SO))) {
        const int t = mm * 256 + ll;                       // IMAD
        d = hh + (t >> SO);                                // LEA.HI.SX32 (hh arrives pre-scaled)
    } else if constexpr (SO <= 16) {
        const int t = mm * 256 + ll;                       // IMAD
        d = hh * (1 << (16 - SO)) + (t >> SO);             // IMAD/SHL + LEA.HI.SX32
    } else {
        int c = hh * 256 + mm;                             // IMAD
        c += ll >> 8;                                      // LEA.HI.SX32
        d = c >> (SO - 8);                                 // SHF
    }
    return __viaddmin_s32_relu(acc, d, 65535);             // VIADDMNMX.RELU: max(min(acc + d, 65535), 0)
}

__device__ __forceinline__ int dp4a_uu(unsigned a, unsigned b, unsigned c)
{
    unsigned d;
    asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return (int)d;
}

template <int KS, int SO>
__global__ void __launch_bounds__(kThreads, 1) conv_i16_tc2_kernel(const Tc2Params p)
{
    constexpr int K2 = KS * KS;
    constexpr int PAD = KS / 2;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *sW = smem;                                    // kWRing x 8 KB
    unsigned char *sB = sW + kWRing * kWBytes;                   // kBRing x (hi 1 KB | lo 1 KB); a builder pair = two consecutive slots
    int4 *sU = reinterpret_cast<int4 *>(sB + kBRing * 2 * kBBytes);   // chain state between tiles (four-group build)
    uint4 *sXl = reinterpret_cast<uint4 *>(sB + kBRing * 2 * kBBytes + kUBytes);   // two-plane build: [b & 1][tile r][step] x 4 pixels
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(sB + kBRing * 2 * kBBytes + kUBytes + kXlBytes);
    unsigned long long *w_full = bars, *w_empty = w_full + kWRing, *a_full = w_empty + kWRing, *a_empty = a_full + 2,
                       *go = a_empty + 2, *mma_done = go + kBRing, *rd_done = mma_done + kBRing;   // rd_done[kBRing]: tile read out by all four warps of its group (two-plane build)
    // go[r]: tile r of the current K-block may be issued = its activation tile is built (1 arrival, builder) AND its TMEM buffer
    // it % 5 has been read out by the epilogue of tile it-5 (4 arrivals, one per warp; pre-arrived for the first five tiles).
    // The ring has kR = 12 slots = one K-block, a multiple of the number of issuers (4), epilogue groups (3) and builders: every
    // barrier's consecutive phases are then awaited by the SAME warp in program order, which the parity wait needs (a warp that
    // could start waiting two phases ahead would see the previous phase's parity and fall through).
    unsigned *tmem_slot = reinterpret_cast<unsigned *>(rd_done + kBRing + 1);
    int *pxtab = reinterpret_cast<int *>(tmem_slot + 4);         // [48][4]: smem pixel offset for tap rows 0..2, valid flag
    int *rowinfo = pxtab + kPT * 4;                              // [32][2]: per staged row slot: first needed column, prefix of the copy table
    int2 *ctab = reinterpret_cast<int2 *>(rowinfo + 64);         // copy table of one C4 group plane: (global pixel offset, smem pixel offset)
    uint2 *sX = reinterpret_cast<uint2 *>(ctab + p.ctab_cap);    // 2 chunks x GS groups x rows_max x PW pixels

    // the warp index through a shuffle: the compiler then knows it (and every TMEM address / role branch derived from it) is warp-uniform
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    const long long npix = (long long)p.B * p.H * p.W;
    const long long pix0 = (long long)blockIdx.x * kPT;
    const int mtile = blockIdx.y;
    const int zero_slot = p.rows_max - 1;
    const int GS = 1 << p.gs_shift;
    const int chunk_px = GS * p.rows_max * p.PW;
    const long long row_first = pix0 / p.W;                      // global row (frame*H + y) of the first pixel

    if (tid == 0) {
        for (int i = 0; i < kWRing; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], kLLCuda ? 4 + kEpiWarps : 4); }
        for (int i = 0; i < kBRing; ++i) mbar_init(&rd_done[i], 4);
        for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], 4); mbar_init(&a_empty[i], kIssuers); }
        for (int i = 0; i < kBRing; ++i) { mbar_init(&go[i], 5); mbar_init(&mma_done[i], 1); }
        for (int i = 0; i < kBufs; ++i)
            for (int k = 0; k < 4; ++k) mbar_arrive(&go[i]);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kIssuer) {   // the first MMA warp owns the TMEM allocation
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int q = tid; q < kPT; q += kThreads) {
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
    for (int i = tid; i < kBRing * 2 * kBBytes / 4; i += kThreads) reinterpret_cast<unsigned *>(sB)[i] = 0u;
    // Activation copy table.  Staged row slot s holds global row row_first - PAD + s; only the columns some pixel of this CTA reads
    // are copied (a CTA's pixels are consecutive, so on wide images it touches a fraction of each row), and the loader walks a
    // precomputed (source, destination) list instead of doing index arithmetic per element.
    if (tid == 0) {
        const long long pix_last = (pix0 + kPT < npix ? pix0 + kPT : npix) - 1;
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
    __syncthreads();
    {   // rounding row (k = 28) of every lo-plane activation tile: b = 2^min(7, e) where a*b = 2^e is the constant to inject
        const int e = (SO <= 15) ? SO - 1 : SO - 9;   // `half` into LL, or half/256 into M
        const int eb = e < 7 ? e : 7;
        for (int i = tid; i < kBRing * kN; i += kThreads) {
            int slot = i / kN, n = i - slot * kN;
            sB[(slot * 2 + 1) * kBBytes + operand_off(n, 28)] = (unsigned char)(1u << eb);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const unsigned tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // warp-uniform for the compiler: TMEM addresses live in uniform registers

    if (warp >= kEpiWarps) {
        // the two helper warpgroups hand most of their registers to the three epilogue warpgroups
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kHelperRegs));
        if (warp >= kIssuer) {
            // ===== four MMA issuer warps (tile r of every K-block with r = iw mod 4): the whole warp runs the uniform loop,
            // one elected lane issues.  Per tile ONE barrier (go[r]) gates the issue. =====
            const int iw = warp - kIssuer;
            const unsigned long long dB0 = smem_desc(sB);
            constexpr unsigned long long kBStep = (2 * kBBytes) >> 4, kBPlane = kBBytes >> 4;   // descriptor address units (16 B)
            unsigned elected;
            asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(elected));
            const unsigned char *wsrc = p.w + (size_t)mtile * p.nkb * kWBytes;
            if (iw == 0 && elected) {
                for (int b = 0; b < kWRing && b < p.nkb; ++b) {
                    mbar_expect_tx(&w_full[b], kWBytes);
                    bulk_g2s(sW + b * kWBytes, wsrc + (size_t)b * kWBytes, kWBytes, &w_full[b]);
                }
            }
            PROF_DECL
            int tb = iw;                                     // TMEM buffer it % 5 of this warp's next tile
            for (int b = 0; b < p.nkb; ++b) {
                PROF_ADD(4);
                if (iw == 0 && b >= 1 && b + 2 < p.nkb && elected) {   // the smem slot of block b-1 has been copied to TMEM: refill it with block b+2
                    const int s = (b - 1) % kWRing;
                    mbar_wait(&w_empty[s], ((b - 1) / kWRing) & 1);
                    mbar_expect_tx(&w_full[s], kWBytes);
                    bulk_g2s(sW + s * kWBytes, wsrc + (size_t)(b + 2) * kWBytes, kWBytes, &w_full[s]);
                }
                __syncwarp();
                mbar_wait(&a_full[b & 1], (b >> 1) & 1);
                PROF_ADD(0);
                const unsigned ah = tmem + kACol + (b & 1) * 16, al = ah + 8;
#pragma unroll
                for (int j = 0; j < kR / kIssuers; ++j) {
                    const int r = iw + j * kIssuers;
                    mbar_wait_sleep<Y2_TC2_SLEEP_ISSUER>(&go[r], b & 1);
                    PROF_ADD(1);
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    if (elected) {
                        const unsigned long long dBh = dB0 + r * kBStep, dBl = dBh + kBPlane;
                        const unsigned d0 = tmem + tb * kBufCols;
                        umma_i8_ts(d0, ah, dBh, idesc_i8(1, 1), 0);            // HH
#pragma unroll
                        for (int cpy = 1; cpy < hh_copies<SO>(); ++cpy) umma_i8_ts(d0, ah, dBh, idesc_i8(1, 1), 1);   // ... x 2^(16-so), exact in int32
                        umma_i8_ts(d0 + kPlaneCols, ah, dBl, idesc_i8(1, 0), 0);       // M  = hi*lo
                        umma_i8_ts(d0 + kPlaneCols, al, dBh, idesc_i8(0, 1), 1);       //    + lo*hi
                        if constexpr (!kLLCuda) umma_i8_ts(d0 + 2 * kPlaneCols, al, dBl, idesc_i8(0, 0), 0);   // LL
                        umma_commit(&mma_done[r]);                             // epilogue (tile ready) and builders (slot free) wait on it
                        if (j == kR / kIssuers - 1) umma_commit(&a_empty[b & 1]);   // this warp's reads of the weight slot are done
                    }
                    __syncwarp();
                    tb = tb >= kBufs - kIssuers ? tb - (kBufs - kIssuers) : tb + kIssuers;
                    PROF_ADD(3);
                }
            }
            PROF_END;
        } else if (warp < kEpiWarps + kBuilders) {
            // ===== three builder warps: stage activations (cp.async) and write the block-diagonal activation tiles, a pair at a time =====
            const int bw = warp - kEpiWarps;
            const int bt = bw * 32 + lane;              // 0..95
            constexpr int kBT = kBuilders * 32;
            const int nrows = p.rows_max - 1;           // staged band rows (the last slot is the all-zero row)
            const int nchunks = (p.G + GS - 1) >> p.gs_shift;
            auto stage_chunk = [&](int c) {
                uint2 *dst = sX + (c & 1) * chunk_px;
                const int g0 = c << p.gs_shift, ng = min(GS, p.G - g0);
                const int per_group = rowinfo[2 * (p.rows_max - 1) + 1];
                const long long plane = (long long)p.H * p.W;
                for (int idx = bt; idx < ng * per_group; idx += kBT) {
                    const int gg = idx / per_group;
                    const int2 e = ctab[idx - gg * per_group];
                    cp_async8(dst + gg * p.rows_max * p.PW + e.y, p.in + (g0 + gg) * plane + e.x);
                }
                asm volatile("cp.async.commit_group;");
            };
            // one tile = 28 entries (step slot s, pixel px) = column s*4+px; lane e < 28 owns entry e of BOTH tiles of a pair
            const bool has = lane < kSteps * kPx;
            const int s0 = has ? lane / kPx : 0, p0 = lane - (lane / kPx) * kPx;
            const int off0 = operand_off(s0 * kPx + p0, 4 * s0);
            stage_chunk(0);
            int staged = 0, ready = -1;
            PROF_DECL
            for (int b = 0; b < p.nkb; ++b) {
                PROF_ADD(4);
                const int c_first = (min(p.G - 1, (b * kSteps) / K2)) >> p.gs_shift;
                const int c_need = (min(p.G - 1, (b * kSteps + kSteps - 1) / K2)) >> p.gs_shift;
                if (staged + 1 < nchunks && staged <= c_first) {
                    bar_sync_named(1, kBT);             // every builder is past K-block b-1: nobody reads chunk staged-1 any more
                    stage_chunk(staged + 1);
                    ++staged;
                }
                if (ready < c_need) {
                    if (staged > c_need) asm volatile("cp.async.wait_group 1;" ::: "memory");
                    else asm volatile("cp.async.wait_group 0;" ::: "memory");
                    bar_sync_named(1, kBT);             // all builder warps see each other's copies
                    ready = c_need;
                }
                PROF_ADD(0);
                const int sg0 = b * kSteps + s0;
                const bool live0 = has && sg0 < p.G * K2;
                const int g0 = live0 ? sg0 / K2 : 0, t0 = sg0 - g0 * K2;
                const int ti0 = live0 ? t0 / KS : 0, tj0 = live0 ? t0 - ti0 * KS : 0;
                const uint2 *xs0 = sX + ((g0 >> p.gs_shift) & 1) * chunk_px + (g0 & (GS - 1)) * p.rows_max * p.PW + tj0;
#pragma unroll
                for (int jj = 0; jj < kR / 2 / kBuilders; ++jj) {
                    const int j = bw + jj * kBuilders;      // pair index in the K-block: ring slots 2j, 2j+1
                    PROF_ADD(4);
                    if (b >= 1 && !kRdHandshake) {          // the MMAs of the previous K-block's tiles in both slots have read them
                        mbar_wait(&mma_done[2 * j], (b - 1) & 1);    // (implied by rd_done below when the handshake is on: a tile is read out after its MMAs)
                        mbar_wait(&mma_done[2 * j + 1], (b - 1) & 1);
                    }
                    if constexpr (kRdHandshake) {
                        // every warp of the tiles' epilogue groups has READ OUT the previous K-block's tiles of these slots.  This (a) keeps
                        // mma_done[r] from completing a second phase before a late warp has tested the first (seven tiles in flight leave
                        // one hop of slack otherwise - measured as hangs), and (b) protects the compact lo bytes: a group reads out tile
                        // (b-1, r) only after it has finished computing tile (b-2, r), whose copy this build overwrites
                        if (b >= 1) {
                            mbar_wait_sleep<Y2_TC2_SLEEP_BUILDER>(&rd_done[2 * j], (b - 1) & 1);
                            mbar_wait_sleep<Y2_TC2_SLEEP_BUILDER>(&rd_done[2 * j + 1], (b - 1) & 1);
                        }
                    }
                    PROF_ADD(1);
                    unsigned char *bh = sB + (j * 4) * kBBytes;
                    if (has) {
                        unsigned hi0 = 0, lo0 = 0, hi1 = 0, lo1 = 0;
                        if (live0) {
                            const uint2 xa = xs0[pxtab[(2 * j * kPx + p0) * 4 + ti0]];
                            const uint2 xb = xs0[pxtab[((2 * j + 1) * kPx + p0) * 4 + ti0]];
                            hi0 = __byte_perm(xa.x, xa.y, 0x7531);
                            lo0 = __byte_perm(xa.x, xa.y, 0x6420);
                            hi1 = __byte_perm(xb.x, xb.y, 0x7531);
                            lo1 = __byte_perm(xb.x, xb.y, 0x6420);
                        }
                        *reinterpret_cast<unsigned *>(bh + off0) = hi0;
                        *reinterpret_cast<unsigned *>(bh + kBBytes + off0) = lo0;
                        *reinterpret_cast<unsigned *>(bh + 2 * kBBytes + off0) = hi1;
                        *reinterpret_cast<unsigned *>(bh + 3 * kBBytes + off0) = lo1;
                        if constexpr (kLLCuda) {            // entry e = step*4 + pixel of tiles 2j, 2j+1
                            unsigned *xl = reinterpret_cast<unsigned *>(sXl) + ((b & 1) * kBRing + 2 * j) * 32 + lane;
                            xl[0] = lo0;
                            xl[32] = lo1;
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    mbar_arrive_if(&go[2 * j], lane == 0);
                    mbar_arrive_if(&go[2 * j + 1], lane == 0);
                    PROF_ADD(2);
                }
            }
            PROF_END;
        }
    } else {
        // ===== epilogue warps: thread = one output channel (TMEM lane); warp group kg = warp/4 takes the tiles r = kg (mod kGroups) =====
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kEpiRegs));
        const int q4 = warp & 3, kg = warp >> 2;
        const int row = q4 * 32 + lane;
        const int m = mtile * kM + row;
        const unsigned lane_base = tmem + ((unsigned)(q4 * 32) << 16);
        int U[kUShared ? 1 : kTPG][kPx];
        int4 *myU = sU + tid;                       // [tile of the group] at stride kEpiWarps * 32
        {
            long long bv = (m < p.OFM) ? (long long)p.bias[m] : 0;
            long long base = round_shift64(bv, p.sb);
            const long long rb = (1LL << (33 - SO)) + 2;      // |(P + half) >> so| <= 2^(33-so): clamping the bias term there cannot change clamp16(bias + r)
            long long boff = base + 32768;
            if (boff > 65535 + rb) boff = 65535 + rb;
            if (boff < -rb) boff = -rb;
            if constexpr (kUShared) {
#pragma unroll
                for (int r = 0; r < kTPG; ++r) myU[r * kEpiWarps * 32] = make_int4((int)boff, (int)boff, (int)boff, (int)boff);
            } else {
#pragma unroll
                for (int r = 0; r < kTPG; ++r)
#pragma unroll
                    for (int j = 0; j < kPx; ++j) U[r][j] = (int)boff;
            }
        }
        // copy one K-block of weights shared memory -> registers -> tensor memory (A operand slot bn & 1)
        auto stage_weights = [&](int bn) {
            const int s = bn % kWRing, a = bn & 1;
            mbar_wait(&w_full[s], (bn / kWRing) & 1);
            const uint4 *src = reinterpret_cast<const uint4 *>(sW + s * kWBytes + row * 64);
            const int sw = (row >> 1) & 3;
            const uint4 c0 = src[0 ^ sw], c1 = src[1 ^ sw], c2 = src[2 ^ sw], c3 = src[3 ^ sw];
            if (bn >= 2) mbar_wait(&a_empty[a], ((bn >> 1) - 1) & 1);   // the MMAs of block bn-2 have read this slot
            asm volatile("tcgen05.fence::after_thread_sync;");
            tmem_st8(lane_base + kACol + a * 16, c0, c1);
            tmem_st8(lane_base + kACol + a * 16 + 8, c2, c3);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;");
            __syncwarp();
            if (lane == 0) { mbar_arrive(&w_empty[s]); mbar_arrive(&a_full[a]); }
        };
        PROF_DECL
        if (kg == 0) stage_weights(0);
        int tb = kg % kBufs;                        // TMEM buffer it % 5 of this group's next tile (stride kGroups)
        for (int b = 0; b < p.nkb; ++b) {
            PROF_ADD(4);
            if (b + 1 < p.nkb && (b + 1) % kGroups == kg) stage_weights(b + 1);
            unsigned wl[8];                         // two-plane build: the lo bytes of this channel's weights for the 7 steps of the K-block
            if constexpr (kLLCuda) {
                const int s = b % kWRing;
                mbar_wait(&w_full[s], (b / kWRing) & 1);
                const uint4 *src = reinterpret_cast<const uint4 *>(sW + s * kWBytes + row * 64);
                const int sw = (row >> 1) & 3;
                const uint4 c2 = src[2 ^ sw], c3 = src[3 ^ sw];
                wl[0] = c2.x; wl[1] = c2.y; wl[2] = c2.z; wl[3] = c2.w; wl[4] = c3.x; wl[5] = c3.y; wl[6] = c3.z; wl[7] = 0u;
                __syncwarp();
                if (lane == 0) mbar_arrive(&w_empty[s]);
            }
            PROF_ADD(0);
#pragma unroll
            for (int rr = 0; rr < kTPG; ++rr) {
                const int r = kGroups * rr + kg;
                PROF_ADD(4);
                mbar_wait_sleep<Y2_TC2_SLEEP_EPI>(&mma_done[r], b & 1);
                PROF_ADD(1);
                asm volatile("tcgen05.fence::after_thread_sync;");
                const unsigned base = lane_base + tb * kBufCols;
                int hh[32], mm[32], ll[32];      // column n = step*4 + pixel
                if constexpr (kLLCuda) {         // HH and M only: 56 registers
                    tmem_ld16(base, hh); tmem_ld8(base + 16, hh + 16); tmem_ld4(base + 24, hh + 24);
                    tmem_ld16(base + kN, mm); tmem_ld8(base + kN + 16, mm + 16); tmem_ld4(base + kN + 24, mm + 24);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    reg_fence16(hh); reg_fence12(hh + 16); reg_fence16(mm); reg_fence12(mm + 16);
                } else if constexpr (kUShared) {        // only the 28 live columns: 84 registers
                    tmem_ld28(base, hh);
                    tmem_ld28(base + kN, mm);
                    tmem_ld28(base + 2 * kN, ll);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    reg_fence16(hh); reg_fence12(hh + 16); reg_fence16(mm); reg_fence12(mm + 16); reg_fence16(ll); reg_fence12(ll + 16);
                } else {
                    tmem_ld16(base, hh); tmem_ld16(base + 16, hh + 16);
                    tmem_ld16(base + kN, mm); tmem_ld16(base + kN + 16, mm + 16);
                    tmem_ld16(base + 2 * kN, ll); tmem_ld16(base + 2 * kN + 16, ll + 16);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    reg_fence16(hh); reg_fence16(hh + 16); reg_fence16(mm); reg_fence16(mm + 16); reg_fence16(ll); reg_fence16(ll + 16);
                }
                // everything is in registers: tile it+5 (ring slot r+5 mod 12) may overwrite this TMEM buffer
                asm volatile("tcgen05.fence::before_thread_sync;");
                __syncwarp();
                mbar_arrive_if(&go[r + kBufs < kR ? r + kBufs : r + kBufs - kR], lane == 0);
                if constexpr (kRdHandshake) mbar_arrive_if(&rd_done[r], lane == 0);
                PROF_ADD(2);
                if constexpr (kLLCuda) {
                    constexpr unsigned kHalfLL = SO <= 15 ? 1u << (SO - 1) : 0u;   // so >= 16: the constant enters through M (K row 28)
                    const int4 u = myU[rr * kEpiWarps * 32];
                    U[0][0] = u.x; U[0][1] = u.y; U[0][2] = u.z; U[0][3] = u.w;
                    const uint4 *xlp = sXl + ((b & 1) * kBRing + r) * 8;
#pragma unroll
                    for (int sidx = 0; sidx < kSteps; ++sidx) {
                        const uint4 xv = xlp[sidx];     // same address in every lane: one broadcast read
                        const unsigned xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                        for (int px = 0; px < kPx; ++px) {
                            const int n = sidx * kPx + px;
                            U[0][px] = tc2_step<SO>(U[0][px], hh[n], mm[n], dp4a_uu(wl[sidx], xs[px], kHalfLL));
                        }
                    }
                    myU[rr * kEpiWarps * 32] = make_int4(U[0][0], U[0][1], U[0][2], U[0][3]);
                } else if constexpr (kUShared) {
                    const int4 u = myU[rr * kEpiWarps * 32];
                    U[0][0] = u.x; U[0][1] = u.y; U[0][2] = u.z; U[0][3] = u.w;
#pragma unroll
                    for (int n = 0; n < kSteps * kPx; ++n) U[0][n % kPx] = tc2_step<SO>(U[0][n % kPx], hh[n], mm[n], ll[n]);
                    myU[rr * kEpiWarps * 32] = make_int4(U[0][0], U[0][1], U[0][2], U[0][3]);
                } else {
#pragma unroll
                    for (int n = 0; n < kSteps * kPx; ++n) U[rr][n % kPx] = tc2_step<SO>(U[rr][n % kPx], hh[n], mm[n], ll[n]);
                }
                tb = tb >= kBufs - kGroups ? tb - (kBufs - kGroups) : tb + kGroups;
                PROF_ADD(3);
            }
        }
        PROF_END;
        if (m < p.OFM) {
#pragma unroll
            for (int rr = 0; rr < kTPG; ++rr) {
                int Uo[kPx];
                if constexpr (kUShared) {
                    const int4 u = myU[rr * kEpiWarps * 32];
                    Uo[0] = u.x; Uo[1] = u.y; Uo[2] = u.z; Uo[3] = u.w;
                } else {
#pragma unroll
                    for (int j = 0; j < kPx; ++j) Uo[j] = U[kUShared ? 0 : rr][j];
                }
#pragma unroll
                for (int j = 0; j < kPx; ++j) {
                    const long long gp = pix0 + (kGroups * rr + kg) * kPx + j;
                    if (gp >= npix) continue;
                    const long long grow = gp / p.W;
                    const int x = (int)(gp - grow * p.W);
                    const long long f = grow / p.H;
                    const int y = (int)(grow - f * p.H);
                    int a = Uo[j] - 32768;
                    if (p.leaky && a < 0) a = a / 10;
                    p.out[f * p.out_frame_stride + (((long long)(m >> 2) * p.H + y) * p.W + x) * 4 + (m & 3)] = (int16_t)a;
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == kIssuer) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

// Weight tiles for the tensor-core path from one layer of the reference's reorganised blob.
// Output: [mtile][kblock][row 0..127][hi plane 32 B | lo plane 32 B], the four 16-byte chunks of a row stored at
// chunk ^ ((row>>1)&3) (conflict-free LDS.128 by 32 consecutive rows); K byte 28 = the rounding constant's weight-side factor.
__global__ void wprep_tc2_kernel(const int16_t *__restrict__ blob, unsigned char *__restrict__ dst, int ifm, int ofm, int ksize,
                                 int TM, int TN, int nkb, int so, long long total)
{
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int k2 = ksize * ksize;
    int k = idx & 31;
    long long r = idx >> 5;
    int ml = r % kM; r /= kM;
    int b = r % nkb;
    int mtile = r / nkb;
    int m = mtile * kM + ml;
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

template <int KS, int SO>
void launch_one(const Tc2Params &p, dim3 grid, size_t smem, cudaStream_t st)
{
    // set on every launch (a microsecond): the attribute is per device and a host may drive several GPUs from several threads,
    // so a cached "already configured" flag would be a data race for nothing
    cudaFuncSetAttribute(conv_i16_tc2_kernel<KS, SO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    conv_i16_tc2_kernel<KS, SO><<<grid, kThreads, smem, st>>>(p);
}

template <int KS>
bool dispatch_so(const Tc2Params &p, int so, dim3 grid, size_t smem, cudaStream_t st)
{
    switch (so) {
#define Y2_TC2_CASE(S) case S: launch_one<KS, S>(p, grid, smem, st); return true;
        Y2_TC2_CASE(8) Y2_TC2_CASE(9) Y2_TC2_CASE(10) Y2_TC2_CASE(11) Y2_TC2_CASE(12) Y2_TC2_CASE(13) Y2_TC2_CASE(14) Y2_TC2_CASE(15)
        Y2_TC2_CASE(16) Y2_TC2_CASE(17) Y2_TC2_CASE(18) Y2_TC2_CASE(19) Y2_TC2_CASE(20) Y2_TC2_CASE(21) Y2_TC2_CASE(22)
#undef Y2_TC2_CASE
    default: return false;
    }
}

}  // namespace

size_t wprep_tc2_bytes(int ifm, int ofm, int ksize)
{
    const int nkb = ceil_div(ceil_div(ifm, 4) * ksize * ksize, kSteps);
    return (size_t)ceil_div(ofm, kM) * nkb * kWBytes;
}

void launch_wprep_tc2(const int16_t *blob, void *dst, int ifm, int ofm, int ksize, int TM, int TN, int so, cudaStream_t st)
{
    const int nkb = ceil_div(ceil_div(ifm, 4) * ksize * ksize, kSteps);
    const long long total = (long long)ceil_div(ofm, kM) * nkb * kM * 32;
    wprep_tc2_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(blob, (unsigned char *)dst, ifm, ofm, ksize, TM, TN, nkb, so, total);
}

// Returns 1 when launched, -1 when the shape/shift is not eligible for the tensor-core path.
int launch_conv_i16_tc2(const ConvFastParams &cp, int ksize, cudaStream_t st, const char **variant)
{
    if ((ksize != 1 && ksize != 3) || cp.so < 8 || cp.so > 22) return -1;
    Tc2Params p{};
    p.in = (const uint2 *)cp.in; p.out = (int16_t *)cp.out; p.w = (const unsigned char *)cp.w; p.bias = (const int16_t *)cp.bias;
    p.B = cp.B; p.H = cp.H; p.W = cp.W; p.G = cp.G; p.OFM = cp.OFM;
    p.in_frame_stride = cp.in_frame_stride; p.out_frame_stride = cp.out_frame_stride;
    p.sb = cp.sb; p.leaky = cp.leaky;
    p.nkb = ceil_div(cp.G * ksize * ksize, kSteps);
    p.PW = cp.W + ksize - 1;
    // 48 consecutive pixels touch at most ceil(47/W)+1 rows; + halo rows + the zero row
    p.rows_max = (kPT - 1) / cp.W + 2 + (ksize - 1) + 1;
    if (p.rows_max > 30) return -1;                     // rowinfo[] holds 32 staged rows
    p.ctab_cap = (p.rows_max - 1) * cp.W;
    const size_t fixed = (size_t)kWRing * kWBytes + (size_t)kBRing * 2 * kBBytes + kUBytes + kXlBytes + 768 + kPT * 16 + 256 + (size_t)p.ctab_cap * 8;
    const size_t per_group = (size_t)p.rows_max * p.PW * 8;
    int gs = (int)((200 * 1024 - fixed) / (2 * per_group));
    if (gs < 1) return -1;
    int sh = 0;
    while ((2 << sh) <= gs && (2 << sh) <= 16) ++sh;   // largest power of two <= min(gs, 16)
    p.gs_shift = sh;
    gs = 1 << sh;
    size_t smem = fixed + 2 * per_group * gs + 1024;
    if (smem < 120 * 1024) smem = 120 * 1024;          // one CTA per SM: a CTA allocates all 512 TMEM columns
    dim3 grid((unsigned)(((long long)cp.B * cp.H * cp.W + kPT - 1) / kPT), ceil_div(cp.OFM, kM));
    const bool ok = ksize == 3 ? dispatch_so<3>(p, cp.so, grid, smem, st) : dispatch_so<1>(p, cp.so, grid, smem, st);
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
    }
#endif
    if (variant) *variant = ksize == 3 ? "conv_i16_tc2<3>" : "conv_i16_tc2<1>";
    return 1;
}

}  // namespace y2
