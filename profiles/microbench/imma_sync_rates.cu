// Micro-benchmark: is the legacy register-accumulator int8 MMA (mma.sync, SASS IMMA) fast enough on B200 to feed the
// exact round-and-saturate step WITHOUT a tensor-memory round trip?  One "pair" = a 16-channel x 4-step weight fragment against
// two block-diagonal activation fragments (2 x {HH, HL, LH, LL}) = 8 MMAs -> 8 exact steps per lane.
//   mode 0: m16n8k16 s8 MMAs only (8 per pair, accumulators discarded into a xor)
//   mode 1: m16n8k32 s8 MMAs only
//   mode 2: m16n8k16 MMAs + the 4-instruction exact step on their outputs (what a kernel would run)
//   mode 3: the 4-instruction step alone on register operands (the ALU ceiling, as tc_epilogue_rates mode 1)
//   mode 4: m16n8k32 MMAs + the step (8 steps per K block, 4 B fragments per A fragment)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define DEVFN __device__ __forceinline__
DEVFN int mad(int a, int b, int c) { int d; asm volatile("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }

#define MMA16(TA, TB, d, a, b, c)                                                                                             \
    asm volatile("mma.sync.aligned.m16n8k16.row.col.s32." TA "." TB ".s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%7,%8,%9,%10};"      \
                 : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(b), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]))
#define MMA32(TA, TB, d, a, b0, b1, c)                                                                                        \
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32." TA "." TB ".s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};" \
                 : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1),      \
                   "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]))

constexpr int NPAIR = 8;      // independent (weight fragment, pixel quad) pairs per warp = 16 chains per lane
constexpr int ITERS = 1024;
constexpr int so = 14;

DEVFN int step(int acc, int hh, int m, int ll)
{
    int t = mad(m, 256, ll);
    int a = mad(hh, 1 << (16 - so), acc);
    return __vimin_s32_relu(a + (t >> so), 65535);
}

template <int MODE>
__global__ void __launch_bounds__(256) k(int *out, int seed)
{
    int ahi[NPAIR][4], alo[NPAIR][4], bhi[NPAIR][4], blo[NPAIR][4], acc[NPAIR][2];
    const int zero[4] = {0, 0, 0, 0};
    const int half[4] = {1 << (so - 1), 1 << (so - 1), 1 << (so - 1), 1 << (so - 1)};
#pragma unroll
    for (int p = 0; p < NPAIR; ++p) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            ahi[p][j] = seed * (p + 1) + threadIdx.x * (j + 3); alo[p][j] = seed * (p + 5) - threadIdx.x * (j + 1);
            bhi[p][j] = seed * (p + 9) + threadIdx.x; blo[p][j] = seed * (p + 11) ^ threadIdx.x;
        }
        acc[p][0] = p; acc[p][1] = p + threadIdx.x;
    }
    int sink = 0;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int p = 0; p < NPAIR; ++p) {
            if (MODE == 3) {
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    acc[p][0] = step(acc[p][0], ahi[p][s], bhi[p][s], blo[p][s]);
                    acc[p][1] = step(acc[p][1], alo[p][s], blo[p][s], bhi[p][s]);
                }
            } else if (MODE == 0 || MODE == 2) {
#pragma unroll
                for (int bt = 0; bt < 2; ++bt) {
                    int hh[4], m[4], ll[4];
                    MMA16("s8", "s8", hh, ahi[p], bhi[p][bt], zero);
                    MMA16("s8", "u8", m, ahi[p], blo[p][bt], zero);
                    MMA16("u8", "s8", m, alo[p], bhi[p][bt], m);
                    MMA16("u8", "u8", ll, alo[p], blo[p][bt], half);
                    if (MODE == 0) {
                        sink ^= hh[0] ^ hh[1] ^ hh[2] ^ hh[3] ^ m[0] ^ m[1] ^ m[2] ^ m[3] ^ ll[0] ^ ll[1] ^ ll[2] ^ ll[3];
                    } else {
                        acc[p][0] = step(acc[p][0], hh[0], m[0], ll[0]);
                        acc[p][1] = step(acc[p][1], hh[2], m[2], ll[2]);
                        acc[p][0] = step(acc[p][0], hh[1], m[1], ll[1]);
                        acc[p][1] = step(acc[p][1], hh[3], m[3], ll[3]);
                    }
                }
            } else {   // MODE 1 / 4: k32, 4 B fragments (8 steps) per A fragment
#pragma unroll
                for (int bt = 0; bt < 4; ++bt) {
                    int hh[4], m[4], ll[4];
                    MMA32("s8", "s8", hh, ahi[p], bhi[p][bt], bhi[p][(bt + 1) & 3], zero);
                    MMA32("s8", "u8", m, ahi[p], blo[p][bt], blo[p][(bt + 1) & 3], zero);
                    MMA32("u8", "s8", m, alo[p], bhi[p][bt], bhi[p][(bt + 1) & 3], m);
                    MMA32("u8", "u8", ll, alo[p], blo[p][bt], blo[p][(bt + 1) & 3], half);
                    if (MODE == 1) {
                        sink ^= hh[0] ^ hh[1] ^ hh[2] ^ hh[3] ^ m[0] ^ m[1] ^ m[2] ^ m[3] ^ ll[0] ^ ll[1] ^ ll[2] ^ ll[3];
                    } else {
                        acc[p][0] = step(acc[p][0], hh[0], m[0], ll[0]);
                        acc[p][1] = step(acc[p][1], hh[2], m[2], ll[2]);
                        acc[p][0] = step(acc[p][0], hh[1], m[1], ll[1]);
                        acc[p][1] = step(acc[p][1], hh[3], m[3], ll[3]);
                    }
                }
            }
        }
#pragma unroll
        for (int p = 0; p < NPAIR; ++p)
#pragma unroll
            for (int j = 0; j < 4; ++j) asm volatile("" : "+r"(bhi[p][j]), "+r"(blo[p][j]));   // new activations every K block
    }
#pragma unroll
    for (int p = 0; p < NPAIR; ++p) sink += acc[p][0] + acc[p][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = sink;
}

template <int MODE>
void run(const char *name, int nsm, int ctas_per_sm, int *out, int REP)
{
    int grid = nsm * ctas_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, 256>>>(out, 3);
    cudaEventRecord(e0);
    for (int r = 0; r < REP; ++r) k<MODE><<<grid, 256>>>(out, 3);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= REP;
    const int steps_per_pair = (MODE == 1 || MODE == 4) ? 16 : 8;      // per lane
    const int mma_per_pair = (MODE == 3) ? 0 : (MODE == 1 || MODE == 4) ? 16 : 8;
    double lane_steps = (double)grid * 256 * ITERS * NPAIR * steps_per_pair;
    double warp_mma = (double)grid * 8 * ITERS * NPAIR * mma_per_pair;
    double sec = ms * 1e-3, clk = 1.965e9;
    printf("{\"mix\": \"%s\", \"ctas_per_sm\": %d, \"ms\": %.3f, \"T_steps_per_s\": %.3f, \"cycles_per_warp_step_per_smsp\": %.2f, "
           "\"cycles_per_mma_per_smsp\": %.2f}\n",
           name, ctas_per_sm, ms, lane_steps / sec * 1e-12, sec * clk / (lane_steps / 32 / nsm / 4),
           warp_mma > 0 ? sec * clk / (warp_mma / nsm / 4) : 0.0);
}

int main(int argc, char **argv)
{
    int REP = argc > 1 ? atoi(argv[1]) : 20;
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int nsm = p.multiProcessorCount;
    int *out;
    cudaMalloc(&out, sizeof(int) * nsm * 8 * 256);
    printf("{\"device\": \"%s\", \"sms\": %d}\n", p.name, nsm);
    for (int occ = 1; occ <= 2; ++occ) {
        run<0>("IMMA m16n8k16 only (8 per pair)", nsm, occ, out, REP);
        run<1>("IMMA m16n8k32 only (16 per pair)", nsm, occ, out, REP);
        run<3>("4-instr exact step only", nsm, occ, out, REP);
        run<2>("IMMA m16n8k16 + 4-instr exact step", nsm, occ, out, REP);
        run<4>("IMMA m16n8k32 + 4-instr exact step", nsm, occ, out, REP);
    }
    return cudaDeviceSynchronize() != cudaSuccess;
}
