// Micro-benchmark: candidate CUDA-core epilogue steps for the tcgen05 INT16 conv path. Operands (the
// three int8-plane partial sums HH, M, LL of one step) sit in registers; what is measured is the
// issue rate of the recombine + round + saturate sequence per chain step, on all SMs.
//   mode 0: 5-instr scaled step   IMAD, LEA.HI, IMAD, LOP3, VIMNMX.RELU           (2 FMA + 3 ALU)
//   mode 1: 4-instr split step    IMAD t=256M+LL, IMAD a=HH*2^(16-so)+acc, LEA.HI a+(t>>so), VIMNMX.RELU   (2 FMA + 2 ALU), so <= 16
//   mode 2: mode 1 with the HH add on the ALU pipe (LEA)                              (1 FMA + 3 ALU)
//   mode 3: IMAD.HI x4 (is the high-half multiply full rate?)
//   mode 4: 4-instr general step  IMAD c=256HH+M, LEA.HI c+=(LL>>8), IMAD.HI acc+=(c>>k), VIMNMX.RELU      (2 FMA(1 HI) + 2 ALU)
//   mode 5: float step            FFMA.RM, FADD, FFMA.RM.SAT, FMNMX                                        (3 FMA + 1 ALU)
//   mode 6: no-saturation fast step   IMAD t=256M+LL, LEA.HI acc+=(t>>so)                                  (1 FMA + 1 ALU)   [round 2]
//   mode 7: same with the recombination on the ALU pipe   LEA t=LL+(M<<8), LEA.HI acc+=(t>>so)             (2 ALU)
//   mode 8: fast step + per-7-step range check (IADD, ISETP.LE.U32 and-chained, one VOTE per 28 steps)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define DEVFN __device__ __forceinline__
DEVFN int mad(int a, int b, int c) { int d; asm volatile("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
DEVFN int madhi(int a, int b, int c) { int d; asm volatile("mad.hi.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
DEVFN float fma_rm(float a, float b, float c) { float d; asm volatile("fma.rm.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
DEVFN float fma_rm_sat(float a, float b, float c) { float d; asm volatile("fma.rm.sat.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
DEVFN float fadd(float a, float b) { float d; asm volatile("add.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }

constexpr int NCH = 16;
constexpr int ITERS = 2048;

template <int MODE>
__global__ void __launch_bounds__(256) k(int *out, int a0, int b0, int so_rt, long long *cycles)
{
    constexpr int so = 14;
    int acc[NCH], hh[NCH], mm[NCH], ll[NCH];
    float facc[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        acc[i] = i + threadIdx.x; hh[i] = a0 * (i + 1) + threadIdx.x; mm[i] = a0 * (i + 7) - threadIdx.x; ll[i] = b0 * (i + 3) + threadIdx.x;
        facc[i] = (float)acc[i] * 1e-5f;
    }
    const int k2 = so - 8, nmask = ~((1 << k2) - 1), ubound = 65535 << k2;
    const int hmul = 1 << (16 - so), himul = 1 << (32 - (so_rt - 8));
    const float s10 = 1.0f / 1024.0f, C = 12582912.0f, sk = 1.0f / (float)(1 << (k2 + 14));
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            if (MODE == 0) {
                int t1 = mad(mm[i], 256, ll[i]);
                int c = acc[i] + (t1 >> 8);
                int v = mad(hh[i], 256, c);
                acc[i] = __vimin_s32_relu(v & nmask, ubound);
            } else if (MODE == 1) {
                int t = mad(mm[i], 256, ll[i]);
                int a = mad(hh[i], hmul, acc[i]);
                acc[i] = __vimin_s32_relu(a + (t >> so), 65535);
            } else if (MODE == 2) {
                int t = mad(mm[i], 256, ll[i]);
                int a = acc[i] + (hh[i] << 2);
                acc[i] = __vimin_s32_relu(a + (t >> so), 65535);
            } else if (MODE == 3) {
                acc[i] = madhi(hh[i], himul, acc[i]); acc[i] = madhi(mm[i], himul, acc[i]);
                acc[i] = madhi(ll[i], himul, acc[i]); acc[i] = madhi(hh[i], hmul, acc[i]);
            } else if (MODE == 4) {
                int c = mad(hh[i], 256, mm[i]);
                c = c + (ll[i] >> 8);
                int a = madhi(c, himul, acc[i]);
                acc[i] = __vimin_s32_relu(a, 65535);
            } else if (MODE == 6) {
                int t = mad(mm[i], 256, ll[i]);
                acc[i] += t >> so;
            } else if (MODE == 7) {
                int t = ll[i] + (mm[i] << 8);
                acc[i] += t >> so;
            } else if (MODE == 8) {
                int t = mad(mm[i], 256, ll[i]);
                acc[i] += t >> so;
            } else if (MODE == 5) {
                float q = fma_rm(__int_as_float(ll[i]), s10, C);
                float t = fadd(__int_as_float(hh[i]), q);
                float s = fma_rm_sat(t, sk, facc[i]);
                facc[i] = fminf(s, 0.99998474f);
            }
        }
        if (MODE == 8 && (it % 7) == 6) {     // one range check per chain per 7 steps; a failing warp would take the exact path
            bool ok = true;
#pragma unroll
            for (int i = 0; i < NCH; ++i) ok = ok && (unsigned)(acc[i] - hmul) <= (unsigned)ubound;
            if (!__all_sync(0xffffffffu, ok)) { acc[0] ^= 1; }
        }
#pragma unroll
        for (int i = 0; i < NCH; ++i) asm volatile("" : "+r"(hh[i]), "+r"(mm[i]), "+r"(ll[i]));   // opaque: new partial sums every step
    }
    long long t1 = clock64();
    int s = 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) s += acc[i] + __float_as_int(facc[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int instr_per_unit, int nsm, int ctas_per_sm, int *out, long long *cyc, int REP)
{
    int grid = nsm * ctas_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, 256>>>(out, 3, 5, 14, cyc);
    cudaEventRecord(e0);
    for (int r = 0; r < REP; ++r) k<MODE><<<grid, 256>>>(out, 3, 5, 14, cyc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= REP;
    double units_per_s = (double)grid * 256 * ITERS * NCH / (ms * 1e-3);
    printf("{\"mix\": \"%s\", \"instr_per_unit\": %d, \"ctas_per_sm\": %d, \"ms\": %.3f, \"G_units_per_s\": %.1f, "
           "\"units_per_clk_per_sm_at_1965\": %.2f, \"cycles_per_warp_step_per_smsp\": %.2f}\n",
           name, instr_per_unit, ctas_per_sm, ms, units_per_s * 1e-9, units_per_s / nsm / 1.965e9, 128.0 / (units_per_s / nsm / 1.965e9));
}

int main(int argc, char **argv)
{
    int REP = argc > 1 ? atoi(argv[1]) : 40;
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int nsm = p.multiProcessorCount;
    int *out; long long *cyc;
    cudaMalloc(&out, sizeof(int) * nsm * 8 * 256); cudaMalloc(&cyc, sizeof(long long) * 4096);
    printf("{\"device\": \"%s\", \"sms\": %d}\n", p.name, nsm);
    for (int occ = 2; occ <= 4; occ += 2) {
        run<0>("tc step v1: IMAD, LEA.HI, IMAD, LOP3, VIMNMX.RELU", 5, nsm, occ, out, cyc, REP);
        run<1>("tc step v2 (so<=16): IMAD, IMAD, LEA.HI, VIMNMX.RELU", 4, nsm, occ, out, cyc, REP);
        run<2>("tc step v2b: IMAD, LEA, LEA.HI, VIMNMX.RELU", 4, nsm, occ, out, cyc, REP);
        run<3>("IMAD.HI x4", 4, nsm, occ, out, cyc, REP);
        run<4>("tc step v3 (general): IMAD, LEA.HI, IMAD.HI, VIMNMX.RELU", 4, nsm, occ, out, cyc, REP);
        run<5>("float step: FFMA.RM, FADD, FFMA.RM.SAT, FMNMX", 4, nsm, occ, out, cyc, REP);
        run<6>("fast step (no saturation possible): IMAD, LEA.HI", 2, nsm, occ, out, cyc, REP);
        run<7>("fast step, ALU only: LEA, LEA.HI", 2, nsm, occ, out, cyc, REP);
        run<8>("fast step + range check per 7 steps", 2, nsm, occ, out, cyc, REP);
    }
    return cudaDeviceSynchronize() != cudaSuccess;
}
