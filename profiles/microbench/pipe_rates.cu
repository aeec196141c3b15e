// Micro-benchmark: issue rates of the integer instructions the exact INT16 conv step is made of
// (IDP.2A, SHF, VIADDMNMX.RELU, IMAD) and of the 7-instruction step itself, on all SMs.
// Output: warp-instructions per clock per SM for each mix -> the ALU-issue roofline of the
// bit-exact datapath (DESIGN.md "exactness-adjusted roofline").
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define DEVFN __device__ __forceinline__
DEVFN int dp2a_lo_su(int a, unsigned b, int c) { int d; asm volatile("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
DEVFN int dp2a_hi_su(int a, unsigned b, int c) { int d; asm volatile("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
DEVFN int dp2a_lo_ss(int a, int b, int c) { int d; asm volatile("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
DEVFN int dp2a_hi_ss(int a, int b, int c) { int d; asm volatile("dp2a.hi.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
DEVFN int shr(int a, int k) { int d; asm volatile("shr.s32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(k)); return d; }
DEVFN int mad(int a, int b, int c) { int d; asm volatile("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }

constexpr int NCH = 16;   // independent chains per thread
constexpr int ITERS = 2048;

__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

template <int MODE>
__global__ void __launch_bounds__(256) k(int *out, int a0, int b0, int k2, long long *cycles, long long *nanos)
{
    int acc[NCH], xa[NCH], xb[NCH], x0 = a0 + threadIdx.x, x1 = a0 * 3 + threadIdx.x, w0 = b0, w1 = b0 ^ 0x55aa;
#pragma unroll
    for (int i = 0; i < NCH; ++i) { acc[i] = i + threadIdx.x; xa[i] = a0 * (i + 1) + threadIdx.x; xb[i] = a0 * (i + 7) - threadIdx.x; }
    unsigned long long g0 = gtimer();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            if (MODE == 0) {            // IDP.2A only (4 per step)
                int p = dp2a_lo_su(x0, w0, acc[i]); p = dp2a_hi_su(x1, w0, p); p = dp2a_lo_ss(x0, w1, p); acc[i] = dp2a_hi_ss(x1, w1, p);
            } else if (MODE == 1) {     // SHF only
                acc[i] = shr(acc[i] + 0, k2); acc[i] = shr(acc[i], k2); acc[i] = shr(acc[i], k2); acc[i] = shr(acc[i], k2);
            } else if (MODE == 2) {     // VIADDMNMX.RELU only
                acc[i] = __viaddmin_s32_relu(acc[i], x0, 65535); acc[i] = __viaddmin_s32_relu(acc[i], x1, 65535);
                acc[i] = __viaddmin_s32_relu(acc[i], w0, 65535); acc[i] = __viaddmin_s32_relu(acc[i], w1, 65535);
            } else if (MODE == 3) {     // IMAD only
                acc[i] = mad(x0, w0, acc[i]); acc[i] = mad(x1, w1, acc[i]); acc[i] = mad(x0, w1, acc[i]); acc[i] = mad(x1, w0, acc[i]);
            } else if (MODE == 4) {     // the 7-instruction exact step
                int plo = dp2a_lo_su(xa[i], w0, 8192); plo = dp2a_hi_su(xb[i], w0, plo);
                int phi = dp2a_lo_ss(xa[i], w1, shr(plo, 8)); phi = dp2a_hi_ss(xb[i], w1, phi);
                acc[i] = __viaddmin_s32_relu(acc[i], shr(phi, k2), 65535);
            } else if (MODE == 6) {     // VIMNMX.RELU only
                acc[i] = __vimin_s32_relu(acc[i] ^ x0, 65535); acc[i] = __vimin_s32_relu(acc[i] ^ x1, 65535);
                acc[i] = __vimin_s32_relu(acc[i] ^ w0, 65535); acc[i] = __vimin_s32_relu(acc[i] ^ w1, 65535);
            } else if (MODE == 7) {     // scaled-state exact step: 4 IDP.2A + LEA.HI.SX32 + LOP3 + VIMNMX.RELU
                int plo = dp2a_lo_su(xa[i], w0, 8192); plo = dp2a_hi_su(xb[i], w0, plo);
                int phi = dp2a_lo_ss(xa[i], w1, acc[i] + (plo >> 8)); phi = dp2a_hi_ss(xb[i], w1, phi);
                acc[i] = __vimin_s32_relu(phi & k2, 65535 << 6);
            } else if (MODE == 5) {     // 6-instruction IMAD step (32-bit products, single shift)
                int p = mad(xa[i], w0, 8192); p = mad(xb[i], w1, p); p = mad(xa[i], w1, p); p = mad(xb[i], w0, p);
                acc[i] = __viaddmin_s32_relu(acc[i], shr(p, k2), 65535);
            }
        }
        x0 += it; w1 ^= it; w0 += 3;
    }
    long long t1 = clock64();
    unsigned long long g1 = gtimer();
    int s = 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) { cycles[blockIdx.x] = t1 - t0; nanos[blockIdx.x] = (long long)(g1 - g0); }
}

template <int MODE>
void run(const char *name, int instr_per_unit, int nsm, int ctas_per_sm, int *out, long long *cyc, long long *ns, int REP)
{
    int grid = nsm * ctas_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, 256>>>(out, 3, 5, 6, cyc, ns);
    cudaEventRecord(e0);
    for (int r = 0; r < REP; ++r) k<MODE><<<grid, 256>>>(out, 3, 5, 6, cyc, ns);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= REP;
    long long h[4096]; cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
    long long hn[4096]; cudaMemcpy(hn, ns, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    double avgn = 0; for (int i = 0; i < grid; ++i) avgn += hn[i]; avgn /= grid;
    double warp_instr_per_sm = (double)ctas_per_sm * 8 * ITERS * NCH * instr_per_unit;
    double units_per_s = (double)grid * 256 * ITERS * NCH / (ms * 1e-3);
    printf("{\"mix\": \"%s\", \"instr_per_unit\": %d, \"ctas_per_sm\": %d, \"ms\": %.3f, \"cycles\": %.0f, "
           "\"warp_instr_per_clk_per_sm\": %.3f, \"units_per_clk_per_sm\": %.3f, \"G_units_per_s\": %.1f, \"clk_mhz_eff\": %.0f, \"sm_clk_mhz_globaltimer\": %.0f}\n",
           name, instr_per_unit, ctas_per_sm, ms, avg, warp_instr_per_sm / avg, warp_instr_per_sm / avg * 32 / instr_per_unit,
           units_per_s * 1e-9, avg / (ms * 1e-3) * 1e-6, avg / avgn * 1e3);
}

int main(int argc, char **argv)
{
    int REP = argc > 1 ? atoi(argv[1]) : 40;
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int nsm = p.multiProcessorCount;
    int *out; long long *cyc, *ns;
    cudaMalloc(&out, sizeof(int) * nsm * 8 * 256); cudaMalloc(&cyc, sizeof(long long) * 4096); cudaMalloc(&ns, sizeof(long long) * 4096);
    printf("{\"device\": \"%s\", \"sms\": %d}\n", p.name, nsm);
    for (int occ = 2; occ <= 4; occ += 2) {
        run<0>("IDP.2A x4", 4, nsm, occ, out, cyc, ns, REP);
        run<1>("SHF x4", 4, nsm, occ, out, cyc, ns, REP);
        run<2>("VIADDMNMX.RELU x4", 4, nsm, occ, out, cyc, ns, REP);
        run<3>("IMAD x4", 4, nsm, occ, out, cyc, ns, REP);
        run<4>("exact step: 4 IDP.2A + 2 SHF + 1 VIADDMNMX", 7, nsm, occ, out, cyc, ns, REP);
        run<5>("imad step: 4 IMAD + 1 SHF + 1 VIADDMNMX", 6, nsm, occ, out, cyc, ns, REP);
        run<6>("LOP3 + VIMNMX.RELU x4", 8, nsm, occ, out, cyc, ns, REP);
        run<7>("scaled exact step: 4 IDP.2A + LEA.HI.SX32 + LOP3 + VIMNMX.RELU", 7, nsm, occ, out, cyc, ns, REP);
    }
    return cudaDeviceSynchronize() != cudaSuccess;
}
