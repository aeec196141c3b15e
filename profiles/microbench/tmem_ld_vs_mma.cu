// Micro-benchmark: latency of a tile read-out (6 x tcgen05.ld.32x32b.x16 + tcgen05.wait::ld, what one epilogue warp of the
// int16 tcgen05 conv does per tile) while 0..4 other warps keep tcgen05.mma kind::i8 (M=128, N=32, A in TMEM) queued.
// Answers: does the TMEM read path wait behind queued MMAs?
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long smem_desc(const void *p)
{
    unsigned long long d = (unsigned long long)((smem_u32(p) >> 4) & 0x3FFF);
    d |= (unsigned long long)(128 >> 4) << 16;
    d |= (unsigned long long)(256 >> 4) << 32;
    d |= 1ull << 46;
    return d;
}
__device__ __forceinline__ void mbar_wait(void *bar, unsigned parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_ld16(unsigned taddr, int *r)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}

// warps 0..NR-1: readers (NR = 4, 8 or 12; quadrant = warp % 4); warps NR..NR+NI-1: MMA issuers with QD groups of 4 MMAs in flight each
template <int NR, int NI, int QD>
__global__ void __launch_bounds__((NR + NI) * 32, 1) k(long long *cycles, int *sink, int iters, long long *mma_groups)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ unsigned tmem_base;
    __shared__ unsigned long long bars[32];
    __shared__ volatile int stop;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<unsigned *>(smem)[i] = i * 2654435761u;
    if (threadIdx.x == 0) {
        stop = 0;
        for (int i = 0; i < 32; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const unsigned tmem = tmem_base;
    if (warp < NR) {
        const unsigned base = tmem + ((unsigned)((warp & 3) * 32) << 16) + (warp >> 2) * 96;   // reads columns 0..287, MMAs write 288..479
        int acc = 0;
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            int a[16], b[16], c[16], d[16], e[16], f[16];
            tmem_ld16(base, a); tmem_ld16(base + 16, b); tmem_ld16(base + 32, c); tmem_ld16(base + 48, d); tmem_ld16(base + 64, e); tmem_ld16(base + 80, f);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc ^= a[0] ^ b[3] ^ c[5] ^ d[7] ^ e[11] ^ f[15];
        }
        long long t1 = clock64();
        if (lane == 0) cycles[blockIdx.x * NR + warp] = t1 - t0;
        sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
        __syncwarp();
        if (warp == 0 && lane == 0) stop = 1;
    } else {
        const int iw = warp - NR;
        constexpr unsigned idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(32 >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
        unsigned elected;
        asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(elected));
        const unsigned long long dB = smem_desc(smem + 16 * 1024 + iw * 4096);
        const unsigned d0 = tmem + 288 + iw * 48, a0 = tmem + 480;
        for (int g = 0; !stop; ++g) {
            if (elected) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d0 + (j & 1) * 16),
                                 "r"(a0 + (j >> 1) * 8), "l"(dB + (unsigned long long)(j * 64)), "r"(idesc), "r"(0u)
                                 : "memory");
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[iw * 8 + (g & 7)])) : "memory");
            }
            __syncwarp();
            if (g >= QD - 1) mbar_wait(&bars[iw * 8 + ((g - (QD - 1)) & 7)], ((g - (QD - 1)) >> 3) & 1);
            if (lane == 0) mma_groups[blockIdx.x * 4 + iw] = g + 1;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

template <int NR, int NI, int QD>
void run(int nsm, long long *cyc, int *sink)
{
    const int iters = 2000;
    cudaFuncSetAttribute(k<NR, NI, QD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
    long long *mg; cudaMalloc(&mg, sizeof(long long) * 4 * nsm); cudaMemset(mg, 0, sizeof(long long) * 4 * nsm);
    k<NR, NI, QD><<<nsm, (NR + NI) * 32, 120 * 1024>>>(cyc, sink, iters, mg);
    cudaError_t err = cudaDeviceSynchronize();
    long long h[2048];
    cudaMemcpy(h, cyc, sizeof(long long) * nsm * NR, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < nsm * NR; ++i) avg += h[i];
    avg /= nsm * NR;
    long long hg[1024]; cudaMemcpy(hg, mg, sizeof(long long) * 4 * nsm, cudaMemcpyDeviceToHost);
    double groups = 0; for (int i = 0; i < 4 * nsm; ++i) groups += hg[i];
    groups /= nsm;   // 4-MMA groups per SM during the readers' run
    printf("{\"reader_warps\": %d, \"mma_warps\": %d, \"mma_groups_in_flight_per_warp\": %d, \"err\": \"%s\", \"cycles_per_tile_readout\": %.1f, \"cycles_per_mma_per_sm\": %.1f}\n", NR, NI, QD,
           cudaGetErrorString(err), avg / iters, groups > 0 ? avg / (groups * 4) : 0.0);
    cudaFree(mg);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int nsm = p.multiProcessorCount;
    long long *cyc; int *sink;
    cudaMalloc(&cyc, sizeof(long long) * 2048); cudaMalloc(&sink, sizeof(int) * nsm * 1024);
    run<4, 0, 1>(nsm, cyc, sink); run<12, 0, 1>(nsm, cyc, sink);
    run<4, 1, 1>(nsm, cyc, sink); run<4, 1, 4>(nsm, cyc, sink);
    run<4, 4, 1>(nsm, cyc, sink); run<4, 4, 2>(nsm, cyc, sink);
    run<12, 4, 1>(nsm, cyc, sink); run<12, 4, 2>(nsm, cyc, sink); run<12, 2, 1>(nsm, cyc, sink); run<8, 2, 2>(nsm, cyc, sink); run<8, 4, 2>(nsm, cyc, sink);
    return 0;
}
