// Micro-benchmark: how fast can tcgen05.mma kind::i8 (M=128, K=32) instructions with a small N be issued and
// retired, as a function of N, of the A operand's home (shared memory descriptor vs tensor memory) and of the
// number of issuing warps?  Answers whether the block-diagonal int16 conv (N = 32..64 per instruction, four
// instructions per tile) is bounded by MMA issue/retire rather than by its CUDA-core epilogue.
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
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}

template <int N, bool TS, int NW>
__global__ void __launch_bounds__(NW * 32, 1) k(long long *cycles, int groups)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ unsigned tmem_base;
    __shared__ unsigned long long bars[16];   // [warp][in-flight slot]
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<unsigned *>(smem)[i] = i * 2654435761u;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 16; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
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
    constexpr unsigned idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
    unsigned elected;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(elected));
    const unsigned long long dA = smem_desc(smem), dB = smem_desc(smem + 16 * 1024 + warp * 8192);
    // each warp writes its own accumulator columns: 2 x N columns per warp (<= 128 per warp for N = 64)
    const unsigned d0 = tmem + warp * 2 * N, a0 = tmem + 480;
    long long t0 = clock64();
    for (int g = 0; g < groups; ++g) {
        if (elected) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (TS)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d0 + (j & 1) * N),
                                 "r"(a0 + (j >> 1) * 8), "l"(dB + (unsigned long long)(j * 64)), "r"(idesc), "r"(0u)
                                 : "memory");
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d0 + (j & 1) * N),
                                 "l"(dA + (unsigned long long)((j >> 1) * 256)), "l"(dB + (unsigned long long)(j * 64)), "r"(idesc), "r"(0u)
                                 : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[warp * 4 + (g & 3)])) : "memory");
        }
        __syncwarp();
        if (g >= 3) mbar_wait(&bars[warp * 4 + ((g - 3) & 3)], ((g - 3) >> 2) & 1);   // at most 4 groups (16 instructions) in flight per warp
    }
    for (int g = groups > 3 ? groups - 3 : 0; g < groups; ++g) mbar_wait(&bars[warp * 4 + (g & 3)], (g >> 2) & 1);
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) cycles[blockIdx.x * NW + warp] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

template <int N, bool TS, int NW>
void run(int nsm, long long *cyc)
{
    const int groups = 4096;
    cudaFuncSetAttribute(k<N, TS, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
    k<N, TS, NW><<<nsm, NW * 32, 120 * 1024>>>(cyc, groups);
    k<N, TS, NW><<<nsm, NW * 32, 120 * 1024>>>(cyc, groups);
    cudaError_t err = cudaDeviceSynchronize();
    long long h[1024];
    cudaMemcpy(h, cyc, sizeof(long long) * nsm * NW, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < nsm * NW; ++i) avg += h[i];
    avg /= nsm * NW;
    printf("{\"N\": %d, \"A\": \"%s\", \"issuing_warps\": %d, \"err\": \"%s\", \"cycles_per_mma_per_warp\": %.1f, \"cycles_per_mma_per_sm\": %.1f, \"floor_N_over_2\": %d}\n", N,
           TS ? "tmem" : "smem", NW, cudaGetErrorString(err), avg / (groups * 4.0), avg / (groups * 4.0 * NW), N / 2);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int nsm = p.multiProcessorCount;
    long long *cyc; cudaMalloc(&cyc, sizeof(long long) * 1024);
    run<32, true, 1>(nsm, cyc);  run<32, true, 2>(nsm, cyc);  run<32, true, 3>(nsm, cyc);
    run<48, true, 1>(nsm, cyc);  run<48, true, 2>(nsm, cyc);
    run<64, true, 1>(nsm, cyc);  run<64, true, 2>(nsm, cyc);
    run<32, false, 1>(nsm, cyc); run<32, false, 2>(nsm, cyc);
    run<48, false, 1>(nsm, cyc); run<64, false, 1>(nsm, cyc); run<64, false, 2>(nsm, cyc);
    return 0;
}
