// Micro-benchmark: tcgen05.ld (TMEM -> registers) bandwidth per SM as a function of the number of
// warps reading, to decide whether a tcgen05 path whose CUDA-core epilogue must read three int32
// partial sums per round-and-saturate step can be fed from TMEM.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

template <int NW>
__global__ void __launch_bounds__(NW * 32, 1) k(int *out, long long *cycles, int iters)
{
    __shared__ unsigned tmem_base;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        unsigned dst = (unsigned)__cvta_generic_to_shared(&tmem_base);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(dst));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const unsigned base = tmem_base + (((unsigned)(warp & 3) * 32u) << 16);
    int acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            unsigned r[32];
            unsigned addr = base + (unsigned)(((it * 4 + c) * 32 + (warp >> 2) * 64) & 511 & ~31);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                  "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                  "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                  "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(addr));
            asm volatile("tcgen05.wait::ld.sync.aligned;");
            acc ^= r[0] ^ r[13] ^ r[31];
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
}

template <int NW>
void run(int nsm, int *out, long long *cyc)
{
    const int iters = 4096;
    k<NW><<<nsm, NW * 32>>>(out, cyc, iters);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<NW><<<nsm, NW * 32>>>(out, cyc, iters);
    cudaEventRecord(e1);
    cudaError_t err = cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[256]; cudaMemcpy(h, cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < nsm; ++i) avg += h[i]; avg /= nsm;
    double bytes_per_sm = (double)NW * iters * 4 * 32 * 32 * 4;
    printf("{\"warps\": %d, \"err\": \"%s\", \"ms\": %.3f, \"cycles\": %.0f, \"tmem_ld_bytes_per_clk_per_sm\": %.1f, \"GBps_per_sm\": %.1f}\n",
           NW, cudaGetErrorString(err), ms, avg, bytes_per_sm / avg, bytes_per_sm / (ms * 1e-3) / 1e9);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int nsm = p.multiProcessorCount;
    int *out; long long *cyc;
    cudaMalloc(&out, sizeof(int) * nsm * 1024); cudaMalloc(&cyc, sizeof(long long) * 256);
    run<4>(nsm, out, cyc);
    run<8>(nsm, out, cyc);
    run<16>(nsm, out, cyc);
    return cudaDeviceSynchronize() != cudaSuccess;
}
