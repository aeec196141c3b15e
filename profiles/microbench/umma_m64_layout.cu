// Layout probe: where does a tcgen05.mma kind::i8 with M = 64 (cta_group::1) put its 64 accumulator rows in tensor memory,
// can the D address carry a lane offset of 16 (so that TWO M = 64 products fill all 128 lanes), and where does it read the
// rows of an A operand held in tensor memory?  A[m][0] = m + 1, B[n][0] = n + 1  =>  D[m][n] = (m + 1) * (n + 1).
// Output: per (experiment, TMEM lane) the value found in column 0 and column 1 (row = value of column 0 - 1).
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
__host__ __device__ inline int operand_off(int row, int k) { return (((row >> 3) * 2 + (k >> 4)) * 8 + (row & 7)) * 16 + (k & 15); }
__device__ __forceinline__ void mbar_wait(void *bar, unsigned parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
constexpr int kN = 16;
__host__ __device__ constexpr unsigned idesc(int M) { return (2u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(kN >> 3) << 17) | ((unsigned)(M >> 4) << 24); }

// exp 0: SS, M=64, D lane 0.  exp 1: SS, M=64, D lane 16.  exp 2: TS, M=64, A rows stored like exp 0's D rows, D lane 0.
// exp 3: TS, M=64, D lane 16, A at lane offset 16 too.  exp 4: TS, M=64, D lane 16, A at lane offset 0.  exp 5: SS M=128 reference.
__global__ void __launch_bounds__(128, 1) probe(int *out, int exp, const int *rowlane)
{
    __shared__ __align__(1024) unsigned char sA[4096];
    __shared__ __align__(1024) unsigned char sB[1024];
    __shared__ unsigned tmem_base;
    __shared__ unsigned long long bar;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 4096; i += 128) sA[i] = 0;
    for (int i = threadIdx.x; i < 1024; i += 128) sB[i] = 0;
    __syncthreads();
    if (threadIdx.x < 128) sA[operand_off(threadIdx.x, 0)] = (unsigned char)(threadIdx.x < 64 || exp == 5 ? (threadIdx.x % 100) + 1 : 0);
    if (threadIdx.x < kN) sB[operand_off(threadIdx.x, 0)] = (unsigned char)(threadIdx.x + 1);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const unsigned tmem = tmem_base;
    const unsigned mylanes = tmem + ((unsigned)(warp * 32) << 16);
    // poison D (columns 0..15) with -1, zero the A region (columns 32..39)
    {
        int m1 = -1, z = 0;
        for (int c = 0; c < 16; c += 4)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%1,%1,%1};" ::"r"(mylanes + c), "r"(m1) : "memory");
        for (int c = 32; c < 40; c += 4)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%1,%1,%1};" ::"r"(mylanes + c), "r"(z) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    if (exp >= 2 && exp <= 4) {
        // A operand in TMEM: lane L holds row rowlane-inverse: the host passes rowlane[r] = TMEM lane of row r (from exp 0)
        // K = 32 int8 = 8 columns; only k = 0 is non-zero: byte 0 of column 0
        const int aoff = (exp == 3) ? 16 : 0;
        for (int r = 0; r < 64; ++r) {
            const int L = rowlane[r] + aoff;
            if (L == warp * 32 + lane) { /* marker */ }
        }
        int v = 0;
        for (int r = 0; r < 64; ++r) if (rowlane[r] + aoff == warp * 32 + lane) v = r + 1;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(mylanes + 32), "r"(v) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (threadIdx.x == 0) {
        const unsigned dlane = (exp == 1 || exp == 3 || exp == 4) ? 16u : 0u;
        const unsigned d = tmem + (dlane << 16);
        const unsigned long long dB = smem_desc(sB), dA = smem_desc(sA);
        const int M = exp == 5 ? 128 : 64;
        if (exp >= 2 && exp <= 4) {
            const unsigned a = tmem + 32 + ((exp == 3 ? 16u : 0u) << 16);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(dB), "r"(idesc(M)) : "memory");
        } else {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(dA), "l"(dB), "r"(idesc(M)) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;");
    int r0, r1, r2, r3;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(mylanes));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    out[threadIdx.x * 2] = r0;
    out[threadIdx.x * 2 + 1] = r1;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem));
}

int main()
{
    int *d_out, *d_rl, h[256], rowlane[64];
    cudaMalloc(&d_out, sizeof(h));
    cudaMalloc(&d_rl, sizeof(rowlane));
    for (int r = 0; r < 64; ++r) rowlane[r] = (r / 16) * 32 + r % 16;     // guess, replaced by what exp 0 shows
    for (int exp = 0; exp <= 5; ++exp) {
        cudaMemcpy(d_rl, rowlane, sizeof(rowlane), cudaMemcpyHostToDevice);
        cudaMemset(d_out, 0, sizeof(h));
        probe<<<1, 128>>>(d_out, exp, d_rl);
        cudaError_t e = cudaDeviceSynchronize();
        printf("{\"exp\": %d, \"err\": \"%s\", \"lanes\": [", exp, cudaGetErrorString(e));
        if (e != cudaSuccess) { printf("]}\n"); return 1; }
        cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
        for (int l = 0; l < 128; ++l) printf("%s[%d,%d]", l ? "," : "", h[2 * l], h[2 * l + 1]);
        printf("]}\n");
        if (exp == 0) {
            int found = 0;
            for (int l = 0; l < 128; ++l)
                if (h[2 * l] >= 1 && h[2 * l] <= 64 && h[2 * l + 1] == 2 * h[2 * l]) { rowlane[h[2 * l] - 1] = l; ++found; }
            printf("{\"exp0_rows_found\": %d}\n", found);
        }
    }
    return 0;
}
