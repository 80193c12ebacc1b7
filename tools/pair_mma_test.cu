// Micro-test of tcgen05.mma.cta_group::2 (CTA pair, M = 256): which half of the B operand each CTA of the pair supplies,
// and that descriptors / commit multicast / TMEM layout behave as cnn_k9 / cnn_a20 would need.  Not product code.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/bin/pair_mma_test tools/pair_mma_test.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major SWIZZLE_NONE descriptor: core matrix = 8 rows x 16 B (128 B contiguous); LBO = bytes between the two K chunks,
// SBO = bytes between 8-row groups
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
           (1ull << 46);
}

template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) pair_kernel(const __half *A, const __half *B, float *D, int b_split_mode) {
    // A: [256][16] row-major global; B: [N][16] row-major global (D = A * B^T); D: [256][N]
    extern __shared__ __align__(1024) unsigned char smem[];
    __half *sa = reinterpret_cast<__half *>(smem);                 // 128 rows: planes [2 K chunks][128 rows][8 halfs]
    __half *sb = reinterpret_cast<__half *>(smem + 8192);          // N/2 rows:  planes [2][N/2][8]
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 16384);
    uint32_t *tm = reinterpret_cast<uint32_t *>(smem + 16400);
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int tid = threadIdx.x;
    for (int i = tid; i < 128 * 16; i += 128) {
        const int r = i / 16, k = i % 16;
        sa[(k / 8) * 128 * 8 + r * 8 + (k % 8)] = A[(rank * 128 + r) * 16 + k];
    }
    for (int i = tid; i < (N / 2) * 16; i += 128) {
        const int r = i / 16, k = i % 16;
        const int n = b_split_mode == 0 ? (int)rank * (N / 2) + r : (1 - (int)rank) * (N / 2) + r;
        sb[(k / 8) * (N / 2) * 8 + r * 8 + (k % 8)] = B[n * 16 + k];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tm)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tm;
    if (rank == 0 && tid == 0) {
        const uint64_t ad = make_desc(smem_u32(sa), 128 * 16, 128);
        const uint64_t bd = make_desc(smem_u32(sb), (N / 2) * 16, 128);
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((256u >> 4) << 24);
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
    }
    // every CTA waits on its own barrier (the commit is multicast to both)
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(bar)), "r"(0u) : "memory");
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int warp = tid >> 5, lane = tid & 31;
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                       "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) D[(size_t)(rank * 128 + warp * 32 + lane) * N + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

template <int N>
static void run(int mode) {
    std::vector<__half> A(256 * 16), B(N * 16);
    std::vector<float> Af(256 * 16), Bf(N * 16);
    for (int i = 0; i < 256 * 16; ++i) { Af[i] = (float)((i * 7) % 13 - 6); A[i] = __float2half(Af[i]); }
    for (int i = 0; i < N * 16; ++i) { Bf[i] = (float)((i * 5) % 11 - 5); B[i] = __float2half(Bf[i]); }
    __half *dA, *dB;
    float *dD;
    cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, 256 * N * 4);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, 256 * N * 4);
    cudaFuncSetAttribute(pair_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    pair_kernel<N><<<2, 128, 32768>>>(dA, dB, dD, mode);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> D(256 * N);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0, bad_lo = 0, bad_hi = 0;
    for (int r = 0; r < 256; ++r)
        for (int n = 0; n < N; ++n) {
            float ref = 0;
            for (int k = 0; k < 16; ++k) ref += Af[r * 16 + k] * Bf[n * 16 + k];
            if (D[r * N + n] != ref) { ++bad; (r < 128 ? bad_lo : bad_hi)++; }
        }
    printf("N=%d b_split_mode=%d: %s, mismatches %d of %d (rows 0-127: %d, rows 128-255: %d)  D[0][0]=%g D[0][N/2]=%g D[128][0]=%g\n", N, mode,
           cudaGetErrorString(e), bad, 256 * N, bad_lo, bad_hi, D[0], D[N / 2], D[128 * N]);
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
}

int main() {
    run<64>(0);
    run<64>(1);
    run<32>(0);
    run<32>(1);
    return 0;
}
