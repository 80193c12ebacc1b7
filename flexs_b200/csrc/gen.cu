// K5 candidate-generation helpers: batched random mutation and per-position argmax decode.
// Both are HBM-bound byte/float streaming kernels (1 B in + 1 B out per residue for mutate;
// 4*row_stride B in + 1 B out per residue for decode).
#include <algorithm>

#include "common.cuh"

namespace {

// Philox-4x32-10 (Salmon et al., SC'11), the counter-based generator also used by cuRAND/torch.
struct Philox {
    uint32_t key[2];
    __device__ __forceinline__ static void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    }
    __device__ __forceinline__ void operator()(uint64_t counter, uint64_t subsequence, uint32_t (&out)[4]) const {
        uint32_t c[4] = {(uint32_t)counter, (uint32_t)(counter >> 32), (uint32_t)subsequence,
                         (uint32_t)(subsequence >> 32)};
        uint32_t k0 = key[0], k1 = key[1];
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            round(c, k0, k1);
            k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
        }
        out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
    }
};

// generate_random_mutant (sequence_utils.py:87-108): per residue, `random.random() < mu` then
// `random.choice(alphabet)` (uniform over ALL residues, may re-draw the same one).
__global__ void __launch_bounds__(256) mutate_kernel(const uint8_t *__restrict__ parents, int64_t total,
                                                     int alphabet_size, float mu, uint64_t seed,
                                                     uint64_t subsequence, uint8_t *__restrict__ children) {
    Philox rng;
    rng.key[0] = (uint32_t)seed; rng.key[1] = (uint32_t)(seed >> 32);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    // one Philox call serves two residues: (u_mutate, u_choice) x 2
    for (int64_t pair = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pair * 2 < total; pair += stride) {
        uint32_t r[4];
        rng((uint64_t)pair, subsequence, r);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t pos = pair * 2 + h;
            if (pos >= total) break;
            uint8_t v = parents[pos];
            const float u = (float)(r[2 * h] >> 8) * (1.0f / 16777216.0f);  // uniform [0,1), 24 bits
            if (u < mu) v = (uint8_t)(((uint64_t)r[2 * h + 1] * (uint64_t)alphabet_size) >> 32);
            children[pos] = v;
        }
    }
}

// np.argmax(x, axis=-1): first maximum wins (cmaes.py:61-67, environments/dyna_ppo.py:144-147)
__global__ void __launch_bounds__(256) argmax_kernel(const float *__restrict__ x, int64_t rows, int row_stride,
                                                     int alphabet_size, uint8_t *__restrict__ idx) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += stride) {
        const float *p = x + r * row_stride;
        float best = __ldg(p);
        int bi = 0;
        for (int a = 1; a < alphabet_size; ++a) {
            const float v = __ldg(p + a);
            // numpy treats NaN as the maximum (first NaN wins)
            if (best == best && (v > best || v != v)) { best = v; bi = a; }
        }
        idx[r] = (uint8_t)bi;
    }
}

int grid_for(int64_t work) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return (int)std::max<int64_t>(1, std::min<int64_t>((work + 255) / 256, (int64_t)sms * 8));
}

}  // namespace

extern "C" {

int flexs_mutate_dev(const uint8_t *d_parents, int64_t n, int seq_len, int alphabet_size, float mu,
                     uint64_t seed, uint64_t subsequence, uint8_t *d_children, void *stream) {
    FX_REQUIRE(n >= 0 && seq_len >= 1, "bad sizes");
    FX_REQUIRE(alphabet_size >= 1 && alphabet_size <= 255, "bad alphabet size");
    if (n == 0) return FLEXS_OK;
    FX_REQUIRE(d_parents && d_children, "null buffer");
    const int64_t total = n * seq_len;
    mutate_kernel<<<grid_for((total + 1) / 2), 256, 0, (cudaStream_t)stream>>>(d_parents, total, alphabet_size, mu,
                                                                                 seed, subsequence, d_children);
    FX_CUDA(cudaGetLastError());
    return FLEXS_OK;
}

int flexs_argmax_decode_dev(const float *d_x, int64_t n, int seq_len, int row_stride, int alphabet_size,
                            uint8_t *d_idx, void *stream) {
    FX_REQUIRE(n >= 0 && seq_len >= 1, "bad sizes");
    FX_REQUIRE(alphabet_size >= 1 && alphabet_size <= 255 && row_stride >= alphabet_size, "bad alphabet size / stride");
    if (n == 0) return FLEXS_OK;
    FX_REQUIRE(d_x && d_idx, "null buffer");
    const int64_t rows = n * seq_len;
    argmax_kernel<<<grid_for(rows), 256, 0, (cudaStream_t)stream>>>(d_x, rows, row_stride, alphabet_size, d_idx);
    FX_CUDA(cudaGetLastError());
    return FLEXS_OK;
}

}  // extern "C"
