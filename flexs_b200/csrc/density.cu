// K8: sequence-density penalty of the DyNA-PPO environment (environments/dyna_ppo.py:106-114, :155-160):
//     density(s) = sum over previously seen sequences o with 0 < d(s, o) <= radius of fitness(o) / d(s, o),
// d = Levenshtein distance (editdistance.eval in the reference).  The reference walks a Python dict of every sequence
// seen so far for every sequence of a batch; here one CTA owns a new sequence (staged in shared memory), its threads
// sweep the seen set, each pair runs the banded dynamic programme (band |i - j| <= radius, 2 * radius + 1 cells per
// row) and leaves as soon as the whole band exceeds the radius — a handful of rows for unrelated sequences.  The per-
// sequence sum is reduced in a fixed order in double precision: deterministic, no atomics.
#include <algorithm>

#include "common.cuh"

namespace {

constexpr int NT = 256;
constexpr int MAXR = 4;

template <int R>
__device__ __forceinline__ int banded_distance(const uint8_t *__restrict__ a, const uint8_t *__restrict__ b, int L) {
    // prev[d + R] = D[i-1][i-1 + d], d in [-R, R]; equal lengths, so the answer is D[L][L] = band centre
    int prev[2 * R + 1], cur[2 * R + 1];
#pragma unroll
    for (int d = 0; d <= 2 * R; ++d) prev[d] = (d >= R) ? d - R : R + 1;   // row 0: D[0][j] = j
    for (int i = 1; i <= L; ++i) {
        const int ca = a[i - 1];
        int best = R + 1;
#pragma unroll
        for (int d = 0; d <= 2 * R; ++d) {
            const int j = i + d - R;
            int v = R + 1;
            if (j == 0) v = min(i, R + 1);
            else if (j > 0 && j <= L) {
                const int sub = prev[d] + (ca != (int)b[j - 1]);                 // D[i-1][j-1]: same diagonal
                const int del = (d + 1 <= 2 * R) ? prev[d + 1] + 1 : R + 1;      // D[i-1][j]
                const int ins = (d >= 1) ? cur[d - 1] + 1 : R + 1;               // D[i][j-1]
                v = min(min(sub, del), min(ins, R + 1));
            }
            cur[d] = v;
            best = min(best, v);
        }
        if (best > R) return R + 1;
#pragma unroll
        for (int d = 0; d <= 2 * R; ++d) prev[d] = cur[d];
    }
    return prev[R];
}

template <int R>
__global__ void __launch_bounds__(NT) density_kernel(const uint8_t *__restrict__ fresh, int64_t n_new,
                                                     const uint8_t *__restrict__ seen, const double *__restrict__ fit,
                                                     int64_t n_seen, int L, double *__restrict__ out) {
    extern __shared__ unsigned char s_row[];
    __shared__ double s_part[NT];
    for (int64_t i = blockIdx.x; i < n_new; i += gridDim.x) {
        __syncthreads();
        for (int t = threadIdx.x; t < L; t += NT) s_row[t] = fresh[i * L + t];
        __syncthreads();
        double acc = 0.0;
        for (int64_t j = threadIdx.x; j < n_seen; j += NT) {
            const int d = banded_distance<R>(s_row, seen + j * L, L);
            if (d != 0 && d <= R) acc += fit[j] / (double)d;
        }
        s_part[threadIdx.x] = acc;
        __syncthreads();
        for (int off = NT / 2; off > 0; off >>= 1) {
            if (threadIdx.x < off) s_part[threadIdx.x] += s_part[threadIdx.x + off];
            __syncthreads();
        }
        if (threadIdx.x == 0) out[i] = s_part[0];
    }
}

}  // namespace

extern "C" {

int flexs_edit_density_dev(const uint8_t *d_new, int64_t n_new, const uint8_t *d_seen, const double *d_seen_fitness,
                           int64_t n_seen, int seq_len, int radius, double *d_out, void *stream) {
    FX_REQUIRE(n_new >= 0 && n_seen >= 0 && seq_len >= 1, "bad sizes");
    FX_REQUIRE(radius >= 1 && radius <= MAXR, "radius must be in [1, 4]");
    if (n_new == 0) return FLEXS_OK;
    FX_REQUIRE(d_new && d_out && (n_seen == 0 || (d_seen && d_seen_fitness)), "null buffer");
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = (int)std::min<int64_t>(n_new, (int64_t)sms * 8);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t smem = (size_t)seq_len;
    switch (radius) {
        case 1: density_kernel<1><<<grid, NT, smem, s>>>(d_new, n_new, d_seen, d_seen_fitness, n_seen, seq_len, d_out); break;
        case 2: density_kernel<2><<<grid, NT, smem, s>>>(d_new, n_new, d_seen, d_seen_fitness, n_seen, seq_len, d_out); break;
        case 3: density_kernel<3><<<grid, NT, smem, s>>>(d_new, n_new, d_seen, d_seen_fitness, n_seen, seq_len, d_out); break;
        default: density_kernel<4><<<grid, NT, smem, s>>>(d_new, n_new, d_seen, d_seen_fitness, n_seen, seq_len, d_out); break;
    }
    FX_CUDA(cudaGetLastError());
    return FLEXS_OK;
}

}  // extern "C"
