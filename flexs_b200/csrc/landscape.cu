// K6 table landscapes: ground-truth landscapes that are pure tables, evaluated where the candidates already are.
//   * additive  — AdditiveAAVPackaging._get_raw_fitness / _fitness_function
//                 (flexs/landscapes/additive_aav_packaging.py:101-118): a per-position table gather-sum in
//                 float64, in position order (the reference's Python float loop), normalise, add noise, clip at 0.
//   * lookup    — TFBinding._fitness_function (flexs/landscapes/tf_binding.py:43-44): a dictionary keyed by the
//                 sequence, stored as a dense table indexed by the base-A packed residues.
// Byte/integer work + one dependent float64 add chain per sequence; HBM traffic is L bytes in + 8 bytes out per
// sequence, the tables live in shared memory (or L2 when too large).  Sums are sequential __dadd_rn so the result is
// bit-identical to CPython's left-to-right float additions.
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace {

struct ColumnOfChar {
    unsigned char col[256];  // 0xFF: character has no table column
};

constexpr int LS_THREADS = 128;

template <bool TABLE_IN_SMEM>
__global__ void __launch_bounds__(LS_THREADS) additive_kernel(const uint8_t *__restrict__ seq, int64_t n, int L,
                                                              ColumnOfChar lut, int use_lut, int ncols,
                                                              const double *__restrict__ table, double offset,
                                                              double denom, const double *__restrict__ noise,
                                                              double *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char ls_smem[];
    double *stab = reinterpret_cast<double *>(ls_smem);
    __shared__ uint8_t scol[256];
    for (int i = threadIdx.x; i < 256; i += LS_THREADS) scol[i] = use_lut ? lut.col[i] : (uint8_t)(i < ncols ? i : 0xFF);
    if (TABLE_IN_SMEM)
        for (int i = threadIdx.x; i < L * ncols; i += LS_THREADS) stab[i] = __ldg(table + i);
    __syncthreads();
    const double *tab = TABLE_IN_SMEM ? stab : table;
    const int64_t stride = (int64_t)gridDim.x * LS_THREADS;
    const uintptr_t buf0 = reinterpret_cast<uintptr_t>(seq), buf1 = buf0 + (uintptr_t)(n * L);
    for (int64_t s = (int64_t)blockIdx.x * LS_THREADS + threadIdx.x; s < n; s += stride) {
        // A thread streams its own sequence in aligned 16-byte words (both halves of every 32-byte sector are used by
        // consecutive loads of the same thread), not byte by byte: byte loads at a 90-byte lane stride cost a sector each.
        const uintptr_t g0 = reinterpret_cast<uintptr_t>(seq) + (uintptr_t)(s * L), g1 = g0 + (uintptr_t)L;
        double total = 0.0;
        for (uintptr_t wa = g0 & ~(uintptr_t)15; wa < g1; wa += 16) {
            uint32_t w4[4] = {0u, 0u, 0u, 0u};
            if (wa >= buf0 && wa + 16 <= buf1) {
                const uint4 q = __ldg(reinterpret_cast<const uint4 *>(wa));
                w4[0] = q.x; w4[1] = q.y; w4[2] = q.z; w4[3] = q.w;
            } else {  // the word straddles an end of the buffer: stay inside it
                for (int k = 0; k < 16; ++k) {
                    const uintptr_t ga = wa + (uintptr_t)k;
                    if (ga >= buf0 && ga < buf1) w4[k >> 2] |= (uint32_t)__ldg(reinterpret_cast<const uint8_t *>(ga)) << (8 * (k & 3));
                }
            }
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const uintptr_t ga = wa + (uintptr_t)k;
                if (ga < g0 || ga >= g1) continue;
                const int i = (int)(ga - g0);
                const int c = scol[(w4[k >> 2] >> (8 * (k & 3))) & 0xFF];
                // `if s in self.data[pos]: total_fitness += ...` (additive_aav_packaging.py:103-105)
                if (c != 0xFF) total = __dadd_rn(total, TABLE_IN_SMEM ? tab[i * ncols + c] : __ldg(tab + i * ncols + c));
            }
        }
        // (raw + mfm * max_possible) / (max_possible * (mfm + 1)) + noise, then max(0, .)   (:107, :112-116)
        double f = __ddiv_rn(__dadd_rn(total, offset), denom);
        if (noise) f = __dadd_rn(f, noise[s]);
        out[s] = (f > 0.0) ? f : 0.0;  // Python max(0, x): 0 unless x > 0 (also for NaN)
    }
}

__global__ void __launch_bounds__(256) lookup_kernel(const uint8_t *__restrict__ seq, int64_t n, int L, ColumnOfChar lut,
                                                     int use_lut, int base, const double *__restrict__ table,
                                                     double *__restrict__ out) {
    __shared__ uint8_t scol[256];
    for (int i = threadIdx.x; i < 256; i += 256) scol[i] = use_lut ? lut.col[i] : (uint8_t)(i < base ? i : 0xFF);
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (int64_t s = (int64_t)blockIdx.x * 256 + threadIdx.x; s < n; s += stride) {
        const uint8_t *p = seq + s * L;
        int64_t key = 0;
        bool known = true;
        for (int i = 0; i < L; ++i) {
            const int c = scol[__ldg(p + i)];
            known = known && (c != 0xFF);
            key = key * base + (c & 0x7F);
        }
        // a sequence that is not a key of the dictionary is reported as NaN; the host raises KeyError (tf_binding.py:44)
        out[s] = known ? __ldg(table + key) : __longlong_as_double(0x7ff8000000000000ll);
    }
}

int blocks_for(int64_t n, int threads) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t want = (n + threads - 1) / threads;
    return (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)sms * 8));
}

}  // namespace

extern "C" {

int flexs_additive_score_dev(const uint8_t *d_seq, int64_t n, int seq_len, const uint8_t *h_column_of_char, int ncols,
                             const double *d_table, double offset, double denom, const double *d_noise, double *d_out,
                             void *stream) {
    FX_REQUIRE(n >= 0 && seq_len >= 1, "bad sizes");
    FX_REQUIRE(ncols >= 1 && ncols < 255, "ncols must be in [1, 254]");
    if (n == 0) return FLEXS_OK;
    FX_REQUIRE(d_seq && d_table && d_out, "null buffer");
    ColumnOfChar lut;
    for (int i = 0; i < 256; ++i) {
        lut.col[i] = h_column_of_char ? h_column_of_char[i] : 0xFF;
        FX_REQUIRE(lut.col[i] == 0xFF || lut.col[i] < ncols, "column_of_char entry out of range");
    }
    const size_t bytes = (size_t)seq_len * ncols * sizeof(double);
    const int grid = blocks_for(n, LS_THREADS);
    cudaStream_t s = (cudaStream_t)stream;
    if (bytes <= 200 * 1024) {
        FX_CUDA(cudaFuncSetAttribute(additive_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        additive_kernel<true><<<grid, LS_THREADS, bytes, s>>>(d_seq, n, seq_len, lut, h_column_of_char != nullptr, ncols,
                                                              d_table, offset, denom, d_noise, d_out);
    } else {
        additive_kernel<false><<<grid, LS_THREADS, 0, s>>>(d_seq, n, seq_len, lut, h_column_of_char != nullptr, ncols,
                                                           d_table, offset, denom, d_noise, d_out);
    }
    FX_CUDA(cudaGetLastError());
    return FLEXS_OK;
}

int flexs_lookup_score_dev(const uint8_t *d_seq, int64_t n, int seq_len, const uint8_t *h_column_of_char, int base,
                           const double *d_table, int64_t table_len, double *d_out, void *stream) {
    FX_REQUIRE(n >= 0 && seq_len >= 1, "bad sizes");
    FX_REQUIRE(base >= 1 && base <= 127, "base must be in [1, 127]");
    int64_t need = 1;
    for (int i = 0; i < seq_len; ++i) {
        need *= base;
        FX_REQUIRE(need <= ((int64_t)1 << 40), "table too large");
    }
    FX_REQUIRE(table_len == need, "table_len must equal base ** seq_len");
    if (n == 0) return FLEXS_OK;
    FX_REQUIRE(d_seq && d_table && d_out, "null buffer");
    ColumnOfChar lut;
    for (int i = 0; i < 256; ++i) {
        lut.col[i] = h_column_of_char ? h_column_of_char[i] : 0xFF;
        FX_REQUIRE(lut.col[i] == 0xFF || lut.col[i] < base, "column_of_char entry out of range");
    }
    lookup_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(d_seq, n, seq_len, lut, h_column_of_char != nullptr,
                                                                         base, d_table, d_out);
    FX_CUDA(cudaGetLastError());
    return FLEXS_OK;
}

}  // extern "C"
