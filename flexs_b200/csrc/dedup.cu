// K3b: exact de-duplication of a candidate batch before the ranking.
// The reference ranks the KEYS of a dict (adalead.py:157 children, cmaes.py:112-115 seen, dyna_ppo.py:310-314
// all_seqs): a sequence proposed twice competes once.  A 1M-candidate screen over the 65 536 8-mers is >= 93 %
// duplicates, so the sharded screen must do the same or its top-k is k copies of one winner.
//
// Open-addressing hash table keyed by the sequence itself: a slot is ONE 64-bit word (24-bit hash tag | 40-bit index
// of the row that claimed it), so a claim is a single CAS and a later row that meets an equal tag always finds the
// representative to compare its bytes against (no second field to publish).  Equal rows meet in the same slot and
// atomicMin their index there; the survivor of a class is its LOWEST index, as in a dict filled in order.  Exact: a
// hash collision only costs a probe, never merges different sequences.  HBM-bound: each row is read ~2-3 times.
#include <cstdint>

#include "common.cuh"

namespace {

constexpr int NT = 256;
constexpr unsigned long long EMPTY = ~0ull;

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {  // splitmix64 finaliser
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27; x *= 0x94d049bb133111ebull;
    x ^= x >> 31;
    return x;
}

__device__ __forceinline__ unsigned long long hash_row(const uint8_t *__restrict__ row, int L) {
    unsigned long long h = 0x9e3779b97f4a7c15ull ^ (unsigned long long)L;
    int i = 0;
    for (; i + 8 <= L; i += 8) {
        unsigned long long w = 0;
#pragma unroll
        for (int b = 0; b < 8; ++b) w |= (unsigned long long)row[i + b] << (8 * b);
        h = mix64(h ^ w);
    }
    unsigned long long w = 0;
    for (int b = 0; i + b < L; ++b) w |= (unsigned long long)row[i + b] << (8 * b);
    return mix64(h ^ w ^ 0xabcdef);
}

__device__ __forceinline__ bool rows_equal(const uint8_t *__restrict__ a, const uint8_t *__restrict__ b, int L) {
    for (int i = 0; i < L; ++i)
        if (a[i] != b[i]) return false;
    return true;
}

__global__ void dedup_clear_kernel(unsigned long long *slots, long long *minidx, int64_t cap) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < cap; i += (int64_t)gridDim.x * blockDim.x) {
        slots[i] = EMPTY;
        minidx[i] = 0x7fffffffffffffffll;
    }
}

// every row finds (or founds) the slot of its class and leaves its index there if it is the lowest so far
__global__ void dedup_insert_kernel(const uint8_t *__restrict__ idx, int64_t n, int L, unsigned long long *slots,
                                    long long *minidx, int64_t cap, unsigned int *slot_of) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint8_t *row = idx + i * L;
        const unsigned long long h = hash_row(row, L);
        const unsigned long long tag = h >> 40;                               // 24 bits
        const unsigned long long mine = (tag << 40) | (unsigned long long)i;  // i < 2^40
        int64_t s = (int64_t)(h & (unsigned long long)(cap - 1));
        for (;;) {
            unsigned long long cur = slots[s];
            if (cur == EMPTY) {
                const unsigned long long old = atomicCAS(&slots[s], EMPTY, mine);
                cur = (old == EMPTY) ? mine : old;
            }
            if ((cur >> 40) == tag) {
                const int64_t rep = (int64_t)(cur & ((1ull << 40) - 1));
                if (rep == i || rows_equal(row, idx + rep * L, L)) break;   // this is the slot of my class
            }
            s = (s + 1) & (cap - 1);
        }
        atomicMin(reinterpret_cast<long long *>(&minidx[s]), (long long)i);
        slot_of[i] = (unsigned int)s;
    }
}

__global__ void dedup_mask_kernel(const float *__restrict__ scores, int64_t n, const long long *__restrict__ minidx,
                                  const unsigned int *__restrict__ slot_of, float *__restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (minidx[slot_of[i]] == (long long)i) ? scores[i] : -INFINITY;
}

__global__ void dedup_rep_kernel(int64_t n, const long long *__restrict__ minidx, const unsigned int *__restrict__ slot_of,
                                 long long *__restrict__ rep) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        rep[i] = minidx[slot_of[i]];
}

int64_t capacity_for(int64_t n) {
    int64_t cap = 1024;
    while (cap < 2 * n) cap <<= 1;
    return cap;
}

}  // namespace

extern "C" {

int64_t flexs_dedup_workspace_bytes(int64_t n) {
    if (n < 0 || n >= (1ll << 31)) return FLEXS_EINVAL;
    const int64_t cap = capacity_for(n);
    return cap * 16 + ((n * 4 + 15) / 16) * 16;
}

int flexs_dedup_scores_dev(const uint8_t *d_idx, int64_t n, int seq_len, const float *d_scores, float *d_scores_out,
                           void *d_work, void *stream) {
    FX_REQUIRE(n >= 0 && n < (1ll << 31), "n must be in [0, 2^31)");
    if (n == 0) return FLEXS_OK;
    FX_REQUIRE(d_idx && d_scores && d_scores_out && d_work, "null buffer");
    FX_REQUIRE(seq_len >= 1, "seq_len must be positive");
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t cap = capacity_for(n);
    unsigned long long *slots = reinterpret_cast<unsigned long long *>(d_work);
    long long *minidx = reinterpret_cast<long long *>(slots + cap);
    unsigned int *slot_of = reinterpret_cast<unsigned int *>(minidx + cap);
    const int grid_cap = (int)std::min<int64_t>((cap + NT - 1) / NT, 148 * 8);
    const int grid_n = (int)std::min<int64_t>((n + NT - 1) / NT, 148 * 8);
    dedup_clear_kernel<<<grid_cap, NT, 0, s>>>(slots, minidx, cap);
    dedup_insert_kernel<<<grid_n, NT, 0, s>>>(d_idx, n, seq_len, slots, minidx, cap, slot_of);
    dedup_mask_kernel<<<grid_n, NT, 0, s>>>(d_scores, n, minidx, slot_of, d_scores_out);
    FX_CUDA(cudaGetLastError());
    return FLEXS_OK;
}

int flexs_dedup_representatives_dev(const uint8_t *d_idx, int64_t n, int seq_len, int64_t *d_rep, void *d_work,
                                    void *stream) {
    FX_REQUIRE(n >= 0 && n < (1ll << 31), "n must be in [0, 2^31)");
    if (n == 0) return FLEXS_OK;
    FX_REQUIRE(d_idx && d_rep && d_work, "null buffer");
    FX_REQUIRE(seq_len >= 1, "seq_len must be positive");
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t cap = capacity_for(n);
    unsigned long long *slots = reinterpret_cast<unsigned long long *>(d_work);
    long long *minidx = reinterpret_cast<long long *>(slots + cap);
    unsigned int *slot_of = reinterpret_cast<unsigned int *>(minidx + cap);
    const int grid_cap = (int)std::min<int64_t>((cap + NT - 1) / NT, 148 * 8);
    const int grid_n = (int)std::min<int64_t>((n + NT - 1) / NT, 148 * 8);
    dedup_clear_kernel<<<grid_cap, NT, 0, s>>>(slots, minidx, cap);
    dedup_insert_kernel<<<grid_n, NT, 0, s>>>(d_idx, n, seq_len, slots, minidx, cap, slot_of);
    dedup_rep_kernel<<<grid_n, NT, 0, s>>>(n, minidx, slot_of, reinterpret_cast<long long *>(d_rep));
    FX_CUDA(cudaGetLastError());
    return FLEXS_OK;
}

}  // extern "C"
