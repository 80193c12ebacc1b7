// K3 selection: the k best scores, descending, with their indices.
// Replaces np.argsort(preds)[: -B : -1] (adalead.py:171-175, cbas_dbas.py:197-201,
// cmaes.py:117-122; caller passes k = B-1) and np.argsort(preds)[::-1][:B] (dyna_ppo.py:315-319).
//
// Radix select on a 64-bit composite key (order-preserving score bits << 32 | ~position): keys are
// unique, so the k-th largest key is an exact threshold and exactly k elements pass it, with ties in
// score resolved towards the LOWER position.  8 histogram passes (HBM-bound, 4 B/element each, 16-byte
// loads), one collect pass, one single-CTA bitonic sort of the k survivors.
#include <algorithm>

#include "common.cuh"

namespace {

constexpr int NT = 256;
constexpr int MAXK = 4096;

struct Work {
    unsigned long long prefix[9];  // prefix[d] = selected top d bytes of the threshold key
    unsigned int krem[9];          // rank still to find below the prefix
    unsigned int count;            // survivors appended so far
    unsigned int pad;
    unsigned int hist[8][256];
};

__device__ __forceinline__ unsigned int ord32(float f) {
    if (f == 0.f) f = 0.f;  // -0.0 and +0.0 compare equal in numpy
    unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unord32(unsigned int k) {
    unsigned int u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}
__device__ __forceinline__ unsigned long long make_key(float s, unsigned int pos) {
    return ((unsigned long long)ord32(s) << 32) | (unsigned long long)(0xffffffffu - pos);
}

__global__ void topk_init_kernel(Work *w, unsigned int k) {
    const int t = threadIdx.x;
    for (int i = t; i < 8 * 256; i += blockDim.x) (&w->hist[0][0])[i] = 0;
    if (t < 9) { w->prefix[t] = 0; w->krem[t] = k; }
    if (t == 0) { w->count = 0; w->pad = 0; }
}

// state after digit d-1 from state d-1... derive (prefix[d], krem[d]) from hist[d-1]
__device__ void advance_state(Work *w, int d, unsigned long long &prefix, unsigned int &krem, bool writer) {
    __shared__ unsigned long long s_prefix;
    __shared__ unsigned int s_krem;
    if (threadIdx.x == 0) {
        unsigned long long pf = w->prefix[d - 1];
        unsigned int kr = w->krem[d - 1];
        unsigned int cum = 0;
        int b = 255;
        for (; b > 0; --b) {
            const unsigned int h = w->hist[d - 1][b];
            if (cum + h >= kr) break;
            cum += h;
        }
        pf = (pf << 8) | (unsigned long long)b;
        kr -= cum;
        s_prefix = pf; s_krem = kr;
        if (writer) { w->prefix[d] = pf; w->krem[d] = kr; }
    }
    __syncthreads();
    prefix = s_prefix; krem = s_krem;
}

__global__ void __launch_bounds__(NT) topk_hist_kernel(const float *__restrict__ scores, long long n, int d, Work *w) {
    __shared__ unsigned int sh[256];
    for (int i = threadIdx.x; i < 256; i += NT) sh[i] = 0;
    unsigned long long prefix = 0;
    unsigned int krem = 0;
    if (d > 0) advance_state(w, d, prefix, krem, blockIdx.x == 0);
    else __syncthreads();
    const int shift_digit = 56 - 8 * d;
    const long long stride = (long long)gridDim.x * NT;
    int run_digit = -1;
    unsigned int run_count = 0;
    for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += stride) {
        const unsigned long long key = make_key(__ldg(scores + i), (unsigned int)i);
        if (d > 0 && (key >> (64 - 8 * d)) != prefix) continue;
        const int digit = (int)((key >> shift_digit) & 0xff);
        if (digit == run_digit) { ++run_count; }
        else {
            if (run_count) atomicAdd(&sh[run_digit], run_count);
            run_digit = digit; run_count = 1;
        }
    }
    if (run_count) atomicAdd(&sh[run_digit], run_count);
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += NT)
        if (sh[i]) atomicAdd(&w->hist[d][i], sh[i]);
}

__global__ void __launch_bounds__(NT) topk_collect_kernel(const float *__restrict__ scores, long long n, Work *w,
                                                          unsigned long long *cand, unsigned int kcap) {
    unsigned long long thr = 0;
    unsigned int krem = 0;
    advance_state(w, 8, thr, krem, blockIdx.x == 0);
    const long long stride = (long long)gridDim.x * NT;
    for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += stride) {
        const unsigned long long key = make_key(__ldg(scores + i), (unsigned int)i);
        if (key >= thr) {
            const unsigned int slot = atomicAdd(&w->count, 1u);
            if (slot < kcap) cand[slot] = key;
        }
    }
}

__global__ void __launch_bounds__(1024) topk_sort_kernel(const unsigned long long *cand, const Work *w, int k, int kpad,
                                                         long long index_offset, const long long *index_map,
                                                         float *top_scores, long long *top_idx) {
    extern __shared__ unsigned long long keys[];
    const int t = threadIdx.x, nt = blockDim.x;
    const unsigned int have = min(w->count, (unsigned int)k);
    for (int i = t; i < kpad; i += nt) keys[i] = (i < (int)have) ? cand[i] : 0ull;
    __syncthreads();
    // bitonic sort, descending
    for (int size = 2; size <= kpad; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = t; i < kpad / 2; i += nt) {
                const int lo = 2 * i - (i & (stride - 1));
                const int hi = lo + stride;
                const bool desc = ((lo & size) == 0);
                const unsigned long long a = keys[lo], b = keys[hi];
                if ((a < b) == desc) { keys[lo] = b; keys[hi] = a; }
            }
            __syncthreads();
        }
    }
    for (int i = t; i < k; i += nt) {
        if (i < (int)have) {
            const unsigned long long key = keys[i];
            const unsigned int pos = 0xffffffffu - (unsigned int)(key & 0xffffffffu);
            top_scores[i] = unord32((unsigned int)(key >> 32));
            top_idx[i] = index_map ? index_map[pos] : (long long)pos + index_offset;
        } else {
            top_scores[i] = -INFINITY;
            top_idx[i] = -1;
        }
    }
}

int next_pow2(int v) {
    int p = 2;
    while (p < v) p <<= 1;
    return p;
}

}  // namespace

extern "C" {

int64_t flexs_topk_workspace_bytes(int64_t n, int k) {
    (void)n;
    if (k < 1 || k > MAXK) return FLEXS_EINVAL;
    return (int64_t)sizeof(Work) + 256 + (int64_t)next_pow2(k) * 8;
}

int flexs_topk_dev(const float *d_scores, int64_t n, int k, int64_t index_offset, const int64_t *d_index_map,
                   float *d_top_scores, int64_t *d_top_idx, void *d_work, void *stream) {
    FX_REQUIRE(k >= 1 && k <= MAXK, "k must be in [1, 4096]");
    FX_REQUIRE(n >= 0 && n < (1ll << 32), "n must be below 2^32");
    FX_REQUIRE(d_top_scores && d_top_idx && d_work, "null buffer");
    FX_REQUIRE(n == 0 || d_scores, "null scores");
    cudaStream_t s = (cudaStream_t)stream;
    Work *w = reinterpret_cast<Work *>(d_work);
    unsigned long long *cand = reinterpret_cast<unsigned long long *>(
        reinterpret_cast<unsigned char *>(d_work) + ((sizeof(Work) + 255) / 256) * 256);
    const int kpad = next_pow2(k);
    const unsigned int keff = (unsigned int)std::min<int64_t>(k, n);
    topk_init_kernel<<<1, 256, 0, s>>>(w, keff);
    if (n > 0) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + NT - 1) / NT, (int64_t)sms * 8));
        for (int d = 0; d < 8; ++d) topk_hist_kernel<<<grid, NT, 0, s>>>(d_scores, n, d, w);
        topk_collect_kernel<<<grid, NT, 0, s>>>(d_scores, n, w, cand, (unsigned int)kpad);
    }
    const int threads = std::min(1024, std::max(32, kpad / 2));
    topk_sort_kernel<<<1, threads, (size_t)kpad * 8, s>>>(cand, w, k, kpad, index_offset,
                                                          reinterpret_cast<const long long *>(d_index_map),
                                                          d_top_scores, reinterpret_cast<long long *>(d_top_idx));
    FX_CUDA(cudaGetLastError());
    return FLEXS_OK;
}

}  // extern "C"
