// K3c: single-launch exact top-k, optionally over DISTINCT rows, and the one-kernel merge of the all-gathered
// per-shard lists of a multi-GPU screen.
//
// Replaces the final ranking of propose_sequences — np.argsort(preds)[: -B : -1] over the keys of a dict
// (adalead.py:157,171-175; cbas_dbas.py:197-201; cmaes.py:112-122) and np.argsort(preds)[::-1][:B] (dyna_ppo.py:310-319)
// — for batches where the 11-launch radix select of topk.cu plus the hash de-duplication of dedup.cu dominate a step
// (a 1M-candidate screen of 8-mers: 0.11 ms of selection behind 0.013 ms of scoring).
//
// topk_select_kernel (cooperative launch, one CTA per SM): radix select on the 64-bit key
// (order-preserving score bits << 32 | ~position) with 12-bit digits; a level histograms the keys that match the prefix
// found so far, a grid barrier, every CTA locates the bin of the wanted rank; the descent stops as soon as "everything
// above the bin + the bin" fits the candidate buffer (8192 keys) — one level for any score distribution that is not a
// wall of ties — and the surviving keys are collected.  CTA 0 sorts them (bitonic, shared memory) and writes the winners.
//
// Distinct rows ("unique"): the reference ranks dict keys, a sequence proposed twice competes once.  Hashing all n rows
// (dedup.cu) reads the whole batch two to three times; here the select asks for the best max(k, 4096) rows and
// de-duplicates only those, in rank order: equal rows have equal scores (the surrogate is deterministic), so the first
// k distinct rows of the ranked candidates ARE the k best distinct rows of the batch whenever the candidates hold k
// distinct rows.  When they do not (a batch that is almost all repeats) status[0] = 1 and the caller falls back to
// dedup.cu + topk.cu; nothing is ever approximated.
//
// screen_merge_kernel (one CTA): the gathered messages of all ranks ([k] int64 index | [k] float score | [k][L] rows
// each) -> sort, drop sequences that reached the top-k of two shards (the copy with the lower global index survives),
// first k.  Replaces ~25 small launches (unpack, dedup, radix select) behind the collective.
#include <cooperative_groups.h>

#include <algorithm>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int NT = 512;
constexpr int CAP = 8192;          // candidate keys the final sort takes
constexpr int KSLACK = 4096;       // rank asked for in unique mode (room for repeats among the best rows)
constexpr int NBIN = 4096, DIGIT = 12, NLEVEL = 6;

struct SelWork {
    unsigned int hist[NLEVEL][NBIN];
    unsigned int count;
    unsigned int pad[3];
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

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {  // splitmix64 finaliser
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27; x *= 0x94d049bb133111ebull;
    x ^= x >> 31;
    return x;
}
__device__ __forceinline__ unsigned long long hash_row(const uint8_t *__restrict__ row, int L) {
    unsigned long long h = 0x9e3779b97f4a7c15ull ^ (unsigned long long)L;
    int i = 0;
    if ((reinterpret_cast<uintptr_t>(row) & 3) == 0) {   // word loads where the row allows (same value as the byte path)
        const uint32_t *w32 = reinterpret_cast<const uint32_t *>(row);
#pragma unroll 4
        for (; i + 8 <= L; i += 8) h = mix64(h ^ ((unsigned long long)w32[i >> 2] | ((unsigned long long)w32[(i >> 2) + 1] << 32)));
    }
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
    int i = 0;
    if (((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 3) == 0)
        for (; i + 4 <= L; i += 4)
            if (*reinterpret_cast<const uint32_t *>(a + i) != *reinterpret_cast<const uint32_t *>(b + i)) return false;
    for (; i < L; ++i)
        if (a[i] != b[i]) return false;
    return true;
}

// in-place bitonic sort of `len` (a power of two) 64-bit keys in shared memory; all NT threads call it
template <bool DESC>
__device__ void bitonic_sort(unsigned long long *keys, int len) {
    for (int size = 2; size <= len; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = threadIdx.x; i < len / 2; i += NT) {
                const int lo = 2 * i - (i & (stride - 1));
                const int hi = lo + stride;
                const bool up = ((lo & size) == 0);
                const unsigned long long a = keys[lo], b = keys[hi];
                if (((a < b) == up) == DESC) { keys[lo] = b; keys[hi] = a; }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ int pow2_at_least(int v) {
    int p = 2;
    while (p < v) p <<= 1;
    return p;
}

__device__ __forceinline__ unsigned long long globaltimer_ns();

// ---- final stage, one CTA ----------------------------------------------------------------------------------------
// bin of rank `krem` (counted from the top) in a 4096-bin histogram: suffix sums, 8 consecutive bins per thread.
// All NT threads call it; returns through shared memory (bin, count above the bin, count in the bin).
__device__ void find_bin(const unsigned int *hist, bool global, unsigned int krem, unsigned int &bin, unsigned int &above,
                         unsigned int &cnt) {
    __shared__ unsigned int s_part[NT], s_res[3];
    const int t = threadIdx.x;
    constexpr int PER = NBIN / NT;
    unsigned int loc[PER], tot = 0;
#pragma unroll
    for (int i = 0; i < PER; ++i) { loc[i] = global ? __ldcg(hist + t * PER + i) : hist[t * PER + i]; tot += loc[i]; }
    s_part[t] = tot;
    __syncthreads();
    for (int off = 1; off < NT; off <<= 1) {   // suffix[t] = sum of s_part[t ..]
        const unsigned int v = (t + off < NT) ? s_part[t + off] : 0u;
        __syncthreads();
        s_part[t] += v;
        __syncthreads();
    }
    unsigned int cum = s_part[t] - tot;  // bins above my span
    if (cum < krem && cum + tot >= krem) {
#pragma unroll
        for (int i = PER - 1; i >= 0; --i) {
            if (cum + loc[i] >= krem) { s_res[0] = (unsigned int)(t * PER + i); s_res[1] = cum; s_res[2] = loc[i]; break; }
            cum += loc[i];
        }
    }
    __syncthreads();
    bin = s_res[0]; above = s_res[1]; cnt = s_res[2];
    __syncthreads();
}

// The candidates a grid-level descent leaves can be thousands although only the best `need` matter: the same radix
// descent, inside the CTA over the keys in shared memory, narrows them to at most NARROW keys before anything is sorted
// (sorting 8192 keys costs 91 bitonic stages of 8 exchanges per thread; 2048 keys cost 66 stages of 2).
constexpr int NARROW = 2048;
__device__ int narrow(const unsigned long long *src, int m, unsigned int need, unsigned long long *dst, unsigned int *hist) {
    __shared__ unsigned int s_cnt;
    const int t = threadIdx.x;
    unsigned long long prefix = 0, thr_prefix = 0;
    int bits_done = 0, thr_bits = 0;
    unsigned int krem = need, above_total = 0;
    for (int level = 0; level < NLEVEL; ++level) {
        const int dbits = min(DIGIT, 64 - bits_done), shift = 64 - bits_done - dbits;
        for (int i = t; i < NBIN; i += NT) hist[i] = 0;
        __syncthreads();
        for (int i0 = t & ~31; i0 < m; i0 += NT) {          // a warp over 32 consecutive keys, equal digits added once
            const int i = i0 + (t & 31);
            bool live = i < m;
            unsigned int digit = 0;
            if (live) {
                const unsigned long long key = src[i];
                live = !(bits_done > 0 && (key >> (64 - bits_done)) != prefix);
                digit = (unsigned int)(key >> shift) & ((1u << dbits) - 1u);
            }
            const unsigned int peers = __match_any_sync(0xffffffffu, live ? digit : 0xffffffffu);
            if (live && (t & 31) == __ffs(peers) - 1) atomicAdd(&hist[digit], (unsigned int)__popc(peers));
        }
        __syncthreads();
        unsigned int b, above, cnt;
        find_bin(hist, false, krem, b, above, cnt);
        thr_prefix = (prefix << dbits) | b;
        thr_bits = bits_done + dbits;
        if (above_total + above + cnt <= (unsigned int)NARROW || thr_bits == 64) break;
        prefix = thr_prefix; bits_done = thr_bits; krem -= above; above_total += above;
    }
    if (t == 0) s_cnt = 0;
    __syncthreads();
    for (int i0 = t & ~31; i0 < m; i0 += NT) {   // one shared-memory atomic per warp, not per selected key
        const int i = i0 + (t & 31);
        const unsigned long long key = i < m ? src[i] : 0ull;
        const bool sel = i < m && (key >> (64 - thr_bits)) >= thr_prefix;
        const unsigned int mask = __ballot_sync(0xffffffffu, sel);
        unsigned int base = 0;
        if ((t & 31) == 0 && mask) base = atomicAdd(&s_cnt, (unsigned int)__popc(mask));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (sel) dst[base + __popc(mask & ((1u << (t & 31)) - 1u))] = key;
    }
    __syncthreads();
    const int m1 = (int)s_cnt;
    __syncthreads();
    return m1;
}

// keys[0 .. m) sorted descending (rank order).  Marks repeats among growing prefixes (starting at P0) until k distinct
// rows are found; rank_of[0 .. found) receives the ranks of the survivors in order.  L == 0: no de-duplication.
template <class RowOf>
__device__ int pick_distinct(const unsigned long long *keys, unsigned long long *hs, unsigned short *rank_of, int m, int P0,
                             int k, int L, RowOf row_of) {
    __shared__ int s_found, s_scan[NT];
    const int t = threadIdx.x;
    if (L == 0) {
        const int found = min(m, k);
        for (int i = t; i < found; i += NT) rank_of[i] = (unsigned short)i;
        __syncthreads();
        return found;
    }
    // Fast path.  Equal rows carry equal scores, so in rank order a row can only repeat an earlier member of its own run
    // of equal scores: the head of a run is always new, everyone else walks back through the run (usually one step: the
    // copy next to it) comparing rows.  No hashing, no second sort.  A run longer than RUN_MAX (a wall of tied scores
    // over different rows) sends the whole call to the hash path below.
    {
        constexpr int RUN_MAX = 64;
        __shared__ int s_long;
        if (t == 0) s_long = 0;
        __syncthreads();
        for (int i = t; i < m; i += NT) {
            const unsigned int sc = (unsigned int)(keys[i] >> 32);
            bool dup = false;
            if (i > 0 && (unsigned int)(keys[i - 1] >> 32) == sc) {
                const uint8_t *mine = row_of(0xffffffffu - (unsigned int)(keys[i] & 0xffffffffu));
                int j = i - 1, steps = 0;
                for (; j >= 0 && (unsigned int)(keys[j] >> 32) == sc && !dup && steps < RUN_MAX; --j, ++steps)
                    dup = rows_equal(mine, row_of(0xffffffffu - (unsigned int)(keys[j] & 0xffffffffu)), L);
                if (!dup && steps == RUN_MAX && j >= 0 && (unsigned int)(keys[j] >> 32) == sc) s_long = 1;
            }
            rank_of[i] = dup ? 1 : 0;
        }
        __syncthreads();
        if (!s_long) {
            const int span = (m + NT - 1) / NT, lo = t * span, hi = min(m, lo + span);
            int mine_cnt = 0;
            for (int i = lo; i < hi; ++i) mine_cnt += rank_of[i] ? 0 : 1;
            s_scan[t] = mine_cnt;
            __syncthreads();
            for (int off = 1; off < NT; off <<= 1) {
                const int v = (t >= off) ? s_scan[t - off] : 0;
                __syncthreads();
                s_scan[t] += v;
                __syncthreads();
            }
            if (t == NT - 1) s_found = s_scan[t];
            int base = s_scan[t] - mine_cnt;
            unsigned int *surv = reinterpret_cast<unsigned int *>(hs);
            for (int i = lo; i < hi; ++i)
                if (!rank_of[i] && base < k) surv[base++] = (unsigned int)i;
            __syncthreads();
            const int found = min(s_found, k);
            for (int i = t; i < found; i += NT) rank_of[i] = (unsigned short)surv[i];
            __syncthreads();
            return found;
        }
        __syncthreads();
    }
    for (int P = min(m, P0);; P = min(m, P * 2)) {
        const int plen = pow2_at_least(max(P, 2));
        for (int i = t; i < plen; i += NT) {
            if (i < P) {
                const unsigned int pos = 0xffffffffu - (unsigned int)(keys[i] & 0xffffffffu);
                hs[i] = (hash_row(row_of(pos), L) & ~0x1fffull) | (unsigned long long)i;   // 51 hash bits | rank (13 bits)
            } else hs[i] = ~0ull;
        }
        __syncthreads();
        bitonic_sort<false>(hs, plen);   // ascending: equal hashes adjacent, by rank
        // a candidate is a repeat iff an earlier member of its hash run has the same row (the run is walked from its
        // head: a batch of identical rows costs one comparison per row, not a quadratic number)
        for (int i = t; i < P; i += NT) {
            const unsigned long long e = hs[i];
            const int rk = (int)(e & 0x1fffull);
            bool dup = false;
            int j = i;
            while (j > 0 && (hs[j - 1] >> 13) == (e >> 13)) --j;
            if (j < i) {
                const uint8_t *mine = row_of(0xffffffffu - (unsigned int)(keys[rk] & 0xffffffffu));
                for (; j < i && !dup; ++j) {
                    const int rj = (int)(hs[j] & 0x1fffull);
                    dup = rows_equal(mine, row_of(0xffffffffu - (unsigned int)(keys[rj] & 0xffffffffu)), L);
                }
            }
            rank_of[rk] = dup ? 1 : 0;   // scratch: repeat flag by rank
        }
        __syncthreads();
        // ordered compaction of the survivors: block scan over ranks, each thread owns a contiguous span
        const int span = (P + NT - 1) / NT, lo = t * span, hi = min(P, lo + span);
        int mine_cnt = 0;
        for (int i = lo; i < hi; ++i) mine_cnt += rank_of[i] ? 0 : 1;
        s_scan[t] = mine_cnt;
        __syncthreads();
        for (int off = 1; off < NT; off <<= 1) {
            const int v = (t >= off) ? s_scan[t - off] : 0;
            __syncthreads();
            s_scan[t] += v;
            __syncthreads();
        }
        if (t == NT - 1) s_found = s_scan[t];
        int base = s_scan[t] - mine_cnt;
        __syncthreads();
        unsigned int *surv = reinterpret_cast<unsigned int *>(hs);   // the hash order is spent: survivors' ranks go here
        for (int i = lo; i < hi; ++i)
            if (!rank_of[i] && base < k) surv[base++] = (unsigned int)i;
        __syncthreads();
        const int found = min(s_found, k);
        if (found >= k || P >= m) {
            for (int i = t; i < found; i += NT) rank_of[i] = (unsigned short)surv[i];
            __syncthreads();
            return found;
        }
        __syncthreads();
    }
}

// k0[0 .. m) hold the candidates (any order); k1 and k2 are scratch of CAP keys each.  RowOf(pos) -> the row of the
// candidate whose key decodes to position pos (L == 0: rank rows, not distinct rows); IdxOf(pos) -> the index to report.
// Writes k winners (score desc, position asc), their rows when top_rows != nullptr, and the number found.
template <class RowOf, class IdxOf>
__device__ void finalize(unsigned long long *k0, unsigned long long *k1, unsigned long long *k2, unsigned short *rank_of,
                         int m, int k, int L, RowOf row_of, IdxOf idx_of, float *top_scores, long long *top_idx,
                         uint8_t *top_rows, int *n_found, int *diag = nullptr) {
    const int t = threadIdx.x;
    const unsigned long long td0 = globaltimer_ns();
    unsigned long long td1 = td0, td2 = td0, td3 = td0;
    const unsigned long long *sorted = nullptr;
    int found = -1;
    const int need = (L == 0) ? k : max(k, 1024);   // rows the first attempt ranks (unique: room for repeats)
    if (need <= 1024 && m > NARROW) {
        const int m1 = narrow(k0, m, (unsigned int)min(m, need), k1, reinterpret_cast<unsigned int *>(k2));
        td1 = globaltimer_ns();
        const int len1 = pow2_at_least(max(m1, 2));
        for (int i = m1 + t; i < len1; i += NT) k1[i] = 0ull;
        __syncthreads();
        bitonic_sort<true>(k1, len1);
        td2 = globaltimer_ns();
        const int f = pick_distinct(k1, k2, rank_of, m1, max(256, pow2_at_least(2 * k)), k, L, row_of);
        td3 = globaltimer_ns();
        if (f >= k || m1 >= m) { found = f; sorted = k1; }
    }
    if (found < 0) {   // few candidates, a large k, or a narrowed set that is mostly repeats: rank all of them
        const int len = pow2_at_least(max(m, 2));
        for (int i = m + t; i < len; i += NT) k0[i] = 0ull;
        __syncthreads();
        bitonic_sort<true>(k0, len);
        found = pick_distinct(k0, k2, rank_of, m, max(256, pow2_at_least(2 * k)), k, L, row_of);
        sorted = k0;
    }
    for (int i = t; i < k; i += NT) {
        if (i < found) {
            const unsigned long long key = sorted[rank_of[i]];
            const unsigned int pos = 0xffffffffu - (unsigned int)(key & 0xffffffffu);
            top_scores[i] = unord32((unsigned int)(key >> 32));
            top_idx[i] = idx_of(pos);
        } else {
            top_scores[i] = -INFINITY;
            top_idx[i] = -1;
        }
    }
    if (top_rows != nullptr && L > 0) {
        for (int i = t; i < k * L; i += NT) {
            const int w = i / L, b = i - w * L;
            uint8_t v = 0;
            if (w < found) v = row_of(0xffffffffu - (unsigned int)(sorted[rank_of[w]] & 0xffffffffu))[b];
            top_rows[i] = v;
        }
    }
    if (t == 0) *n_found = found;
    if (t == 0 && diag != nullptr) {   // ns: narrowing, sort of the narrowed set, de-duplication, everything after
        diag[0] = (int)(td1 - td0); diag[1] = (int)(td2 - td1); diag[2] = (int)(td3 - td2);
    }
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Visit every score once, a warp at a time over 128 consecutive elements (one 16-byte load per lane: four loads in
// flight per thread hide the L2 latency a one-element-per-iteration loop is bound by).  fn(score, index, live) is called
// by all 32 lanes together, four times per visit (live = the element exists), so it may use warp collectives.
template <class Fn>
__device__ __forceinline__ void for_each_score(const float *__restrict__ scores, long long n, Fn fn) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * NT + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * NT) >> 5;
    const bool vec = (reinterpret_cast<uintptr_t>(scores) & 15) == 0;
    for (long long base = warp * 128; base < n; base += nwarps * 128) {
        const long long i0 = base + lane * 4;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (vec && i0 + 3 < n) {
            const float4 q = __ldg(reinterpret_cast<const float4 *>(scores + i0));
            v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) if (i0 + e < n) v[e] = __ldg(scores + i0 + e);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) fn(v[e], i0 + e, i0 + e < n);
    }
}

struct SelParams {
    const float *scores;
    long long n;
    int k, unique;
    long long index_offset;
    const uint8_t *rows;
    int row_len;
    float *top_scores;
    long long *top_idx;
    uint8_t *top_rows;
    int *status;          // int[8]; [0] = 1: fewer than k distinct rows among the candidates although the batch has more rows
    SelWork *work;
    unsigned long long *cand;
};

__global__ void __launch_bounds__(NT, 1) topk_select_kernel(const SelParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem_raw);       // CAP keys; first NBIN words double as the level histogram
    unsigned long long *k1 = keys + CAP, *k2 = k1 + CAP;
    unsigned short *rank_of = reinterpret_cast<unsigned short *>(k2 + CAP);
    unsigned int *sh = reinterpret_cast<unsigned int *>(smem_raw);
    cg::grid_group grid = cg::this_grid();
    const int t = threadIdx.x;
    const long long stride = (long long)gridDim.x * NT;
    const long long gtid = (long long)blockIdx.x * NT + t;
    SelWork *w = p.work;

    const unsigned long long t_start = globaltimer_ns();
    // the workspace cleans itself: no separate init launch
    for (long long i = gtid; i < (long long)(NLEVEL * NBIN); i += stride) (&w->hist[0][0])[i] = 0;
    if (gtid == 0) w->count = 0;
    grid.sync();

    // unique: ask for 8 rows per wanted sequence (at least 4096), never more than the final sort takes
    const unsigned int want = (unsigned int)min(p.n, (long long)(p.unique ? min(CAP, max(8 * p.k, KSLACK)) : p.k));
    unsigned long long prefix = 0;
    int bits_done = 0;
    unsigned int krem = want, above_total = 0;
    unsigned long long thr_prefix = 0;
    int thr_bits = 0;
    if (p.n > 0) {
        for (int level = 0; level < NLEVEL; ++level) {
            const int dbits = min(DIGIT, 64 - bits_done);
            const int shift = 64 - bits_done - dbits;
            for (int i = t; i < NBIN; i += NT) sh[i] = 0;
            __syncthreads();
            // Scores of one batch crowd into a few bins of the first digit (sign, exponent, 3 mantissa bits): the warp
            // first agrees on who holds equal digits (match.any) and one lane adds the whole group — a handful of
            // shared-memory atomics per warp and load instead of 32 colliding ones.
            for_each_score(p.scores, p.n, [&](float sc, long long i, bool live) {
                unsigned int digit = 0;
                if (live) {
                    const unsigned long long key = make_key(sc, (unsigned int)i);
                    live = !(bits_done > 0 && (key >> (64 - bits_done)) != prefix);
                    digit = (unsigned int)(key >> shift) & ((1u << dbits) - 1u);
                }
                const unsigned int peers = __match_any_sync(0xffffffffu, live ? digit : 0xffffffffu);
                if (live && (unsigned int)(t & 31) == (unsigned int)(__ffs(peers) - 1)) atomicAdd(&sh[digit], (unsigned int)__popc(peers));
            });
            __syncthreads();
            for (int i = t; i < NBIN; i += NT)
                if (sh[i]) atomicAdd(&w->hist[level][i], sh[i]);
            grid.sync();
            // every CTA finds the bin that holds rank `krem` counting from the top
            unsigned int b, above, cnt;
            find_bin(w->hist[level], true, krem, b, above, cnt);
            thr_prefix = (prefix << dbits) | b;
            thr_bits = bits_done + dbits;
            if (above_total + above + cnt <= (unsigned int)CAP || thr_bits == 64) break;
            prefix = thr_prefix; bits_done = thr_bits; krem -= above; above_total += above;
        }
        // collect every key at or above the threshold prefix
        for_each_score(p.scores, p.n, [&](float sc, long long i, bool live) {
            if (!live) return;
            const unsigned long long key = make_key(sc, (unsigned int)i);
            if ((key >> (64 - thr_bits)) >= thr_prefix) {
                const unsigned int slot = atomicAdd(&w->count, 1u);
                if (slot < (unsigned int)CAP) p.cand[slot] = key;
            }
        });
    }
    grid.sync();
    if (blockIdx.x != 0) return;
    const unsigned long long t_collected = globaltimer_ns();
    const int m = (int)min(__ldcg(&w->count), (unsigned int)CAP);
    for (int i = t; i < m; i += NT) keys[i] = __ldcg(p.cand + i);
    __syncthreads();
    __shared__ int s_nfound;
    const uint8_t *rows = p.rows;
    const int L = (p.unique && rows != nullptr) ? p.row_len : 0;
    const long long off = p.index_offset;
    finalize(keys, k1, k2, rank_of, m, p.k, L,
             [rows, L2 = p.row_len](unsigned int pos) { return rows + (size_t)pos * L2; },
             [off](unsigned int pos) { return (long long)pos + off; },
             p.top_scores, p.top_idx, p.top_rows, &s_nfound, p.status ? p.status + 5 : nullptr);
    if (rows == nullptr && p.top_rows != nullptr)   // an empty shard still sends a well-defined message
        for (int i = t; i < p.k * p.row_len; i += NT) p.top_rows[i] = 0;
    // rows of the winners are wanted even when ranking rows rather than distinct rows
    if (!p.unique && p.top_rows != nullptr && p.rows != nullptr) {
        __syncthreads();
        for (int i = t; i < p.k * p.row_len; i += NT) {
            const int wi = i / p.row_len, b = i - wi * p.row_len;
            const long long gi = p.top_idx[wi];
            p.top_rows[i] = gi >= 0 ? rows[(size_t)(gi - off) * p.row_len + b] : (uint8_t)0;
        }
    }
    __syncthreads();
    if (t == 0 && p.status != nullptr) {
        p.status[0] = (p.unique && s_nfound < p.k && (long long)m < p.n) ? 1 : 0;
        // diagnostics: radix levels used, candidates sorted, ns spent selecting / in the final one-CTA stage
        p.status[1] = thr_bits / DIGIT + (thr_bits % DIGIT ? 1 : 0);
        p.status[2] = m;
        p.status[3] = (int)(t_collected - t_start);
        p.status[4] = (int)(globaltimer_ns() - t_collected);
    }
}

struct MergeParams {
    const unsigned char *gathered;   // world messages
    int world, k, row_len;
    long long msg_bytes, off_score, off_rows;
    float *top_scores;
    long long *top_idx;
    uint8_t *top_rows;
};

__global__ void __launch_bounds__(NT, 1) screen_merge_kernel(const MergeParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem_raw);
    unsigned long long *k1 = keys + CAP, *k2 = k1 + CAP;
    unsigned short *rank_of = reinterpret_cast<unsigned short *>(k2 + CAP);
    __shared__ int s_m, s_nfound;
    const int t = threadIdx.x, total = p.world * p.k;
    if (t == 0) s_m = 0;
    __syncthreads();
    // position j = rank-major order of the gathered lists: shards own increasing index ranges and each list is sorted
    // (score desc, index asc), so ties by position are ties by global index
    for (int j0 = t & ~31; j0 < total; j0 += NT) {
        const int j = j0 + (t & 31);
        bool live = j < total;
        unsigned long long key = 0;
        if (live) {
            const int r = j / p.k, i = j - r * p.k;
            const unsigned char *msg = p.gathered + (size_t)r * p.msg_bytes;
            live = reinterpret_cast<const long long *>(msg)[i] >= 0;   // < 0: an absent winner (fewer than k candidates on that shard)
            if (live) key = make_key(reinterpret_cast<const float *>(msg + p.off_score)[i], (unsigned int)j);
        }
        const unsigned int mask = __ballot_sync(0xffffffffu, live);
        int base = 0;
        if ((t & 31) == 0 && mask) base = atomicAdd(&s_m, __popc(mask));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (live) keys[base + __popc(mask & ((1u << (t & 31)) - 1u))] = key;
    }
    __syncthreads();
    const unsigned char *g = p.gathered;
    const int k = p.k, L = p.row_len;
    const long long mb = p.msg_bytes, orow = p.off_rows;
    finalize(keys, k1, k2, rank_of, s_m, k, L,
             [g, k, L, mb, orow](unsigned int j) {
                 const int r = (int)j / k, i = (int)j - r * k;
                 return reinterpret_cast<const uint8_t *>(g + (size_t)r * mb + orow + (size_t)i * L);
             },
             [g, k, mb](unsigned int j) {
                 const int r = (int)j / k, i = (int)j - r * k;
                 return reinterpret_cast<const long long *>(g + (size_t)r * mb)[i];
             },
             p.top_scores, p.top_idx, p.top_rows, &s_nfound);
}

constexpr size_t FINAL_SMEM = (size_t)CAP * 8 * 3 + (size_t)CAP * 2;

}  // namespace

extern "C" {

int64_t flexs_topk_select_workspace_bytes(void) { return (int64_t)((sizeof(SelWork) + 255) / 256 * 256 + (size_t)CAP * 8); }

int flexs_topk_select_dev(const float *d_scores, int64_t n, int k, int64_t index_offset, const uint8_t *d_rows,
                          int row_len, int unique, float *d_top_scores, int64_t *d_top_idx, uint8_t *d_top_rows,
                          int *d_status, void *d_work, void *stream) {
    FX_REQUIRE(k >= 1 && k <= 4096, "k must be in [1, 4096]");
    FX_REQUIRE(n >= 0 && n < (1ll << 32), "n must be below 2^32");
    FX_REQUIRE(d_top_scores && d_top_idx && d_work, "null buffer");
    FX_REQUIRE(n == 0 || d_scores, "null scores");
    FX_REQUIRE(!unique || n == 0 || (d_rows && row_len >= 1), "unique ranking needs the rows");
    FX_REQUIRE(d_top_rows == nullptr || n == 0 || (d_rows && row_len >= 1), "winner rows need the rows");
    SelParams p;
    p.scores = d_scores; p.n = n; p.k = k; p.unique = unique ? 1 : 0; p.index_offset = index_offset;
    p.rows = n > 0 ? d_rows : nullptr; p.row_len = row_len;
    p.top_scores = d_top_scores; p.top_idx = reinterpret_cast<long long *>(d_top_idx); p.top_rows = d_top_rows;
    p.status = d_status;
    p.work = reinterpret_cast<SelWork *>(d_work);
    p.cand = reinterpret_cast<unsigned long long *>(reinterpret_cast<unsigned char *>(d_work) + (sizeof(SelWork) + 255) / 256 * 256);
    int dev = 0, sms = 148;
    FX_CUDA(cudaGetDevice(&dev));
    FX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    FX_CUDA(cudaFuncSetAttribute(topk_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FINAL_SMEM));
    // a small batch does not need the whole GPU; every CTA of a cooperative launch must be resident (1 per SM here)
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + NT * 8 - 1) / (NT * 8), sms));
    void *args[] = {&p};
    FX_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void *>(topk_select_kernel), dim3(grid), dim3(NT), args, FINAL_SMEM,
                                        (cudaStream_t)stream));
    return FLEXS_OK;
}

int64_t flexs_screen_message_bytes(int k, int seq_len) {
    if (k < 1 || seq_len < 0) return FLEXS_EINVAL;
    return ((int64_t)k * 12 + (int64_t)k * seq_len + 15) / 16 * 16;
}

int flexs_screen_merge_dev(const void *d_gathered, int world, int k, int seq_len, float *d_top_scores,
                           int64_t *d_top_idx, uint8_t *d_top_rows, void *stream) {
    FX_REQUIRE(world >= 1 && k >= 1 && seq_len >= 0, "bad sizes");
    FX_REQUIRE((int64_t)world * k <= CAP, "world * k must not exceed 8192");
    FX_REQUIRE(d_gathered && d_top_scores && d_top_idx, "null buffer");
    MergeParams p;
    p.gathered = reinterpret_cast<const unsigned char *>(d_gathered);
    p.world = world; p.k = k; p.row_len = seq_len;
    p.msg_bytes = flexs_screen_message_bytes(k, seq_len);
    p.off_score = (long long)k * 8;
    p.off_rows = (long long)k * 12;
    p.top_scores = d_top_scores; p.top_idx = reinterpret_cast<long long *>(d_top_idx); p.top_rows = d_top_rows;
    FX_CUDA(cudaFuncSetAttribute(screen_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FINAL_SMEM));
    screen_merge_kernel<<<1, NT, FINAL_SMEM, (cudaStream_t)stream>>>(p);
    FX_CUDA(cudaGetLastError());
    return FLEXS_OK;
}

}  // extern "C"
