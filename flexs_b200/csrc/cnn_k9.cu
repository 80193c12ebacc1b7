// K1e (tcgen05 + L2-resident table; A = 4, k = 5, k3 = 3, F = 32, H <= 112, 8 <= L <= ~175) — the large-batch kernel of
// the 4-letter-alphabet shapes (north star: 100-mers; also the 8-mer and 14-mer configurations).
//
// With a 4-letter alphabet the first two layers of cnn.py:23-40 are a function of 9 residues:
//   h2[o, :] = relu(b2 + sum_j W2[j]^T relu(b1 + conv1(x[o+j-2 .. o+j+2])))   (conv1 valid, conv2 "same")
// so conv1 o ReLU o conv2 o ReLU is ONE 128-byte row of a 4^9-entry table (32 channels, already split into the fp16
// hi/lo operand planes of conv3 and pre-scaled), plus four small tables for the two positions at either end of a
// sequence where "same" padding truncates the window (4^7 + 4^8 entries each side).  The table (54.5 MB per ensemble
// member) is rebuilt by k9_build_kernel whenever the weights change and stays resident in the 126 MB L2; the per-row
// work of two conv layers becomes one 128-byte gather that goes straight into the tensor core's operand buffer
// (cp.async, no registers).  What is left for the tensor pipe is conv3 (3 taps) and the dense head.
//
// Row mapping.  MMA row i = 8c + b of a 128-row tile is output position 16q + c of sequence ("stream") b of an item of
// 8 sequences: an 8-row group holds the same position of 8 sequences, a tap is +1 group = +1024 B (swizzle-aligned),
// and a thread of the epilogue sees ONE sequence for the whole item, so GlobalMaxPooling1D is a running fmaxf in
// registers (bias, scale and ReLU once per item; the warps merge an item through a staged reduction, no atomics)
// instead of a REDUX round per tile and sequence.
//
// Roles (18 warps): 0-7 producers (warp w owns ring slot w: it packs the 24 residue words of its own tile -> table index
// -> 128-byte gathers into a SWIZZLE_128B operand slot), 8-15 conv3 epilogue (two sets of 4 warps on alternate items,
// running max in registers), 16-17 issue the MMAs of alternate tiles.  Eight operand slots and eight TMEM accumulators
// (64 columns each); the roles only meet through mbarriers: there is no CTA-wide barrier inside the kernel and no
// barrier at a group boundary either (the residue buffers are double-buffered and re-armed by whichever producer warp
// leaves a group last).  The pooled features leave 8 sequences at a time into a [32][128] fp32 tile per group in global
// memory; cnn_k9_dense_kernel turns those tiles into scores with two tcgen05 GEMMs per tile, its five stages as five
// warp roles with two tiles in flight.
//
// CTA pairs (the shipped form, cnn_k9_pair_kernel): two CTAs of a cluster run the same tile index of two groups as ONE
// M = 256 tcgen05.mma.cta_group::2 — each SM feeds its own 128 rows and HALF of the weight columns, which takes 14 % off
// the operand traffic that bounds the kernel (see issue_conv3_tile_pair and DESIGN.md 5.1).
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "umma2_layout.cuh"

namespace {

using namespace u2;

constexpr int NT = 576, NPROD = 8, MMAW = 16, NMMA = 2;  // warps: 0-7 producers, 8-15 epilogue, 16-17 MMA issue
constexpr int RING = 8;                 // operand slots == TMEM accumulators == producer warps
constexpr int CM = 16 + K3 - 1;         // 8-row groups per tile (16 positions + k3-1 behind them)
constexpr int SLOT = CM * 1024;         // rows of 128 B = one table entry: hi ch 0-15 | hi 16-31 | lo 0-15 | lo 16-31,
                                        // K-major SWIZZLE_128B (16-byte chunk j of row r sits at chunk j ^ (r & 7))
constexpr int GS = DSLOTS;  // sequences per group = per feature tile of the dense head

// table segments (entries of 128 B): interior | o = 0 | o = 1 | o = T-2 | o = T-1
constexpr int N_MAIN = 1 << 18, N_E7 = 1 << 14, N_E8 = 1 << 16;
constexpr int ENT_EL0 = N_MAIN, ENT_EL1 = ENT_EL0 + N_E7, ENT_ER1 = ENT_EL1 + N_E8, ENT_ER0 = ENT_ER1 + N_E8;
constexpr int N_ENT = ENT_ER0 + N_E7;
constexpr size_t TAB_BYTES = (size_t)N_ENT * 128;

struct K9Params {
    const uint8_t *idx;        // [n][L] residues of this launch
    float *feat;               // [n_groups][32][128] pooled features (workspace)
    const float *weights;      // this member's fp32 block (b3)
    const unsigned char *uw;   // this member's operand blob of cnn_umma2 (UW3, scales)
    const unsigned char *tab;  // this member's table [N_ENT][128 B]
    const int *tab_ovf;        // raised by the builder when an entry left the fp16 window
    int *overflow_flag;
    int64_t n, n_groups;
    fx::CnnOffsets o;
    int L, T, nti, idx_slot;
    long long *prof;
    int dbg;  // profiling knobs (FLEXS_UMMA_DBG bitmask, PROF build only): 1 skip the MMAs, 2 skip the gathers, 4 skip the epilogue math
};

struct DenseParams {
    const float *feat;
    float *out;
    const unsigned char *uw;
    int *overflow_flag;
    int64_t n, n_groups;
    int mem, M;
};

struct Offs {
    int mbar, tm, done, b3, uw3, idx, stage, ring;
    size_t total;
};

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

__host__ __device__ inline Offs carve(const K9Params &p) {
    Offs o;
    size_t off = 0;
    auto take = [&](size_t bytes, size_t align) {
        off = align_up(off, align);
        const size_t r = off;
        off += bytes;
        return (int)r;
    };
    o.mbar = take(64 * 8, 16); o.tm = take(16, 16); o.done = take(16, 16); o.b3 = take(F * 4, 16);
    o.uw3 = take((size_t)K3 * UWTAP, 128);
    o.idx = take((size_t)2 * p.idx_slot, 16);  // residues of two groups: this one and the next
    o.stage = take((size_t)2 * 2 * 4 * 256 * 4, 16);  // epilogue merge buffers: [set][double buffer][4 warps][256 maxima]
    o.ring = take((size_t)RING * SLOT, 1024);
    o.total = off;
    return o;
}

// ---- table builder -------------------------------------------------------------------------------------------
// One warp per entry: lane g evaluates conv1 + ReLU for channel g at the (up to five) h1 positions of the window,
// then lane f accumulates conv2 for filter f in fp32.  Entry = split_fp16(ASCALE * relu(h2)) as
// [hi: 32 channels][lo: 32 channels] halves = the 4 + 4 16-byte chunks the operand planes take.
__global__ void __launch_bounds__(256) k9_build_kernel(const float *__restrict__ w, const fx::CnnOffsets o,
                                                       unsigned char *__restrict__ tab, int *__restrict__ ovf) {
    __shared__ __align__(16) float W2s[K * F * F];
    __shared__ float W1s[K * ALPHA * F], b1s[F], b2s[F];
    __shared__ __align__(16) float h1s[8][K * F];
    const int tid = threadIdx.x, lane = tid & 31, wl = tid >> 5;
    for (int i = tid; i < K * F * F; i += 256) W2s[i] = __ldg(w + o.w2 + i);
    for (int i = tid; i < K * ALPHA * F; i += 256) W1s[i] = __ldg(w + o.w1 + i);
    if (tid < F) { b1s[tid] = __ldg(w + o.b1 + tid); b2s[tid] = __ldg(w + o.b2 + tid); }
    __syncthreads();
    __half *th = reinterpret_cast<__half *>(tab);
    bool bad = false;
    for (int e = blockIdx.x * 8 + wl; e < N_ENT; e += gridDim.x * 8) {
        int len = 9, joff = 0, code = e;
        if (e >= ENT_ER0) { len = 7; code = e - ENT_ER0; }
        else if (e >= ENT_ER1) { len = 8; code = e - ENT_ER1; }
        else if (e >= ENT_EL1) { len = 8; joff = 1; code = e - ENT_EL1; }
        else if (e >= ENT_EL0) { len = 7; joff = 2; code = e - ENT_EL0; }
        const int nv = len - (K - 1);  // h1 positions inside the window
        for (int i = 0; i < nv; ++i) {
            float v = b1s[lane];
#pragma unroll
            for (int mm = 0; mm < K; ++mm) {
                const int r = (code >> (2 * (len - 1 - (i + mm)))) & 3;
                v += W1s[(mm * ALPHA + r) * F + lane];
            }
            h1s[wl][i * F + lane] = fmaxf(v, 0.f);
        }
        __syncwarp();
        float acc = b2s[lane];
        for (int i = 0; i < nv; ++i) {
            const float *wrow = W2s + (size_t)(i + joff) * F * F + lane;
#pragma unroll
            for (int g = 0; g < F; g += 4) {
                const float4 h = *reinterpret_cast<const float4 *>(&h1s[wl][i * F + g]);
                acc = fmaf(h.x, wrow[(g + 0) * F], acc);
                acc = fmaf(h.y, wrow[(g + 1) * F], acc);
                acc = fmaf(h.z, wrow[(g + 2) * F], acc);
                acc = fmaf(h.w, wrow[(g + 3) * F], acc);
            }
        }
        const float x = fmaxf(acc, 0.f) * ASCALE;
        if (!(x <= 60000.f)) bad = true;  // beyond the fp16 window (or NaN): the fp32 kernel recomputes
        const __half hi = __float2half_rn(x);
        const __half lo = __float2half_rn(x - __half2float(hi));
        th[(size_t)e * 64 + lane] = hi;
        th[(size_t)e * 64 + 32 + lane] = lo;
        __syncwarp();
    }
    if (bad) atomicExch(ovf, 1);
}

// ---- forward: conv kernel -----------------------------------------------------------------------------------
// packed fp32 add (sm_100 FADD2): two accumulator halves of two filters per instruction
__device__ __forceinline__ void add_f32x2(float &x0, float &x1, float a0, float a1, float b0, float b1) {
    asm("{\n\t.reg .b64 ra, rb, rc;\n\t"
        "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
        "add.rn.f32x2 rc, ra, rb;\n\t"
        "mov.b64 {%0, %1}, rc;\n\t}"
        : "=f"(x0), "=f"(x1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}


// A operand: K-major SWIZZLE_128B, the canonical UMMA layout: 8-row groups of 1024 B (SBO), a K step of 16 channels is
// 32 B further into the row (start address + 32 B; the hardware applies the XOR to the final address), a tap is one
// group = +1024 B.  hi word: SBO = 1024 B, version 1, layout type 2 (bits 61-63).
constexpr uint32_t A_DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);

// one 128-row tile of conv3: taps x 2 K steps x {A_hi x [W_hi|W_lo], A_lo x W_hi}; all 32 lanes call it
__device__ __forceinline__ void issue_conv3_tile(uint32_t a_slot_addr, uint32_t w_addr, uint32_t d_tmem) {
    const uint32_t a0 = desc_lo(a_slot_addr, 16), b0 = desc_lo(w_addr, UWKC);
#pragma unroll
    for (int j = 0; j < K3; ++j) {
#pragma unroll
        for (int kp = 0; kp < 2; ++kp) {
            const uint32_t a_hi = a0 + (((uint32_t)j * 1024u + (uint32_t)kp * 32u) >> 4);
            const uint32_t a_lo = a_hi + (64u >> 4);
            const uint32_t bd = b0 + (((uint32_t)j * UWTAP + (uint32_t)(2 * kp) * UWKC) >> 4);
            umma_f16_elect(d_tmem, a_hi, A_DESC_HI, bd, DESC_HI, IDESC_N64, (j | kp) ? 1u : 0u);
            umma_f16_elect(d_tmem, a_lo, A_DESC_HI, bd, DESC_HI, IDESC_N32, 1u);
        }
    }
}

// ---- CTA-pair variant (cta_group::2): two CTAs of a cluster run the same tile index on their own 128 rows as ONE
// M = 256 MMA.  The B operand of a pair MMA is split by columns: CTA rank r supplies filters [r N/2, (r+1) N/2) from
// the SAME shared-memory offset of its own SM (tools/pair_mma_test.cu measures exactly this), so each SM reads half of
// the weight planes per tile — the operand traffic that bounds the single-CTA kernel drops from 11 KB to 9.5 KB per tile.
// Weight planes per CTA: W64 [tap][chunk][32 rows of this rank's half of hi|lo][16 B], W32 [tap][chunk][16 rows of
// this rank's half of hi][16 B].
constexpr int PW64_CH = 32 * 16, PW64_TAP = 4 * PW64_CH, PW32_CH = 16 * 16, PW32_TAP = 4 * PW32_CH;
constexpr int PW32_OFF = K3 * PW64_TAP;
constexpr uint32_t IDESC_P64 = (1u << 4) | ((64u >> 3) << 17) | ((256u >> 4) << 24);
constexpr uint32_t IDESC_P32 = (1u << 4) | ((32u >> 3) << 17) | ((256u >> 4) << 24);
static_assert(PW32_OFF + K3 * PW32_TAP <= K3 * UWTAP, "pair weight planes fit the single-CTA staging area");

__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the leader CTA's copy of a shared-memory object, usable with the shared::cluster forms
__device__ __forceinline__ uint32_t leader_addr(const void *local) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(fxd::smem_u32(local)));
    return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    // default semantics (release at CTA scope): a cluster-scope release costs MEMBAR.ALL.GPU + two error barriers per
    // arrival (measured: +60 % kernel time).  What the arrival publishes is this SM's own shared memory / TMEM state,
    // complete before the arrive issues (cp.async.wait_all + proxy fence, tcgen05.wait::ld), and it is consumed by this
    // SM's half of the pair MMA, which the leader can only start after it has seen the arrival.
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(fxd::smem_u32(dst_smem)),
                 "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_pair_elect(uint32_t d_tmem, uint32_t a_lo32, uint32_t a_hi32, uint32_t b_lo32,
                                                    uint32_t b_hi32, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "elect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %6, 0;\n\t"
        "@e tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo32), "r"(a_hi32), "r"(b_lo32), "r"(b_hi32), "r"(idesc), "r"(accumulate) : "memory");
}
// the arrival lands on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair_elect(uint64_t *bar) {
    asm volatile(
        "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
        ::"r"(fxd::smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void issue_conv3_tile_pair(uint32_t a_slot_addr, uint32_t w_addr, uint32_t d_tmem) {
    const uint32_t a0 = desc_lo(a_slot_addr, 16);
    const uint32_t b64 = desc_lo(w_addr, PW64_CH), b32 = desc_lo(w_addr + PW32_OFF, PW32_CH);
#pragma unroll
    for (int j = 0; j < K3; ++j) {
#pragma unroll
        for (int kp = 0; kp < 2; ++kp) {
            const uint32_t a_hi = a0 + (((uint32_t)j * 1024u + (uint32_t)kp * 32u) >> 4);
            const uint32_t a_lo = a_hi + (64u >> 4);
            const uint32_t bd64 = b64 + (((uint32_t)j * PW64_TAP + (uint32_t)(2 * kp) * PW64_CH) >> 4);
            const uint32_t bd32 = b32 + (((uint32_t)j * PW32_TAP + (uint32_t)(2 * kp) * PW32_CH) >> 4);
            umma_f16_pair_elect(d_tmem, a_hi, A_DESC_HI, bd64, DESC_HI, IDESC_P64, (j | kp) ? 1u : 0u);
            umma_f16_pair_elect(d_tmem, a_lo, A_DESC_HI, bd32, DESC_HI, IDESC_P32, 1u);
        }
    }
}

__device__ __forceinline__ void issue_idx_load(const K9Params &p, uint8_t *dst, uint64_t *bar, int64_t group) {
    const int64_t first = group * GS;
    const int64_t cnt = min((int64_t)GS, p.n - first);
    const uintptr_t g0 = reinterpret_cast<uintptr_t>(p.idx) + (uintptr_t)(first * p.L);
    const uintptr_t a0 = g0 & ~(uintptr_t)15;
    const uintptr_t a1 = (g0 + (uintptr_t)(cnt * p.L) + 15) & ~(uintptr_t)15;
    const uint32_t bytes = (uint32_t)(a1 - a0);
    fxd::mbar_arrive_expect_tx(bar, bytes);
    // the residues are read once: keep them from displacing the table in L2
    asm volatile(
        "{\n\t.reg .b64 pol;\n\t"
        "createpolicy.fractional.L2::evict_first.b64 pol, 1.0;\n\t"
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], pol;\n\t}"
        ::"r"(fxd::smem_u32(dst)), "l"(reinterpret_cast<const void *>(a0)), "r"(bytes), "r"(fxd::smem_u32(bar))
        : "memory");
}

template <bool PROF, bool PAIR>
__device__ __forceinline__ void k9_body(const K9Params &p) {
    auto now = [] { return PROF ? clock64() : 0ll; };  // phase timers exist only in the FLEXS_UMMA_PROF=1 instantiation
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const Offs of = carve(p);
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem_raw + of.mbar);
    uint64_t *mbar_idx = mbar, *full = mbar + 8, *empty = mbar + 16, *tfull = mbar + 24, *tempty = mbar + 40;  // tfull: [2 sets][8]
    uint32_t *tmem_addr_s = reinterpret_cast<uint32_t *>(smem_raw + of.tm);
    float *b3 = reinterpret_cast<float *>(smem_raw + of.b3);
    unsigned char *uw3 = smem_raw + of.uw3;
    uint32_t *idx_done = reinterpret_cast<uint32_t *>(smem_raw + of.done);  // producer warps finished with residue buffer b
    unsigned char *ring = smem_raw + of.ring;

    const int tid = threadIdx.x, lane = tid & 31;
    // broadcast from lane 0: tells the compiler the role branches are warp-uniform, which keeps the MMA issuers'
    // descriptor arithmetic on the uniform datapath (UTCHMMA every ~4 uniform instructions instead of ~16 with R2UR moves)
    const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int L = p.L, T = p.T, nti = p.nti;

    // Work units.  Single CTA: unit = CTA, one group of 128 sequences per step.  Pair: unit = cluster of two CTAs that
    // walk the groups 2u, 2u + 1 in lockstep (same tile index at the same time); the leader's group is the earlier and
    // therefore the fuller one, so it sets the tile count, and the peer pads with absent streams (zero rows).
    const uint32_t rank = PAIR ? cluster_rank() : 0u;
    const int64_t W = PAIR ? 2 : 1;
    const int64_t gb0 = (int64_t)(PAIR ? (blockIdx.x >> 1) : blockIdx.x) * W;
    const int64_t gstep = (int64_t)(PAIR ? (gridDim.x >> 1) : gridDim.x) * W;
    auto mine_of = [&](int64_t gb) -> int {  // sequences of this CTA's group of the unit step at gb (0: no group)
        const int64_t left = p.n - (gb + rank) * GS;
        return (int)max((int64_t)0, min((int64_t)GS, left));
    };
    auto lead_of = [&](int64_t gb) -> int { return (int)min((int64_t)GS, p.n - gb * GS); };

    if (tid == 0) {
        fxd::mbar_init(&mbar_idx[0], 1); fxd::mbar_init(&mbar_idx[1], 1);
        idx_done[0] = 0; idx_done[1] = 0;
        for (int i = 0; i < RING; ++i) {
            // pair: the leader's "operands ready" / "accumulator drained" barriers collect both CTAs' arrivals
            fxd::mbar_init(&full[i], PAIR ? 2 : 1); fxd::mbar_init(&empty[i], 1);
            fxd::mbar_init(&tfull[i], 1); fxd::mbar_init(&tfull[8 + i], 1); fxd::mbar_init(&tempty[i], PAIR ? 8 : 4);
        }
        fxd::fence_mbar_init();
    }
    if (wid == 0) { if (PAIR) tmem_alloc_pair(tmem_addr_s, 512); else tmem_alloc(tmem_addr_s, 512); }
    const float inv3 = __ldg(reinterpret_cast<const float *>(p.uw + OFF_SCAL) + 1);
    for (int i = tid; i < F; i += NT) b3[i] = __ldg(p.weights + p.o.b3 + i);
    if (!PAIR) {
        for (int i = tid; i < K3 * UWTAP / 16; i += NT)
            reinterpret_cast<uint4 *>(uw3)[i] = __ldg(reinterpret_cast<const uint4 *>(p.uw + OFF_UW3) + i);
    } else {
        // this rank's column halves of the weight planes: rows 32 r .. 32 r + 31 of [hi|lo], rows 16 r .. 16 r + 15 of hi
        const uint4 *src = reinterpret_cast<const uint4 *>(p.uw + OFF_UW3);
        for (int i = tid; i < K3 * 4 * 32; i += NT) {
            const int blk = i >> 5, row = i & 31;
            reinterpret_cast<uint4 *>(uw3)[i] = __ldg(src + blk * 64 + 32 * (int)rank + row);
        }
        for (int i = tid; i < K3 * 4 * 16; i += NT) {
            const int blk = i >> 4, row = i & 15;
            reinterpret_cast<uint4 *>(uw3 + PW32_OFF)[i] = __ldg(src + blk * 64 + 16 * (int)rank + row);
        }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();  // both CTAs' barriers exist before anyone arrives on the peer's
    tc_fence_after();
    const uint32_t tmem_base = *tmem_addr_s;
    const uint32_t ring_addr = fxd::smem_u32(ring), uw3_addr = fxd::smem_u32(uw3);
    if (tid == 0) {
        if (mine_of(gb0) > 0) issue_idx_load(p, smem_raw + of.idx, &mbar_idx[0], gb0 + rank);
        if (mine_of(gb0 + gstep) > 0) issue_idx_load(p, smem_raw + of.idx + p.idx_slot, &mbar_idx[1], gb0 + gstep + rank);
    }

    uint32_t kt = 0;  // tiles before the current group: tile k uses slot = accumulator = k & 7
    uint32_t gi = 0;  // groups before the current one: residue buffer = gi & 1, its barrier's parity = (gi >> 1) & 1
    long long pt[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const long long t_begin = now();

    if (wid < NPROD) {
        // =========================== producers: residues -> table rows -> operand slot ===========================
        // lane = 8 rr + j moves 16-byte chunk j of the rows of streams rr and rr + 4: eight adjacent lanes read one whole
        // 128-byte entry (one L2 request of 4 sectors) and write the 8 chunks of one swizzled row = 8 bank groups.
        const int rr = lane >> 3, jch = lane & 7;
        const uint32_t slot_addr = ring_addr + (uint32_t)wid * SLOT;
        const uint32_t dst0 = (uint32_t)(rr * 128 + ((jch ^ rr) << 4));
        const uint32_t dst1 = (uint32_t)((rr + 4) * 128 + ((jch ^ (rr + 4)) << 4));
        const unsigned char *tabj = p.tab + jch * 16;
        // lane 3 b + w (< 24) packs word q - 1 + w of stream b for the tile: 16 residues, 2 bits each, first residue in the
        // top bits.  Every warp builds the 24 words of its own tile, so the producer warps never wait for each other.
        const int wsb = lane / 3, www = lane - 3 * wsb;
        const uint32_t full_mine = PAIR ? leader_addr(&full[wid]) : 0u;
        for (int64_t gb = gb0; gb < p.n_groups; gb += gstep, ++gi) {
            const int64_t first = (gb + rank) * GS;
            const int s_grp = mine_of(gb);
            const uint32_t ntiles = (uint32_t)(((lead_of(gb) + 7) >> 3) * nti);
            const uint32_t buf = gi & 1u;
            if (s_grp > 0) fxd::mbar_wait(&mbar_idx[buf], (gi >> 1) & 1);  // (a CTA without a group is in its last step)
            const uint32_t sidx = fxd::smem_u32(smem_raw + of.idx) + buf * (uint32_t)p.idx_slot +
                                  (uint32_t)((reinterpret_cast<uintptr_t>(p.idx) + (uintptr_t)(first * L)) & 15);
            for (uint32_t tl = ((uint32_t)wid - kt) & 7u; tl < ntiles; tl += NPROD) {
                const uint32_t use = (kt + tl) >> 3;
                const long long q0 = now();
                if (use > 0) fxd::mbar_wait(&empty[wid], (use - 1) & 1);  // the MMAs that read this slot retired
                const long long q1 = now();
                const int item = (int)tl / nti, q = (int)tl - item * nti;
                if (PROF && (p.dbg & 2)) {
                    if (lane == 0) { if (PAIR) mbar_arrive_remote(full_mine); else mbar_arrive(&full[wid]); }
                    continue;
                }
                // residues 16(q-1) .. 16(q+2)-1 of the lane's two streams (word index + 1 in the padded array)
                const int sl0 = item * 8 + rr, sl1 = sl0 + 4;
                const bool ok0 = sl0 < s_grp, ok1 = sl1 < s_grp;
                uint32_t word = 0;
                {
                    const int sq = item * 8 + wsb, wv = q - 1 + www;
                    if (lane < 24 && sq < s_grp && wv >= 0 && wv * 16 < L) {
                        // 16 residue bytes at an arbitrary byte offset: five aligned words, funnel-shifted into four
                        const uint32_t a = sidx + (uint32_t)(sq * L + wv * 16), sh = (a & 3u) * 8u;
                        uint32_t w[5];
#pragma unroll
                        for (int i = 0; i < 5; ++i)
                            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w[i]) : "r"((a & ~3u) + 4u * i));
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const uint32_t x = __funnelshift_r(w[i], w[i + 1], sh) & 0x03030303u;
                            // residues at bits 0, 8, 16, 24 -> one byte r0 r1 r2 r3 (a multiply gathers them in the top byte)
                            word = (word << 8) | ((x * 0x40100401u) >> 24);
                        }
                        const int cnt = min(16, L - wv * 16);  // the last word of a row: drop the next row's residues
                        word &= 0xffffffffu << (2 * (16 - cnt));
                    }
                }
                const uint32_t u0 = __shfl_sync(0xffffffffu, word, 3 * rr), u1 = __shfl_sync(0xffffffffu, word, 3 * rr + 1);
                const uint32_t u2 = __shfl_sync(0xffffffffu, word, 3 * rr + 2), v0 = __shfl_sync(0xffffffffu, word, 3 * rr + 12);
                const uint32_t v1 = __shfl_sync(0xffffffffu, word, 3 * rr + 13), v2 = __shfl_sync(0xffffffffu, word, 3 * rr + 14);
                // the window of input row c (h2 position o = 16 q + c - 1) starts at residue o - 2, i.e. 2 (c + 13) bits
                // into w0:w1:w2 — a compile-time shift
                auto code_of = [](int c, uint32_t x0, uint32_t x1, uint32_t x2) -> uint32_t {
                    if (c < 3) return __funnelshift_l(x1, x0, 2 * (c + 13)) >> 14;
                    if (c <= 10) return (x1 << (2 * (c - 3))) >> 14;
                    return __funnelshift_l(x2, x1, 2 * (c - 3)) >> 14;
                };
                if (q > 0 && 16 * q + 16 <= T - 3 && item * 8 + 7 < s_grp) {
                    // interior tile of a complete item (4 of 6 tiles of a 100-mer): every row is a full 9-residue window
                    // of an existing sequence — code, one multiply-add for the address, copy
#pragma unroll
                    for (int c = 0; c < CM; ++c) {
                        const unsigned char *src0 = tabj + (size_t)code_of(c, u0, u1, u2) * 128;
                        const unsigned char *src1 = tabj + (size_t)code_of(c, v0, v1, v2) * 128;
                        const uint32_t grp = slot_addr + (uint32_t)(c * 1024);
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(grp + dst0), "l"(src0) : "memory");
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(grp + dst1), "l"(src1) : "memory");
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < CM; ++c) {
                        const int o = 16 * q + c - 1;
                        const uint32_t code0 = code_of(c, u0, u1, u2), code1 = code_of(c, v0, v1, v2);
                        // truncated windows: o = 0, 1 read zeros in front (the code is already that of the short window),
                        // o = T-2, T-1 drop the residues past the end
                        uint32_t base = 0, rsh = 0;
                        if (o <= 1) base = (o == 0 ? ENT_EL0 : ENT_EL1);
                        if (o >= T - 2) { base = (o == T - 2) ? ENT_ER1 : ENT_ER0; rsh = (o == T - 2) ? 2 : 4; }
                        const bool pos_ok = o >= 0 && o < T;
                        const bool r0 = pos_ok && ok0, r1 = pos_ok && ok1;
                        const unsigned char *src0 = tabj + (r0 ? (size_t)(base + (code0 >> rsh)) * 128 : (size_t)0);
                        const unsigned char *src1 = tabj + (r1 ? (size_t)(base + (code1 >> rsh)) * 128 : (size_t)0);
                        const uint32_t grp = slot_addr + (uint32_t)(c * 1024);
                        // src-size 0 -> cp.async zero-fills ("same" padding of conv3, rows past the sequence, absent streams)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(grp + dst0), "l"(src0), "r"(r0 ? 16u : 0u) : "memory");
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(grp + dst1), "l"(src1), "r"(r1 ? 16u : 0u) : "memory");
                    }
                }
                asm volatile("cp.async.wait_all;" ::: "memory");
                fence_async_smem();
                __syncwarp();
                if (lane == 0) { if (PAIR) mbar_arrive_remote(full_mine); else mbar_arrive(&full[wid]); }
                if (PROF && tid == 0) { pt[2] += q1 - q0; pt[3] += now() - q1; }
            }
            kt += ntiles;
            if (PROF && tid == 0) pt[7] += ntiles;
            // The last of the eight warps to finish the group hands its residue buffer to the bulk copy of the group after
            // the next one (nobody waits: the copy has a whole group's time to land).
            __syncwarp();
            if (lane == 0) {
                if (atomicAdd(&idx_done[buf], 1u) == NPROD - 1) {
                    atomicExch(&idx_done[buf], 0u);
                    if (mine_of(gb + 2 * gstep) > 0) {
                        fence_async_smem();  // the warps' reads of the buffer precede the async-proxy write
                        issue_idx_load(p, smem_raw + of.idx + buf * p.idx_slot, &mbar_idx[buf], gb + 2 * gstep + rank);
                    }
                }
            }
        }
    } else if (wid < MMAW) {
        // =========================== conv3 epilogue: running max per sequence ===========================
        // Two sets of 4 warps; set `par` owns the items of parity par with all their tiles, so an item is merged inside
        // one set.  The eight accumulators are shared (tile k uses k & 7), which means a set does NOT see every use of an
        // accumulator — and a parity wait cannot tell phase n from n + 2 — so each set has its own "accumulator full"
        // barriers: the MMA warp commits a tile to the barrier of the set that owns it, and the set flips one private
        // parity bit per accumulator.  Warp (lq, par) owns TMEM lanes 32 lq .. 32 lq + 31 (positions 4 lq .. 4 lq + 3 of
        // the 8 streams) and all 32 filters.
        // max_t relu(a_t * inv3 + b) == relu(max_t(a_t) * inv3 + b) (inv3 > 0): per tile only one packed add of the two
        // accumulator halves and one fmaxf per filter; scale, bias and ReLU wait for the item's end.
        const int lq = wid & 3, par = (wid >> 2) & 1;
        // TMEM lane 32 lq + lane is MMA row 8 c + b: position c of stream b
        const int c = 4 * lq + (lane >> 3), b = lane & 7;
        const uint32_t tlane = tmem_base + ((uint32_t)(lq * 32) << 16);
        const bool up8 = (lane & 8) != 0, up16 = (lane & 16) != 0;
        float *stage_s = reinterpret_cast<float *>(smem_raw + of.stage) + par * 2 * 4 * 256;  // [2 buffers][4 warps][4 fg][8 b][8 f]
        uint64_t *tfull_mine = tfull + 8 * par;
        const uint32_t tempty_lead = PAIR ? leader_addr(&tempty[0]) : 0u;
        uint32_t nflush = 0;    // items of this set so far: staging buffer = nflush & 1
        uint32_t phase_bits = 0;  // bit a: parity of the next phase of this set's barrier for accumulator a
        float mx[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) mx[j] = -INFINITY;
        for (int64_t gb = gb0; gb < p.n_groups; gb += gstep, ++gi) {
            const int64_t g = gb + rank;
            const int s_grp = mine_of(gb);
            const int nitems = (lead_of(gb) + 7) >> 3;  // pair: the peer visits (and frees) every tile of the unit
            const uint32_t ntiles = (uint32_t)(nitems * nti);
            float *featT = p.feat + (size_t)g * F * GS;  // [32][128] tile of this group, written 8 sequences at a time
            // one tile: accumulator -> registers -> running maxima
            auto visit = [&](uint32_t k, int q) {
                const uint32_t acc = k & 7u;
                const long long w0 = now();
                fxd::mbar_wait(&tfull_mine[acc], (phase_bits >> acc) & 1u);
                phase_bits ^= 1u << acc;
                const long long w1 = now();
                tc_fence_after();
                const bool valid = !(PROF && (p.dbg & 4)) && 16 * q + c < T;  // rows past the sequence: last tile only
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    uint32_t v[16], v2[16];
                    tmem_ld16_nowait(tlane + acc * 64u + (uint32_t)(hf * 16), v);
                    tmem_ld16_nowait(tlane + acc * 64u + 32u + (uint32_t)(hf * 16), v2);
                    tmem_ld_wait();
                    if (hf == 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {  // the accumulator is in registers: hand it back
                            if (PAIR) mbar_arrive_remote(tempty_lead + acc * 8u); else mbar_arrive(&tempty[acc]);
                        }
                    }
                    if (valid) {
#pragma unroll
                        for (int j = 0; j < 16; j += 2) {
                            float a0, a1;
                            add_f32x2(a0, a1, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v2[j]),
                                      __uint_as_float(v2[j + 1]));
                            mx[hf * 16 + j] = fmaxf(mx[hf * 16 + j], a0);
                            mx[hf * 16 + j + 1] = fmaxf(mx[hf * 16 + j + 1], a1);
                        }
                    }
                }
                if (PROF && tid == 8 * 32) { pt[4] += w1 - w0; pt[1] += now() - w1; }
            };
            // GlobalMaxPooling1D, step 1: the 4 lanes of stream b merge with a halving butterfly — after the xor-8 step a
            // lane keeps filters 16 (lane bit 3) + 0..15, after the xor-16 step 8 of those (filters 8 fg .. 8 fg + 7) —
            // and park them in this warp's 256-float slot of a staging buffer, [fg][b][8]
            auto butterfly_to = [&](float *slot) {
                float k16[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float send = up8 ? mx[j] : mx[j + 16], keep = up8 ? mx[j + 16] : mx[j];
                    k16[j] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 8));
                }
                float k8[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float send = up16 ? k16[j] : k16[j + 8], keep = up16 ? k16[j + 8] : k16[j];
                    k8[j] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 16));
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) mx[j] = -INFINITY;
                const int fg = (up8 ? 2 : 0) + (up16 ? 1 : 0);
                float4 *wr = reinterpret_cast<float4 *>(slot + (fg * 8 + b) * 8);
                // the two 16-byte halves of a stream's 8 maxima swap places for streams 4-7: a quarter-warp (8 streams, 32 B
                // apart) then covers all eight 16-byte bank groups per store instead of four twice (ncu source view: 2-way)
                const int sw = (b >> 2) & 1;
                wr[sw] = make_float4(k8[0], k8[1], k8[2], k8[3]);
                wr[sw ^ 1] = make_float4(k8[4], k8[5], k8[6], k8[7]);
            };
            // Step 2: the 4 warps of the set (4 positions each) meet in the set's double-buffered staging area — plain
            // stores, a 128-thread barrier, warp lq reduces filters 8 lq .. 8 lq + 7.  Step 3: scale, bias, ReLU, one plain
            // store per feature: no atomics anywhere.
            for (int item = par; item < nitems; item += 2) {
                const uint32_t k0 = kt + (uint32_t)(item * nti);
                for (int q = 0; q < nti; ++q) visit(k0 + (uint32_t)q, q);
                const long long f0 = now();
                float *stg = stage_s + (int)(nflush & 1) * 4 * 256;
                butterfly_to(stg + lq * 256);
                const long long f1 = now();
                if (par == 0) asm volatile("bar.sync 3, 128;" ::: "memory");
                else asm volatile("bar.sync 4, 128;" ::: "memory");
                const long long f2 = now();
                const int pr = lane >> 3, f = 8 * lq + 2 * pr, sl = item * 8 + b;
                const float *rd = stg + (lq * 8 + b) * 8 + ((2 * pr) ^ (((b >> 2) & 1) << 2));   // the writers' half swap
                float2 t = *reinterpret_cast<const float2 *>(rd);
#pragma unroll
                for (int w4 = 1; w4 < 4; ++w4) {
                    const float2 o2 = *reinterpret_cast<const float2 *>(rd + w4 * 256);
                    t.x = fmaxf(t.x, o2.x); t.y = fmaxf(t.y, o2.y);
                }
                // a warp writes 8 rows of 32 bytes (8 sequences of one filter): whole sectors
                if (sl < s_grp) {
                    featT[f * GS + sl] = fmaxf(fmaf(t.x, inv3, b3[f]), 0.f);
                    featT[(f + 1) * GS + sl] = fmaxf(fmaf(t.y, inv3, b3[f + 1]), 0.f);
                }
                ++nflush;
                if (PROF && tid == 8 * 32) { pt[8] += f1 - f0; pt[9] += f2 - f1; pt[10] += now() - f2; }
            }
            kt += ntiles;
        }
    } else {
        // =========================== MMA issuers ===========================
        // Two warps take alternate tiles: issuing a tile (two barrier waits, 12 MMAs, two commits) costs more issue time
        // than the tensor pipe needs to run it, and tiles are independent (own slot, own accumulator; a commit tracks
        // the MMAs of the committing thread).
        // Pair: only the leader CTA issues; its MMAs drive both SMs' tensor cores and its commits arrive in both CTAs.
        const uint32_t me = (uint32_t)(wid - MMAW);
        for (int64_t gb = gb0; gb < p.n_groups && rank == 0; gb += gstep) {
            const uint32_t ntiles = (uint32_t)(((lead_of(gb) + 7) >> 3) * nti);
            for (uint32_t tl = (me - kt) & (NMMA - 1); tl < ntiles; tl += NMMA) {
                const uint32_t k = kt + tl, s = k & 7u, use = k >> 3;
                const long long m0 = now();
                fxd::mbar_wait_warp(&full[s], use & 1);
                const long long m1 = now();
                if (use > 0) fxd::mbar_wait_warp(&tempty[s], (use - 1) & 1);  // accumulator drained by the epilogue
                if (PROF && lane == 0 && wid == MMAW) { pt[5] += m1 - m0; pt[6] += now() - m1; }
                tc_fence_after();
                const uint32_t set_bar = 8u * ((tl / (uint32_t)nti) & 1u) + s;  // the epilogue set that owns the tile's item
                if (PAIR) {
                    if (!(PROF && (p.dbg & 1))) issue_conv3_tile_pair(ring_addr + s * SLOT, uw3_addr, tmem_base + s * 64u);
                    umma_commit_pair_elect(&tfull[set_bar]);
                    umma_commit_pair_elect(&empty[s]);
                } else {
                    if (!(PROF && (p.dbg & 1))) issue_conv3_tile(ring_addr + s * SLOT, uw3_addr, tmem_base + s * 64u);
                    umma_commit_elect(&tfull[set_bar]);
                    umma_commit_elect(&empty[s]);
                }
            }
            kt += ntiles;
        }
    }
    if (PROF && tid == 0) pt[0] = now() - t_begin;
    if (blockIdx.x == 0 && tid == 0 && __ldg(p.tab_ovf) != 0) atomicExch(p.overflow_flag, 1);
    if (PROF && p.prof != nullptr && (tid == 0 || tid == 8 * 32 || tid == MMAW * 32))
        for (int i = 0; i < 12; ++i)
            if (pt[i]) atomicAdd(reinterpret_cast<unsigned long long *>(&p.prof[(size_t)blockIdx.x * 12 + i]), (unsigned long long)pt[i]);
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();  // no CTA leaves (or frees TMEM) while its partner can still reach it
    if (wid == 0) { if (PAIR) tmem_dealloc_pair(tmem_base, 512); else tmem_dealloc(tmem_base, 512); }
}

template <bool PROF>
__global__ void __launch_bounds__(NT, 1) cnn_k9_kernel(const K9Params p) { k9_body<PROF, false>(p); }
template <bool PROF>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NT, 1) cnn_k9_pair_kernel(const K9Params p) { k9_body<PROF, true>(p); }

// ---- forward: dense head kernel -----------------------------------------------------------------------------
// One [32][128] feature tile per step: Dense(H,relu) -> Dense(H,relu) -> Dense(1) -> nan_to_num -> ensemble accumulate
// (cnn.py:49-52, keras_model.py:77-79, ensemble.py:54-59) as two tcgen05 GEMMs with the fp16 hi/lo split, weights
// staged once per CTA.  The five stages of a tile are five warp roles that meet only through mbarriers, so two tiles
// are in flight and the tensor pipe runs GEMM 2 of tile i while the epilogue warps turn tile i + 1 into its operand
// (the first version ran the stages one after the other between CTA barriers: 5900 cycles per tile, 1600 of them MMA):
//   0-3    features (global, coalesced; next tile prefetched in registers) -> fp16 hi/lo planes X1[i & 1]
//   16     GEMM 1: X1 x B1 -> TMEM columns 0..223
//   4-11   epilogue 1: bias, ReLU, split -> planes X2[i & 1]
//   17     GEMM 2: X2 x B2 -> TMEM columns 256..479
//   12-15  epilogue 2: bias, ReLU, dot with the output weights, nan_to_num, ensemble accumulate -> scores
constexpr int DNT = 576, DW_E1 = 4, DW_E2 = 12, DW_G1 = 16, DW_G2 = 17;
constexpr int D_OFF_MBAR = 0, D_OFF_TM = 128, D_OFF_DV = 256;
constexpr int D_OFF_B1 = 2048, D_OFF_B2 = D_OFF_B1 + 4 * DBK, D_OFF_X1 = D_OFF_B2 + 14 * DBK;
constexpr int D_OFF_X2 = D_OFF_X1 + 2 * 8 * DPLANE, D_SMEM = D_OFF_X2 + 2 * 28 * DPLANE + 1024;
static_assert(D_OFF_DV + DV_FLOATS * 4 <= D_OFF_B1 && D_OFF_X1 % 1024 == 0, "dense kernel shared-memory map");
// mbarriers
constexpr int DB_X1F = 0, DB_X1E = 2, DB_A1F = 4, DB_A1E = 5, DB_X2F = 6, DB_X2E = 8, DB_A2F = 10, DB_A2E = 11;

__global__ void __launch_bounds__(DNT, 1) cnn_k9_dense_kernel(const DenseParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + D_OFF_MBAR);
    uint32_t *tmem_addr_s = reinterpret_cast<uint32_t *>(smem_raw + D_OFF_TM);
    float *dv = reinterpret_cast<float *>(smem_raw + D_OFF_DV);  // bd1 * ASCALE | bd2 | wd3 | inv_d1s, inv_d2, bd3
    unsigned char *db1 = smem_raw + D_OFF_B1, *db2 = smem_raw + D_OFF_B2;
    unsigned char *dx1 = smem_raw + D_OFF_X1, *dx2 = smem_raw + D_OFF_X2;
    const int tid = threadIdx.x, lane = tid & 31, wid = __shfl_sync(0xffffffffu, tid >> 5, 0);
    if (tid == 0) {
        for (int b = 0; b < 2; ++b) {
            fxd::mbar_init(&bar[DB_X1F + b], 4); fxd::mbar_init(&bar[DB_X1E + b], 1);
            fxd::mbar_init(&bar[DB_X2F + b], 8); fxd::mbar_init(&bar[DB_X2E + b], 1);
        }
        fxd::mbar_init(&bar[DB_A1F], 1); fxd::mbar_init(&bar[DB_A1E], 8);
        fxd::mbar_init(&bar[DB_A2F], 1); fxd::mbar_init(&bar[DB_A2E], 4);
        fxd::fence_mbar_init();
    }
    if (wid == 0) tmem_alloc(tmem_addr_s, 512);
    {
        const float *gdv = reinterpret_cast<const float *>(p.uw + OFF_DV);
        for (int i = tid; i < DV_FLOATS; i += DNT) dv[i] = __ldg(gdv + i);
        for (int i = tid; i < 4 * DBK / 16; i += DNT)
            reinterpret_cast<uint4 *>(db1)[i] = __ldg(reinterpret_cast<const uint4 *>(p.uw + OFF_DB1) + i);
        for (int i = tid; i < 14 * DBK / 16; i += DNT)
            reinterpret_cast<uint4 *>(db2)[i] = __ldg(reinterpret_cast<const uint4 *>(p.uw + OFF_DB2) + i);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_addr_s;
    const int64_t g0 = blockIdx.x, gs = gridDim.x;
    float xmax = 0.f;  // largest activation written as fp16 (range guard)

    if (wid < DW_E1) {
        // ---- features -> X1 planes: thread = sequence slot, 32 coalesced loads per tile ----
        const int slot = tid;
        float cur[F], nxt[F];
        auto load = [&](int64_t g, float (&v)[F]) {
            const bool ok = g < p.n_groups && g * GS + slot < p.n;  // slots past the batch hold stale features
            const float *src = p.feat + (size_t)g * F * GS + slot;
#pragma unroll
            for (int f = 0; f < F; ++f) v[f] = ok ? __ldg(src + f * GS) : 0.f;
        };
        // The tiles were written by the conv kernel up to a chunk (136 MB > L2) ago: they come from DRAM.  One bulk prefetch
        // per tile pulls it into L2 three tiles ahead; the register prefetch one tile ahead then only pays the L2 latency
        // (measured without: the whole pipeline waited ~2000 cycles per tile for this warp's loads).
        auto prefetch_l2 = [&](int64_t g) {
            if (tid == 0 && g < p.n_groups)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.feat + (size_t)g * F * GS), "r"(F * GS * 4) : "memory");
        };
        prefetch_l2(g0 + gs); prefetch_l2(g0 + 2 * gs);
        load(g0, cur);
        uint32_t i = 0;
        for (int64_t g = g0; g < p.n_groups; g += gs, ++i) {
            const uint32_t b = i & 1u;
            prefetch_l2(g + 3 * gs);
            load(g + gs, nxt);
            if (i >= 2) fxd::mbar_wait(&bar[DB_X1E + b], ((i >> 1) - 1) & 1);  // GEMM 1 of tile i - 2 has read this buffer
            unsigned char *x1 = dx1 + b * 8 * DPLANE;
#pragma unroll
            for (int cchunk = 0; cchunk < 4; ++cchunk) {
                float x[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) x[q] = cur[cchunk * 8 + q] * ASCALE;
                uint4 hi4, lo4;
                split8(x, hi4, lo4, xmax);
                *reinterpret_cast<uint4 *>(x1 + (size_t)cchunk * DPLANE + slot * 16) = hi4;
                *reinterpret_cast<uint4 *>(x1 + (size_t)(4 + cchunk) * DPLANE + slot * 16) = lo4;
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar[DB_X1F + b]);
#pragma unroll
            for (int f = 0; f < F; ++f) cur[f] = nxt[f];
        }
    } else if (wid < DW_E2) {
        // ---- epilogue 1: accumulator 1 -> bias, ReLU, split -> X2 planes ----
        const int lq = wid & 3, half = (wid - DW_E1) >> 2, slot = 32 * lq + lane;
        const uint32_t tl = tmem_base + ((uint32_t)(lq * 32) << 16);
        const float inv_d1s = dv[3 * DH];
        uint32_t i = 0;
        for (int64_t g = g0; g < p.n_groups; g += gs, ++i) {
            const uint32_t b = i & 1u;
            fxd::mbar_wait(&bar[DB_A1F], i & 1);
            if (i >= 2) fxd::mbar_wait(&bar[DB_X2E + b], ((i >> 1) - 1) & 1);  // GEMM 2 of tile i - 2 has read this buffer
            tc_fence_after();
            unsigned char *x2 = dx2 + b * 28 * DPLANE;
            // the TMEM loads of chunk c + 1 are in flight while chunk c is converted (the stages of a tile form one
            // dependency chain G1 -> E1 -> G2 -> E2 with two tiles in flight: its latency, not its issue slots, sets the rate)
            uint32_t va[2][8], vb[2][8];
            tmem_ld8_nowait(tl + (uint32_t)(half * 7 * 8), va[0]);
            tmem_ld8_nowait(tl + (uint32_t)(DH + half * 7 * 8), vb[0]);
#pragma unroll
            for (int c7 = 0; c7 < 7; ++c7) {
                const int cchunk = half * 7 + c7, cur = c7 & 1;
                const float4 b0 = *reinterpret_cast<const float4 *>(dv + cchunk * 8);
                const float4 b1 = *reinterpret_cast<const float4 *>(dv + cchunk * 8 + 4);
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                tmem_ld_wait();
                if (c7 < 6) {
                    tmem_ld8_nowait(tl + (uint32_t)((cchunk + 1) * 8), va[cur ^ 1]);
                    tmem_ld8_nowait(tl + (uint32_t)(DH + (cchunk + 1) * 8), vb[cur ^ 1]);
                } else {  // the accumulator is in registers: GEMM 1 of the next tile may overwrite it
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar[DB_A1E]);
                }
                float x[8];
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    x[q] = fmaxf(fmaf(__uint_as_float(va[cur][q]) + __uint_as_float(vb[cur][q]), inv_d1s, bb[q]), 0.f);
                uint4 hi4, lo4;
                split8(x, hi4, lo4, xmax);
                *reinterpret_cast<uint4 *>(x2 + (size_t)cchunk * DPLANE + slot * 16) = hi4;
                *reinterpret_cast<uint4 *>(x2 + (size_t)(14 + cchunk) * DPLANE + slot * 16) = lo4;
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar[DB_X2F + b]);
        }
    } else if (wid < DW_G1) {
        // ---- epilogue 2: accumulator 2 -> bias, ReLU, dot with the output weights -> score ----
        const int lq = wid & 3, slot = 32 * lq + lane;
        const uint32_t tl = tmem_base + ((uint32_t)(lq * 32) << 16) + 256u;
        const float inv_d2 = dv[3 * DH + 1], bd3v = dv[3 * DH + 2];
        uint32_t i = 0;
        for (int64_t g = g0; g < p.n_groups; g += gs, ++i) {
            fxd::mbar_wait(&bar[DB_A2F], i & 1);
            tc_fence_after();
            float sum = 0.f;
            uint32_t va[2][8], vb[2][8];
            tmem_ld8_nowait(tl, va[0]);
            tmem_ld8_nowait(tl + (uint32_t)DH, vb[0]);
#pragma unroll
            for (int cchunk = 0; cchunk < 14; ++cchunk) {
                const int cur = cchunk & 1;
                const float4 b0 = *reinterpret_cast<const float4 *>(dv + DH + cchunk * 8);
                const float4 b1 = *reinterpret_cast<const float4 *>(dv + DH + cchunk * 8 + 4);
                const float4 w0 = *reinterpret_cast<const float4 *>(dv + 2 * DH + cchunk * 8);
                const float4 w1 = *reinterpret_cast<const float4 *>(dv + 2 * DH + cchunk * 8 + 4);
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
                tmem_ld_wait();
                if (cchunk < 13) {   // next chunk's loads fly while this one is reduced
                    tmem_ld8_nowait(tl + (uint32_t)((cchunk + 1) * 8), va[cur ^ 1]);
                    tmem_ld8_nowait(tl + (uint32_t)(DH + (cchunk + 1) * 8), vb[cur ^ 1]);
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float d2 = fmaxf(fmaf(__uint_as_float(va[cur][q]) + __uint_as_float(vb[cur][q]), inv_d2, bb[q]), 0.f);
                    sum = fmaf(d2, ww[q], sum);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar[DB_A2E]);
            // Dense(1) bias, nan_to_num (keras_model.py:77), ensemble mean (ensemble.py:24)
            const int64_t seq = g * GS + slot;
            if (seq < p.n) {
                const float y = fxd::nan_to_num(sum + bd3v);
                float tot = (p.mem == 0) ? y : p.out[seq] + y;
                if (p.M > 1 && p.mem == p.M - 1) tot = tot / (float)p.M;
                p.out[seq] = tot;
            }
        }
    } else if (wid == DW_G1) {
        uint32_t i = 0;
        for (int64_t g = g0; g < p.n_groups; g += gs, ++i) {
            const uint32_t b = i & 1u;
            fxd::mbar_wait_warp(&bar[DB_X1F + b], (i >> 1) & 1);
            if (i >= 1) fxd::mbar_wait_warp(&bar[DB_A1E], (i - 1) & 1);
            tc_fence_after();
            issue_dense_layer<2, 4>(fxd::smem_u32(dx1 + b * 8 * DPLANE), fxd::smem_u32(db1), tmem_base);
            umma_commit_elect(&bar[DB_X1E + b]);
            umma_commit_elect(&bar[DB_A1F]);
        }
    } else {
        uint32_t i = 0;
        for (int64_t g = g0; g < p.n_groups; g += gs, ++i) {
            const uint32_t b = i & 1u;
            fxd::mbar_wait_warp(&bar[DB_X2F + b], (i >> 1) & 1);
            if (i >= 1) fxd::mbar_wait_warp(&bar[DB_A2E], (i - 1) & 1);
            tc_fence_after();
            issue_dense_layer<7, 14>(fxd::smem_u32(dx2 + b * 28 * DPLANE), fxd::smem_u32(db2), tmem_base + 256u);
            umma_commit_elect(&bar[DB_X2E + b]);
            umma_commit_elect(&bar[DB_A2F]);
        }
    }
    if (xmax > 60000.f) atomicExch(p.overflow_flag, 1);
    tc_fence_before();
    __syncthreads();
    if (wid == 0) tmem_dealloc(tmem_base, 512);
}

static bool plan(const flexs_model *m, K9Params &p) {
    p.o = fx::cnn_offsets(m);
    p.L = m->L;
    p.T = m->L - m->K + 1;
    p.nti = (p.T + 15) / 16;
    p.idx_slot = (int)align_up((size_t)GS * m->L + 32, 16);
    return (int64_t)carve(p).total + 1024 <= m->max_smem_optin && D_SMEM <= m->max_smem_optin;
}

static int prepare(flexs_model *m, cudaStream_t s) {
    int rc = fx::prepare_cnn_umma2(m);
    if (rc != FLEXS_OK) return rc;
    if (m->k9_ready) return FLEXS_OK;
    FX_CUDA(cudaSetDevice(m->device));
    if (!m->d_k9_tab) FX_CUDA(cudaMalloc(&m->d_k9_tab, TAB_BYTES * m->M));
    if (!m->d_k9_ovf) FX_CUDA(cudaMalloc(&m->d_k9_ovf, sizeof(int)));
    FX_CUDA(cudaMemsetAsync(m->d_k9_ovf, 0, sizeof(int), s));
    const fx::CnnOffsets o = fx::cnn_offsets(m);
    for (int mem = 0; mem < m->M; ++mem) {
        k9_build_kernel<<<m->sm_count * 4, 256, 0, s>>>(m->d_weights + (int64_t)mem * m->member_floats, o,
                                                        reinterpret_cast<unsigned char *>(m->d_k9_tab) + (size_t)mem * TAB_BYTES,
                                                        m->d_k9_ovf);
        FX_CUDA(cudaGetLastError());
        m->launches += 1;
    }
    // the model may be driven from several streams (score_host alternates two): the table must be complete
    // before any of them reads it
    FX_CUDA(cudaStreamSynchronize(s));
    m->k9_ready = true;
    return FLEXS_OK;
}

}  // namespace

namespace fx {

bool cnn_k9_supported(const flexs_model *m) {
    if (m->kind != FLEXS_KIND_CNN || m->F != 32 || m->K != 5 || m->A != 4 || m->H > DH) return false;
    if (m->L - m->K + 1 < 4) return false;  // the four truncated-window tables assume distinct positions 0, 1, T-2, T-1
    if (!cnn_umma2_supported(m)) return false;  // operand blob, small-batch kernel and fp32 fall-back
    K9Params p;
    return plan(m, p);
}

int launch_dense_tiles(flexs_model *m, const float *feat, float *out, const unsigned char *uw, int *flag, int64_t n,
                       int mem, cudaStream_t s) {
    const int64_t n_groups = (n + GS - 1) / GS;
    FX_REQUIRE(m->H <= DH && D_SMEM <= m->max_smem_optin, "dense-head kernel needs H <= 112");
    FX_CUDA(cudaFuncSetAttribute(cnn_k9_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, D_SMEM));
    DenseParams dp{feat, out, uw, flag, n, n_groups, mem, m->M};
    const int grid = (int)std::min<int64_t>(n_groups, m->sm_count);
    cnn_k9_dense_kernel<<<grid, DNT, D_SMEM, s>>>(dp);
    FX_CUDA(cudaGetLastError());
    m->launches += 1;
    return FLEXS_OK;
}

int launch_cnn_k9(flexs_model *m, const uint8_t *d_idx, int64_t n, float *d_out, cudaStream_t s) {
    K9Params p;
    FX_REQUIRE(cnn_k9_supported(m) && plan(m, p), "shape not supported by the table + tcgen05 kernel");
    int rc = prepare(m, s);
    if (rc != FLEXS_OK) return rc;
    if (!m->umma_weights_ok) return launch_cnn_tiled(m, d_idx, n, d_out, s);  // non-finite weights: fp32 path
    // feature workspace of this stream: one [32][128] fp32 tile per group of a chunk.  A chunk is a whole number of
    // waves of the persistent grid (56 groups per SM, ~1.06 M sequences, 136 MB).
    const int64_t chunk_groups = (int64_t)m->sm_count * 56;
    const int64_t n_groups = (n + GS - 1) / GS;
    const size_t ws_bytes = (size_t)std::min(n_groups, chunk_groups) * F * GS * sizeof(float);
    flexs_model::StreamWs *ws = nullptr;
    rc = stream_workspace(m, s, ws_bytes, &ws);
    if (rc != FLEXS_OK) return rc;
    static const bool prof = std::getenv("FLEXS_UMMA_PROF") && std::getenv("FLEXS_UMMA_PROF")[0] == '1';
    p.dbg = std::getenv("FLEXS_UMMA_DBG") ? std::atoi(std::getenv("FLEXS_UMMA_DBG")) : 0;
    p.feat = reinterpret_cast<float *>(ws->ptr);
    p.tab_ovf = m->d_k9_ovf;
    p.overflow_flag = ws->flag;
    const size_t smem = carve(p).total + 1024;
    // CTA pairs (cta_group::2) by default; FLEXS_K9_PAIR=0, or a device that cannot co-schedule a pair, keeps the
    // single-CTA kernel
    static const bool want_pair = !(std::getenv("FLEXS_K9_PAIR") && std::getenv("FLEXS_K9_PAIR")[0] == '0');
    if (want_pair && m->k9_pair_units < 0) {
        auto pk = prof ? cnn_k9_pair_kernel<true> : cnn_k9_pair_kernel<false>;
        FX_CUDA(cudaFuncSetAttribute(pk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * (unsigned)(m->sm_count / 2)); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = smem;
        int ncl = 0;
        if (cudaOccupancyMaxActiveClusters(&ncl, pk, &cfg) != cudaSuccess) { ncl = 0; (void)cudaGetLastError(); }
        m->k9_pair_units = std::min(ncl, m->sm_count / 2);
    }
    const bool pair = want_pair && m->k9_pair_units > 0;
    auto kernel = pair ? (prof ? cnn_k9_pair_kernel<true> : cnn_k9_pair_kernel<false>)
                       : (prof ? cnn_k9_kernel<true> : cnn_k9_kernel<false>);
    FX_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int max_units = pair ? m->k9_pair_units : m->sm_count;
    FX_CUDA(cudaFuncSetAttribute(cnn_k9_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, D_SMEM));
    FX_CUDA(cudaMemsetAsync(ws->flag, 0, sizeof(int), s));
    for (int64_t g0 = 0; g0 < n_groups; g0 += chunk_groups) {
        const int64_t first = g0 * GS, cnt = std::min(n - first, chunk_groups * GS);
        p.idx = d_idx + first * m->L;
        p.n = cnt;
        p.n_groups = (cnt + GS - 1) / GS;
        const int dgrid = (int)std::min<int64_t>(p.n_groups, m->sm_count);
        const int grid = pair ? 2 * (int)std::min<int64_t>((p.n_groups + 1) / 2, max_units)
                              : (int)std::min<int64_t>(p.n_groups, max_units);
        for (int mem = 0; mem < m->M; ++mem) {
            p.weights = m->d_weights + (int64_t)mem * m->member_floats;
            p.uw = reinterpret_cast<const unsigned char *>(m->d_umma2_w) + (size_t)mem * UW_MEMBER_BYTES;
            p.tab = reinterpret_cast<const unsigned char *>(m->d_k9_tab) + (size_t)mem * TAB_BYTES;
            p.prof = nullptr;
            if (prof) {
                FX_CUDA(cudaMalloc(&p.prof, (size_t)grid * 12 * sizeof(long long)));
                FX_CUDA(cudaMemset(p.prof, 0, (size_t)grid * 12 * sizeof(long long)));
            }
            kernel<<<grid, NT, smem, s>>>(p);
            FX_CUDA(cudaGetLastError());
            DenseParams dp{p.feat, d_out + first, p.uw, ws->flag, cnt, p.n_groups, mem, m->M};
            cnn_k9_dense_kernel<<<dgrid, DNT, D_SMEM, s>>>(dp);
            FX_CUDA(cudaGetLastError());
            m->launches += 2;
            if (prof) {
                FX_CUDA(cudaStreamSynchronize(s));
                std::vector<long long> h((size_t)grid * 12);
                FX_CUDA(cudaMemcpy(h.data(), p.prof, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
                cudaFree(p.prof);
                double a[12] = {0};
                for (int b = 0; b < grid; ++b) for (int i = 0; i < 12; ++i) a[i] += (double)h[(size_t)b * 12 + i] / grid;
                const double nt = a[7] > 0 ? a[7] : 1;
                fprintf(stderr, "[k9 prof] n=%lld grid=%d tiles/CTA=%.0f | cycles per tile %.0f | producer warp 0 (per own tile = 1/8 "
                                "of tiles): wait slot %.0f, gather %.0f | epilogue warp 8 (per own tile = 1/2 of tiles): wait MMA %.0f, "
                                "load+max %.0f, per item: butterfly %.0f, set barrier %.0f, merge+store %.0f | MMA warp 16 (per own tile = 1/2 of tiles): "
                                "wait operands %.0f, wait accumulator %.0f\n",
                        (long long)cnt, grid, a[7], a[0] / nt, a[2] / nt * 8, a[3] / nt * 8, a[4] / nt * 2, a[1] / nt * 2,
                        a[8] / nt * 2 * p.nti, a[9] / nt * 2 * p.nti, a[10] / nt * 2 * p.nti, a[5] / nt * 2, a[6] / nt * 2);
            }
        }
    }
    // fp16 range guard: the gated FFMA kernel recomputes the batch iff the flag was raised
    return launch_cnn_tiled_gated(m, d_idx, n, d_out, ws->flag, s);
}

}  // namespace fx
