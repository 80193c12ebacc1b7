// K1f (tcgen05; A = 20, k = 5, k3 = 19, F = 32, H <= 112) — the kernel of the protein shapes (GFP 237/238, AAV 90/735):
// BASELINE configs[3] and [4].  Successor of cnn_umma.cu (phase-serial, 16-byte-misaligned taps, conv1 on shared-memory
// tables), rebuilt on the architecture of cnn_k9.cu.
//
// Row mapping.  A CTA works on an *item* of 8 sequences ("streams") in lockstep.  MMA row 8c + b of a 128-row tile is
// position c (of 16) of stream b, so an 8-row group is the same position of the 8 streams, a convolution tap is
// +1 group = +1024 bytes (aligned to the SWIZZLE_128B atom: no operand fetch straddles a line, the flaw of the
// 16-byte row shift of cnn_umma.cu), and an epilogue thread sees ONE stream: GlobalMaxPooling1D (cnn.py:48) is a running
// fmaxf in registers.
//
// Rolling activation rings.  The streams advance through the sequence in chunks of 16 positions.  h1 (conv1 output) and
// h2 (conv2 output) live in two shared-memory rings of 1 KB groups ([8 streams][hi 32 ch | lo 32 ch] fp16, K-major
// SWIZZLE_128B: the canonical tcgen05 A operand).  A conv tile reads a window of 16 + taps - 1 consecutive groups that
// starts at a chunk boundary; the head of each ring is mirrored behind its end so every window is contiguous.  Nothing
// is ever recomputed (no halo): "same" padding (cnn.py:33-47; TF rule left = (k-1)/2) is the zero groups before position
// 0 and after position T-1, which the epilogues write as zeros.
//      h1 group v  <->  position v - 11      conv2 tile q reads h1 groups [16q, 16q + 20)   ring: 2 (or 3) chunks + 4 mirrored
//      h2 group u  <->  position u - 9       conv3 tile q reads h2 groups [16q, 16q + 34)   ring: 3 chunks + 18 mirrored
//
// Roles (14 warps, no CTA-wide barrier inside the loop; everything meets through mbarriers):
//   0-3   conv1: h1 = relu(T012[x0,x1,x2] + T34[x3,x4]) — two 128-byte gathers per position from L2-resident tables
//         (bias folded in; 20^3 and 20^2 entries, 1 MB + 50 KB per member), split to fp16 hi/lo, stored swizzled
//   12    issues the conv2 MMAs  (5 taps x 2 K-steps x {A_hi x [W_hi|W_lo] (N=64), A_lo x W_hi (N=32)})
//   4-7   conv2 epilogue: TMEM -> bias, ReLU, split -> h2 ring
//   13    issues the conv3 MMAs  (19 taps; the last two wait for the third chunk of the window)
//   8-11  conv3 epilogue: TMEM -> running max per stream; per item a staged merge -> pooled features
// The pooled features leave as [32][128] fp32 tiles; the dense head is cnn_k9_dense_kernel (launch_dense_tiles).
// The rings keep rolling across items, so the tensor pipe only drains at the end of the launch.
//
// CTA pairs (the shipped form, cnn_a20_pair_kernel): the two CTAs of a cluster take items 2u and 2u + 1 and run them tile
// by tile as M = 256 tcgen05.mma.cta_group::2 MMAs issued by the leader's warps 12 / 13; every barrier an issuer waits on
// lives in the leader and collects the warps of both CTAs, every barrier an issuer signals is a multicast commit.  Each SM
// fetches half of the weight columns per MMA (+15-19 % throughput).
//
// Precision: the fp16 hi/lo split of cnn_umma.cu (three products per MAC, FP32 accumulation in TMEM); activations above
// 60000/8 raise the per-stream flag and the gated FP32 kernel recomputes the batch.
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "operand_prep.h"
#include "umma2_layout.cuh"

namespace {

using namespace u2;

constexpr int KC3 = 19, NA = 20;
constexpr int NT = 448;                     // 14 warps
constexpr int W_E2 = 4, W_E3 = 8, W_I2 = 12, W_I3 = 13;
constexpr int R1M = K - 1;                                      // h1 ring: r1c (2 or 3) chunks + 4 mirrored groups
constexpr int R2C = 3, R2M = KC3 - 1, R2G = R2C * 16 + R2M;    // h2 ring: 66 groups
constexpr int GS = DSLOTS;                  // sequences per feature tile of the dense kernel

// operand blob of one member: [u2-compatible prefix: only the dense-head parts are filled] | UW2 | UW3 | scales | T012 | T34
constexpr int A20_OFF_UW2 = OFF_TBIG, A20_OFF_UW3 = A20_OFF_UW2 + K * UWTAP, A20_OFF_SCAL = A20_OFF_UW3 + KC3 * UWTAP;
constexpr int A20_OFF_T012 = (A20_OFF_SCAL + 16 + 255) / 256 * 256;
constexpr int A20_OFF_T34 = A20_OFF_T012 + NA * NA * NA * F * 4;
constexpr int A20_MEMBER_BYTES = (A20_OFF_T34 + NA * NA * F * 4 + 255) / 256 * 256;

// shared memory map (the rings need 1024-byte alignment: SWIZZLE_128B atoms)
constexpr int S_BAR = 0, S_TM = 512, S_B2S = 640, S_B3 = 768;
constexpr int S_W2 = 1024, S_W3 = S_W2 + K * UWTAP, S_R1 = S_W3 + KC3 * UWTAP;
static_assert(S_R1 % 1024 == 0, "rings must be aligned to the swizzle atom");
// behind the h1 ring of r1c chunks: the h2 ring, the merge staging buffer (4 warps x 256 maxima), the residue buffer(s)
__host__ __device__ constexpr int s_r2(int r1c) { return S_R1 + (r1c * 16 + R1M) * 1024; }
__host__ __device__ constexpr int s_stage(int r1c) { return s_r2(r1c) + R2G * 1024; }
__host__ __device__ constexpr int s_idx(int r1c) { return s_stage(r1c) + 4 * 256 * 4; }

// mbarriers
constexpr int B_IDX = 0, B_H1F = 2, B_H1E = 5, B_A2F = 8, B_A2E = 10, B_H2F = 12, B_H2E = 15, B_A3F = 18, B_A3E = 20;

struct A20Params {
    const uint8_t *idx;        // [n][L] residues
    float *feat;               // [tiles][32][128] pooled features (workspace)
    const float *weights;      // this member's fp32 block (b2, b3)
    const unsigned char *uw;   // this member's operand blob
    int *overflow_flag;
    int64_t n, n_items;
    fx::CnnOffsets o;
    int L, T, nt3, nc2, nlive2, nc1, idx_slot;
    int r1c, idx_nbuf;   // h1 ring chunks (3 when shared memory allows: conv2 runs a chunk further ahead), residue buffers
    long long *prof;  // FLEXS_UMMA_PROF=1: per-CTA wait counters of each role
};

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// A operand (both rings): K-major SWIZZLE_128B, 8-row groups of 1024 B (SBO); a K step of 16 channels is +32 B, the lo
// half +64 B, a tap +1024 B.  hi word: SBO = 1024 B, version 1, layout type 2.
constexpr uint32_t A_DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);

// taps [J0, J1) of one 128-row tile: per tap 2 K steps x {A_hi x [W_hi|W_lo], A_lo x W_hi}.  All 32 lanes call it.
template <int J0, int J1>
__device__ __forceinline__ void issue_taps(uint32_t a_win_addr, uint32_t w_addr, uint32_t d_tmem) {
    // redux.sync lands in a uniform register: the descriptor arithmetic below then stays on the uniform datapath
    // (base + immediate) instead of six R2UR moves per MMA
    const uint32_t a0 = __reduce_or_sync(0xffffffffu, desc_lo(a_win_addr, 16));
    const uint32_t b0 = __reduce_or_sync(0xffffffffu, desc_lo(w_addr, UWKC));
    d_tmem = __reduce_or_sync(0xffffffffu, d_tmem);
#pragma unroll
    for (int j = J0; j < J1; ++j) {
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

// ---- CTA-pair variant (cta_group::2; see cnn_k9.cu and tools/pair_mma_test.cu): two CTAs of a cluster run the same
// tile of their own items as ONE M = 256 MMA.  CTA rank r supplies filters [r N/2, (r+1) N/2) of the B operand from
// the same shared-memory offset of its own SM, so each SM fetches half of the weight planes per MMA.  Planes per CTA
// and layer: W64 [tap][chunk][32 rows: this rank's half of hi|lo][16 B], then W32 [tap][chunk][16 rows: half of hi].
constexpr int PW64_CH = 32 * 16, PW64_TAP = 4 * PW64_CH, PW32_CH = 16 * 16, PW32_TAP = 4 * PW32_CH;
constexpr uint32_t IDESC_P64 = (1u << 4) | ((64u >> 3) << 17) | ((256u >> 4) << 24);
constexpr uint32_t IDESC_P32 = (1u << 4) | ((32u >> 3) << 17) | ((256u >> 4) << 24);

__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t leader_addr(const void *local) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(fxd::smem_u32(local)));
    return r;
}
// arrival on the leader CTA's barrier.  CTA-scope release (the default): what it publishes is this SM's own shared
// memory / TMEM state, complete before the arrive issues and consumed by this SM's half of the pair MMA; a
// cluster-scope release costs a MEMBAR.ALL.GPU per arrival (measured on cnn_k9: +60 % kernel time).
template <bool PAIR>
__device__ __forceinline__ void arrive_at_issuer(uint64_t *local_bar, uint32_t leader_bar) {
    if (PAIR) asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(leader_bar) : "memory");
    else mbar_arrive(local_bar);
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
// commit whose arrival lands on the barrier at this offset in BOTH CTAs of the pair
template <bool PAIR>
__device__ __forceinline__ void commit_to(uint64_t *bar) {
    if (PAIR) {
        asm volatile(
            "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
            "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
            ::"r"(fxd::smem_u32(bar)), "h"((uint16_t)3) : "memory");
    } else {
        umma_commit_elect(bar);
    }
}
// pair version of issue_taps: w_addr = this layer's W64 planes, w32_addr = its W32 planes
template <int J0, int J1>
__device__ __forceinline__ void issue_taps_pair(uint32_t a_win_addr, uint32_t w_addr, uint32_t w32_addr, uint32_t d_tmem) {
    const uint32_t a0 = __reduce_or_sync(0xffffffffu, desc_lo(a_win_addr, 16));
    const uint32_t b64 = __reduce_or_sync(0xffffffffu, desc_lo(w_addr, PW64_CH));
    const uint32_t b32 = __reduce_or_sync(0xffffffffu, desc_lo(w32_addr, PW32_CH));
    d_tmem = __reduce_or_sync(0xffffffffu, d_tmem);
#pragma unroll
    for (int j = J0; j < J1; ++j) {
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

__device__ __forceinline__ void issue_idx_load(const A20Params &p, uint8_t *dst, uint64_t *bar, int64_t item) {
    const int64_t first = item * 8;
    const int64_t cnt = min((int64_t)8, p.n - first);
    const uintptr_t g0 = reinterpret_cast<uintptr_t>(p.idx) + (uintptr_t)(first * p.L);
    const uintptr_t a0 = g0 & ~(uintptr_t)15;
    const uintptr_t a1 = (g0 + (uintptr_t)(cnt * p.L) + 15) & ~(uintptr_t)15;
    const uint32_t bytes = (uint32_t)(a1 - a0);
    fxd::mbar_arrive_expect_tx(bar, bytes);
    fxd::bulk_g2s(dst, reinterpret_cast<const void *>(a0), bytes, bar);
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint4 &v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <bool PROF, bool PAIR>
__device__ __forceinline__ void a20_body(const A20Params &p) {
    auto now = [] { return PROF ? clock64() : 0ll; };  // phase timers exist only in the FLEXS_UMMA_PROF=1 instantiation
    long long pt[4] = {0, 0, 0, 0};
    const long long t_begin = now();
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + S_BAR);
    uint32_t *tmem_addr_s = reinterpret_cast<uint32_t *>(smem_raw + S_TM);
    float *b2s = reinterpret_cast<float *>(smem_raw + S_B2S);
    float *b3 = reinterpret_cast<float *>(smem_raw + S_B3);
    const int tid = threadIdx.x, lane = tid & 31;
    // broadcast from lane 0: tells the compiler the role branches below are warp-uniform, which keeps the issuing warps'
    // descriptor arithmetic on the uniform datapath
    const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int L = p.L, T = p.T;

    if (tid == 0) {
        fxd::mbar_init(&bar[B_IDX], 1); fxd::mbar_init(&bar[B_IDX + 1], 1);
        // pair: the barriers the issuers wait on live in the leader CTA and collect the warps of both CTAs
        const uint32_t nw = PAIR ? 8 : 4;
        for (int i = 0; i < 2; ++i) {
            fxd::mbar_init(&bar[B_A2F + i], 1); fxd::mbar_init(&bar[B_A2E + i], nw);
            fxd::mbar_init(&bar[B_A3F + i], 1); fxd::mbar_init(&bar[B_A3E + i], nw);
        }
        for (int i = 0; i < 3; ++i) {
            fxd::mbar_init(&bar[B_H1F + i], nw); fxd::mbar_init(&bar[B_H1E + i], 1);
            fxd::mbar_init(&bar[B_H2F + i], nw); fxd::mbar_init(&bar[B_H2E + i], 1);
        }
        fxd::fence_mbar_init();
    }
    // conv2 accumulators at columns 0 / 64, conv3 at 128 / 192
    if (wid == 0) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(fxd::smem_u32(tmem_addr_s)), "r"(256u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            tmem_alloc(tmem_addr_s, 256);
        }
    }
    // Work units.  Single CTA: unit = CTA, one item per step.  Pair: unit = cluster; its CTAs take the items 2u, 2u + 1
    // of a step and run them tile by tile in lockstep.  An item past the end (odd item count) is all padding.
    const uint32_t rank = PAIR ? cluster_rank() : 0u;
    const int64_t unit0 = PAIR ? (int64_t)(blockIdx.x >> 1) * 2 : (int64_t)blockIdx.x;
    const int64_t stride = PAIR ? (int64_t)(gridDim.x >> 1) * 2 : (int64_t)gridDim.x;
    const uint32_t lbar = PAIR ? leader_addr(bar) : 0u;  // the leader's barrier array
    const float *scal = reinterpret_cast<const float *>(p.uw + A20_OFF_SCAL);
    const float inv2s = __ldg(scal), inv3 = __ldg(scal + 1);
    for (int i = tid; i < F; i += NT) {
        b2s[i] = __ldg(p.weights + p.o.b2 + i) * ASCALE;
        b3[i] = __ldg(p.weights + p.o.b3 + i);
    }
    if (!PAIR) {
        for (int i = tid; i < (K + KC3) * UWTAP / 16; i += NT)
            reinterpret_cast<uint4 *>(smem_raw + S_W2)[i] = __ldg(reinterpret_cast<const uint4 *>(p.uw + A20_OFF_UW2) + i);
    } else {
        // this rank's column halves: rows 32 r .. 32 r + 31 of every [hi|lo] block, rows 16 r .. 16 r + 15 of its hi part
        for (int layer = 0; layer < 2; ++layer) {
            const int taps = layer ? KC3 : K;
            const uint4 *src = reinterpret_cast<const uint4 *>(p.uw + (layer ? A20_OFF_UW3 : A20_OFF_UW2));
            uint4 *d64 = reinterpret_cast<uint4 *>(smem_raw + (layer ? S_W3 : S_W2));
            uint4 *d32 = d64 + taps * PW64_TAP / 16;
            for (int i = tid; i < taps * 4 * 32; i += NT) d64[i] = __ldg(src + (i >> 5) * 64 + 32 * (int)rank + (i & 31));
            for (int i = tid; i < taps * 4 * 16; i += NT) d32[i] = __ldg(src + (i >> 4) * 64 + 16 * (int)rank + (i & 15));
        }
    }
    const uint32_t r1c = (uint32_t)p.r1c;
    const int S_R2 = s_r2(p.r1c), S_STAGE = s_stage(p.r1c), S_IDX = s_idx(p.r1c);
    for (int i = tid; i < (S_STAGE - S_R1) / 16; i += NT)
        reinterpret_cast<uint4 *>(smem_raw + S_R1)[i] = make_uint4(0, 0, 0, 0);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();  // both CTAs' barriers exist before anyone arrives on the leader's
    tc_fence_after();
    const uint32_t tmem_base = *tmem_addr_s;
    const uint32_t r1_addr = fxd::smem_u32(smem_raw + S_R1), r2_addr = fxd::smem_u32(smem_raw + S_R2);
    const uint32_t w2_addr = fxd::smem_u32(smem_raw + S_W2), w3_addr = fxd::smem_u32(smem_raw + S_W3);
    float xmax = 0.f;  // largest activation written as fp16 (range guard)

    if (wid < W_E2) {
        // =========================== conv1 producers ===========================
        // lane = 4 b + cc moves 8 channels (cc) of stream b: the 4 lanes of a stream read one whole 128-byte table row with
        // adjacent lanes (moving them 8 lanes apart made the gathers — which sit on the conv2 -> conv1 -> conv2 chain of
        // the two-chunk ring — slower by more than the stores gained: -9 % overall).  A 128-bit shared-memory store is
        // processed per quarter-warp (8 adjacent lanes = two rows b): writing "hi" from all lanes puts both rows on the
        // same four bank groups (every store replayed once — ncu source view of the first version,
        // profiles/r02_a20_ncu_summary.txt).  Odd rows therefore store their lo chunk first and their hi chunk second:
        // chunk (4 + cc) ^ b lies in the other four bank groups, and each store instruction covers all eight.
        const int b = lane >> 2, cc = lane & 3;
        const float *t012 = reinterpret_cast<const float *>(p.uw + A20_OFF_T012) + cc * 8;
        const float *t34 = reinterpret_cast<const float *>(p.uw + A20_OFF_T34) + cc * 8;
        const uint32_t row_off = (uint32_t)(b * 128), hi_off = (uint32_t)((cc ^ b) << 4), lo_off = (uint32_t)(((4 + cc) ^ b) << 4);
        const bool odd = (b & 1) != 0;
        const uint32_t first_off = odd ? lo_off : hi_off, second_off = odd ? hi_off : lo_off;
        if (tid == 0 && unit0 + rank < p.n_items) issue_idx_load(p, smem_raw + S_IDX, &bar[B_IDX], unit0 + rank);
        uint32_t g1 = 0, itc = 0;
        for (int64_t ib = unit0; ib < p.n_items; ib += stride, ++itc) {
            const int64_t item = ib + rank;
            const bool real = item < p.n_items;  // (a CTA without an item is in its last step)
            // every producer is done with the previous item: its residue buffer is free.  Two buffers: the NEXT item's
            // residues are fetched now; one buffer (long sequences, where the third h1 chunk is worth more): this item's.
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const uint32_t ibuf = p.idx_nbuf == 2 ? (itc & 1u) : 0u;
            if (p.idx_nbuf == 2) {
                if (tid == 0 && item + stride < p.n_items)
                    issue_idx_load(p, smem_raw + S_IDX + (ibuf ^ 1u) * p.idx_slot, &bar[B_IDX + (ibuf ^ 1u)], item + stride);
            } else if (tid == 0 && itc > 0 && real) {
                issue_idx_load(p, smem_raw + S_IDX, &bar[B_IDX], item);
            }
            if (real) fxd::mbar_wait(&bar[B_IDX + ibuf], (p.idx_nbuf == 2 ? (itc >> 1) : itc) & 1);
            const int nvalid = (int)min((int64_t)8, p.n - item * 8);
            const uint8_t *sidx = smem_raw + S_IDX + ibuf * p.idx_slot +
                                  ((reinterpret_cast<uintptr_t>(p.idx) + (uintptr_t)(item * 8 * L)) & 15) + (b < nvalid ? b : 0) * L;
            for (int qc = 0; qc < p.nc1; ++qc, ++g1) {
                const uint32_t use1 = g1 / r1c, slot = g1 - use1 * r1c;
                float4 ta[4][2], tb[4][2];
                bool ok[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int t = 16 * qc + wid + 4 * i - 11;
                    ok[i] = b < nvalid && t >= 0 && t < T;
                    if (ok[i]) {
                        const uint8_t *ip = sidx + t;
                        const uint32_t x0 = min((uint32_t)ip[0], 19u), x1 = min((uint32_t)ip[1], 19u), x2 = min((uint32_t)ip[2], 19u);
                        const uint32_t x3 = min((uint32_t)ip[3], 19u), x4 = min((uint32_t)ip[4], 19u);
                        const float4 *pa = reinterpret_cast<const float4 *>(t012 + (size_t)((x0 * NA + x1) * NA + x2) * F);
                        const float4 *pb = reinterpret_cast<const float4 *>(t34 + (size_t)(x3 * NA + x4) * F);
                        ta[i][0] = __ldg(pa); ta[i][1] = __ldg(pa + 1);
                        tb[i][0] = __ldg(pb); tb[i][1] = __ldg(pb + 1);
                    }
                }
                // the gathers are in flight before the ring slot is waited for: their L2 latency is off the
                // conv2 -> conv1 -> conv2 dependency chain of the two-chunk ring
                const long long w0 = now();
                if (use1 > 0) fxd::mbar_wait(&bar[B_H1E + slot], (use1 - 1) & 1);
                pt[0] += now() - w0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uint4 hi4 = make_uint4(0, 0, 0, 0), lo4 = make_uint4(0, 0, 0, 0);
                    if (ok[i]) {
                        const float x[8] = {fmaxf(ta[i][0].x + tb[i][0].x, 0.f), fmaxf(ta[i][0].y + tb[i][0].y, 0.f),
                                            fmaxf(ta[i][0].z + tb[i][0].z, 0.f), fmaxf(ta[i][0].w + tb[i][0].w, 0.f),
                                            fmaxf(ta[i][1].x + tb[i][1].x, 0.f), fmaxf(ta[i][1].y + tb[i][1].y, 0.f),
                                            fmaxf(ta[i][1].z + tb[i][1].z, 0.f), fmaxf(ta[i][1].w + tb[i][1].w, 0.f)};
                        split8(x, hi4, lo4, xmax);
                    }
                    const uint32_t gr = slot * 16u + (uint32_t)(wid + 4 * i);
                    const uint32_t row = r1_addr + gr * 1024u + row_off;
                    const uint4 first4 = odd ? lo4 : hi4, second4 = odd ? hi4 : lo4;
                    st_shared_v4(row + first_off, first4);
                    st_shared_v4(row + second_off, second4);
                    if (gr < (uint32_t)R1M) {  // head of the ring, mirrored behind its end
                        st_shared_v4(row + r1c * 16 * 1024 + first_off, first4);
                        st_shared_v4(row + r1c * 16 * 1024 + second_off, second4);
                    }
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) arrive_at_issuer<PAIR>(&bar[B_H1F + slot], lbar + (B_H1F + slot) * 8u);
            }
        }
    } else if (wid < W_E3) {
        // =========================== conv2 epilogue: accumulator -> h2 ring ===========================
        // warp lq owns TMEM lanes 32 lq .. 32 lq + 31 = positions 4 lq .. 4 lq + 3 of the chunk, all 8 streams
        const int lq = wid & 3, c = 4 * lq + (lane >> 3), b = lane & 7;
        const uint32_t tlane = tmem_base + ((uint32_t)(lq * 32) << 16);
        uint32_t g2 = 0, a2 = 0;
        for (int64_t ib = unit0; ib < p.n_items; ib += stride) {
            const int64_t item = ib + rank;
            const int nvalid = (int)min((int64_t)8, p.n - item * 8);
            for (int qc = 0; qc < p.nc2; ++qc, ++g2) {
                const uint32_t slot = g2 % 3u;
                uint4 hi4[4], lo4[4];
                if (qc < p.nlive2) {
                    const uint32_t a = a2 & 1u;
                    const long long w0 = now();
                    fxd::mbar_wait(&bar[B_A2F + a], (a2 >> 1) & 1);
                    pt[0] += now() - w0;
                    tc_fence_after();
                    const int pos = 16 * qc + c - 9;
                    const bool valid = b < nvalid && pos >= 0 && pos < T;
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        uint32_t v[16], v2[16];
                        tmem_ld16_nowait(tlane + a * 64u + (uint32_t)(hf * 16), v);
                        tmem_ld16_nowait(tlane + a * 64u + 32u + (uint32_t)(hf * 16), v2);
                        tmem_ld_wait();
                        if (hf == 1) {
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) arrive_at_issuer<PAIR>(&bar[B_A2E + a], lbar + (B_A2E + a) * 8u);  // the accumulator is in registers
                        }
#pragma unroll
                        for (int h8 = 0; h8 < 2; ++h8) {
                            // the biases as two 16-byte broadcasts (scalar loads were 32 wavefronts per thread and tile)
                            const float4 ba = *reinterpret_cast<const float4 *>(b2s + hf * 16 + h8 * 8);
                            const float4 bb = *reinterpret_cast<const float4 *>(b2s + hf * 16 + h8 * 8 + 4);
                            const float bias[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
                            float x[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const int col = h8 * 8 + i;
                                const float acc = __uint_as_float(v[col]) + __uint_as_float(v2[col]);
                                x[i] = valid ? fmaxf(fmaf(acc, inv2s, bias[i]), 0.f) : 0.f;
                            }
                            split8(x, hi4[hf * 2 + h8], lo4[hf * 2 + h8], xmax);
                        }
                    }
                    ++a2;
                } else {
                    // a chunk past the end of the sequence: the right-hand "same" padding of conv3
#pragma unroll
                    for (int j = 0; j < 4; ++j) { hi4[j] = make_uint4(0, 0, 0, 0); lo4[j] = make_uint4(0, 0, 0, 0); }
                }
                const long long w1 = now();
                if (g2 >= 3) fxd::mbar_wait(&bar[B_H2E + slot], ((g2 / 3u) - 1) & 1);  // the conv3 tiles reading this slot retired
                pt[1] += now() - w1;
                const uint32_t gr = slot * 16u + (uint32_t)c;
                const uint32_t row = r2_addr + gr * 1024u + (uint32_t)(b * 128);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    st_shared_v4(row + (uint32_t)((j ^ b) << 4), hi4[j]);
                    st_shared_v4(row + (uint32_t)(((4 + j) ^ b) << 4), lo4[j]);
                }
                if (gr < (uint32_t)R2M) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        st_shared_v4(row + R2C * 16 * 1024 + (uint32_t)((j ^ b) << 4), hi4[j]);
                        st_shared_v4(row + R2C * 16 * 1024 + (uint32_t)(((4 + j) ^ b) << 4), lo4[j]);
                    }
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) arrive_at_issuer<PAIR>(&bar[B_H2F + slot], lbar + (B_H2F + slot) * 8u);
            }
        }
    } else if (wid < W_I2) {
        // =========================== conv3 epilogue: running max per stream ===========================
        // max_t relu(a_t * inv3 + b) == relu(max_t(a_t) * inv3 + b) (inv3 > 0): per tile one add of the two accumulator
        // halves and one fmaxf per filter; scale, bias and ReLU wait for the item's end.
        const int lq = wid & 3, c = 4 * lq + (lane >> 3), b = lane & 7;
        const uint32_t tlane = tmem_base + ((uint32_t)(lq * 32) << 16) + 128u;
        const bool up8 = (lane & 8) != 0, up16 = (lane & 16) != 0;
        float *stg = reinterpret_cast<float *>(smem_raw + S_STAGE);  // [4 warps][4 fg][8 b][8 f]
        uint32_t t3 = 0;
        float mx[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) mx[j] = -INFINITY;
        for (int64_t ib = unit0; ib < p.n_items; ib += stride) {
            const int64_t item = ib + rank;
            const int nvalid = (int)min((int64_t)8, p.n - item * 8);
            for (int q = 0; q < p.nt3; ++q, ++t3) {
                const uint32_t a = t3 & 1u;
                const long long w0 = now();
                fxd::mbar_wait(&bar[B_A3F + a], (t3 >> 1) & 1);
                pt[0] += now() - w0;
                tc_fence_after();
                const bool valid = 16 * q + c < T;
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    uint32_t v[16], v2[16];
                    tmem_ld16_nowait(tlane + a * 64u + (uint32_t)(hf * 16), v);
                    tmem_ld16_nowait(tlane + a * 64u + 32u + (uint32_t)(hf * 16), v2);
                    tmem_ld_wait();
                    if (hf == 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) arrive_at_issuer<PAIR>(&bar[B_A3E + a], lbar + (B_A3E + a) * 8u);
                    }
                    if (valid) {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            mx[hf * 16 + j] = fmaxf(mx[hf * 16 + j], __uint_as_float(v[j]) + __uint_as_float(v2[j]));
                    }
                }
            }
            // GlobalMaxPooling1D: the 4 lanes of stream b merge with a halving butterfly (after the xor-8 step a lane keeps
            // filters 16 (lane bit 3) + 0..15, after the xor-16 step 8 of those), park them in this warp's 256-float slot
            // [fg][b][8]; the 4 warps meet in the double-buffered staging area; warp lq then reduces filters 8 lq .. 8 lq + 7.
            {
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
                float4 *wr = reinterpret_cast<float4 *>(stg + lq * 256 + (fg * 8 + b) * 8);
                // the two 16-byte halves of a stream's 8 maxima swap places for streams 4-7: a quarter-warp (8 streams, 32 B
                // apart) then covers all eight 16-byte bank groups per store instead of four twice (ncu source view: 2-way)
                const int sw = (b >> 2) & 1;
                wr[sw] = make_float4(k8[0], k8[1], k8[2], k8[3]);
                wr[sw ^ 1] = make_float4(k8[4], k8[5], k8[6], k8[7]);
            }
            asm volatile("bar.sync 2, 128;" ::: "memory");
            const int pr = lane >> 3, f = 8 * lq + 2 * pr;
            const float *rd = stg + (lq * 8 + b) * 8 + ((2 * pr) ^ (((b >> 2) & 1) << 2));   // the writers' half swap
            float2 t = *reinterpret_cast<const float2 *>(rd);
#pragma unroll
            for (int w4 = 1; w4 < 4; ++w4) {
                const float2 o2 = *reinterpret_cast<const float2 *>(rd + w4 * 256);
                t.x = fmaxf(t.x, o2.x); t.y = fmaxf(t.y, o2.y);
            }
            if (b < nvalid) {
                const int64_t seq = item * 8 + b;
                float *dst = p.feat + (size_t)(seq >> 7) * (F * GS) + (size_t)(seq & (GS - 1));
                dst[(size_t)f * GS] = fmaxf(fmaf(t.x, inv3, b3[f]), 0.f);
                dst[(size_t)(f + 1) * GS] = fmaxf(fmaf(t.y, inv3, b3[f + 1]), 0.f);
            }
            asm volatile("bar.sync 2, 128;" ::: "memory");   // everyone has read the staging buffer: the next item may write it
        }
    } else if (wid == W_I2) {
        // =========================== conv2 MMA issue ===========================
        // (pair: only the leader issues; its MMAs drive both SMs' tensor cores and its commits arrive in both CTAs)
        const uint32_t w2b_addr = w2_addr + K * PW64_TAP;
        uint32_t g1 = 0, a2 = 0;
        for (int64_t ib = unit0; ib < p.n_items && rank == 0; ib += stride) {
            for (int qc = 0; qc < p.nlive2; ++qc, ++g1, ++a2) {
                const uint32_t use1 = g1 / r1c, s0 = g1 - use1 * r1c, a = a2 & 1u;
                const uint32_t use1n = (g1 + 1) / r1c, s1 = (g1 + 1) - use1n * r1c;
                const long long w0 = now();
                fxd::mbar_wait_warp(&bar[B_H1F + s0], use1 & 1);
                fxd::mbar_wait_warp(&bar[B_H1F + s1], use1n & 1);  // the window ends 4 groups into the next chunk
                const long long w1 = now();
                if (a2 >= 2) fxd::mbar_wait_warp(&bar[B_A2E + a], ((a2 >> 1) - 1) & 1);
                const long long w2 = now();
                pt[0] += w1 - w0; pt[1] += w2 - w1;
                tc_fence_after();
                if (PAIR) issue_taps_pair<0, K>(r1_addr + s0 * 16u * 1024u, w2_addr, w2b_addr, tmem_base + a * 64u);
                else issue_taps<0, K>(r1_addr + s0 * 16u * 1024u, w2_addr, tmem_base + a * 64u);
                commit_to<PAIR>(&bar[B_A2F + a]);
                commit_to<PAIR>(&bar[B_H1E + s0]);
                pt[2] += now() - w2;
            }
            // the item's last h1 chunk is only ever the 4-group tail of the last window: release it as well
            commit_to<PAIR>(&bar[B_H1E + g1 % r1c]);
            ++g1;
        }
    } else {
        // =========================== conv3 MMA issue ===========================
        const uint32_t w3b_addr = w3_addr + KC3 * PW64_TAP;
        uint32_t g2 = 0, t3 = 0;
        for (int64_t ib = unit0; ib < p.n_items && rank == 0; ib += stride) {
            for (int q = 0; q < p.nt3; ++q, ++t3) {
                const uint32_t G = g2 + (uint32_t)q, s0 = G % 3u, a = t3 & 1u;
                const long long w0 = now();
                fxd::mbar_wait_warp(&bar[B_H2F + s0], (G / 3u) & 1);
                fxd::mbar_wait_warp(&bar[B_H2F + (G + 1) % 3u], ((G + 1) / 3u) & 1);
                const long long w1 = now();
                if (t3 >= 2) fxd::mbar_wait_warp(&bar[B_A3E + a], ((t3 >> 1) - 1) & 1);
                const long long w2 = now();
                tc_fence_after();
                const uint32_t win = r2_addr + s0 * 16u * 1024u, d = tmem_base + 128u + a * 64u;
                if (PAIR) issue_taps_pair<0, KC3 - 2>(win, w3_addr, w3b_addr, d);
                else issue_taps<0, KC3 - 2>(win, w3_addr, d);
                // taps 17 and 18 reach into the third chunk of the window (its first two groups)
                const long long w3 = now();
                fxd::mbar_wait_warp(&bar[B_H2F + (G + 2) % 3u], ((G + 2) / 3u) & 1);
                const long long w4 = now();
                tc_fence_after();
                if (PAIR) issue_taps_pair<KC3 - 2, KC3>(win, w3_addr, w3b_addr, d);
                else issue_taps<KC3 - 2, KC3>(win, w3_addr, d);
                commit_to<PAIR>(&bar[B_A3F + a]);
                commit_to<PAIR>(&bar[B_H2E + s0]);
                pt[0] += w1 - w0; pt[1] += w2 - w1; pt[2] += w4 - w3; pt[3] += (w3 - w2) + (now() - w4);
            }
            // the two trailing chunks of the item were never the first chunk of a window: release them here
            commit_to<PAIR>(&bar[B_H2E + (g2 + (uint32_t)p.nt3) % 3u]);
            commit_to<PAIR>(&bar[B_H2E + (g2 + (uint32_t)p.nt3 + 1) % 3u]);
            g2 += (uint32_t)p.nc2;
        }
    }
    if (xmax > 60000.f) atomicExch(p.overflow_flag, 1);
    if (PROF && p.prof != nullptr && lane == 0 && (wid == 0 || wid == W_E2 || wid == W_E3 || wid == W_I2 || wid == W_I3)) {
        const int role = wid == 0 ? 0 : wid == W_E2 ? 1 : wid == W_E3 ? 2 : wid == W_I2 ? 3 : 4;
        long long *dst = p.prof + (size_t)blockIdx.x * 24 + role * 4;
        for (int i = 0; i < 4; ++i) dst[i] = pt[i];
        if (role == 0) p.prof[(size_t)blockIdx.x * 24 + 20] = now() - t_begin;
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();  // no CTA leaves (or frees TMEM) while its partner can still reach it
    if (wid == 0) {
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
        else tmem_dealloc(tmem_base, 256);
    }
}

template <bool PROF>
__global__ void __launch_bounds__(NT, 1) cnn_a20_kernel(const A20Params p) { a20_body<PROF, false>(p); }
template <bool PROF>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NT, 1) cnn_a20_pair_kernel(const A20Params p) { a20_body<PROF, true>(p); }

static bool plan(const flexs_model *m, A20Params &p) {
    p.o = fx::cnn_offsets(m);
    p.L = m->L;
    p.T = m->L - m->K + 1;
    p.nt3 = (p.T + 15) / 16;            // conv3 tiles per item
    p.nc2 = p.nt3 + 2;                  // h2 chunks per item: the last window ends 2 groups into chunk nt3 + 1
    p.nlive2 = (p.T + 8) / 16 + 1;      // h2 chunks with a position inside the sequence (the others are zero padding)
    p.nc1 = p.nlive2 + 1;               // h1 chunks per item: a conv2 window ends 4 groups into the next chunk
    p.idx_slot = (int)align_up((size_t)8 * m->L + 32, 16);
    // Three h1 chunks when shared memory allows (conv2 runs a chunk further ahead of conv1), two residue buffers (one for
    // very long sequences).  With single-CTA MMAs the third chunk bought nothing (the operand fetch of the tensor pipe was
    // the limit either way); with CTA pairs it is worth +3.4 % on gfp237 and +1.3 % on aav735.  FLEXS_A20_R1C=2 for A/B runs.
    static const bool two = std::getenv("FLEXS_A20_R1C") && std::getenv("FLEXS_A20_R1C")[0] == '2';
    const int opts[4][2] = {{two ? 2 : 3, 2}, {two ? 2 : 3, 1}, {2, 2}, {2, 1}};
    for (const auto &o : opts) {
        p.r1c = o[0]; p.idx_nbuf = o[1];
        if ((int64_t)s_idx(p.r1c) + p.idx_nbuf * p.idx_slot + 1024 <= m->max_smem_optin) return true;
    }
    return false;
}

static int prepare(flexs_model *m) {
    if (m->a20_ready) return FLEXS_OK;
    const fx::CnnOffsets o = fx::cnn_offsets(m);
    std::vector<float> host((size_t)m->member_floats * m->M);
    FX_CUDA(cudaSetDevice(m->device));
    FX_CUDA(cudaMemcpy(host.data(), m->d_weights, host.size() * sizeof(float), cudaMemcpyDeviceToHost));
    std::vector<unsigned char> blob((size_t)A20_MEMBER_BYTES * m->M, 0);
    m->umma_weights_ok = true;
    for (int mem = 0; mem < m->M; ++mem) {
        const float *w = host.data() + (size_t)mem * m->member_floats;
        unsigned char *dst = blob.data() + (size_t)mem * A20_MEMBER_BYTES;
        const float inv2 = prep::fill_conv_planes(w + o.w2, K, dst + A20_OFF_UW2, &m->umma_weights_ok);
        const float inv3 = prep::fill_conv_planes(w + o.w3, KC3, dst + A20_OFF_UW3, &m->umma_weights_ok);
        float *scal = reinterpret_cast<float *>(dst + A20_OFF_SCAL);
        scal[0] = inv2 * ASCALE;  // the conv2 epilogue emits activations pre-scaled by ASCALE
        scal[1] = inv3;
        prep::fill_dense_head(w, o, m->H, dst, &m->umma_weights_ok);
        // conv1 (cnn.py:25-32) as two gathers: T012[x0,x1,x2] = b1 + W1[0,x0] + W1[1,x1] + W1[2,x2], T34[x3,x4] =
        // W1[3,x3] + W1[4,x4], both pre-scaled by ASCALE (a power of two: exact)
        const float *w1 = w + o.w1, *b1 = w + o.b1;  // w1 (k, A, F)
        float *t012 = reinterpret_cast<float *>(dst + A20_OFF_T012), *t34 = reinterpret_cast<float *>(dst + A20_OFF_T34);
        for (int x0 = 0; x0 < NA; ++x0)
            for (int x1 = 0; x1 < NA; ++x1)
                for (int x2 = 0; x2 < NA; ++x2) {
                    float *e = t012 + (size_t)((x0 * NA + x1) * NA + x2) * F;
                    for (int f = 0; f < F; ++f) {
                        const float v = ((w1[(0 * NA + x0) * F + f] + w1[(1 * NA + x1) * F + f]) + w1[(2 * NA + x2) * F + f]) + b1[f];
                        if (!std::isfinite(v)) m->umma_weights_ok = false;
                        e[f] = v * ASCALE;
                    }
                }
        for (int x3 = 0; x3 < NA; ++x3)
            for (int x4 = 0; x4 < NA; ++x4)
                for (int f = 0; f < F; ++f) {
                    const float v = w1[(3 * NA + x3) * F + f] + w1[(4 * NA + x4) * F + f];
                    if (!std::isfinite(v)) m->umma_weights_ok = false;
                    t34[(size_t)(x3 * NA + x4) * F + f] = v * ASCALE;
                }
    }
    if (!m->d_a20_w) FX_CUDA(cudaMalloc(&m->d_a20_w, blob.size()));
    FX_CUDA(cudaMemcpy(m->d_a20_w, blob.data(), blob.size(), cudaMemcpyHostToDevice));
    m->a20_ready = true;
    return FLEXS_OK;
}

}  // namespace

namespace fx {

bool cnn_a20_supported(const flexs_model *m) {
    if (m->kind != FLEXS_KIND_CNN || m->F != 32 || m->K != 5 || m->A != NA || m->H > DH) return false;
    if (!cnn_tiled_supported(m)) return false;  // the fp16-overflow fall-back path
    A20Params p;
    return plan(m, p);
}

int launch_cnn_a20(flexs_model *m, const uint8_t *d_idx, int64_t n, float *d_out, cudaStream_t s) {
    A20Params p;
    FX_REQUIRE(cnn_a20_supported(m) && plan(m, p), "shape not supported by the A = 20 tcgen05 kernel");
    int rc = prepare(m);
    if (rc != FLEXS_OK) return rc;
    if (!m->umma_weights_ok) return launch_cnn_tiled(m, d_idx, n, d_out, s);  // non-finite weights: fp32 path
    // feature workspace of this stream: one [32][128] fp32 tile per 128 sequences of a chunk (<= 136 MB)
    const int64_t chunk_tiles = (int64_t)m->sm_count * 56;
    const int64_t n_tiles = (n + GS - 1) / GS;
    flexs_model::StreamWs *ws = nullptr;
    rc = stream_workspace(m, s, (size_t)std::min(n_tiles, chunk_tiles) * F * GS * sizeof(float), &ws);
    if (rc != FLEXS_OK) return rc;
    p.feat = reinterpret_cast<float *>(ws->ptr);
    p.overflow_flag = ws->flag;
    const size_t smem = (size_t)s_idx(p.r1c) + p.idx_nbuf * p.idx_slot + 1024;
    static const bool prof = std::getenv("FLEXS_UMMA_PROF") && std::getenv("FLEXS_UMMA_PROF")[0] == '1';
    // CTA pairs (cta_group::2) by default; FLEXS_A20_PAIR=0, or a device that cannot co-schedule a pair, keeps the
    // single-CTA kernel
    static const bool want_pair = !(std::getenv("FLEXS_A20_PAIR") && std::getenv("FLEXS_A20_PAIR")[0] == '0');
    if (want_pair && m->a20_pair_units < 0) {
        auto pk = prof ? cnn_a20_pair_kernel<true> : cnn_a20_pair_kernel<false>;
        FX_CUDA(cudaFuncSetAttribute(pk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * (unsigned)(m->sm_count / 2)); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = smem;
        int ncl = 0;
        if (cudaOccupancyMaxActiveClusters(&ncl, pk, &cfg) != cudaSuccess) { ncl = 0; (void)cudaGetLastError(); }
        m->a20_pair_units = std::min(ncl, m->sm_count / 2);
    }
    const bool pair = want_pair && m->a20_pair_units > 0;
    auto kernel = pair ? (prof ? cnn_a20_pair_kernel<true> : cnn_a20_pair_kernel<false>)
                       : (prof ? cnn_a20_kernel<true> : cnn_a20_kernel<false>);
    FX_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int max_units = pair ? m->a20_pair_units : m->sm_count;
    FX_CUDA(cudaMemsetAsync(ws->flag, 0, sizeof(int), s));
    for (int64_t t0 = 0; t0 < n_tiles; t0 += chunk_tiles) {
        const int64_t first = t0 * GS, cnt = std::min(n - first, chunk_tiles * GS);
        p.idx = d_idx + first * m->L;
        p.n = cnt;
        p.n_items = (cnt + 7) / 8;
        const int grid = pair ? 2 * (int)std::min<int64_t>((p.n_items + 1) / 2, max_units)
                              : (int)std::min<int64_t>(p.n_items, max_units);
        for (int mem = 0; mem < m->M; ++mem) {
            p.weights = m->d_weights + (int64_t)mem * m->member_floats;
            p.uw = reinterpret_cast<const unsigned char *>(m->d_a20_w) + (size_t)mem * A20_MEMBER_BYTES;
            p.prof = nullptr;
            if (prof) {
                FX_CUDA(cudaMalloc(&p.prof, (size_t)grid * 24 * sizeof(long long)));
                FX_CUDA(cudaMemset(p.prof, 0, (size_t)grid * 24 * sizeof(long long)));
            }
            kernel<<<grid, NT, smem, s>>>(p);
            FX_CUDA(cudaGetLastError());
            m->launches += 1;
            if (prof) {
                FX_CUDA(cudaStreamSynchronize(s));
                std::vector<long long> h((size_t)grid * 24);
                FX_CUDA(cudaMemcpy(h.data(), p.prof, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
                cudaFree(p.prof);
                double a[24] = {0};
                for (int b = 0; b < grid; ++b) for (int i = 0; i < 24; ++i) a[i] += (double)h[(size_t)b * 24 + i] / grid;
                const double it = (double)p.n_items / grid;  // items per CTA
                fprintf(stderr, "[a20 prof] n=%lld grid=%d items/CTA=%.1f | cycles per item %.0f (conv3 tiles %d, conv2 tiles %d) | "
                                "conv1 warp 0: wait ring %.0f | conv2 epilogue: wait MMA %.0f, wait ring %.0f | conv3 epilogue: wait MMA %.0f | "
                                "conv2 issue: wait h1 %.0f, wait accumulator %.0f, issue %.0f | conv3 issue: wait h2 %.0f, wait accumulator %.0f, "
                                "wait 3rd chunk %.0f, issue %.0f\n",
                        (long long)cnt, grid, it, a[20] / it, p.nt3, p.nlive2, a[0] / it, a[4] / it, a[5] / it, a[8] / it, a[12] / it,
                        a[13] / it, a[14] / it, a[16] / it, a[17] / it, a[18] / it, a[19] / it);
            }
            rc = launch_dense_tiles(m, p.feat, d_out + first, p.uw, ws->flag, cnt, mem, s);
            if (rc != FLEXS_OK) return rc;
        }
    }
    // fp16 range guard: the gated FFMA kernel recomputes the batch iff the flag was raised
    return launch_cnn_tiled_gated(m, d_idx, n, d_out, ws->flag, s);
}

}  // namespace fx
