// K1e (tcgen05 + L2-resident table; A = 4, k = 5, k3 = 3, F = 32, H <= 112, T >= 16) — the large-batch kernel of the
// north-star shape (100-mers over a 4-letter alphabet).
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
// 8 sequences: a core matrix holds the same position of 8 sequences, a tap is +1 core matrix = +128 B (aligned), and
// a thread of the epilogue sees ONE sequence for the whole item, so GlobalMaxPooling1D is a running fmaxf in registers
// (two shuffles and four shared atomics per item instead of a REDUX round per tile and sequence).
//
// Roles (17 warps): 0-7 producers (warp w owns ring slot w: residues -> table index -> gathers), 8-15 conv3 epilogue,
// 16 issues the MMAs.  Eight operand slots and eight TMEM accumulators (64 columns each) keep all three busy; every
// 128 sequences the pipeline drains and the dense head (u2::dense_head_umma) runs on the idle ring memory.
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

constexpr int NT = 544, NPROD = 8, MMAW = 16;
constexpr int RING = 8;                 // operand slots == TMEM accumulators == producer warps
constexpr int CM = 16 + K3 - 1;         // core matrices per tile and plane (16 positions + k3-1 behind them)
constexpr int PLANE = CM * 128;         // bytes per channel-chunk plane of a slot
constexpr int SLOT = 8 * PLANE;         // 4 hi + 4 lo planes
constexpr int GS = DSLOTS, SBP = GS + 4;  // sequences per dense-head batch, padded feature row
static_assert(RING * SLOT >= DS_TOTAL, "dense scratch must fit the operand ring");

// table segments (entries of 128 B): interior | o = 0 | o = 1 | o = T-2 | o = T-1
constexpr int N_MAIN = 1 << 18, N_E7 = 1 << 14, N_E8 = 1 << 16;
constexpr int ENT_EL0 = N_MAIN, ENT_EL1 = ENT_EL0 + N_E7, ENT_ER1 = ENT_EL1 + N_E8, ENT_ER0 = ENT_ER1 + N_E8;
constexpr int N_ENT = ENT_ER0 + N_E7;
constexpr size_t TAB_BYTES = (size_t)N_ENT * 128;

struct K9Params {
    const uint8_t *idx;
    float *out;
    const float *weights;
    const unsigned char *uw;   // cnn_umma2's operand blob (UW3, scales, dense planes)
    const unsigned char *tab;  // [M][N_ENT][128 B]
    const int *tab_ovf;        // raised by the builder when an entry left the fp16 window
    int *overflow_flag;
    int64_t n, n_groups, member_floats, uw_member_bytes;
    fx::CnnOffsets o;
    int M, L, T, nti, idx_slot;
    long long *prof;
};

struct Offs {
    int mbar, tm, b3, uw3, i0, i1, feat, ring;
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
    o.mbar = take(64 * 8, 16); o.tm = take(16, 16); o.b3 = take(F * 4, 16);
    o.uw3 = take((size_t)K3 * UWTAP, 128);
    o.i0 = take(p.idx_slot, 16); o.i1 = take(p.idx_slot, 16);
    o.feat = take((size_t)F * SBP * 4, 16);
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

// ---- forward -------------------------------------------------------------------------------------------------
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

__global__ void __launch_bounds__(NT, 1) cnn_k9_kernel(const K9Params p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const Offs of = carve(p);
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem_raw + of.mbar);
    uint64_t *mbar_idx = mbar, *dbar = mbar + 2, *full = mbar + 8, *empty = mbar + 16, *tfull = mbar + 24, *tempty = mbar + 32;
    uint32_t *tmem_addr_s = reinterpret_cast<uint32_t *>(smem_raw + of.tm);
    float *b3 = reinterpret_cast<float *>(smem_raw + of.b3);
    unsigned char *uw3 = smem_raw + of.uw3;
    float *featT = reinterpret_cast<float *>(smem_raw + of.feat);
    unsigned char *ring = smem_raw + of.ring;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int L = p.L, T = p.T, nti = p.nti;

    if (tid == 0) {
        fxd::mbar_init(&mbar_idx[0], 1); fxd::mbar_init(&mbar_idx[1], 1);
        fxd::mbar_init(dbar, 1);
        for (int i = 0; i < RING; ++i) {
            fxd::mbar_init(&full[i], 1); fxd::mbar_init(&empty[i], 1);
            fxd::mbar_init(&tfull[i], 1); fxd::mbar_init(&tempty[i], 8);
        }
        fxd::fence_mbar_init();
    }
    if (wid == 0) tmem_alloc(tmem_addr_s, 512);
    for (int i = tid; i < F * SBP; i += NT) featT[i] = 0.f;
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_addr_s;
    const uint32_t ring_addr = fxd::smem_u32(ring), uw3_addr = fxd::smem_u32(uw3);

    uint32_t kt = 0;   // tiles done (running over groups and members): slot = accumulator = kt & 7
    uint32_t gi = 0;   // groups done: residue buffer = gi & 1
    uint32_t dph = 0;  // completed phases of the dense-head barrier
    float xmax = 0.f;  // largest dense-head activation written as fp16 (range guard)
    long long pt[8] = {0, 0, 0, 0, 0, 0, 0, 0};

    for (int mem = 0; mem < p.M; ++mem) {
        const float *w = p.weights + (int64_t)mem * p.member_floats;
        const unsigned char *uw = p.uw + (int64_t)mem * p.uw_member_bytes;
        const unsigned char *tab = p.tab + (size_t)mem * TAB_BYTES;
        const float inv3 = __ldg(reinterpret_cast<const float *>(uw + OFF_SCAL) + 1);
        __syncthreads();
        for (int i = tid; i < F; i += NT) b3[i] = __ldg(w + p.o.b3 + i);
        for (int i = tid; i < K3 * UWTAP / 16; i += NT)
            reinterpret_cast<uint4 *>(uw3)[i] = __ldg(reinterpret_cast<const uint4 *>(uw + OFF_UW3) + i);
        fence_async_smem();
        __syncthreads();
        if (tid == 0 && (int64_t)blockIdx.x < p.n_groups)
            issue_idx_load(p, smem_raw + ((gi & 1) ? of.i1 : of.i0), &mbar_idx[gi & 1], blockIdx.x);

        for (int64_t g = blockIdx.x; g < p.n_groups; g += gridDim.x, ++gi) {
            const int64_t first = g * GS;
            const int s_grp = (int)min((int64_t)GS, p.n - first);
            const uint32_t ntiles = (uint32_t)(((s_grp + 7) >> 3) * nti);
            const int buf = gi & 1;
            const long long tg0 = clock64();

            if (wid < NPROD) {
                // =========================== producers: residues -> table rows -> operand slot ===========================
                const int b = lane & 7, qq = lane >> 3;
                const uint32_t slot_addr = ring_addr + (uint32_t)wid * SLOT;
                fxd::mbar_wait(&mbar_idx[buf], (gi >> 1) & 1);
                const uint8_t *sidx = smem_raw + (buf ? of.i1 : of.i0) +
                                      ((reinterpret_cast<uintptr_t>(p.idx) + (uintptr_t)(first * L)) & 15);
                for (uint32_t tl = ((uint32_t)wid - kt) & 7u; tl < ntiles; tl += NPROD) {
                    const uint32_t use = (kt + tl) >> 3;
                    const long long q0 = clock64();
                    if (use > 0) fxd::mbar_wait(&empty[wid], (use - 1) & 1);  // the MMAs that read this slot retired
                    const long long q1 = clock64();
                    const int item = (int)tl / nti, q = (int)tl - item * nti;
                    const int sl = item * 8 + b;
                    const uint8_t *sq = sidx + sl * L;
                    const bool stream_ok = sl < s_grp;
#pragma unroll
                    for (int bt = 0; bt < (CM + 3) / 4; ++bt) {
                        // lane (b, qq) looks up the entry of input row c = 4 bt + qq of stream b: h2 position o = 16 q + c - 1
                        const int c = 4 * bt + qq;
                        const int o = 16 * q + c - 1;
                        int ent = -1;  // -1: zero row ("same" padding of conv3, rows past the sequence, absent streams)
                        if (c < CM && stream_ok && o >= 0 && o < T) {
                            int start = o - 2, len = 9, base = 0;
                            if (o == 0) { start = 0; len = 7; base = ENT_EL0; }
                            else if (o == 1) { start = 0; len = 8; base = ENT_EL1; }
                            else if (o == T - 2) { len = 8; base = ENT_ER1; }
                            else if (o == T - 1) { len = 7; base = ENT_ER0; }
                            int code = 0;
#pragma unroll
                            for (int mm = 0; mm < 9; ++mm)
                                if (mm < len) code = code * 4 + (sq[start + mm] & 3);
                            ent = base + code;
                        }
                        // core matrix c2 = 4 bt + cc: its 8 rows (streams) x 8 chunks of 16 B = 2 cp.async per lane:
                        // lane (b, qq) moves chunk qq (hi plane qq) and chunk 4 + qq (lo plane qq) of stream b's row
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) {
                            const int c2 = 4 * bt + cc;
                            if (c2 >= CM) break;
                            const int e = __shfl_sync(0xffffffffu, ent, cc * 8 + b);
                            const unsigned char *src = tab + (e >= 0 ? (size_t)e * 128 : (size_t)0) + qq * 16;
                            const uint32_t nbytes = e >= 0 ? 16u : 0u;  // 0 -> cp.async zero-fills
                            const uint32_t dst = slot_addr + (uint32_t)(qq * PLANE + c2 * 128 + b * 16);
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + 4u * PLANE), "l"(src + 64),
                                         "r"(nbytes) : "memory");
                        }
                    }
                    asm volatile("cp.async.wait_all;" ::: "memory");
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full[wid]);
                    if (tid == 0) { pt[2] += q1 - q0; pt[3] += clock64() - q1; }
                }
            } else if (wid < 16) {
                // =========================== conv3 epilogue: bias, ReLU, running max per sequence ===========================
                const int lq = wid & 3, ch = (wid >> 2) & 1;
                // TMEM lane 32 lq + lane is MMA row 8 c + b: position c of stream b
                const int c = 4 * lq + (lane >> 3), b = lane & 7, qq = lane >> 3;
                const uint32_t tlane = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(ch * 16);
                float bb[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) bb[j] = b3[ch * 16 + j];
                float mx[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) mx[j] = 0.f;
                int q = 0, item = 0;
                for (uint32_t tl = 0; tl < ntiles; ++tl) {
                    const uint32_t k = kt + tl, acc = k & 7u;
                    const long long w0 = clock64();
                    fxd::mbar_wait(&tfull[acc], (k >> 3) & 1);
                    if (tid == 8 * 32) pt[4] += clock64() - w0;
                    tc_fence_after();
                    uint32_t v[16], v2[16];
                    tmem_ld16_nowait(tlane + acc * 64u, v);
                    tmem_ld16_nowait(tlane + acc * 64u + 32u, v2);
                    tmem_ld_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[acc]);  // the accumulator is in registers: hand it back
                    const bool valid = 16 * q + c < T;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float a = __uint_as_float(v[j]) + __uint_as_float(v2[j]);
                        const float y = fmaxf(fmaf(a, inv3, bb[j]), 0.f);
                        mx[j] = fmaxf(mx[j], valid ? y : 0.f);
                    }
                    if (++q == nti) {
                        // GlobalMaxPooling1D: the 4 lanes of stream b (and the 4 warps of this half) merge their maxima
                        const int sl = item * 8 + b;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            float t = fmaxf(mx[j], __shfl_xor_sync(0xffffffffu, mx[j], 8));
                            t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, 16));
                            if ((j >> 2) == qq && sl < s_grp)  // y >= 0: uint order == float order
                                atomicMax(reinterpret_cast<unsigned int *>(featT) + (size_t)(ch * 16 + j) * SBP + sl,
                                          __float_as_uint(t));
                            mx[j] = 0.f;
                        }
                        q = 0; ++item;
                    }
                }
            } else {
                // =========================== MMA issuer ===========================
                {   // prefetch the next group's residues into the other buffer: its last readers (the producers of the
                    // previous group) finished before the dense-head barrier this warp has passed
                    const int64_t next = g + gridDim.x;
                    if (lane == 0 && next < p.n_groups)
                        issue_idx_load(p, smem_raw + (buf ? of.i0 : of.i1), &mbar_idx[buf ^ 1], next);
                    __syncwarp();
                }
                for (uint32_t tl = 0; tl < ntiles; ++tl) {
                    const uint32_t k = kt + tl, s = k & 7u, use = k >> 3;
                    const long long m0 = clock64();
                    fxd::mbar_wait(&full[s], use & 1);
                    const long long m1 = clock64();
                    if (use > 0) fxd::mbar_wait(&tempty[s], (use - 1) & 1);  // accumulator drained by the epilogue
                    if (lane == 0) { pt[5] += m1 - m0; pt[6] += clock64() - m1; }
                    tc_fence_after();
                    issue_conv_tile<K3, PLANE>(ring_addr + s * SLOT, uw3_addr, tmem_base + s * 64u);
                    umma_commit_elect(&tfull[s]);
                    umma_commit_elect(&empty[s]);
                }
            }
            kt += ntiles;
            // ---- drain, dense head on the group's features, reset featT ----
            tc_fence_before();
            __syncthreads();
            const long long tg1 = clock64();
            dense_head_umma<NT, MMAW>(ring, featT, SBP, GS, s_grp, uw, tmem_base, dbar, dph, xmax, mem, p.M, p.out,
                                      [first](int sl) { return (long long)(first + sl); });
            for (int i = tid; i < F * SBP; i += NT) featT[i] = 0.f;
            __syncthreads();
            if (tid == 0) { pt[0] += tg1 - tg0; pt[1] += clock64() - tg1; pt[7] += ntiles; }
        }
    }
    if (xmax > 60000.f) atomicExch(p.overflow_flag, 1);
    if (blockIdx.x == 0 && tid == 0 && __ldg(p.tab_ovf) != 0) atomicExch(p.overflow_flag, 1);
    if (p.prof != nullptr && (tid == 0 || tid == 8 * 32 || tid == MMAW * 32))
        for (int i = 0; i < 8; ++i)
            if (pt[i]) atomicAdd(reinterpret_cast<unsigned long long *>(&p.prof[(size_t)blockIdx.x * 8 + i]), (unsigned long long)pt[i]);
    tc_fence_before();
    __syncthreads();
    if (wid == 0) tmem_dealloc(tmem_base, 512);
}

static bool plan(const flexs_model *m, K9Params &p) {
    p.o = fx::cnn_offsets(m);
    p.M = m->M;
    p.L = m->L;
    p.T = m->L - m->K + 1;
    p.nti = (p.T + 15) / 16;
    p.member_floats = m->member_floats;
    p.uw_member_bytes = UW_MEMBER_BYTES;
    p.idx_slot = (int)align_up((size_t)GS * m->L + 32, 16);
    return (int64_t)carve(p).total + 1024 <= m->max_smem_optin;
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
    if (m->L - m->K + 1 < 16) return false;  // short sequences: cnn_umma2's dense row packing wastes fewer MMA rows
    if (!cnn_umma2_supported(m)) return false;  // operand blob, small-batch kernel and fp32 fall-back
    K9Params p;
    return plan(m, p);
}

int launch_cnn_k9(flexs_model *m, const uint8_t *d_idx, int64_t n, float *d_out, cudaStream_t s) {
    K9Params p;
    FX_REQUIRE(cnn_k9_supported(m) && plan(m, p), "shape not supported by the table + tcgen05 kernel");
    int rc = prepare(m, s);
    if (rc != FLEXS_OK) return rc;
    if (!m->umma_weights_ok) return launch_cnn_tiled(m, d_idx, n, d_out, s);  // non-finite weights: fp32 path
    p.idx = d_idx; p.out = d_out; p.weights = m->d_weights; p.n = n;
    p.uw = reinterpret_cast<const unsigned char *>(m->d_umma2_w);
    p.tab = reinterpret_cast<const unsigned char *>(m->d_k9_tab);
    p.tab_ovf = m->d_k9_ovf;
    p.overflow_flag = m->d_flag;
    p.n_groups = (n + GS - 1) / GS;
    const size_t smem = carve(p).total + 1024;
    const int grid = (int)std::min<int64_t>(p.n_groups, m->sm_count);
    static const bool prof = std::getenv("FLEXS_UMMA_PROF") && std::getenv("FLEXS_UMMA_PROF")[0] == '1';
    p.prof = nullptr;
    if (prof) {
        FX_CUDA(cudaMalloc(&p.prof, (size_t)grid * 8 * sizeof(long long)));
        FX_CUDA(cudaMemset(p.prof, 0, (size_t)grid * 8 * sizeof(long long)));
    }
    FX_CUDA(cudaMemsetAsync(m->d_flag, 0, sizeof(int), s));
    FX_CUDA(cudaFuncSetAttribute(cnn_k9_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cnn_k9_kernel<<<grid, NT, smem, s>>>(p);
    FX_CUDA(cudaGetLastError());
    m->launches += 1;
    if (prof) {
        FX_CUDA(cudaStreamSynchronize(s));
        std::vector<long long> h((size_t)grid * 8);
        FX_CUDA(cudaMemcpy(h.data(), p.prof, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        cudaFree(p.prof);
        double a[8] = {0};
        for (int b = 0; b < grid; ++b) for (int i = 0; i < 8; ++i) a[i] += (double)h[(size_t)b * 8 + i] / grid;
        const double nt = a[7] > 0 ? a[7] : 1;
        fprintf(stderr, "[k9 prof] n=%lld grid=%d tiles/CTA=%.0f | cycles per tile: pipeline %.0f, dense+drain %.0f | "
                        "producer warp 0 (per own tile = 1/8 of tiles): wait slot %.0f, gather %.0f | epilogue warp 8: wait MMA %.0f | "
                        "MMA warp: wait operands %.0f, wait accumulator %.0f\n",
                (long long)n, grid, a[7], a[0] / nt, a[1] / nt, a[2] / nt * 8, a[3] / nt * 8, a[4] / nt, a[5] / nt, a[6] / nt);
    }
    // fp16 range guard: the gated FFMA kernel recomputes the batch iff the flag was raised
    return launch_cnn_tiled_gated(m, d_idx, n, d_out, m->d_flag, s);
}

}  // namespace fx
