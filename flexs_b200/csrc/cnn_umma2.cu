// K1 (tcgen05 variant, A = 4 shapes: k = 5, k3 = 3, F = 32) — warp-specialised, pipelined.
//
// Same mathematics as cnn_umma.cu (fp16 hi/lo split operands, FP32 accumulation in TMEM, taps as
// start-address offsets of the A descriptor) with three changes that the first ncu capture asked for
// (profiles/r01_umma_v1_*.txt: tensor pipe 9.6 % active, phases serialised, 73 cycles per MMA):
//
//  1. 128-byte aligned taps.  In the plain layout a tap moves the descriptor start by 16 B, so every
//     8-row core matrix straddles two 128 B lines and the A read costs double.  Here MMA row i of a
//     128-row tile is conv position 16*(i%8) + i/8: core matrix a holds positions {a, a+16, ...}, a
//     tap is +1 core matrix = +128 B, always aligned.  Core matrices a+j >= 16 live in a small "wrap"
//     region behind the tile (k-1 extra core matrices whose rows are shifted by one), written twice
//     by the producers (12-25 % of the rows).
//  2. [W_hi | W_lo] fused along N: one N=64 MMA gives hi*hi and hi*lo, one N=32 MMA adds lo*hi; the
//     activation planes are read twice per (tap, channel pair) instead of three times.  The epilogue
//     adds the two 32-column halves.
//  3. Roles: warps 0-7 epilogue (TMEM -> bias/ReLU/split -> A2, or -> max -> featT), warp 8 issues
//     MMAs, warps 9-15 produce conv1 for the NEXT chunk while the tensor core works on this one.
//     Hand-offs are mbarriers; the CTA only meets at __syncthreads around the dense head.
//
// conv1 for A = 4 is two table gathers: T012[a0,a1,a2] + T34[a3,a4] (bias and the activation scale
// folded in), 80 rows of 32 floats, instead of five.
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "dense_head.cuh"
#include "operand_prep.h"
#include "umma2_layout.cuh"

namespace {

using namespace u2;
constexpr int NT = 544;
// 17 warps: 0-7 conv2 epilogue (E2) of chunk n-1 then conv1 of chunk n, 8-15 conv3 epilogue (E3), 16 MMA issuer.
// Measured alternatives (profiles/r01_umma_v2_roles.txt): conv1 on its own 3 warps (14.1k cycles/chunk), conv1
// shared by all 16 warps (15.0k), conv1 with E3 (13.5k); this split is the fastest (12.3k).  Every generic
// shared-memory access competes with the tensor core's operand fetch, which is the real limiter.
constexpr int NEPI = 8, NPROD = 8, PRODW0 = 0, MMAW = 16;
constexpr int NTILE = 4;
constexpr int WR1 = K - 1, WR2 = K3 - 1;                      // wrap core matrices
constexpr int TS1 = (16 + WR1) * 128, TS2 = (16 + WR2) * 128;  // bytes per tile per plane
constexpr int PL1 = NTILE * TS1, PL2 = NTILE * TS2;            // bytes per plane
constexpr int TP = 36;  // padded row (floats) of the conv1 gather tables: rows land in different bank groups
constexpr int ROUT = (NTILE * 128 - (K3 - 1)) & ~3;  // conv3 output rows per chunk (508)

struct U2Params {
    const uint8_t *idx;
    float *out;
    const float *weights;
    const unsigned char *uw;  // per member: UW2 | UW3 | T012 | T34 | inv2s, inv3
    int *overflow_flag;
    int64_t n, n_items, member_floats, uw_member_bytes;
    fx::CnnDims d;
    fx::CnnOffsets o;
    int M;
    int P, hl, S;
    int sbcap, sbp, idx_slot, rs_bytes, stage;
    int dense_umma;  // dense head on the tensor cores (H <= 112)
    long long *prof;
    int dbg;  // profiling knobs (FLEXS_UMMA_DBG bitmask): 1 skip E3 reduction, 2 skip conv1 math, 4 skip E2 split/stores
};

// dense scratch inside the (idle) activation buffers
static_assert(DS_TOTAL <= 8 * (PL1 + PL2), "dense scratch must fit the activation buffers");

struct Offs {
    int mbar, tm, b, t012, t34, uw2, uw3, i0, i1, rs, feat, slot, a1, a2;
    size_t total;
};

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Byte offsets inside dynamic shared memory.  The kernel forms every pointer as smem_raw + offset in
// its own scope: passing pointers through a struct made the compiler fall back to generic LD/ST.
__host__ __device__ inline Offs carve(const U2Params &p) {
    Offs o;
    size_t off = 0;
    auto take = [&](size_t bytes, size_t align) {
        off = align_up(off, align);
        const size_t r = off;
        off += bytes;
        return (int)r;
    };
    o.mbar = take(16 * 8, 16); o.tm = take(16, 16); o.b = take(2 * F * 4, 16);
    o.t012 = take(64 * TP * 4, 16); o.t34 = take(16 * TP * 4, 16);
    o.uw2 = take((size_t)K * UWTAP, 128); o.uw3 = take((size_t)K3 * UWTAP, 128);
    o.i0 = take(p.idx_slot, 16); o.i1 = take(p.idx_slot, 16); o.rs = take(p.rs_bytes, 16);
    o.feat = take((size_t)F * p.sbp * 4, 16); o.slot = take((size_t)p.sbcap * 8, 16);
    o.a1 = take((size_t)8 * PL1, 1024); o.a2 = take((size_t)8 * PL2, 1024);
    o.total = off;
    return o;
}



__device__ __forceinline__ void issue_idx_load(const U2Params &p, uint8_t *dst, uint64_t *bar, int64_t item) {
    const int64_t first = item * p.S;
    const int64_t cnt = min((int64_t)p.S, p.n - first);
    const uintptr_t g0 = reinterpret_cast<uintptr_t>(p.idx) + (uintptr_t)(first * p.d.L);
    const uintptr_t a0 = g0 & ~(uintptr_t)15;
    const uintptr_t a1 = (g0 + (uintptr_t)(cnt * p.d.L) + 15) & ~(uintptr_t)15;
    const uint32_t bytes = (uint32_t)(a1 - a0);
    fxd::mbar_arrive_expect_tx(bar, bytes);
    fxd::bulk_g2s(dst, reinterpret_cast<const void *>(a0), bytes, bar);
}

struct Chunk {
    int64_t item, first;
    int s_item, rows_item, c0, nout, ntile2, ntile3, slot0;
    uint32_t n, iter;
    bool first_chunk;
};

// every role walks the same (item, chunk) sequence of a group
template <class Fn>
__device__ __forceinline__ void walk_group(const U2Params &p, int64_t item_begin, int64_t item_end, uint32_t n0,
                                           uint32_t iter0, Fn &&fn) {
    Chunk c;
    c.n = n0; c.iter = iter0; c.slot0 = 0;
    for (c.item = item_begin; c.item < item_end; c.item += gridDim.x, ++c.iter) {
        c.first = c.item * p.S;
        c.s_item = (int)min((int64_t)p.S, p.n - c.first);
        c.rows_item = c.s_item * p.P;
        for (c.c0 = 0; c.c0 < c.rows_item; c.c0 += ROUT, ++c.n) {
            c.nout = min(ROUT, c.rows_item - c.c0);
            c.ntile3 = (c.nout + 127) >> 7;
            c.ntile2 = min(NTILE, (c.nout + K3 - 1 + 127) >> 7);
            c.first_chunk = (c.c0 == 0);
            fn(c);
        }
        c.slot0 += c.s_item;
    }
}


__global__ void __launch_bounds__(NT, 1) cnn_umma2_kernel(const U2Params p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const Offs of = carve(p);
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem_raw + of.mbar);
    uint64_t *mbar_idx = mbar, *a1_full = mbar + 2, *a2_full = mbar + 3, *e3_done = mbar + 4, *c2 = mbar + 5, *c3 = mbar + 9, *dbar = mbar + 13;
    uint32_t *tmem_addr_s = reinterpret_cast<uint32_t *>(smem_raw + of.tm);
    float *b2s = reinterpret_cast<float *>(smem_raw + of.b), *b3 = b2s + F;
    float *t012 = reinterpret_cast<float *>(smem_raw + of.t012), *t34 = reinterpret_cast<float *>(smem_raw + of.t34);
    unsigned char *uw2 = smem_raw + of.uw2, *uw3 = smem_raw + of.uw3;
    uint8_t *rowseq = smem_raw + of.rs;
    float *featT = reinterpret_cast<float *>(smem_raw + of.feat);
    long long *slot_seq = reinterpret_cast<long long *>(smem_raw + of.slot);
    unsigned char *a1 = smem_raw + of.a1, *a2 = smem_raw + of.a2;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int P = p.P, hl = p.hl, L = p.d.L, T = p.d.T;
    const int pl2 = p.d.pl2, pl3 = p.d.pl3;

    if (tid == 0) {
        fxd::mbar_init(&mbar_idx[0], 1); fxd::mbar_init(&mbar_idx[1], 1);
        fxd::mbar_init(dbar, 1);
        fxd::mbar_init(a1_full, NPROD); fxd::mbar_init(a2_full, NEPI); fxd::mbar_init(e3_done, 8);
        for (int i = 0; i < NTILE; ++i) { fxd::mbar_init(&c2[i], 1); fxd::mbar_init(&c3[i], 1); }
        fxd::fence_mbar_init();
    }
    if (wid == 0) tmem_alloc(tmem_addr_s, 512);
    for (int i = tid; i < 8 * PL1 / 16; i += NT) reinterpret_cast<uint4 *>(a1)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < 8 * PL2 / 16; i += NT) reinterpret_cast<uint4 *>(a2)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < F * p.sbp; i += NT) featT[i] = 0.f;
    // rowseq[rho]: sequence-in-item of item row rho if it is a real conv position, else 0xFF
    for (int rho = tid; rho < p.S * P; rho += NT) {
        const int s = rho / P, t = rho - s * P - hl;
        rowseq[rho] = (t >= 0 && t < T) ? (uint8_t)s : (uint8_t)0xFF;
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_addr_s;
    const uint32_t a1_addr = fxd::smem_u32(a1), a2_addr = fxd::smem_u32(a2);
    const uint32_t uw2_addr = fxd::smem_u32(uw2), uw3_addr = fxd::smem_u32(uw3);

    uint32_t n = 0, iter = 0;  // chunks done, items done (running over members and groups)
    uint32_t dph = 0;          // completed phases of the dense-head barrier
    float xmax = 0.f;          // largest activation written as fp16 (range guard)
    long long pt[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};

    for (int mem = 0; mem < p.M; ++mem) {
        const float *w = p.weights + (int64_t)mem * p.member_floats;
        const unsigned char *uw = p.uw + (int64_t)mem * p.uw_member_bytes;
        const unsigned char *tbig = uw + OFF_TBIG;  // conv1 as a single gather (global, L2-resident, 128 KB)
        const float inv2s = __ldg(reinterpret_cast<const float *>(uw + OFF_SCAL));
        const float inv3 = __ldg(reinterpret_cast<const float *>(uw + OFF_SCAL) + 1);
        __syncthreads();
        for (int i = tid; i < F; i += NT) {
            b2s[i] = __ldg(w + p.o.b2 + i) * ASCALE;
            b3[i] = __ldg(w + p.o.b3 + i);
        }
        for (int i = tid; i < K * UWTAP / 16; i += NT)
            reinterpret_cast<uint4 *>(uw2)[i] = __ldg(reinterpret_cast<const uint4 *>(uw) + i);
        for (int i = tid; i < K3 * UWTAP / 16; i += NT)
            reinterpret_cast<uint4 *>(uw3)[i] = __ldg(reinterpret_cast<const uint4 *>(uw + OFF_UW3) + i);
        for (int i = tid; i < 64 * F / 4; i += NT)
            reinterpret_cast<float4 *>(t012 + (i >> 3) * TP)[i & 7] = __ldg(reinterpret_cast<const float4 *>(uw + OFF_T012) + i);
        for (int i = tid; i < 16 * F / 4; i += NT)
            reinterpret_cast<float4 *>(t34 + (i >> 3) * TP)[i & 7] = __ldg(reinterpret_cast<const float4 *>(uw + OFF_T34) + i);
        fence_async_smem();
        __syncthreads();

        int64_t item = blockIdx.x;
        if (tid == PRODW0 * 32 && item < p.n_items) issue_idx_load(p, smem_raw + ((iter & 1) ? of.i1 : of.i0), &mbar_idx[iter & 1], item);
        while (item < p.n_items) {
            // ---- the group of items whose features fit the dense-head batch ----
            int64_t g_end = item;
            int g_slots = 0;
            uint32_t g_chunks = 0, g_items = 0;
            while (g_end < p.n_items) {
                const int s = (int)min((int64_t)p.S, p.n - g_end * p.S);
                if (g_slots + s > p.sbcap) break;
                g_slots += s;
                g_chunks += (uint32_t)((s * P + ROUT - 1) / ROUT);
                ++g_items;
                g_end += gridDim.x;
            }
            const long long tg0 = clock64();

            // conv1 -> A1 for one chunk; shared by all 16 worker warps (warp `wid` takes units wid, wid+16, ...)
            const int pw = wid - PRODW0, ptid = tid - PRODW0 * 32;
            auto conv1 = [&](const Chunk &c) {
                    const int buf = c.iter & 1;
                    const long long q0 = clock64();
                    if (c.n > 0) fxd::mbar_wait(&c2[NTILE - 1], (c.n - 1) & 1);  // A1 free again
                    const long long q1 = clock64();
                    if (c.first_chunk) {
                        for (int i = ptid; i < c.s_item; i += NPROD * 32) slot_seq[c.slot0 + i] = c.first + i;
                        fxd::mbar_wait(&mbar_idx[buf], (c.iter >> 1) & 1);
                    }
                    const long long q2 = clock64();
                    const uint8_t *sidx = smem_raw + (buf ? of.i1 : of.i0) +
                                          ((reinterpret_cast<uintptr_t>(p.idx) + (uintptr_t)(c.first * L)) & 15);
                    // One unit = 32 MMA rows of a tile (lane -> row i = 8a + b -> conv position 16b + a, so a warp's
                    // 16-byte stores are contiguous), plus one last unit for the k-1 rows behind the last tile.
                    // A row of h1 is ONE 128-byte entry of a 4^5-entry table (bias, ReLU, activation scale and the fp16
                    // hi/lo split folded in at prepare time) gathered from L2: no shared-memory table reads compete with
                    // the tensor core's operand stream.  A warp owns up to three units and issues all their gathers
                    // before the first store, so the L2 latency is paid once per chunk, not once per unit.
                    // The gathers are cp.async (LDGSTS) 16-byte copies straight into the operand planes: no registers,
                    // no shared-memory table reads, and every unit of the warp is in flight before the first one lands.
                    const int nunits = c.ntile2 * 4 + 1;
                    const uint32_t a1s = a1_addr;
                    for (int u = pw; u < nunits; u += NPROD) {
                        int tile, a, b;
                        if (u < c.ntile2 * 4) {
                            const int i = (u & 3) * 32 + lane;
                            tile = u >> 2; a = i >> 3; b = i & 7;
                        } else {
                            if (lane >= K - 1) continue;
                            tile = c.ntile2; a = lane; b = 0;
                        }
                        const int rho = c.c0 - pl3 - pl2 + tile * 128 + 16 * b + a;
                        const int s = (rho >= 0 && rho < c.rows_item) ? rowseq[rho] : 0xFF;
                        const int moff = (tile < NTILE) ? tile * TS1 + a * 128 + b * 16 : -1;
                        int woff = -1;
                        if (a < WR1) {
                            const int tt = (b == 0) ? tile - 1 : tile, bb = (b == 0) ? 7 : b - 1;
                            if (tt >= 0 && tt < NTILE) woff = tt * TS1 + (16 + a) * 128 + bb * 16;
                        }
                        const unsigned char *tp = tbig;
                        uint32_t nbytes = 0;                       // 0 -> cp.async zero-fills (halo / padding rows)
                        if (s != 0xFF && !(p.dbg & 2)) {
                            const uint8_t *ip = sidx + s * L + (rho - s * P - hl);
                            const int i5 = ((((ip[0] * ALPHA + ip[1]) * ALPHA + ip[2]) * ALPHA + ip[3]) * ALPHA) + ip[4];
                            tp = tbig + (size_t)i5 * 128;
                            nbytes = 16;
                        }
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const uint32_t plane = a1s + (uint32_t)q * PL1;   // q = 0..3 hi chunks, 4..7 lo chunks
                            if (moff >= 0)
                                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(plane + moff), "l"(tp + q * 16),
                                             "r"(nbytes) : "memory");
                            if (woff >= 0)
                                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(plane + woff), "l"(tp + q * 16),
                                             "r"(nbytes) : "memory");
                        }
                    }
                    asm volatile("cp.async.wait_all;" ::: "memory");
                    const long long q3 = clock64();
                    fence_async_smem();
                    __syncwarp();
                    const long long q4 = clock64();
                    if (lane == 0) mbar_arrive(a1_full);
                    if (ptid == 0) { pt[8] += q1 - q0; pt[9] += clock64() - q1; pt[14] += q2 - q1; pt[15] += q3 - q2; pt[1] += q4 - q3; }
                };
            if (wid < 8) {
                // ====== warps 0-7: conv2 epilogue (E2) + conv1 producer ======
                const int lq = wid & 3, ch = wid >> 2;
                // TMEM lane i = 32*lq + lane is MMA row i = 8a + b, i.e. conv position 16b + a of the tile
                const int ea = 4 * lq + (lane >> 3), eb = lane & 7;
                const int pos = 16 * eb + ea;
                const int e_main = ea * 128 + eb * 16;                       // inside a tile of a plane
                const bool e_wrap = ea < WR2;                                // this row also feeds a wrap slot
                const int e_wrap_off = (16 + ea) * 128 + ((eb == 0) ? 7 : eb - 1) * 16;
                const uint32_t tlane = tmem_base + ((uint32_t)(lq * 32) << 16);
                unsigned char *a2h = a2 + (size_t)(ch * 2) * PL2;            // hi plane of this warp's first chunk
                auto e2 = [&](const Chunk &c) {
                    const uint32_t par = c.n & 1;
                    // ---- E2: conv2 accumulators -> bias, ReLU, mask, split -> A2 ----
                    for (int t = 0; t < c.ntile2; ++t) {
                        const long long w0 = clock64();
                        fxd::mbar_wait(&c2[t], par);
                        if (t == 0 && c.n > 0) fxd::mbar_wait(&c3[NTILE - 1], (c.n - 1) & 1);  // A2 free again
                        if (tid == 0) pt[2] += clock64() - w0;
                        tc_fence_after();
                        uint32_t v[16], v2[16];
                        tmem_ld16_nowait(tlane + (uint32_t)(t * 64 + ch * 16), v);
                        tmem_ld16_nowait(tlane + (uint32_t)(t * 64 + 32 + ch * 16), v2);
                        const int rho = c.c0 - pl3 + t * 128 + pos;
                        const bool valid = (rho >= 0) && (rho < c.rows_item) && (rowseq[rho] != 0xFF);
                        const int wrap_tile = (eb == 0) ? t - 1 : t;
                        tmem_ld_wait();
                        if (p.dbg & 4) continue;
#pragma unroll
                        for (int cc = 0; cc < 2; ++cc) {
                            float x[8];
                            const float4 bA = *reinterpret_cast<const float4 *>(b2s + ch * 16 + cc * 8);
                            const float4 bB = *reinterpret_cast<const float4 *>(b2s + ch * 16 + cc * 8 + 4);
                            const float bb[8] = {bA.x, bA.y, bA.z, bA.w, bB.x, bB.y, bB.z, bB.w};
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const float acc = __uint_as_float(v[cc * 8 + q]) + __uint_as_float(v2[cc * 8 + q]);
                                x[q] = fmaxf(fmaf(acc, inv2s, bb[q]), 0.f);
                            }
                            uint4 hi4, lo4;
                            float lmax = 0.f;
                            split8(x, hi4, lo4, lmax);
                            // masked rows may hold anything (stale tails of the buffers): keep them out of
                            // both the activations and the range guard
                            if (valid) xmax = fmaxf(xmax, lmax);
                            else { hi4 = make_uint4(0, 0, 0, 0); lo4 = hi4; }
                            unsigned char *ph = a2h + (size_t)cc * PL2 + t * TS2, *plo = ph + (size_t)4 * PL2;
                            *reinterpret_cast<uint4 *>(ph + e_main) = hi4;
                            *reinterpret_cast<uint4 *>(plo + e_main) = lo4;
                            if (e_wrap && wrap_tile >= 0) {
                                const int wo = e_wrap_off + (wrap_tile - t) * TS2;
                                *reinterpret_cast<uint4 *>(ph + wo) = hi4;
                                *reinterpret_cast<uint4 *>(plo + wo) = lo4;
                            }
                        }
                    }
                    fence_async_smem();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(a2_full);
                };
                Chunk prev;
                bool have_prev = false;
                walk_group(p, item, g_end, n, iter, [&](const Chunk &c) {
                    if (have_prev) e2(prev);   // E2(n-1) ...
                    conv1(c);                  // ... then h1 of chunk n while the tensor core runs conv3(n-1)
                    prev = c; have_prev = true;
                });
                if (have_prev) e2(prev);
            } else if (wid < 16) {
                // ====== warps 8-15: conv3 epilogue (E3) ======
                const int lq = wid & 3, ch = (wid >> 2) & 1;
                const int ea = 4 * lq + (lane >> 3), eb = lane & 7;
                const int pos = 16 * eb + ea;  // conv position of TMEM lane 32*lq + lane inside a tile
                const uint32_t tlane = tmem_base + ((uint32_t)(lq * 32) << 16);
                auto e3 = [&](const Chunk &c) {
                    const uint32_t par = c.n & 1;
                    // ---- E3: conv3 accumulators -> bias, ReLU, mask -> max per sequence -> featT ----
                    for (int t = 0; t < c.ntile3; ++t) {
                        const long long w0 = clock64();
                        fxd::mbar_wait(&c3[t], par);
                        if (tid == 8 * 32) pt[4] += clock64() - w0;
                        tc_fence_after();
                        uint32_t v[16], v2[16];
                        tmem_ld16_nowait(tlane + 256u + (uint32_t)(t * 64 + ch * 16), v);
                        tmem_ld16_nowait(tlane + 256u + (uint32_t)(t * 64 + 32 + ch * 16), v2);
                        const int r = t * 128 + pos;  // output row of this chunk
                        const int s = (r < c.nout) ? rowseq[c.c0 + r] : 0xFF;
                        const bool valid = (s != 0xFF);
                        tmem_ld_wait();
                        uint32_t bits[16];
#pragma unroll
                        for (int g4 = 0; g4 < 4; ++g4) {
                            const float4 bq = *reinterpret_cast<const float4 *>(b3 + ch * 16 + g4 * 4);
                            const float bb[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const float acc = __uint_as_float(v[g4 * 4 + q]) + __uint_as_float(v2[g4 * 4 + q]);
                                const float y = fmaxf(fmaf(acc, inv3, bb[q]), 0.f);
                                bits[g4 * 4 + q] = valid ? __float_as_uint(y) : 0u;  // y >= 0: uint order == float order
                            }
                        }
                        // GlobalMaxPooling1D over the rows of each sequence present in this warp: one REDUX per
                        // filter, lane q keeps filter q, then a single 16-address shared atomic per sequence
                        unsigned todo = (p.dbg & 1) ? 0u : __ballot_sync(0xffffffffu, valid);
                        if (P < 48) {
                            // short sequences: a tile holds many of them, each owning only a lane or two of this warp,
                            // so per-lane shared atomics (<= 2-way conflicts) beat one REDUX round per sequence
                            if (valid && todo) {
                                unsigned int *dst = reinterpret_cast<unsigned int *>(featT) + (size_t)(ch * 16) * p.sbp + c.slot0 + s;
#pragma unroll
                                for (int q = 0; q < 16; ++q) atomicMax(dst + (size_t)q * p.sbp, bits[q]);
                            }
                            todo = 0;
                        }
                        while (todo) {
                            const int leader = __ffs(todo) - 1;
                            const int s_l = __shfl_sync(0xffffffffu, s, leader);
                            const unsigned seg = __ballot_sync(0xffffffffu, valid && (s == s_l));
                            uint32_t keep = 0;
#pragma unroll
                            for (int q = 0; q < 16; ++q) {
                                const uint32_t red = __reduce_max_sync(0xffffffffu, (seg >> lane) & 1u ? bits[q] : 0u);
                                if (lane == q) keep = red;
                            }
                            if (lane < 16)
                                atomicMax(reinterpret_cast<unsigned int *>(featT) + (size_t)(ch * 16 + lane) * p.sbp +
                                              c.slot0 + s_l, keep);
                            todo &= ~seg;
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(e3_done);
                };
                walk_group(p, item, g_end, n, iter, [&](const Chunk &c) { e3(c); });
            } else {
                // =========================== MMA issuer ===========================
                walk_group(p, item, g_end, n, iter, [&](const Chunk &c) {
                    const uint32_t par = c.n & 1;
                    if (c.first_chunk) {
                        // Prefetch the next item's residues into the other slot: its last readers (conv1 of the previous
                        // item) arrived on a1_full before this warp got here.  Issued from this warp because it is about
                        // to wait anyway; on a producer warp the bulk-copy issue cost ~0.8k cycles of the critical path.
                        const int64_t next = c.item + gridDim.x;
                        const int buf = c.iter & 1;
                        if (lane == 0 && next < p.n_items)
                            issue_idx_load(p, smem_raw + (buf ? of.i0 : of.i1), &mbar_idx[buf ^ 1], next);
                        __syncwarp();
                    }
                    const long long m0 = clock64();
                    fxd::mbar_wait(a1_full, par);
                    const long long m1 = clock64();
                    tc_fence_after();
#pragma unroll
                    for (int t = 0; t < NTILE; ++t) {
                        if (t < c.ntile2)
                            issue_conv_tile<K, PL1>(a1_addr + (uint32_t)t * TS1, uw2_addr, tmem_base + (uint32_t)t * 64u);
                        umma_commit_elect(&c2[t]);
                    }
                    const long long m2 = clock64();
                    fxd::mbar_wait(a2_full, par);
                    const long long m3 = clock64();
                    if (c.n > 0) fxd::mbar_wait(e3_done, (c.n - 1) & 1);  // conv3 accumulators drained
                    const long long m4 = clock64();
                    if (lane == 0) { pt[10] += m1 - m0; pt[11] += m3 - m2; pt[12] += m4 - m3; pt[13] += m2 - m1; }
                    tc_fence_after();
#pragma unroll
                    for (int t = 0; t < NTILE; ++t) {
                        if (t < c.ntile3)
                            issue_conv_tile<K3, PL2>(a2_addr + (uint32_t)t * TS2, uw3_addr, tmem_base + 256u + (uint32_t)t * 64u);
                        umma_commit_elect(&c3[t]);
                    }
                });
            }
            n += g_chunks; iter += g_items; item = g_end;
            // ---- drain, dense head on the group's features, reset featT ----
            tc_fence_before();
            __syncthreads();
            const long long tg1 = clock64();
            if (p.dense_umma) {
                // tensor-core dense head on the idle activation buffers (u2::dense_head_umma, shared with cnn_k9.cu)
                dense_head_umma<NT, MMAW, true>(a1, featT, p.sbp, p.sbcap, g_slots, uw, tmem_base, dbar, dph, xmax, mem, p.M,
                                                p.out, [slot_seq](int sl) { return slot_seq[sl]; });
            } else {
                fxd::DenseArgs da{w + p.o.wd1, w + p.o.bd1, w + p.o.wd2, w + p.o.bd2, w + p.o.wd3, w + p.o.bd3,
                                  featT, reinterpret_cast<float *>(a1), slot_seq, p.out,
                                  F, p.d.H, p.sbp, g_slots, mem, p.M, p.stage};
                fxd::dense_head_flush<NT>(da);
            }
            for (int i = tid; i < F * p.sbp; i += NT) featT[i] = 0.f;
            __syncthreads();
            if (tid == 0) { pt[0] += tg1 - tg0; pt[5] += clock64() - tg1; pt[6] += g_chunks; }
        }
    }
    if (xmax > 60000.f) atomicExch(p.overflow_flag, 1);
    if (p.prof != nullptr && (tid == 0 || tid == 8 * 32 || tid == PRODW0 * 32 || tid == MMAW * 32))  // warp 0 times the conv2 waits, warp 8 the conv3 waits
        for (int i = 0; i < 16; ++i)
            if (pt[i]) atomicAdd(reinterpret_cast<unsigned long long *>(&p.prof[(size_t)blockIdx.x * 16 + i]), (unsigned long long)pt[i]);
    tc_fence_before();
    __syncthreads();
    if (wid == 0) tmem_dealloc(tmem_base, 512);
}

static bool plan(const flexs_model *m, U2Params &p) {
    p.d = fx::cnn_dims(m);
    p.o = fx::cnn_offsets(m);
    p.M = m->M;
    p.member_floats = m->member_floats;
    p.uw_member_bytes = UW_MEMBER_BYTES;
    p.hl = std::max(p.d.pl2, p.d.pl3);
    const int hr = std::max(p.d.pr2, p.d.pr3);
    p.P = (p.hl + p.d.T + hr + 3) & ~3;
    const size_t abytes = (size_t)8 * (PL1 + PL2);
    p.dense_umma = (p.d.H <= DH) ? 1 : 0;
    int sbcap = p.dense_umma ? DSLOTS : 64;
    while (!p.dense_umma && sbcap >= 8 && fxd::dense_scratch_floats(F, p.d.H, sbcap + 4, false) * 4 > abytes) sbcap -= 8;
    if (sbcap < 8) return false;
    p.sbcap = sbcap; p.sbp = sbcap + 4;
    p.stage = (!p.dense_umma && fxd::dense_scratch_floats(F, p.d.H, p.sbp, true) * 4 <= abytes) ? 1 : 0;
    p.S = std::max(1, std::min(ROUT / p.P, sbcap));
    p.idx_slot = (int)align_up((size_t)p.S * p.d.L + 32, 16);
    p.rs_bytes = (int)align_up((size_t)p.S * p.P, 16);
    return (int64_t)carve(p).total + 1024 <= m->max_smem_optin;
}

static int prepare(flexs_model *m, const U2Params &p) {
    if (m->umma2_ready) return FLEXS_OK;
    std::vector<float> host((size_t)m->member_floats * m->M);
    FX_CUDA(cudaMemcpy(host.data(), m->d_weights, host.size() * sizeof(float), cudaMemcpyDeviceToHost));
    std::vector<unsigned char> blob((size_t)UW_MEMBER_BYTES * m->M, 0);
    m->umma_weights_ok = true;
    for (int mem = 0; mem < m->M; ++mem) {
        const float *w = host.data() + (size_t)mem * m->member_floats;
        unsigned char *dst = blob.data() + (size_t)mem * UW_MEMBER_BYTES;
        float inv[2];
        inv[0] = prep::fill_conv_planes(w + p.o.w2, K, dst, &m->umma_weights_ok);
        inv[1] = prep::fill_conv_planes(w + p.o.w3, K3, dst + OFF_UW3, &m->umma_weights_ok);
        // conv1 gather tables for A = 4: T012[a0,a1,a2] (with bias) and T34[a3,a4], scaled by ASCALE
        const float *w1 = w + p.o.w1, *b1 = w + p.o.b1;  // w1 (k, A, F)
        float *t012 = reinterpret_cast<float *>(dst + OFF_T012), *t34 = reinterpret_cast<float *>(dst + OFF_T34);
        for (int a0 = 0; a0 < ALPHA; ++a0)
            for (int a1 = 0; a1 < ALPHA; ++a1)
                for (int a2 = 0; a2 < ALPHA; ++a2)
                    for (int f = 0; f < F; ++f) {
                        const float v = ((w1[(0 * ALPHA + a0) * F + f] + w1[(1 * ALPHA + a1) * F + f]) +
                                         w1[(2 * ALPHA + a2) * F + f]) + b1[f];
                        t012[((a0 * ALPHA + a1) * ALPHA + a2) * F + f] = v * ASCALE;
                    }
        for (int a3 = 0; a3 < ALPHA; ++a3)
            for (int a4 = 0; a4 < ALPHA; ++a4)
                for (int f = 0; f < F; ++f)
                    t34[(a3 * ALPHA + a4) * F + f] = (w1[(3 * ALPHA + a3) * F + f] + w1[(4 * ALPHA + a4) * F + f]) * ASCALE;
        {   // conv1 as one gather: entry[a0..a4] = split(relu(T012 + T34)) for all 32 channels
            __half *tb = reinterpret_cast<__half *>(dst + OFF_TBIG);
            for (int i5 = 0; i5 < 1024; ++i5) {
                const int i012 = i5 >> 4, i34 = i5 & 15;
                for (int f = 0; f < F; ++f) {
                    const float x = std::max(t012[i012 * F + f] + t34[i34 * F + f], 0.f);
                    if (!(x <= 60000.f)) m->umma_weights_ok = false;  // beyond the fp16 window (or NaN): fp32 kernel
                    const __half hi = __float2half_rn(x);
                    const __half lo = __float2half_rn(x - __half2float(hi));
                    tb[(size_t)i5 * 64 + (f >> 3) * 8 + (f & 7)] = hi;
                    tb[(size_t)i5 * 64 + 32 + (f >> 3) * 8 + (f & 7)] = lo;
                }
            }
        }
        float *tail = reinterpret_cast<float *>(dst + OFF_SCAL);
        tail[0] = inv[0] * ASCALE;  // conv2 epilogue emits activations pre-scaled by ASCALE
        tail[1] = inv[1];
        // dense head operands (used when H <= 112): Wd1 (F,H) and Wd2 (H,H) as [k chunk][n hi|lo][8 k] planes
        const int H = p.d.H;
        if (H <= DH) prep::fill_dense_head(w, p.o, H, dst, &m->umma_weights_ok);
    }
    FX_CUDA(cudaSetDevice(m->device));
    if (!m->d_umma2_w) FX_CUDA(cudaMalloc(&m->d_umma2_w, blob.size()));
    FX_CUDA(cudaMemcpy(m->d_umma2_w, blob.data(), blob.size(), cudaMemcpyHostToDevice));
    m->umma2_ready = true;
    return FLEXS_OK;
}

}  // namespace

namespace fx {

int prepare_cnn_umma2(flexs_model *m) {
    U2Params p;
    FX_REQUIRE(plan(m, p), "shape not supported by the pipelined UMMA kernel");
    return prepare(m, p);
}

bool cnn_umma2_supported(const flexs_model *m) {
    if (m->kind != FLEXS_KIND_CNN || m->F != 32 || m->K != 5 || m->A != 4) return false;
    if (!cnn_tiled_supported(m)) return false;  // the fp16-overflow fall-back path
    U2Params p;
    return plan(m, p);
}

int launch_cnn_umma2(flexs_model *m, const uint8_t *d_idx, int64_t n, float *d_out, cudaStream_t s) {
    U2Params p;
    FX_REQUIRE(cnn_umma2_supported(m) && plan(m, p), "shape not supported by the pipelined UMMA kernel");
    int rc = prepare(m, p);
    if (rc != FLEXS_OK) return rc;
    if (!m->umma_weights_ok) return launch_cnn_tiled(m, d_idx, n, d_out, s);  // non-finite weights: fp32 path
    p.idx = d_idx; p.out = d_out; p.weights = m->d_weights; p.n = n;
    p.uw = reinterpret_cast<const unsigned char *>(m->d_umma2_w);
    flexs_model::StreamWs *ws = nullptr;  // the fp16-overflow flag is per stream: chunks of score_host run concurrently
    rc = stream_workspace(m, s, 0, &ws);
    if (rc != FLEXS_OK) return rc;
    p.overflow_flag = ws->flag;
    p.n_items = (n + p.S - 1) / p.S;
    const size_t smem = carve(p).total + 1024;
    const int grid = (int)std::min<int64_t>(p.n_items, m->sm_count);
    static const bool prof = std::getenv("FLEXS_UMMA_PROF") && std::getenv("FLEXS_UMMA_PROF")[0] == '1';
    p.dbg = std::getenv("FLEXS_UMMA_DBG") ? std::atoi(std::getenv("FLEXS_UMMA_DBG")) : 0;
    p.prof = nullptr;
    if (prof) {
        FX_CUDA(cudaMalloc(&p.prof, (size_t)grid * 16 * sizeof(long long)));
        FX_CUDA(cudaMemset(p.prof, 0, (size_t)grid * 16 * sizeof(long long)));
    }
    FX_CUDA(cudaMemsetAsync(ws->flag, 0, sizeof(int), s));
    FX_CUDA(cudaFuncSetAttribute(cnn_umma2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cnn_umma2_kernel<<<grid, NT, smem, s>>>(p);
    FX_CUDA(cudaGetLastError());
    m->launches += 1;
    if (prof) {
        FX_CUDA(cudaStreamSynchronize(s));
        std::vector<long long> h((size_t)grid * 16);
        FX_CUDA(cudaMemcpy(h.data(), p.prof, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        cudaFree(p.prof);
        double a[16] = {0};
        for (int b = 0; b < grid; ++b) for (int i = 0; i < 16; ++i) a[i] += (double)h[(size_t)b * 16 + i] / grid;
        const double ch = a[6] > 0 ? a[6] : 1;
        fprintf(stderr, "[umma2 prof] n=%lld grid=%d chunks/CTA=%.0f | cycles per chunk: pipeline %.0f (epilogue warp 0 "
                        "waiting on conv2 MMAs %.0f, conv3 MMAs %.0f), dense+drain %.0f | producer: wait A1 free %.0f, conv1 %.0f | "
                        "MMA warp: wait A1 %.0f, issue conv2 %.0f, wait A2 %.0f, wait E3 %.0f | conv1 parts: idx wait %.0f, rows %.0f, fence %.0f\n",
                (long long)n, grid, a[6], a[0] / ch, a[2] / ch, a[4] / ch, a[5] / ch, a[8] / ch, a[9] / ch, a[10] / ch, a[13] / ch,
                a[11] / ch, a[12] / ch, a[14] / ch, a[15] / ch, a[1] / ch);
    }
    // fp16 range guard: the gated FFMA kernel recomputes the batch iff the flag was raised
    return launch_cnn_tiled_gated(m, d_idx, n, d_out, ws->flag, s);
}

}  // namespace fx
