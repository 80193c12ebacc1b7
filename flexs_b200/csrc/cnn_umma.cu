// K1 (tcgen05 variant): fused CNN forward with conv2/conv3 on the 5th-gen tensor cores.
//
// Same row-space formulation as cnn_tiled.cu (each sequence owns P rows, zero halo rows give the
// "same" padding, a tap is a +1 row shift), but the two implicit GEMMs
//     conv2: [rows, 5*32] x [5*32, 32]      conv3: [rows, k3*32] x [k3*32, 32]
// run as tcgen05.mma (M=128 rows, N=32 filters, K=16 channels per instruction) with FP32
// accumulators in TMEM.  The A operand is the activation buffer itself, stored in shared memory in
// the UMMA K-major no-swizzle canonical layout with a 16-byte row pitch per 8-channel chunk
// ("plane"): row r of chunk c sits at plane[c] + 16*r, so the im2col of tap j is just a +16*j byte
// start-address offset in the matrix descriptor — nothing is copied.
//
// Precision: fp16 inputs would lose the 1e-4 contract, so every operand is split x = hi + lo
// (two fp16 values, 22 significand bits together) and each product runs as three MMAs
// (hi*hi + hi*lo + lo*hi) into the same FP32 accumulator: fp32-level accuracy at 1/3 of the fp16
// tensor rate.  Operands are pre-scaled by powers of two (weights per layer to ~2^14, activations
// by 2^3) so the lo parts stay in fp16's normal range; the epilogue undoes the scale exactly.
// Activations above 60000/8 would overflow fp16: the epilogue raises a flag and the launcher's
// gated FFMA kernel (cnn_tiled.cu) recomputes the batch in that case.
//
// Per chunk of up to ntile*128 rows:
//   conv1 gather-add (CUDA cores, all warps)            -> A1 planes (fp16 hi/lo)
//   warp 8 issues conv2 MMAs (one elected lane), tcgen05.commit per tile -> mbarrier
//   warps 0-7: tcgen05.ld, bias+ReLU+mask, split        -> A2 planes
//   warp 8 issues conv3 MMAs
//   warps 0-7: tcgen05.ld, bias+ReLU+mask, warp REDUX max per sequence -> featT (smem atomics)
//   every <=64 sequences: dense head (dense_head.cuh)    -> out
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "dense_head.cuh"

namespace {

constexpr int F = 32;
constexpr int NT = 512;
constexpr int W1P = 36;
constexpr float ASCALE = 8.f;
constexpr int WPLANE = 1024;   // bytes of one 8-channel weight plane: [hi 32 filters | lo 32 filters] x 16 B
constexpr int WTAP = 2 * WPLANE;  // half a tap (kept so that a tap = 2 * WTAP = 4 chunks x 1 KB, as sized below)

struct UmmaParams {
    const uint8_t *idx;
    float *out;
    const float *weights;
    const unsigned char *uw;  // per member: W2 planes | W3 planes | inv2s, inv3 (floats)
    int *overflow_flag;
    int64_t n, n_items, member_floats, uw_member_bytes;
    fx::CnnDims d;
    fx::CnnOffsets o;
    int M;
    int P, hl, S;
    int ntile, rout, rows1, rows2;
    int sbcap, sbp, idx_slot;
    int stage;  // dense head stages Wd1/Wd2 in shared memory
    int swap_lbo_sbo;  // debug knob (FLEXS_UMMA_SWAP=1)
    long long *prof;   // debug: per-CTA phase cycle counters (FLEXS_UMMA_PROF=1)
};

struct Smem {
    uint64_t *mbar_idx;  // [2]
    uint64_t *mbar_c2;   // [4]
    uint64_t *mbar_c3;   // [4]
    uint32_t *tmem_addr;
    float *b1, *b2s, *b3;
    float *w1;
    unsigned char *uw2, *uw3;
    uint8_t *idx[2];
    float *featT;
    long long *slot_seq;
    unsigned char *a1, *a2;
};

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

__host__ __device__ inline size_t carve(unsigned char *base, const UmmaParams &p, Smem *s) {
    size_t off = 0;
    auto take = [&](size_t bytes, size_t align) {
        off = align_up(off, align);
        size_t o = off;
        off += bytes;
        return o;
    };
    size_t o_mbar = take(16 * 8, 16);
    size_t o_tm = take(16, 16);
    size_t o_b = take(3 * F * 4, 16);
    size_t o_w1 = take((size_t)p.d.K * p.d.A * W1P * 4, 16);
    size_t o_uw2 = take((size_t)p.d.K * 2 * WTAP, 128);
    size_t o_uw3 = take((size_t)p.d.K3 * 2 * WTAP, 128);
    size_t o_i0 = take(p.idx_slot, 16);
    size_t o_i1 = take(p.idx_slot, 16);
    size_t o_feat = take((size_t)F * p.sbp * 4, 16);
    size_t o_slot = take((size_t)p.sbcap * 8, 16);
    size_t o_a1 = take((size_t)8 * p.rows1 * 16, 128);
    size_t o_a2 = take((size_t)8 * p.rows2 * 16, 128);
    if (s) {
        s->mbar_idx = reinterpret_cast<uint64_t *>(base + o_mbar);
        s->mbar_c2 = s->mbar_idx + 2;
        s->mbar_c3 = s->mbar_idx + 6;
        s->tmem_addr = reinterpret_cast<uint32_t *>(base + o_tm);
        s->b1 = reinterpret_cast<float *>(base + o_b);
        s->b2s = s->b1 + F;
        s->b3 = s->b1 + 2 * F;
        s->w1 = reinterpret_cast<float *>(base + o_w1);
        s->uw2 = base + o_uw2;
        s->uw3 = base + o_uw3;
        s->idx[0] = base + o_i0;
        s->idx[1] = base + o_i1;
        s->featT = reinterpret_cast<float *>(base + o_feat);
        s->slot_seq = reinterpret_cast<long long *>(base + o_slot);
        s->a1 = base + o_a1;
        s->a2 = base + o_a2;
    }
    return off;
}

// ---- tcgen05 wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(fxd::smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
    // reached by the whole (converged) MMA warp, issued by one elected lane: keeps the descriptor arithmetic on
    // the uniform datapath instead of a single thread's R2UR chain (see cnn_umma2.cu / profiles)
    asm volatile(
        "{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile(
        "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        ::"r"(fxd::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor, K-major, SWIZZLE_NONE: core matrix = 8 rows x 16 bytes.
// lbo = byte distance between the two 8-element K chunks of one MMA, sbo = between 8-row groups.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

// kind::f16 instruction descriptor: D=F32 (bit 4), A=B=F16 (0), K-major both, N=32, M=128
constexpr uint32_t IDESC = (1u << 4) | ((32u >> 3) << 17) | ((128u >> 4) << 24);    // N = 32: A_lo x W_hi
constexpr uint32_t IDESC64 = (1u << 4) | ((64u >> 3) << 17) | ((128u >> 4) << 24);  // N = 64: A_hi x [W_hi | W_lo]

// one tile (128 rows) of an implicit-GEMM conv: taps x 2 channel pairs x {A_hi x [W_hi|W_lo] (N=64), A_lo x W_hi (N=32)}.
// The two 32-column halves of the accumulator are added in the epilogue; the activation planes are read twice per
// (tap, channel pair) instead of three times.
__device__ __forceinline__ void issue_conv_tile(uint32_t a_base, uint32_t a_plane, uint32_t w_base, int taps,
                                                int row0, uint32_t d_tmem, bool swap) {
    uint32_t first = 0;
    for (int j = 0; j < taps; ++j) {
#pragma unroll
        for (int kp = 0; kp < 2; ++kp) {
            const uint32_t a_hi = a_base + (uint32_t)(2 * kp) * a_plane + (uint32_t)(row0 + j) * 16u;
            const uint32_t a_lo = a_hi + 4u * a_plane;
            const uint32_t b_addr = w_base + (uint32_t)(j * 4 + 2 * kp) * WPLANE;
            const uint64_t bd = swap ? make_desc(b_addr, 128, WPLANE) : make_desc(b_addr, WPLANE, 128);
            umma_f16(d_tmem, swap ? make_desc(a_hi, 128, a_plane) : make_desc(a_hi, a_plane, 128), bd, IDESC64, first);
            umma_f16(d_tmem, swap ? make_desc(a_lo, 128, a_plane) : make_desc(a_lo, a_plane, 128), bd, IDESC, 1u);
            first = 1;
        }
    }
}

__device__ __forceinline__ void issue_idx_load(const UmmaParams &p, const Smem &sm, int64_t item, int buf) {
    const int64_t first = item * p.S;
    const int64_t cnt = min((int64_t)p.S, p.n - first);
    const uintptr_t g0 = reinterpret_cast<uintptr_t>(p.idx) + (uintptr_t)(first * p.d.L);
    const uintptr_t a0 = g0 & ~(uintptr_t)15;
    const uintptr_t a1 = (g0 + (uintptr_t)(cnt * p.d.L) + 15) & ~(uintptr_t)15;
    const uint32_t bytes = (uint32_t)(a1 - a0);
    fxd::mbar_arrive_expect_tx(&sm.mbar_idx[buf], bytes);
    fxd::bulk_g2s(sm.idx[buf], reinterpret_cast<const void *>(a0), bytes, &sm.mbar_idx[buf]);
}

// split 8 scaled activations into fp16 hi / lo and store the two 16-byte rows
__device__ __forceinline__ void store_split8(const float (&x)[8], unsigned char *hi_row, unsigned char *lo_row,
                                             bool &overflow) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float a = x[2 * i], b = x[2 * i + 1];
        overflow |= (a > 60000.f) | (b > 60000.f);
        const __half2 h = __floats2half2_rn(a, b);
        const float2 back = __half22float2(h);
        const __half2 l = __floats2half2_rn(a - back.x, b - back.y);
        hi[i] = *reinterpret_cast<const uint32_t *>(&h);
        lo[i] = *reinterpret_cast<const uint32_t *>(&l);
    }
    *reinterpret_cast<uint4 *>(hi_row) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4 *>(lo_row) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

__global__ void __launch_bounds__(NT, 1) cnn_umma_kernel(const UmmaParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    Smem sm;
    carve(smem_raw, p, &sm);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NWARP = NT / 32;
    const int T = p.d.T, P = p.P, hl = p.hl, L = p.d.L, A = p.d.A, K = p.d.K, K3 = p.d.K3;
    const int pl2 = p.d.pl2, pl3 = p.d.pl3;
    const uint32_t pl1_bytes = (uint32_t)p.rows1 * 16u, pl2_bytes = (uint32_t)p.rows2 * 16u;
    const uint32_t tmem_cols = 512;  // conv2: 4 tiles x 64 columns, conv3: the other 256

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) fxd::mbar_init(&sm.mbar_idx[i], 1);
        for (int i = 0; i < 4; ++i) { fxd::mbar_init(&sm.mbar_c2[i], 1); fxd::mbar_init(&sm.mbar_c3[i], 1); }
        fxd::fence_mbar_init();
    }
    if (wid == 0) tmem_alloc(sm.tmem_addr, tmem_cols);
    // zero both activation buffers once (tail rows are read by masked output rows only)
    for (int i = tid; i < 8 * p.rows1; i += NT) reinterpret_cast<uint4 *>(sm.a1)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < 8 * p.rows2; i += NT) reinterpret_cast<uint4 *>(sm.a2)[i] = make_uint4(0, 0, 0, 0);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *sm.tmem_addr;
    const uint32_t a1_addr = fxd::smem_u32(sm.a1), a2_addr = fxd::smem_u32(sm.a2);
    const uint32_t uw2_addr = fxd::smem_u32(sm.uw2), uw3_addr = fxd::smem_u32(sm.uw3);

    uint32_t iter = 0;   // idx loads consumed
    uint32_t chunk = 0;  // chunks processed (parity of the MMA barriers)
    bool overflow = false;
    long long pt[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // conv1, c2 phase, c2 wait, c3 phase, c3 wait, dense, chunks
    for (int mem = 0; mem < p.M; ++mem) {
        const float *w = p.weights + (int64_t)mem * p.member_floats;
        const unsigned char *uw = p.uw + (int64_t)mem * p.uw_member_bytes;
        const int w2b = K * 2 * WTAP, w3b = K3 * 2 * WTAP;
        const float inv2s = __ldg(reinterpret_cast<const float *>(uw + w2b + w3b));      // descale * ASCALE
        const float inv3 = __ldg(reinterpret_cast<const float *>(uw + w2b + w3b) + 1);   // descale
        __syncthreads();
        for (int i = tid; i < F; i += NT) {
            sm.b1[i] = __ldg(w + p.o.b1 + i);
            sm.b2s[i] = __ldg(w + p.o.b2 + i) * ASCALE;
            sm.b3[i] = __ldg(w + p.o.b3 + i);
        }
        for (int i = tid; i < K * A * F; i += NT) {
            const int row = i / F, f = i - row * F;
            sm.w1[row * W1P + f] = __ldg(w + p.o.w1 + i);
        }
        for (int i = tid; i < w2b / 16; i += NT)
            reinterpret_cast<uint4 *>(sm.uw2)[i] = __ldg(reinterpret_cast<const uint4 *>(uw) + i);
        for (int i = tid; i < w3b / 16; i += NT)
            reinterpret_cast<uint4 *>(sm.uw3)[i] = __ldg(reinterpret_cast<const uint4 *>(uw + w2b) + i);
        fence_async_smem();  // weights were written through the generic proxy, UMMA reads them through the async proxy
        __syncthreads();

        int nslots = 0;
        int64_t item = blockIdx.x;
        if (tid == 0 && item < p.n_items) issue_idx_load(p, sm, item, iter & 1);
        for (; item < p.n_items; item += gridDim.x, ++iter) {
            const int buf = iter & 1;
            const int64_t next = item + gridDim.x;
            if (tid == 0 && next < p.n_items) issue_idx_load(p, sm, next, buf ^ 1);
            const int64_t first = item * p.S;
            const int s_item = (int)min((int64_t)p.S, p.n - first);
            const int rows_item = s_item * P;
            if (nslots + s_item > p.sbcap) {
                fxd::DenseArgs da{w + p.o.wd1, w + p.o.bd1, w + p.o.wd2, w + p.o.bd2, w + p.o.wd3, w + p.o.bd3,
                                  sm.featT, reinterpret_cast<float *>(sm.a1), sm.slot_seq, p.out,
                                  F, p.d.H, p.sbp, nslots, mem, p.M, p.stage};
                const long long d0 = clock64();
                fxd::dense_head_flush<NT>(da);
                pt[5] += clock64() - d0;
                nslots = 0;
            }
            for (int i = tid; i < F * s_item; i += NT) sm.featT[(i / s_item) * p.sbp + nslots + (i % s_item)] = 0.f;
            for (int i = tid; i < s_item; i += NT) sm.slot_seq[nslots + i] = first + i;
            fxd::mbar_wait(&sm.mbar_idx[buf], (iter >> 1) & 1);
            const uint8_t *sidx = sm.idx[buf] + ((reinterpret_cast<uintptr_t>(p.idx) + (uintptr_t)(first * L)) & 15);

            for (int c0 = 0; c0 < rows_item; c0 += p.rout, ++chunk) {
                const int c1 = min(c0 + p.rout, rows_item);
                const int nout = c1 - c0;
                const int ntile3 = (nout + 127) >> 7;                      // conv3 tiles with live rows
                const int ntile2 = min(p.ntile, (nout + K3 - 1 + 127) >> 7);  // conv2 tiles conv3 needs
                const int nrows1 = ntile2 * 128 + K - 1;
                const uint32_t par = chunk & 1;
                const long long tA = clock64();
                // ---- conv1: gather-add -> A1 (rows [c0 - pl3 - pl2, ...)) ----
                for (int u = wid; u < 4 * ((nrows1 + 31) >> 5); u += NWARP) {
                    const int c = u & 3, r = (u >> 2) * 32 + lane;
                    if (r >= nrows1) continue;
                    const int rho = c0 - pl3 - pl2 + r;
                    const int sh = rho + 4 * P;
                    const int s = sh / P - 4;
                    const int t = sh - (s + 4) * P - hl;
                    const bool valid = (rho >= 0) && (rho < rows_item) && (t >= 0) && (t < T);
                    float x[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) x[i] = 0.f;
                    if (valid) {
                        const uint8_t *ip = sidx + s * L + t;
                        for (int j = 0; j < K; ++j) {
                            const float *tp = sm.w1 + ((size_t)j * A + ip[j]) * W1P + c * 8;
                            const float4 a0 = *reinterpret_cast<const float4 *>(tp);
                            const float4 a1 = *reinterpret_cast<const float4 *>(tp + 4);
                            x[0] += a0.x; x[1] += a0.y; x[2] += a0.z; x[3] += a0.w;
                            x[4] += a1.x; x[5] += a1.y; x[6] += a1.z; x[7] += a1.w;
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i) x[i] = fmaxf(x[i] + sm.b1[c * 8 + i], 0.f) * ASCALE;
                    }
                    store_split8(x, sm.a1 + (size_t)c * pl1_bytes + (size_t)r * 16,
                                 sm.a1 + (size_t)(4 + c) * pl1_bytes + (size_t)r * 16, overflow);
                }
                fence_async_smem();
                tc_fence_before();
                __syncthreads();
                const long long tB = clock64();
                // ---- conv2 on the tensor cores ----
                if (wid == 8) {  // every barrier completes one phase per chunk, used tile or not (parity = chunk & 1)
                    tc_fence_after();
                    for (int t = 0; t < 4; ++t) {
                        if (t < ntile2)
                            issue_conv_tile(a1_addr, pl1_bytes, uw2_addr, K, t * 128, tmem_base + (uint32_t)(t * 64),
                                            p.swap_lbo_sbo != 0);
                        umma_commit(&sm.mbar_c2[t]);
                    }
                }
                if (wid < 8) {
                    const int lq = wid & 3, ch = wid >> 2;
                    for (int t = 0; t < ntile2; ++t) {
                        const long long w0 = clock64();
                        fxd::mbar_wait(&sm.mbar_c2[t], par);
                        pt[2] += clock64() - w0;
                        tc_fence_after();
                        uint32_t v[16], vlo[16];
                        tmem_ld16(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(t * 64 + ch * 16), v);
                        tmem_ld16(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(t * 64 + 32 + ch * 16), vlo);
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(vlo[i]));
                        const int r = t * 128 + lq * 32 + lane;  // A2 row
                        const int rho = c0 - pl3 + r;
                        const int sh = rho + 4 * P;
                        const int s = sh / P - 4;
                        const int tt = sh - (s + 4) * P - hl;
                        const bool valid = (rho >= 0) && (rho < rows_item) && (tt >= 0) && (tt < T);
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
                            float x[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float acc = __uint_as_float(v[half * 8 + i]);
                                x[i] = valid ? fmaxf(fmaf(acc, inv2s, sm.b2s[ch * 16 + half * 8 + i]), 0.f) : 0.f;
                            }
                            const int c = ch * 2 + half;
                            store_split8(x, sm.a2 + (size_t)c * pl2_bytes + (size_t)r * 16,
                                         sm.a2 + (size_t)(4 + c) * pl2_bytes + (size_t)r * 16, overflow);
                        }
                    }
                }
                fence_async_smem();
                tc_fence_before();
                __syncthreads();
                const long long tC = clock64();
                // ---- conv3 on the tensor cores ----
                if (wid == 8) {
                    tc_fence_after();
                    for (int t = 0; t < 4; ++t) {
                        if (t < ntile3)
                            issue_conv_tile(a2_addr, pl2_bytes, uw3_addr, K3, t * 128, tmem_base + 256u + (uint32_t)(t * 64),
                                            p.swap_lbo_sbo != 0);
                        umma_commit(&sm.mbar_c3[t]);
                    }
                }
                if (wid < 8) {
                    const int lq = wid & 3, ch = wid >> 2;
                    for (int t = 0; t < ntile3; ++t) {
                        const long long w0 = clock64();
                        fxd::mbar_wait(&sm.mbar_c3[t], par);
                        pt[4] += clock64() - w0;
                        tc_fence_after();
                        uint32_t v[16], vlo[16];
                        tmem_ld16(tmem_base + ((uint32_t)(lq * 32) << 16) + 256u + (uint32_t)(t * 64 + ch * 16), v);
                        tmem_ld16(tmem_base + ((uint32_t)(lq * 32) << 16) + 256u + (uint32_t)(t * 64 + 32 + ch * 16), vlo);
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(vlo[i]));
                        const int r = t * 128 + lq * 32 + lane;  // output row of this chunk
                        const int rho = c0 + r;
                        const int s = rho / P;
                        const int tt = rho - s * P - hl;
                        const bool inrange = r < nout;
                        const bool valid = inrange && (tt >= 0) && (tt < T);
                        uint32_t bits[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float y = valid ? fmaxf(fmaf(__uint_as_float(v[i]), inv3, sm.b3[ch * 16 + i]), 0.f) : 0.f;
                            bits[i] = __float_as_uint(y);  // y >= 0: uint order == float order
                        }
                        // GlobalMaxPooling1D: max over the rows of each sequence present in this warp
                        unsigned todo = __ballot_sync(0xffffffffu, inrange);
                        while (todo) {
                            const int leader = __ffs(todo) - 1;
                            const int s_l = __shfl_sync(0xffffffffu, s, leader);
                            const bool mine = inrange && (s == s_l);
                            const unsigned seg = __ballot_sync(0xffffffffu, mine);
                            if (mine) {
                                uint32_t red[16];
#pragma unroll
                                for (int i = 0; i < 16; ++i) red[i] = __reduce_max_sync(seg, bits[i]);
                                if (lane == leader) {
                                    unsigned int *dst = reinterpret_cast<unsigned int *>(sm.featT) +
                                                        (size_t)(ch * 16) * p.sbp + nslots + s_l;
#pragma unroll
                                    for (int i = 0; i < 16; ++i) atomicMax(dst + (size_t)i * p.sbp, red[i]);
                                }
                            }
                            todo &= ~seg;
                        }
                    }
                }
                tc_fence_before();
                __syncthreads();
                const long long tD = clock64();
                pt[0] += tB - tA; pt[1] += tC - tB; pt[3] += tD - tC; pt[6] += 1;
            }
            nslots += s_item;
        }
        if (nslots > 0) {
            fxd::DenseArgs da{w + p.o.wd1, w + p.o.bd1, w + p.o.wd2, w + p.o.bd2, w + p.o.wd3, w + p.o.bd3,
                              sm.featT, reinterpret_cast<float *>(sm.a1), sm.slot_seq, p.out,
                              F, p.d.H, p.sbp, nslots, mem, p.M, p.stage};
            fxd::dense_head_flush<NT>(da);
        }
    }
    if (overflow) atomicExch(p.overflow_flag, 1);
    if (p.prof != nullptr && tid == 0)
        for (int i = 0; i < 8; ++i) p.prof[(size_t)blockIdx.x * 8 + i] = pt[i];
    tc_fence_before();
    __syncthreads();
    if (wid == 0) tmem_dealloc(tmem_base, tmem_cols);
}

static bool plan(const flexs_model *m, UmmaParams &p) {
    p.d = fx::cnn_dims(m);
    p.o = fx::cnn_offsets(m);
    p.M = m->M;
    p.member_floats = m->member_floats;
    const int K = p.d.K, K3 = p.d.K3, T = p.d.T;
    p.hl = std::max(p.d.pl2, p.d.pl3);
    const int hr = std::max(p.d.pr2, p.d.pr3);
    p.P = (p.hl + T + hr + 3) & ~3;
    p.uw_member_bytes = (int64_t)align_up((size_t)(K + K3) * 2 * WTAP + 16, 256);
    for (int ntile = 4; ntile >= 1; --ntile) {
        p.ntile = ntile;
        p.rout = (ntile * 128 - (K3 - 1)) & ~3;
        if (p.rout < 4) continue;
        p.rows1 = (int)align_up(ntile * 128 + K - 1 + 8, 8);
        p.rows2 = (int)align_up(ntile * 128 + K3 - 1 + 8, 8);
        int sbcap = 64;
        const size_t abytes = (size_t)8 * (p.rows1 + p.rows2) * 16;
        while (sbcap >= 8 && (size_t)2 * p.d.H * (sbcap + 4) * 4 > abytes) sbcap -= 8;
        if (sbcap < 8) continue;
        p.sbcap = sbcap; p.sbp = sbcap + 4;
        p.stage = fxd::dense_scratch_floats(F, p.d.H, p.sbp, true) * 4 <= abytes ? 1 : 0;
        p.S = std::max(1, std::min(p.rout / p.P, sbcap));
        p.idx_slot = (int)align_up((size_t)p.S * p.d.L + 32, 16);
        if ((int64_t)carve(nullptr, p, nullptr) + 1024 <= m->max_smem_optin) return true;
    }
    return false;
}

}  // namespace

namespace fx {

bool cnn_umma_supported(const flexs_model *m) {
    if (m->kind != FLEXS_KIND_CNN || m->F != 32 || m->K != 5) return false;
    if (m->K3 != 3 && m->K3 != 19) return false;
    if (!cnn_tiled_supported(m)) return false;  // the overflow fall-back path
    UmmaParams p;
    return plan(m, p);
}

// Re-lay conv2/conv3 weights of every member as fp16 UMMA planes [tap][chunk][hi filters | lo filters][8 ch].
int prepare_cnn_umma(flexs_model *m) {
    if (m->umma_ready) return FLEXS_OK;
    UmmaParams p;
    FX_REQUIRE(plan(m, p), "shape not supported by the UMMA kernel");
    const int K = p.d.K, K3 = p.d.K3;
    std::vector<float> host((size_t)m->member_floats * m->M);
    FX_CUDA(cudaMemcpy(host.data(), m->d_weights, host.size() * sizeof(float), cudaMemcpyDeviceToHost));
    std::vector<unsigned char> blob((size_t)p.uw_member_bytes * m->M, 0);
    m->umma_weights_ok = true;
    for (int mem = 0; mem < m->M; ++mem) {
        const float *w = host.data() + (size_t)mem * m->member_floats;
        unsigned char *dst = blob.data() + (size_t)mem * p.uw_member_bytes;
        float inv[2];
        for (int layer = 0; layer < 2; ++layer) {
            const int taps = layer == 0 ? K : K3;
            const float *src = w + (layer == 0 ? p.o.w2 : p.o.w3);  // (taps, in g, out f)
            float mx = 0.f;
            for (int i = 0; i < taps * F * F; ++i) {
                if (!std::isfinite(src[i])) m->umma_weights_ok = false;
                mx = std::max(mx, std::fabs(src[i]));
            }
            int e = 0;
            if (mx > 0.f && std::isfinite(mx)) e = 14 - (int)std::floor(std::log2(mx));  // scaled max in [2^14, 2^15)
            e = std::max(-24, std::min(e, 40));
            const float scale = std::ldexp(1.f, e);
            inv[layer] = std::ldexp(1.f, -e) / ASCALE;  // input activations carry ASCALE
            __half *planes = reinterpret_cast<__half *>(dst + (layer == 0 ? 0 : (size_t)K * 2 * WTAP));
            for (int j = 0; j < taps; ++j)
                for (int g = 0; g < F; ++g)
                    for (int f = 0; f < F; ++f) {
                        const float v = src[((size_t)j * F + g) * F + f] * scale;
                        const __half hi = __float2half_rn(v);
                        const __half lo = __float2half_rn(v - __half2float(hi));
                        const size_t base = ((size_t)j * 4 + (g >> 3)) * (WPLANE / 2) + (g & 7);
                        planes[base + (size_t)f * 8] = hi;
                        planes[base + (size_t)(32 + f) * 8] = lo;
                    }
        }
        float *tail = reinterpret_cast<float *>(dst + (size_t)(K + K3) * 2 * WTAP);
        tail[0] = inv[0] * ASCALE;  // conv2 epilogue writes activations pre-scaled by ASCALE
        tail[1] = inv[1];
    }
    FX_CUDA(cudaSetDevice(m->device));
    if (!m->d_umma_w) FX_CUDA(cudaMalloc(&m->d_umma_w, blob.size()));
    FX_CUDA(cudaMemcpy(m->d_umma_w, blob.data(), blob.size(), cudaMemcpyHostToDevice));
    m->umma_ready = true;
    return FLEXS_OK;
}

int launch_cnn_umma(flexs_model *m, const uint8_t *d_idx, int64_t n, float *d_out, cudaStream_t s) {
    // A = 4 shapes run the pipelined kernel of cnn_umma2.cu; this file's kernel serves k3 = 19
    static const bool force_v1 = std::getenv("FLEXS_UMMA_V1") && std::getenv("FLEXS_UMMA_V1")[0] == '1';
    if (!force_v1 && cnn_umma2_supported(m)) return launch_cnn_umma2(m, d_idx, n, d_out, s);
    if (!force_v1 && cnn_a20_supported(m)) return launch_cnn_a20(m, d_idx, n, d_out, s);  // k3 = 19: cnn_a20.cu
    UmmaParams p;
    FX_REQUIRE(cnn_umma_supported(m) && plan(m, p), "shape not supported by the UMMA kernel");
    int rc = prepare_cnn_umma(m);
    if (rc != FLEXS_OK) return rc;
    if (!m->umma_weights_ok) return launch_cnn_tiled(m, d_idx, n, d_out, s);  // non-finite weights: fp32 path
    p.idx = d_idx; p.out = d_out; p.weights = m->d_weights; p.n = n;
    p.uw = reinterpret_cast<const unsigned char *>(m->d_umma_w);
    flexs_model::StreamWs *ws = nullptr;  // the fp16-overflow flag is per stream: chunks of score_host run concurrently
    rc = stream_workspace(m, s, 0, &ws);
    if (rc != FLEXS_OK) return rc;
    p.overflow_flag = ws->flag;
    p.n_items = (n + p.S - 1) / p.S;
    static const bool swap = std::getenv("FLEXS_UMMA_SWAP") && std::getenv("FLEXS_UMMA_SWAP")[0] == '1';
    p.swap_lbo_sbo = swap ? 1 : 0;
    const size_t smem = carve(nullptr, p, nullptr) + 1024;
    const int grid = (int)std::min<int64_t>(p.n_items, m->sm_count);
    static const bool prof = std::getenv("FLEXS_UMMA_PROF") && std::getenv("FLEXS_UMMA_PROF")[0] == '1';
    p.prof = nullptr;
    if (prof) {
        FX_CUDA(cudaMalloc(&p.prof, (size_t)grid * 8 * sizeof(long long)));
        FX_CUDA(cudaMemset(p.prof, 0, (size_t)grid * 8 * sizeof(long long)));
    }
    FX_CUDA(cudaMemsetAsync(ws->flag, 0, sizeof(int), s));
    FX_CUDA(cudaFuncSetAttribute(cnn_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cnn_umma_kernel<<<grid, NT, smem, s>>>(p);
    FX_CUDA(cudaGetLastError());
    m->launches += 1;
    if (prof) {
        FX_CUDA(cudaStreamSynchronize(s));
        std::vector<long long> h((size_t)grid * 8);
        FX_CUDA(cudaMemcpy(h.data(), p.prof, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        cudaFree(p.prof);
        double a[8] = {0};
        for (int b = 0; b < grid; ++b) for (int i = 0; i < 8; ++i) a[i] += (double)h[(size_t)b * 8 + i] / grid;
        const double ch = a[6] > 0 ? a[6] : 1;
        fprintf(stderr, "[umma prof] n=%lld grid=%d chunks/CTA=%.0f | cycles per chunk: conv1 %.0f, conv2 phase %.0f (mma wait %.0f), "
                        "conv3 phase %.0f (mma wait %.0f), dense (amortised) %.0f\n",
                (long long)n, grid, a[6], a[0] / ch, a[1] / ch, a[2] / ch, a[3] / ch, a[4] / ch, a[5] / ch);
    }
    // fp16 range guard: the gated FFMA kernel recomputes the batch iff the flag was raised
    return launch_cnn_tiled_gated(m, d_idx, n, d_out, ws->flag, s);
}

}  // namespace fx
