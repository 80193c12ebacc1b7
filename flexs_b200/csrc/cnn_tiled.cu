// K1 (FP32 FFMA variant): fused CNN forward for the paper-script shapes (F = 32, k = 5,
// A in {4, 20}) — cnn.py:23-54 evaluated end to end inside one persistent kernel:
//
//   uint8 residue indices --1-D bulk async copy (TMA engine)--> shared memory
//   conv1  gather-add from a [k][A][F] table (the one-hot never exists)          -> h1 (smem)
//   conv2  register-tiled FFMA implicit GEMM, 4 rows x 8 filters per thread      -> h2 (smem)
//   conv3  same tile shape, ReLU + max over the thread's 4 rows                   -> pmax (smem)
//   global max over time per (sequence, filter)                                   -> featT (smem)
//   dense 32->H->H->1 (batched over up to 64 sequences), nan_to_num, ensemble mean -> out (HBM)
//
// HBM traffic is L bytes in + 4 bytes out per sequence; everything else lives on chip.
//
// Row space.  A CTA works on an "item" of S consecutive sequences.  Each sequence owns P rows:
// hl zero rows, its T conv positions, then zero rows up to P (P % 4 == 0), so the "same" padding
// of conv2/conv3 is just the neighbouring zero rows and a tap is a +1 row shift.  Activations are
// stored channel-major [g][row] so a thread's 4 consecutive rows (+taps) are 16-byte vector loads
// that are conflict-free across a warp, and the weight vector for (tap, g) is a warp broadcast.
#include <algorithm>

#include "common.cuh"
#include "dense_head.cuh"

namespace {

constexpr int F = 32;
constexpr int FPT = 8;        // filters per thread
constexpr int NFG = F / FPT;  // filter groups
constexpr int W1P = 36;       // padded row (floats) of the conv1 gather table
constexpr int NT = 512;

struct TiledParams {
    const uint8_t *idx;
    float *out;
    const float *weights;
    const int *gate;  // optional: run only if *gate != 0
    int64_t n, n_items, member_floats;
    fx::CnnDims d;
    fx::CnnOffsets o;
    int M;
    int P, hl, S;
    int rcap;    // rows capacity of h1/h2 (multiple of 128)
    int rout;    // conv3 output rows per chunk
    int sbcap;   // feature slots batched for the dense head (multiple of 8)
    int sbp;     // slot pitch of the dense buffers = sbcap + 4
    int idx_slot;  // bytes per idx staging slot
    int stage;     // dense head stages Wd1/Wd2 in shared memory
};

struct Smem {
    uint64_t *mbar;   // [2]
    float *b1, *b2, *b3;
    float *w1, *w2, *w3;
    uint8_t *idx[2];
    float *featT;     // [32][sbp]
    long long *slot_seq;  // [sbcap]
    float *h1, *h2;   // [32][rcap] each, contiguous; pmax and the dense scratch alias them
};

__device__ __forceinline__ Smem carve(unsigned char *base, const TiledParams &p) {
    Smem s;
    size_t off = 0;
    s.mbar = reinterpret_cast<uint64_t *>(base + off); off += 16;
    s.b1 = reinterpret_cast<float *>(base + off); off += F * 4;
    s.b2 = reinterpret_cast<float *>(base + off); off += F * 4;
    s.b3 = reinterpret_cast<float *>(base + off); off += F * 4;
    s.w1 = reinterpret_cast<float *>(base + off); off += (size_t)p.d.K * p.d.A * W1P * 4;
    s.w2 = reinterpret_cast<float *>(base + off); off += (size_t)p.d.K * F * F * 4;
    s.w3 = reinterpret_cast<float *>(base + off); off += (size_t)p.d.K3 * F * F * 4;
    s.idx[0] = base + off; off += p.idx_slot;
    s.idx[1] = base + off; off += p.idx_slot;
    s.featT = reinterpret_cast<float *>(base + off); off += (size_t)F * p.sbp * 4;
    s.slot_seq = reinterpret_cast<long long *>(base + off); off += (size_t)p.sbcap * 8;
    s.h1 = reinterpret_cast<float *>(base + off); off += (size_t)F * p.rcap * 4;
    s.h2 = reinterpret_cast<float *>(base + off);
    return s;
}

static size_t smem_bytes(const TiledParams &p) {
    return 16 + 3 * F * 4 + (size_t)p.d.K * p.d.A * W1P * 4 + (size_t)p.d.K * F * F * 4 +
           (size_t)p.d.K3 * F * F * 4 + 2 * (size_t)p.idx_slot + (size_t)F * p.sbp * 4 +
           (size_t)p.sbcap * 8 + 2 * (size_t)F * p.rcap * 4;
}

// acc[r][f] += sum_j sum_g x[g][4q + r + j] * w[j][g][f0 + f]
template <int KW>
__device__ __forceinline__ void conv_quad(const float *__restrict__ xs, int pitch,
                                          const float *__restrict__ ws, float (&acc)[4][FPT]) {
    constexpr int NX4 = (4 + KW - 1 + 3) / 4;
#pragma unroll 1
    for (int g = 0; g < F; ++g) {
        float x[NX4 * 4];
#pragma unroll
        for (int i = 0; i < NX4; ++i) {
            const float4 v = *reinterpret_cast<const float4 *>(xs + (size_t)g * pitch + 4 * i);
            x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
        }
#pragma unroll
        for (int j = 0; j < KW; ++j) {
            float w[FPT];
#pragma unroll
            for (int i = 0; i < FPT / 4; ++i) {
                const float4 v = *reinterpret_cast<const float4 *>(ws + ((size_t)j * F + g) * F + 4 * i);
                w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
            }
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int f = 0; f < FPT; ++f) acc[r][f] = fmaf(x[r + j], w[f], acc[r][f]);
        }
    }
}

// (sequence-in-item, conv position, validity) of 4 consecutive item rows starting at rho0
struct Quad {
    int s[4], t[4];
    bool v[4];
};
__device__ __forceinline__ Quad quad_info(int rho0, int rows_item, int P, int hl, int T) {
    Quad q;
    // rho0 may be slightly negative; shift into the non-negative range for the division
    const int sh = rho0 + 4 * P;
    int s = sh / P - 4;
    int r = sh - (s + 4) * P;  // row inside the sequence, 0..P-1
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int rho = rho0 + i;
        q.s[i] = s; q.t[i] = r - hl;
        q.v[i] = (rho >= 0) && (rho < rows_item) && (q.t[i] >= 0) && (q.t[i] < T);
        if (++r == P) { r = 0; ++s; }
    }
    return q;
}

__device__ __forceinline__ void issue_idx_load(const TiledParams &p, const Smem &sm, int64_t item, int buf) {
    const int64_t first = item * p.S;
    const int64_t cnt = min((int64_t)p.S, p.n - first);
    const uintptr_t g0 = reinterpret_cast<uintptr_t>(p.idx) + (uintptr_t)(first * p.d.L);
    const uintptr_t a0 = g0 & ~(uintptr_t)15;
    const uintptr_t a1 = (g0 + (uintptr_t)(cnt * p.d.L) + 15) & ~(uintptr_t)15;
    const uint32_t bytes = (uint32_t)(a1 - a0);
    fxd::mbar_arrive_expect_tx(&sm.mbar[buf], bytes);
    fxd::bulk_g2s(sm.idx[buf], reinterpret_cast<const void *>(a0), bytes, &sm.mbar[buf]);
}

template <int K, int K3>
__global__ void __launch_bounds__(NT, 1) cnn_tiled_kernel(const TiledParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    if (p.gate != nullptr && *p.gate == 0) return;  // uniform across the grid
    const Smem sm = carve(smem_raw, p);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NWARP = NT / 32;
    const int T = p.d.T, P = p.P, hl = p.hl, L = p.d.L, A = p.d.A;
    const int pl2 = p.d.pl2, pl3 = p.d.pl3;
    const int rcap = p.rcap;
    const int pmp = rcap / 4 + 1;  // pitch of pmax[f][quad] (odd -> conflict-free column reads)
    float *pmax = sm.h1;

    if (tid == 0) {
        fxd::mbar_init(&sm.mbar[0], 1);
        fxd::mbar_init(&sm.mbar[1], 1);
        fxd::fence_mbar_init();
    }
    __syncthreads();

    uint32_t iter = 0;  // running count of idx loads consumed (selects buffer + parity)
    for (int mem = 0; mem < p.M; ++mem) {
        const float *w = p.weights + (int64_t)mem * p.member_floats;
        // ---- stage this member's conv weights in shared memory ----
        __syncthreads();
        for (int i = tid; i < F; i += NT) {
            sm.b1[i] = __ldg(w + p.o.b1 + i);
            sm.b2[i] = __ldg(w + p.o.b2 + i);
            sm.b3[i] = __ldg(w + p.o.b3 + i);
        }
        for (int i = tid; i < K * A * F; i += NT) {
            const int row = i / F, f = i - row * F;
            sm.w1[row * W1P + f] = __ldg(w + p.o.w1 + i);
        }
        for (int i = tid; i < K * F * F / 4; i += NT)
            reinterpret_cast<float4 *>(sm.w2)[i] = __ldg(reinterpret_cast<const float4 *>(w + p.o.w2) + i);
        for (int i = tid; i < K3 * F * F / 4; i += NT)
            reinterpret_cast<float4 *>(sm.w3)[i] = __ldg(reinterpret_cast<const float4 *>(w + p.o.w3) + i);
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
                                  sm.featT, sm.h1, sm.slot_seq, p.out, F, p.d.H, p.sbp, nslots, mem, p.M, p.stage};
                fxd::dense_head_flush<NT>(da);
                nslots = 0;
            }
            for (int i = tid; i < F * s_item; i += NT) sm.featT[(i / s_item) * p.sbp + nslots + (i % s_item)] = 0.f;
            for (int i = tid; i < s_item; i += NT) sm.slot_seq[nslots + i] = first + i;
            fxd::mbar_wait(&sm.mbar[buf], (iter >> 1) & 1);
            const uint8_t *sidx = sm.idx[buf] +
                                  ((reinterpret_cast<uintptr_t>(p.idx) + (uintptr_t)(first * L)) & 15);

            for (int c0 = 0; c0 < rows_item; c0 += p.rout) {
                const int c1 = min(c0 + p.rout, rows_item);
                const int nq3 = (c1 - c0) >> 2;
                const int nq2 = (c1 - c0 + K3 - 1 + 3) >> 2;
                const int nq1 = nq2 + (K - 1 + 3) / 4;
                // ---- conv1: gather-add, rows [c0 - pl3 - pl2, ...) ----
                for (int u = wid; u < NFG * ((nq1 + 31) >> 5); u += NWARP) {
                    const int f0 = (u & (NFG - 1)) * FPT, q = (u >> 2) * 32 + lane;
                    if (q >= nq1) continue;
                    const Quad qi = quad_info(c0 - pl3 - pl2 + 4 * q, rows_item, P, hl, T);
                    float acc[4][FPT];
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
#pragma unroll
                        for (int f = 0; f < FPT; ++f) acc[r][f] = 0.f;
                        if (qi.v[r]) {
                            const uint8_t *ip = sidx + qi.s[r] * L + qi.t[r];
#pragma unroll
                            for (int j = 0; j < K; ++j) {
                                const float *tp = sm.w1 + ((size_t)j * A + ip[j]) * W1P + f0;
                                const float4 a0 = *reinterpret_cast<const float4 *>(tp);
                                const float4 a1 = *reinterpret_cast<const float4 *>(tp + 4);
                                acc[r][0] += a0.x; acc[r][1] += a0.y; acc[r][2] += a0.z; acc[r][3] += a0.w;
                                acc[r][4] += a1.x; acc[r][5] += a1.y; acc[r][6] += a1.z; acc[r][7] += a1.w;
                            }
                        }
                    }
#pragma unroll
                    for (int f = 0; f < FPT; ++f) {
                        const float b = sm.b1[f0 + f];
                        float4 v;
                        v.x = qi.v[0] ? fmaxf(acc[0][f] + b, 0.f) : 0.f;
                        v.y = qi.v[1] ? fmaxf(acc[1][f] + b, 0.f) : 0.f;
                        v.z = qi.v[2] ? fmaxf(acc[2][f] + b, 0.f) : 0.f;
                        v.w = qi.v[3] ? fmaxf(acc[3][f] + b, 0.f) : 0.f;
                        *reinterpret_cast<float4 *>(sm.h1 + (size_t)(f0 + f) * rcap + 4 * q) = v;
                    }
                }
                __syncthreads();
                // ---- conv2: rows [c0 - pl3, ...) ----
                for (int u = wid; u < NFG * ((nq2 + 31) >> 5); u += NWARP) {
                    const int f0 = (u & (NFG - 1)) * FPT, q = (u >> 2) * 32 + lane;
                    if (q >= nq2) continue;
                    const Quad qi = quad_info(c0 - pl3 + 4 * q, rows_item, P, hl, T);
                    float acc[4][FPT];
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int f = 0; f < FPT; ++f) acc[r][f] = 0.f;
                    if (qi.v[0] | qi.v[1] | qi.v[2] | qi.v[3])
                        conv_quad<K>(sm.h1 + 4 * q, rcap, sm.w2 + f0, acc);
#pragma unroll
                    for (int f = 0; f < FPT; ++f) {
                        const float b = sm.b2[f0 + f];
                        float4 v;
                        v.x = qi.v[0] ? fmaxf(acc[0][f] + b, 0.f) : 0.f;
                        v.y = qi.v[1] ? fmaxf(acc[1][f] + b, 0.f) : 0.f;
                        v.z = qi.v[2] ? fmaxf(acc[2][f] + b, 0.f) : 0.f;
                        v.w = qi.v[3] ? fmaxf(acc[3][f] + b, 0.f) : 0.f;
                        *reinterpret_cast<float4 *>(sm.h2 + (size_t)(f0 + f) * rcap + 4 * q) = v;
                    }
                }
                __syncthreads();
                // ---- conv3 + ReLU + max over the thread's 4 rows: rows [c0, c1) ----
                for (int u = wid; u < NFG * ((nq3 + 31) >> 5); u += NWARP) {
                    const int f0 = (u & (NFG - 1)) * FPT, q = (u >> 2) * 32 + lane;
                    if (q >= nq3) continue;
                    const Quad qi = quad_info(c0 + 4 * q, rows_item, P, hl, T);
                    float acc[4][FPT];
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int f = 0; f < FPT; ++f) acc[r][f] = 0.f;
                    if (qi.v[0] | qi.v[1] | qi.v[2] | qi.v[3])
                        conv_quad<K3>(sm.h2 + 4 * q, rcap, sm.w3 + f0, acc);
#pragma unroll
                    for (int f = 0; f < FPT; ++f) {
                        const float b = sm.b3[f0 + f];
                        float mx = 0.f;  // ReLU output is >= 0, so 0 is the identity of the max
#pragma unroll
                        for (int r = 0; r < 4; ++r) mx = qi.v[r] ? fmaxf(mx, acc[r][f] + b) : mx;
                        pmax[(size_t)(f0 + f) * pmp + q] = mx;
                    }
                }
                __syncthreads();
                // ---- GlobalMaxPooling1D: fold this chunk's quads into featT ----
                {
                    const int sA = c0 / P, sB = (c1 - 1) / P;
                    for (int e = tid; e < (sB - sA + 1) * F; e += NT) {
                        const int s = sA + (e >> 5), f = e & 31;
                        const int qa = (max(s * P, c0) - c0) >> 2, qb = (min((s + 1) * P, c1) - c0) >> 2;
                        float mx = sm.featT[f * p.sbp + nslots + s];
                        for (int q = qa; q < qb; ++q) mx = fmaxf(mx, pmax[(size_t)f * pmp + q]);
                        sm.featT[f * p.sbp + nslots + s] = mx;
                    }
                }
                __syncthreads();
            }
            nslots += s_item;
        }
        if (nslots > 0) {
            fxd::DenseArgs da{w + p.o.wd1, w + p.o.bd1, w + p.o.wd2, w + p.o.bd2, w + p.o.wd3, w + p.o.bd3,
                              sm.featT, sm.h1, sm.slot_seq, p.out, F, p.d.H, p.sbp, nslots, mem, p.M, p.stage};
            fxd::dense_head_flush<NT>(da);
        }
    }
}

static bool plan(const flexs_model *m, TiledParams &p) {
    p.d = fx::cnn_dims(m);
    p.o = fx::cnn_offsets(m);
    p.M = m->M;
    p.member_floats = m->member_floats;
    const int K = p.d.K, K3 = p.d.K3, T = p.d.T;
    p.hl = std::max(p.d.pl2, p.d.pl3);
    const int hr = std::max(p.d.pr2, p.d.pr3);
    p.P = (p.hl + T + hr + 3) & ~3;
    for (int rcap = 512; rcap >= 128; rcap -= 128) {
        p.rcap = rcap;
        const int nq2max = rcap / 4 - (K - 1 + 3) / 4;
        p.rout = (4 * nq2max - (K3 - 1)) & ~3;
        if (p.rout < 4) continue;
        // dense scratch (2 x [H][sbp]) aliases h1+h2
        int sbcap = 64;
        while (sbcap >= 8 && (size_t)2 * p.d.H * (sbcap + 4) > (size_t)2 * F * rcap) sbcap -= 8;
        if (sbcap < 8) continue;
        p.sbcap = sbcap; p.sbp = sbcap + 4;
        p.stage = fxd::dense_scratch_floats(F, p.d.H, p.sbp, true) <= (size_t)2 * F * rcap ? 1 : 0;
        p.S = std::max(1, std::min(p.rout / p.P, sbcap));
        p.idx_slot = ((p.S * p.d.L + 32) + 15) & ~15;
        if ((int64_t)smem_bytes(p) <= m->max_smem_optin) return true;
    }
    return false;
}

}  // namespace

namespace fx {

bool cnn_tiled_supported(const flexs_model *m) {
    if (m->kind != FLEXS_KIND_CNN || m->F != 32 || m->K != 5) return false;
    if (m->K3 != 3 && m->K3 != 19) return false;
    TiledParams p;
    return plan(m, p);
}

int launch_cnn_tiled(flexs_model *m, const uint8_t *d_idx, int64_t n, float *d_out, cudaStream_t s) {
    return launch_cnn_tiled_gated(m, d_idx, n, d_out, nullptr, s);
}

int launch_cnn_tiled_gated(flexs_model *m, const uint8_t *d_idx, int64_t n, float *d_out, const int *d_gate,
                           cudaStream_t s) {
    TiledParams p;
    FX_REQUIRE(cnn_tiled_supported(m) && plan(m, p), "shape not supported by the tiled CNN kernel");
    p.idx = d_idx; p.out = d_out; p.weights = m->d_weights; p.n = n; p.gate = d_gate;
    p.n_items = (n + p.S - 1) / p.S;
    const size_t smem = smem_bytes(p);
    const int grid = (int)std::min<int64_t>(p.n_items, m->sm_count);
    if (m->K3 == 3) {
        FX_CUDA(cudaFuncSetAttribute(cnn_tiled_kernel<5, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cnn_tiled_kernel<5, 3><<<grid, NT, smem, s>>>(p);
    } else {
        FX_CUDA(cudaFuncSetAttribute(cnn_tiled_kernel<5, 19>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cnn_tiled_kernel<5, 19><<<grid, NT, smem, s>>>(p);
    }
    FX_CUDA(cudaGetLastError());
    m->launches += 1;
    return FLEXS_OK;
}

}  // namespace fx
