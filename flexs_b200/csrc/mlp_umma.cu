// K2b (tcgen05): MLP forward (mlp.py:21-31) with EVERY layer on the tensor cores, H <= 112.
//
// Flatten -> Dense(H, relu) on a one-hot input is a genuine contraction over K1 = L * A inputs of which L are one: the
// FFMA kernel (mlp.cu) evaluates it as L row gathers per sequence and is bound by L1 bandwidth (one 4-byte load per add).
// Here a tile of 128 sequences is the M side of tcgen05.mma: the one-hot is built on the fly in shared memory as an fp16
// operand — 0 and 1 are exact, so only the weights need the hi/lo split and the layer costs ONE N = 224 MMA per 16
// inputs ([W_hi | W_lo] fused along N) instead of the three products of a float x float layer — and W1 streams through
// shared memory in blocks of 32 inputs (bulk async copies, double-buffered, L2-resident).  Layers 2 and 3 are the
// K = 112 GEMM of the CNN's dense head (umma2_layout.cuh: fp16 hi/lo activations x fp16 hi/lo weights, three products),
// the output layer a dot product in the last epilogue, followed by nan_to_num (keras_model.py:77) and the ensemble mean
// (ensemble.py:54-59).  FP32 accumulation in TMEM throughout.
//
// 9 warps: 0-7 build the one-hot blocks and run the epilogues, 8 issues the MMAs.  Building block b + 1 overlaps the
// MMAs of block b (two buffers, completion through tcgen05.commit); the three layers of a tile are sequential.
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

constexpr int NT = 288, MMAW = 8, NEPI = 256;
constexpr int TS = DSLOTS;                  // sequences per tile
constexpr int KB_CHUNKS = 4;                // 8-input chunks per streamed block (32 inputs = 2 K steps)
constexpr int ABLK = KB_CHUNKS * DPLANE, BBLK = KB_CHUNKS * DBK;

// operand blob of one member: W1 planes (padded to whole blocks) | W2 planes | W3 planes | vectors
constexpr int MV_FLOATS = 4 * DH + 8;       // b1 | b2*ASCALE.. see prepare | w4 | scalars

// shared memory map
constexpr int S_BAR = 0, S_TM = 64, S_VEC = 128;                       // mbarriers, tmem address, vectors (MV_FLOATS floats)
constexpr int S_B2 = 2048, S_B3 = S_B2 + 14 * DBK;                      // resident weight planes of layers 2 and 3
constexpr int S_X = S_B3 + 14 * DBK;                                    // activation planes (28 x DPLANE) — aliased by the
constexpr int S_XEND = S_X + 28 * DPLANE;                               //   layer-1 streaming buffers: 2 x (ABLK + BBLK)
constexpr int S_PART = S_XEND, S_IDX = S_PART + 2 * TS * 4;
static_assert(2 * (ABLK + BBLK) <= 28 * DPLANE, "streaming buffers must fit the activation planes they alias");
static_assert(S_B2 % 1024 == 0 && S_X % 128 == 0, "operand alignment");

struct MlpUParams {
    const uint8_t *idx;
    float *out;
    const unsigned char *uw;   // this member's blob
    int *overflow_flag;
    int64_t n, n_tiles;
    int L, A, H, K1, nblk, mem, M;
    int off_w2, off_w3, off_vec;   // blob offsets
};

__global__ void __launch_bounds__(NT, 1) mlp_umma_kernel(const MlpUParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + S_BAR);    // [0,1] B block landed, [2,3] block MMAs retired, [4] layer done
    uint32_t *tmem_addr_s = reinterpret_cast<uint32_t *>(smem_raw + S_TM);
    float *vec = reinterpret_cast<float *>(smem_raw + S_VEC);
    float *dpart = reinterpret_cast<float *>(smem_raw + S_PART);
    uint8_t *sidx = smem_raw + S_IDX;
    const int tid = threadIdx.x, lane = tid & 31;
    const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform for the compiler: MMA issue on the uniform datapath
    const int L = p.L, A = p.A, K1 = p.K1;

    if (tid == 0) {
        for (int i = 0; i < 5; ++i) fxd::mbar_init(&bar[i], 1);
        fxd::fence_mbar_init();
    }
    if (wid == 0) tmem_alloc(tmem_addr_s, 512);
    for (int i = tid; i < MV_FLOATS; i += NT) vec[i] = __ldg(reinterpret_cast<const float *>(p.uw + p.off_vec) + i);
    for (int i = tid; i < 14 * DBK / 16; i += NT) {
        reinterpret_cast<uint4 *>(smem_raw + S_B2)[i] = __ldg(reinterpret_cast<const uint4 *>(p.uw + p.off_w2) + i);
        reinterpret_cast<uint4 *>(smem_raw + S_B3)[i] = __ldg(reinterpret_cast<const uint4 *>(p.uw + p.off_w3) + i);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_addr_s;
    const uint32_t x_addr = fxd::smem_u32(smem_raw + S_X), b2_addr = fxd::smem_u32(smem_raw + S_B2), b3_addr = fxd::smem_u32(smem_raw + S_B3);
    const float inv1 = vec[4 * DH], inv2s = vec[4 * DH + 1], inv3 = vec[4 * DH + 2], b4 = vec[4 * DH + 3];
    const float *b1v = vec, *b2v = vec + DH, *b3v = vec + 2 * DH, *w4v = vec + 3 * DH;

    uint32_t nblocks = 0;   // streamed blocks so far (buffer = nblocks & 1, its barriers' phase = (nblocks >> 1) & 1)
    uint32_t nlayer = 0;    // completions of bar[4]
    float xmax = 0.f;
    const int lq = wid & 3, half = wid >> 2, slot = 32 * lq + lane;   // epilogue roles of warps 0-7
    const uint32_t tl = tmem_base + ((uint32_t)(lq * 32) << 16);

    for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const int64_t first = tile * TS;
        const int cnt = (int)min((int64_t)TS, p.n - first);
        // residues of the tile (the previous tile's last reader was its layer-1 build, two barriers ago)
        for (int i = tid; i < TS * L; i += NT) sidx[i] = (i < cnt * L) ? __ldg(p.idx + first * L + i) : (uint8_t)0xff;
        __syncthreads();

        // ---------------- layer 1: one-hot [128, K1] x [W1_hi | W1_lo], streamed in blocks of 32 inputs ----------------
        for (int blk = 0; blk < p.nblk; ++blk, ++nblocks) {
            const uint32_t buf = nblocks & 1u;
            unsigned char *abuf = smem_raw + S_X + buf * (ABLK + BBLK), *bbuf = abuf + ABLK;
            if (nblocks >= 2) fxd::mbar_wait(&bar[2 + buf], ((nblocks >> 1) - 1) & 1);   // the MMAs that read this buffer retired
            if (tid == 0) {
                fxd::mbar_arrive_expect_tx(&bar[buf], BBLK);
                fxd::bulk_g2s(bbuf, p.uw + (size_t)blk * BBLK, BBLK, &bar[buf]);
            }
            if (wid < MMAW) {
                // one 16-byte row of 8 inputs per (slot, chunk): input k = l * A + c is 1 iff residue l of the sequence is c
                for (int task = tid; task < TS * KB_CHUNKS; task += NEPI) {
                    const int s = task & (TS - 1), j = task >> 7;
                    const int k0 = (blk * KB_CHUNKS + j) * 8;
                    const uint8_t *row = sidx + s * L;
                    uint32_t h[4] = {0, 0, 0, 0};
                    int l = k0 / A, c = k0 - l * A;
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        if (k0 + e < K1 && row[l] == c) h[e >> 1] |= 0x3C00u << (16 * (e & 1));   // fp16 1.0
                        if (++c == A) { c = 0; ++l; }
                    }
                    *reinterpret_cast<uint4 *>(abuf + j * DPLANE + s * 16) = make_uint4(h[0], h[1], h[2], h[3]);
                }
                fence_async_smem();
            }
            __syncthreads();
            if (wid == MMAW) {
                fxd::mbar_wait_warp(&bar[buf], (nblocks >> 1) & 1);
                tc_fence_after();
                const uint32_t a0 = desc_lo(fxd::smem_u32(abuf), DPLANE), b0 = desc_lo(fxd::smem_u32(bbuf), DBK);
#pragma unroll
                for (int ks = 0; ks < KB_CHUNKS / 2; ++ks)
                    umma_f16_elect(tmem_base, a0 + (((uint32_t)(2 * ks) * DPLANE) >> 4), DESC_HI,
                                   b0 + (((uint32_t)(2 * ks) * DBK) >> 4), DESC_HI, IDESC_DN, (blk | ks) ? 1u : 0u);
                umma_commit_elect(&bar[2 + buf]);
                if (blk == p.nblk - 1) umma_commit_elect(&bar[4]);
            }
        }
        // ---------------- epilogue 1: bias, ReLU, split -> activation planes (the streaming buffers are dead now) ----------------
        if (wid < MMAW) {
            fxd::mbar_wait(&bar[4], nlayer & 1);
            tc_fence_after();
            for (int c7 = 0; c7 < 7; ++c7) {
                const int cchunk = half * 7 + c7;
                uint32_t va[8], vb[8];
                tmem_ld8_nowait(tl + (uint32_t)(cchunk * 8), va);
                tmem_ld8_nowait(tl + (uint32_t)(DH + cchunk * 8), vb);
                tmem_ld_wait();
                float x[8];
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    x[q] = fmaxf(fmaf(__uint_as_float(va[q]) + __uint_as_float(vb[q]), inv1, b1v[cchunk * 8 + q]), 0.f) * ASCALE;
                uint4 hi4, lo4;
                split8(x, hi4, lo4, xmax);
                *reinterpret_cast<uint4 *>(smem_raw + S_X + (size_t)cchunk * DPLANE + slot * 16) = hi4;
                *reinterpret_cast<uint4 *>(smem_raw + S_X + (size_t)(14 + cchunk) * DPLANE + slot * 16) = lo4;
            }
        }
        ++nlayer;
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        // ---------------- layer 2 ----------------
        if (wid == MMAW) {
            tc_fence_after();
            issue_dense_layer<7, 14>(x_addr, b2_addr, tmem_base + 256u);
            umma_commit_elect(&bar[4]);
        } else {
            fxd::mbar_wait(&bar[4], nlayer & 1);
            tc_fence_after();
            for (int c7 = 0; c7 < 7; ++c7) {
                const int cchunk = half * 7 + c7;
                uint32_t va[8], vb[8];
                tmem_ld8_nowait(tl + 256u + (uint32_t)(cchunk * 8), va);
                tmem_ld8_nowait(tl + 256u + (uint32_t)(DH + cchunk * 8), vb);
                tmem_ld_wait();
                float x[8];
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    x[q] = fmaxf(fmaf(__uint_as_float(va[q]) + __uint_as_float(vb[q]), inv2s, b2v[cchunk * 8 + q]), 0.f);
                uint4 hi4, lo4;
                split8(x, hi4, lo4, xmax);
                *reinterpret_cast<uint4 *>(smem_raw + S_X + (size_t)cchunk * DPLANE + slot * 16) = hi4;
                *reinterpret_cast<uint4 *>(smem_raw + S_X + (size_t)(14 + cchunk) * DPLANE + slot * 16) = lo4;
            }
        }
        ++nlayer;
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        // ---------------- layer 3 + output layer ----------------
        if (wid == MMAW) {
            tc_fence_after();
            issue_dense_layer<7, 14>(x_addr, b3_addr, tmem_base);
            umma_commit_elect(&bar[4]);
        } else {
            fxd::mbar_wait(&bar[4], nlayer & 1);
            tc_fence_after();
            float sum = 0.f;
            for (int c7 = 0; c7 < 7; ++c7) {
                const int cchunk = half * 7 + c7;
                uint32_t va[8], vb[8];
                tmem_ld8_nowait(tl + (uint32_t)(cchunk * 8), va);
                tmem_ld8_nowait(tl + (uint32_t)(DH + cchunk * 8), vb);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float d3 = fmaxf(fmaf(__uint_as_float(va[q]) + __uint_as_float(vb[q]), inv3, b3v[cchunk * 8 + q]), 0.f);
                    sum = fmaf(d3, w4v[cchunk * 8 + q], sum);
                }
            }
            dpart[half * TS + slot] = sum;
        }
        ++nlayer;
        tc_fence_before();
        __syncthreads();
        for (int sl = tid; sl < cnt; sl += NT) {
            const float y = fxd::nan_to_num(dpart[sl] + dpart[TS + sl] + b4);
            float tot = (p.mem == 0) ? y : p.out[first + sl] + y;
            if (p.M > 1 && p.mem == p.M - 1) tot = tot / (float)p.M;
            p.out[first + sl] = tot;
        }
        __syncthreads();
    }
    if (xmax > 60000.f) atomicExch(p.overflow_flag, 1);
    tc_fence_before();
    __syncthreads();
    if (wid == 0) tmem_dealloc(tmem_base, 512);
}

struct Plan {
    int K1, nblk, off_w2, off_w3, off_vec, member_bytes;
    size_t smem;
};

static bool plan(const flexs_model *m, Plan &pl) {
    if (m->kind != FLEXS_KIND_MLP || m->H > DH) return false;
    pl.K1 = m->L * m->A;
    pl.nblk = (pl.K1 + 8 * KB_CHUNKS - 1) / (8 * KB_CHUNKS);
    pl.off_w2 = pl.nblk * BBLK;
    pl.off_w3 = pl.off_w2 + 14 * DBK;
    pl.off_vec = pl.off_w3 + 14 * DBK;
    pl.member_bytes = (pl.off_vec + MV_FLOATS * 4 + 255) / 256 * 256;
    pl.smem = (size_t)S_IDX + (size_t)TS * m->L + 1024;
    return (int64_t)pl.smem <= m->max_smem_optin;
}

// W1 (K1, H) -> planes [k chunk][n: 0..DH-1 hi | DH..2DH-1 lo][8 k] over nblk whole blocks (zero padded); no ASCALE
// descale: the one-hot input is exact
static float fill_w1(const float *src, int K1, int H, int nblk, unsigned char *dst, bool *finite) {
    const int e = prep::scale_exponent(src, (size_t)K1 * H, finite);
    const float scale = std::ldexp(1.f, e);
    __half *planes = reinterpret_cast<__half *>(dst);
    for (int k = 0; k < K1; ++k)
        for (int o = 0; o < H; ++o) {
            const float v = src[(size_t)k * H + o] * scale;
            const __half hi = __float2half_rn(v);
            const __half lo = __float2half_rn(v - __half2float(hi));
            const size_t base = (size_t)(k >> 3) * (DBK / 2) + (k & 7);
            planes[base + (size_t)o * 8] = hi;
            planes[base + (size_t)(DH + o) * 8] = lo;
        }
    (void)nblk;
    return std::ldexp(1.f, -e);
}

static int prepare(flexs_model *m, const Plan &pl) {
    if (m->mlp_ready) return FLEXS_OK;
    const fx::MlpOffsets o = fx::mlp_offsets(m);
    std::vector<float> host((size_t)m->member_floats * m->M);
    FX_CUDA(cudaSetDevice(m->device));
    FX_CUDA(cudaMemcpy(host.data(), m->d_weights, host.size() * sizeof(float), cudaMemcpyDeviceToHost));
    std::vector<unsigned char> blob((size_t)pl.member_bytes * m->M, 0);
    m->umma_weights_ok = true;
    const int H = m->H;
    for (int mem = 0; mem < m->M; ++mem) {
        const float *w = host.data() + (size_t)mem * m->member_floats;
        unsigned char *dst = blob.data() + (size_t)mem * pl.member_bytes;
        const float inv1 = fill_w1(w + o.w1, pl.K1, H, pl.nblk, dst, &m->umma_weights_ok);
        const float inv2 = prep::fill_dense_planes(w + o.w2, H, H, dst + pl.off_w2, &m->umma_weights_ok);
        const float inv3 = prep::fill_dense_planes(w + o.w3, H, H, dst + pl.off_w3, &m->umma_weights_ok);
        float *v = reinterpret_cast<float *>(dst + pl.off_vec);
        for (int i = 0; i < H; ++i) {
            v[i] = w[o.b1 + i];
            v[DH + i] = w[o.b2 + i] * ASCALE;   // the layer-2 epilogue emits activations pre-scaled by ASCALE
            v[2 * DH + i] = w[o.b3 + i];
            v[3 * DH + i] = w[o.w4 + i];
            if (!std::isfinite(w[o.b1 + i]) || !std::isfinite(w[o.b2 + i]) || !std::isfinite(w[o.b3 + i]) || !std::isfinite(w[o.w4 + i]))
                m->umma_weights_ok = false;
        }
        v[4 * DH] = inv1;               // layer 1: exact one-hot input, output scaled by ASCALE in the epilogue
        v[4 * DH + 1] = inv2 * ASCALE;  // layer 2 consumes ASCALE-scaled input and emits ASCALE-scaled output
        v[4 * DH + 2] = inv3;
        v[4 * DH + 3] = w[o.b4];
    }
    if (!m->d_mlp_w) FX_CUDA(cudaMalloc(&m->d_mlp_w, blob.size()));
    FX_CUDA(cudaMemcpy(m->d_mlp_w, blob.data(), blob.size(), cudaMemcpyHostToDevice));
    m->mlp_ready = true;
    return FLEXS_OK;
}

}  // namespace

namespace fx {

bool mlp_umma_supported(const flexs_model *m) {
    Plan pl;
    return plan(m, pl);
}

int launch_mlp_umma(flexs_model *m, const uint8_t *d_idx, int64_t n, float *d_out, cudaStream_t s) {
    Plan pl;
    FX_REQUIRE(plan(m, pl), "shape not supported by the tcgen05 MLP kernel (H <= 112, 128 * L bytes of residues in shared memory)");
    int rc = prepare(m, pl);
    if (rc != FLEXS_OK) return rc;
    if (!m->umma_weights_ok) return launch_mlp(m, d_idx, n, d_out, s);   // non-finite weights: fp32 path
    flexs_model::StreamWs *ws = nullptr;
    rc = stream_workspace(m, s, 0, &ws);
    if (rc != FLEXS_OK) return rc;
    FX_CUDA(cudaMemsetAsync(ws->flag, 0, sizeof(int), s));
    MlpUParams p;
    p.idx = d_idx; p.out = d_out; p.overflow_flag = ws->flag; p.n = n; p.n_tiles = (n + TS - 1) / TS;
    p.L = m->L; p.A = m->A; p.H = m->H; p.K1 = pl.K1; p.nblk = pl.nblk; p.M = m->M;
    p.off_w2 = pl.off_w2; p.off_w3 = pl.off_w3; p.off_vec = pl.off_vec;
    FX_CUDA(cudaFuncSetAttribute(mlp_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    const int grid = (int)std::min<int64_t>(p.n_tiles, m->sm_count);
    for (int mem = 0; mem < m->M; ++mem) {
        p.mem = mem;
        p.uw = reinterpret_cast<const unsigned char *>(m->d_mlp_w) + (size_t)mem * pl.member_bytes;
        mlp_umma_kernel<<<grid, NT, pl.smem, s>>>(p);
        FX_CUDA(cudaGetLastError());
        m->launches += 1;
    }
    // fp16 range guard: the gated FFMA kernel recomputes the batch iff the flag was raised
    return launch_mlp_gated(m, d_idx, n, d_out, ws->flag, s);
}

}  // namespace fx
