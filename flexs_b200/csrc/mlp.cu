// K2: fused MLP forward (mlp.py:21-31): Flatten -> Dense(H,relu) x3 -> Dense(1), plus the
// squeeze/nan_to_num of keras_model.py:77-79 and the Ensemble mean (ensemble.py:54-59).
// Layer 1 on a one-hot input is a gather: out[o] = b1[o] + sum_l W1[l*A + idx[l]][o]; the
// float one-hot and the (L*A)-wide GEMM are never formed.  A CTA scores 64 sequences at a time,
// activations stay in shared memory as [channel][slot]; each thread owns one output channel for
// 8 sequences so every weight it loads (coalesced across the warp, L1/L2 resident) is used 8x.
#include <algorithm>

#include "common.cuh"

namespace {

constexpr int NT = 256;
constexpr int SB = 64;   // sequences per CTA batch
constexpr int SBP = 68;  // slot pitch (floats): 16B-aligned rows, conflict-free float4 stores

struct MlpParams {
    const uint8_t *idx;
    float *out;
    const float *weights;
    int64_t n, n_batches, member_floats;
    fx::MlpOffsets o;
    int L, A, H, M;
    const int *gate;  // non-null: every CTA returns at once unless *gate != 0 (fp16-overflow fall-back of mlp_umma.cu)
};

__device__ __forceinline__ void dense_relu(const float *__restrict__ w, const float *__restrict__ b,
                                           const float *__restrict__ xT, float *__restrict__ yT,
                                           int in, int H, int nsg) {
    const int HP = (H + 31) & ~31;
    for (int wk = threadIdx.x; wk < HP * nsg; wk += NT) {
        const int o = wk % HP, sg = wk / HP;
        if (o >= H) continue;
        float acc[8];
        const float bb = __ldg(b + o);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = bb;
#pragma unroll 4
        for (int g = 0; g < in; ++g) {
            const float wv = __ldg(w + (size_t)g * H + o);
            const float4 x0 = *reinterpret_cast<const float4 *>(xT + (size_t)g * SBP + sg * 8);
            const float4 x1 = *reinterpret_cast<const float4 *>(xT + (size_t)g * SBP + sg * 8 + 4);
            acc[0] = fmaf(wv, x0.x, acc[0]); acc[1] = fmaf(wv, x0.y, acc[1]);
            acc[2] = fmaf(wv, x0.z, acc[2]); acc[3] = fmaf(wv, x0.w, acc[3]);
            acc[4] = fmaf(wv, x1.x, acc[4]); acc[5] = fmaf(wv, x1.y, acc[5]);
            acc[6] = fmaf(wv, x1.z, acc[6]); acc[7] = fmaf(wv, x1.w, acc[7]);
        }
        *reinterpret_cast<float4 *>(yT + (size_t)o * SBP + sg * 8) =
            make_float4(fmaxf(acc[0], 0.f), fmaxf(acc[1], 0.f), fmaxf(acc[2], 0.f), fmaxf(acc[3], 0.f));
        *reinterpret_cast<float4 *>(yT + (size_t)o * SBP + sg * 8 + 4) =
            make_float4(fmaxf(acc[4], 0.f), fmaxf(acc[5], 0.f), fmaxf(acc[6], 0.f), fmaxf(acc[7], 0.f));
    }
}

__global__ void __launch_bounds__(NT) mlp_kernel(const MlpParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (p.gate != nullptr && __ldg(p.gate) == 0) return;
    const int L = p.L, A = p.A, H = p.H;
    float *xa = reinterpret_cast<float *>(smem_raw);  // [H][SBP]
    float *xb = xa + (size_t)H * SBP;                  // [H][SBP]
    uint8_t *sidx = reinterpret_cast<uint8_t *>(xb + (size_t)H * SBP);  // [SB][L]
    const int tid = threadIdx.x;
    const int HP = (H + 31) & ~31;

    for (int64_t batch = blockIdx.x; batch < p.n_batches; batch += gridDim.x) {
        const int64_t first = batch * SB;
        const int cnt = (int)min((int64_t)SB, p.n - first);
        const int nsg = (cnt + 7) >> 3;
        __syncthreads();
        for (int i = tid; i < SB * L; i += NT) sidx[i] = (i < cnt * L) ? p.idx[first * L + i] : 0;
        for (int mem = 0; mem < p.M; ++mem) {
            const float *w = p.weights + (int64_t)mem * p.member_floats;
            __syncthreads();
            // layer 1: gather-add of W1 rows
            for (int wk = tid; wk < HP * nsg; wk += NT) {
                const int o = wk % HP, sg = wk / HP;
                if (o >= H) continue;
                float acc[8];
                const float bb = __ldg(w + p.o.b1 + o);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = bb;
                const uint8_t *ip = sidx + (size_t)sg * 8 * L;
                for (int l = 0; l < L; ++l) {
                    const float *wl = w + p.o.w1 + (size_t)l * A * H + o;
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[i] += __ldg(wl + (size_t)ip[i * L + l] * H);
                }
                *reinterpret_cast<float4 *>(xa + (size_t)o * SBP + sg * 8) =
                    make_float4(fmaxf(acc[0], 0.f), fmaxf(acc[1], 0.f), fmaxf(acc[2], 0.f), fmaxf(acc[3], 0.f));
                *reinterpret_cast<float4 *>(xa + (size_t)o * SBP + sg * 8 + 4) =
                    make_float4(fmaxf(acc[4], 0.f), fmaxf(acc[5], 0.f), fmaxf(acc[6], 0.f), fmaxf(acc[7], 0.f));
            }
            __syncthreads();
            dense_relu(w + p.o.w2, w + p.o.b2, xa, xb, H, H, nsg);
            __syncthreads();
            dense_relu(w + p.o.w3, w + p.o.b3, xb, xa, H, H, nsg);
            __syncthreads();
            for (int slot = tid; slot < cnt; slot += NT) {
                float acc = 0.f;
#pragma unroll 4
                for (int g = 0; g < H; ++g) acc = fmaf(xa[(size_t)g * SBP + slot], __ldg(w + p.o.w4 + g), acc);
                const float y = fxd::nan_to_num(acc + __ldg(w + p.o.b4));
                float tot = (mem == 0) ? y : p.out[first + slot] + y;
                if (p.M > 1 && mem == p.M - 1) tot = tot / (float)p.M;
                p.out[first + slot] = tot;
            }
        }
    }
}

}  // namespace

namespace fx {

int launch_mlp(flexs_model *m, const uint8_t *d_idx, int64_t n, float *d_out, cudaStream_t s) {
    return launch_mlp_gated(m, d_idx, n, d_out, nullptr, s);
}

int launch_mlp_gated(flexs_model *m, const uint8_t *d_idx, int64_t n, float *d_out, const int *d_gate, cudaStream_t s) {
    MlpParams p;
    p.gate = d_gate;
    p.idx = d_idx; p.out = d_out; p.weights = m->d_weights; p.n = n;
    p.n_batches = (n + SB - 1) / SB;
    p.member_floats = m->member_floats; p.o = mlp_offsets(m);
    p.L = m->L; p.A = m->A; p.H = m->H; p.M = m->M;
    const size_t smem = (size_t)2 * m->H * SBP * 4 + (size_t)SB * m->L + 16;
    FX_REQUIRE((int64_t)smem <= m->max_smem_optin, "MLP hidden size / sequence length too large for shared memory");
    FX_CUDA(cudaFuncSetAttribute(mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    FX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, mlp_kernel, NT, smem));
    const int64_t grid = std::max<int64_t>(1, std::min<int64_t>(p.n_batches, (int64_t)m->sm_count * std::max(1, occ)));
    mlp_kernel<<<(unsigned)grid, NT, smem, s>>>(p);
    FX_CUDA(cudaGetLastError());
    m->launches += 1;
    return FLEXS_OK;
}

}  // namespace fx
