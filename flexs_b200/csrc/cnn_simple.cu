// K1 (shape-generic variant): one CTA per sequence, every layer of cnn.py:23-54 evaluated with
// activations in shared memory.  Exists so that EVERY shape the reference constructor accepts
// (e.g. tests/test_models.py:56-63: L=3, F=1, H=1, k=2) runs on the GPU; the F=32 shapes the
// paper scripts use go through cnn_tiled.cu / cnn_umma.cu instead.
#include <algorithm>

#include "common.cuh"

namespace {

struct SimpleParams {
    const uint8_t *idx;
    float *out;
    const float *weights;  // [M][member_floats]
    int64_t n;
    int64_t member_floats;
    fx::CnnDims d;
    fx::CnnOffsets o;
    int M;
};

__global__ void __launch_bounds__(256) cnn_simple_kernel(SimpleParams p) {
    extern __shared__ __align__(16) float smem[];
    const fx::CnnDims d = p.d;
    const int T = d.T, F = d.F, H = d.H;
    float *h1 = smem;                  // [T][F]
    float *h2 = h1 + (size_t)T * F;    // [T][F]
    float *feat = h2 + (size_t)T * F;  // [F]
    float *d1 = feat + F;              // [H]
    float *d2 = d1 + H;                // [H]
    uint8_t *sidx = reinterpret_cast<uint8_t *>(d2 + H);  // [L]
    const int tid = threadIdx.x, nt = blockDim.x;

    for (int64_t seq = blockIdx.x; seq < p.n; seq += gridDim.x) {
        __syncthreads();
        for (int i = tid; i < d.L; i += nt) sidx[i] = p.idx[seq * d.L + i];
        __syncthreads();
        float total = 0.f;
        for (int mem = 0; mem < p.M; ++mem) {
            const float *w = p.weights + (int64_t)mem * p.member_floats;
            const float *w1 = w + p.o.w1, *b1 = w + p.o.b1, *w2 = w + p.o.w2, *b2 = w + p.o.b2;
            const float *w3 = w + p.o.w3, *b3 = w + p.o.b3, *wd1 = w + p.o.wd1, *bd1 = w + p.o.bd1;
            const float *wd2 = w + p.o.wd2, *bd2 = w + p.o.bd2, *wd3 = w + p.o.wd3, *bd3 = w + p.o.bd3;
            // conv1 (valid): the one-hot input selects one (A,F) row of W1 per tap
            for (int e = tid; e < T * F; e += nt) {
                const int t = e / F, f = e - t * F;
                float acc = 0.f;
                for (int j = 0; j < d.K; ++j) acc += __ldg(w1 + ((size_t)j * d.A + sidx[t + j]) * F + f);
                h1[e] = fmaxf(acc + __ldg(b1 + f), 0.f);
            }
            __syncthreads();
            // conv2 (same)
            for (int e = tid; e < T * F; e += nt) {
                const int t = e / F, f = e - t * F;
                float acc = 0.f;
                for (int j = 0; j < d.K; ++j) {
                    const int s = t + j - d.pl2;
                    if (s < 0 || s >= T) continue;
                    const float *x = h1 + (size_t)s * F;
                    const float *wj = w2 + (size_t)j * F * F + f;
                    for (int g = 0; g < F; ++g) acc = fmaf(x[g], __ldg(wj + (size_t)g * F), acc);
                }
                h2[e] = fmaxf(acc + __ldg(b2 + f), 0.f);
            }
            __syncthreads();
            // conv3 (same, width A-1) -> written over h1, then global max over time
            for (int e = tid; e < T * F; e += nt) {
                const int t = e / F, f = e - t * F;
                float acc = 0.f;
                for (int j = 0; j < d.K3; ++j) {
                    const int s = t + j - d.pl3;
                    if (s < 0 || s >= T) continue;
                    const float *x = h2 + (size_t)s * F;
                    const float *wj = w3 + (size_t)j * F * F + f;
                    for (int g = 0; g < F; ++g) acc = fmaf(x[g], __ldg(wj + (size_t)g * F), acc);
                }
                h1[e] = fmaxf(acc + __ldg(b3 + f), 0.f);
            }
            __syncthreads();
            for (int f = tid; f < F; f += nt) {
                float mx = h1[f];
                for (int t = 1; t < T; ++t) mx = fmaxf(mx, h1[(size_t)t * F + f]);
                feat[f] = mx;
            }
            __syncthreads();
            for (int o = tid; o < H; o += nt) {
                float acc = 0.f;
                for (int g = 0; g < F; ++g) acc = fmaf(feat[g], __ldg(wd1 + (size_t)g * H + o), acc);
                d1[o] = fmaxf(acc + __ldg(bd1 + o), 0.f);
            }
            __syncthreads();
            for (int o = tid; o < H; o += nt) {
                float acc = 0.f;
                for (int g = 0; g < H; ++g) acc = fmaf(d1[g], __ldg(wd2 + (size_t)g * H + o), acc);
                d2[o] = fmaxf(acc + __ldg(bd2 + o), 0.f);
            }
            __syncthreads();
            if (tid < 32) {
                float acc = 0.f;
                for (int g = tid; g < H; g += 32) acc = fmaf(d2[g], __ldg(wd3 + g), acc);
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
                total += fxd::nan_to_num(acc + __ldg(bd3));
            }
            __syncthreads();
        }
        // Ensemble default combine (ensemble.py:24): np.mean over members, fp32, divide by M
        if (tid == 0) p.out[seq] = (p.M == 1) ? total : total / (float)p.M;
    }
}

}  // namespace

namespace fx {

int launch_cnn_simple(flexs_model *m, const uint8_t *d_idx, int64_t n, float *d_out, cudaStream_t s) {
    SimpleParams p;
    p.idx = d_idx; p.out = d_out; p.weights = m->d_weights; p.n = n;
    p.member_floats = m->member_floats; p.d = cnn_dims(m); p.o = cnn_offsets(m); p.M = m->M;
    size_t smem = sizeof(float) * ((size_t)2 * p.d.T * p.d.F + p.d.F + 2 * (size_t)p.d.H) + p.d.L + 16;
    FX_REQUIRE((int64_t)smem <= m->max_smem_optin,
               "sequence too long for the shape-generic CNN kernel (activations must fit shared memory)");
    FX_CUDA(cudaFuncSetAttribute(cnn_simple_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t grid = std::min<int64_t>(n, (int64_t)m->sm_count * 8);
    cnn_simple_kernel<<<(unsigned)grid, 256, smem, s>>>(p);
    FX_CUDA(cudaGetLastError());
    m->launches += 1;
    return FLEXS_OK;
}

}  // namespace fx
