// K9: the VAE generator of CbAS / DbAS — train, decode, reconstruction log-probability.
//
// Replaces flexs/utils/VAE_utils.py:28-217: VAEModel (encoder Dense-ELU -> Dropout(0.3) -> Dense-ELU -> BatchNorm ->
// Dense-ELU -> z_mean / z_log_var -> sampling; decoder Dense-ELU x2 -> Dropout(0.3) -> Dense-ELU -> Dense-sigmoid), its
// train_step (loss = sum_d BCE + KL, :75-92), the compile() of :127 (Adam lr 1e-4, clipvalue 0.5) as driven by fit (:141-151),
// generate()'s decoder pass (:71-74, :158-160) and calculate_log_probability (:189-217).
//
// CbAS refits the generator after every 100 proposals on a data set of a few hundred to a few thousand sequences in
// mini-batches of TEN (cbas_dbas.py:183, VAE_utils.py:104): thousands of optimiser steps of almost no arithmetic each.
// This is a latency problem: everything a step needs lives on the device (weights, Adam moments, the epoch's permutation,
// Philox for the dropout masks and the latent noise), a step is ~45 small launches on one stream with no host
// synchronisation, and the host only reads the epoch's mean loss back (early stopping, :139).  The input is one-hot, so
// the first layer is a row gather and its weight gradient a scatter by residue; nothing else is special.  Deterministic
// (no atomics: every reduction has one owner thread or a fixed-shape tree).
//
// Semantics that are choices (TensorFlow is absent, the reference's tests assert no values — "parity unpinned" like the
// surrogates'): sample weights multiply the per-sample loss, the batch loss is their sum / batch size; BatchNorm uses
// the biased batch variance for normalisation and for the moving average (Keras non-fused path, momentum 0.99, eps 1e-3).
// The test suite states the same step in float64 (torch autograd on the CPU) and pins one step against it.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "common.cuh"

struct flexs_vae {
    int device = 0, L = 0, A = 0, D = 0, I = 0, Z = 0;
    std::vector<int64_t> sizes, offs;   // 22 arrays, Keras get_weights() order
    int64_t total = 0;
    float *w = nullptr, *g = nullptr, *m = nullptr, *v = nullptr;   // weights, gradients, Adam moments
    int64_t step = 0;
    float *ws = nullptr;                // activations of one batch
    int64_t ws_floats = 0;
    int maxB = 0;
    long long *d_state = nullptr;       // fit(): [0] optimiser steps done, [1] offset into the permutation array (device-side
                                        // so that one captured CUDA graph replays for every mini-batch)
    cudaStream_t fit_stream = nullptr;  // graphs cannot be captured on the legacy default stream
};

namespace {

constexpr int NARR = 22;
enum { W1, B1, W2, B2, BNG, BNB, BNM, BNV, W3, B3, WM, BM, WV, BV, W4, B4, W5, B5, W6, B6, W7, B7 };
constexpr float DROP = 0.3f, BN_EPS = 1e-3f, BN_MOM = 0.99f;
constexpr float ADAM_LR = 1e-4f, ADAM_B1 = 0.9f, ADAM_B2 = 0.999f, ADAM_EPS = 1e-7f, CLIP = 0.5f;

__device__ __forceinline__ void philox(uint64_t counter, uint64_t key, uint32_t (&out)[4]) {
    uint32_t c[4] = {(uint32_t)counter, (uint32_t)(counter >> 32), 0x243F6A88u, 0x85A308D3u};
    uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

// dropout keep masks (value 1/(1-p) or 0) and standard-normal latent noise for one step
__global__ void k_noise(float *mask1, float *mask2, float *eps, int nmask, int neps, uint64_t seed, uint64_t step,
                        const long long *d_state) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t r[4];
    if (d_state) step = (uint64_t)d_state[0] + 1;
    if (i < nmask) {
        philox((uint64_t)i, seed ^ (step * 0x9E3779B97F4A7C15ull), r);
        mask1[i] = ((r[0] >> 8) * (1.0f / 16777216.0f) >= DROP) ? 1.f / (1.f - DROP) : 0.f;
        mask2[i] = ((r[1] >> 8) * (1.0f / 16777216.0f) >= DROP) ? 1.f / (1.f - DROP) : 0.f;
    }
    if (i < neps) {
        philox((uint64_t)i + (1ull << 40), seed ^ (step * 0x9E3779B97F4A7C15ull), r);
        const float u1 = ((r[0] >> 8) + 1) * (1.0f / 16777216.0f), u2 = (r[1] >> 8) * (1.0f / 16777216.0f);
        eps[i] = sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);   // Box-Muller
    }
}

__device__ __forceinline__ float act_fwd(float z, int act) {
    if (act == 1) return z > 0.f ? z : expm1f(z);           // ELU, alpha = 1
    if (act == 2) return 1.f / (1.f + expf(-z));            // sigmoid
    return z;
}

// layer 1 on the one-hot input: y[b,o] = act(b1[o] + sum_l W1[l*A + idx[row(b), l], o]) * mask
__global__ void k_gather_fwd(const uint8_t *idx, const int *perm, const long long *d_state, const float *w, const float *bias,
                             const float *mask, float *y, int B, int L, int A, int out, int act) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * out) return;
    const int b = i / out, o = i - b * out;
    if (perm && d_state) perm += d_state[1];
    const uint8_t *row = idx + (size_t)(perm ? perm[b] : b) * L;
    float acc = bias[o];
    for (int l = 0; l < L; ++l) acc += w[(size_t)(l * A + row[l]) * out + o];
    float v = act_fwd(acc, act);
    if (mask) v *= mask[i];
    y[i] = v;
}

// y[b,o] = act(bias[o] + sum_i x[b,i] W[i,o]) * mask
__global__ void k_dense_fwd(const float *x, const float *w, const float *bias, const float *mask, float *y, int B, int in, int out,
                            int act) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * out) return;
    const int b = i / out, o = i - b * out;
    const float *xr = x + (size_t)b * in;
    float acc = bias[o];
    for (int k = 0; k < in; ++k) acc = fmaf(xr[k], w[(size_t)k * out + o], acc);
    float v = act_fwd(acc, act);
    if (mask) v *= mask[i];
    y[i] = v;
}

// BatchNormalization over the batch axis, one thread per feature.  train: batch statistics (biased variance), moving
// averages updated; else: moving statistics.  xhat and 1/std are kept for the backward pass.
__global__ void k_bn_fwd(const float *x, const float *gamma, const float *beta, float *mov_mean, float *mov_var, float *xhat,
                         float *inv_std, float *y, int B, int I, int train) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= I) return;
    float mean, var;
    if (train) {
        float s = 0.f;
        for (int b = 0; b < B; ++b) s += x[(size_t)b * I + f];
        mean = s / B;
        float q = 0.f;
        for (int b = 0; b < B; ++b) { const float d = x[(size_t)b * I + f] - mean; q = fmaf(d, d, q); }
        var = q / B;
        mov_mean[f] = BN_MOM * mov_mean[f] + (1.f - BN_MOM) * mean;
        mov_var[f] = BN_MOM * mov_var[f] + (1.f - BN_MOM) * var;
    } else {
        mean = mov_mean[f]; var = mov_var[f];
    }
    const float is = rsqrtf(var + BN_EPS);
    if (inv_std) inv_std[f] = is;
    for (int b = 0; b < B; ++b) {
        const float xh = (x[(size_t)b * I + f] - mean) * is;
        if (xhat) xhat[(size_t)b * I + f] = xh;
        y[(size_t)b * I + f] = fmaf(xh, gamma[f], beta[f]);
    }
}

__global__ void k_sample(const float *zm, const float *zlv, const float *eps, float *z, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) z[i] = zm[i] + expf(0.5f * zlv[i]) * (eps ? eps[i] : 0.f);
}

// Loss of VAEModel.train_step (:78-86) per sample, and its gradients at the network's outputs:
//   recon_b = sum_d BCE(x_bd, out_bd)  (= original_dim * mean_d, keras clips probabilities to [1e-7, 1 - 1e-7])
//   kl_b    = -0.5 * mean_z(1 + lv - m^2 - exp(lv));   loss = sum_b weight_b (recon_b + kl_b) / B
// gpre7[b,d] = weight_b / B * (out - x) (sigmoid + BCE; zero where the clip is active), gzm / gzlv receive the KL part.
__global__ void k_loss(const float *out, const uint8_t *idx, const int *perm, const long long *d_state, const float *weights,
                       const float *zm, const float *zlv, float *gpre7, float *gzm, float *gzlv, double *loss_accum, int B, int L,
                       int A, int Z) {
    __shared__ double s_part[256];
    const int D = L * A;
    if (perm && d_state) perm += d_state[1];
    double acc = 0.0;
    for (int i = threadIdx.x; i < B * D; i += blockDim.x) {
        const int b = i / D, d = i - b * D;
        const int row = perm ? perm[b] : b;
        const float x = (idx[(size_t)row * L + d / A] == d % A) ? 1.f : 0.f;
        const float wb = weights[row] / B;
        const float o = out[i], oc = fminf(fmaxf(o, 1e-7f), 1.f - 1e-7f);
        acc += (double)(weights[row] * -(x * logf(oc) + (1.f - x) * logf(1.f - oc)));
        gpre7[i] = (o > 1e-7f && o < 1.f - 1e-7f) ? wb * (o - x) : 0.f;
    }
    for (int i = threadIdx.x; i < B * Z; i += blockDim.x) {
        const int b = i / Z;
        const int row = perm ? perm[b] : b;
        const float m = zm[i], lv = zlv[i], wb = weights[row] / B;
        acc += (double)(weights[row] * (-0.5f / Z) * (1.f + lv - m * m - expf(lv)));
        gzm[i] = wb * m / Z;
        gzlv[i] = wb * (-0.5f / Z) * (1.f - expf(lv));
    }
    s_part[threadIdx.x] = acc;
    __syncthreads();
    for (int off = blockDim.x / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) s_part[threadIdx.x] += s_part[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) *loss_accum += s_part[0];   // sum over the batch; the host divides by the samples seen
}

// gW[i,o] = sum_b x[b,i] gy[b,o] ; gb[o] = sum_b gy[b,o]
__global__ void k_dense_bwd_w(const float *x, const float *gy, float *gw, float *gb, int B, int in, int out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < in * out) {
        const int k = i / out, o = i - k * out;
        float acc = 0.f;
        for (int b = 0; b < B; ++b) acc = fmaf(x[(size_t)b * in + k], gy[(size_t)b * out + o], acc);
        gw[i] = acc;
    } else if (i < in * out + out) {
        const int o = i - in * out;
        float acc = 0.f;
        for (int b = 0; b < B; ++b) acc += gy[(size_t)b * out + o];
        gb[o] = acc;
    }
}

// layer-1 weight gradient: gW1[l*A + a, o] = sum_{b : idx[row(b), l] == a} gy[b,o]
__global__ void k_gather_bwd_w(const uint8_t *idx, const int *perm, const long long *d_state, const float *gy, float *gw, float *gb,
                               int B, int L, int A, int out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int D = L * A;
    if (perm && d_state) perm += d_state[1];
    if (i < D * out) {
        const int k = i / out, o = i - k * out, l = k / A, a = k - l * A;
        float acc = 0.f;
        for (int b = 0; b < B; ++b)
            if (idx[(size_t)(perm ? perm[b] : b) * L + l] == a) acc += gy[(size_t)b * out + o];
        gw[i] = acc;
    } else if (i < D * out + out) {
        const int o = i - D * out;
        float acc = 0.f;
        for (int b = 0; b < B; ++b) acc += gy[(size_t)b * out + o];
        gb[o] = acc;
    }
}

// gx[b,i] = sum_o gy[b,o] W[i,o], then through what produced x: * mask, * act'(.) with act' from the OUTPUT value
// (ELU: y > 0 ? 1 : y + 1, y before the mask)
__global__ void k_dense_bwd_x(const float *w, const float *gy, const float *mask, const float *y_unmasked, float *gx, int B, int in,
                              int out, int act_in) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * in) return;
    const int b = i / in, k = i - b * in;
    const float *wr = w + (size_t)k * out, *gr = gy + (size_t)b * out;
    float acc = 0.f;
    for (int o = 0; o < out; ++o) acc = fmaf(gr[o], wr[o], acc);
    if (mask) acc *= mask[i];
    if (act_in == 1) { const float y = y_unmasked[i]; acc *= (y > 0.f ? 1.f : y + 1.f); }
    gx[i] = acc;
}

// BatchNorm backward, one thread per feature: gy -> gx (in place into gx), ggamma, gbeta; then through the ELU before it
__global__ void k_bn_bwd(const float *gy, const float *xhat, const float *inv_std, const float *gamma, const float *y_prev,
                         float *gx, float *ggamma, float *gbeta, int B, int I) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= I) return;
    float sg = 0.f, sgx = 0.f;
    for (int b = 0; b < B; ++b) { const float g = gy[(size_t)b * I + f]; sg += g; sgx = fmaf(g, xhat[(size_t)b * I + f], sgx); }
    ggamma[f] = sgx; gbeta[f] = sg;
    const float k = gamma[f] * inv_std[f] / B;
    for (int b = 0; b < B; ++b) {
        const size_t i = (size_t)b * I + f;
        float v = k * (B * gy[i] - sg - xhat[i] * sgx);
        const float y = y_prev[i];                       // the ELU output that fed the BatchNorm
        v *= (y > 0.f ? 1.f : y + 1.f);
        gx[i] = v;
    }
}

// reparameterisation: z = m + exp(lv / 2) eps  ->  gm += gz ; glv += gz * 0.5 * exp(lv / 2) * eps
__global__ void k_sample_bwd(const float *gz, const float *zlv, const float *eps, float *gzm, float *gzlv, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    gzm[i] += gz[i];
    gzlv[i] += gz[i] * 0.5f * expf(0.5f * zlv[i]) * eps[i];
}

// gx = gx_a + gx_b (the two heads share their input)
__global__ void k_add(const float *a, const float *b, float *out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] + b[i];
}

// Keras Adam with clipvalue: g clipped to [-0.5, 0.5] elementwise; lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t)
__global__ void k_adam_clip(float *w, const float *g, float *m, float *v, int64_t count, float lr_t, const long long *d_state,
                            int64_t skip_a, int64_t skip_b) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count || (i >= skip_a && i < skip_b)) return;   // the BatchNorm moving statistics are not trained
    if (d_state) {
        const double t = (double)(d_state[0] + 1);
        lr_t = (float)((double)ADAM_LR * sqrt(1.0 - pow((double)ADAM_B2, t)) / (1.0 - pow((double)ADAM_B1, t)));
    }
    const float gi = fminf(fmaxf(g[i], -CLIP), CLIP);
    const float mi = ADAM_B1 * m[i] + (1.f - ADAM_B1) * gi;
    const float vi = ADAM_B2 * v[i] + (1.f - ADAM_B2) * gi * gi;
    m[i] = mi; v[i] = vi;
    w[i] -= lr_t * mi / (sqrtf(vi) + ADAM_EPS);
}

// log-probability of reconstructing each sequence (:189-217): sum_l log(1e-9 + out[l, x_l] / sum_a out[l, a]), nan_to_num
__global__ void k_log_prob(const float *out, const uint8_t *idx, double *lp, int n, int L, int A) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    double acc = 0.0;
    for (int l = 0; l < L; ++l) {
        const float *p = out + ((size_t)b * L + l) * A;
        float s = 0.f;
        for (int a = 0; a < A; ++a) s += p[a];
        acc += log(1e-9 + (double)(p[idx[(size_t)b * L + l]] / s));
    }
    if (acc != acc) acc = 0.0;
    if (acc > 1.7976931348623157e308) acc = 1.7976931348623157e308;
    if (acc < -1.7976931348623157e308) acc = -1.7976931348623157e308;
    lp[b] = acc;
}

// end of a graph-replayed step: one more optimiser step done, the next mini-batch starts B rows further into the permutation
__global__ void k_advance(long long *d_state, int B, int steps) {
    d_state[0] += steps;
    d_state[1] += B;
}

__global__ void k_mul(const float *a, const float *b, float *out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] * b[i];
}

inline int blocks(int64_t n) { return (int)std::max<int64_t>(1, (n + 255) / 256); }

// activations and gradient scratch of one batch of B rows
struct Ws {
    double *loss;
    float *h1, *h1m, *h2, *xhat, *istd, *bn, *h3, *zm, *zlv, *z, *g1, *g2, *g2m, *g3, *out, *mask1, *mask2, *eps;
    float *ga, *gb, *gzm, *gzlv, *gz;
};

int64_t ws_floats_for(const flexs_vae *v, int B) {
    const int64_t I = v->I, D = v->D, Z = v->Z, W = std::max(I, D);
    // 12 [B, I] arrays, out [B, D], two gradient buffers [B, max(I, D)], 7 [B, Z] arrays, istd [I], the loss; every array is
    // rounded up to 4 floats by carve()
    return 64 + (int64_t)B * (12 * I + D + 2 * W + 7 * Z) + I + 32 * 4;
}

Ws carve(const flexs_vae *v, int B) {
    Ws w;
    float *p = v->ws;
    const int64_t I = v->I, D = v->D, Z = v->Z, W = std::max(I, D);
    auto take = [&](int64_t n) { float *r = p; p += (n + 3) / 4 * 4; return r; };
    w.loss = reinterpret_cast<double *>(take(4));
    w.h1 = take(B * I); w.h1m = take(B * I); w.h2 = take(B * I); w.xhat = take(B * I); w.istd = take(I); w.bn = take(B * I);
    w.h3 = take(B * I); w.zm = take(B * Z); w.zlv = take(B * Z); w.z = take(B * Z);
    w.g1 = take(B * I); w.g2 = take(B * I); w.g2m = take(B * I); w.g3 = take(B * I); w.out = take(B * D);
    w.mask1 = take(B * I); w.mask2 = take(B * I); w.eps = take(B * Z);
    w.ga = take(B * W); w.gb = take(B * W); w.gzm = take(B * Z); w.gzlv = take(B * Z); w.gz = take(B * Z);
    return w;
}

int ensure_ws(flexs_vae *v, int B) {
    const int64_t need = ws_floats_for(v, B);
    if (need <= v->ws_floats) return FLEXS_OK;
    if (v->ws) { FX_CUDA(cudaDeviceSynchronize()); FX_CUDA(cudaFree(v->ws)); v->ws = nullptr; v->ws_floats = 0; }
    FX_CUDA(cudaMalloc(&v->ws, need * sizeof(float)));
    v->ws_floats = need;
    return FLEXS_OK;
}

#define AT(a) (v->w + v->offs[a])
#define GR(a) (v->g + v->offs[a])

// decoder for B rows of z: Dense-ELU x2 -> Dropout -> Dense-ELU -> Dense-sigmoid
void decode(flexs_vae *v, const Ws &w, const float *z, int B, bool train, cudaStream_t s) {
    const int I = v->I, D = v->D, Z = v->Z;
    k_dense_fwd<<<blocks((int64_t)B * I), 256, 0, s>>>(z, AT(W4), AT(B4), nullptr, w.g1, B, Z, I, 1);
    k_dense_fwd<<<blocks((int64_t)B * I), 256, 0, s>>>(w.g1, AT(W5), AT(B5), nullptr, w.g2, B, I, I, 1);
    const float *x6 = w.g2;
    if (train) { k_mul<<<blocks((int64_t)B * I), 256, 0, s>>>(w.g2, w.mask2, w.g2m, B * I); x6 = w.g2m; }
    k_dense_fwd<<<blocks((int64_t)B * I), 256, 0, s>>>(x6, AT(W6), AT(B6), nullptr, w.g3, B, I, I, 1);
    k_dense_fwd<<<blocks((int64_t)B * D), 256, 0, s>>>(w.g3, AT(W7), AT(B7), nullptr, w.out, B, I, D, 2);
}

// encoder + sampling + decoder for B rows.  train: dropout masks and BatchNorm batch statistics; eps == nullptr: z = z_mean
void forward(flexs_vae *v, const Ws &w, const uint8_t *idx, const int *perm, int B, bool train, const float *eps, cudaStream_t s,
             const long long *d_state = nullptr) {
    const int I = v->I, Z = v->Z, L = v->L, A = v->A;
    k_gather_fwd<<<blocks((int64_t)B * I), 256, 0, s>>>(idx, perm, d_state, AT(W1), AT(B1), nullptr, w.h1, B, L, A, I, 1);
    const float *x2 = w.h1;
    if (train) { k_mul<<<blocks((int64_t)B * I), 256, 0, s>>>(w.h1, w.mask1, w.h1m, B * I); x2 = w.h1m; }
    k_dense_fwd<<<blocks((int64_t)B * I), 256, 0, s>>>(x2, AT(W2), AT(B2), nullptr, w.h2, B, I, I, 1);
    k_bn_fwd<<<blocks(I), 256, 0, s>>>(w.h2, AT(BNG), AT(BNB), AT(BNM), AT(BNV), w.xhat, w.istd, w.bn, B, I, train ? 1 : 0);
    k_dense_fwd<<<blocks((int64_t)B * I), 256, 0, s>>>(w.bn, AT(W3), AT(B3), nullptr, w.h3, B, I, I, 1);
    k_dense_fwd<<<blocks((int64_t)B * Z), 256, 0, s>>>(w.h3, AT(WM), AT(BM), nullptr, w.zm, B, I, Z, 0);
    k_dense_fwd<<<blocks((int64_t)B * Z), 256, 0, s>>>(w.h3, AT(WV), AT(BV), nullptr, w.zlv, B, I, Z, 0);
    k_sample<<<blocks((int64_t)B * Z), 256, 0, s>>>(w.zm, w.zlv, eps, w.z, B * Z);
    decode(v, w, w.z, B, train, s);
}

// gradients of the loss for the batch that forward(train) just ran; adds sum_b weight_b * loss_b to *w.loss
void backward(flexs_vae *v, const Ws &w, const uint8_t *idx, const int *perm, const float *weights, int B, cudaStream_t s,
              const long long *d_state = nullptr) {
    const int I = v->I, D = v->D, Z = v->Z, L = v->L, A = v->A;
    const int nBI = blocks((int64_t)B * I), nBZ = blocks((int64_t)B * Z);
    k_loss<<<1, 256, 0, s>>>(w.out, idx, perm, d_state, weights, w.zm, w.zlv, w.ga, w.gzm, w.gzlv, w.loss, B, L, A, Z);   // ga = d loss / d pre7
    // decoder, last layer first; "gp" = gradient at a layer's pre-activation
    k_dense_bwd_w<<<blocks((int64_t)I * D + D), 256, 0, s>>>(w.g3, w.ga, GR(W7), GR(B7), B, I, D);
    k_dense_bwd_x<<<nBI, 256, 0, s>>>(AT(W7), w.ga, nullptr, w.g3, w.gb, B, I, D, 1);        // gb = gp6
    k_dense_bwd_w<<<blocks((int64_t)I * I + I), 256, 0, s>>>(w.g2m, w.gb, GR(W6), GR(B6), B, I, I);
    k_dense_bwd_x<<<nBI, 256, 0, s>>>(AT(W6), w.gb, w.mask2, w.g2, w.ga, B, I, I, 1);        // ga = gp5 (through the dropout)
    k_dense_bwd_w<<<blocks((int64_t)I * I + I), 256, 0, s>>>(w.g1, w.ga, GR(W5), GR(B5), B, I, I);
    k_dense_bwd_x<<<nBI, 256, 0, s>>>(AT(W5), w.ga, nullptr, w.g1, w.gb, B, I, I, 1);        // gb = gp4
    k_dense_bwd_w<<<blocks((int64_t)Z * I + I), 256, 0, s>>>(w.z, w.gb, GR(W4), GR(B4), B, Z, I);
    k_dense_bwd_x<<<nBZ, 256, 0, s>>>(AT(W4), w.gb, nullptr, nullptr, w.gz, B, Z, I, 0);     // gz
    // reparameterisation and the two heads (gzm / gzlv already hold the KL part)
    k_sample_bwd<<<nBZ, 256, 0, s>>>(w.gz, w.zlv, w.eps, w.gzm, w.gzlv, B * Z);
    k_dense_bwd_w<<<blocks((int64_t)I * Z + Z), 256, 0, s>>>(w.h3, w.gzm, GR(WM), GR(BM), B, I, Z);
    k_dense_bwd_w<<<blocks((int64_t)I * Z + Z), 256, 0, s>>>(w.h3, w.gzlv, GR(WV), GR(BV), B, I, Z);
    k_dense_bwd_x<<<nBI, 256, 0, s>>>(AT(WM), w.gzm, nullptr, w.h3, w.ga, B, I, Z, 1);
    k_dense_bwd_x<<<nBI, 256, 0, s>>>(AT(WV), w.gzlv, nullptr, w.h3, w.gb, B, I, Z, 1);
    k_add<<<nBI, 256, 0, s>>>(w.ga, w.gb, w.ga, B * I);                                       // ga = gp3
    k_dense_bwd_w<<<blocks((int64_t)I * I + I), 256, 0, s>>>(w.bn, w.ga, GR(W3), GR(B3), B, I, I);
    k_dense_bwd_x<<<nBI, 256, 0, s>>>(AT(W3), w.ga, nullptr, nullptr, w.gb, B, I, I, 0);     // gb = d loss / d BatchNorm output
    k_bn_bwd<<<blocks(I), 256, 0, s>>>(w.gb, w.xhat, w.istd, AT(BNG), w.h2, w.ga, GR(BNG), GR(BNB), B, I);   // ga = gp2
    k_dense_bwd_w<<<blocks((int64_t)I * I + I), 256, 0, s>>>(w.h1m, w.ga, GR(W2), GR(B2), B, I, I);
    k_dense_bwd_x<<<nBI, 256, 0, s>>>(AT(W2), w.ga, w.mask1, w.h1, w.gb, B, I, I, 1);        // gb = gp1
    k_gather_bwd_w<<<blocks((int64_t)D * I + I), 256, 0, s>>>(idx, perm, d_state, w.gb, GR(W1), GR(B1), B, L, A, I);
}

// d_state == nullptr: the host counts the steps (train_step_dev); else the step number lives on the device (fit_dev)
void adam(flexs_vae *v, cudaStream_t s, const long long *d_state = nullptr) {
    float lr_t = 0.f;
    if (!d_state) {
        v->step += 1;
        const double t = (double)v->step;
        lr_t = (float)(ADAM_LR * std::sqrt(1.0 - std::pow((double)ADAM_B2, t)) / (1.0 - std::pow((double)ADAM_B1, t)));
    }
    k_adam_clip<<<blocks(v->total), 256, 0, s>>>(v->w, v->g, v->m, v->v, v->total, lr_t, d_state, v->offs[BNM],
                                                  v->offs[BNV] + v->sizes[BNV]);
}

// one optimiser step on the B rows perm[d_state[1] ..] with everything step-dependent read from d_state: capturable
void graph_step(flexs_vae *v, const Ws &w, const uint8_t *idx, const int *perm, const float *weights, int B, uint64_t seed,
                cudaStream_t s) {
    k_noise<<<blocks((int64_t)B * v->I), 256, 0, s>>>(w.mask1, w.mask2, w.eps, B * v->I, B * v->Z, seed, 0, v->d_state);
    forward(v, w, idx, perm, B, true, w.eps, s, v->d_state);
    backward(v, w, idx, perm, weights, B, s, v->d_state);
    adam(v, s, v->d_state);
    k_advance<<<1, 1, 0, s>>>(v->d_state, B, 1);
}

}  // namespace

extern "C" {

int flexs_vae_create(int device, int seq_len, int alphabet_size, int intermediate_dim, int latent_dim, flexs_vae_t **out) {
    FX_REQUIRE(out != nullptr, "out is null");
    *out = nullptr;
    FX_REQUIRE(seq_len >= 1 && alphabet_size >= 2 && alphabet_size <= 255, "bad sequence shape");
    FX_REQUIRE(intermediate_dim >= 1 && intermediate_dim <= 4096 && latent_dim >= 1 && latent_dim <= 64, "bad layer sizes");
    int ndev = 0;
    FX_CUDA(cudaGetDeviceCount(&ndev));
    FX_REQUIRE(device >= 0 && device < ndev, "device index out of range");
    flexs_vae *v = new flexs_vae();
    v->device = device; v->L = seq_len; v->A = alphabet_size; v->D = seq_len * alphabet_size; v->I = intermediate_dim; v->Z = latent_dim;
    const int64_t D = v->D, I = v->I, Z = v->Z;
    v->sizes = {D * I, I, I * I, I, I, I, I, I, I * I, I, I * Z, Z, I * Z, Z, Z * I, I, I * I, I, I * I, I, I * D, D};
    int64_t off = 0;
    for (int64_t sz : v->sizes) { v->offs.push_back(off); off += (sz + 3) / 4 * 4; }
    v->total = off;
    cudaSetDevice(device);
    const size_t bytes = sizeof(float) * (size_t)v->total;
    if (cudaMalloc(&v->w, bytes) != cudaSuccess || cudaMalloc(&v->g, bytes) != cudaSuccess ||
        cudaMalloc(&v->m, bytes) != cudaSuccess || cudaMalloc(&v->v, bytes) != cudaSuccess) {
        flexs_vae_destroy(v);
        return fx::cuda_fail(cudaGetLastError(), "cudaMalloc (VAE)");
    }
    cudaMemset(v->w, 0, bytes); cudaMemset(v->g, 0, bytes); cudaMemset(v->m, 0, bytes); cudaMemset(v->v, 0, bytes);
    *out = v;
    return FLEXS_OK;
}

void flexs_vae_destroy(flexs_vae_t *v) {
    if (!v) return;
    cudaSetDevice(v->device);
    cudaFree(v->w); cudaFree(v->g); cudaFree(v->m); cudaFree(v->v); cudaFree(v->ws); cudaFree(v->d_state);
    if (v->fit_stream) cudaStreamDestroy(v->fit_stream);
    delete v;
}

int flexs_vae_num_arrays(const flexs_vae_t *v) { return v ? NARR : FLEXS_EINVAL; }

int64_t flexs_vae_array_size(const flexs_vae_t *v, int i) {
    if (!v || i < 0 || i >= NARR) return FLEXS_EINVAL;
    return v->sizes[i];
}

static int copy_arrays(flexs_vae_t *v, float *base, float *const *h_arrays, bool to_device) {
    FX_REQUIRE(v && h_arrays, "null argument");
    FX_CUDA(cudaSetDevice(v->device));
    for (int i = 0; i < NARR; ++i) {
        FX_REQUIRE(h_arrays[i] != nullptr, "null array");
        if (to_device) FX_CUDA(cudaMemcpy(base + v->offs[i], h_arrays[i], sizeof(float) * v->sizes[i], cudaMemcpyHostToDevice));
        else FX_CUDA(cudaMemcpy(h_arrays[i], base + v->offs[i], sizeof(float) * v->sizes[i], cudaMemcpyDeviceToHost));
    }
    return FLEXS_OK;
}

int flexs_vae_set_weights(flexs_vae_t *v, const float *const *h_arrays) {
    return copy_arrays(v, v ? v->w : nullptr, const_cast<float *const *>(h_arrays), true);
}
int flexs_vae_get_weights(flexs_vae_t *v, float *const *h_arrays) { return copy_arrays(v, v ? v->w : nullptr, h_arrays, false); }
int flexs_vae_get_gradients(flexs_vae_t *v, float *const *h_arrays) { return copy_arrays(v, v ? v->g : nullptr, h_arrays, false); }

int flexs_vae_reset_optimizer(flexs_vae_t *v) {
    FX_REQUIRE(v, "null VAE");
    FX_CUDA(cudaSetDevice(v->device));
    FX_CUDA(cudaMemset(v->m, 0, sizeof(float) * v->total));
    FX_CUDA(cudaMemset(v->v, 0, sizeof(float) * v->total));
    v->step = 0;
    return FLEXS_OK;
}

int flexs_vae_train_step_dev(flexs_vae_t *v, const uint8_t *d_idx, const float *d_weights, int64_t n, const float *d_mask1,
                             const float *d_mask2, const float *d_eps, float *h_loss, void *stream) {
    FX_REQUIRE(v && d_idx && d_weights && d_mask1 && d_mask2 && d_eps, "null argument");
    FX_REQUIRE(n >= 2 && n <= 4096, "batch of 2..4096 sequences (BatchNorm needs two)");
    FX_CUDA(cudaSetDevice(v->device));
    cudaStream_t s = (cudaStream_t)stream;
    int rc = ensure_ws(v, (int)n);
    if (rc != FLEXS_OK) return rc;
    Ws w = carve(v, (int)n);
    const int B = (int)n;
    FX_CUDA(cudaMemcpyAsync(w.mask1, d_mask1, sizeof(float) * B * v->I, cudaMemcpyDeviceToDevice, s));
    FX_CUDA(cudaMemcpyAsync(w.mask2, d_mask2, sizeof(float) * B * v->I, cudaMemcpyDeviceToDevice, s));
    FX_CUDA(cudaMemcpyAsync(w.eps, d_eps, sizeof(float) * B * v->Z, cudaMemcpyDeviceToDevice, s));
    FX_CUDA(cudaMemsetAsync(w.loss, 0, sizeof(double), s));
    forward(v, w, d_idx, nullptr, B, true, w.eps, s);
    backward(v, w, d_idx, nullptr, d_weights, B, s);
    adam(v, s);
    double h = 0.0;
    FX_CUDA(cudaMemcpyAsync(&h, w.loss, sizeof(double), cudaMemcpyDeviceToHost, s));
    FX_CUDA(cudaStreamSynchronize(s));
    FX_CUDA(cudaGetLastError());
    if (h_loss) *h_loss = (float)(h / (double)B);
    return FLEXS_OK;
}

int flexs_vae_fit_dev(flexs_vae_t *v, const uint8_t *d_idx, const float *d_weights, int64_t n_train, int batch_size, int epochs,
                      int patience, uint64_t seed, float *h_losses, int *epochs_run, void *stream) {
    FX_REQUIRE(v && d_idx && d_weights, "null argument");
    FX_REQUIRE(n_train >= 2 && n_train < (1ll << 31), "need at least two training sequences");
    FX_REQUIRE(batch_size >= 2 && batch_size <= 4096 && epochs >= 0, "bad batch_size / epochs");
    FX_CUDA(cudaSetDevice(v->device));
    cudaStream_t s = (cudaStream_t)stream;
    const int Bmax = (int)std::min<int64_t>(batch_size, n_train);
    int rc = ensure_ws(v, Bmax);
    if (rc != FLEXS_OK) return rc;
    if (epochs_run) *epochs_run = 0;
    if (epochs == 0) return FLEXS_OK;
    // the epochs' permutations (Fisher-Yates on a splitmix64 stream), uploaded once
    std::vector<int> perm((size_t)n_train * epochs);
    uint64_t st = seed * 0x9E3779B97F4A7C15ull + 0x243F6A8885A308D3ull;
    auto next = [&]() { st += 0x9E3779B97F4A7C15ull; uint64_t z = st; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
                        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); };
    for (int e = 0; e < epochs; ++e) {
        int *p = perm.data() + (size_t)e * n_train;
        for (int64_t i = 0; i < n_train; ++i) p[i] = (int)i;
        for (int64_t i = n_train - 1; i > 0; --i) std::swap(p[i], p[(int64_t)(next() % (uint64_t)(i + 1))]);
    }
    int *d_perm = nullptr;
    FX_CUDA(cudaMalloc(&d_perm, sizeof(int) * perm.size()));
    if (!v->d_state) FX_CUDA(cudaMalloc(&v->d_state, 2 * sizeof(long long)));
    if (!v->fit_stream) FX_CUDA(cudaStreamCreateWithFlags(&v->fit_stream, cudaStreamNonBlocking));
    cudaStream_t fs = v->fit_stream;
    // the fit runs on its own stream (a graph cannot be captured on the legacy default stream), ordered after the caller's
    cudaEvent_t ev;
    FX_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    FX_CUDA(cudaEventRecord(ev, s));
    FX_CUDA(cudaStreamWaitEvent(fs, ev, 0));
    FX_CUDA(cudaMemcpyAsync(d_perm, perm.data(), sizeof(int) * perm.size(), cudaMemcpyHostToDevice, fs));
    const long long state0[2] = {(long long)v->step, 0};
    FX_CUDA(cudaMemcpyAsync(v->d_state, state0, sizeof(state0), cudaMemcpyHostToDevice, fs));
    Ws w = carve(v, Bmax);
    // ONE optimiser step (~45 launches) captured as a CUDA graph and replayed for every full mini-batch: the step number and
    // the position in the permutation are device-side state, so nothing in the graph changes between replays
    const int64_t n_full = n_train / batch_size;
    const int B_last = (int)(n_train - n_full * batch_size);
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    if (n_full > 0) {
        FX_CUDA(cudaStreamBeginCapture(fs, cudaStreamCaptureModeThreadLocal));
        graph_step(v, w, d_idx, d_perm, d_weights, batch_size, seed, fs);
        FX_CUDA(cudaStreamEndCapture(fs, &graph));
        FX_CUDA(cudaGraphInstantiate(&exec, graph, nullptr, nullptr, 0));
    }
    double best = 1e300;
    int bad = 0, ran = 0;
    int64_t steps_done = 0;
    for (int e = 0; e < epochs; ++e) {
        FX_CUDA(cudaMemsetAsync(w.loss, 0, sizeof(double), fs));
        int64_t counted = 0;
        for (int64_t b = 0; b < n_full; ++b) FX_CUDA(cudaGraphLaunch(exec, fs));
        counted += n_full * batch_size;
        steps_done += n_full;
        if (B_last >= 2) {   // the ragged last mini-batch: same kernels, launched directly
            graph_step(v, w, d_idx, d_perm, d_weights, B_last, seed, fs);
            counted += B_last;
            steps_done += 1;
        } else if (B_last == 1) {
            k_advance<<<1, 1, 0, fs>>>(v->d_state, 1, 0);   // BatchNorm needs more than one sample in training mode: skipped
        }
        double h = 0.0;
        FX_CUDA(cudaMemcpyAsync(&h, w.loss, sizeof(double), cudaMemcpyDeviceToHost, fs));
        FX_CUDA(cudaStreamSynchronize(fs));
        const double epoch_loss = h / (double)std::max<int64_t>(counted, 1);
        if (h_losses) h_losses[e] = (float)epoch_loss;
        ran = e + 1;
        if (epoch_loss < best - 1e-12) { best = epoch_loss; bad = 0; }
        else if (patience > 0 && ++bad >= patience) break;   // EarlyStopping(monitor="loss", patience=3), VAE_utils.py:139
    }
    v->step += steps_done;
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
    FX_CUDA(cudaEventRecord(ev, fs));
    FX_CUDA(cudaStreamWaitEvent(s, ev, 0));   // later work on the caller's stream sees the trained weights
    cudaEventDestroy(ev);
    cudaFree(d_perm);
    FX_CUDA(cudaGetLastError());
    if (epochs_run) *epochs_run = ran;
    return FLEXS_OK;
}

int flexs_vae_decode_dev(flexs_vae_t *v, const float *d_z, int64_t n, float *d_out, void *stream) {
    FX_REQUIRE(v && d_z && d_out, "null argument");
    FX_REQUIRE(n >= 1 && n <= 65536, "1..65536 latent vectors");
    FX_CUDA(cudaSetDevice(v->device));
    cudaStream_t s = (cudaStream_t)stream;
    int rc = ensure_ws(v, (int)n);
    if (rc != FLEXS_OK) return rc;
    Ws w = carve(v, (int)n);
    decode(v, w, d_z, (int)n, false, s);
    FX_CUDA(cudaMemcpyAsync(d_out, w.out, sizeof(float) * n * v->D, cudaMemcpyDeviceToDevice, s));
    FX_CUDA(cudaGetLastError());
    return FLEXS_OK;
}

int flexs_vae_log_prob_dev(flexs_vae_t *v, const uint8_t *d_idx, int64_t n, const float *d_eps, double *d_logp, void *stream) {
    FX_REQUIRE(v && d_idx && d_logp, "null argument");
    FX_REQUIRE(n >= 1, "no sequences");
    FX_CUDA(cudaSetDevice(v->device));
    cudaStream_t s = (cudaStream_t)stream;
    const int chunk = 4096;
    int rc = ensure_ws(v, (int)std::min<int64_t>(n, chunk));
    if (rc != FLEXS_OK) return rc;
    for (int64_t start = 0; start < n; start += chunk) {
        const int B = (int)std::min<int64_t>(chunk, n - start);
        Ws w = carve(v, B);
        forward(v, w, d_idx + start * v->L, nullptr, B, false, d_eps ? d_eps + start * v->Z : nullptr, s);
        k_log_prob<<<blocks(B), 256, 0, s>>>(w.out, d_idx + start * v->L, d_logp + start, B, v->L, v->A);
    }
    FX_CUDA(cudaGetLastError());
    return FLEXS_OK;
}

}  // extern "C"
