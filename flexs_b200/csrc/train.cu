// K4 training: replaces keras Model.fit as called from KerasModel.train (keras_model.py:49-67) with
// the compile() settings of cnn.py:56 / mlp.py:33 — MSE loss, Adam(lr 1e-3, beta 0.9/0.999, eps 1e-7),
// mini-batches of `batch_size`, reshuffled every epoch, Dropout(0.25) after the second Dense of the
// CNN (cnn.py:51).  Weights and Adam moments live on the device and persist across calls, like the
// compiled Keras model's (explorer.py:157-160 retrains every round on the whole history).
//
// The data set is at most ~1000 measured sequences (B-1 per round), so this is a latency problem, not
// a throughput one: 80 optimiser steps of ~25 small launches on one stream with a single sync at the
// end.  Everything is deterministic (no atomics).  The reductions over (sample, position) — the conv
// weight/bias gradients — are two-stage: one CTA per sample group accumulates its partial sums in shared
// memory tiles, then a second kernel adds the partials in fixed group order.
#include <algorithm>
#include <cmath>
#include <random>
#include <vector>

#include "common.cuh"

namespace {

using fx::CnnDims; using fx::CnnOffsets; using fx::MlpOffsets; using fx::cnn_dims; using fx::cnn_offsets; using fx::mlp_offsets;

constexpr int TPB = 256;
constexpr float DROP_RATE = 0.25f;
constexpr float ADAM_LR = 1e-3f, ADAM_B1 = 0.9f, ADAM_B2 = 0.999f, ADAM_EPS = 1e-7f;

inline int grid_for(int64_t work) { return (int)std::max<int64_t>(1, std::min<int64_t>((work + TPB - 1) / TPB, 148 * 16)); }

// ---------------------------------------------------------------- Philox for the dropout mask
__device__ __forceinline__ void philox4(uint32_t k0, uint32_t k1, uint64_t ctr, uint64_t sub, uint32_t (&o)[4]) {
    uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)sub, (uint32_t)(sub >> 32)};
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    o[0] = c[0]; o[1] = c[1]; o[2] = c[2]; o[3] = c[3];
}

__global__ void k_dropout_mask(float *mask, int64_t count, uint64_t seed, uint64_t step) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q * 4 < count; q += stride) {
        uint32_t r[4];
        philox4((uint32_t)seed, (uint32_t)(seed >> 32), (uint64_t)q, step, r);
        for (int i = 0; i < 4 && q * 4 + i < count; ++i)
            mask[q * 4 + i] = ((float)(r[i] >> 8) * (1.0f / 16777216.0f) >= DROP_RATE) ? 1.f : 0.f;
    }
}

// ---------------------------------------------------------------- forward kernels
// rows of the mini-batch are perm[b] (perm == nullptr: identity)
__device__ __forceinline__ int64_t src_row(const int *perm, int b) { return perm ? perm[b] : b; }

__global__ void k_conv1_fwd(const uint8_t *idx, const int *perm, const float *w1, const float *b1, float *h1, int B,
                            int L, int A, int T, int F, int K) {
    const int64_t total = (int64_t)B * T * F, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int f = (int)(e % F), t = (int)((e / F) % T), b = (int)(e / ((int64_t)F * T));
        const uint8_t *ip = idx + src_row(perm, b) * L + t;
        float acc = b1[f];
        for (int j = 0; j < K; ++j) acc += w1[((size_t)j * A + ip[j]) * F + f];
        h1[e] = fmaxf(acc, 0.f);
    }
}

// Keras Conv1D(padding="same"), channels-last, kernel (K, F, F); output = relu(z)
__global__ void k_conv_fwd(const float *x, const float *w, const float *bias, float *h, int B, int T, int F, int K, int pl) {
    const int64_t total = (int64_t)B * T * F, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int f = (int)(e % F), t = (int)((e / F) % T), b = (int)(e / ((int64_t)F * T));
        float acc = bias[f];
        for (int j = 0; j < K; ++j) {
            const int s = t + j - pl;
            if (s < 0 || s >= T) continue;
            const float *xr = x + ((size_t)b * T + s) * F;
            const float *wr = w + (size_t)j * F * F + f;
            for (int g = 0; g < F; ++g) acc = fmaf(xr[g], wr[(size_t)g * F], acc);
        }
        h[e] = fmaxf(acc, 0.f);
    }
}

__global__ void k_gmax(const float *h3, float *p, int *am, int B, int T, int F) {
    const int64_t total = (int64_t)B * F, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int f = (int)(e % F), b = (int)(e / F);
        const float *col = h3 + (size_t)b * T * F + f;
        float best = col[0];
        int bi = 0;
        for (int t = 1; t < T; ++t) {
            const float v = col[(size_t)t * F];
            if (v > best) { best = v; bi = t; }  // first maximum wins (np.argmax)
        }
        p[e] = best; am[e] = bi;
    }
}

__global__ void k_gather_fwd(const uint8_t *idx, const int *perm, const float *w1, const float *b1, float *h, int B, int L,
                             int A, int H) {  // MLP layer 1 on a one-hot input
    const int64_t total = (int64_t)B * H, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int o = (int)(e % H), b = (int)(e / H);
        const uint8_t *ip = idx + src_row(perm, b) * L;
        float acc = b1[o];
        for (int l = 0; l < L; ++l) acc += w1[((size_t)l * A + ip[l]) * H + o];
        h[e] = fmaxf(acc, 0.f);
    }
}

__global__ void k_dense_fwd(const float *x, const float *w, const float *bias, float *y, int B, int in, int out, int relu) {
    const int64_t total = (int64_t)B * out, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int o = (int)(e % out), b = (int)(e / out);
        float acc = bias[o];
        const float *xr = x + (size_t)b * in;
        for (int g = 0; g < in; ++g) acc = fmaf(xr[g], w[(size_t)g * out + o], acc);
        y[e] = relu ? fmaxf(acc, 0.f) : acc;
    }
}

// out[b] = keep(d2)[b] . wd3 + bd3 ; also gout = 2/B (out - y) and the batch's summed squared error
__global__ void k_head(const float *d2, const float *mask, const float *wd3, const float *bd3, const float *labels,
                       const int *perm, float *outv, float *gout, double *sse, int B, int H) {
    __shared__ double red[TPB];
    double local = 0.0;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
        float acc = bd3[0];
        for (int g = 0; g < H; ++g) {
            const float keep = mask ? mask[(size_t)b * H + g] * (1.f / (1.f - DROP_RATE)) : 1.f;
            acc = fmaf(d2[(size_t)b * H + g] * keep, wd3[g], acc);
        }
        const float err = acc - labels[src_row(perm, b)];
        outv[b] = acc;
        gout[b] = 2.f * err / (float)B;
        local += (double)err * (double)err;
    }
    red[threadIdx.x] = local;
    __syncthreads();
    for (int s = TPB / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) sse[blockIdx.x] = red[0];
}

// ---------------------------------------------------------------- backward kernels
// last layer: gw[g] = sum_b keep*d2[b,g]*gout[b]; gb = sum_b gout[b]; gd2[b,g] = gout[b]*wd3[g]*keep*(d2>0)
// One CTA per hidden unit g (CTA H handles the bias): threads stride over the batch, then a fixed-shape shared-memory
// tree adds the TPB partial sums — deterministic, and no thread walks the whole batch alone.
__global__ void __launch_bounds__(TPB) k_head_bwd(const float *d2, const float *mask, const float *wd3, const float *gout,
                                                  float *gwd3, float *gbd3, float *gd2, int B, int H) {
    __shared__ float red[TPB];
    const int g = blockIdx.x;
    float acc = 0.f;
    if (g < H) {
        const float wg = wd3[g];
        for (int b = threadIdx.x; b < B; b += TPB) {
            const float keep = mask ? mask[(size_t)b * H + g] * (1.f / (1.f - DROP_RATE)) : 1.f;
            const float v = d2[(size_t)b * H + g];
            acc = fmaf(v * keep, gout[b], acc);
            gd2[(size_t)b * H + g] = (v > 0.f) ? gout[b] * wg * keep : 0.f;
        }
    } else {
        for (int b = threadIdx.x; b < B; b += TPB) acc += gout[b];
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = TPB / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (g < H) gwd3[g] = red[0]; else gbd3[0] = red[0];
    }
}

// gw[in,out] = sum_b x[b,in] * gy[b,out]; gb[out] = sum_b gy[b,out]
__global__ void k_dense_bwd_w(const float *x, const float *gy, float *gw, float *gb, int B, int in, int out) {
    const int64_t total = (int64_t)(in + 1) * out, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int o = (int)(e % out), g = (int)(e / out);
        float acc = 0.f;
        if (g < in) {
            for (int b = 0; b < B; ++b) acc = fmaf(x[(size_t)b * in + g], gy[(size_t)b * out + o], acc);
            gw[(size_t)g * out + o] = acc;
        } else {
            for (int b = 0; b < B; ++b) acc += gy[(size_t)b * out + o];
            gb[o] = acc;
        }
    }
}

// gx[b,in] = (x[b,in] > 0 or !relu_in) * sum_out gy[b,out] * w[in,out]
__global__ void k_dense_bwd_x(const float *x, const float *w, const float *gy, float *gx, int B, int in, int out, int relu_in) {
    const int64_t total = (int64_t)B * in, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int g = (int)(e % in), b = (int)(e / in);
        float acc = 0.f;
        const float *gr = gy + (size_t)b * out, *wr = w + (size_t)g * out;
        for (int o = 0; o < out; ++o) acc = fmaf(gr[o], wr[o], acc);
        gx[e] = (!relu_in || x[e] > 0.f) ? acc : 0.f;
    }
}

// GlobalMaxPooling backward fused with conv3's ReLU: gz3[b,t,f] = (t == am[b,f] && h3 > 0) ? gp[b,f] : 0
__global__ void k_gmax_bwd(const float *h3, const int *am, const float *gp, float *gz3, int B, int T, int F) {
    const int64_t total = (int64_t)B * T * F, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int f = (int)(e % F), t = (int)((e / F) % T), b = (int)(e / ((int64_t)F * T));
        gz3[e] = (am[(size_t)b * F + f] == t && h3[e] > 0.f) ? gp[(size_t)b * F + f] : 0.f;
    }
}

// MLP layer-1 weight gradient: gw1[l*A+a, o] = sum_{b : idx[b,l]==a} gu1[b,o]; gb1[o] = sum_b gu1[b,o]
__global__ void k_gather_bwd_w(const uint8_t *idx, const int *perm, const float *gu1, float *gw1, float *gb1, int B, int L,
                               int A, int H) {
    const int64_t total = (int64_t)(L * A + 1) * H, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int o = (int)(e % H);
        const int64_t row = e / H;
        float acc = 0.f;
        if (row < (int64_t)L * A) {
            const int l = (int)(row / A), a = (int)(row % A);
            for (int b = 0; b < B; ++b)
                if (idx[src_row(perm, b) * L + l] == a) acc += gu1[(size_t)b * H + o];
            gw1[e] = acc;
        } else {
            for (int b = 0; b < B; ++b) acc += gu1[(size_t)b * H + o];
            gb1[o] = acc;
        }
    }
}

// ---------------------------------------------------------------- two-stage (sample-group) gradient kernels
constexpr int GROUPS_MAX = 296;   // partial-sum slots: 2 CTAs per SM
constexpr int TCHUNK = 128;       // positions per shared-memory tile

// Stage 1 of the conv weight/bias gradient.  CTA `grp` owns samples grp, grp+G, ... and writes
//   part[grp][j,g,f] = sum_{b in group} sum_t x[b,t+j-pl,g] * gz[b,t,f],   part[grp][K*F*F + f] = sum gz[b,t,f]
// x and gz tiles (TCHUNK positions, x with a K-1 halo, zero outside [0,T)) are staged in shared memory.
__global__ void __launch_bounds__(TPB) k_conv_bwd_w_part(const float *__restrict__ x, const float *__restrict__ gz,
                                                         float *__restrict__ part, int B, int T, int F, int K, int pl) {
    extern __shared__ __align__(16) float tsm[];
    float *xs = tsm;                                   // [(TCHUNK + K - 1)][F]
    float *gs = tsm + (size_t)(TCHUNK + K - 1) * F;    // [TCHUNK][F]
    const int G = gridDim.x, grp = blockIdx.x, tid = threadIdx.x;
    const int count = K * F * F + F;
    float *mine = part + (size_t)grp * count;
    for (int e = tid; e < count; e += TPB) mine[e] = 0.f;
    const bool vec = (F % 4) == 0;
    for (int b = grp; b < B; b += G) {
        for (int c0 = 0; c0 < T; c0 += TCHUNK) {
            const int tc = min(TCHUNK, T - c0);
            __syncthreads();
            for (int e = tid; e < (tc + K - 1) * F; e += TPB) {
                const int s = c0 - pl + e / F;
                xs[e] = (s >= 0 && s < T) ? x[((size_t)b * T + s) * F + (e % F)] : 0.f;
            }
            for (int e = tid; e < tc * F; e += TPB) gs[e] = gz[((size_t)b * T + c0) * F + e];
            __syncthreads();
            for (int j = 0; j < K; ++j) {
                if (vec) {
                    const int fq = F / 4;
                    for (int u = tid; u < F * fq; u += TPB) {
                        const int g = u / fq, f0 = (u % fq) * 4;
                        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                        for (int t = 0; t < tc; ++t) {
                            const float xv = xs[(t + j) * F + g];
                            const float4 gv = *reinterpret_cast<const float4 *>(gs + t * F + f0);
                            acc.x = fmaf(xv, gv.x, acc.x); acc.y = fmaf(xv, gv.y, acc.y);
                            acc.z = fmaf(xv, gv.z, acc.z); acc.w = fmaf(xv, gv.w, acc.w);
                        }
                        float *o = mine + ((size_t)j * F + g) * F + f0;   // this thread is the only writer of these four
                        o[0] += acc.x; o[1] += acc.y; o[2] += acc.z; o[3] += acc.w;
                    }
                } else {
                    for (int u = tid; u < F * F; u += TPB) {
                        const int g = u / F, f = u % F;
                        float acc = 0.f;
                        for (int t = 0; t < tc; ++t) acc = fmaf(xs[(t + j) * F + g], gs[t * F + f], acc);
                        mine[((size_t)j * F + g) * F + f] += acc;
                    }
                }
            }
            for (int f = tid; f < F; f += TPB) {
                float acc = 0.f;
                for (int t = 0; t < tc; ++t) acc += gs[t * F + f];
                mine[(size_t)K * F * F + f] += acc;
            }
        }
    }
}

// Stage 1 of the conv1 (one-hot input) gradient: part[grp][j,a,f] = sum_{b,t : idx[b,t+j]==a} gz1[b,t,f]; then gb1.
// Thread (j,f) is the only writer of column [j][.][f] of the shared accumulator, so the scatter needs no atomics.
__global__ void __launch_bounds__(TPB) k_conv1_bwd_w_part(const uint8_t *__restrict__ idx, const int *__restrict__ perm,
                                                          const float *__restrict__ gz1, float *__restrict__ part, int B,
                                                          int L, int A, int T, int F, int K) {
    extern __shared__ __align__(16) float tsm[];   // [K][A][F] + [F]
    const int G = gridDim.x, grp = blockIdx.x, tid = threadIdx.x;
    const int count = K * A * F + F;
    for (int e = tid; e < count; e += TPB) tsm[e] = 0.f;
    __syncthreads();
    for (int u = tid; u < K * F; u += TPB) {
        const int j = u / F, f = u % F;
        float bsum = 0.f;
        for (int b = grp; b < B; b += G) {
            const uint8_t *ip = idx + src_row(perm, b) * L + j;
            const float *gr = gz1 + (size_t)b * T * F + f;
            for (int t = 0; t < T; ++t) {
                const float v = gr[(size_t)t * F];
                tsm[((size_t)j * A + ip[t]) * F + f] += v;
                bsum += v;
            }
        }
        if (j == 0) tsm[(size_t)K * A * F + f] = bsum;
    }
    __syncthreads();
    for (int e = tid; e < count; e += TPB) part[(size_t)grp * count + e] = tsm[e];
}

// Stage 2: out[e] = part[0][e] + part[1][e] + ... in group order (fixed for a given batch size).
__global__ void k_sum_parts(const float *__restrict__ part, int G, int count, int split, float *__restrict__ out_a,
                            float *__restrict__ out_b) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += stride) {
        float acc = 0.f;
        for (int g = 0; g < G; ++g) acc += part[(size_t)g * count + e];
        if (e < split) out_a[e] = acc; else out_b[e - split] = acc;
    }
}

// conv data gradient fused with the producer's ReLU, one CTA per sample: weights (padded rows) and a gz tile with a
// K-1 halo in shared memory;  gx[b,s,g] = (x[b,s,g] > 0) * sum_j sum_f gz[b,s-j+pl,f] * w[j,g,f]
__global__ void __launch_bounds__(TPB) k_conv_bwd_x_tile(const float *__restrict__ x, const float *__restrict__ w,
                                                         const float *__restrict__ gz, float *__restrict__ gx, int B,
                                                         int T, int F, int K, int pl) {
    extern __shared__ __align__(16) float tsm[];
    const int FP = F + 1;
    float *ws = tsm;                               // [K][F][F+1]
    float *gs = tsm + (size_t)K * F * FP;          // [(TCHUNK + K - 1)][F], row r holds t = c0 + pl - (K-1) + r
    const int tid = threadIdx.x;
    for (int e = tid; e < K * F * F; e += TPB) ws[(size_t)(e / F) * FP + (e % F)] = w[e];
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        for (int c0 = 0; c0 < T; c0 += TCHUNK) {
            const int tc = min(TCHUNK, T - c0);
            __syncthreads();
            for (int e = tid; e < (tc + K - 1) * F; e += TPB) {
                const int t = c0 + pl - (K - 1) + e / F;
                gs[e] = (t >= 0 && t < T) ? gz[((size_t)b * T + t) * F + (e % F)] : 0.f;
            }
            __syncthreads();
            for (int e = tid; e < tc * F; e += TPB) {
                const int sl = e / F, g = e % F;
                const size_t ge = ((size_t)b * T + c0 + sl) * F + g;
                float acc = 0.f;
                if (x[ge] > 0.f) {
                    for (int j = 0; j < K; ++j) {
                        const float *gr = gs + (size_t)(sl + (K - 1) - j) * F;
                        const float *wr = ws + ((size_t)j * F + g) * FP;
                        for (int f = 0; f < F; ++f) acc = fmaf(gr[f], wr[f], acc);
                    }
                }
                gx[ge] = acc;
            }
        }
    }
}

// Keras Adam (non-amsgrad): lr_t = lr*sqrt(1-b2^t)/(1-b1^t); w -= lr_t * m / (sqrt(v) + eps)
__global__ void k_adam(float *w, const float *g, float *m, float *v, int64_t count, float lr_t) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        const float gi = g[i];
        const float mi = ADAM_B1 * m[i] + (1.f - ADAM_B1) * gi;
        const float vi = ADAM_B2 * v[i] + (1.f - ADAM_B2) * gi * gi;
        m[i] = mi; v[i] = vi;
        w[i] -= lr_t * mi / (sqrtf(vi) + ADAM_EPS);
    }
}

// ---------------------------------------------------------------- host side
struct Ws {  // workspace carved out of one allocation, sized for batch B
    float *h1, *h2, *h3, *ga, *gb, *p, *gp, *d1, *d2, *gd1, *gd2, *outv, *gout, *mask, *grads;
    int *am;
    double *sse;
};

int64_t ws_floats(const flexs_model *m, int B) {
    const int64_t T = m->kind == FLEXS_KIND_CNN ? (m->L - m->K + 1) : 0, F = m->F, H = m->H;
    int64_t n = 0;
    n += 5 * (int64_t)B * T * F;          // h1 h2 h3 ga gb
    n += 2 * (int64_t)B * std::max<int64_t>(F, 1);  // p gp
    n += (int64_t)B * std::max<int64_t>(F, 1);      // am (as ints)
    n += 7 * (int64_t)B * H;              // d1 d2 d3(mlp) gd1 gd2 gd3 mask
    n += 2 * (int64_t)B;                  // outv gout
    n += m->member_floats;                // grads
    if (m->kind == FLEXS_KIND_CNN)        // per-group partial sums of the conv gradients
        n += (int64_t)GROUPS_MAX * (std::max<int64_t>({(int64_t)m->K * F * F, (int64_t)m->K3 * F * F, (int64_t)m->K * m->A * F}) + F);
    n += 2 * 64;                          // sse (doubles)
    return n + 256;
}

int ensure_training_state(flexs_model *m, int B) {
    FX_CUDA(cudaSetDevice(m->device));
    const size_t wbytes = sizeof(float) * m->member_floats * m->M;
    if (!m->d_adam_m) {
        FX_CUDA(cudaMalloc(&m->d_adam_m, wbytes));
        FX_CUDA(cudaMalloc(&m->d_adam_v, wbytes));
        FX_CUDA(cudaMemset(m->d_adam_m, 0, wbytes));
        FX_CUDA(cudaMemset(m->d_adam_v, 0, wbytes));
    }
    const int64_t need = ws_floats(m, B) * (int64_t)sizeof(float);
    if (m->train_ws_bytes < need) {
        cudaFree(m->train_ws);
        m->train_ws = nullptr; m->train_ws_bytes = 0;
        FX_CUDA(cudaMalloc(&m->train_ws, need));
        m->train_ws_bytes = need;
    }
    return FLEXS_OK;
}

// One optimiser step of one member on the mini-batch rows perm[0..B) (perm may be null = rows 0..B).
// `mask`: dropout mask [B,H] or null.  Adds the batch's summed squared error into *d_sse_accum (double).
int train_step(flexs_model *m, int member, const uint8_t *d_idx, const float *d_labels, const int *d_perm, int B,
               const float *d_mask, double *d_sse_out, cudaStream_t s) {
    const int L = m->L, A = m->A, H = m->H;
    float *w = m->d_weights + (int64_t)member * m->member_floats;
    float *am_ = m->d_adam_m + (int64_t)member * m->member_floats;
    float *av_ = m->d_adam_v + (int64_t)member * m->member_floats;
    float *base = reinterpret_cast<float *>(m->train_ws);
    float *grads;
    if (m->kind == FLEXS_KIND_CNN) {
        const CnnDims d = cnn_dims(m);
        const CnnOffsets o = cnn_offsets(m);
        const int T = d.T, F = d.F, K = d.K, K3 = d.K3;
        const int64_t btf = (int64_t)B * T * F;
        float *h1 = base, *h2 = h1 + btf, *h3 = h2 + btf, *ga = h3 + btf, *gb = ga + btf;
        float *p = gb + btf, *gp = p + (int64_t)B * F;
        int *am = reinterpret_cast<int *>(gp + (int64_t)B * F);
        float *d1 = reinterpret_cast<float *>(am + (int64_t)B * F), *d2 = d1 + (int64_t)B * H;
        float *gd1 = d2 + (int64_t)B * H, *gd2 = gd1 + (int64_t)B * H;
        float *outv = gd2 + (int64_t)B * H, *gout = outv + B;
        grads = gout + B;
        // ---- forward ----
        k_conv1_fwd<<<grid_for(btf), TPB, 0, s>>>(d_idx, d_perm, w + o.w1, w + o.b1, h1, B, L, A, T, F, K);
        k_conv_fwd<<<grid_for(btf), TPB, 0, s>>>(h1, w + o.w2, w + o.b2, h2, B, T, F, K, d.pl2);
        k_conv_fwd<<<grid_for(btf), TPB, 0, s>>>(h2, w + o.w3, w + o.b3, h3, B, T, F, K3, d.pl3);
        k_gmax<<<grid_for((int64_t)B * F), TPB, 0, s>>>(h3, p, am, B, T, F);
        k_dense_fwd<<<grid_for((int64_t)B * H), TPB, 0, s>>>(p, w + o.wd1, w + o.bd1, d1, B, F, H, 1);
        k_dense_fwd<<<grid_for((int64_t)B * H), TPB, 0, s>>>(d1, w + o.wd2, w + o.bd2, d2, B, H, H, 1);
        k_head<<<1, TPB, 0, s>>>(d2, d_mask, w + o.wd3, w + o.bd3, d_labels, d_perm, outv, gout, d_sse_out, B, H);
        // ---- backward ----
        k_head_bwd<<<H + 1, TPB, 0, s>>>(d2, d_mask, w + o.wd3, gout, grads + o.wd3, grads + o.bd3, gd2, B, H);
        k_dense_bwd_w<<<grid_for((int64_t)(H + 1) * H), TPB, 0, s>>>(d1, gd2, grads + o.wd2, grads + o.bd2, B, H, H);
        k_dense_bwd_x<<<grid_for((int64_t)B * H), TPB, 0, s>>>(d1, w + o.wd2, gd2, gd1, B, H, H, 1);
        k_dense_bwd_w<<<grid_for((int64_t)(F + 1) * H), TPB, 0, s>>>(p, gd1, grads + o.wd1, grads + o.bd1, B, F, H);
        k_dense_bwd_x<<<grid_for((int64_t)B * F), TPB, 0, s>>>(p, w + o.wd1, gd1, gp, B, F, H, 0);
        k_gmax_bwd<<<grid_for(btf), TPB, 0, s>>>(h3, am, gp, ga, B, T, F);                                  // ga = gz3
        float *part = grads + m->member_floats;
        const int G = std::min(B, GROUPS_MAX);
        const size_t sm_max = sizeof(float) * std::max({(size_t)(2 * TCHUNK + std::max(K, K3) - 1) * F,
                                                        (size_t)std::max(K, K3) * F * (F + 1) + (size_t)(TCHUNK + std::max(K, K3) - 1) * F,
                                                        (size_t)K * A * F + F});
        FX_REQUIRE(sm_max <= 200 * 1024, "training tiles do not fit shared memory (num_filters too large)");
        if (sm_max > 48 * 1024) {  // per device, cheap: no caching across devices
            FX_CUDA(cudaFuncSetAttribute(k_conv_bwd_w_part, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            FX_CUDA(cudaFuncSetAttribute(k_conv_bwd_x_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            FX_CUDA(cudaFuncSetAttribute(k_conv1_bwd_w_part, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        }
        auto conv_bwd = [&](const float *xin, const float *wgt, const float *gzz, float *gxx, int64_t ow, int64_t ob,
                            int k, int pl) {
            const int cnt = k * F * F + F;
            const size_t sm_w = sizeof(float) * ((size_t)(TCHUNK + k - 1) * F + (size_t)TCHUNK * F);
            k_conv_bwd_w_part<<<G, TPB, sm_w, s>>>(xin, gzz, part, B, T, F, k, pl);
            k_sum_parts<<<grid_for(cnt), TPB, 0, s>>>(part, G, cnt, k * F * F, grads + ow, grads + ob);
            const size_t sm_x = sizeof(float) * ((size_t)k * F * (F + 1) + (size_t)(TCHUNK + k - 1) * F);
            k_conv_bwd_x_tile<<<std::min(B, 148 * 4), TPB, sm_x, s>>>(xin, wgt, gzz, gxx, B, T, F, k, pl);
        };
        conv_bwd(h2, w + o.w3, ga, gb, o.w3, o.b3, K3, d.pl3);                                                // gb = gz2
        conv_bwd(h1, w + o.w2, gb, ga, o.w2, o.b2, K, d.pl2);                                                 // ga = gz1
        {
            const int cnt = K * A * F + F;
            k_conv1_bwd_w_part<<<G, TPB, sizeof(float) * cnt, s>>>(d_idx, d_perm, ga, part, B, L, A, T, F, K);
            k_sum_parts<<<grid_for(cnt), TPB, 0, s>>>(part, G, cnt, K * A * F, grads + o.w1, grads + o.b1);
        }
        m->launches += 21;
    } else {
        const MlpOffsets o = mlp_offsets(m);
        float *x1 = base, *x2 = x1 + (int64_t)B * H, *x3 = x2 + (int64_t)B * H;
        float *g1 = x3 + (int64_t)B * H, *g2 = g1 + (int64_t)B * H, *g3 = g2 + (int64_t)B * H;
        float *outv = g3 + (int64_t)B * H, *gout = outv + B;
        grads = gout + B;
        k_gather_fwd<<<grid_for((int64_t)B * H), TPB, 0, s>>>(d_idx, d_perm, w + o.w1, w + o.b1, x1, B, L, A, H);
        k_dense_fwd<<<grid_for((int64_t)B * H), TPB, 0, s>>>(x1, w + o.w2, w + o.b2, x2, B, H, H, 1);
        k_dense_fwd<<<grid_for((int64_t)B * H), TPB, 0, s>>>(x2, w + o.w3, w + o.b3, x3, B, H, H, 1);
        k_head<<<1, TPB, 0, s>>>(x3, nullptr, w + o.w4, w + o.b4, d_labels, d_perm, outv, gout, d_sse_out, B, H);
        k_head_bwd<<<H + 1, TPB, 0, s>>>(x3, nullptr, w + o.w4, gout, grads + o.w4, grads + o.b4, g3, B, H);
        k_dense_bwd_w<<<grid_for((int64_t)(H + 1) * H), TPB, 0, s>>>(x2, g3, grads + o.w3, grads + o.b3, B, H, H);
        k_dense_bwd_x<<<grid_for((int64_t)B * H), TPB, 0, s>>>(x2, w + o.w3, g3, g2, B, H, H, 1);
        k_dense_bwd_w<<<grid_for((int64_t)(H + 1) * H), TPB, 0, s>>>(x1, g2, grads + o.w2, grads + o.b2, B, H, H);
        k_dense_bwd_x<<<grid_for((int64_t)B * H), TPB, 0, s>>>(x1, w + o.w2, g2, g1, B, H, H, 1);
        k_gather_bwd_w<<<grid_for((int64_t)(L * A + 1) * H), TPB, 0, s>>>(d_idx, d_perm, g1, grads + o.w1, grads + o.b1, B, L, A, H);
        m->launches += 10;
    }
    // ---- Adam ----
    const int64_t t = ++m->adam_step[member];
    const double lr_t = (double)ADAM_LR * std::sqrt(1.0 - std::pow((double)ADAM_B2, (double)t)) /
                        (1.0 - std::pow((double)ADAM_B1, (double)t));
    k_adam<<<grid_for(m->member_floats), TPB, 0, s>>>(w, grads, am_, av_, m->member_floats, (float)lr_t);
    m->launches += 1;
    FX_CUDA(cudaGetLastError());
    m->umma_ready = false;
    m->umma2_ready = false;
    m->k9_ready = false;
    m->enum_ready = false;
    return FLEXS_OK;
}

float *mask_buffer(flexs_model *m, int B) {  // lives behind everything else in the workspace
    float *base = reinterpret_cast<float *>(m->train_ws);
    return base + (ws_floats(m, B) - 256 - 2 * 64 - (int64_t)B * m->H);
}
double *sse_buffer(flexs_model *m, int B) {
    float *base = reinterpret_cast<float *>(m->train_ws);
    uintptr_t a = reinterpret_cast<uintptr_t>(base + (ws_floats(m, B) - 256 - 2 * 64));
    return reinterpret_cast<double *>((a + 7) & ~(uintptr_t)7);  // the 256-float tail leaves room to align
}

}  // namespace

using namespace fx;

extern "C" {

int flexs_model_train_step_dev(flexs_model_t *m, int member, const uint8_t *d_idx, const float *d_labels, int64_t n,
                               const float *d_dropout_mask, float *h_loss, void *stream) {
    FX_REQUIRE(m && d_idx && d_labels, "null argument");
    FX_REQUIRE(member >= 0 && member < m->M, "member out of range");
    FX_REQUIRE(n >= 1 && n <= 65536, "batch of 1..65536 sequences");
    FX_REQUIRE(d_dropout_mask == nullptr || m->kind == FLEXS_KIND_CNN, "only the CNN has a Dropout layer (cnn.py:51)");
    FX_CUDA(cudaSetDevice(m->device));
    cudaStream_t s = (cudaStream_t)stream;
    int rc = ensure_training_state(m, (int)n);
    if (rc != FLEXS_OK) return rc;
    double *sse = sse_buffer(m, (int)n);
    rc = train_step(m, member, d_idx, d_labels, nullptr, (int)n, d_dropout_mask, sse, s);
    if (rc != FLEXS_OK) return rc;
    double h = 0.0;
    FX_CUDA(cudaMemcpyAsync(&h, sse, sizeof(double), cudaMemcpyDeviceToHost, s));
    FX_CUDA(cudaStreamSynchronize(s));
    if (h_loss) *h_loss = (float)(h / (double)n);
    return FLEXS_OK;
}

int flexs_model_fit_dev(flexs_model_t *m, const uint8_t *d_idx, const float *d_labels, int64_t n, int batch_size,
                        int epochs, uint64_t seed, float *h_losses, void *stream) {
    FX_REQUIRE(m && d_idx && d_labels, "null argument");
    FX_REQUIRE(n >= 1 && n < (1ll << 31), "bad n");
    FX_REQUIRE(batch_size >= 1 && epochs >= 0, "bad batch_size / epochs");
    FX_CUDA(cudaSetDevice(m->device));
    cudaStream_t s = (cudaStream_t)stream;
    const int B = (int)std::min<int64_t>(batch_size, n);
    int rc = ensure_training_state(m, B);
    if (rc != FLEXS_OK) return rc;
    const int nbatch = (int)((n + B - 1) / B);
    // per (member, epoch) permutations, uploaded once; per (member, epoch, batch) SSE slots read back once
    const int64_t total_steps = (int64_t)m->M * epochs * nbatch;
    if (total_steps == 0) return FLEXS_OK;
    int *d_perm = nullptr;
    double *d_sse = nullptr;
    FX_CUDA(cudaMalloc(&d_perm, sizeof(int) * (size_t)n * m->M * std::max(epochs, 1)));
    FX_CUDA(cudaMalloc(&d_sse, sizeof(double) * (size_t)total_steps));
    std::vector<int> perm((size_t)n * m->M * epochs);
    for (int mem = 0; mem < m->M; ++mem) {
        std::mt19937_64 rng(seed * 0x9E3779B97F4A7C15ull + (uint64_t)mem + 1);
        for (int e = 0; e < epochs; ++e) {
            int *pp = perm.data() + ((size_t)mem * epochs + e) * n;
            for (int64_t i = 0; i < n; ++i) pp[i] = (int)i;
            std::shuffle(pp, pp + n, rng);  // Keras fit(shuffle=True): new order every epoch
        }
    }
    FX_CUDA(cudaMemcpyAsync(d_perm, perm.data(), sizeof(int) * perm.size(), cudaMemcpyHostToDevice, s));
    float *mask = (m->kind == FLEXS_KIND_CNN) ? mask_buffer(m, B) : nullptr;
    int64_t step = 0;
    for (int mem = 0; mem < m->M && rc == FLEXS_OK; ++mem)
        for (int e = 0; e < epochs && rc == FLEXS_OK; ++e)
            for (int b = 0; b < nbatch && rc == FLEXS_OK; ++b, ++step) {
                const int64_t start = (int64_t)b * B;
                const int cur = (int)std::min<int64_t>(B, n - start);
                const int *pp = d_perm + ((size_t)mem * epochs + e) * n + start;
                if (mask) {
                    k_dropout_mask<<<grid_for(((int64_t)cur * m->H + 3) / 4), TPB, 0, s>>>(
                        mask, (int64_t)cur * m->H, seed ^ 0xD1B54A32D192ED03ull, (uint64_t)(m->adam_step[mem] + 1) * 131 + mem);
                    m->launches += 1;
                }
                rc = train_step(m, mem, d_idx, d_labels, pp, cur, mask, d_sse + step, s);
            }
    std::vector<double> sse((size_t)total_steps, 0.0);
    if (rc == FLEXS_OK) {
        cudaError_t e1 = cudaMemcpyAsync(sse.data(), d_sse, sizeof(double) * sse.size(), cudaMemcpyDeviceToHost, s);
        cudaError_t e2 = cudaStreamSynchronize(s);
        if (e1 != cudaSuccess || e2 != cudaSuccess) rc = cuda_fail(e1 != cudaSuccess ? e1 : e2, "fit readback");
    }
    cudaFree(d_perm);
    cudaFree(d_sse);
    if (rc != FLEXS_OK) return rc;
    if (h_losses) {
        step = 0;
        for (int mem = 0; mem < m->M; ++mem)
            for (int e = 0; e < epochs; ++e) {
                double acc = 0.0;
                for (int b = 0; b < nbatch; ++b, ++step) acc += sse[step];
                h_losses[(size_t)mem * epochs + e] = (float)(acc / (double)n);  // mean squared error over the epoch
            }
    }
    return FLEXS_OK;
}

int flexs_model_get_optimizer_state(flexs_model_t *m, int member, float *const *h_m, float *const *h_v, int64_t *step) {
    FX_REQUIRE(m, "null model");
    FX_REQUIRE(member >= 0 && member < m->M, "member out of range");
    FX_CUDA(cudaSetDevice(m->device));
    if (step) *step = m->adam_step[member];
    for (size_t i = 0; i < m->arr_sizes.size(); ++i) {
        const size_t bytes = sizeof(float) * m->arr_sizes[i];
        const int64_t off = (int64_t)member * m->member_floats + m->arr_offs[i];
        if (h_m && h_m[i]) {
            if (m->d_adam_m) FX_CUDA(cudaMemcpy(h_m[i], m->d_adam_m + off, bytes, cudaMemcpyDeviceToHost));
            else std::fill(h_m[i], h_m[i] + m->arr_sizes[i], 0.f);
        }
        if (h_v && h_v[i]) {
            if (m->d_adam_v) FX_CUDA(cudaMemcpy(h_v[i], m->d_adam_v + off, bytes, cudaMemcpyDeviceToHost));
            else std::fill(h_v[i], h_v[i] + m->arr_sizes[i], 0.f);
        }
    }
    return FLEXS_OK;
}

int flexs_model_reset_optimizer(flexs_model_t *m) {
    FX_REQUIRE(m, "null model");
    FX_CUDA(cudaSetDevice(m->device));
    const size_t wbytes = sizeof(float) * m->member_floats * m->M;
    if (m->d_adam_m) {
        FX_CUDA(cudaMemset(m->d_adam_m, 0, wbytes));
        FX_CUDA(cudaMemset(m->d_adam_v, 0, wbytes));
    }
    std::fill(m->adam_step.begin(), m->adam_step.end(), 0);
    return FLEXS_OK;
}

}  // extern "C"
