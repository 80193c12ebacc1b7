// Dense head shared by the fused CNN kernels: GlobalMaxPool features [32][slots] ->
// Dense(H,relu) -> Dense(H,relu) -> Dense(1) -> nan_to_num -> ensemble accumulate (cnn.py:49-52,
// keras_model.py:77-79, ensemble.py:54-59).  FP32 FFMA; each thread owns one output channel for
// 8 sequence slots so every weight it loads (coalesced across the warp) is used 8 times.
#pragma once
#include "common.cuh"

namespace fxd {

struct DenseArgs {
    const float *wd1, *bd1, *wd2, *bd2, *wd3, *bd3;  // global, Keras layout
    const float *featT;   // smem [F][sbp]
    float *scratch;       // smem, 2 * H * sbp floats (+ F*H + H*H floats when weights are staged)
    const long long *slot_seq;  // smem [nslots] -> global sequence index
    float *out;           // global scores
    int F, H, sbp, nslots, mem, M;
    int stage_weights;    // != 0: scratch has room to stage Wd1/Wd2 (the carve-out leaves too little
                          // L1 for the 53 KB of dense weights to stay cached between flushes)
};

template <int NT, bool WSMEM>
__device__ __forceinline__ void dense_layer_relu(const float *__restrict__ w, const float *__restrict__ b,
                                                 const float *__restrict__ xT, float *__restrict__ yT, int in,
                                                 int H, int sbp, int nsg) {
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
            const float wv = WSMEM ? w[(size_t)g * H + o] : __ldg(w + (size_t)g * H + o);
            const float4 x0 = *reinterpret_cast<const float4 *>(xT + (size_t)g * sbp + sg * 8);
            const float4 x1 = *reinterpret_cast<const float4 *>(xT + (size_t)g * sbp + sg * 8 + 4);
            acc[0] = fmaf(wv, x0.x, acc[0]); acc[1] = fmaf(wv, x0.y, acc[1]);
            acc[2] = fmaf(wv, x0.z, acc[2]); acc[3] = fmaf(wv, x0.w, acc[3]);
            acc[4] = fmaf(wv, x1.x, acc[4]); acc[5] = fmaf(wv, x1.y, acc[5]);
            acc[6] = fmaf(wv, x1.z, acc[6]); acc[7] = fmaf(wv, x1.w, acc[7]);
        }
        *reinterpret_cast<float4 *>(yT + (size_t)o * sbp + sg * 8) =
            make_float4(fmaxf(acc[0], 0.f), fmaxf(acc[1], 0.f), fmaxf(acc[2], 0.f), fmaxf(acc[3], 0.f));
        *reinterpret_cast<float4 *>(yT + (size_t)o * sbp + sg * 8 + 4) =
            make_float4(fmaxf(acc[4], 0.f), fmaxf(acc[5], 0.f), fmaxf(acc[6], 0.f), fmaxf(acc[7], 0.f));
    }
}

// All NT threads of the CTA must call this (it contains __syncthreads).
template <int NT>
__device__ void dense_head_flush(const DenseArgs &a) {
    const int nsg = (a.nslots + 7) >> 3;
    float *d1T = a.scratch;
    float *d2T = a.scratch + (size_t)a.H * a.sbp;
    __syncthreads();
    if (a.stage_weights) {
        float *s1 = d2T + (size_t)a.H * a.sbp;          // [F][H]
        float *s2 = s1 + (((size_t)a.F * a.H + 3) & ~(size_t)3);  // [H][H]
        for (int i = threadIdx.x; i < a.F * a.H; i += NT) s1[i] = __ldg(a.wd1 + i);
        for (int i = threadIdx.x; i < a.H * a.H; i += NT) s2[i] = __ldg(a.wd2 + i);
        __syncthreads();
        dense_layer_relu<NT, true>(s1, a.bd1, a.featT, d1T, a.F, a.H, a.sbp, nsg);
        __syncthreads();
        dense_layer_relu<NT, true>(s2, a.bd2, d1T, d2T, a.H, a.H, a.sbp, nsg);
    } else {
        dense_layer_relu<NT, false>(a.wd1, a.bd1, a.featT, d1T, a.F, a.H, a.sbp, nsg);
        __syncthreads();
        dense_layer_relu<NT, false>(a.wd2, a.bd2, d1T, d2T, a.H, a.H, a.sbp, nsg);
    }
    __syncthreads();
    for (int slot = threadIdx.x; slot < a.nslots; slot += NT) {
        float acc = 0.f;
#pragma unroll 4
        for (int g = 0; g < a.H; ++g) acc = fmaf(d2T[(size_t)g * a.sbp + slot], __ldg(a.wd3 + g), acc);
        const float y = nan_to_num(acc + __ldg(a.bd3));
        const long long seq = a.slot_seq[slot];
        // Ensemble default combine (ensemble.py:24): ((s0 + s1) + s2 ...) / M in fp32
        float tot = (a.mem == 0) ? y : a.out[seq] + y;
        if (a.M > 1 && a.mem == a.M - 1) tot = tot / (float)a.M;
        a.out[seq] = tot;
    }
    __syncthreads();
}

// floats of scratch the dense head needs (with / without weight staging)
__host__ __device__ inline size_t dense_scratch_floats(int F, int H, int sbp, bool staged) {
    size_t n = (size_t)2 * H * sbp;
    if (staged) n += (((size_t)F * H + 3) & ~(size_t)3) + (size_t)H * H;
    return n;
}

}  // namespace fxd
