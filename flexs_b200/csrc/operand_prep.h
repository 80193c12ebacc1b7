// Host-side re-layout of fp32 Keras weights into the fp16 hi/lo operand planes the tcgen05 kernels consume
// (shared by cnn_umma2.cu, cnn_a20.cu and mlp_umma.cu).  Layouts are documented in umma2_layout.cuh.
#pragma once
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>

#include "umma2_layout.cuh"

namespace prep {

// power-of-two scale that brings max|w| into [2^14, 2^15): the lo halves stay in fp16's normal range
inline int scale_exponent(const float *src, size_t count, bool *finite) {
    float mx = 0.f;
    for (size_t i = 0; i < count; ++i) {
        if (!std::isfinite(src[i])) *finite = false;
        mx = std::max(mx, std::fabs(src[i]));
    }
    int e = 0;
    if (mx > 0.f && std::isfinite(mx)) e = 14 - (int)std::floor(std::log2(mx));
    return std::max(-24, std::min(e, 40));
}

// Conv kernel (taps, in g, out f), F = 32 -> planes [tap][chunk g/8][n: 0-31 hi, 32-63 lo][g%8] (u2::UWTAP bytes per tap).
// Returns the descale factor for activations that carry ASCALE.
inline float fill_conv_planes(const float *src, int taps, unsigned char *dst, bool *finite) {
    using namespace u2;
    const int e = scale_exponent(src, (size_t)taps * F * F, finite);
    const float scale = std::ldexp(1.f, e);
    __half *planes = reinterpret_cast<__half *>(dst);
    for (int j = 0; j < taps; ++j)
        for (int g = 0; g < F; ++g)
            for (int f = 0; f < F; ++f) {
                const float v = src[((size_t)j * F + g) * F + f] * scale;
                const __half hi = __float2half_rn(v);
                const __half lo = __float2half_rn(v - __half2float(hi));
                const size_t base = (size_t)j * (UWTAP / 2) + (size_t)(g >> 3) * (UWKC / 2) + (g & 7);
                planes[base + (size_t)f * 8] = hi;
                planes[base + (size_t)(32 + f) * 8] = lo;
            }
    return std::ldexp(1.f, -e) / ASCALE;
}

// One dense layer (in, out), out <= DH -> [k chunk][n: 0..DH-1 hi | DH..2DH-1 lo][8 k] planes (u2::DBK bytes per chunk).
inline float fill_dense_planes(const float *src, int kin, int H, unsigned char *dst, bool *finite) {
    using namespace u2;
    const int e = scale_exponent(src, (size_t)kin * H, finite);
    const float scale = std::ldexp(1.f, e);
    __half *planes = reinterpret_cast<__half *>(dst);
    for (int k = 0; k < kin; ++k)
        for (int o = 0; o < H; ++o) {
            const float v = src[(size_t)k * H + o] * scale;
            const __half hi = __float2half_rn(v);
            const __half lo = __float2half_rn(v - __half2float(hi));
            const size_t base = (size_t)(k >> 3) * (DBK / 2) + (k & 7);
            planes[base + (size_t)o * 8] = hi;
            planes[base + (size_t)(DH + o) * 8] = lo;
        }
    return std::ldexp(1.f, -e) / ASCALE;
}

// The dense head of a CNN member (cnn.py:49-52) at the u2 blob offsets OFF_DB1 / OFF_DB2 / OFF_DV of `blob`.
inline void fill_dense_head(const float *w, const fx::CnnOffsets &o, int H, unsigned char *blob, bool *finite) {
    using namespace u2;
    const float d1 = fill_dense_planes(w + o.wd1, F, H, blob + OFF_DB1, finite);
    const float d2 = fill_dense_planes(w + o.wd2, H, H, blob + OFF_DB2, finite);
    float *dv = reinterpret_cast<float *>(blob + OFF_DV);
    for (int i = 0; i < H; ++i) {
        dv[i] = w[o.bd1 + i] * ASCALE;
        dv[DH + i] = w[o.bd2 + i];
        dv[2 * DH + i] = w[o.wd3 + i];
    }
    dv[3 * DH] = d1 * ASCALE;  // layer-1 epilogue emits activations pre-scaled by ASCALE
    dv[3 * DH + 1] = d2;
    dv[3 * DH + 2] = w[o.bd3];
}

}  // namespace prep
