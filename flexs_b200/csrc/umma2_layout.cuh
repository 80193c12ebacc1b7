// Shared pieces of the tcgen05 kernels for the A = 4 shapes (cnn_umma2.cu, cnn_k9.cu): operand-blob layout,
// PTX wrappers (tcgen05 / mbarrier / descriptors), the per-tile MMA issue loops and the fp16 hi/lo split.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace u2 {

constexpr int F = 32, K = 5, K3 = 3, ALPHA = 4;
constexpr int UWTAP = 4 * 64 * 16;  // one tap: 4 channel chunks x ([hi|lo] 64 filters) x 16 B
constexpr int UWKC = 64 * 16;       // one channel chunk of a tap
constexpr float ASCALE = 8.f;
constexpr uint32_t IDESC_N64 = (1u << 4) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t IDESC_N32 = (1u << 4) | ((32u >> 3) << 17) | ((128u >> 4) << 24);

// per-member operand blob built by prepare_cnn_umma2: UW2 | UW3 | T012 | T34 | inv2s, inv3 | dense planes | vectors | TBIG
constexpr int OFF_UW3 = K * UWTAP, OFF_T012 = OFF_UW3 + K3 * UWTAP, OFF_T34 = OFF_T012 + 64 * F * 4;
constexpr int OFF_SCAL = OFF_T34 + 16 * F * 4;
// tensor-core dense head (H <= 112): [128 sequence slots] x [32 -> 112 -> 112] as two UMMA GEMMs with the same
// fp16 hi/lo split; weights as [k chunk][n: hi 0..111 | lo 112..223][8 k] planes, vectors padded to 112
constexpr int DH = 112, DN = 2 * DH, DSLOTS = 128, DPLANE = DSLOTS * 16, DBK = DN * 16;
constexpr int OFF_DB1 = (OFF_SCAL + 16 + 255) / 256 * 256, OFF_DB2 = OFF_DB1 + 4 * DBK, OFF_DV = OFF_DB2 + 14 * DBK;
constexpr int DV_FLOATS = 3 * DH + 4;  // bd1*ASCALE | bd2 | wd3 | inv_d1s, inv_d2, bd3
constexpr int OFF_TBIG = (OFF_DV + DV_FLOATS * 4 + 255) / 256 * 256;  // 4^5 entries x (4 chunks hi | 4 chunks lo) x 16 B
constexpr int UW_MEMBER_BYTES = OFF_TBIG + 1024 * 128;
// dense-head scratch (offsets inside the idle activation buffers)
constexpr int DS_X1 = 0, DS_B1 = DS_X1 + 8 * DPLANE, DS_X2 = DS_B1 + 4 * DBK, DS_B2 = DS_X2 + 28 * DPLANE;
constexpr int DS_PART = DS_B2 + 14 * DBK, DS_DV = DS_PART + 2 * DSLOTS * 4, DS_TOTAL = DS_DV + (DV_FLOATS * 4 + 15) / 16 * 16;
constexpr uint32_t IDESC_DN = (1u << 4) | ((uint32_t)(DN >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t IDESC_DH = (1u << 4) | ((uint32_t)(DH >> 3) << 17) | ((128u >> 4) << 24);

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(fxd::smem_u32(dst_smem)),
                 "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(fxd::smem_u32(bar)) : "memory");
}
// tcgen05.mma issued by ONE elected lane, but reached by the whole (converged) warp: control flow stays
// warp-uniform, so the descriptor arithmetic runs on the uniform datapath instead of one thread's
// R2UR-latency-bound chain (the first v2 profile spent ~2000 single-thread instructions per chunk there).
__device__ __forceinline__ void umma_f16_elect(uint32_t d_tmem, uint32_t a_lo32, uint32_t a_hi32, uint32_t b_lo32,
                                               uint32_t b_hi32, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "elect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %6, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo32), "r"(a_hi32), "r"(b_lo32), "r"(b_hi32), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint64_t *bar) {
    asm volatile(
        "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        ::"r"(fxd::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (verified on hardware by the v1 kernel):
//   lo word: bits 0-13 start address >> 4, bits 16-29 LBO >> 4 (bytes between the two 8-element K chunks)
//   hi word: bits 0-13 SBO >> 4 (bytes between consecutive 8-row core matrices), bit 14 = version 1
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr, uint32_t lbo) { return (smem_addr >> 4) | ((lbo >> 4) << 16); }
constexpr uint32_t DESC_HI = (128u >> 4) | (1u << 14);

// one 128-row tile of an implicit-GEMM conv: taps x 2 channel pairs x {A_hi x [W_hi|W_lo], A_lo x W_hi}.
// Called by all 32 lanes of the MMA warp; every offset is a compile-time constant added to two base words.
template <int TAPS, uint32_t A_PLANE>
__device__ __forceinline__ void issue_conv_tile(uint32_t a_tile_addr, uint32_t w_addr, uint32_t d_tmem) {
    const uint32_t a0 = desc_lo(a_tile_addr, A_PLANE), b0 = desc_lo(w_addr, UWKC);
#pragma unroll
    for (int j = 0; j < TAPS; ++j) {
#pragma unroll
        for (int kp = 0; kp < 2; ++kp) {
            const uint32_t a_hi = a0 + (((uint32_t)(2 * kp) * A_PLANE + (uint32_t)j * 128u) >> 4);
            const uint32_t a_lo = a_hi + ((4u * A_PLANE) >> 4);
            const uint32_t bd = b0 + (((uint32_t)j * UWTAP + (uint32_t)(2 * kp) * UWKC) >> 4);
            umma_f16_elect(d_tmem, a_hi, DESC_HI, bd, DESC_HI, IDESC_N64, (j | kp) ? 1u : 0u);
            umma_f16_elect(d_tmem, a_lo, DESC_HI, bd, DESC_HI, IDESC_N32, 1u);
        }
    }
}

// one dense layer: [128 slots, 16*KP] x [16*KP, 112] with X = hi + lo planes (NPL planes per split)
template <int KP, int NPL>
__device__ __forceinline__ void issue_dense_layer(uint32_t x_addr, uint32_t b_addr, uint32_t d_tmem) {
    const uint32_t a0 = desc_lo(x_addr, DPLANE), b0 = desc_lo(b_addr, DBK);
#pragma unroll
    for (int kp = 0; kp < KP; ++kp) {
        const uint32_t a_hi = a0 + (((uint32_t)(2 * kp) * DPLANE) >> 4);
        const uint32_t a_lo = a_hi + (((uint32_t)NPL * DPLANE) >> 4);
        const uint32_t bd = b0 + (((uint32_t)(2 * kp) * DBK) >> 4);
        umma_f16_elect(d_tmem, a_hi, DESC_HI, bd, DESC_HI, IDESC_DN, kp ? 1u : 0u);
        umma_f16_elect(d_tmem, a_lo, DESC_HI, bd, DESC_HI, IDESC_DH, 1u);
    }
}

// 8 fp32 -> fp16 hi row + lo row; mx tracks the largest value seen (fp16 range guard)
__device__ __forceinline__ void split8(const float (&x)[8], uint4 &hi4, uint4 &lo4, float &mx) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float a = x[2 * i], b = x[2 * i + 1];
        mx = fmaxf(mx, fmaxf(a, b));
        const __half2 h = __floats2half2_rn(a, b);
        const float2 back = __half22float2(h);
        const __half2 l = __floats2half2_rn(a - back.x, b - back.y);
        hi[i] = *reinterpret_cast<const uint32_t *>(&h);
        lo[i] = *reinterpret_cast<const uint32_t *>(&l);
    }
    hi4 = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    lo4 = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// Dense head on the tensor cores (H <= 112): features [32][sbp] of up to 128 sequence slots -> Dense(H,relu) ->
// Dense(H,relu) -> Dense(1) -> nan_to_num -> ensemble accumulate (cnn.py:49-52, keras_model.py:77-79,
// ensemble.py:54-59) as two UMMA GEMMs with the fp16 hi/lo split.  `scratch` (>= DS_TOTAL bytes, 1024-aligned) is
// the idle activation memory; TMEM columns 0..223 and 256..479 must be free.  Every thread of the CTA calls this
// (it contains __syncthreads); warps 0-7 run the epilogues, warp MMAW issues.  seq_of(slot) -> global sequence.
// Stage the dense weight planes and the bias / output-weight vectors of one member into the scratch.
template <int NT>
__device__ __forceinline__ void dense_stage_weights(unsigned char *scratch, const unsigned char *uw) {
    const int tid = threadIdx.x;
    const float *gdv = reinterpret_cast<const float *>(uw + OFF_DV);
    float *dv = reinterpret_cast<float *>(scratch + DS_DV);
    for (int i = tid; i < 3 * DH; i += NT) dv[i] = __ldg(gdv + i);
    for (int i = tid; i < 4 * DBK / 16; i += NT)
        reinterpret_cast<uint4 *>(scratch + DS_B1)[i] = __ldg(reinterpret_cast<const uint4 *>(uw + OFF_DB1) + i);
    for (int i = tid; i < 14 * DBK / 16; i += NT)
        reinterpret_cast<uint4 *>(scratch + DS_B2)[i] = __ldg(reinterpret_cast<const uint4 *>(uw + OFF_DB2) + i);
}

// STAGE: stage the weights on every call (the scratch is shared with other work between calls); otherwise the
// caller ran dense_stage_weights once and the scratch is dedicated.
template <int NT, int MMAW, bool STAGE, class SeqOf>
__device__ __forceinline__ void dense_head_umma(unsigned char *scratch, const float *featT, int sbp, int sbcap,
                                                int g_slots, const unsigned char *uw, uint32_t tmem_base,
                                                uint64_t *dbar, uint32_t &dph, float &xmax, int mem, int M,
                                                float *out, SeqOf seq_of) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform for the compiler: MMA issue on the uniform datapath
    unsigned char *dx1 = scratch + DS_X1, *db1 = scratch + DS_B1, *dx2 = scratch + DS_X2, *db2 = scratch + DS_B2;
    float *dpart = reinterpret_cast<float *>(scratch + DS_PART);
    const float *gdv = reinterpret_cast<const float *>(uw + OFF_DV);
    float *dv = reinterpret_cast<float *>(scratch + DS_DV);  // bias / output-weight vectors, staged in smem
    const float inv_d1s = __ldg(gdv + 3 * DH), inv_d2 = __ldg(gdv + 3 * DH + 1), bd3v = __ldg(gdv + 3 * DH + 2);
    // (1) stage the dense weight planes, turn the features into the A operand of layer 1
    if (STAGE) dense_stage_weights<NT>(scratch, uw);
    for (int t = tid; t < 4 * DSLOTS; t += NT) {
        const int slot = t & (DSLOTS - 1), cchunk = t >> 7;
        float x[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) x[q] = (slot < sbcap) ? featT[(cchunk * 8 + q) * sbp + slot] * ASCALE : 0.f;
        uint4 hi4, lo4;
        split8(x, hi4, lo4, xmax);
        *reinterpret_cast<uint4 *>(dx1 + (size_t)cchunk * DPLANE + slot * 16) = hi4;
        *reinterpret_cast<uint4 *>(dx1 + (size_t)(4 + cchunk) * DPLANE + slot * 16) = lo4;
    }
    fence_async_smem();
    __syncthreads();
    const int lq = wid & 3, half = wid >> 2, slot = 32 * lq + lane;
    const uint32_t tl = tmem_base + ((uint32_t)(lq * 32) << 16);
    // (2) layer 1 on the tensor cores, (3) bias + ReLU + split -> A operand of layer 2
    if (wid == MMAW) {
        tc_fence_after();
        issue_dense_layer<2, 4>(fxd::smem_u32(dx1), fxd::smem_u32(db1), tmem_base);
        umma_commit_elect(dbar);
    } else if (wid < 8) {
        fxd::mbar_wait(dbar, dph & 1);
        tc_fence_after();
        for (int c7 = 0; c7 < 7; ++c7) {
            const int cchunk = half * 7 + c7;
            uint32_t va[8], vb[8];
            tmem_ld8_nowait(tl + (uint32_t)(cchunk * 8), va);
            tmem_ld8_nowait(tl + (uint32_t)(DH + cchunk * 8), vb);
            const float4 b0 = *reinterpret_cast<const float4 *>(dv + cchunk * 8);
            const float4 b1 = *reinterpret_cast<const float4 *>(dv + cchunk * 8 + 4);
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            tmem_ld_wait();
            float x[8];
#pragma unroll
            for (int q = 0; q < 8; ++q)
                x[q] = fmaxf(fmaf(__uint_as_float(va[q]) + __uint_as_float(vb[q]), inv_d1s, bb[q]), 0.f);
            uint4 hi4, lo4;
            split8(x, hi4, lo4, xmax);
            *reinterpret_cast<uint4 *>(dx2 + (size_t)cchunk * DPLANE + slot * 16) = hi4;
            *reinterpret_cast<uint4 *>(dx2 + (size_t)(14 + cchunk) * DPLANE + slot * 16) = lo4;
        }
    }
    ++dph;
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    // (4) layer 2, (5) bias + ReLU + dot with the output weights
    if (wid == MMAW) {
        tc_fence_after();
        issue_dense_layer<7, 14>(fxd::smem_u32(dx2), fxd::smem_u32(db2), tmem_base + 256u);
        umma_commit_elect(dbar);
    } else if (wid < 8) {
        fxd::mbar_wait(dbar, dph & 1);
        tc_fence_after();
        float sum = 0.f;
        for (int c7 = 0; c7 < 7; ++c7) {
            const int cchunk = half * 7 + c7;
            uint32_t va[8], vb[8];
            tmem_ld8_nowait(tl + 256u + (uint32_t)(cchunk * 8), va);
            tmem_ld8_nowait(tl + 256u + (uint32_t)(DH + cchunk * 8), vb);
            const float4 b0 = *reinterpret_cast<const float4 *>(dv + DH + cchunk * 8);
            const float4 b1 = *reinterpret_cast<const float4 *>(dv + DH + cchunk * 8 + 4);
            const float4 w0 = *reinterpret_cast<const float4 *>(dv + 2 * DH + cchunk * 8);
            const float4 w1 = *reinterpret_cast<const float4 *>(dv + 2 * DH + cchunk * 8 + 4);
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float d2 = fmaxf(fmaf(__uint_as_float(va[q]) + __uint_as_float(vb[q]), inv_d2, bb[q]), 0.f);
                sum = fmaf(d2, ww[q], sum);
            }
        }
        dpart[half * DSLOTS + slot] = sum;
    }
    ++dph;
    tc_fence_before();
    __syncthreads();
    // (6) Dense(1) bias, nan_to_num (keras_model.py:77), ensemble mean (ensemble.py:24)
    for (int sl = tid; sl < g_slots; sl += NT) {
        const float y = fxd::nan_to_num(dpart[sl] + dpart[DSLOTS + sl] + bd3v);
        const long long seq = seq_of(sl);
        float tot = (mem == 0) ? y : out[seq] + y;
        if (M > 1 && mem == M - 1) tot = tot / (float)M;
        out[seq] = tot;
    }
    __syncthreads();
}

}  // namespace u2
