// Shared declarations for the flexs_b200 native library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/flexs_b200.h"

namespace fx {

void set_error(const std::string &msg);
int cuda_fail(cudaError_t e, const char *what);

#define FX_CUDA(call)                                            \
    do {                                                         \
        cudaError_t e__ = (call);                                \
        if (e__ != cudaSuccess) return fx::cuda_fail(e__, #call); \
    } while (0)

#define FX_REQUIRE(cond, msg)           \
    do {                                \
        if (!(cond)) {                  \
            fx::set_error(msg);         \
            return FLEXS_EINVAL;        \
        }                               \
    } while (0)

// Geometry of one CNN member (cnn.py:23-54) in the notation of DESIGN.md.
struct CnnDims {
    int L, A, F, H, K, K3, T;
    int pl2, pr2, pl3, pr3;  // "same" padding of conv2 / conv3 (TF rule: left=(k-1)/2)
};

// Offsets (in floats) of the 12 Keras arrays inside one member's weight block.
struct CnnOffsets {
    int64_t w1, b1, w2, b2, w3, b3, wd1, bd1, wd2, bd2, wd3, bd3, total;
};

struct MlpOffsets {
    int64_t w1, b1, w2, b2, w3, b3, w4, b4, total;
};

}  // namespace fx

// Opaque model object behind the C ABI.
struct flexs_model {
    int kind = 0, device = 0;
    int L = 0, A = 0, F = 0, H = 0, K = 0, K3 = 0, M = 1;
    int variant = FLEXS_VARIANT_AUTO;
    int sm_count = 148;
    int max_smem_optin = 0;
    int64_t launches = 0;

    std::vector<int64_t> arr_sizes;  // per-member array element counts, Keras order
    std::vector<int64_t> arr_offs;   // prefix offsets (floats) inside a member block
    int64_t member_floats = 0;
    float *d_weights = nullptr;      // [M][member_floats], Keras layout, fp32

    // derived operand layouts (rebuilt by set_weights when a variant needs them)
    void *d_umma_w = nullptr;        // fp16 hi/lo canonical-layout conv weights, all members
    bool umma_ready = false;
    void *d_umma2_w = nullptr;       // same for cnn_umma2.cu ([hi|lo] fused N=64 planes + gather tables)
    bool umma2_ready = false;
    bool umma_weights_ok = true;     // false: non-finite conv weights, use the fp32 kernel
    void *d_k9_tab = nullptr;        // cnn_k9.cu: conv1 o conv2 as a table over 9 residues, [M][425984][128 B]
    int *d_k9_ovf = nullptr;         // raised by the table builder when an entry left the fp16 window
    bool k9_ready = false;
    int k9_pair_units = -1;   // CTA pairs of cnn_k9_pair_kernel the device can hold at once (-1: not asked yet, 0: none)
    float *d_enum_tab = nullptr;     // enum_table.cu: scores of all A^L sequences (A^L <= 2^20)
    bool enum_ready = false;
    // Per-stream scratch of the tcgen05 kernels: the fp16-overflow flag (raised by a kernel, read by the gated FP32
    // re-computation enqueued behind it on the same stream) and the pooled-feature tiles that travel from a conv kernel
    // to the dense-head kernel.  One per stream in use, so concurrent chunks of score_host never share a flag.
    struct StreamWs { cudaStream_t stream; void *ptr; size_t bytes; int *flag; };
    std::vector<StreamWs> stream_ws;
    void *d_a20_w = nullptr;         // cnn_a20.cu: operand blob (dense planes, conv planes, conv1 gather tables), all members
    bool a20_ready = false;
    int a20_pair_units = -1;  // the same for cnn_a20_pair_kernel
    void *d_mlp_w = nullptr;         // mlp_umma.cu: operand blob, all members
    bool mlp_ready = false;

    // Adam state for K4 (same layout as d_weights) and the 1-based step counter per member
    float *d_adam_m = nullptr, *d_adam_v = nullptr;
    std::vector<int64_t> adam_step;
    void *train_ws = nullptr;        // training workspace (activations + grads)
    int64_t train_ws_bytes = 0;

    // score_host staging: slots of pinned host + device buffers, one stream each
    static constexpr int NSLOT = 3;  // three chunks in flight: copy in, compute, copy out, with slack for jitter
    cudaStream_t streams[NSLOT] = {};
    cudaEvent_t slot_done[NSLOT] = {};
    uint8_t *h_pin_chars[NSLOT] = {};
    float *h_pin_out[NSLOT] = {};
    uint8_t *d_chars[NSLOT] = {};
    uint8_t *d_idx[NSLOT] = {};
    float *d_out[NSLOT] = {};
    int64_t *d_status = nullptr;     // [NSLOT][2]
    int64_t *h_status = nullptr;     // pinned mirror
    int64_t host_chunk = 0;          // sequences per staging slot
};

namespace fx {

CnnDims cnn_dims(const flexs_model *m);
CnnOffsets cnn_offsets(const flexs_model *m);
MlpOffsets mlp_offsets(const flexs_model *m);

// kernels' host launchers (each returns a FLEXS_* code and bumps m->launches)
int launch_encode(const uint8_t *d_chars, int64_t n_bytes, const char *alphabet, int a,
                  uint8_t *d_idx, int64_t *d_status, cudaStream_t s);
int bits_per_residue(int alphabet_size);
int launch_unpack(const uint8_t *d_packed, int64_t n, int L, int a, uint8_t *d_idx, int64_t *d_status, cudaStream_t s);
int launch_pack(const uint8_t *d_idx, int64_t n, int L, int a, uint8_t *d_packed, cudaStream_t s);
int launch_cnn_simple(flexs_model *m, const uint8_t *d_idx, int64_t n, float *d_out, cudaStream_t s);
int launch_cnn_tiled(flexs_model *m, const uint8_t *d_idx, int64_t n, float *d_out, cudaStream_t s);
// same kernel, but every CTA returns immediately unless *d_gate != 0 (fp16-overflow fall-back)
int launch_cnn_tiled_gated(flexs_model *m, const uint8_t *d_idx, int64_t n, float *d_out, const int *d_gate,
                           cudaStream_t s);
bool cnn_tiled_supported(const flexs_model *m);
int launch_cnn_umma(flexs_model *m, const uint8_t *d_idx, int64_t n, float *d_out, cudaStream_t s);
bool cnn_umma_supported(const flexs_model *m);
int prepare_cnn_umma(flexs_model *m);
// pipelined tcgen05 kernel for the A = 4 shapes (cnn_umma2.cu)
bool cnn_umma2_supported(const flexs_model *m);
int launch_cnn_umma2(flexs_model *m, const uint8_t *d_idx, int64_t n, float *d_out, cudaStream_t s);
// table (conv1 o conv2 over 9 residues, L2-resident) + tcgen05 conv3/dense kernel for large A = 4 batches (cnn_k9.cu)
bool cnn_k9_supported(const flexs_model *m);
int launch_cnn_k9(flexs_model *m, const uint8_t *d_idx, int64_t n, float *d_out, cudaStream_t s);
int prepare_cnn_umma2(flexs_model *m);
// AUTO picks the table kernel from this batch size on (building the table costs about as much as scoring
// 3e4 sequences with cnn_umma2), or, once the table of the current weights exists, from the size at which its
// 128-sequence groups occupy enough SMs to beat cnn_umma2's finer-grained items
constexpr int64_t K9_MIN_N = 65536, K9_MIN_N_READY = 8192;
// whole-model table over all A^L sequences (enum_table.cu); enum_space() is 0 when A^L > 2^20
int64_t enum_space(const flexs_model *m);
int launch_enum(flexs_model *m, const uint8_t *d_idx, int64_t n, float *d_out, cudaStream_t s);
// the fused kernels, without the whole-model table in front of them
int forward_direct(flexs_model *m, const uint8_t *d_idx, int64_t n, float *d_out, cudaStream_t s);
int launch_mlp(flexs_model *m, const uint8_t *d_idx, int64_t n, float *d_out, cudaStream_t s);
int launch_mlp_gated(flexs_model *m, const uint8_t *d_idx, int64_t n, float *d_out, const int *d_gate, cudaStream_t s);
// every MLP layer on tcgen05: one-hot x [W_hi | W_lo] streamed from L2 for layer 1, K = 112 GEMMs for layers 2-3 (mlp_umma.cu)
bool mlp_umma_supported(const flexs_model *m);
int launch_mlp_umma(flexs_model *m, const uint8_t *d_idx, int64_t n, float *d_out, cudaStream_t s);
// k3 = 19 (A = 20) shapes: stream-interleaved rows, rolling activation rings, conv2 + conv3 on tcgen05 (cnn_a20.cu)
bool cnn_a20_supported(const flexs_model *m);
int launch_cnn_a20(flexs_model *m, const uint8_t *d_idx, int64_t n, float *d_out, cudaStream_t s);
// per-stream scratch (created on first use; grown to `bytes` of feature space, zero-filled when (re)allocated)
int stream_workspace(flexs_model *m, cudaStream_t s, size_t bytes, flexs_model::StreamWs **out);
// dense head over [32][128] pooled-feature tiles (cnn_k9.cu): Dense(H,relu) x2 -> Dense(1) -> nan_to_num -> ensemble mean;
// `uw` is a member blob with the u2 dense offsets (OFF_DB1 / OFF_DB2 / OFF_DV) filled in
int launch_dense_tiles(flexs_model *m, const float *feat, float *out, const unsigned char *uw, int *flag, int64_t n,
                       int mem, cudaStream_t s);

}  // namespace fx

// ---- small device helpers shared by the kernels -------------------------------------------
#ifdef __CUDACC__
namespace fxd {

// np.nan_to_num on float32 (keras_model.py:77): NaN -> 0, +-inf -> +-FLT_MAX
__device__ __forceinline__ float nan_to_num(float x) {
    if (x != x) return 0.f;
    if (x > 3.4028234663852886e38f) return 3.4028234663852886e38f;
    if (x < -3.4028234663852886e38f) return -3.4028234663852886e38f;
    return x;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP) ----------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)
                 : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    // the suspend-time hint lets the hardware park the warp instead of spinning through the
    // issue slots the working warps need (first v2 profile: 19 % of all instructions were spins)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    // Slow path with a watchdog: no wait of these kernels lasts longer than a launch (milliseconds).  A protocol bug
    // must surface as a launch failure the caller sees, not as a GPU that never comes back.  The timer is read once
    // per 64 failed attempts so the loop costs the working warps no more issue slots than a bare retry.
    unsigned long long t0 = 0, t1;
    unsigned int spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 63u) == 0) {
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t0 == 0) t0 = t1;
            else if (t1 - t0 > 4000000000ull) __trap();
        }
    }
}
// The same wait for a warp that must stay CONVERGED afterwards (the MMA-issuing warps): the loop exits on a vote, a
// warp-uniform condition, so the compiler keeps the code behind it on the uniform datapath (descriptor arithmetic in
// uniform registers, no per-MMA R2UR moves under an election predicate).
__device__ __forceinline__ void mbar_wait_warp(uint64_t *bar, uint32_t parity) {
    if (__all_sync(0xffffffffu, mbar_try_wait(bar, parity))) return;
    unsigned long long t0 = 0, t1;
    unsigned int spins = 0;
    while (!__all_sync(0xffffffffu, mbar_try_wait(bar, parity))) {
        if ((++spins & 63u) == 0) {
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t0 == 0) t0 = t1;
            else if (t1 - t0 > 4000000000ull) __trap();
        }
    }
}
// global -> shared bulk copy; dst/src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                         uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

}  // namespace fxd
#endif
