// C ABI of flexs_b200 (include/flexs_b200.h): model objects, weight I/O, dispatch and the
// host-buffer scoring call.  No CPU compute path exists in this file or behind it.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>

#include "common.cuh"

namespace fx {

static thread_local std::string g_last_error;

void set_error(const std::string &msg) { g_last_error = msg; }

int cuda_fail(cudaError_t e, const char *what) {
    g_last_error = std::string("CUDA error ") + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) +
                   ") in " + what;
    cudaGetLastError();  // clear sticky-less errors
    return FLEXS_ECUDA;
}

CnnDims cnn_dims(const flexs_model *m) {
    CnnDims d;
    d.L = m->L; d.A = m->A; d.F = m->F; d.H = m->H; d.K = m->K; d.K3 = m->K3;
    d.T = m->L - m->K + 1;
    d.pl2 = (m->K - 1) / 2;  d.pr2 = (m->K - 1) - d.pl2;
    d.pl3 = (m->K3 - 1) / 2; d.pr3 = (m->K3 - 1) - d.pl3;
    return d;
}

CnnOffsets cnn_offsets(const flexs_model *m) {
    CnnOffsets o;
    const std::vector<int64_t> &p = m->arr_offs;
    o.w1 = p[0]; o.b1 = p[1]; o.w2 = p[2]; o.b2 = p[3]; o.w3 = p[4]; o.b3 = p[5];
    o.wd1 = p[6]; o.bd1 = p[7]; o.wd2 = p[8]; o.bd2 = p[9]; o.wd3 = p[10]; o.bd3 = p[11];
    o.total = m->member_floats;
    return o;
}

MlpOffsets mlp_offsets(const flexs_model *m) {
    MlpOffsets o;
    const std::vector<int64_t> &p = m->arr_offs;
    o.w1 = p[0]; o.b1 = p[1]; o.w2 = p[2]; o.b2 = p[3]; o.w3 = p[4]; o.b3 = p[5];
    o.w4 = p[6]; o.b4 = p[7];
    o.total = m->member_floats;
    return o;
}

static int finish_create(flexs_model *m) {
    int64_t off = 0;
    m->arr_offs.clear();
    for (int64_t s : m->arr_sizes) {
        m->arr_offs.push_back(off);
        off += (s + 3) / 4 * 4;  // keep every array 16-byte aligned inside the block
    }
    m->member_floats = off;
    int ndev = 0;
    FX_CUDA(cudaGetDeviceCount(&ndev));
    FX_REQUIRE(m->device >= 0 && m->device < ndev, "device index out of range");
    FX_CUDA(cudaSetDevice(m->device));
    FX_CUDA(cudaDeviceGetAttribute(&m->sm_count, cudaDevAttrMultiProcessorCount, m->device));
    FX_CUDA(cudaDeviceGetAttribute(&m->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin,
                                   m->device));
    size_t bytes = sizeof(float) * m->member_floats * m->M;
    FX_CUDA(cudaMalloc(&m->d_weights, bytes));
    FX_CUDA(cudaMemset(m->d_weights, 0, bytes));
    m->adam_step.assign(m->M, 0);
    return FLEXS_OK;
}

}  // namespace fx

using namespace fx;

extern "C" {

int flexs_abi_version(void) { return FLEXS_B200_ABI_VERSION; }

const char *flexs_last_error(void) { return g_last_error.c_str(); }

int flexs_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount");
    return n;
}

int flexs_cnn_create(int device, int seq_len, int alphabet_size, int num_filters, int hidden_size,
                     int kernel_size, int n_members, flexs_model_t **out) {
    FX_REQUIRE(out != nullptr, "out is null");
    *out = nullptr;
    FX_REQUIRE(alphabet_size >= 2 && alphabet_size <= 255, "alphabet_size must be in [2,255]");
    FX_REQUIRE(kernel_size >= 1 && seq_len >= kernel_size,
               "seq_len must be >= kernel_size (valid conv, cnn.py:25-32)");
    FX_REQUIRE(num_filters >= 1 && hidden_size >= 1 && n_members >= 1, "bad sizes");
    flexs_model *m = new flexs_model();
    m->kind = FLEXS_KIND_CNN; m->device = device;
    m->L = seq_len; m->A = alphabet_size; m->F = num_filters; m->H = hidden_size;
    m->K = kernel_size; m->K3 = alphabet_size - 1; m->M = n_members;
    const int64_t k = m->K, a = m->A, f = m->F, h = m->H, k3 = m->K3;
    m->arr_sizes = {k * a * f, f, k * f * f, f, k3 * f * f, f, f * h, h, h * h, h, h, 1};
    int rc = finish_create(m);
    if (rc != FLEXS_OK) { flexs_model_destroy(m); return rc; }
    *out = m;
    return FLEXS_OK;
}

int flexs_mlp_create(int device, int seq_len, int alphabet_size, int hidden_size, int n_members,
                     flexs_model_t **out) {
    FX_REQUIRE(out != nullptr, "out is null");
    *out = nullptr;
    FX_REQUIRE(alphabet_size >= 2 && alphabet_size <= 255, "alphabet_size must be in [2,255]");
    FX_REQUIRE(seq_len >= 1 && hidden_size >= 1 && n_members >= 1, "bad sizes");
    flexs_model *m = new flexs_model();
    m->kind = FLEXS_KIND_MLP; m->device = device;
    m->L = seq_len; m->A = alphabet_size; m->H = hidden_size; m->M = n_members;
    const int64_t d = (int64_t)m->L * m->A, h = m->H;
    m->arr_sizes = {d * h, h, h * h, h, h * h, h, h, 1};
    int rc = finish_create(m);
    if (rc != FLEXS_OK) { flexs_model_destroy(m); return rc; }
    *out = m;
    return FLEXS_OK;
}

void flexs_model_destroy(flexs_model_t *m) {
    if (!m) return;
    cudaSetDevice(m->device);
    cudaFree(m->d_weights);
    cudaFree(m->d_umma_w);
    cudaFree(m->d_umma2_w);
    cudaFree(m->d_k9_tab);
    cudaFree(m->d_k9_ovf);
    for (auto &w : m->stream_ws) { cudaFree(w.ptr); cudaFree(w.flag); }
    cudaFree(m->d_a20_w);
    cudaFree(m->d_mlp_w);
    cudaFree(m->d_enum_tab);
    cudaFree(m->d_adam_m);
    cudaFree(m->d_adam_v);
    cudaFree(m->train_ws);
    for (int i = 0; i < flexs_model::NSLOT; ++i) {
        if (m->streams[i]) cudaStreamDestroy(m->streams[i]);
        if (m->slot_done[i]) cudaEventDestroy(m->slot_done[i]);
        cudaFreeHost(m->h_pin_chars[i]);
        cudaFreeHost(m->h_pin_out[i]);
        cudaFree(m->d_chars[i]);
        cudaFree(m->d_idx[i]);
        cudaFree(m->d_out[i]);
    }
    cudaFree(m->d_status);
    cudaFreeHost(m->h_status);
    delete m;
}

int flexs_model_num_arrays(const flexs_model_t *m) { return m ? (int)m->arr_sizes.size() : FLEXS_EINVAL; }

int64_t flexs_model_array_size(const flexs_model_t *m, int i) {
    if (!m || i < 0 || i >= (int)m->arr_sizes.size()) return FLEXS_EINVAL;
    return m->arr_sizes[i];
}

int flexs_model_set_weights(flexs_model_t *m, int member, const float *const *h_arrays) {
    FX_REQUIRE(m && h_arrays, "null argument");
    FX_REQUIRE(member >= 0 && member < m->M, "member out of range");
    FX_CUDA(cudaSetDevice(m->device));
    float *base = m->d_weights + (int64_t)member * m->member_floats;
    for (size_t i = 0; i < m->arr_sizes.size(); ++i) {
        FX_REQUIRE(h_arrays[i] != nullptr, "null weight array");
        FX_CUDA(cudaMemcpy(base + m->arr_offs[i], h_arrays[i], sizeof(float) * m->arr_sizes[i],
                           cudaMemcpyHostToDevice));
    }
    m->umma_ready = false;
    m->umma2_ready = false;
    m->k9_ready = false;
    m->enum_ready = false;
    m->a20_ready = false;
    m->mlp_ready = false;
    return FLEXS_OK;
}

int flexs_model_get_weights(flexs_model_t *m, int member, float *const *h_arrays) {
    FX_REQUIRE(m && h_arrays, "null argument");
    FX_REQUIRE(member >= 0 && member < m->M, "member out of range");
    FX_CUDA(cudaSetDevice(m->device));
    const float *base = m->d_weights + (int64_t)member * m->member_floats;
    for (size_t i = 0; i < m->arr_sizes.size(); ++i) {
        FX_REQUIRE(h_arrays[i] != nullptr, "null weight array");
        FX_CUDA(cudaMemcpy(h_arrays[i], base + m->arr_offs[i], sizeof(float) * m->arr_sizes[i],
                           cudaMemcpyDeviceToHost));
    }
    return FLEXS_OK;
}

int flexs_model_set_variant(flexs_model_t *m, int variant) {
    FX_REQUIRE(m, "null model");
    FX_REQUIRE(variant >= FLEXS_VARIANT_AUTO && variant <= FLEXS_VARIANT_ENUM, "unknown variant");
    if (variant == FLEXS_VARIANT_ENUM) FX_REQUIRE(enum_space(m) > 0, "ENUM variant needs A^L <= 2^20");
    if (m->kind == FLEXS_KIND_CNN) {
        if (variant == FLEXS_VARIANT_TILED) FX_REQUIRE(cnn_tiled_supported(m), "TILED variant needs F=32, k=5, A in {4,20}");
        if (variant == FLEXS_VARIANT_UMMA) FX_REQUIRE(cnn_umma_supported(m), "UMMA variant not available for this shape");
        if (variant == FLEXS_VARIANT_UMMA_LUT) FX_REQUIRE(cnn_k9_supported(m), "UMMA_LUT variant needs A=4, F=32, k=5, H<=112, 8 <= L <= ~175");
    } else {
        FX_REQUIRE(variant == FLEXS_VARIANT_AUTO || variant == FLEXS_VARIANT_ENUM || variant == FLEXS_VARIANT_TILED ||
                   variant == FLEXS_VARIANT_UMMA, "MLP kernels: TILED (FP32 FFMA), UMMA (tcgen05), ENUM (whole-model table)");
        if (variant == FLEXS_VARIANT_UMMA) FX_REQUIRE(mlp_umma_supported(m), "UMMA variant of the MLP needs H <= 112 and L <= ~550");
    }
    m->variant = variant;
    return FLEXS_OK;
}

// variant of the fused kernels (never ENUM)
static int direct_variant(const flexs_model *m, int64_t n) {
    if (m->kind != FLEXS_KIND_CNN) {
        if (m->variant == FLEXS_VARIANT_TILED || m->variant == FLEXS_VARIANT_UMMA) return m->variant;
        return mlp_umma_supported(m) ? FLEXS_VARIANT_UMMA : FLEXS_VARIANT_TILED;
    }
    if (m->variant != FLEXS_VARIANT_AUTO && m->variant != FLEXS_VARIANT_ENUM) return m->variant;
    if (cnn_k9_supported(m) && n >= (m->k9_ready ? K9_MIN_N_READY : K9_MIN_N)) return FLEXS_VARIANT_UMMA_LUT;
    if (cnn_umma_supported(m)) return FLEXS_VARIANT_UMMA;
    if (cnn_tiled_supported(m)) return FLEXS_VARIANT_TILED;
    return FLEXS_VARIANT_SIMPLE;
}

int flexs_model_active_variant(const flexs_model_t *m, int64_t n) {
    if (!m) return FLEXS_EINVAL;
    if (m->variant == FLEXS_VARIANT_ENUM) return FLEXS_VARIANT_ENUM;
    // AUTO: a batch at least as large as the sequence space pays for scoring the whole space once
    const int64_t space = enum_space(m);
    if (m->variant == FLEXS_VARIANT_AUTO && space > 0 && (m->enum_ready || n >= space)) return FLEXS_VARIANT_ENUM;
    return direct_variant(m, n);
}

int64_t flexs_model_launch_count(const flexs_model_t *m) { return m ? m->launches : FLEXS_EINVAL; }

int flexs_encode_dev(const uint8_t *d_chars, int64_t n_bytes, const char *alphabet,
                     int alphabet_size, uint8_t *d_idx, int64_t *d_status, void *stream) {
    FX_REQUIRE(alphabet && alphabet_size >= 1 && alphabet_size <= 255, "bad alphabet");
    FX_REQUIRE(n_bytes >= 0, "negative size");
    FX_REQUIRE(d_status != nullptr, "d_status is null");
    FX_REQUIRE(n_bytes == 0 || (d_chars && d_idx), "null buffer");
    return launch_encode(d_chars, n_bytes, alphabet, alphabet_size, d_idx, d_status,
                         (cudaStream_t)stream);
}

int flexs_model_forward_dev(flexs_model_t *m, const uint8_t *d_idx, int64_t n, float *d_out,
                            void *stream) {
    FX_REQUIRE(m, "null model");
    FX_REQUIRE(n >= 0, "negative n");
    if (n == 0) return FLEXS_OK;
    FX_REQUIRE(d_idx && d_out, "null buffer");
    FX_CUDA(cudaSetDevice(m->device));  // weights, tables and workspaces live on the model's device
    cudaStream_t s = (cudaStream_t)stream;
    if (flexs_model_active_variant(m, n) == FLEXS_VARIANT_ENUM) return launch_enum(m, d_idx, n, d_out, s);
    return forward_direct(m, d_idx, n, d_out, s);
}

}  // extern "C"

namespace fx {

int stream_workspace(flexs_model *m, cudaStream_t s, size_t bytes, flexs_model::StreamWs **out) {
    flexs_model::StreamWs *ws = nullptr;
    for (auto &w : m->stream_ws) if (w.stream == s) ws = &w;
    if (!ws) {
        m->stream_ws.push_back({s, nullptr, 0, nullptr});
        ws = &m->stream_ws.back();
        FX_CUDA(cudaMalloc(&ws->flag, sizeof(int)));
    }
    if (ws->bytes < bytes) {
        FX_CUDA(cudaStreamSynchronize(s));
        if (ws->ptr) FX_CUDA(cudaFree(ws->ptr));
        ws->ptr = nullptr; ws->bytes = 0;
        FX_CUDA(cudaMalloc(&ws->ptr, bytes));
        FX_CUDA(cudaMemsetAsync(ws->ptr, 0, bytes, s));  // slots past the end of a batch are read (never reported)
        ws->bytes = bytes;
    }
    *out = ws;
    return FLEXS_OK;
}

int forward_direct(flexs_model *m, const uint8_t *d_idx, int64_t n, float *d_out, cudaStream_t s) {
    if (m->kind == FLEXS_KIND_MLP)
        return direct_variant(m, n) == FLEXS_VARIANT_UMMA ? launch_mlp_umma(m, d_idx, n, d_out, s) : launch_mlp(m, d_idx, n, d_out, s);
    switch (direct_variant(m, n)) {
        case FLEXS_VARIANT_UMMA_LUT: return launch_cnn_k9(m, d_idx, n, d_out, s);
        case FLEXS_VARIANT_UMMA: return launch_cnn_umma(m, d_idx, n, d_out, s);
        case FLEXS_VARIANT_TILED: return launch_cnn_tiled(m, d_idx, n, d_out, s);
        default: return launch_cnn_simple(m, d_idx, n, d_out, s);
    }
}

}  // namespace fx

extern "C" {

// Host-buffer scoring: chunk the batch, and for each chunk H2D(chars) -> encode -> forward ->
// D2H(scores), rotating over NSLOT slots/streams so copies overlap compute.
static int ensure_host_staging(flexs_model *m, int64_t n) {
    // A slot holds up to ~64 MiB of residue characters: large enough that launch overheads and the tail wave of the
    // persistent kernels stay below a few percent of a chunk, small enough that the first copy (which nothing
    // overlaps) is a small part of a multi-million-sequence call.  Slots grow on demand: a model that only ever
    // scores an explorer's 20-string batches pins a few KB, not 64 MiB.
    int64_t cap = (64ll << 20) / std::max(1, m->L);
    cap = std::max<int64_t>(1024, std::min<int64_t>(cap, 1 << 20)) / 128 * 128;
    const int64_t chunk = std::min(cap, std::max<int64_t>(1024, (n + 127) / 128 * 128));
    if (!m->streams[0]) {
        for (int i = 0; i < flexs_model::NSLOT; ++i) {
            FX_CUDA(cudaStreamCreateWithFlags(&m->streams[i], cudaStreamNonBlocking));
            FX_CUDA(cudaEventCreateWithFlags(&m->slot_done[i], cudaEventDisableTiming));
        }
        FX_CUDA(cudaMalloc(&m->d_status, 2 * flexs_model::NSLOT * sizeof(int64_t)));
        FX_CUDA(cudaMallocHost(&m->h_status, 2 * flexs_model::NSLOT * sizeof(int64_t)));
    }
    if (chunk <= m->host_chunk) return FLEXS_OK;
    for (int i = 0; i < flexs_model::NSLOT; ++i) {
        FX_CUDA(cudaStreamSynchronize(m->streams[i]));
        cudaFreeHost(m->h_pin_chars[i]); cudaFreeHost(m->h_pin_out[i]);
        cudaFree(m->d_chars[i]); cudaFree(m->d_idx[i]); cudaFree(m->d_out[i]);
        m->h_pin_chars[i] = nullptr; m->h_pin_out[i] = nullptr; m->d_chars[i] = nullptr; m->d_idx[i] = nullptr; m->d_out[i] = nullptr;
    }
    m->host_chunk = 0;
    for (int i = 0; i < flexs_model::NSLOT; ++i) {
        FX_CUDA(cudaMallocHost(&m->h_pin_chars[i], chunk * m->L));
        FX_CUDA(cudaMallocHost(&m->h_pin_out[i], chunk * sizeof(float)));
        FX_CUDA(cudaMalloc(&m->d_chars[i], chunk * m->L + 16));
        FX_CUDA(cudaMalloc(&m->d_idx[i], chunk * m->L + 16));
        FX_CUDA(cudaMalloc(&m->d_out[i], chunk * sizeof(float)));
    }
    m->host_chunk = chunk;
    return FLEXS_OK;
}

// `alphabet` != NULL: h_in holds n * L residue characters (flexs_model_score_host); NULL: n packed rows of
// ceil(L * bits / 8) bytes (flexs_model_score_host_packed).  Same pipeline, a different first kernel.
static int score_host_impl(flexs_model_t *m, const char *h_chars, int64_t n, const char *alphabet,
                           float *h_out, int64_t *bad_pos) {
    FX_REQUIRE(m, "null argument");
    FX_REQUIRE(n >= 0, "negative n");
    if (bad_pos) *bad_pos = -1;
    if (n == 0) return FLEXS_OK;
    FX_REQUIRE(h_chars && h_out, "null buffer");
    FX_CUDA(cudaSetDevice(m->device));
    int rc = ensure_host_staging(m, n);
    if (rc != FLEXS_OK) return rc;
    // One kernel choice for the whole call: AUTO is resolved on the full batch and held for every chunk (a short last
    // chunk would otherwise fall to the small-batch kernel and its scores would differ in the last bits), so scores
    // do not depend on where a sequence falls in the batch.
    struct VariantHold {
        flexs_model *m;
        int saved;
        ~VariantHold() { m->variant = saved; }
    } hold{m, m->variant};
    if (m->variant == FLEXS_VARIANT_AUTO) {
        const int v = flexs_model_active_variant(m, n);
        if (v == FLEXS_VARIANT_ENUM || v == FLEXS_VARIANT_UMMA_LUT) m->variant = v;
    }
    const int64_t L = m->L;
    const bool packed = alphabet == nullptr;
    const int64_t row_bytes = packed ? (L * bits_per_residue(m->A) + 7) / 8 : L;  // bytes per sequence on the wire
    // A chunk is a whole number of waves of the persistent kernels (sm_count groups of 128 sequences), close to
    // FLEXS_HOST_CHUNK_MB of residue characters: small enough that the first copy in and the last compute (which nothing
    // overlaps) are a small part of the call, large enough to amortise the per-chunk launches.
    static const int64_t target_mb = std::getenv("FLEXS_HOST_CHUNK_MB") ? std::atoll(std::getenv("FLEXS_HOST_CHUNK_MB")) : 64;
    const int64_t wave = (int64_t)m->sm_count * 128;
    int64_t per = std::max<int64_t>(1, (target_mb << 20) / std::max<int64_t>(1, L * wave)) * wave;
    per = std::min(per, m->host_chunk);
    const int64_t want = (n + per - 1) / per;
    const int64_t chunk = std::min<int64_t>(m->host_chunk, want <= 1 ? (n + 127) / 128 * 128 : per);
    // Buffers the caller already page-locked (cudaHostAlloc / cudaHostRegister / torch pin_memory) are used
    // directly by the async copies; pageable ones go through the model's pinned staging slots.
    auto is_pinned = [](const void *ptr) {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, ptr) != cudaSuccess) { cudaGetLastError(); return false; }
        return attr.type == cudaMemoryTypeHost;
    };
    const bool in_pinned = is_pinned(h_chars), out_pinned = is_pinned(h_out);
    // Chunk schedule: nothing overlaps the first copy in, so a multi-chunk call starts with a chunk of 1/8 of the full
    // size and doubles up to it; the full-size chunks amortise the per-chunk launches and the tail of the persistent kernels.
    std::vector<int64_t> sched;  // start offsets
    {
        int64_t pos = 0, cur = chunk;
        if (n > chunk) cur = std::max<int64_t>(wave, chunk / 8 / wave * wave);
        while (pos < n) {
            sched.push_back(pos);
            pos += std::min(cur, n - pos);
            cur = std::min(chunk, cur * 2);
        }
        sched.push_back(n);
    }
    const int64_t nchunks = (int64_t)sched.size() - 1;
    int64_t first_bad = std::numeric_limits<int64_t>::max();
    // slot bookkeeping: what is in flight in each slot
    int64_t inflight_start[flexs_model::NSLOT], inflight_cnt[flexs_model::NSLOT];
    for (int i = 0; i < flexs_model::NSLOT; ++i) { inflight_start[i] = -1; inflight_cnt[i] = 0; }
    auto drain = [&](int slot) -> int {
        if (inflight_start[slot] < 0) return FLEXS_OK;
        FX_CUDA(cudaEventSynchronize(m->slot_done[slot]));
        const int64_t *st = m->h_status + 2 * slot;
        if (st[0] != 0) first_bad = std::min(first_bad, inflight_start[slot] * L + st[1]);  // flat residue position
        if (!out_pinned) std::memcpy(h_out + inflight_start[slot], m->h_pin_out[slot], inflight_cnt[slot] * sizeof(float));
        inflight_start[slot] = -1;
        return FLEXS_OK;
    };
    // Whatever way this call ends, no copy may still be in flight on the caller's buffers (or on a slot the next call
    // reuses) when it returns: an early error return synchronises every slot stream first.
    struct DrainAll {
        flexs_model *m;
        ~DrainAll() { for (int i = 0; i < flexs_model::NSLOT; ++i) if (m->streams[i]) cudaStreamSynchronize(m->streams[i]); }
    } drain_all{m};
    for (int64_t c = 0; c < nchunks; ++c) {
        const int slot = (int)(c % flexs_model::NSLOT);
        rc = drain(slot);
        if (rc != FLEXS_OK) return rc;
        const int64_t start = sched[c], cnt = sched[c + 1] - start;
        cudaStream_t s = m->streams[slot];
        const void *src = h_chars + start * row_bytes;
        if (!in_pinned) { std::memcpy(m->h_pin_chars[slot], src, cnt * row_bytes); src = m->h_pin_chars[slot]; }
        FX_CUDA(cudaMemcpyAsync(m->d_chars[slot], src, cnt * row_bytes, cudaMemcpyHostToDevice, s));
        rc = packed ? launch_unpack(m->d_chars[slot], cnt, m->L, m->A, m->d_idx[slot], m->d_status + 2 * slot, s)
                    : launch_encode(m->d_chars[slot], cnt * L, alphabet, m->A, m->d_idx[slot], m->d_status + 2 * slot, s);
        if (rc != FLEXS_OK) return rc;
        m->launches += 1;
        rc = flexs_model_forward_dev(m, m->d_idx[slot], cnt, m->d_out[slot], s);
        if (rc != FLEXS_OK) return rc;
        FX_CUDA(cudaMemcpyAsync(out_pinned ? (void *)(h_out + start) : (void *)m->h_pin_out[slot], m->d_out[slot],
                                cnt * sizeof(float), cudaMemcpyDeviceToHost, s));
        FX_CUDA(cudaMemcpyAsync(m->h_status + 2 * slot, m->d_status + 2 * slot, 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
        FX_CUDA(cudaEventRecord(m->slot_done[slot], s));
        inflight_start[slot] = start; inflight_cnt[slot] = cnt;
    }
    for (int slot = 0; slot < flexs_model::NSLOT; ++slot) {
        rc = drain(slot);
        if (rc != FLEXS_OK) return rc;
    }
    if (first_bad != std::numeric_limits<int64_t>::max()) {
        if (bad_pos) *bad_pos = first_bad;
        set_error((packed ? "packed residue value outside the alphabet at flat position "
                          : "character outside the alphabet at flat position ") + std::to_string(first_bad));
        return FLEXS_EALPHABET;
    }
    return FLEXS_OK;
}

int flexs_model_score_host(flexs_model_t *m, const char *h_chars, int64_t n, const char *alphabet,
                           float *h_out, int64_t *bad_pos) {
    FX_REQUIRE(alphabet, "null alphabet");
    return score_host_impl(m, h_chars, n, alphabet, h_out, bad_pos);
}

int flexs_model_score_host_packed(flexs_model_t *m, const uint8_t *h_packed, int64_t n, float *h_out, int64_t *bad_pos) {
    return score_host_impl(m, reinterpret_cast<const char *>(h_packed), n, nullptr, h_out, bad_pos);
}

int flexs_bits_per_residue(int alphabet_size) {
    FX_REQUIRE(alphabet_size >= 2 && alphabet_size <= 255, "alphabet_size must be in [2,255]");
    return bits_per_residue(alphabet_size);
}

int64_t flexs_packed_row_bytes(int seq_len, int alphabet_size) {
    if (seq_len < 0 || alphabet_size < 2 || alphabet_size > 255) return FLEXS_EINVAL;
    return ((int64_t)seq_len * bits_per_residue(alphabet_size) + 7) / 8;
}

int flexs_unpack_dev(const uint8_t *d_packed, int64_t n, int seq_len, int alphabet_size, uint8_t *d_idx,
                     int64_t *d_status, void *stream) {
    FX_REQUIRE(alphabet_size >= 2 && alphabet_size <= 255 && seq_len >= 1 && n >= 0, "bad sizes");
    FX_REQUIRE(d_status != nullptr, "d_status is null");
    FX_REQUIRE(n == 0 || (d_packed && d_idx), "null buffer");
    return launch_unpack(d_packed, n, seq_len, alphabet_size, d_idx, d_status, (cudaStream_t)stream);
}

int flexs_pack_dev(const uint8_t *d_idx, int64_t n, int seq_len, int alphabet_size, uint8_t *d_packed, void *stream) {
    FX_REQUIRE(alphabet_size >= 2 && alphabet_size <= 255 && seq_len >= 1 && n >= 0, "bad sizes");
    FX_REQUIRE(n == 0 || (d_packed && d_idx), "null buffer");
    return launch_pack(d_idx, n, seq_len, alphabet_size, d_packed, (cudaStream_t)stream);
}

}  // extern "C"
