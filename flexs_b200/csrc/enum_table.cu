// K7: whole-model table for tiny sequence spaces (A^L <= 2^20: 8-mers and 10-mers over 4 letters, the TF-binding
// configurations).  A surrogate is a pure function of the sequence, and a virtual screen of a million candidates over
// a space of 4^8 = 65 536 sequences scores every sequence ~16 times: evaluate the model ONCE on all A^L sequences with
// the regular fused kernels (any kind: CNN, MLP, ensembles), then a screen is one gather per candidate — HBM-bound,
// L bytes in and 4 bytes out.  The reference's explorers memoise scores in Python dicts for the same reason
// (adalead.py:157, cmaes.py:85-90).  The table is rebuilt when the weights change.
#include <cstdint>

#include "common.cuh"

namespace {

constexpr int64_t ENUM_MAX = 1 << 20;

__global__ void enum_fill_kernel(uint8_t *idx, int64_t total, int L, int A) {
    // sequence number i = big-endian base-A number formed by its residues
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t v = i;
        for (int p = L - 1; p >= 0; --p) {
            idx[i * L + p] = (uint8_t)(v % A);
            v /= A;
        }
    }
}

__global__ void enum_lookup_kernel(const uint8_t *__restrict__ idx, int64_t n, int L, int A,
                                   const float *__restrict__ table, int64_t total, float *__restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint8_t *s = idx + i * L;
        int64_t code = 0;
        for (int p = 0; p < L; ++p) code = code * A + min((int)s[p], A - 1);  // a residue >= A cannot leave the table
        out[i] = __ldg(table + code);
    }
}

}  // namespace

namespace fx {

int64_t enum_space(const flexs_model *m) {
    int64_t total = 1;
    for (int p = 0; p < m->L; ++p) {
        total *= m->A;
        if (total > ENUM_MAX) return 0;
    }
    return total;
}

int launch_enum(flexs_model *m, const uint8_t *d_idx, int64_t n, float *d_out, cudaStream_t s) {
    const int64_t total = enum_space(m);
    FX_REQUIRE(total > 0, "sequence space too large for a whole-model table (A^L must be <= 2^20)");
    FX_CUDA(cudaSetDevice(m->device));
    if (!m->enum_ready) {
        if (!m->d_enum_tab) FX_CUDA(cudaMalloc(&m->d_enum_tab, sizeof(float) * total));
        uint8_t *d_all = nullptr;
        FX_CUDA(cudaMalloc(&d_all, (size_t)total * m->L + 16));
        enum_fill_kernel<<<(int)std::min<int64_t>((total + 255) / 256, 1184), 256, 0, s>>>(d_all, total, m->L, m->A);
        FX_CUDA(cudaGetLastError());
        m->launches += 1;
        int rc = forward_direct(m, d_all, total, m->d_enum_tab, s);
        // the model may be driven from several streams: the table must be complete before any of them reads it
        cudaError_t e = cudaStreamSynchronize(s);
        cudaFree(d_all);
        if (rc != FLEXS_OK) return rc;
        if (e != cudaSuccess) return cuda_fail(e, "cudaStreamSynchronize");
        m->enum_ready = true;
    }
    enum_lookup_kernel<<<(int)std::min<int64_t>((n + 255) / 256, 1184), 256, 0, s>>>(d_idx, n, m->L, m->A, m->d_enum_tab, total, d_out);
    FX_CUDA(cudaGetLastError());
    m->launches += 1;
    return FLEXS_OK;
}

}  // namespace fx
