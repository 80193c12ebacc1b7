// K0 encode: residue characters -> residue indices (alphabet.index(ch)).
// Replaces string_to_one_hot (flexs/utils/sequence_utils.py:32-47); the float one-hot the
// reference builds (8*L*A bytes per sequence) is never materialised: downstream kernels
// gather weight rows by index instead.  HBM-bound byte work: 1 B read + 1 B written per
// residue, 16-byte vector accesses, grid sized in multiples of the SM count.
#include <algorithm>

#include "common.cuh"

namespace {

struct Alphabet {
    unsigned char ch[256];
    int n;
};

__global__ void encode_init_kernel(int64_t *status) {
    status[0] = 0;
    status[1] = 0x7fffffffffffffffll;
}

__global__ void __launch_bounds__(256) encode_kernel(const uint8_t *__restrict__ chars, int64_t n_bytes,
                                                     Alphabet alpha, uint8_t *__restrict__ idx,
                                                     int64_t *__restrict__ status) {
    __shared__ uint8_t lut[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = 0xff;
    __syncthreads();
    // first index wins, like str.index
    if (threadIdx.x == 0)
        for (int i = alpha.n - 1; i >= 0; --i) lut[alpha.ch[i]] = (uint8_t)i;
    __syncthreads();

    int64_t bad_count = 0, bad_first = 0x7fffffffffffffffll;
    // the vector body needs both pointers 16-byte aligned; otherwise everything is "tail"
    const bool aligned = ((reinterpret_cast<uintptr_t>(chars) | reinterpret_cast<uintptr_t>(idx)) & 15) == 0;
    const int64_t nvec = aligned ? n_bytes / 16 : 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint4 *src = reinterpret_cast<const uint4 *>(chars);
    uint4 *dst = reinterpret_cast<uint4 *>(idx);
    for (int64_t v = gtid; v < nvec; v += stride) {
        uint4 in = __ldg(src + v);
        uint32_t w[4] = {in.x, in.y, in.z, in.w};
        uint32_t o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint32_t r = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                uint32_t code = lut[(w[i] >> (8 * b)) & 0xff];
                if (code == 0xff) {
                    bad_count++;
                    int64_t pos = v * 16 + i * 4 + b;
                    bad_first = pos < bad_first ? pos : bad_first;
                    code = 0;
                }
                r |= code << (8 * b);
            }
            o[i] = r;
        }
        dst[v] = make_uint4(o[0], o[1], o[2], o[3]);
    }
    for (int64_t p = nvec * 16 + gtid; p < n_bytes; p += stride) {
        uint32_t code = lut[chars[p]];
        if (code == 0xff) {
            bad_count++;
            bad_first = p < bad_first ? p : bad_first;
            code = 0;
        }
        idx[p] = (uint8_t)code;
    }
    if (bad_count) {
        atomicAdd(reinterpret_cast<unsigned long long *>(status), (unsigned long long)bad_count);
        atomicMin(reinterpret_cast<long long *>(status + 1), (long long)bad_first);
    }
}

}  // namespace

namespace fx {

int launch_encode(const uint8_t *d_chars, int64_t n_bytes, const char *alphabet, int a,
                  uint8_t *d_idx, int64_t *d_status, cudaStream_t s) {
    Alphabet alpha;
    alpha.n = a;
    for (int i = 0; i < a; ++i) alpha.ch[i] = (unsigned char)alphabet[i];
    encode_init_kernel<<<1, 1, 0, s>>>(d_status);
    if (n_bytes > 0) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        int64_t want = (n_bytes / 16 + 255) / 256;
        int64_t grid = std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)sms * 8));
        encode_kernel<<<(unsigned)grid, 256, 0, s>>>(d_chars, n_bytes, alpha, d_idx, d_status);
    }
    FX_CUDA(cudaGetLastError());
    return FLEXS_OK;
}

}  // namespace fx
