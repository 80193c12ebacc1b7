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

// ---- packed residues: the wire format of the host boundary -------------------------------------------------------
// A sequence travels over PCIe as a little-endian bit stream of ceil(log2 A) bits per residue (2 bits for DNA/RNA, 5 for
// proteins), rows padded to whole bytes: 25 bytes instead of 100 for a 100-mer.  unpack_kernel turns the rows back into
// one byte per residue for the forward kernels (HBM-bound byte work: reads bits/8, writes 1 byte per residue; values
// >= A — impossible for a row packed from the alphabet — are reported through `status` like bad characters are).
__global__ void __launch_bounds__(256) unpack_kernel(const uint8_t *__restrict__ packed, int64_t n, int L, int A, int bits,
                                                     int row_bytes, uint8_t *__restrict__ idx, int64_t *__restrict__ status) {
    const int64_t total = n * (int64_t)L;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 16;
    const uint32_t mask = (1u << bits) - 1u;
    int64_t bad_count = 0, bad_first = 0x7fffffffffffffffll;
    const bool vec = (reinterpret_cast<uintptr_t>(idx) & 15) == 0;
    for (int64_t p0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16; p0 < total; p0 += stride) {
        int64_t row = p0 / L;
        int i = (int)(p0 - row * L);
        const uint8_t *src = packed + row * row_bytes;
        const int cnt = (int)min((int64_t)16, total - p0);
        uint32_t o[4] = {0, 0, 0, 0};
        for (int k = 0; k < cnt; ++k) {
            const int bit = i * bits, byte = bit >> 3, sh = bit & 7;
            uint32_t w = src[byte];
            if (sh + bits > 8) w |= (uint32_t)src[byte + 1] << 8;
            uint32_t code = (w >> sh) & mask;
            if (code >= (uint32_t)A) {
                bad_count++;
                bad_first = min(bad_first, p0 + k);
                code = 0;
            }
            o[k >> 2] |= code << (8 * (k & 3));
            if (++i == L) { i = 0; src += row_bytes; }
        }
        if (vec && cnt == 16) *reinterpret_cast<uint4 *>(idx + p0) = make_uint4(o[0], o[1], o[2], o[3]);
        else for (int k = 0; k < cnt; ++k) idx[p0 + k] = (uint8_t)(o[k >> 2] >> (8 * (k & 3)));
    }
    if (bad_count) {
        atomicAdd(reinterpret_cast<unsigned long long *>(status), (unsigned long long)bad_count);
        atomicMin(reinterpret_cast<long long *>(status + 1), (long long)bad_first);
    }
}

// residue indices -> packed rows (device-generated candidates on their way to the host); one thread per output byte
__global__ void __launch_bounds__(256) pack_kernel(const uint8_t *__restrict__ idx, int64_t n, int L, int bits, int row_bytes,
                                                   uint8_t *__restrict__ packed) {
    const int64_t total = n * (int64_t)row_bytes;
    const uint32_t mask = (1u << bits) - 1u;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = q / row_bytes;
        const int byte = (int)(q - row * row_bytes);
        const uint8_t *src = idx + row * L;
        const int first = (byte * 8) / bits, last = min(L - 1, (byte * 8 + 7) / bits);
        uint32_t out = 0;
        for (int i = first; i <= last; ++i) {
            const int rel = i * bits - byte * 8;  // position of residue i's bit 0 relative to this byte
            const uint32_t v = (uint32_t)src[i] & mask;
            out |= rel >= 0 ? (v << rel) : (v >> (-rel));
        }
        packed[q] = (uint8_t)out;
    }
}

}  // namespace

namespace fx {

int bits_per_residue(int a) {
    int bits = 1;
    while ((1 << bits) < a) ++bits;
    return bits;
}

static int grid_for(int64_t work_items) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return (int)std::max<int64_t>(1, std::min<int64_t>((work_items + 255) / 256, (int64_t)sms * 8));
}

int launch_unpack(const uint8_t *d_packed, int64_t n, int L, int a, uint8_t *d_idx, int64_t *d_status, cudaStream_t s) {
    const int bits = bits_per_residue(a), row_bytes = (L * bits + 7) / 8;
    encode_init_kernel<<<1, 1, 0, s>>>(d_status);
    if (n > 0) unpack_kernel<<<grid_for((n * L + 15) / 16), 256, 0, s>>>(d_packed, n, L, a, bits, row_bytes, d_idx, d_status);
    FX_CUDA(cudaGetLastError());
    return FLEXS_OK;
}

int launch_pack(const uint8_t *d_idx, int64_t n, int L, int a, uint8_t *d_packed, cudaStream_t s) {
    const int bits = bits_per_residue(a), row_bytes = (L * bits + 7) / 8;
    if (n > 0) pack_kernel<<<grid_for(n * row_bytes), 256, 0, s>>>(d_idx, n, L, bits, row_bytes, d_packed);
    FX_CUDA(cudaGetLastError());
    return FLEXS_OK;
}

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
