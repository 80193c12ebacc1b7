// K3d: the sharded screen's one exchange step as one-sided stores over NVLink peer memory (screen.py, overlap mode).
//
// The reference ranks one candidate list (adalead.py:171-175); sharded over G GPUs every rank needs the G per-shard top-k
// messages (select.cu).  An NCCL all_gather is a rendezvous: every rank's stream blocks until the slowest rank of that step
// has arrived, and at 8 GPUs the per-step wait (the GPUs of a box do not run a 7 ms forward equally fast) costs more than
// the 16 us the collective itself takes.  Here the exchange never blocks the sender: after its selection launch a rank
// WRITES its message into a slot of every peer's mailbox (cudaIpc-mapped peer memory, 16-byte stores over NVLink, then a
// release store of the step number at system scope); the merge of step i is issued one step later and spins (acquire loads,
// with a 20 s watchdog) only until the messages of step i are all in — by then they normally are.
// Ranks drift by up to a step instead of meeting after every forward.
//
// Mailbox of a rank: data [depth][world][msg_bytes] | flags [depth][world] uint32 (last step written into the slot).
// depth = 4: a slot written at step i was merged by its owner at step i - 3 at the latest before the sender could have
// received the owner's step i - 2 message, which the sender waited for at step i - 1.
#include <cstring>

#include "common.cuh"

namespace {

__host__ __device__ inline int64_t flag_offset(int64_t msg_bytes, int world, int depth) {
    return ((int64_t)depth * world * msg_bytes + 127) / 128 * 128;
}

struct PushParams {
    const uint4 *msg;
    unsigned char *const *peers;   // device array: base of every rank's mailbox (own included)
    int64_t msg_bytes, flag_off;
    int rank, world, slot;
    unsigned int seq;
};

__global__ void __launch_bounds__(256) screen_push_kernel(const PushParams p) {
    unsigned char *base = p.peers[blockIdx.x];
    uint4 *dst = reinterpret_cast<uint4 *>(base + ((int64_t)p.slot * p.world + p.rank) * p.msg_bytes);
    for (int64_t i = threadIdx.x; i < p.msg_bytes / 16; i += blockDim.x) dst[i] = p.msg[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int *flag = reinterpret_cast<unsigned int *>(base + p.flag_off) + p.slot * p.world + p.rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(p.seq) : "memory");
    }
}

__global__ void screen_wait_kernel(const unsigned int *flags, int world, unsigned int seq, int *status) {
    const int r = threadIdx.x;
    if (r >= world) return;
    unsigned long long t0 = 0, t1;
    unsigned int spins = 0, v;
    for (;;) {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + r) : "memory");
        if ((int)(v - seq) >= 0) break;   // (wrap-safe: step numbers only grow)
        if ((++spins & 63u) == 0) {
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t0 == 0) t0 = t1;
            else if (t1 - t0 > 20000000000ull) {  // 20 s: a peer died — surface it instead of hanging the stream (ranks that
                                                   // are merely late, e.g. a first launch that loads modules, must not trip it)
                if (status) atomicExch(status, 2);
                __trap();
            }
        }
    }
}

}  // namespace

extern "C" {

int64_t flexs_peer_mailbox_bytes(int64_t msg_bytes, int world, int depth) {
    if (msg_bytes <= 0 || msg_bytes % 16 || world < 1 || depth < 1) return -1;
    return flag_offset(msg_bytes, world, depth) + (int64_t)depth * world * 4 + 128;
}

int flexs_peer_alloc(int64_t bytes, void **d_ptr, unsigned char *handle64) {
    FX_REQUIRE(bytes > 0 && d_ptr && handle64, "bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    void *ptr = nullptr;
    FX_CUDA(cudaMalloc(&ptr, (size_t)bytes));
    FX_CUDA(cudaMemset(ptr, 0, (size_t)bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
    if (e != cudaSuccess) { cudaFree(ptr); FX_CUDA(e); }
    std::memcpy(handle64, &h, 64);
    *d_ptr = ptr;
    return FLEXS_OK;
}

int flexs_peer_open(const unsigned char *handle64, void **d_ptr) {
    FX_REQUIRE(handle64 && d_ptr, "bad arguments");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, 64);
    FX_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return FLEXS_OK;
}

int flexs_peer_close(void *d_ptr) {
    if (d_ptr) FX_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return FLEXS_OK;
}

int flexs_peer_free(void *d_ptr) {
    if (d_ptr) FX_CUDA(cudaFree(d_ptr));
    return FLEXS_OK;
}

int flexs_screen_push_dev(const void *d_msg, int64_t msg_bytes, int rank, int world, int slot, int depth, uint32_t seq,
                          const void *d_peer_bases, void *stream) {
    FX_REQUIRE(d_msg && d_peer_bases && msg_bytes > 0 && msg_bytes % 16 == 0, "bad message");
    FX_REQUIRE(world >= 1 && rank >= 0 && rank < world && depth >= 1 && slot >= 0 && slot < depth, "bad rank / slot");
    PushParams p;
    p.msg = reinterpret_cast<const uint4 *>(d_msg);
    p.peers = reinterpret_cast<unsigned char *const *>(d_peer_bases);
    p.msg_bytes = msg_bytes; p.flag_off = flag_offset(msg_bytes, world, depth);
    p.rank = rank; p.world = world; p.slot = slot; p.seq = seq;
    screen_push_kernel<<<world, 256, 0, (cudaStream_t)stream>>>(p);
    FX_CUDA(cudaGetLastError());
    return FLEXS_OK;
}

int flexs_screen_wait_dev(const void *d_mailbox, int64_t msg_bytes, int world, int slot, int depth, uint32_t seq,
                          int *d_status, void *stream) {
    FX_REQUIRE(d_mailbox && msg_bytes > 0 && msg_bytes % 16 == 0, "bad mailbox");
    FX_REQUIRE(world >= 1 && world <= 1024 && depth >= 1 && slot >= 0 && slot < depth, "bad world / slot");
    const unsigned int *flags =
        reinterpret_cast<const unsigned int *>(reinterpret_cast<const unsigned char *>(d_mailbox) + flag_offset(msg_bytes, world, depth)) +
        (int64_t)slot * world;
    screen_wait_kernel<<<1, ((world + 31) / 32) * 32, 0, (cudaStream_t)stream>>>(flags, world, seq, d_status);
    FX_CUDA(cudaGetLastError());
    return FLEXS_OK;
}

}  // extern "C"
