// K1 (tcgen05 variant) — placeholder until the UMMA kernel lands; reports "not supported" so AUTO
// dispatch uses the FFMA kernel.
#include "common.cuh"

namespace fx {
bool cnn_umma_supported(const flexs_model *) { return false; }
int prepare_cnn_umma(flexs_model *) { return FLEXS_OK; }
int launch_cnn_umma(flexs_model *, const uint8_t *, int64_t, float *, cudaStream_t) {
    set_error("UMMA variant not built");
    return FLEXS_EINVAL;
}
}  // namespace fx
