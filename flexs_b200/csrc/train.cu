// K4 training — placeholder (implemented later this round).
#include "common.cuh"

extern "C" {
int flexs_model_fit_dev(flexs_model_t *, const uint8_t *, const float *, int64_t, int, int, uint64_t, float *, void *) {
    fx::set_error("fit not implemented yet");
    return FLEXS_EINVAL;
}
int flexs_model_train_step_dev(flexs_model_t *, int, const uint8_t *, const float *, int64_t, const float *, float *, void *) {
    fx::set_error("train_step not implemented yet");
    return FLEXS_EINVAL;
}
}
