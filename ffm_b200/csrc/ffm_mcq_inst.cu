// MC-Q model kernel: instantiations and picker.
#include "ffm_mcq_kernel.cuh"
#include "ffm_internal.h"

namespace ffm {
const void* pick_mcq_kernel(bool f64, int threads) {
    if (threads == 128) return f64 ? (const void*)ffm_mcq_rollout_kernel<double, 128> : (const void*)ffm_mcq_rollout_kernel<float, 128>;
    return f64 ? (const void*)ffm_mcq_rollout_kernel<double, 256> : (const void*)ffm_mcq_rollout_kernel<float, 256>;
}
}  // namespace ffm
