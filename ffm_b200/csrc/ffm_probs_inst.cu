// Move-probability probe kernel (parity tests): instantiations and picker.
#include "ffm_core_kernel.cuh"
#include "ffm_internal.h"

namespace ffm {
namespace {
template <typename S, int NBR>
const void* probs_pick_dff(bool dff) {
    return dff ? (const void*)core_move_probs_kernel<S, NBR, true> : (const void*)core_move_probs_kernel<S, NBR, false>;
}
}  // namespace
const void* pick_probs_kernel(bool f64, int nbr, bool dff) {
    return f64 ? (nbr == 4 ? probs_pick_dff<double, 4>(dff) : probs_pick_dff<double, 8>(dff))
               : (nbr == 4 ? probs_pick_dff<float, 4>(dff) : probs_pick_dff<float, 8>(dff));
}
}  // namespace ffm
