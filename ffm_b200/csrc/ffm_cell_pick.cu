// Dispatch over the instantiation units of the cell-centric rollout kernel.
#include "ffm_internal.h"

namespace ffm {
const void* pick_cell_kernel_f32_c1(bool small, int nbr, bool dff, bool fs, int threads);
const void* pick_cell_kernel_f64_c1(bool small, int nbr, bool dff, bool fs, int threads);

const void* pick_cell_kernel(bool f64, bool small, int nbr, bool dff, bool fs, int threads, int cluster) {
    if (cluster == 1) return f64 ? pick_cell_kernel_f64_c1(small, nbr, dff, fs, threads) : pick_cell_kernel_f32_c1(small, nbr, dff, fs, threads);
    return nullptr;
}
}  // namespace ffm
