// Dispatch over the instantiation units of the cell-centric rollout kernel.
#include "ffm_internal.h"

namespace ffm {
const void* pick_cell_kernel_f32_c1(bool small, int nbr, bool dff, bool fs, int threads);
const void* pick_cell_kernel_f64_c1(bool small, int nbr, bool dff, bool fs, int threads);
#define FFM_DECL_CL(cl)                                                                              \
    const void* pick_cell_kernel_f32_c##cl(bool small, int nbr, bool dff, bool fs, int threads);     \
    const void* pick_cell_kernel_f64_c##cl(bool small, int nbr, bool dff, bool fs, int threads);
FFM_DECL_CL(2) FFM_DECL_CL(4) FFM_DECL_CL(8)
#undef FFM_DECL_CL

const void* pick_cell_kernel(bool f64, bool small, int nbr, bool dff, bool fs, int threads, int cluster) {
    if (cluster == 1) return f64 ? pick_cell_kernel_f64_c1(small, nbr, dff, fs, threads) : pick_cell_kernel_f32_c1(small, nbr, dff, fs, threads);
    if (cluster == 2) return f64 ? pick_cell_kernel_f64_c2(small, nbr, dff, fs, threads) : pick_cell_kernel_f32_c2(small, nbr, dff, fs, threads);
    if (cluster == 4) return f64 ? pick_cell_kernel_f64_c4(small, nbr, dff, fs, threads) : pick_cell_kernel_f32_c4(small, nbr, dff, fs, threads);
    if (cluster == 8) return f64 ? pick_cell_kernel_f64_c8(small, nbr, dff, fs, threads) : pick_cell_kernel_f32_c8(small, nbr, dff, fs, threads);
    return nullptr;
}
}  // namespace ffm
