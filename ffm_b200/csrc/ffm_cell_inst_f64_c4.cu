// Instantiations of the cell-centric base-model rollout kernel: double scores, one cluster of 4 CTAs per episode.
#include "ffm_cell_kernel.cuh"
#include "ffm_cell_inst.inl"
namespace ffm { const void* pick_cell_kernel_f64_c4(bool small, int nbr, bool dff, bool fs, int threads) { return cpick_ent<double, 4>(small, nbr, dff, fs, threads); } }
