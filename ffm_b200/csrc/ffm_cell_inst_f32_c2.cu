// Instantiations of the cell-centric base-model rollout kernel: float scores, one cluster of 2 CTAs per episode.
#include "ffm_cell_kernel.cuh"
#include "ffm_cell_inst.inl"
namespace ffm { const void* pick_cell_kernel_f32_c2(bool small, int nbr, bool dff, bool fs, int threads) { return cpick_ent<float, 2>(small, nbr, dff, fs, threads); } }
