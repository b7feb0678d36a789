// Instantiations of the (round-1, pedestrian-centric) base-model rollout kernel for float32 scores.
#include "ffm_core_kernel.cuh"
#include "ffm_core_inst.inl"
namespace ffm { const void* pick_core_kernel_f32(bool small, int nbr, bool dff, bool fs, int threads) { return pick_pos<float>(small, nbr, dff, fs, threads); } }
