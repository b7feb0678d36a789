// Host-side glue between the translation units of libffm_b200 (not part of the C ABI).
// Each kernel family lives in its own .cu and exposes kernel pickers / launchers here; ffm_api.cu owns
// the handles and the extern "C" surface of include/ffm_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>

namespace ffm {

// stores the message ffm_last_error() returns for the calling thread; returns `code`
int set_error(int code, const char* fmt, va_list ap);

struct HStats;

// ---- base CA model (ffm_cell_kernel.cuh) ----------------------------------------------------------
// cluster = CTAs per episode (1, 2, 4, 8): the map is split into row bands held in distributed shared memory
const void* pick_cell_kernel(bool f64, bool small, int nbr, bool dff, bool fields_in_smem, int threads, int cluster);
const void* pick_probs_kernel(bool f64, int nbr, bool dff);
// round-1 pedestrian-centric kernel (ffm_core_kernel.cuh), kept selectable for A/B measurements: FFM_KERNEL=ped
const void* pick_core_kernel_f32(bool small, int nbr, bool dff, bool fields_in_smem, int threads);
const void* pick_core_kernel_f64(bool small, int nbr, bool dff, bool fields_in_smem, int threads);

// ---- unified / trained models (ffm_unified_kernel.cuh) --------------------------------------------
const void* pick_unified_kernel(bool f64, int nbr, bool fields_in_smem, int threads, bool actor);
// dF (may be null): per-state "key touched in this sync" marks (0 / >0), folded into v_seen; a state with dN > 0 also
// gets its H row marked present when the actor learns (Hm != null)
cudaError_t launch_apply_deltas(double* V, double* dV, double* dN, double* dF, double alpha_v, double* Hm, double* dH,
                                uint8_t* h_seen, uint8_t* v_seen, int S, int A, HStats* hstats, double* blk_lo, double* blk_hi,
                                int* blk_any, int blocks, cudaStream_t st);
cudaError_t launch_rescan_hstats(const double* Hm, const uint8_t* h_seen, int S, int A, HStats* hstats, double* blk_lo,
                                 double* blk_hi, int* blk_any, int blocks, cudaStream_t st);
cudaError_t launch_rollout_returns(const float* reward, const int32_t* len, int B, int T, int N, double gamma, double* G,
                                   cudaStream_t st);

// ---- MC-Q model (ffm_mcq_kernel.cuh) ---------------------------------------------------------------
const void* pick_mcq_kernel(bool f64, int threads);

}  // namespace ffm
