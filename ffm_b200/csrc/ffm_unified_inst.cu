// Unified / trained model kernels: instantiations, pickers and the launchers of the table kernels.
#include "ffm_unified_kernel.cuh"
#include "ffm_internal.h"

namespace ffm {
namespace {
template <typename S, int NBR, bool FS, bool ACTOR>
const void* upick_threads(int threads) {
    if (threads >= 256) return (const void*)ffm_unified_rollout_kernel<S, NBR, FS, 256, ACTOR>;
    if (threads <= 64) return (const void*)ffm_unified_rollout_kernel<S, NBR, FS, 64, ACTOR>;
    return (const void*)ffm_unified_rollout_kernel<S, NBR, FS, 128, ACTOR>;
}
template <typename S, int NBR, bool ACTOR>
const void* upick_fs(bool fs, int threads) { return fs ? upick_threads<S, NBR, true, ACTOR>(threads) : upick_threads<S, NBR, false, ACTOR>(threads); }
template <typename S, bool ACTOR>
const void* upick_nbr(int nbr, bool fs, int threads) { return nbr == 4 ? upick_fs<S, 4, ACTOR>(fs, threads) : upick_fs<S, 8, ACTOR>(fs, threads); }
}  // namespace

const void* pick_unified_kernel(bool f64, int nbr, bool fs, int threads, bool actor) {
    // the actor / trained modes always score in float32 (the SFF is cast, ffm_unified.py:72-76): no float64 actor variants
    if (actor) return upick_nbr<float, true>(nbr, fs, threads);
    return f64 ? upick_nbr<double, false>(nbr, fs, threads) : upick_nbr<float, false>(nbr, fs, threads);
}

cudaError_t launch_apply_deltas(double* V, double* dV, double* dN, double* dF, double alpha_v, double* Hm, double* dH,
                                uint8_t* h_seen, uint8_t* v_seen, int S, int A, HStats* hstats, double* blk_lo, double* blk_hi,
                                int* blk_any, int blocks, cudaStream_t st) {
    unified_apply_deltas_kernel<<<blocks, 256, 0, st>>>(V, dV, dN, dF, alpha_v, Hm, dH, h_seen, v_seen, S, A, blk_lo, blk_hi, blk_any);
    if (Hm != nullptr) unified_finish_stats_kernel<<<1, 32, 0, st>>>(hstats, blk_lo, blk_hi, blk_any, blocks);
    return cudaGetLastError();
}

cudaError_t launch_rescan_hstats(const double* Hm, const uint8_t* h_seen, int S, int A, HStats* hstats, double* blk_lo,
                                 double* blk_hi, int* blk_any, int blocks, cudaStream_t st) {
    unified_apply_deltas_kernel<<<blocks, 256, 0, st>>>(nullptr, nullptr, nullptr, nullptr, 0.0, const_cast<double*>(Hm), nullptr,
                                                        const_cast<uint8_t*>(h_seen), nullptr, S, A, blk_lo, blk_hi, blk_any);
    unified_finish_stats_kernel<<<1, 32, 0, st>>>(hstats, blk_lo, blk_hi, blk_any, blocks);
    return cudaGetLastError();
}

cudaError_t launch_rollout_returns(const float* reward, const int32_t* len, int B, int T, int N, double gamma, double* G,
                                   cudaStream_t st) {
    const long long total = (long long)B * ((N + 3) / 4);
    const int blocks = (int)((total + 255) / 256 < 148LL * 8 ? (total + 255) / 256 : 148LL * 8);
    rollout_returns_kernel<<<blocks, 256, 0, st>>>(reward, len, B, T, N, gamma, G);
    return cudaGetLastError();
}

}  // namespace ffm
