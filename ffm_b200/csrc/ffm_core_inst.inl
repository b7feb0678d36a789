// template-parameter dispatch of ffm_core_rollout_kernel (included by the per-dtype instantiation units)
namespace ffm { namespace {
template <typename S, typename PosT, int NBR, bool DFF, bool FS>
const void* pick_threads(int threads) {
    if (threads == 1024) return (const void*)ffm_core_rollout_kernel<S, PosT, NBR, DFF, FS, 1024>;
    if (threads == 128) return (const void*)ffm_core_rollout_kernel<S, PosT, NBR, DFF, FS, 128>;
    return (const void*)ffm_core_rollout_kernel<S, PosT, NBR, DFF, FS, 256>;
}
template <typename S, typename PosT, int NBR, bool DFF>
const void* pick_fs(bool fs, int threads) {
    return fs ? pick_threads<S, PosT, NBR, DFF, true>(threads) : pick_threads<S, PosT, NBR, DFF, false>(threads);
}
template <typename S, typename PosT, int NBR>
const void* pick_dff(bool dff, bool fs, int threads) {
    return dff ? pick_fs<S, PosT, NBR, true>(fs, threads) : pick_fs<S, PosT, NBR, false>(fs, threads);
}
template <typename S, typename PosT>
const void* pick_nbr(int nbr, bool dff, bool fs, int threads) {
    return nbr == 4 ? pick_dff<S, PosT, 4>(dff, fs, threads) : pick_dff<S, PosT, 8>(dff, fs, threads);
}
template <typename S>
const void* pick_pos(bool small, int nbr, bool dff, bool fs, int threads) {
    return small ? pick_nbr<S, uint16_t>(nbr, dff, fs, threads) : pick_nbr<S, uint32_t>(nbr, dff, fs, threads);
}
} }
