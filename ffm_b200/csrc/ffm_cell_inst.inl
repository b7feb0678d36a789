// template-parameter dispatch of ffm_cell_rollout_kernel for one score dtype / one cluster size
// (included by the instantiation units with FFM_CELL_S and FFM_CELL_CL defined)
namespace ffm { namespace {
template <typename S, typename EntT, int NBR, bool DFF, bool FS, int CL>
const void* cpick_threads(int threads) {
    if constexpr (CL > 1) {   // cluster variants: 512 threads per CTA (1024 measured slower: 64 registers, spills)
        if constexpr (CL == 2) {   // 1024 threads only pay in the 2-CTA form (measured on C3); larger clusters: 512
            if (threads == 1024) return (const void*)ffm_cell_rollout_kernel<S, EntT, NBR, DFF, FS, 1024, CL>;
        }
        return (const void*)ffm_cell_rollout_kernel<S, EntT, NBR, DFF, FS, 512, CL>;
    } else {
        if (threads == 1024) return (const void*)ffm_cell_rollout_kernel<S, EntT, NBR, DFF, FS, 1024, 1>;
        if (threads == 128) return (const void*)ffm_cell_rollout_kernel<S, EntT, NBR, DFF, FS, 128, 1>;
        return (const void*)ffm_cell_rollout_kernel<S, EntT, NBR, DFF, FS, 256, 1>;
    }
}
template <typename S, typename EntT, int NBR, bool DFF, int CL>
const void* cpick_fs(bool fs, int threads) {
    return fs ? cpick_threads<S, EntT, NBR, DFF, true, CL>(threads) : cpick_threads<S, EntT, NBR, DFF, false, CL>(threads);
}
template <typename S, typename EntT, int NBR, int CL>
const void* cpick_dff(bool dff, bool fs, int threads) {
    return dff ? cpick_fs<S, EntT, NBR, true, CL>(fs, threads) : cpick_fs<S, EntT, NBR, false, CL>(fs, threads);
}
template <typename S, typename EntT, int CL>
const void* cpick_nbr(int nbr, bool dff, bool fs, int threads) {
    return nbr == 4 ? cpick_dff<S, EntT, 4, CL>(dff, fs, threads) : cpick_dff<S, EntT, 8, CL>(dff, fs, threads);
}
template <typename S, int CL>
const void* cpick_ent(bool small, int nbr, bool dff, bool fs, int threads) {
    return small ? cpick_nbr<S, uint16_t, CL>(nbr, dff, fs, threads) : cpick_nbr<S, uint32_t, CL>(nbr, dff, fs, threads);
}
} }
