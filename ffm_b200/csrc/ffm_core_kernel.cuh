// Persistent rollout kernel of the base floor-field CA: one CTA per episode, all steps in-kernel.
//
// Reproduces, per step, FloorFieldModel.step() + update_dff() of the reference
// (model/ffm_core.py:36-117) and, around it, the loop of run() (ffm_core.py:119-126):
//
//   phase A  candidate build, forced exit, SFF/DFF move probabilities, keyed draw -> target cell
//            (ffm_core.py:40-88)
//   phase B  same-target conflicts: lone claimant moves; k >= 2 claimants -> coin, then the
//            floor(u*k)-th claimant in ascending agent index moves; DFF footprint += 1
//            (ffm_core.py:90-98)
//   phase C  exit removal as a STABLE compaction, so alive rank == the reference's array index
//            (ffm_core.py:101-102), occupancy/owner grid update, optional trajectory row
//   phase D  DFF decay + diffusion with NumPy's float32 rounding sequence (ffm_core.py:106-117)
//
// HBM is touched only by the prologue (fields + positions in), the epilogue (state out) and the
// optional trajectory rows; everything a step reads or writes lives in shared memory when the
// fields fit (FIELDS_IN_SMEM), otherwise SFF scores / DFF stay in global memory (L2-resident).
//
// Shared-memory state of an episode
//   grid  u16[HW + 2*(W+1)]  bits 15..14 cell type (0 free, 1 wall, 2 exit, 3 other), bits 13..0
//                            1 + alive rank of the pedestrian standing there (0 = empty); a guard
//                            band of W+1 "wall" entries on both ends absorbs the neighbour reads
//                            of border (exit) cells
//   score S[HW]              -k_S * sff  (S = float | double, the dtype NumPy computes in)
//   dffA/dffB f32[HW]        dynamic floor field, ping-pong
//   posA/posB PosT[n_max]    linear cell per pedestrian, alive-rank order, ping-pong
//                            (PosT = u16 when H*W <= 65536, else u32)
//   tgt   PosT[n_max]        requested cell this step (all-ones = no request)
//   nxt   PosT[n_max]        cell after conflict resolution
//   wcnt  u32[n_max/32+2]    survivors per 32-pedestrian group (compaction scan input)
#pragma once
#include "ffm_device.cuh"

namespace ffm {

constexpr uint32_t TYPE_SHIFT = 14;
constexpr uint32_t OCC_MASK = 0x3FFFu;
constexpr uint32_t TYPE_FREE = 0, TYPE_WALL = 1, TYPE_EXIT = 2, TYPE_OTHER = 3;
constexpr int MAX_PEDS = 16382;

struct RolloutParams {
    int H, W, HW, n_max, B;
    int max_steps;
    const uint16_t* type_grid;   // [HW + 2*(W+1)] type bits only, guard band included
    const void* score;           // [HW] S
    float kd, c0, c1, thr;
    uint32_t* pos;               // [B][n_max] linear cells (always u32 in HBM)
    int32_t* n_alive;            // [B]
    int32_t* t_done;             // [B]
    unsigned long long* ped_steps;  // [B]
    float* dff;                  // [B][HW]   (global home of the DFF)
    float* dff_tmp;              // [B][HW]   second buffer when the fields stay in global memory
    unsigned long long seed;
    uint32_t episode_base;
    const double* move_draws;    // [B][draw_steps][n_max] or null
    const double* conflict_draws;  // [B][draw_steps][HW][2] or null
    int draw_steps, draw_first;
    uint32_t* traj;              // [B][traj_steps][n_max] or null
    int32_t* traj_n;             // [B][traj_steps]
    int traj_steps;
};

struct SmemLayout {
    uint32_t grid, score, dffA, dffB, posA, posB, tgt, nxt, wcnt, total;
};

__host__ __device__ inline uint32_t align16(uint32_t x) { return (x + 15u) & ~15u; }

__host__ __device__ inline SmemLayout make_layout(int HW, int W, int n_max, int sizeof_score, bool dff,
                                                  bool fields_in_smem) {
    const uint32_t ps = (HW <= 65536) ? 2u : 4u;   // sizeof(PosT)
    SmemLayout L;
    uint32_t o = 0;
    L.score = o; if (fields_in_smem) o = align16(o + (uint32_t)HW * sizeof_score);
    L.dffA = o;  if (fields_in_smem && dff) o = align16(o + (uint32_t)HW * 4u);
    L.dffB = o;  if (fields_in_smem && dff) o = align16(o + (uint32_t)HW * 4u);
    L.grid = o;  o = align16(o + (uint32_t)(HW + 2 * (W + 1)) * 2u);
    L.posA = o;  o = align16(o + (uint32_t)n_max * ps);
    L.posB = o;  o = align16(o + (uint32_t)n_max * ps);
    L.tgt = o;   o = align16(o + (uint32_t)n_max * ps);
    L.nxt = o;   o = align16(o + (uint32_t)n_max * ps);
    L.wcnt = o;  o = align16(o + (uint32_t)(n_max / 32 + 2) * 4u);
    L.total = o;
    return L;
}

// neighbour offsets in the reference's order (ffm_core.py:30 / :32-34)
template <int NBR> __device__ __forceinline__ int nbr_dr(int k);
template <int NBR> __device__ __forceinline__ int nbr_dc(int k);
template <> __device__ __forceinline__ int nbr_dr<4>(int k) { return k == 0 ? -1 : (k == 1 ? 1 : 0); }
template <> __device__ __forceinline__ int nbr_dc<4>(int k) { return k == 2 ? -1 : (k == 3 ? 1 : 0); }
template <> __device__ __forceinline__ int nbr_dr<8>(int k) { return k < 3 ? -1 : (k < 5 ? 0 : 1); }
template <> __device__ __forceinline__ int nbr_dc<8>(int k) {
    return (k == 0 || k == 3 || k == 5) ? -1 : ((k == 1 || k == 6) ? 0 : 1);
}

template <typename S, typename PosT, int NBR, bool DFF, bool FIELDS_IN_SMEM, int THREADS>
__global__ void __launch_bounds__(THREADS)
ffm_core_rollout_kernel(const RolloutParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int e = blockIdx.x;
    const int W = P.W, HW = P.HW, H = P.H;
    const int G = W + 1;  // guard band
    const SmemLayout L = make_layout(HW, W, P.n_max, (int)sizeof(S), DFF, FIELDS_IN_SMEM);

    uint16_t* grid = reinterpret_cast<uint16_t*>(smem_raw + L.grid) + G;   // grid[-G .. HW+G)
    constexpr uint32_t NONE_CELL = (uint32_t)(PosT)~(PosT)0;
    PosT* posA = reinterpret_cast<PosT*>(smem_raw + L.posA);
    PosT* posB = reinterpret_cast<PosT*>(smem_raw + L.posB);
    PosT* tgt = reinterpret_cast<PosT*>(smem_raw + L.tgt);
    PosT* nxt = reinterpret_cast<PosT*>(smem_raw + L.nxt);
    uint32_t* wcnt = reinterpret_cast<uint32_t*>(smem_raw + L.wcnt);

    const S* score;
    float* dffA = nullptr;
    float* dffB = nullptr;
    float* dff_home = DFF ? P.dff + (size_t)e * HW : nullptr;
    if (FIELDS_IN_SMEM) {
        S* s_sm = reinterpret_cast<S*>(smem_raw + L.score);
        const S* s_g = reinterpret_cast<const S*>(P.score);
        for (int c = tid; c < HW; c += THREADS) s_sm[c] = s_g[c];
        score = s_sm;
        if (DFF) {
            dffA = reinterpret_cast<float*>(smem_raw + L.dffA);
            dffB = reinterpret_cast<float*>(smem_raw + L.dffB);
            for (int c = tid; c < HW; c += THREADS) dffA[c] = dff_home[c];
        }
    } else {
        score = reinterpret_cast<const S*>(P.score);
        if (DFF) {
            dffA = dff_home;
            dffB = P.dff_tmp + (size_t)e * HW;
        }
    }

    // ---- prologue: occupancy/owner grid and positions ------------------------------------------
    for (int c = tid; c < HW + 2 * G; c += THREADS) grid[c - G] = P.type_grid[c];
    int n = P.n_alive[e];
    const int t0 = P.t_done[e];
    uint32_t* gpos = P.pos + (size_t)e * P.n_max;
    for (int i = tid; i < n; i += THREADS) posA[i] = (PosT)gpos[i];
    __syncthreads();
    for (int i = tid; i < n; i += THREADS) grid[posA[i]] |= (uint16_t)(i + 1);
    __syncthreads();

    const uint32_t episode = P.episode_base + (uint32_t)e;
    const double* mv_draws = P.move_draws ? P.move_draws + (size_t)e * P.draw_steps * P.n_max : nullptr;
    const double* cf_draws = P.conflict_draws ? P.conflict_draws + (size_t)e * P.draw_steps * HW * 2 : nullptr;

    int off[NBR];
#pragma unroll
    for (int k = 0; k < NBR; ++k) off[k] = nbr_dr<NBR>(k) * W + nbr_dc<NBR>(k);

    unsigned long long ped_steps = 0;
    int tl = 0;
    for (; tl < P.max_steps && n > 0; ++tl) {
        const uint32_t t = (uint32_t)(t0 + tl);
        ped_steps += (unsigned long long)n;
        const int di = (int)t - P.draw_first;   // row in the injected-draw buffers
        const bool inj = di >= 0 && di < P.draw_steps;

        // ================= phase A: choose a target cell ========================================
        for (int i = tid; i < n; i += THREADS) {
            const uint32_t c = posA[i];
            uint32_t m = 0, ex = 0;
#pragma unroll
            for (int k = 0; k < NBR; ++k) {
                const uint32_t g = grid[(int)c + off[k]];
                // passable (map 0 or 3, ffm_core.py:52-53) and not occupied at time t (:57-60)
                const bool is_exit = g == (TYPE_EXIT << TYPE_SHIFT);
                if (g == 0u || is_exit) m |= 1u << k;
                if (is_exit) ex |= 1u << k;
            }
            uint32_t target = NONE_CELL;
            if (m != 0u) {
                if ((grid[c] >> TYPE_SHIFT) == TYPE_EXIT) ex |= 1u << NBR;   // "stay" joins the candidates (:64)
                if (ex != 0u) {
                    // forced exit: first exit cell in candidate order, no draw (:66-72)
                    const int k = __ffs(ex) - 1;
                    target = (k == NBR) ? c : (uint32_t)((int)c + off[k]);
                } else {
                    const uint32_t mfull = m | (1u << NBR);
                    const int ncand = __popc(mfull);
                    S p[NBR + 1];
                    S mx = neg_inf<S>();
#pragma unroll
                    for (int k = 0; k <= NBR; ++k) {
                        p[k] = neg_inf<S>();
                        if ((mfull >> k) & 1u) {
                            const int cc = (k == NBR) ? (int)c : (int)c + off[k];
                            S s = score[cc];                                   // -k_S * sff
                            if (DFF) s = add_rn(s, (S)mul_rn(P.kd, dffA[cc]));  // + k_D * dff  (:77)
                            p[k] = s;
                            mx = max_t(mx, s);
                        }
                    }
#pragma unroll
                    for (int k = 0; k <= NBR; ++k)
                        if ((mfull >> k) & 1u) p[k] = exp_t(add_rn(p[k], -mx));  // exp(score - max) (:80)
                    const S sum = np_sum_masked<S, NBR + 1>(p, mfull, ncand);      // probs.sum() (:81)
                    if (isfinite(sum) && sum != (S)0) {                            // (:82)
                        double tot = 0.0;
#pragma unroll
                        for (int k = 0; k <= NBR; ++k)
                            if ((mfull >> k) & 1u) {
                                p[k] = div_rn(p[k], sum);                          // probs /= sum (:83)
                                tot = __dadd_rn(tot, (double)p[k]);                // choice(): cdf = cumsum(p)
                            }
                        const double u = (inj && mv_draws)
                                             ? mv_draws[(size_t)di * P.n_max + i]
                                             : draw_u0(P.seed, episode, t, STREAM_MOVE, (uint32_t)i);
                        // searchsorted(cdf / cdf[-1], u, 'right') == #{j : cdf_j / tot <= u}
                        double run = 0.0;
                        int j = 0;
#pragma unroll
                        for (int k = 0; k <= NBR; ++k)
                            if ((mfull >> k) & 1u) {
                                run = __dadd_rn(run, (double)p[k]);
                                j += (__ddiv_rn(run, tot) <= u) ? 1 : 0;
                            }
                        if (j >= ncand) j = ncand - 1;
                        const int slot = (int)__fns(mfull, 0, j + 1);
                        target = (slot == NBR) ? c : (uint32_t)((int)c + off[slot]);
                    }
                }
            }
            tgt[i] = (PosT)target;
        }
        __syncthreads();

        // ================= phase B: resolve same-target conflicts ===============================
        for (int base = 0; base < n; base += THREADS) {
            const int i = base + tid;
            const bool active = i < n;
            bool kept = false;
            if (active) {
                const uint32_t c = posA[i];
                const uint32_t T = tgt[i];
                uint32_t newc = c;
                if (T != NONE_CELL) {
                    bool moved;
                    if (T == c) {
                        moved = true;   // nobody else can request an occupied cell: lone claimant
                    } else {
                        int k = 0, r = 0;
#pragma unroll
                        for (int q = 0; q < NBR; ++q) {
                            const uint32_t occ = grid[(int)T + off[q]] & OCC_MASK;
                            if (occ != 0u) {
                                const int j = (int)occ - 1;
                                if (tgt[j] == T) { ++k; r += (j < i) ? 1 : 0; }
                            }
                        }
                        if (k == 1) {
                            moved = true;                                         // (:91-93)
                        } else {
                            Draw2 d;
                            if (inj && cf_draws) {
                                d.u0 = cf_draws[((size_t)di * HW + T) * 2];
                                d.u1 = cf_draws[((size_t)di * HW + T) * 2 + 1];
                            } else {
                                d = draw2(P.seed, episode, t, STREAM_CONFLICT, T);
                            }
                            // coin (:95), then agents[int(u * k)] in ascending agent index (:96)
                            moved = (d.u0 < 0.5) && ((int)(d.u1 * (double)k) == r);
                        }
                    }
                    if (moved) {
                        newc = T;
                        if (DFF) dffA[c] = __fadd_rn(dffA[c], 1.0f);               // footprint (:93,98)
                    }
                }
                kept = (grid[newc] >> TYPE_SHIFT) != TYPE_EXIT;                    // (:101)
                nxt[i] = (PosT)newc;
            }
            const uint32_t bal = __ballot_sync(0xffffffffu, kept);
            if (lane == 0 && i < n) wcnt[i >> 5] = __popc(bal);
        }
        __syncthreads();

        // ================= phase C: stable compaction + grid update =============================
        const int ngroups = (n + 31) >> 5;
        int n_new = 0;
        for (int base = 0; base < n; base += THREADS) {
            const int i = base + tid;
            const bool active = i < n;
            const int v = i >> 5;   // warp-uniform
            // exclusive prefix of wcnt[0..v) and grand total, by warp-wide reduction
            int before = 0, total = 0;
            for (int w0 = 0; w0 < ngroups; w0 += 32) {
                const int w = w0 + lane;
                const int x = (w < ngroups) ? (int)wcnt[w] : 0;
                int xb = (w < v) ? x : 0, xt = x;
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) {
                    xb += __shfl_xor_sync(0xffffffffu, xb, s);
                    xt += __shfl_xor_sync(0xffffffffu, xt, s);
                }
                before += xb;
                total += xt;
            }
            n_new = total;
            const uint32_t newc = active ? (uint32_t)nxt[i] : 0u;
            // only this thread writes grid[newc] in this phase, so its type bits are stable
            const bool kept = active && (grid[newc] >> TYPE_SHIFT) != TYPE_EXIT;
            const uint32_t bal = __ballot_sync(0xffffffffu, kept);
            if (active) {
                const uint32_t c = posA[i];
                if (newc != c) grid[c] &= (uint16_t)(3u << TYPE_SHIFT);
                if (kept) {
                    const int ni = before + __popc(bal & ((1u << lane) - 1u));
                    posB[ni] = (PosT)newc;
                    grid[newc] = (uint16_t)((grid[newc] & (3u << TYPE_SHIFT)) | (uint32_t)(ni + 1));
                }
            }
        }

        // ================= phase D: DFF decay + diffusion =======================================
        if (DFF) {
            // new = c0 * dff (ffm_core.py:109); the neighbour terms read this scaled field (:111)
            for (int c = tid; c < HW; c += THREADS) dffA[c] = __fmul_rn(P.c0, dffA[c]);
            __syncthreads();
            for (int c = tid; c < HW; c += THREADS) {
                const int r = c / W, col = c - r * W;
                float acc = dffA[c];
#pragma unroll
                for (int k = 0; k < NBR; ++k) {
                    const int rr = r + nbr_dr<NBR>(k), cc = col + nbr_dc<NBR>(k);
                    const float v = (rr >= 0 && rr < H && cc >= 0 && cc < W) ? dffA[rr * W + cc] : 0.0f;
                    acc = __fadd_rn(acc, __fmul_rn(P.c1, v));                      // (:112-113)
                }
                if (acc < P.thr) acc = 0.0f;                                       // (:116-117)
                dffB[c] = acc;
            }
            float* tmp = dffA; dffA = dffB; dffB = tmp;
        }
        __syncthreads();

        // trajectory row: positions after this step, alive-rank order (ffm_core.py:125)
        if (P.traj != nullptr && tl < P.traj_steps) {
            uint32_t* row = P.traj + ((size_t)e * P.traj_steps + tl) * P.n_max;
            for (int i = tid; i < n_new; i += THREADS) row[i] = posB[i];
            if (tid == 0) P.traj_n[(size_t)e * P.traj_steps + tl] = n_new;
        }
        PosT* tp = posA; posA = posB; posB = tp;
        n = n_new;
    }

    // ---- epilogue: state back to HBM ------------------------------------------------------------
    for (int i = tid; i < n; i += THREADS) gpos[i] = posA[i];
    if (DFF) {
        if (dffA != dff_home)
            for (int c = tid; c < HW; c += THREADS) dff_home[c] = dffA[c];
    }
    if (tid == 0) {
        P.n_alive[e] = n;
        P.t_done[e] = t0 + tl;
        P.ped_steps[e] += ped_steps;
    }
}

}  // namespace ffm
