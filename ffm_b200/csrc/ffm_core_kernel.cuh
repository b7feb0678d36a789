// Persistent rollout kernel of the base floor-field CA: one CTA per episode, all steps in-kernel.
//
// Reproduces, per step, FloorFieldModel.step() + update_dff() of the reference
// (model/ffm_core.py:36-117) and, around it, the loop of run() (ffm_core.py:119-126).
//
// The reference walks the pedestrians one by one; here a step is four block-wide phases over
// shared-memory state, with the expensive work compacted onto dense work lists so that warps run
// with full lanes even when most of a packed crowd cannot move:
//
//   A1 (every live slot)   8/4 neighbour cells -> candidate mask (passable and unoccupied at time t,
//                          ffm_core.py:52-60).  No candidate: no request, no draw (:63).  An exit
//                          among the candidates: forced request, no draw (:66-72) -> request list.
//                          Otherwise -> draw list.
//   A2 (draw list)         score = -k_S*sff + k_D*dff, exp(score - max), NumPy-ordered sum,
//                          normalise, float64 CDF, keyed uniform -> target (:74-88) -> request list
//   B  (request list)      same-target conflicts: lone claimant moves; k >= 2 claimants -> coin, then
//                          the floor(u*k)-th claimant in ascending agent index moves; DFF footprint
//                          += 1 (:90-98).  Claimants are found by looking at the owners of the
//                          target's neighbour cells (no atomics, deterministic).
//   C  (moved requests)    apply: owner grid, position; pedestrians that reached an exit clear their
//                          alive bit (:101-102)
//   D  (every cell)        DFF decay + diffusion with NumPy's float32 rounding sequence (:106-117)
//
// Pedestrians keep a stable SLOT; the reference's array index (which shifts when somebody leaves,
// :102) is the slot's alive rank = prefix popcount over the alive bitmap, recomputed only in steps
// where somebody left.  Slots are physically re-packed when a quarter of them are dead.
//
// HBM is touched only by the prologue (fields + positions in), the epilogue (state out) and the
// optional trajectory rows; everything a step reads or writes lives in shared memory when the
// fields fit (FIELDS_IN_SMEM), otherwise SFF scores / DFF stay in global memory (L2-resident).
//
// Shared-memory state of an episode
//   grid  u16[HW + 2*(W+1)]  bits 15..14 cell type (0 free, 1 blocked, 2 exit, 3 free cell next to an
//                            exit), bits 13..0 = 1 + slot of the pedestrian standing there, 0 = empty,
//                            0x3FFF on blocked cells (walls look "occupied": one test per neighbour);
//                            a guard band of W+1 blocked entries on both ends absorbs neighbour reads
//                            of border cells
//   claim u4[HW]             requests per target cell this step (nibbles, 8 cells per word; swept clean in C)
//   score S[HW]              -k_S * sff  (S = float | double, the dtype NumPy computes in)
//   dffA/dffB f32[HW]        dynamic floor field, ping-pong
//   pos   PosT[n_max]        linear cell per slot (PosT = u16 when H*W <= 65536, else u32)
//   tgt   PosT[n_max]        per slot: candidate mask between A1 and A2, then the requested cell
//                            (all-ones = no request)
//   list  u16[n_max]         ONE work list: A1 pushes the slots that must draw from the front and the
//                            forced requests from the back; A2 turns each draw entry into a request
//                            entry in place (0xFFFF = no request); | 0x8000 once the request is granted
//   alive u32[n_max/32+1]    alive bitmap;  wpre u32[n_max/32+1] exclusive prefix popcounts
#pragma once
#include "ffm_device.cuh"
#include "ffm_dff_stencil.cuh"

namespace ffm {

constexpr uint32_t TYPE_SHIFT = 14;
constexpr uint32_t OCC_MASK = 0x3FFFu;
constexpr uint32_t TYPE_BITS = 3u << TYPE_SHIFT;
constexpr uint32_t TYPE_FREE = 0, TYPE_WALL = 1, TYPE_EXIT = 2, TYPE_NEAR_EXIT = 3;
constexpr uint32_t WALL_CELL = (TYPE_WALL << TYPE_SHIFT) | OCC_MASK;
constexpr int MAX_PEDS = 16380;   // owner ids 1..16380; 0x3FFE / 0x3FFF mark blocked cells

struct RolloutParams {
    int H, W, HW, n_max, B;
    int max_steps;
    const uint16_t* type_grid;   // [HW + 2*(W+1)] type bits only, guard band included
    const void* score;           // [HW] S
    float kd, c0, c1, thr;
    uint32_t* pos;               // [B][n_max] linear cells, alive-rank order (always u32 in HBM)
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
    int32_t* err;                // device validation flag (128: two pedestrians on one cell)
};

struct SmemLayout {
    uint32_t score, dffA, dffB, grid, claim, pos, tgt, list, alive, wpre, ctr, bar, total;
};

__host__ __device__ inline uint32_t align16(uint32_t x) { return (x + 15u) & ~15u; }

__host__ __device__ inline SmemLayout make_layout(int HW, int W, int n_max, int sizeof_score, bool dff,
                                                  bool fields_in_smem) {
    const uint32_t ps = (HW <= 65536) ? 2u : 4u;   // sizeof(PosT)
    const uint32_t nw = (uint32_t)(n_max + 31) / 32 + 1;
    SmemLayout L;
    uint32_t o = 0;
    L.score = o; if (fields_in_smem) o = align16(o + (uint32_t)HW * sizeof_score);
    L.dffA = o;  if (fields_in_smem && dff) o = align16(o + (uint32_t)HW * 4u);
    L.dffB = o;  if (fields_in_smem && dff) o = align16(o + (uint32_t)HW * 4u);
    L.grid = o;  o = align16(o + (uint32_t)(HW + 2 * (W + 1)) * 2u);
    L.claim = o; o = align16(o + ((uint32_t)HW + 7u) / 8u * 4u);
    L.pos = o;   o = align16(o + (uint32_t)n_max * ps);
    L.tgt = o;   o = align16(o + (uint32_t)n_max * ps);
    L.list = o;  o = align16(o + (uint32_t)n_max * 2u);
    L.alive = o; o = align16(o + nw * 4u);
    L.wpre = o;  o = align16(o + nw * 4u);
    L.ctr = o;   o = align16(o + 8u * 4u);
    L.bar = o;   o = align16(o + 8u);
    L.total = o;
    return L;
}

// neighbour offsets in the reference's order (ffm_core.py:30 / :32-34)
template <int NBR> __device__ __forceinline__ int nbr_dr(int k);
template <int NBR> __device__ __forceinline__ int nbr_dc(int k);
template <> __device__ __forceinline__ int nbr_dr<4>(int k) { return k == 0 ? -1 : (k == 1 ? 1 : 0); }
template <> __device__ __forceinline__ int nbr_dc<4>(int k) { return k == 2 ? -1 : (k == 3 ? 1 : 0); }
template <> __device__ __forceinline__ int nbr_dr<8>(int k) { const int q = k + (k >= 4 ? 1 : 0); return q / 3 - 1; }
template <> __device__ __forceinline__ int nbr_dc<8>(int k) { const int q = k + (k >= 4 ? 1 : 0); return q % 3 - 1; }
template <int NBR> __device__ __forceinline__ int nbr_off(int k, int W) { return nbr_dr<NBR>(k) * W + nbr_dc<NBR>(k); }

// same, for a neighbour index only known at run time (2-bit lookup tables of dr+1 / dc+1)
template <int NBR> __device__ __forceinline__ int nbr_off_rt(int k, int W);
template <> __device__ __forceinline__ int nbr_off_rt<8>(int k, int W) {
    const int dr = (int)((0xA940u >> (2 * k)) & 3u) - 1, dc = (int)((0x9224u >> (2 * k)) & 3u) - 1;
    return dr * W + dc;
}
template <> __device__ __forceinline__ int nbr_off_rt<4>(int k, int W) {
    const int dr = (int)((0x58u >> (2 * k)) & 3u) - 1, dc = (int)((0x85u >> (2 * k)) & 3u) - 1;
    return dr * W + dc;
}

__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// Warp-aggregated append of `value` to a shared-memory list (one atomic per warp).  All 32 lanes
// must call it; lanes with pred == false append nothing.
// dir = +1 fills list[0], list[1], ...; dir = -1 fills list[last], list[last-1], ...
template <typename T>
__device__ __forceinline__ void warp_append(bool pred, T value, T* list, uint32_t* counter, int lane, int dir = 1, int last = 0) {
    const uint32_t bal = __ballot_sync(0xffffffffu, pred);
    if (bal == 0u) return;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(counter, (uint32_t)__popc(bal));
    base = __shfl_sync(0xffffffffu, base, 0);
    const int k = (int)base + __popc(bal & lanemask_lt());
    if (pred) list[dir > 0 ? k : last - k] = value;
}

// Unnormalised move probabilities of one pedestrian (ffm_core.py:74-80): candidates in COMPACTED order (set bits
// of the neighbour mask, then "stay"), e_j = exp(score_j - max_j score_j) with score = -k_S*sff + k_D*dff in the
// dtype NumPy uses.  Returns sum_j e_j (float64); cell[j] / e[j] are filled for j < ncand.  The reference then
// normalises (p = e / sum(e), :83) and np.random.choice picks the first j with cumsum(p)[j]/cumsum(p)[-1] > u
// (:84) -- the first j with E_j > u * E_n (E = running sums of e) up to the rounding of the normalisation
// (< 1e-7), which, like the <= 2 ulp difference between NumPy's exp and CUDA's, only matters for a draw closer
// than that to a CDF boundary; the parity bar excludes those draws (tests/helpers.py MARGIN_GUARD).
template <typename S, int NBR, bool DFF>
__device__ __forceinline__ double move_weights(uint32_t mm, int ncand, int c, int W, const S* score, const float* dffA, float kd,
                                               int (&cell)[NBR + 1], S (&e)[NBR + 1]) {
    S mx = neg_inf<S>();
#pragma unroll
    for (int j = 0; j <= NBR; ++j)
        if (j < ncand) {
            int cc = c;
            if (j < ncand - 1) {
                const int k = __ffs(mm) - 1;
                mm &= mm - 1u;
                cc = c + nbr_off_rt<NBR>(k, W);
            }
            cell[j] = cc;
            S sc = score[cc];                                    // -k_S * sff
            if (DFF) sc = add_rn(sc, (S)mul_rn(kd, dffA[cc]));   // + k_D * dff   (:77)
            e[j] = sc;
            mx = max_t(mx, sc);
        }
    double tot = 0.0;
#pragma unroll
    for (int j = 0; j <= NBR; ++j)
        if (j < ncand) {
            e[j] = exp_t(add_rn(e[j], -mx));                     // exp(score - max) (:80)
            tot += (double)e[j];
        }
    return tot;
}

// update_dff (ffm_core.py:106-117) in one pass: out = threshold(s + sum_k c1 * s[nb_k]) with s = c0 * in,
// each product and sum rounded separately in the reference's neighbour order.  A thread walks down a
// column strip with a 3-row register window, so a cell costs 3 loads instead of 9 and every product
// c1 * s is formed once; out-of-map neighbours contribute exactly 0 (np.pad of the scaled field, :111).
template <int NBR>
__device__ __forceinline__ void dff_decay_diffuse(const float* __restrict__ in, float* __restrict__ out, int H, int W,
                                                  float c0, float c1, float thr, int tid, const StencilGeom& geom) {
    if ((W & 3) == 0 && W >= 32) {   // rows are 16-byte aligned and long enough to fill lanes: the vectorised walk (ffm_dff_stencil.cuh), same arithmetic; tiny maps (12x12: 3 column groups) keep the scalar walk, measured faster there
        dff_stencil_v4<NBR>([&](int r) -> const float* { return in + (size_t)r * W; }, [](int) -> const float* { return nullptr; },
                            [&](int r) -> float* { return out + (size_t)r * W; }, 0, H, W, c0, c1, thr, tid, geom);
        return;
    }
    const int cw = geom.cw;                              // columns handled per sweep
    const int bands = geom.bands;                        // row bands working in parallel on one column sweep
    const int rpb = geom.srpb;
    const int band = geom.band, colb = geom.colb;
    if (band >= bands) return;
    const int r0 = band * rpb, r1 = min(H, r0 + rpb);
    for (int col = colb; col < W; col += cw) {
        const bool hasl = col > 0, hasr = col + 1 < W;
        float up[3], uc[3], un[3], sc = 0.0f;            // c1*s of rows r-1, r, r+1 (left, centre, right); s of the centre
        auto load_row = [&](int r, float (&u)[3], float& s_centre) {
            float d0 = 0.0f, d1 = 0.0f, d2 = 0.0f;
            if (r >= 0 && r < H) {
                const float* row = in + (size_t)r * W + col;
                d1 = row[0];
                if (hasl) d0 = row[-1];
                if (hasr) d2 = row[1];
            }
            const float s0 = __fmul_rn(c0, d0), s1 = __fmul_rn(c0, d1), s2 = __fmul_rn(c0, d2);   // (:109)
            u[0] = __fmul_rn(c1, s0); u[1] = __fmul_rn(c1, s1); u[2] = __fmul_rn(c1, s2);         // (:113)
            s_centre = s1;
        };
        float dummy;
        load_row(r0 - 1, up, dummy);
        load_row(r0, uc, sc);
        for (int r = r0; r < r1; ++r) {
            float sn;
            load_row(r + 1, un, sn);
            float acc = sc;
            if (NBR == 8) {   // (-1,-1) (-1,0) (-1,1) (0,-1) (0,1) (1,-1) (1,0) (1,1)
                acc = __fadd_rn(acc, up[0]); acc = __fadd_rn(acc, up[1]); acc = __fadd_rn(acc, up[2]);
                acc = __fadd_rn(acc, uc[0]); acc = __fadd_rn(acc, uc[2]);
                acc = __fadd_rn(acc, un[0]); acc = __fadd_rn(acc, un[1]); acc = __fadd_rn(acc, un[2]);
            } else {          // (-1,0) (1,0) (0,-1) (0,1)
                acc = __fadd_rn(acc, up[1]); acc = __fadd_rn(acc, un[1]);
                acc = __fadd_rn(acc, uc[0]); acc = __fadd_rn(acc, uc[2]);
            }
            if (acc < thr) acc = 0.0f;                                                            // (:116-117)
            out[(size_t)r * W + col] = acc;
#pragma unroll
            for (int k = 0; k < 3; ++k) { up[k] = uc[k]; uc[k] = un[k]; }
            sc = sn;
        }
    }
}

template <typename S, typename PosT, int NBR, bool DFF, bool FIELDS_IN_SMEM, int THREADS>
__global__ void __launch_bounds__(THREADS, (THREADS <= 256 && sizeof(S) == 4) ? 1536 / THREADS : 1)
ffm_core_rollout_kernel(const RolloutParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int e = blockIdx.x;
    const int W = P.W, HW = P.HW, H = P.H;
    const int G = W + 1;  // guard band
    const SmemLayout L = make_layout(HW, W, P.n_max, (int)sizeof(S), DFF, FIELDS_IN_SMEM);
    constexpr uint32_t NONE_CELL = (uint32_t)(PosT)~(PosT)0;
    constexpr uint32_t EXIT_EMPTY = TYPE_EXIT << TYPE_SHIFT;

    uint16_t* grid = reinterpret_cast<uint16_t*>(smem_raw + L.grid) + G;   // grid[-G .. HW+G)
    PosT* pos = reinterpret_cast<PosT*>(smem_raw + L.pos);
    PosT* tgt = reinterpret_cast<PosT*>(smem_raw + L.tgt);
    uint16_t* list = reinterpret_cast<uint16_t*>(smem_raw + L.list);
    const int list_last = P.n_max - 1;
    uint32_t* alive = reinterpret_cast<uint32_t*>(smem_raw + L.alive);
    uint32_t* wpre = reinterpret_cast<uint32_t*>(smem_raw + L.wpre);
    uint32_t* ctr = reinterpret_cast<uint32_t*>(smem_raw + L.ctr);   // [parity][n_work, n_req, n_exit, -]
    uint32_t* claim32 = reinterpret_cast<uint32_t*>(smem_raw + L.claim);
    const int claim_words = (HW + 7) / 8;

    // ---- prologue: stage the episode's fields with TMA bulk copies (cp.async.bulk + mbarrier) ------
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem_raw + L.bar);
    const S* score;
    float* dffA = nullptr;
    float* dffB = nullptr;
    float* dff_home = DFF ? P.dff + (size_t)e * HW : nullptr;
    const uint32_t grid_bytes = align16((uint32_t)(HW + 2 * G) * 2u);      // d_type_grid is padded to this size
    const uint32_t score_bytes = (uint32_t)HW * (uint32_t)sizeof(S), dff_bytes = (uint32_t)HW * 4u;
    const bool tma_fields = FIELDS_IN_SMEM && (score_bytes % 16u == 0u) && (dff_bytes % 16u == 0u);
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) {
        uint32_t bytes = grid_bytes;
        if (tma_fields) bytes += score_bytes + (DFF ? dff_bytes : 0u);
        mbar_arrive_expect_tx(bar, bytes);
        bulk_copy_g2s(smem_raw + L.grid, P.type_grid, grid_bytes, bar);
        if (tma_fields) {
            bulk_copy_g2s(smem_raw + L.score, P.score, score_bytes, bar);
            if (DFF) bulk_copy_g2s(smem_raw + L.dffA, dff_home, dff_bytes, bar);
        }
    }
    if (FIELDS_IN_SMEM) {
        S* s_sm = reinterpret_cast<S*>(smem_raw + L.score);
        score = s_sm;
        if (DFF) {
            dffA = reinterpret_cast<float*>(smem_raw + L.dffA);
            dffB = reinterpret_cast<float*>(smem_raw + L.dffB);
        }
        if (!tma_fields) {     // odd sizes: ordinary loads
            const S* s_g = reinterpret_cast<const S*>(P.score);
            for (int c = tid; c < HW; c += THREADS) s_sm[c] = s_g[c];
            if (DFF) for (int c = tid; c < HW; c += THREADS) dffA[c] = dff_home[c];
        }
    } else {
        score = reinterpret_cast<const S*>(P.score);
        if (DFF) {
            dffA = dff_home;
            dffB = P.dff_tmp + (size_t)e * HW;
        }
    }

    // ---- owner grid, positions, alive bitmap -----------------------------------------------------
    int n = P.n_alive[e];        // pedestrians still inside
    int n_slots = n;             // slots in use (live + dead since the last re-pack)
    const int t0 = P.t_done[e];
    uint32_t* gpos = P.pos + (size_t)e * P.n_max;
    for (int i = tid; i < n; i += THREADS) pos[i] = (PosT)gpos[i];
    for (int w = tid; w <= (n + 31) / 32; w += THREADS) {
        const int lo = w * 32;
        alive[w] = (lo + 32 <= n) ? 0xffffffffu : (lo < n ? ((1u << (n - lo)) - 1u) : 0u);
        wpre[w] = (uint32_t)(lo < n ? lo : n);
    }
    if (tid < 8) ctr[tid] = 0u;
    for (int c = tid; c < claim_words; c += THREADS) claim32[c] = 0u;
    mbar_wait(bar, 0);          // the bulk copies have landed (phase 0 of the barrier completed)
    __syncthreads();
    for (int i = tid; i < n; i += THREADS) grid[pos[i]] |= (uint16_t)(i + 1);
    __syncthreads();
    // two pedestrians on one cell would OR their ids together: at least one of them then reads back a foreign id
    for (int i = tid; i < n; i += THREADS)
        if ((grid[pos[i]] & OCC_MASK) != (uint32_t)(i + 1) && P.err != nullptr) atomicOr(P.err, 128);

    const uint32_t episode = P.episode_base + (uint32_t)e;
    const double* mv_draws = P.move_draws ? P.move_draws + (size_t)e * P.draw_steps * P.n_max : nullptr;
    const double* cf_draws = P.conflict_draws ? P.conflict_draws + (size_t)e * P.draw_steps * HW * 2 : nullptr;

    const StencilGeom sgeom = make_stencil_geom(H, W, tid, THREADS);   // DFF stencil geometry, once (not per step)
    unsigned long long ped_steps = 0;
    int tl = 0;
    for (; tl < P.max_steps && n > 0; ++tl) {
        const uint32_t t = (uint32_t)(t0 + tl);
        ped_steps += (unsigned long long)n;
        const int di = (int)t - P.draw_first;   // row in the injected-draw buffers
        const bool inj = di >= 0 && di < P.draw_steps;
        uint32_t* cnt = ctr + ((tl & 1) << 2);          // this step's counters
        if (tid == 0) {                                   // next step's counters (idle this step)
            uint32_t* nx = ctr + (((tl + 1) & 1) << 2);
            nx[0] = 0u; nx[1] = 0u; nx[2] = 0u;
        }

        // ================= A1: candidate masks, forced exits ====================================
        for (int base = 0; base < n_slots; base += THREADS) {
            const int s = base + tid;
            bool need_draw = false, forced = false;
            uint32_t m = 0;
            if (s < n_slots && ((alive[s >> 5] >> (s & 31)) & 1u)) {
                const int c = (int)pos[s];
                // passable (map 0 or 3, ffm_core.py:52-53) and not occupied at time t (:57-60):
                // blocked cells carry 0x3FFF in the owner bits, so "owner bits == 0" is the whole test
#pragma unroll
                for (int k = 0; k < NBR; ++k)
                    if ((grid[c + nbr_off<NBR>(k, W)] & OCC_MASK) == 0u) m |= 1u << k;
                uint32_t target = NONE_CELL;
                if (m != 0u) {
                    const uint32_t own_type = grid[c] >> TYPE_SHIFT;
                    uint32_t ex = 0;
                    if (own_type >= TYPE_EXIT) {          // on an exit, or a free cell next to one (static flag)
                        if (own_type == TYPE_EXIT) ex |= 1u << NBR;   // "stay" joins the candidates (:64)
#pragma unroll
                        for (int k = 0; k < NBR; ++k)
                            if (grid[c + nbr_off<NBR>(k, W)] == EXIT_EMPTY) ex |= 1u << k;
                    }
                    if (ex != 0u) {
                        // forced exit: first exit cell in candidate order, no draw (:66-72)
                        const int k = __ffs(ex) - 1;
                        target = (k == NBR) ? (uint32_t)c : (uint32_t)(c + nbr_off_rt<NBR>(k, W));
                        forced = true;
                        if (k != NBR) atomicAdd(&claim32[target >> 3], 1u << (4 * (target & 7u)));
                    } else {
                        need_draw = true;
                    }
                }
                tgt[s] = need_draw ? (PosT)m : (PosT)target;     // the mask rides in tgt[] until A2
            }
            warp_append<uint16_t>(need_draw, (uint16_t)s, list, &cnt[0], lane);
            if (forced) list[list_last - (int)atomicAdd(&cnt[1], 1u)] = (uint16_t)s;   // rare (next to an exit): no warp aggregation
        }
        __syncthreads();

        // ================= A2: move probabilities and keyed draw (draw list) ====================
        // Candidates are walked in COMPACTED order (set bits of the mask, then "stay"), which is the
        // order of the reference's neighbor_coords array (:54,60,64) -- and keeps lanes busy.
        const int n_work = (int)cnt[0];
        for (int base = 0; base < n_work; base += THREADS) {
            const int wi = base + tid;
            if (wi < n_work) {
                const int s = (int)list[wi];
                uint32_t mm = (uint32_t)tgt[s];
                const int ncand = __popc(mm) + 1;
                const int c = (int)pos[s];
                int cell[NBR + 1];
                S p[NBR + 1];
                const double tot = move_weights<S, NBR, DFF>(mm, ncand, c, W, score, dffA, P.kd, cell, p);
                uint32_t target = NONE_CELL;
                if (isfinite(tot) && tot != 0.0) {                            // (:82)
                    // the reference's array index of this pedestrian = alive rank of its slot
                    const uint32_t rank = wpre[s >> 5] + (uint32_t)__popc(alive[s >> 5] & ((1u << (s & 31)) - 1u));
                    const double u = (inj && mv_draws) ? mv_draws[(size_t)di * P.n_max + rank]
                                                       : draw_u0(P.seed, episode, t, STREAM_MOVE, rank);
                    const double thresh = u * tot;
                    double run = 0.0;
                    target = (uint32_t)c;                                     // E_n > u * E_n always: "stay" is last
                    bool done = false;
#pragma unroll
                    for (int j = 0; j < NBR; ++j)
                        if (j < ncand - 1 && !done) {
                            run += (double)p[j];
                            if (run > thresh) { target = (uint32_t)cell[j]; done = true; }
                        }
                    if (target != (uint32_t)c) atomicAdd(&claim32[target >> 3], 1u << (4 * (target & 7u)));
                } else {
                    list[wi] = (uint16_t)0xFFFFu;                             // no request after all (:82)
                }
                tgt[s] = (PosT)target;
            }
        }
        __syncthreads();

        // ================= B: resolve same-target conflicts (request list) ======================
        const int n_req = n_work + (int)cnt[1];
        for (int j = tid; j < n_req; j += THREADS) {
            const int li = j < n_work ? j : list_last - (j - n_work);
            const int s = (int)list[li];
            if (s == 0xFFFF) continue;
            const int c = (int)pos[s];
            const uint32_t T = tgt[s];
            bool moved;
            if (T == (uint32_t)c) {
                moved = true;   // nobody else can request an occupied cell: lone claimant of "stay"
            } else {
                const int k = (int)((claim32[T >> 3] >> (4 * (T & 7u))) & 0xFu);
                if (k == 1) {
                    moved = true;                                             // (:91-93)
                } else {
                    // rank among the claimants in ascending agent index: the claimants are owners of
                    // T's neighbour cells whose request is T
                    int r = 0;
#pragma unroll
                    for (int q = 0; q < NBR; ++q) {
                        const uint32_t o = (grid[(int)T + nbr_off<NBR>(q, W)] & OCC_MASK) - 1u;
                        if (o < (uint32_t)s && (uint32_t)tgt[o] == T) ++r;    // o < s also excludes empty / blocked
                    }
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
                if (DFF) dffA[c] = __fadd_rn(dffA[c], 1.0f);                  // footprint (:93,98)
                if (T != (uint32_t)c) list[li] = (uint16_t)(s | 0x8000);      // to be applied in C
            }
        }
        __syncthreads();

        // ================= C: apply granted moves, exits ========================================
        for (int base = 0; base < n_req; base += THREADS) {
            const int j = base + tid;
            bool leaves = false;
            if (j < n_req) {
                const uint32_t rv = list[j < n_work ? j : list_last - (j - n_work)];
                if ((rv & 0x8000u) && rv != 0xFFFFu) {
                    const int s = (int)(rv & 0x7FFFu);
                    const int c = (int)pos[s];
                    const uint32_t T = tgt[s];
                    grid[c] &= (uint16_t)TYPE_BITS;
                    if (grid[T] == EXIT_EMPTY) {                               // (:101-102)
                        leaves = true;
                        atomicAnd(&alive[s >> 5], ~(1u << (s & 31)));
                    } else {
                        grid[T] |= (uint16_t)(s + 1);
                        pos[s] = (PosT)T;
                    }
                }
            }
            const uint32_t bal = __ballot_sync(0xffffffffu, leaves);
            if (bal != 0u && lane == 0) atomicAdd(&cnt[2], (uint32_t)__popc(bal));
        }
        for (int c = tid; c < claim_words; c += THREADS) claim32[c] = 0u;    // claim counters clean for the next step
        // DFF decay + diffusion reads the bumped field (phase B wrote it before the last barrier) -> other buffer
        if (DFF) dff_decay_diffuse<NBR>(dffA, dffB, H, W, P.c0, P.c1, P.thr, tid, sgeom);
        __syncthreads();

        const int n_exit = (int)cnt[2];
        if (n_exit > 0) {
            n -= n_exit;
            const int nwords = (n_slots + 31) >> 5;
            if (4 * n <= 3 * n_slots && n_slots > 32) {
                // ---- re-pack: slot := alive rank (stable), tgt[] is free to serve as scratch -----
                if (tid < 32) {
                    uint32_t carry = 0;
                    for (int w0 = 0; w0 < nwords; w0 += 32) {
                        const int w = w0 + lane;
                        const uint32_t x = (w < nwords) ? (uint32_t)__popc(alive[w]) : 0u;
                        uint32_t inc = x;
#pragma unroll
                        for (int d = 1; d < 32; d <<= 1) {
                            const uint32_t y = __shfl_up_sync(0xffffffffu, inc, d);
                            if (lane >= d) inc += y;
                        }
                        if (w < nwords) wpre[w] = carry + inc - x;
                        carry += __shfl_sync(0xffffffffu, inc, 31);
                    }
                }
                __syncthreads();
                for (int s = tid; s < n_slots; s += THREADS)
                    if ((alive[s >> 5] >> (s & 31)) & 1u) {
                        const uint32_t rank = wpre[s >> 5] + (uint32_t)__popc(alive[s >> 5] & ((1u << (s & 31)) - 1u));
                        tgt[rank] = pos[s];
                    }
                __syncthreads();
                for (int i = tid; i < n; i += THREADS) {
                    const PosT c = tgt[i];
                    pos[i] = c;
                    grid[c] = (uint16_t)((grid[c] & TYPE_BITS) | (uint32_t)(i + 1));
                }
                for (int w = tid; w <= (n + 31) / 32; w += THREADS) {
                    const int lo = w * 32;
                    alive[w] = (lo + 32 <= n) ? 0xffffffffu : (lo < n ? ((1u << (n - lo)) - 1u) : 0u);
                    wpre[w] = (uint32_t)(lo < n ? lo : n);
                }
                n_slots = n;
                __syncthreads();
            } else {
                // ---- alive ranks changed: refresh the word prefix -------------------------------
                if (tid < 32) {
                    uint32_t carry = 0;
                    for (int w0 = 0; w0 < nwords; w0 += 32) {
                        const int w = w0 + lane;
                        const uint32_t x = (w < nwords) ? (uint32_t)__popc(alive[w]) : 0u;
                        uint32_t inc = x;
#pragma unroll
                        for (int d = 1; d < 32; d <<= 1) {
                            const uint32_t y = __shfl_up_sync(0xffffffffu, inc, d);
                            if (lane >= d) inc += y;
                        }
                        if (w < nwords) wpre[w] = carry + inc - x;
                        carry += __shfl_sync(0xffffffffu, inc, 31);
                    }
                }
                __syncthreads();
            }
        }

        if (DFF) { float* tmp = dffA; dffA = dffB; dffB = tmp; }      // phase D ran alongside C (above)

        // trajectory row: positions after this step, alive-rank order (ffm_core.py:125)
        if (P.traj != nullptr && tl < P.traj_steps) {
            uint32_t* row = P.traj + ((size_t)e * P.traj_steps + tl) * P.n_max;
            for (int s = tid; s < n_slots; s += THREADS)
                if ((alive[s >> 5] >> (s & 31)) & 1u) {
                    const uint32_t rank = wpre[s >> 5] + (uint32_t)__popc(alive[s >> 5] & ((1u << (s & 31)) - 1u));
                    row[rank] = (uint32_t)pos[s];
                }
            if (tid == 0) P.traj_n[(size_t)e * P.traj_steps + tl] = n;
        }
    }

    // ---- epilogue: state back to HBM, alive-rank order -------------------------------------------
    for (int s = tid; s < n_slots; s += THREADS)
        if ((alive[s >> 5] >> (s & 31)) & 1u) {
            const uint32_t rank = wpre[s >> 5] + (uint32_t)__popc(alive[s >> 5] & ((1u << (s & 31)) - 1u));
            gpos[rank] = (uint32_t)pos[s];
        }
    if (DFF && dffA != dff_home) {
        if (tma_fields) {               // shared -> global bulk store of the final DFF
            fence_proxy_async_smem();
            __syncthreads();
            if (tid == 0) { bulk_copy_s2g(dff_home, dffA, dff_bytes); bulk_commit_wait_all(); }
        } else {
            for (int c = tid; c < HW; c += THREADS) dff_home[c] = dffA[c];
        }
    }
    if (tid == 0) {
        P.n_alive[e] = n;
        P.t_done[e] = t0 + tl;
        P.ped_steps[e] += ped_steps;
    }
}

// Probe (parity tests, "move probabilities within 1e-6 relative"): the probability vector every pedestrian of
// the CURRENT state would sample from, computed with the hot path's own arithmetic (move_weights), written in
// SLOT order (neighbours in the reference's order, then "stay"; 0 for non-candidates).
//   kind: 0 = no candidate, no request (:63);  1 = forced exit (one-hot, :66-72);  2 = draws from probs
template <typename S, int NBR, bool DFF>
__global__ void core_move_probs_kernel(const RolloutParams P, double* __restrict__ probs, int32_t* __restrict__ kind) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int e = blockIdx.x, W = P.W, HW = P.HW, G = W + 1;
    uint16_t* grid = reinterpret_cast<uint16_t*>(smem_raw) + G;
    for (int c = threadIdx.x; c < HW + 2 * G; c += blockDim.x) grid[c - G] = P.type_grid[c];
    const int n = P.n_alive[e];
    const uint32_t* gpos = P.pos + (size_t)e * P.n_max;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) grid[gpos[i]] |= (uint16_t)(i + 1);
    __syncthreads();
    const S* score = reinterpret_cast<const S*>(P.score);
    const float* dff = DFF ? P.dff + (size_t)e * HW : nullptr;
    constexpr uint32_t EXIT_EMPTY = TYPE_EXIT << TYPE_SHIFT;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int c = (int)gpos[i];
        double* out = probs + ((size_t)e * P.n_max + i) * (NBR + 1);
#pragma unroll
        for (int k = 0; k <= NBR; ++k) out[k] = 0.0;
        uint32_t m = 0, ex = 0;
#pragma unroll
        for (int k = 0; k < NBR; ++k) {
            const uint32_t g = grid[c + nbr_off<NBR>(k, W)];
            if ((g & OCC_MASK) == 0u) m |= 1u << k;
            if (g == EXIT_EMPTY) ex |= 1u << k;
        }
        int kd = 0;
        if (m != 0u) {
            if (ex != 0u) {
                kd = 1;
                out[__ffs(ex) - 1] = 1.0;
            } else {
                kd = 2;
                int cell[NBR + 1];
                S w[NBR + 1];
                const int ncand = __popc(m) + 1;
                const double tot = move_weights<S, NBR, DFF>(m, ncand, c, W, score, dff, P.kd, cell, w);
                uint32_t mm = m;
#pragma unroll
                for (int j = 0; j <= NBR; ++j)
                    if (j < ncand) {
                        int slot = NBR;
                        if (j < ncand - 1) { slot = __ffs(mm) - 1; mm &= mm - 1u; }
                        out[slot] = (double)w[j] / tot;
                    }
            }
        }
        kind[(size_t)e * P.n_max + i] = kd;
    }
}

}  // namespace ffm
