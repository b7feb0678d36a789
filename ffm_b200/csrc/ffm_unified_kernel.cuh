// Persistent rollout kernel of the unified critic/actor model and of the trained-actor inference
// model: one CTA per episode, all steps in-kernel, tabular V / H in global memory (L2-resident).
//
// Reproduces FloorFieldModelUnified.step() (model/ffm_unified.py:271-606) incl. _encode_state
// (:188-269), _update_critic (:608-670), _get_td_errors (:672-723), _update_actor (:725-777),
// update_dff (:779-798), and FloorFieldModel.step() of model/ffm_trained_core.py:159-331.
//
// Differences from the base model (ffm_core_kernel.cuh) that shape the kernel:
//   * every pedestrian samples every step over a FIXED slot vector (neighbours then "stay"), with
//     a validity mask; scores are taken over ALL slots (walls give -inf) and the max is over all
//     slots (:361-371); "stay" is always valid
//   * an exit among the neighbour slots forces the request and marks will_exit even if the
//     pedestrian later loses the conflict (:334-350)
//   * contested cells always have exactly one winner, floor(u*k)-th claimant (:527-539); all k
//     claimants record k-1 collisions
//   * state = rank code of the four directions + coarse block index (:188-269), dense id
//       (bx*nby + by)*256 + rU*64 + rD*16 + rL*4 + rR
//   * learning: sequential TD(0) over agents in array order on the shared V (:633-665); actor_only
//     recomputes the TD errors with the updated V (:568-574); H[s][a] += alpha_h*delta (:776-777);
//     H is normalised with the min/max over EVERY value in the table incl. zero rows inserted earlier
//     in the same step (:413-439)
//
// Phases per step (block-wide, __syncthreads between):
//   U1  encode state, validity mask, forced exit                       (all pedestrians, parallel)
//   U2  scores -> probabilities -> epsilon-greedy / keyed draw -> target, claim counters
//   B   conflicts (one winner), collision counts, DFF footprints
//   C   apply moves to the owner grid (grid == state_map_next afterwards, :543-546)
//   T1  encode next states (parallel), warm V lines
//   T2  EXACT mode: one thread applies the TD / actor updates in agent order (the reference's
//       semantics);  BATCHED mode: all threads accumulate alpha*delta into delta tables with
//       atomicAdd against the frozen V / H of this launch (synchronous batched TD -- the deltas are
//       summed over episodes / GPUs by the caller between launches)
//   K   stable compaction of the survivors (:601-604), D  DFF update (:606)
#pragma once
#include "ffm_core_kernel.cuh"

namespace ffm {

enum { UMODE_CRITIC = 0, UMODE_ACTOR = 1, UMODE_BOTH = 2, UMODE_TRAINED = 3 };
enum { ULEARN_NONE = 0, ULEARN_EXACT = 1, ULEARN_BATCHED = 2 };

constexpr uint32_t PEDMARK_CELL = (TYPE_WALL << TYPE_SHIFT) | 0x3FFEu;   // map code 1: blocked AND "a pedestrian" to the encoder

struct HStats {           // running min / max over every value of the touched rows of H
    double hmin, hmax;
    int dirty;            // an extreme value moved inwards: rescan before the next use
    int any;              // at least one row exists
};

struct UnifiedDyn { double epsilon; uint32_t episode_base; uint32_t pad; };

struct UnifiedParams {
    int H, W, HW, n_max, B;
    int max_steps;
    int mode, learn;
    int block_size, nby;
    int S;                       // number of states
    const uint16_t* type_grid;   // [HW + 2*(W+1)]
    const void* score;           // [HW] S: -k_S*sff (critic_only)
    float kd, c0, c1, thr;
    double kA, gamma, alpha_v, alpha_h, exit_reward, step_penalty, collision_penalty, epsilon;
    double sff_min, sff_max;     // over the inf->0 float32 SFF (:425-426)
    uint32_t* pos; int32_t* n_alive; int32_t* t_done; unsigned long long* ped_steps;
    float* dff; float* dff_tmp;
    double* V; uint8_t* v_seen;  // [S]
    double* Hm; uint8_t* h_seen; // [S][A], [S]
    double* dV; double* dN; double* dH;   // batched mode: sum of TD errors, visit counts [S]; sum of alpha_h*delta [S][A]
    double* dF;                           // batched mode: > 0 where a V key was touched (read or written) in this sync, or null
    HStats* hstats;
    unsigned long long seed; uint32_t episode_base;
    const double* move_draws; const double* conflict_draws; int draw_steps, draw_first;
    uint32_t* traj; int32_t* traj_n; int traj_steps;
    uint32_t* rec_state; uint8_t* rec_action; float* rec_reward; int32_t* rec_len;   // rollout buffer [B][traj_steps][n_max]
    int32_t* err;                // device validation flag (128: two pedestrians on one cell)
    const struct UnifiedDyn* dyn; // optional device-resident overrides of episode_base / epsilon (CUDA-graph replays: a captured
                                 // launch freezes its by-value parameters, these two change every round)
    uint32_t magic_w, magic_bs;  // ceil(2^32 / W), ceil(2^32 / block_size): x / d = umulhi(x, magic) for x * d < 2^32
};

struct USmemLayout {
    uint32_t score, dffA, dffB, grid, claim, pos, posB, tgt, st, nst, info, orig, origB, td, wcnt, red, misc, total;
};

__host__ __device__ inline USmemLayout make_ulayout(int HW, int W, int n_max, int sizeof_score, bool fields_in_smem) {
    USmemLayout L;
    uint32_t o = 0;
    L.score = o; if (fields_in_smem) o = align16(o + (uint32_t)HW * sizeof_score);
    L.dffA = o;  if (fields_in_smem) o = align16(o + (uint32_t)HW * 4u);
    L.dffB = o;  if (fields_in_smem) o = align16(o + (uint32_t)HW * 4u);
    L.grid = o;  o = align16(o + (uint32_t)(HW + 2 * (W + 1)) * 2u);
    L.claim = o; o = align16(o + (uint32_t)HW);
    L.pos = o;   o = align16(o + (uint32_t)n_max * 4u);
    L.posB = o;  o = align16(o + (uint32_t)n_max * 4u);
    L.tgt = o;   o = align16(o + (uint32_t)n_max * 4u);
    L.st = o;    o = align16(o + (uint32_t)n_max * 4u);
    L.nst = o;   o = align16(o + (uint32_t)n_max * 4u);
    L.info = o;  o = align16(o + (uint32_t)n_max * 4u);
    L.orig = o;  o = align16(o + (uint32_t)n_max * 2u);
    L.origB = o; o = align16(o + (uint32_t)n_max * 2u);
    L.td = o;    o = align16(o + (uint32_t)n_max * 8u);
    L.wcnt = o;  o = align16(o + (uint32_t)(n_max / 32 + 2) * 4u);
    L.red = o;   o = align16(o + 32u * 8u * 2u + 32u * 4u);
    L.misc = o;  o = align16(o + 64u);
    L.total = o;
    return L;
}

// info word of a pedestrian this step
constexpr uint32_t INFO_SLOT_MASK = 0xFu;          // chosen slot
constexpr uint32_t INFO_VALID_SHIFT = 4;           // bits 4..12: validity of the 9 (5) slots
constexpr uint32_t INFO_EXIT = 1u << 13;           // will_exit
constexpr uint32_t INFO_MOVED = 1u << 14;          // request granted
constexpr uint32_t INFO_COLL_SHIFT = 16;           // collision count (k-1), 4 bits

__device__ __forceinline__ bool cell_blocked(uint32_t g) { return (g & OCC_MASK) != 0u; }                   // wall(2) or ped(1)
__device__ __forceinline__ bool cell_is_ped(uint32_t g) { const uint32_t o = g & OCC_MASK; return o != 0u && o != OCC_MASK; }

// _encode_state (ffm_unified.py:188-269) on the owner grid; (r, col) = coordinates of linear cell c
__device__ __forceinline__ uint32_t encode_state(const uint16_t* grid, int c, int r, int col, int H, int W, uint32_t magic_bs, int nby) {
    uint32_t code = 0;
#pragma unroll
    for (int d = 0; d < 4; ++d) {                     // up, down, left, right (:209)
        const int dr = d == 0 ? -1 : (d == 1 ? 1 : 0), dc = d == 2 ? -1 : (d == 3 ? 1 : 0);
        const int o1 = dr * W + dc;
        uint32_t rank = 3;
        const int r1 = r + dr, c1 = col + dc;
        if (r1 < 0 || r1 >= H || c1 < 0 || c1 >= W) {
            rank = 0;                                 // :252-254
        } else if (cell_blocked(grid[c + o1])) {
            rank = 0;                                 // :219-220
        } else {
            bool ped = false;                         // forward diagonals (:224-237)
            if (dr != 0) {
                if (c1 - 1 >= 0) ped |= cell_is_ped(grid[c + o1 - 1]);
                if (c1 + 1 < W) ped |= cell_is_ped(grid[c + o1 + 1]);
            } else {
                if (r1 - 1 >= 0) ped |= cell_is_ped(grid[c + o1 - W]);
                if (r1 + 1 < H) ped |= cell_is_ped(grid[c + o1 + W]);
            }
            if (ped) {
                rank = 1;                             // :239-240
            } else {
                const int r2 = r + 2 * dr, c2 = col + 2 * dc;
                if (r2 < 0 || r2 >= H || c2 < 0 || c2 >= W) rank = 2;              // :249-251
                else if (cell_blocked(grid[c + 2 * o1])) rank = 2;                 // :245-248
            }
        }
        code = code * 4u + rank;
    }
    // block index (x // bs, y // bs) without integer divisions (coordinates are < 4096, bs >= 1)
    const uint32_t bx = magic_bs ? __umulhi((uint32_t)r, magic_bs) : (uint32_t)r;      // magic_bs == 0: block_size 1
    const uint32_t by = magic_bs ? __umulhi((uint32_t)col, magic_bs) : (uint32_t)col;
    return (bx * (uint32_t)nby + by) * 256u + code;
}

// reward of one agent-step (:636-648): Python float arithmetic in the reference's order
__device__ __forceinline__ double agent_reward(const UnifiedParams& P, uint32_t w) {
    double rew = P.step_penalty;
    if (w & INFO_EXIT) rew = __dadd_rn(rew, P.exit_reward);
    return __dadd_rn(rew, __dmul_rn((double)((w >> INFO_COLL_SHIFT) & 0xFu), P.collision_penalty));
}

// ACTOR = false: critic_only (scores from the SFF in its own dtype, no H table): the actor paths (float64 scoring, H rows,
// extremes) compile out, which is worth ~30 registers per thread (c4: 10 instead of 8 CTAs per SM, +10 %; capping the registers further to reach 12
// CTAs measured slower: 4.09e9 vs 4.31e9)
template <typename S, int NBR, bool FIELDS_IN_SMEM, int THREADS, bool ACTOR>
__global__ void __launch_bounds__(THREADS)
ffm_unified_rollout_kernel(const UnifiedParams P) {
    constexpr int A = NBR + 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int e = blockIdx.x;
    const int W = P.W, HW = P.HW, H = P.H;
    const int G = W + 1;
    const USmemLayout L = make_ulayout(HW, W, P.n_max, (int)sizeof(S), FIELDS_IN_SMEM);
    constexpr uint32_t EXIT_EMPTY = TYPE_EXIT << TYPE_SHIFT;
    constexpr uint32_t NO_STATE = 0xFFFFFFFFu;
    const double DINF = __longlong_as_double(0x7ff0000000000000LL);

    uint16_t* grid = reinterpret_cast<uint16_t*>(smem_raw + L.grid) + G;
    uint8_t* claim = smem_raw + L.claim;
    uint32_t* claim32 = reinterpret_cast<uint32_t*>(smem_raw + L.claim);
    uint32_t* pos = reinterpret_cast<uint32_t*>(smem_raw + L.pos);
    uint32_t* posB = reinterpret_cast<uint32_t*>(smem_raw + L.posB);
    uint32_t* tgt = reinterpret_cast<uint32_t*>(smem_raw + L.tgt);
    uint32_t* st = reinterpret_cast<uint32_t*>(smem_raw + L.st);
    uint32_t* nst = reinterpret_cast<uint32_t*>(smem_raw + L.nst);
    uint32_t* info = reinterpret_cast<uint32_t*>(smem_raw + L.info);
    uint16_t* orig = reinterpret_cast<uint16_t*>(smem_raw + L.orig);    // column of the rollout buffer (index at launch)
    uint16_t* origB = reinterpret_cast<uint16_t*>(smem_raw + L.origB);
    double* tdv = reinterpret_cast<double*>(smem_raw + L.td);
    uint32_t* wcnt = reinterpret_cast<uint32_t*>(smem_raw + L.wcnt);
    double* red_lo = reinterpret_cast<double*>(smem_raw + L.red);
    double* red_hi = red_lo + 32;
    int* red_any = reinterpret_cast<int*>(red_hi + 32);
    int* misc = reinterpret_cast<int*>(smem_raw + L.misc);               // [0] first agent inserting a new H row
    double* dmisc = reinterpret_cast<double*>(smem_raw + L.misc + 16);   // [0] hmin, [1] hmax at step start

    const S* score;
    float* dffA; float* dffB;
    float* dff_home = P.dff + (size_t)e * HW;
    if (FIELDS_IN_SMEM) {
        S* s_sm = reinterpret_cast<S*>(smem_raw + L.score);
        const S* s_g = reinterpret_cast<const S*>(P.score);
        for (int c = tid; c < HW; c += THREADS) s_sm[c] = s_g[c];
        score = s_sm;
        dffA = reinterpret_cast<float*>(smem_raw + L.dffA);
        dffB = reinterpret_cast<float*>(smem_raw + L.dffB);
        for (int c = tid; c < HW; c += THREADS) dffA[c] = dff_home[c];
    } else {
        score = reinterpret_cast<const S*>(P.score);
        dffA = dff_home;
        dffB = P.dff_tmp + (size_t)e * HW;
    }
    for (int c = tid; c < HW + 2 * G; c += THREADS) grid[c - G] = P.type_grid[c];
    for (int c = tid; c < (HW + 3) / 4; c += THREADS) claim32[c] = 0u;
    int n = P.n_alive[e];
    const int t0 = P.t_done[e];
    uint32_t* gpos = P.pos + (size_t)e * P.n_max;
    const bool recording = P.rec_reward != nullptr || P.rec_state != nullptr || P.rec_action != nullptr;
    for (int i = tid; i < n; i += THREADS) { pos[i] = gpos[i]; orig[i] = (uint16_t)i; }
    if (P.rec_len != nullptr) for (int i = tid; i < P.n_max; i += THREADS) P.rec_len[(size_t)e * P.n_max + i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += THREADS) grid[pos[i]] |= (uint16_t)(i + 1);
    __syncthreads();
    for (int i = tid; i < n; i += THREADS)      // duplicates OR their ids together: somebody reads back a foreign id
        if ((grid[pos[i]] & OCC_MASK) != (uint32_t)(i + 1) && P.err != nullptr) atomicOr(P.err, 128);

    const uint32_t episode = (P.dyn != nullptr ? P.dyn->episode_base : P.episode_base) + (uint32_t)e;
    const double epsilon = P.dyn != nullptr ? P.dyn->epsilon : P.epsilon;
    const double* mv_draws = P.move_draws ? P.move_draws + (size_t)e * P.draw_steps * P.n_max : nullptr;
    const double* cf_draws = P.conflict_draws ? P.conflict_draws + (size_t)e * P.draw_steps * HW * 2 : nullptr;
    const bool actor = ACTOR;                                                    // == (P.mode != UMODE_CRITIC), checked by the host
    // The tables are written in place only by FFM_LEARN_EXACT (one episode = one CTA per handle).  With frozen tables
    // (NONE) or batched learning many CTAs share them: rows are never inserted by lookups, the extremes of H are the
    // ones found at launch (ffm_rollout refreshes them beforehand when they are stale) and every table write goes
    // through the delta tables.
    const bool exact = P.learn == ULEARN_EXACT;
    const bool trains_h = ACTOR && (P.mode == UMODE_ACTOR || P.mode == UMODE_BOTH);
    const bool inserts_rows = trains_h && exact;                                 // H lookups insert zero rows (:405-410)
    const bool learn_actor = trains_h && P.learn != ULEARN_NONE;
    const bool track_stats = actor && exact;

    unsigned long long ped_steps = 0;
    int tl = 0;
    const StencilGeom sgeom = make_stencil_geom(H, W, tid, THREADS);   // DFF stencil geometry, once (not per step)
    for (; tl < P.max_steps && n > 0; ++tl) {
        const uint32_t t = (uint32_t)(t0 + tl);
        ped_steps += (unsigned long long)n;
        const int di = (int)t - P.draw_first;
        const bool inj = di >= 0 && di < P.draw_steps;

        // ---- extremes of the H table at step start (rescan when an extreme value moved inwards) --
        if (actor) {
            if (track_stats) {                 // uniform across the CTA: tid 0 publishes the flag through shared memory
                if (tid == 0) misc[1] = P.hstats->dirty;
                __syncthreads();
            }
            if (track_stats && misc[1]) {
                double lo = DINF, hi = -DINF;
                int any = 0;
                for (int s = tid; s < P.S; s += THREADS)
                    if (P.h_seen[s]) {
                        any = 1;
#pragma unroll
                        for (int a = 0; a < A; ++a) { const double v = P.Hm[(size_t)s * A + a]; lo = fmin(lo, v); hi = fmax(hi, v); }
                    }
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, d));
                    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, d));
                    any |= __shfl_xor_sync(0xffffffffu, any, d);
                }
                if (lane == 0) { red_lo[warp] = lo; red_hi[warp] = hi; red_any[warp] = any; }
                __syncthreads();
                if (tid == 0) {
                    for (int w = 1; w < THREADS / 32; ++w) { lo = fmin(lo, red_lo[w]); hi = fmax(hi, red_hi[w]); any |= red_any[w]; }
                    P.hstats->hmin = lo; P.hstats->hmax = hi; P.hstats->any = any; P.hstats->dirty = 0;
                }
                __syncthreads();
            }
            if (tid == 0) {
                dmisc[0] = P.hstats->any ? P.hstats->hmin : DINF;
                dmisc[1] = P.hstats->any ? P.hstats->hmax : -DINF;
                misc[0] = 0x7fffffff;
            }
        }
        __syncthreads();

        // ================= U1: state, validity, forced exit =====================================
        for (int i = tid; i < n; i += THREADS) {
            const int c = (int)pos[i];
            const int r = (int)__umulhi((uint32_t)c, P.magic_w), col = c - r * W;
            const uint32_t sid = encode_state(grid, c, r, col, H, W, P.magic_bs, P.nby);     // :293
            st[i] = sid;
            uint32_t valid = 1u << NBR, ex = 0;                       // "stay" is always valid (:318-319)
#pragma unroll
            for (int k = 0; k < NBR; ++k) {
                const uint32_t g = grid[c + nbr_off<NBR>(k, W)];
                if ((g & OCC_MASK) == 0u) valid |= 1u << k;           // in bounds, map 0/3, unoccupied (:304-323)
                if ((g >> TYPE_SHIFT) == TYPE_EXIT) ex |= 1u << k;    // exit among the neighbour slots (:327-332)
            }
            uint32_t w = valid << INFO_VALID_SHIFT;
            uint32_t target = 0xFFFFFFFFu;
            if (ex != 0u) {                                           // :334-350
                const int k = __ffs(ex) - 1;
                target = (uint32_t)(c + nbr_off_rt<NBR>(k, W));
                w |= INFO_EXIT | (uint32_t)k;
                atomicAdd(&claim32[target >> 2], 1u << (8 * (target & 3u)));
            } else if (inserts_rows && !P.h_seen[sid]) {
                atomicMin(&misc[0], i);                               // this agent's lookup inserts a zero row
            }
            tgt[i] = target;
            info[i] = w;
        }
        __syncthreads();

        // ================= U2: scores, probabilities, draw =====================================
        const int first_new = actor ? misc[0] : 0x7fffffff;
        for (int i = tid; i < n; i += THREADS) {
            uint32_t w = info[i];
            if (w & INFO_EXIT) continue;
            const int c = (int)pos[i];
            const uint32_t sid = st[i];
            const uint32_t valid = (w >> INFO_VALID_SHIFT) & ((1u << A) - 1u);
            double e_[A];
            double tot = 0.0;
            if (!actor) {
                // critic_only: -k_S*sff + k_D*dff over all slots, max over all slots (:355-368)
                S sc[A];
                S mx = neg_inf<S>();
#pragma unroll
                for (int k = 0; k < A; ++k) {
                    const int cc = (k == NBR) ? c : c + nbr_off<NBR>(k, W);
                    sc[k] = add_rn(score[cc], (S)mul_rn(P.kd, dffA[cc]));
                    mx = max_t(mx, sc[k]);
                }
#pragma unroll
                for (int k = 0; k < A; ++k) {
                    const S pk = ((valid >> k) & 1u) ? exp_t(add_rn(sc[k], -mx)) : (S)0;   // :368,371
                    e_[k] = (double)pk;
                    tot += e_[k];
                }
            } else if (P.mode == UMODE_TRAINED) {
                // ffm_trained_core.py:229-284, all float32
                float h[A];
                const bool have = P.h_seen[sid] != 0;
#pragma unroll
                for (int k = 0; k < A; ++k) h[k] = have ? (float)P.Hm[(size_t)sid * A + k] : 0.0f;
                const double hmin = dmisc[0], hmax = dmisc[1];
                if (hmax - hmin > 1e-6) {                               // :259
                    const float hm = (float)hmax, den = (float)(hmax - hmin), rng = (float)(P.sff_max - P.sff_min), smn = (float)P.sff_min;
#pragma unroll
                    for (int k = 0; k < A; ++k)
                        h[k] = __fadd_rn(__fmul_rn(__fdiv_rn(__fadd_rn(hm, -h[k]), den), rng), smn);   // :260-263
                }
                float sc[A], mx = neg_inf<float>();
                const float nka = (float)(-P.kA);
#pragma unroll
                for (int k = 0; k < A; ++k) {
                    const int cc = (k == NBR) ? c : c + nbr_off<NBR>(k, W);
                    sc[k] = __fadd_rn(__fmul_rn(nka, h[k]), __fmul_rn(P.kd, dffA[cc]));           // :267-270
                    mx = fmaxf(mx, sc[k]);
                }
#pragma unroll
                for (int k = 0; k < A; ++k) {
                    const float pk = ((valid >> k) & 1u) ? expf(__fadd_rn(sc[k], -mx)) : 0.0f;
                    e_[k] = (double)pk;
                    tot += e_[k];
                }
            } else {
                // actor_only / both: H row (zero row inserted if absent, :405-411), float64 (:411-445)
                double h[A];
                const bool have = P.h_seen[sid] != 0;
#pragma unroll
                for (int k = 0; k < A; ++k) h[k] = have ? P.Hm[(size_t)sid * A + k] : 0.0;
                // table extremes seen by this agent: step-start extremes plus the zero rows inserted by
                // agents up to and including this one (:413-426)
                double hmin = dmisc[0], hmax = dmisc[1];
                if (first_new <= i || (!exact && !have)) { hmin = fmin(hmin, 0.0); hmax = fmax(hmax, 0.0); }
                if (hmax - hmin > 1e-6) {                               // :434
                    const double den = hmax - hmin, rng = P.sff_max - P.sff_min;
#pragma unroll
                    for (int k = 0; k < A; ++k)
                        h[k] = __dadd_rn(__dmul_rn(__ddiv_rn(__dadd_rn(hmax, -h[k]), den), rng), P.sff_min);   // :435-438
                }
                double sc[A], mx = -DINF;
#pragma unroll
                for (int k = 0; k < A; ++k) {
                    const int cc = (k == NBR) ? c : c + nbr_off<NBR>(k, W);
                    sc[k] = __dadd_rn(__dmul_rn(-P.kA, h[k]), (double)__fmul_rn(P.kd, dffA[cc]));   // :442-445
                    mx = fmax(mx, sc[k]);
                }
#pragma unroll
                for (int k = 0; k < A; ++k) {
                    e_[k] = ((valid >> k) & 1u) ? exp(__dadd_rn(sc[k], -mx)) : 0.0;                 // :459,462
                    tot += e_[k];
                }
            }
            if (!(isfinite(tot) && tot > 0.0)) {                        // uniform over the valid slots (:377-384)
                tot = 0.0;
#pragma unroll
                for (int k = 0; k < A; ++k) { e_[k] = ((valid >> k) & 1u) ? 1.0 : 0.0; tot += e_[k]; }
            }
            int slot = -1;
            if (learn_actor && epsilon > 0.0) {                       // epsilon-greedy (:478-495)
                const Draw2 d = draw2(P.seed, episode, t, STREAM_EPS, (uint32_t)i);
                if (d.u0 < epsilon) {
                    const int nv = __popc(valid);
                    slot = (int)__fns(valid, 0, (int)(d.u1 * (double)nv) + 1);
                }
            }
            if (slot < 0) {
                const double u = (inj && mv_draws) ? mv_draws[(size_t)di * P.n_max + i]
                                                   : draw_u0(P.seed, episode, t, STREAM_MOVE, (uint32_t)i);
                const double thresh = u * tot;
                double run = 0.0;
                slot = NBR;
                bool done = false;
#pragma unroll
                for (int k = 0; k < NBR; ++k)
                    if (!done) {
                        run += e_[k];
                        if (run > thresh) { slot = k; done = true; }     // zero-width (invalid) slots are never hit
                    }
            }
            const uint32_t target = (slot == NBR) ? (uint32_t)c : (uint32_t)(c + nbr_off_rt<NBR>(slot, W));
            tgt[i] = target;
            info[i] = w | (uint32_t)slot;
            if (target != (uint32_t)c) atomicAdd(&claim32[target >> 2], 1u << (8 * (target & 3u)));
            if (inserts_rows) P.h_seen[sid] = 1;                         // the lookup inserted the row
        }
        __syncthreads();

        // ================= B: conflicts (one winner per contested cell, :520-539) ===============
        for (int i = tid; i < n; i += THREADS) {
            const int c = (int)pos[i];
            const uint32_t T = tgt[i];
            uint32_t w = info[i];
            bool moved = true;
            if (T != (uint32_t)c) {
                const int k = (int)claim[T];
                if (k > 1) {
                    int r = 0;
#pragma unroll
                    for (int q = 0; q < NBR; ++q) {
                        const uint32_t o = (grid[(int)T + nbr_off<NBR>(q, W)] & OCC_MASK) - 1u;
                        if (o < (uint32_t)i && tgt[o] == T) ++r;
                    }
                    const double u1 = (inj && cf_draws) ? cf_draws[((size_t)di * HW + T) * 2 + 1]
                                                        : draw2(P.seed, episode, t, STREAM_CONFLICT, T).u1;
                    moved = (int)(u1 * (double)k) == r;                  // random.choice(agents) (:530)
                    w |= (uint32_t)(k - 1) << INFO_COLL_SHIFT;           // every claimant records k-1 (:529,535-539)
                }
            }
            if (moved) {
                w |= INFO_MOVED;
                dffA[c] = __fadd_rn(dffA[c], 1.0f);                      // :523-525, :532-534
            }
            info[i] = w;
        }
        if (tid == 0 && track_stats && inserts_rows && misc[0] != 0x7fffffff) {
            // zero rows were inserted by this step's lookups: they take part in the extremes from now on
            HStats* hs = P.hstats;
            if (!hs->any) { hs->any = 1; hs->hmin = 0.0; hs->hmax = 0.0; }
            else { hs->hmin = fmin(hs->hmin, 0.0); hs->hmax = fmax(hs->hmax, 0.0); }
        }
        __syncthreads();

        // ================= C: apply moves -> grid becomes state_map_next (:543-546) =============
        for (int i = tid; i < n; i += THREADS) {
            const int c = (int)pos[i];
            const uint32_t T = tgt[i];
            if (T != (uint32_t)c) claim[T] = 0;
            uint32_t nc = (uint32_t)c;
            if ((info[i] & INFO_MOVED) && T != (uint32_t)c) {
                grid[c] &= (uint16_t)TYPE_BITS;
                if (grid[T] != EXIT_EMPTY) grid[T] |= (uint16_t)(i + 1);   // pedestrians on exits are not marked
                nc = T;
            }
            posB[i] = nc;
            if (recording && tl < P.traj_steps) {                        // rollout buffer row: coalesced over pedestrians
                const size_t at = ((size_t)e * P.traj_steps + tl) * P.n_max + orig[i];
                const uint32_t w = info[i];
                if (P.rec_state) P.rec_state[at] = st[i];
                if (P.rec_action) P.rec_action[at] = (uint8_t)(w & INFO_SLOT_MASK);
                if (P.rec_reward) P.rec_reward[at] = (float)agent_reward(P, w);
                if (P.rec_len) P.rec_len[(size_t)e * P.n_max + orig[i]] = tl + 1;
            }
        }
        __syncthreads();

        // ================= T: learning ==========================================================
        if (P.learn != ULEARN_NONE && P.mode != UMODE_TRAINED) {
            // T1: next states (parallel); reads insert the keys (defaultdict, :658,661)
            for (int i = tid; i < n; i += THREADS) {
                uint32_t ns = NO_STATE;
                if (!(info[i] & INFO_EXIT)) {                            // :651-658
                    const int c = (int)posB[i];
                    const int r = (int)__umulhi((uint32_t)c, P.magic_w), col = c - r * W;
                    ns = encode_state(grid, c, r, col, H, W, P.magic_bs, P.nby);
                    if (exact) P.v_seen[ns] = 1; else if (P.dF) P.dF[ns] = 1.0;
                }
                nst[i] = ns;
                if (exact) P.v_seen[st[i]] = 1; else if (P.dF) P.dF[st[i]] = 1.0;
            }
            __syncthreads();
            if (P.learn == ULEARN_EXACT) {
                // T2: the reference's sequential loop over agents on the shared table (:633-665)
                if (tid == 0) {
                    double* V = P.V;              // one CTA owns the tables in this mode: plain (L1-cached) accesses, program order per thread
                    for (int i = 0; i < n; ++i) {
                        const uint32_t w = info[i];
                        const double rew = agent_reward(P, w);
                        const double v_next = (w & INFO_EXIT) ? 0.0 : V[nst[i]];
                        const double v_cur = V[st[i]];
                        const double td = __dadd_rn(__dadd_rn(rew, __dmul_rn(P.gamma, v_next)), -v_cur);   // :662
                        V[st[i]] = __dadd_rn(v_cur, __dmul_rn(P.alpha_v, td));                             // :665
                        tdv[i] = td;                                                                      // "both": :577-584
                    }
                }
                __syncthreads();
                if (P.mode == UMODE_ACTOR) {
                    // actor_only recomputes the TD errors with the UPDATED table (:568-574); V is fixed now
                    for (int i = tid; i < n; i += THREADS) {
                        const uint32_t w = info[i];
                        const double v_next = (w & INFO_EXIT) ? 0.0 : P.V[nst[i]];
                        tdv[i] = __dadd_rn(__dadd_rn(agent_reward(P, w), __dmul_rn(P.gamma, v_next)), -P.V[st[i]]);
                    }
                    __syncthreads();
                }
                if (learn_actor && tid == 0) {
                    // _update_actor (:745-777) in agent order; keeps the table extremes current
                    HStats* hs = P.hstats;
                    double hmin = hs->hmin, hmax = hs->hmax;
                    int any = hs->any, dirty = hs->dirty;
                    double* Hm = P.Hm;
                    for (int i = 0; i < n; ++i) {
                        const uint32_t w = info[i], sid = st[i];
                        if (!P.h_seen[sid]) {                            // row inserted as zeros (:769-773)
                            P.h_seen[sid] = 1;
                            if (!any) { any = 1; hmin = 0.0; hmax = 0.0; } else { hmin = fmin(hmin, 0.0); hmax = fmax(hmax, 0.0); }
                        }
                        const uint32_t a = w & INFO_SLOT_MASK;
                        if ((w >> (INFO_VALID_SHIFT + a)) & 1u) {         // valid_mask[chosen_idx] (:776)
                            const double old = Hm[(size_t)sid * A + a];
                            const double nw = __dadd_rn(old, __dmul_rn(P.alpha_h, tdv[i]));   // :777
                            Hm[(size_t)sid * A + a] = nw;
                            if (nw < hmin) hmin = nw; else if (old == hmin && nw > old) dirty = 1;
                            if (nw > hmax) hmax = nw; else if (old == hmax && nw < old) dirty = 1;
                        }
                    }
                    hs->hmin = hmin; hs->hmax = hmax; hs->any = any; hs->dirty = dirty;
                }
            } else {
                // BATCHED: synchronous TD against the frozen tables of this launch; deltas are summed with
                // atomics and applied (after the cross-GPU all-reduce) by ffm_unified_apply_deltas
                for (int i = tid; i < n; i += THREADS) {
                    const uint32_t w = info[i], sid = st[i];
                    const double v_next = (w & INFO_EXIT) ? 0.0 : P.V[nst[i]];
                    const double td = __dadd_rn(__dadd_rn(agent_reward(P, w), __dmul_rn(P.gamma, v_next)), -P.V[sid]);
                    atomicAdd(&P.dV[sid], td);
                    atomicAdd(&P.dN[sid], 1.0);
                    if (learn_actor) {                                   // the row becomes present when the deltas are applied
                        const uint32_t a = w & INFO_SLOT_MASK;
                        if ((w >> (INFO_VALID_SHIFT + a)) & 1u) atomicAdd(&P.dH[(size_t)sid * A + a], __dmul_rn(P.alpha_h, td));
                    }
                }
            }
            __syncthreads();
        }

        // ================= K: exit removal, stable compaction (:601-604) ========================
        for (int base = 0; base < n; base += THREADS) {
            const int i = base + tid;
            const bool kept = i < n && grid[posB[i]] != EXIT_EMPTY;
            const uint32_t bal = __ballot_sync(0xffffffffu, kept);
            if (lane == 0 && i < n) wcnt[i >> 5] = __popc(bal);
        }
        __syncthreads();
        const int ngroups = (n + 31) >> 5;
        int n_new = 0;
        for (int base = 0; base < n; base += THREADS) {
            const int i = base + tid;
            const int v = i >> 5;
            int before = 0, total = 0;
            for (int w0 = 0; w0 < ngroups; w0 += 32) {
                const int wq = w0 + lane;
                const int x = (wq < ngroups) ? (int)wcnt[wq] : 0;
                int xb = (wq < v) ? x : 0, xt = x;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    xb += __shfl_xor_sync(0xffffffffu, xb, d);
                    xt += __shfl_xor_sync(0xffffffffu, xt, d);
                }
                before += xb;
                total += xt;
            }
            n_new = total;
            const bool kept = i < n && grid[posB[i]] != EXIT_EMPTY;
            const uint32_t bal = __ballot_sync(0xffffffffu, kept);
            if (kept) {
                const int ni = before + __popc(bal & ((1u << lane) - 1u));
                const uint32_t c = posB[i];
                pos[ni] = c;
                origB[ni] = orig[i];
                grid[c] = (uint16_t)((grid[c] & TYPE_BITS) | (uint32_t)(ni + 1));
            }
        }
        { uint16_t* tmpo = orig; orig = origB; origB = tmpo; }
        // ================= D: DFF decay + diffusion (:779-798) ==================================
        dff_decay_diffuse<NBR>(dffA, dffB, H, W, P.c0, P.c1, P.thr, tid, sgeom);
        { float* tmp = dffA; dffA = dffB; dffB = tmp; }
        __syncthreads();
        n = n_new;
        if (P.traj != nullptr && tl < P.traj_steps) {
            uint32_t* row = P.traj + ((size_t)e * P.traj_steps + tl) * P.n_max;
            for (int i = tid; i < n; i += THREADS) row[i] = pos[i];
            if (tid == 0) P.traj_n[(size_t)e * P.traj_steps + tl] = n;
        }
    }

    for (int i = tid; i < n; i += THREADS) gpos[i] = pos[i];
    if (dffA != dff_home)
        for (int c = tid; c < HW; c += THREADS) dff_home[c] = dffA[c];
    if (tid == 0) {
        P.n_alive[e] = n;
        P.t_done[e] = t0 + tl;
        P.ped_steps[e] += ped_steps;
    }
}

// Discounted returns of a rollout buffer: G[b][t][n] = r[b][t][n] + gamma * G[b][t+1][n] over each path
// (ffm_learning_core.py:262-278, 350-355: `G = r + self.gamma * G`, float64).  HBM-bound streaming kernel:
// one thread owns 4 adjacent pedestrians of an episode, reads float4 reward rows and writes 2 x double2
// return rows walking the time axis backwards; rows are contiguous over pedestrians -> fully coalesced.
static __global__ void __launch_bounds__(256)
rollout_returns_kernel(const float* __restrict__ reward, const int32_t* __restrict__ len, int B, int T, int N, double gamma,
                       double* __restrict__ G) {
    const int groups = (N + 3) / 4;
    const long long total = (long long)B * groups;
    for (long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(x / groups), n0 = (int)(x - (long long)b * groups) * 4;
        const bool vec = (N % 4 == 0);
        int L[4];
        int Lmax = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            L[k] = (n0 + k < N) ? min(len[(size_t)b * N + n0 + k], T) : 0;
            Lmax = max(Lmax, L[k]);
        }
        double g[4] = {0.0, 0.0, 0.0, 0.0};
        const size_t base = (size_t)b * T * N + n0;
        for (int t = T - 1; t >= Lmax; --t) {          // beyond every path's end: zeros
            if (vec) {
                reinterpret_cast<double2*>(G + base + (size_t)t * N)[0] = make_double2(0.0, 0.0);
                reinterpret_cast<double2*>(G + base + (size_t)t * N)[1] = make_double2(0.0, 0.0);
            } else {
                for (int k = 0; k < 4 && n0 + k < N; ++k) G[base + (size_t)t * N + k] = 0.0;
            }
        }
#pragma unroll 4
        for (int t = Lmax - 1; t >= 0; --t) {
            float r[4];
            if (vec) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(reward + base + (size_t)t * N));
                r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
            } else {
                for (int k = 0; k < 4; ++k) r[k] = (n0 + k < N) ? reward[base + (size_t)t * N + k] : 0.0f;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) g[k] = (t < L[k]) ? __dadd_rn((double)r[k], __dmul_rn(gamma, g[k])) : 0.0;
            if (vec) {
                reinterpret_cast<double2*>(G + base + (size_t)t * N)[0] = make_double2(g[0], g[1]);
                reinterpret_cast<double2*>(G + base + (size_t)t * N)[1] = make_double2(g[2], g[3]);
            } else {
                for (int k = 0; k < 4 && n0 + k < N; ++k) G[base + (size_t)t * N + k] = g[k];
            }
        }
    }
}

// Batched learning, between launches.  A state visited n times in the sync with TD errors d_1..d_n against
// the frozen table receives V += (1 - (1 - alpha_v)^n) * mean(d): what n sequential updates towards the same
// targets would give (a plain sum would multiply the step size by n and diverge for well-visited states).
// H accumulates alpha_h * delta like the reference (:777).  Keys touched in the sync (dF > 0) become present in V,
// visited states (dN > 0) get their H row when the actor learns.  Deltas zeroed, extremes of H recomputed.
// With dV == nullptr only the extremes are recomputed (tables loaded by the caller: ffm_tables_set).
static __global__ void unified_apply_deltas_kernel(double* V, double* dV, double* dN, double* dF, double alpha_v, double* Hm, double* dH,
                                            uint8_t* h_seen, uint8_t* v_seen, int S, int A,
                                            double* block_lo, double* block_hi, int* block_any) {
    const double DINF = __longlong_as_double(0x7ff0000000000000LL);
    double lo = DINF, hi = -DINF;
    int any = 0;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < S; s += gridDim.x * blockDim.x) {
        bool visited = false;
        if (dV != nullptr) {
            const double cnt = dN[s];
            visited = cnt > 0.0;
            if (visited) {
                V[s] += (1.0 - pow(1.0 - alpha_v, cnt)) * (dV[s] / cnt);
                dV[s] = 0.0;
                dN[s] = 0.0;
            }
            if (dF != nullptr && dF[s] != 0.0) { v_seen[s] = 1; dF[s] = 0.0; }
        }
        if (Hm != nullptr) {
            if (visited && dH != nullptr) {
                h_seen[s] = 1;
                for (int a = 0; a < A; ++a) {
                    Hm[(size_t)s * A + a] += dH[(size_t)s * A + a];
                    dH[(size_t)s * A + a] = 0.0;
                }
            }
            if (h_seen[s]) {
                any = 1;
                for (int a = 0; a < A; ++a) { const double v = Hm[(size_t)s * A + a]; lo = fmin(lo, v); hi = fmax(hi, v); }
            }
        }
    }
    __shared__ double slo[32], shi[32];
    __shared__ int sany[32];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, d));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, d));
        any |= __shfl_xor_sync(0xffffffffu, any, d);
    }
    if ((threadIdx.x & 31) == 0) { slo[threadIdx.x >> 5] = lo; shi[threadIdx.x >> 5] = hi; sany[threadIdx.x >> 5] = any; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)blockDim.x / 32; ++w) { lo = fmin(lo, slo[w]); hi = fmax(hi, shi[w]); any |= sany[w]; }
        block_lo[blockIdx.x] = lo; block_hi[blockIdx.x] = hi; block_any[blockIdx.x] = any;
    }
}

static __global__ void unified_finish_stats_kernel(HStats* hstats, const double* block_lo, const double* block_hi, const int* block_any, int nblocks) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const double DINF = __longlong_as_double(0x7ff0000000000000LL);
        double lo = DINF, hi = -DINF;
        int any = 0;
        for (int b = 0; b < nblocks; ++b) { lo = fmin(lo, block_lo[b]); hi = fmax(hi, block_hi[b]); any |= block_any[b]; }
        hstats->hmin = lo; hstats->hmax = hi; hstats->any = any; hstats->dirty = 0;
    }
}

}  // namespace ffm
