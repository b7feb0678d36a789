// Persistent rollout kernel of the target-centric Monte-Carlo Q-learning model: one CTA per episode.
//
// Reproduces FloorFieldModel.step(beta) of model/ffm_learning_core.py:145-285 with _combined3x3_at_target
// (:115-140), the reverse Monte-Carlo backup (:262-278), _update_dff (:307-321) and finalize_timeouts (:326-360).
//
// What differs from the other models and shapes the kernel:
//   * von Neumann moves + STOP; every candidate's logit needs ITS OWN state: the 3x3 window around the TARGET
//     (map codes, OOB = 2, +1 on free cells that are occupied -- the agent itself included, :132-135) and the
//     coarse block (tx//3, ty//3).  64-bit key: block * 4^9 + sum v_i * 4^i (row-major window); the table is an
//     open-addressing HASH TABLE in HBM (keys u64, rows float32 [5]; linear probing, insertion by atomicCAS) -- the
//     reference's dict: rows are created by _ensure_qvec (:289-291, find-or-insert), never by the read path
//     (:190-191, find only); only visited states take memory, any map size
//   * logit = beta*(-k_S*sff) + k_D*dff + (1-beta)*k_Q*Q[s][from_dir] in Python floats (:193), also for STOP
//   * contested cells always have one winner; losers overwrite their last reward with -collision_penalty
//     (:253-257); only real moves leave a DFF footprint (:235,247)
//   * per-agent paths (state, action, reward) are kept for the whole episode (column = index at launch); an
//     agent that reaches an exit gets exit_reward on its last record and its path is backed up in reverse:
//     G = r + gamma*G (Python floats), Q[s][a] += alpha*(G - Q[s][a]) in float32 (NEP 50 weak scalars), arrivals
//     in descending index order (:263-267); at the step cap every remaining agent appends a timeout record and
//     is backed up in index order (:326-360)
//   * the DFF update sums the 8 shifted copies of the scaled field first and multiplies once (:307-321)
//
// Decisions of all agents are independent given the table at step start (rows created during the step are zero
// rows and an absent row also reads as 0), so they run in parallel; the backups are the reference's sequential
// loop, executed by one thread (learn = EXACT, one episode per handle).  With learn = NONE the table is frozen
// and any number of episodes runs concurrently.  learn = BATCHED also freezes the table during the rollout but
// records, per episode, the order in which the agents' paths would have been backed up; afterwards either
// mcq_backup_ordered_kernel applies the reference's backups in episode order (exact whenever the policy did not
// depend on Q: beta = 1, i.e. the coverage pretrain and the warm-up episodes of
// run_coverage_pretrain_and_training.py:173-216,313-333) or mcq_accumulate_kernel / mcq_fold_kernel reduce the
// returns per (state, action) and fold them in with the visit count (the synchronous batched form whose deltas
// all-reduce over GPUs).
//
// Coverage pretrain (run_coverage_pretrain_and_training.py:91-166, force_first_step_and_roll): an episode may start
// with a teacher-forced first transition -- the single agent stands on src, the record (state of target T with src
// occupied, from_dir, -step/-stop penalty) is appended, the agent is moved to T with a footprint on src, no DFF
// update and no step counted -- and has its own step cap (SFF(src) + 10), after which finalize_timeouts runs.
#pragma once
#include "ffm_unified_kernel.cuh"

namespace ffm {

constexpr uint32_t MCQ_STATES_PER_BLOCK = 262144u;   // 4^9
enum { RW_STEP = 0, RW_STOP = 1, RW_COLL = 2, RW_EXIT = 3, RW_TIMEOUT = 4 };   // reward codes of a path record

constexpr unsigned long long MCQ_EMPTY = ~0ULL;
constexpr int MCQ_ERR_TABLE_FULL = 64;

struct McqParams {
    int H, W, HW, n_max, B;
    int max_steps;           // steps to run in this launch
    int step_cap;            // params["max_steps"]: finalize_timeouts when the step counter reaches it
    int learn;               // ULEARN_NONE | ULEARN_EXACT | ULEARN_BATCHED (deferred: finish order recorded, table untouched)
    int force_finalize;      // finalize_timeouts() called by the driver before the cap (main_learning.py:96-97)
    int nby;
    const uint16_t* type_grid;
    const void* sff;         // [HW] S: the SFF itself (float32 or float64 as loaded)
    double kS, kD, kQ, beta, alpha, gamma;
    double rw[5];            // reward of each code: -step_pen, -stop_pen, -coll_pen, exit_reward, -timeout_pen
    float c0, c1, thr;
    uint32_t* pos; int32_t* n_alive; int32_t* t_done; unsigned long long* ped_steps;
    float* dff; float* dff_tmp;
    unsigned long long* qkeys; float* Q;        // hash table: [cap] keys (MCQ_EMPTY = free slot), [cap][5] rows
    uint32_t qmask;                             // cap - 1 (cap is a power of two)
    unsigned int* q_count;                      // rows in the table
    int32_t* err;                               // device error flag (MCQ_ERR_TABLE_FULL)
    uint32_t* path_state; uint8_t* path_code;   // [B][path_rows][n_max]: table slot; action | reward code << 4
    int path_rows;                              // step_cap + 2 (a forced first record and the timeout record)
    int32_t* path_len;                          // [B][n_max]
    int32_t* path_shift;                        // [B] 1 when the episode began with a teacher-forced record (rows = step + shift)
    uint16_t* fin_order; int32_t* fin_count;    // [B][n_max], [B]: columns in the order their paths are backed up
    const int32_t* forced_target;               // [B] target cell T of the teacher-forced first transition, -1 = none; or null
    const int32_t* forced_dir;                  // [B] its FROM_* action
    const int32_t* ep_cap;                      // [B] per-episode cap on CA steps, then finalize_timeouts; or null
    uint16_t* path_col;                         // [B][n_max] column of the pedestrian at index i (survives launches)
    unsigned long long seed; uint32_t episode_base;
    const double* move_draws; const double* conflict_draws; int draw_steps, draw_first;
    uint32_t* traj; int32_t* traj_n; int traj_steps;
};

struct MSmemLayout { uint32_t grid, claim, pos, posB, tgt, info, orig, origB, wcnt, arr, misc, total; };

__host__ __device__ inline MSmemLayout make_mlayout(int HW, int W, int n_max) {
    MSmemLayout L;
    uint32_t o = 0;
    L.grid = o;  o = align16(o + (uint32_t)(HW + 2 * (W + 1)) * 2u);
    L.claim = o; o = align16(o + (uint32_t)HW);
    L.pos = o;   o = align16(o + (uint32_t)n_max * 4u);
    L.posB = o;  o = align16(o + (uint32_t)n_max * 4u);
    L.tgt = o;   o = align16(o + (uint32_t)n_max * 4u);
    L.info = o;  o = align16(o + (uint32_t)n_max * 4u);
    L.orig = o;  o = align16(o + (uint32_t)n_max * 2u);
    L.origB = o; o = align16(o + (uint32_t)n_max * 2u);
    L.wcnt = o;  o = align16(o + (uint32_t)(n_max / 32 + 2) * 4u);
    L.arr = o;   o = align16(o + (uint32_t)n_max * 2u);
    L.misc = o;  o = align16(o + 32u);
    L.total = o;
    return L;
}

// value of one window cell (:121-139): map code with OOB = 2, +1 when a free cell is occupied
__device__ __forceinline__ uint32_t mcq_cell_code(uint32_t g) {
    const uint32_t ty = g >> TYPE_SHIFT, occ = g & OCC_MASK;
    if (ty == TYPE_EXIT) return 3u;
    if (ty == TYPE_WALL) return occ == OCC_MASK ? 2u : 1u;          // wall (2) / map code 1
    return occ != 0u ? 1u : 0u;                                     // free cell: occupied or not
}

// 64-bit key of target (tr, tc) (:115-140 + _block_index :112-113)
__device__ __forceinline__ unsigned long long mcq_state(const uint16_t* grid, int tr, int tc, int H, int W, int nby) {
    uint32_t code = 0;
    const bool inner = tr >= 1 && tr < H - 1 && tc >= 1 && tc < W - 1;
    int k = 0;
#pragma unroll
    for (int dr = -1; dr <= 1; ++dr)
#pragma unroll
        for (int dc = -1; dc <= 1; ++dc) {
            uint32_t v = 2u;
            const int a = tr + dr, b = tc + dc;
            if (inner || (a >= 0 && a < H && b >= 0 && b < W)) v = mcq_cell_code(grid[a * W + b]);
            code |= v << (2 * k);
            ++k;
        }
    return (unsigned long long)((tr / 3) * nby + tc / 3) * MCQ_STATES_PER_BLOCK + code;
}

// ---- the Q dict as an open-addressing hash table ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mcq_hash(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return (uint32_t)k;
}
// the read path (self.Q.get(state_key), :190): slot of the key or -1; never inserts.  A key being inserted concurrently
// may be missed: its row is a zero row, and an absent row reads as 0 as well.
__device__ __forceinline__ int mcq_find(const unsigned long long* keys, uint32_t mask, unsigned long long key) {
    uint32_t s = mcq_hash(key) & mask;
    for (uint32_t probe = 0; probe <= mask; ++probe) {
        const unsigned long long cur = __ldcg(keys + s);
        if (cur == key) return (int)s;
        if (cur == MCQ_EMPTY) return -1;
        s = (s + 1u) & mask;
    }
    return -1;
}
// _ensure_qvec (:289-291): slot of the key, inserted (zero row: the table is zero-filled and rows are never deleted) if absent
__device__ __forceinline__ uint32_t mcq_find_or_insert(unsigned long long* keys, uint32_t mask, unsigned long long key,
                                                      unsigned int* count, int32_t* err) {
    uint32_t s = mcq_hash(key) & mask;
    for (uint32_t probe = 0; probe <= mask; ++probe) {
        unsigned long long cur = __ldcg(keys + s);
        if (cur == MCQ_EMPTY) {
            cur = atomicCAS(keys + s, MCQ_EMPTY, key);
            if (cur == MCQ_EMPTY) {
                if (atomicAdd(count, 1u) > (mask >> 1)) atomicOr(err, MCQ_ERR_TABLE_FULL);   // keep the load factor below 1/2
                return s;
            }
        }
        if (cur == key) return s;
        s = (s + 1u) & mask;
    }
    atomicOr(err, MCQ_ERR_TABLE_FULL);
    return 0u;
}

// reverse Monte-Carlo backup of one path (:262-267): Python-float returns, float32 table arithmetic
__device__ __forceinline__ void mcq_backup(const McqParams& P, int e, int col, int len) {
    double G = 0.0;
    const float alpha32 = (float)P.alpha;
    const size_t stride = (size_t)P.n_max, base = (size_t)e * P.path_rows * stride + col;
    float* Q = P.Q;      // one CTA owns the table in the EXACT mode: plain (L1-cached) accesses, program order per thread
    for (int t = len - 1; t >= 0; --t) {
        const uint32_t sid = P.path_state[base + (size_t)t * stride];              // table slot (the row exists: _ensure_qvec ran)
        const uint32_t pc = P.path_code[base + (size_t)t * stride];
        G = __dadd_rn(P.rw[pc >> 4], __dmul_rn(P.gamma, G));                       // G = r + gamma * G
        const size_t qi = (size_t)sid * 5 + (pc & 0xFu);
        const float q = Q[qi];
        Q[qi] = __fadd_rn(q, __fmul_rn(alpha32, __fsub_rn((float)G, q)));           // Q += alpha * (G - Q)   (float32)
    }
}

template <typename S, int THREADS>
__global__ void __launch_bounds__(THREADS)
ffm_mcq_rollout_kernel(const McqParams P) {
    constexpr int NBR = 4, A = 5;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int e = blockIdx.x;
    const int W = P.W, HW = P.HW, H = P.H, G = W + 1;
    const MSmemLayout L = make_mlayout(HW, W, P.n_max);
    constexpr uint32_t EXIT_EMPTY = TYPE_EXIT << TYPE_SHIFT;
    uint16_t* grid = reinterpret_cast<uint16_t*>(smem_raw + L.grid) + G;
    uint8_t* claim = smem_raw + L.claim;
    uint32_t* claim32 = reinterpret_cast<uint32_t*>(smem_raw + L.claim);
    uint32_t* pos = reinterpret_cast<uint32_t*>(smem_raw + L.pos);
    uint32_t* posB = reinterpret_cast<uint32_t*>(smem_raw + L.posB);
    uint32_t* tgt = reinterpret_cast<uint32_t*>(smem_raw + L.tgt);
    uint32_t* info = reinterpret_cast<uint32_t*>(smem_raw + L.info);     // bit0 moved, bit1 arrived, bit2 lost a conflict
    uint16_t* orig = reinterpret_cast<uint16_t*>(smem_raw + L.orig);
    uint16_t* origB = reinterpret_cast<uint16_t*>(smem_raw + L.origB);
    uint32_t* wcnt = reinterpret_cast<uint32_t*>(smem_raw + L.wcnt);
    uint16_t* arr = reinterpret_cast<uint16_t*>(smem_raw + L.arr);       // arrivals of this step (agent indices)
    int* misc = reinterpret_cast<int*>(smem_raw + L.misc);               // [0] number of arrivals

    const S* sff = reinterpret_cast<const S*>(P.sff);
    float* dffA = P.dff + (size_t)e * HW;
    float* dffB = P.dff_tmp + (size_t)e * HW;
    float* dff_home = dffA;
    for (int c = tid; c < HW + 2 * G; c += THREADS) grid[c - G] = P.type_grid[c];
    for (int c = tid; c < (HW + 3) / 4; c += THREADS) claim32[c] = 0u;
    int n = P.n_alive[e];
    const int t0 = P.t_done[e];
    uint32_t* gpos = P.pos + (size_t)e * P.n_max;
    const size_t pstride = (size_t)P.n_max, pbase = (size_t)e * P.path_rows * pstride;
    int32_t* plen = P.path_len + (size_t)e * P.n_max;
    uint16_t* gcol = P.path_col + (size_t)e * P.n_max;
    for (int i = tid; i < n; i += THREADS) { pos[i] = gpos[i]; orig[i] = (t0 == 0) ? (uint16_t)i : gcol[i]; }
    if (t0 == 0) for (int i = tid; i < P.n_max; i += THREADS) plen[i] = 0;
    if (t0 == 0 && tid == 0) { P.path_shift[e] = 0; P.fin_count[e] = 0; }
    __syncthreads();
    for (int i = tid; i < n; i += THREADS) grid[pos[i]] |= (uint16_t)(i + 1);
    __syncthreads();
    for (int i = tid; i < n; i += THREADS)      // duplicates OR their ids together: somebody reads back a foreign id
        if ((grid[pos[i]] & OCC_MASK) != (uint32_t)(i + 1)) atomicOr(P.err, 128);
    // teacher-forced first transition of a coverage-pretrain mini-episode (run_coverage_pretrain_and_training.py:116-147)
    if (t0 == 0 && P.forced_target != nullptr && P.forced_target[e] >= 0 && n == 1 && tid == 0) {
        const int c = (int)pos[0], T = P.forced_target[e], a = P.forced_dir[e];
        const int tr = T / W, tcc = T - tr * W;
        const uint32_t slot = mcq_find_or_insert(P.qkeys, P.qmask, mcq_state(grid, tr, tcc, H, W, P.nby), P.q_count, P.err);   // :128-130
        P.path_state[pbase + orig[0]] = slot;
        P.path_code[pbase + orig[0]] = (uint8_t)(a | ((a == 4 ? RW_STOP : RW_STEP) << 4));   // :135-147
        plen[orig[0]] = 1;
        P.path_shift[e] = 1;
        if (a != 4) {                                     // footprint on src, agent now on T; no DFF update, no step counted
            dffA[c] = __fadd_rn(dffA[c], 1.0f);
            grid[c] &= (uint16_t)TYPE_BITS;
            grid[T] |= (uint16_t)1;
            pos[0] = (uint32_t)T;
        }
    }
    __syncthreads();
    const int shift = P.path_shift[e];
    const int my_cap = P.ep_cap != nullptr ? P.ep_cap[e] : 0x7fffffff;

    const uint32_t episode = P.episode_base + (uint32_t)e;
    const double* mv_draws = P.move_draws ? P.move_draws + (size_t)e * P.draw_steps * P.n_max : nullptr;
    const double* cf_draws = P.conflict_draws ? P.conflict_draws + (size_t)e * P.draw_steps * HW * 2 : nullptr;
    const bool learn = P.learn == ULEARN_EXACT;
    const bool deferred = P.learn == ULEARN_BATCHED;
    uint16_t* fin_order = P.fin_order + (size_t)e * P.n_max;
    const double wq = __dmul_rn(1.0 - P.beta, P.kQ);       // (1 - beta) * k_Q
    const int doff[4] = {-W, W, -1, 1};                     // UP, DOWN, LEFT, RIGHT (:73)
    // finalize_timeouts() (:326-360): every remaining agent appends (state of its own cell, STOP, -timeout_penalty)
    // as record `row` of its path and is backed up, in index order
    auto finalize = [&](int row) {
        for (int i = tid; i < n; i += THREADS) {
            const int c = (int)pos[i];
            const int r = c / W, col = c - r * W;
            const size_t at = pbase + (size_t)row * pstride + orig[i];
            P.path_state[at] = mcq_find_or_insert(P.qkeys, P.qmask, mcq_state(grid, r, col, H, W, P.nby), P.q_count, P.err);
            P.path_code[at] = (uint8_t)(4 | (RW_TIMEOUT << 4));
            plen[orig[i]] = row + 1;
        }
        __syncthreads();
        if (tid == 0 && learn)
            for (int i = 0; i < n; ++i) mcq_backup(P, e, orig[i], row + 1);
        if (tid == 0 && deferred) {
            int f = P.fin_count[e];
            for (int i = 0; i < n; ++i) fin_order[f++] = orig[i];
            P.fin_count[e] = f;
        }
        __syncthreads();
    };
    unsigned long long ped_steps = 0;
    int tl = 0;
    for (; tl < P.max_steps && n > 0 && t0 + tl < my_cap; ++tl) {
        const int tstep = t0 + tl;
        const int prow = tstep + shift;                     // row of this step's path record
        const uint32_t t = (uint32_t)tstep;
        ped_steps += (unsigned long long)n;
        const int di = tstep - P.draw_first;
        const bool inj = di >= 0 && di < P.draw_steps;
        if (tid == 0) misc[0] = 0;

        // ================= M1: candidates, target-centric states, logits, draw =================
        for (int i = tid; i < n; i += THREADS) {
            const int c = (int)pos[i];
            const int r = c / W, col = c - r * W;
            int ccell[A], cact[A];
            unsigned long long csid[A];
            double e_[A];
            int nc = 0;
#pragma unroll
            for (int k = 0; k < NBR; ++k) {
                const uint32_t g = grid[c + doff[k]];
                if ((g & OCC_MASK) == 0u) {                                   // in bounds, passable, unoccupied (:175-181)
                    ccell[nc] = c + doff[k];
                    cact[nc] = (k == 0) ? 1 : (k == 1) ? 0 : (k == 2) ? 3 : 2;  // _dir_to_from (:294-305)
                    ++nc;
                }
            }
            ccell[nc] = c; cact[nc] = 4; ++nc;                                 // STOP (:183)
            double mx = -__longlong_as_double(0x7ff0000000000000LL);
#pragma unroll
            for (int j = 0; j < A; ++j)
                if (j < nc) {
                    const int tc_ = ccell[j];
                    const int tr = tc_ / W, tcc = tc_ - tr * W;
                    const unsigned long long sid = mcq_state(grid, tr, tcc, H, W, P.nby);
                    csid[j] = sid;
                    double q_val = 0.0;                                                                     // (:190-191)
                    if (wq != 0.0) {                       // beta = 1: the Q term is 0 * q (tables hold finite values)
                        const int slot = mcq_find(P.qkeys, P.qmask, sid);
                        if (slot >= 0) q_val = (double)P.Q[(size_t)slot * 5 + cact[j]];
                    }
                    const double a1 = __dmul_rn(P.beta, __dmul_rn(-P.kS, (double)sff[tc_]));
                    const double a2 = __dmul_rn(P.kD, (double)dffA[tc_]);
                    const double lg = __dadd_rn(__dadd_rn(a1, a2), __dmul_rn(wq, q_val));                   // (:193)
                    e_[j] = lg;
                    mx = fmax(mx, lg);
                }
            double tot = 0.0;
#pragma unroll
            for (int j = 0; j < A; ++j)
                if (j < nc) { e_[j] = exp(__dadd_rn(e_[j], -mx)); tot += e_[j]; }
            int chosen = nc - 1;                                                // not finite / <= 0: STOP (:201-203)
            if (isfinite(tot) && tot > 0.0) {
                const double u = (inj && mv_draws) ? mv_draws[(size_t)di * P.n_max + i]
                                                   : draw_u0(P.seed, episode, t, STREAM_MOVE, (uint32_t)i);
                const double thresh = u * tot;
                double run = 0.0;
                bool done = false;
#pragma unroll
                for (int j = 0; j < A - 1; ++j)
                    if (j < nc - 1 && !done) {
                        run += e_[j];
                        if (run > thresh) { chosen = j; done = true; }
                    }
            }
            uint32_t target = (uint32_t)c;
            unsigned long long sidc = csid[0];
            int act = 4;
#pragma unroll
            for (int j = 0; j < A; ++j)
                if (j == chosen) { target = (uint32_t)ccell[j]; sidc = csid[j]; act = cact[j]; }
            tgt[i] = target;
            info[i] = 0u;
            if (target != (uint32_t)c) atomicAdd(&claim32[target >> 2], 1u << (8 * (target & 3u)));
            // path record (:221-222); the loser / exit overwrites happen below
            const size_t at = pbase + (size_t)prow * pstride + orig[i];
            P.path_state[at] = mcq_find_or_insert(P.qkeys, P.qmask, sidc, P.q_count, P.err);   // _ensure_qvec (:221)
            P.path_code[at] = (uint8_t)(act | ((act == 4 ? RW_STOP : RW_STEP) << 4));
            (void)col;
        }
        __syncthreads();

        // ================= B: conflicts (one winner, :229-257) ==================================
        for (int i = tid; i < n; i += THREADS) {
            const int c = (int)pos[i];
            const uint32_t T = tgt[i];
            uint32_t w = 0;
            if (T == (uint32_t)c) {
                w = 0u;                                                          // STOP: nothing moves, no footprint
            } else {
                const int k = (int)claim[T];
                bool win = true;
                if (k > 1) {
                    int rk = 0;
#pragma unroll
                    for (int q = 0; q < NBR; ++q) {
                        const uint32_t o = (grid[(int)T + doff[q]] & OCC_MASK) - 1u;
                        if (o < (uint32_t)i && tgt[o] == T) ++rk;
                    }
                    const double u1 = (inj && cf_draws) ? cf_draws[((size_t)di * HW + T) * 2 + 1]
                                                        : draw2(P.seed, episode, t, STREAM_CONFLICT, T).u1;
                    win = (int)(u1 * (double)k) == rk;
                }
                if (win) {
                    w = 1u;
                    dffA[c] = __fadd_rn(dffA[c], 1.0f);                          // (:235,247)
                    if (grid[T] == EXIT_EMPTY) w |= 2u;                          // arrival (:239,250)
                } else {
                    w = 4u;                                                      // loser: last reward := -collision_penalty
                    const size_t at = pbase + (size_t)prow * pstride + orig[i];
                    P.path_code[at] = (uint8_t)((P.path_code[at] & 0xFu) | (RW_COLL << 4));
                }
            }
            info[i] = w;
        }
        __syncthreads();

        // ================= C: apply; collect arrivals ===========================================
        for (int base = 0; base < n; base += THREADS) {
            const int i = base + tid;
            bool arrived = false;
            if (i < n) {
                const int c = (int)pos[i];
                const uint32_t T = tgt[i];
                if (T != (uint32_t)c) claim[T] = 0;
                uint32_t nc_ = (uint32_t)c;
                const uint32_t w = info[i];
                if (w & 1u) {
                    grid[c] &= (uint16_t)TYPE_BITS;
                    if (!(w & 2u)) grid[T] |= (uint16_t)(i + 1);
                    nc_ = T;
                }
                posB[i] = nc_;
                plen[orig[i]] = prow + 1;
                arrived = (w & 2u) != 0u;
            }
            const uint32_t bal = __ballot_sync(0xffffffffu, arrived);
            if (bal != 0u) {
                int b0 = 0;
                if (lane == 0) b0 = atomicAdd(&misc[0], __popc(bal));
                b0 = __shfl_sync(0xffffffffu, b0, 0);
                if (arrived) arr[b0 + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)i;
            }
        }
        __syncthreads();

        // ================= T: arrivals -> exit reward + reverse MC backup (:263-267) ============
        const int n_arr = misc[0];
        if (n_arr > 0 && tid == 0) {
            // descending agent index (sorted(arrived_indices, reverse=True)); the list is tiny (<= #exit cells)
            for (int a = 0; a < n_arr; ++a)
                for (int b = a + 1; b < n_arr; ++b)
                    if (arr[b] > arr[a]) { const uint16_t x = arr[a]; arr[a] = arr[b]; arr[b] = x; }
            for (int a = 0; a < n_arr; ++a) {
                const int i = arr[a], colm = orig[i];
                const size_t at = pbase + (size_t)prow * pstride + colm;
                P.path_code[at] = (uint8_t)((P.path_code[at] & 0xFu) | (RW_EXIT << 4));   // last reward := exit_reward
                if (learn) mcq_backup(P, e, colm, prow + 1);
                if (deferred) fin_order[P.fin_count[e]++] = (uint16_t)colm;
            }
        }
        __syncthreads();

        // ================= K: stable compaction of the survivors ================================
        for (int base = 0; base < n; base += THREADS) {
            const int i = base + tid;
            const bool kept = i < n && !(info[i] & 2u);
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
                const int wq_ = w0 + lane;
                const int x = (wq_ < ngroups) ? (int)wcnt[wq_] : 0;
                int xb = (wq_ < v) ? x : 0, xt = x;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    xb += __shfl_xor_sync(0xffffffffu, xb, d);
                    xt += __shfl_xor_sync(0xffffffffu, xt, d);
                }
                before += xb;
                total += xt;
            }
            n_new = total;
            const bool kept = i < n && !(info[i] & 2u);
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

        // ================= D: _update_dff (:307-321), always Moore ===============================
        for (int c = tid; c < HW; c += THREADS) {
            const int r = c / W, col = c - r * W;
            float acc = 0.0f;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int rr = r + nbr_dr<8>(k), cc = col + nbr_dc<8>(k);
                const float v = (rr >= 0 && rr < H && cc >= 0 && cc < W) ? __fmul_rn(P.c0, dffA[rr * W + cc]) : 0.0f;
                acc = __fadd_rn(acc, v);                                          // acc += padded[...]
            }
            acc = __fmul_rn(acc, P.c1);                                           // acc *= decay*(1-diffuse)/8
            float out = __fadd_rn(__fmul_rn(P.c0, dffA[c]), acc);                 // base + acc
            if (out < P.thr) out = 0.0f;
            dffB[c] = out;
        }
        { float* tmp = dffA; dffA = dffB; dffB = tmp; }
        __syncthreads();
        n = n_new;

        // ================= timeouts at the step cap (:284-285) ==================================
        if (tstep + 1 >= P.step_cap && n > 0) {
            finalize(prow + 1);
            n = 0;                                                                      // everybody is cleared (:357-360)
        }
        if (P.traj != nullptr && tl < P.traj_steps) {
            uint32_t* row = P.traj + ((size_t)e * P.traj_steps + tl) * P.n_max;
            for (int i = tid; i < n; i += THREADS) row[i] = pos[i];
            if (tid == 0) P.traj_n[(size_t)e * P.traj_steps + tl] = n;
        }
    }

    // finalize_timeouts() called by the driver: explicitly (main_learning.py:96-97) or at the mini-episode's own cap
    // (run_coverage_pretrain_and_training.py:150-162)
    if ((P.force_finalize || t0 + tl >= my_cap) && n > 0 && t0 + tl <= P.step_cap) {
        finalize(t0 + tl + shift);
        n = 0;
    }
    for (int i = tid; i < n; i += THREADS) { gpos[i] = pos[i]; gcol[i] = orig[i]; }
    if (dffA != dff_home)
        for (int c = tid; c < HW; c += THREADS) dff_home[c] = dffA[c];
    if (tid == 0) {
        P.n_alive[e] = n;
        P.t_done[e] = t0 + tl;
        P.ped_steps[e] += ped_steps;
    }
}

// ---- after a BATCHED (deferred) rollout ----------------------------------------------------------------------------------
// The reference's backups of a whole batch, in episode order and, within an episode, in the recorded finish order: exactly
// what running the episodes one after the other on the shared dict does (:262-278, :350-355) whenever their policy did not
// read Q.  Updates to different (row, action) entries commute, so thread k of the single CTA applies the entries whose
// slot falls into its residue class, in sequence; every thread walks all records (broadcast loads) and recomputes G.
static __global__ void __launch_bounds__(1024) mcq_backup_ordered_kernel(const McqParams P) {
    const uint32_t tid = threadIdx.x;
    const float alpha32 = (float)P.alpha;
    const size_t stride = (size_t)P.n_max;
    for (int e = 0; e < P.B; ++e) {
        const int nf = P.fin_count[e];
        for (int q = 0; q < nf; ++q) {
            const int col = P.fin_order[(size_t)e * P.n_max + q];
            const int len = P.path_len[(size_t)e * P.n_max + col];
            const size_t base = (size_t)e * P.path_rows * stride + col;
            double G = 0.0;
            for (int t = len - 1; t >= 0; --t) {
                const uint32_t sid = P.path_state[base + (size_t)t * stride];
                const uint32_t pc = P.path_code[base + (size_t)t * stride];
                G = __dadd_rn(P.rw[pc >> 4], __dmul_rn(P.gamma, G));
                if ((sid & 1023u) == tid) {
                    const size_t qi = (size_t)sid * 5 + (pc & 0xFu);
                    const float qv = P.Q[qi];
                    P.Q[qi] = __fadd_rn(qv, __fmul_rn(alpha32, __fsub_rn((float)G, qv)));
                }
            }
        }
    }
}

// Synchronous batched form: returns summed per (row, action) with their visit counts (one thread per finished path) ...
static __global__ void mcq_accumulate_kernel(const McqParams P, double* dG, double* dN) {
    const size_t stride = (size_t)P.n_max;
    const long long total = (long long)P.B * P.n_max;
    for (long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(x / P.n_max), q = (int)(x - (long long)e * P.n_max);
        if (q >= P.fin_count[e]) continue;
        const int col = P.fin_order[(size_t)e * P.n_max + q];
        const int len = P.path_len[(size_t)e * P.n_max + col];
        const size_t base = (size_t)e * P.path_rows * stride + col;
        double G = 0.0;
        for (int t = len - 1; t >= 0; --t) {
            const uint32_t sid = P.path_state[base + (size_t)t * stride];
            const uint32_t pc = P.path_code[base + (size_t)t * stride];
            G = __dadd_rn(P.rw[pc >> 4], __dmul_rn(P.gamma, G));
            atomicAdd(&dG[(size_t)sid * 5 + (pc & 0xFu)], G);
            atomicAdd(&dN[(size_t)sid * 5 + (pc & 0xFu)], 1.0);
        }
    }
}
// ... and folded in: n sequential updates Q += alpha (G_i - Q) towards targets of mean m move Q by (1 - (1 - alpha)^n)(m - Q)
// when the targets are equal, which is the commutative form used here (the deltas all-reduce over GPUs by key).
static __global__ void mcq_fold_kernel(float* Q, double* dG, double* dN, size_t entries, double alpha) {
    for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < entries; x += (size_t)gridDim.x * blockDim.x) {
        const double n = dN[x];
        if (n > 0.0) {
            const double q = (double)Q[x];
            Q[x] = (float)(q + (1.0 - pow(1.0 - alpha, n)) * (dG[x] / n - q));
            dG[x] = 0.0; dN[x] = 0.0;
        }
    }
}
// exchange by key (the slot of a key differs between ranks): rows touched since the last fold -> (key, sum G[5], n[5])
static __global__ void mcq_export_deltas_kernel(const unsigned long long* keys, double* dG, double* dN, uint32_t cap, unsigned long long* out_keys,
                                         double* out_rows, unsigned int* out_count, unsigned int out_cap, int32_t* err) {
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < cap; s += gridDim.x * blockDim.x) {
        double n[5]; bool any = false;
#pragma unroll
        for (int a = 0; a < 5; ++a) { n[a] = dN[(size_t)s * 5 + a]; any |= n[a] > 0.0; }
        if (!any) continue;
        const unsigned int k = atomicAdd(out_count, 1u);
        if (k >= out_cap) atomicOr(err, 256);                  // the caller's list is too short: reported at the next host read
        if (k < out_cap) {
            out_keys[k] = keys[s];
#pragma unroll
            for (int a = 0; a < 5; ++a) { out_rows[(size_t)k * 10 + a] = dG[(size_t)s * 5 + a]; out_rows[(size_t)k * 10 + 5 + a] = n[a]; }
        }
#pragma unroll
        for (int a = 0; a < 5; ++a) { dG[(size_t)s * 5 + a] = 0.0; dN[(size_t)s * 5 + a] = 0.0; }
    }
}
// one rank's exported list added into the local delta tables (keys are unique within a list: plain adds, so that importing the
// lists in rank order gives every rank bit-identical sums)
static __global__ void mcq_import_deltas_kernel(McqParams P, const unsigned long long* in_keys, const double* in_rows, unsigned int count,
                                         const unsigned int* count_dev, double* dG, double* dN) {
    if (count_dev != nullptr) count = min(*count_dev, count);          // the exporting rank's count travels with its list
    for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) {
        const uint32_t s = mcq_find_or_insert(P.qkeys, P.qmask, in_keys[k], P.q_count, P.err);
#pragma unroll
        for (int a = 0; a < 5; ++a) { dG[(size_t)s * 5 + a] += in_rows[(size_t)k * 10 + a]; dN[(size_t)s * 5 + a] += in_rows[(size_t)k * 10 + 5 + a]; }
    }
}
// `model.Q = shared_Q` (main_learning.py:81): insert n (key, row) pairs into a cleared table
static __global__ void mcq_insert_rows_kernel(McqParams P, const unsigned long long* in_keys, const float* in_rows, long long n) {
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
        const uint32_t s = mcq_find_or_insert(P.qkeys, P.qmask, in_keys[k], P.q_count, P.err);
#pragma unroll
        for (int a = 0; a < 5; ++a) P.Q[(size_t)s * 5 + a] = in_rows[(size_t)k * 5 + a];
    }
}

}  // namespace ffm
