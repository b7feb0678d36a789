// Legacy 13-cell models: the TD(0) critic of model/ffm_ac_core.py and the actor of
// model/ffm_actor_only.py, one CTA per episode, all steps in-kernel, the dict tables as open-addressing hash tables
// in HBM (L2-resident at the sizes the reference reaches: a few thousand keys).
//
// Own translation unit and own handle (ffm_legacy_t): the legacy models share the draw streams, the owner grid
// encoding, the candidate / softmax arithmetic (move_weights) and the DFF stencil with the current models, nothing else.
#include "../../include/ffm_b200.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "ffm_core_kernel.cuh"
#include "ffm_internal.h"

namespace ffm {

constexpr unsigned long long LEG_EMPTY = ~0ull;
constexpr uint32_t LEG_PEDMARK = (TYPE_WALL << TYPE_SHIFT) | 0x3FFEu;   // map code 1: not walkable, reads as 1 in the state
constexpr uint32_t LEG_NONE = 0xFFFFFFFFu;
enum { LEG_ERR_DUP = 128, LEG_ERR_TABLE_FULL = 64 };

struct LegTable {                 // one dict: keys [cap], rows [cap][width] float64 (unseen slots hold the default row)
    unsigned long long* keys;
    double* rows;
    uint32_t mask;
    unsigned int* count;
    uint32_t* order;              // slots in insertion order (rescans walk `count` entries instead of the whole table)
};

struct LegStats { double hmin, hmax; int any, dirty; };   // extremes over every value of the H table

struct LegacyParams {
    int H, W, HW, n_max, B;
    int max_steps, learn, block_size, nby;
    const uint16_t* type_grid;    // [HW + 2*(W+1)]
    const void* score;            // AC: -k_S * sff in the SFF's dtype
    float kd, c0, c1, thr;
    double kA, gamma, alpha_v, alpha_h, exit_reward, step_penalty, collision_penalty, epsilon;
    double sff_min, sff_max;
    uint32_t* pos; int32_t* n_alive; int32_t* t_done; unsigned long long* ped_steps;
    float* dff; float* dff_tmp;
    LegTable V, Ht;
    LegStats* hstats;
    unsigned long long seed; uint32_t episode_base;
    uint32_t* traj; int32_t* traj_n; int traj_steps;
    int32_t* err;
};

__host__ __device__ __forceinline__ uint32_t leg_hash(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return (uint32_t)k;
}

// dict lookup that inserts (defaultdict read / explicit insert): slot of the key.  New slots already hold the default row.
__device__ __forceinline__ uint32_t leg_find_or_insert(const LegTable& T, unsigned long long key, int32_t* err, bool* inserted = nullptr) {
    uint32_t s = leg_hash(key) & T.mask;
    for (uint32_t probe = 0; probe <= T.mask; ++probe) {
        unsigned long long cur = __ldcg(T.keys + s);
        if (cur == LEG_EMPTY) {
            cur = atomicCAS(T.keys + s, LEG_EMPTY, key);
            if (cur == LEG_EMPTY) {
                const unsigned int at = atomicAdd(T.count, 1u);
                if (at > (T.mask >> 1)) atomicOr(err, LEG_ERR_TABLE_FULL);   // keep the load factor below 1/2
                else T.order[at] = s;
                if (inserted) *inserted = true;
                return s;
            }
        }
        if (cur == key) return s;
        s = (s + 1u) & T.mask;
    }
    atomicOr(err, LEG_ERR_TABLE_FULL);
    return 0u;
}
__device__ __forceinline__ int leg_find(const LegTable& T, unsigned long long key) {
    uint32_t s = leg_hash(key) & T.mask;
    for (uint32_t probe = 0; probe <= T.mask; ++probe) {
        const unsigned long long cur = __ldcg(T.keys + s);
        if (cur == key) return (int)s;
        if (cur == LEG_EMPTY) return -1;
        s = (s + 1u) & T.mask;
    }
    return -1;
}

// value of a cell in the reference's state_map / occupancy (0 free, 1 pedestrian, 2 wall, 3 exit)
__device__ __forceinline__ uint32_t leg_cell_code(uint32_t g) {
    const uint32_t type = g >> TYPE_SHIFT, occ = g & OCC_MASK;
    if (type == TYPE_WALL) return occ == 0x3FFEu ? 1u : 2u;
    if (type == TYPE_EXIT) return 3u;
    return occ != 0u ? 1u : 0u;
}

// _encode_state (ffm_ac_core.py:62-109 with OUTSIDE = 2, ffm_actor_only.py:102-148 with OUTSIDE = 0)
template <uint32_t OUTSIDE>
__device__ __forceinline__ unsigned long long leg_key13(const uint16_t* grid, int r, int col, int H, int W, int bs, int nby) {
    uint32_t code = 0;
    int j = 0;
#pragma unroll
    for (int a = -1; a <= 1; ++a)
#pragma unroll
        for (int b = -1; b <= 1; ++b) {
            const int p = r + a, q = col + b;
            const uint32_t v = (p >= 0 && p < H && q >= 0 && q < W) ? leg_cell_code(grid[p * W + q]) : OUTSIDE;
            code |= v << (2 * j);
            ++j;
        }
#pragma unroll
    for (int d = 0; d < 4; ++d) {                     // U2, D2, L2, R2 (:89)
        const int p = r + (d == 0 ? -2 : (d == 1 ? 2 : 0)), q = col + (d == 2 ? -2 : (d == 3 ? 2 : 0));
        const uint32_t v = (p >= 0 && p < H && q >= 0 && q < W) ? leg_cell_code(grid[p * W + q]) : OUTSIDE;
        code |= v << (2 * j);
        ++j;
    }
    return ((unsigned long long)((r / bs) * nby + col / bs) << 26) | code;
}

struct LSmemLayout { uint32_t grid, claim, first, pos, posB, tgt, info, st, nst, td, req, wcnt, misc, total; };

__host__ __device__ inline LSmemLayout make_llayout(int HW, int W, int n_max, bool actor) {
    LSmemLayout L;
    uint32_t o = 0;
    L.grid = o;  o = align16(o + (uint32_t)(HW + 2 * (W + 1)) * 2u);
    L.claim = o; o = align16(o + (uint32_t)HW + 4u);
    L.first = o; if (actor) o = align16(o + (uint32_t)HW * 4u);          // order of the first request per cell
    L.pos = o;   o = align16(o + (uint32_t)n_max * 4u);
    L.posB = o;  o = align16(o + (uint32_t)n_max * 4u);
    L.tgt = o;   o = align16(o + (uint32_t)n_max * 4u);
    L.info = o;  o = align16(o + (uint32_t)n_max * 4u);
    L.st = o;    o = align16(o + (uint32_t)n_max * 4u);
    L.nst = o;   o = align16(o + (uint32_t)n_max * 4u);
    L.td = o;    if (actor) o = align16(o + (uint32_t)n_max * 8u);
    L.req = o;   if (actor) o = align16(o + (uint32_t)n_max * 8u * 4u);  // up to 8 requested cells per pedestrian
    L.wcnt = o;  o = align16(o + (uint32_t)(n_max / 32 + 2) * 4u);
    L.misc = o;  o = align16(o + 64u);
    L.total = o;
    return L;
}

constexpr uint32_t LI_REQ = 1u << 0;            // filed a request
constexpr uint32_t LI_EXIT = 1u << 1;           // will_exit
constexpr uint32_t LI_MOVED = 1u << 2;          // request granted
constexpr uint32_t LI_COLL_SHIFT = 8;           // collision count (k - 1)

// stable compaction of the survivors (keep_mask, ffm_ac_core.py:241-244): pos <- posB without the cells on exits
template <int THREADS>
__device__ __forceinline__ int leg_compact(uint16_t* grid, const uint32_t* posB, uint32_t* pos, uint32_t* wcnt, int n, int tid) {
    constexpr uint32_t EXIT_EMPTY = TYPE_EXIT << TYPE_SHIFT;
    const int lane = tid & 31;
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
            grid[c] = (uint16_t)((grid[c] & TYPE_BITS) | (uint32_t)(ni + 1));
        }
    }
    return n_new;
}

// reward of one agent-step (ffm_ac_core.py:268-281): Python-float arithmetic in the reference's order
__device__ __forceinline__ double leg_reward(const LegacyParams& P, bool will_exit, bool has_coll, int coll) {
    double rew = P.step_penalty;
    if (will_exit) rew = __dadd_rn(rew, P.exit_reward);
    if (has_coll) rew = __dadd_rn(rew, __dmul_rn((double)coll, P.collision_penalty));
    return rew;
}

// ---- model/ffm_ac_core.py ------------------------------------------------------------------------------------------------
// Phases per step (block-wide, __syncthreads between):
//   A  state key -> table slot; candidates (free or exit, unoccupied neighbours in neighbour order, then "stay": :141-164);
//      no candidate -> no request, no draw (:163); an exit among them -> forced request, will_exit (:172-178); else
//      softmax over the candidates and a keyed draw (:187-202)
//   B  conflicts: a contested cell always has one winner, the floor(u*k)-th claimant (:216-219); all k claimants record
//      k-1 collisions (:217-229); granted requests (incl. "stay") leave a DFF footprint (:211-213, :221-223)
//   C  apply the moves: the owner grid becomes state_map_next (:233-236)
//   T  next-state keys -> slots (parallel), then ONE thread applies the TD(0) updates in agent order on the shared
//      table (:264-296) -- the reference's sequential semantics
//   K  stable exit removal (:241-244);  D  DFF decay + diffusion (:298-318)
template <typename S, int NBR, int THREADS>
__global__ void __launch_bounds__(THREADS)
ffm_legacy_ac_kernel(const LegacyParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int e = blockIdx.x;
    const int W = P.W, HW = P.HW, H = P.H;
    const int G = W + 1;
    const LSmemLayout L = make_llayout(HW, W, P.n_max, false);
    constexpr uint32_t EXIT_EMPTY = TYPE_EXIT << TYPE_SHIFT;

    uint16_t* grid = reinterpret_cast<uint16_t*>(smem_raw + L.grid) + G;
    uint8_t* claim = smem_raw + L.claim;
    uint32_t* claim32 = reinterpret_cast<uint32_t*>(smem_raw + L.claim);
    uint32_t* pos = reinterpret_cast<uint32_t*>(smem_raw + L.pos);
    uint32_t* posB = reinterpret_cast<uint32_t*>(smem_raw + L.posB);
    uint32_t* tgt = reinterpret_cast<uint32_t*>(smem_raw + L.tgt);
    uint32_t* info = reinterpret_cast<uint32_t*>(smem_raw + L.info);
    uint32_t* st = reinterpret_cast<uint32_t*>(smem_raw + L.st);
    uint32_t* nst = reinterpret_cast<uint32_t*>(smem_raw + L.nst);
    uint32_t* wcnt = reinterpret_cast<uint32_t*>(smem_raw + L.wcnt);

    const S* score = reinterpret_cast<const S*>(P.score);
    float* dff_home = P.dff + (size_t)e * HW;
    float* dffA = dff_home;
    float* dffB = P.dff_tmp + (size_t)e * HW;
    for (int c = tid; c < HW + 2 * G; c += THREADS) grid[c - G] = P.type_grid[c];
    for (int c = tid; c < HW / 4 + 1; c += THREADS) claim32[c] = 0u;
    int n = P.n_alive[e];
    const int t0 = P.t_done[e];
    uint32_t* gpos = P.pos + (size_t)e * P.n_max;
    for (int i = tid; i < n; i += THREADS) pos[i] = gpos[i];
    __syncthreads();
    for (int i = tid; i < n; i += THREADS) grid[pos[i]] |= (uint16_t)(i + 1);
    __syncthreads();
    for (int i = tid; i < n; i += THREADS)
        if ((grid[pos[i]] & OCC_MASK) != (uint32_t)(i + 1)) atomicOr(P.err, LEG_ERR_DUP);

    const uint32_t episode = P.episode_base + (uint32_t)e;
    const bool learn = P.learn != 0;
    unsigned long long ped_steps = 0;
    int tl = 0;
    const StencilGeom sgeom = make_stencil_geom(H, W, tid, THREADS);
    for (; tl < P.max_steps && n > 0; ++tl) {
        const uint32_t t = (uint32_t)(t0 + tl);
        ped_steps += (unsigned long long)n;

        // ================= A ====================================================================
        for (int i = tid; i < n; i += THREADS) {
            const int c = (int)pos[i];
            const int r = c / W, col = c - r * W;
            if (learn) st[i] = leg_find_or_insert(P.V, leg_key13<2u>(grid, r, col, H, W, P.block_size, P.nby), P.err);   // :129-130
            uint32_t mm = 0, ex = 0;
#pragma unroll
            for (int k = 0; k < NBR; ++k) {
                const uint32_t g = grid[c + nbr_off<NBR>(k, W)];
                if ((g & OCC_MASK) == 0u) {                           // map 0 / 3 and nobody there (:141-160)
                    mm |= 1u << k;
                    if ((g >> TYPE_SHIFT) == TYPE_EXIT) ex |= 1u << k;
                }
            }
            uint32_t w = 0, target = LEG_NONE;
            if (mm != 0u) {                                           // :163
                if (ex != 0u) {                                       // first exit among the candidates (:172-178)
                    target = (uint32_t)(c + nbr_off_rt<NBR>(__ffs(ex) - 1, W));
                    w = LI_REQ | LI_EXIT;
                } else {
                    const int ncand = __popc(mm) + 1;
                    int cell[NBR + 1];
                    S p[NBR + 1];
                    const double tot = move_weights<S, NBR, true>(mm, ncand, c, W, score, dffA, P.kd, cell, p);   // :180-193
                    if (isfinite(tot) && tot != 0.0) {                // :195
                        const double thresh = draw_u0(P.seed, episode, t, STREAM_MOVE, (uint32_t)i) * tot;       // :197
                        double run = 0.0;
                        target = (uint32_t)c;                         // "stay" is the last candidate
                        bool done = false;
#pragma unroll
                        for (int j = 0; j < NBR; ++j)
                            if (j < ncand - 1 && !done) {
                                run += (double)p[j];
                                if (run > thresh) { target = (uint32_t)cell[j]; done = true; }
                            }
                        w = LI_REQ;
                    }
                }
                if (target != LEG_NONE && target != (uint32_t)c) atomicAdd(&claim32[target >> 2], 1u << (8 * (target & 3u)));
            }
            tgt[i] = target;
            info[i] = w;
        }
        __syncthreads();

        // ================= B ====================================================================
        for (int i = tid; i < n; i += THREADS) {
            uint32_t w = info[i];
            if (!(w & LI_REQ)) continue;
            const int c = (int)pos[i];
            const uint32_t T = tgt[i];
            bool moved = true;
            if (T != (uint32_t)c) {
                const int k = (int)claim[T];
                if (k > 1) {
                    int rk = 0;                                        // position among the claimants, in agent order
#pragma unroll
                    for (int q = 0; q < NBR; ++q) {
                        const uint32_t o = (grid[(int)T + nbr_off<NBR>(q, W)] & OCC_MASK) - 1u;
                        if (o < (uint32_t)i && tgt[o] == T) ++rk;
                    }
                    const double u1 = draw2(P.seed, episode, t, STREAM_CONFLICT, T).u1;
                    moved = (int)(u1 * (double)k) == rk;               // random.choice(agents) (:219)
                    w |= (uint32_t)(k - 1) << LI_COLL_SHIFT;
                }
            }
            if (moved) {
                w |= LI_MOVED;
                dffA[c] = __fadd_rn(dffA[c], 1.0f);                    // :211-213, :221-223
            }
            info[i] = w;
        }
        __syncthreads();

        // ================= C ====================================================================
        for (int i = tid; i < n; i += THREADS) {
            const int c = (int)pos[i];
            const uint32_t T = tgt[i];
            const uint32_t w = info[i];
            uint32_t nc = (uint32_t)c;
            if (w & LI_REQ) {
                if (T != (uint32_t)c) claim[T] = 0;
                if ((w & LI_MOVED) && T != (uint32_t)c) {
                    grid[c] &= (uint16_t)TYPE_BITS;
                    if (grid[T] != EXIT_EMPTY) grid[T] |= (uint16_t)(i + 1);
                    nc = T;
                }
            }
            posB[i] = nc;
        }
        __syncthreads();

        // ================= T ====================================================================
        if (learn) {
            for (int i = tid; i < n; i += THREADS) {
                uint32_t ns = LEG_NONE;
                if (!(info[i] & LI_EXIT)) {                            // :283-290
                    const int c = (int)posB[i];
                    const int r = c / W, col = c - r * W;
                    ns = leg_find_or_insert(P.V, leg_key13<2u>(grid, r, col, H, W, P.block_size, P.nby), P.err);
                }
                nst[i] = ns;
            }
            __syncthreads();
            if (tid == 0) {
                double* V = P.V.rows;
                for (int i = 0; i < n; ++i) {                          // :264-296, agent order
                    const uint32_t w = info[i];
                    const double rew = leg_reward(P, (w & LI_EXIT) != 0u, (w & LI_REQ) != 0u, (int)(w >> LI_COLL_SHIFT));
                    const double v_next = (w & LI_EXIT) ? 0.0 : V[nst[i]];
                    const double v_cur = V[st[i]];
                    const double td = __dadd_rn(__dadd_rn(rew, __dmul_rn(P.gamma, v_next)), -v_cur);    // :293
                    V[st[i]] = __dadd_rn(v_cur, __dmul_rn(P.alpha_v, td));                              // :296
                }
            }
            __syncthreads();
        }

        // ================= K, D =================================================================
        const int n_new = leg_compact<THREADS>(grid, posB, pos, wcnt, n, tid);
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

// ---- model/ffm_actor_only.py ---------------------------------------------------------------------------------------------
// The reference's decision block is nested inside the loop that collects the exit flags (:214-355): an agent decides once
// per neighbour slot i, seeing only the exit flags of slots <= i, and files one request per slot.  Requests to one cell form
// a list WITH duplicates, ordered by (agent, slot); cells are resolved in the order of their first request (dict order,
// :360); a later grant overwrites an earlier one of the same agent (:362, :372), every grant leaves a DFF footprint
// (:363-365, :373-375), and the collision count of an agent is the one of the last resolved cell it asked for (:366, :376-380).
// Phases per step:
//   A0 state key -> V slot; validity of the slots; does the H row exist?  (the first agent whose lookup inserts a row decides
//      from where on the table extremes include the zeros of a fresh row, :252-269)
//   A1 H row (inserted as zeros if absent), scores -- flat over the valid slots as soon as ONE slot is invalid (the -inf of
//      :300 trips the isinf test of :304) --, one epsilon-greedy / keyed draw per slot i < first exit slot, one forced exit
//      request per slot i >= first exit slot; claim counters and first-request order per cell
//   B  per agent and distinct requested cell: position of its requests in the cell's list, winner = floor(u*k)-th entry
//      (:370); final cell = granted cell with the latest first request; footprints
//   C  apply the moves (owner grid becomes occupancy_next, :388-391), clear the per-cell scratch
//   T  next-state V slots and H slots of all agents (parallel), then ONE thread: TD(0) in agent order (:434-472), then
//      H[s][a] += alpha_h * delta for the LAST decision of each agent (:495-536), keeping the table extremes current
//   K  exit removal;  D  DFF
constexpr uint32_t AI_ACT_MASK = 0xFu;
constexpr uint32_t AI_VALID_SHIFT = 4;          // 9 bits
constexpr uint32_t AI_EXIT = 1u << 13;
constexpr uint32_t AI_HASCOLL = 1u << 14;       // always set: every agent files requests
constexpr uint32_t AI_COLL_SHIFT = 16;          // 8 bits

template <int NBR, int THREADS>
__global__ void __launch_bounds__(THREADS)
ffm_legacy_actor_kernel(const LegacyParams P) {
    constexpr int A = NBR + 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int e = blockIdx.x;
    const int W = P.W, HW = P.HW, H = P.H;
    const int G = W + 1;
    const LSmemLayout L = make_llayout(HW, W, P.n_max, true);
    constexpr uint32_t EXIT_EMPTY = TYPE_EXIT << TYPE_SHIFT;
    const double DINF = __longlong_as_double(0x7ff0000000000000LL);

    uint16_t* grid = reinterpret_cast<uint16_t*>(smem_raw + L.grid) + G;
    uint8_t* claim = smem_raw + L.claim;
    uint32_t* claim32 = reinterpret_cast<uint32_t*>(smem_raw + L.claim);
    uint32_t* first = reinterpret_cast<uint32_t*>(smem_raw + L.first);
    uint32_t* pos = reinterpret_cast<uint32_t*>(smem_raw + L.pos);
    uint32_t* posB = reinterpret_cast<uint32_t*>(smem_raw + L.posB);
    uint32_t* hs = reinterpret_cast<uint32_t*>(smem_raw + L.tgt);        // H slot of the agent's state (LEG_NONE: no row)
    uint32_t* info = reinterpret_cast<uint32_t*>(smem_raw + L.info);
    uint32_t* st = reinterpret_cast<uint32_t*>(smem_raw + L.st);
    uint32_t* nst = reinterpret_cast<uint32_t*>(smem_raw + L.nst);
    double* tdv = reinterpret_cast<double*>(smem_raw + L.td);
    uint32_t* req = reinterpret_cast<uint32_t*>(smem_raw + L.req);       // [n][8]
    uint32_t* wcnt = reinterpret_cast<uint32_t*>(smem_raw + L.wcnt);
    int* misc = reinterpret_cast<int*>(smem_raw + L.misc);               // [0] first agent inserting an H row, [1] dirty, [2] rows inserted
    double* dmisc = reinterpret_cast<double*>(smem_raw + L.misc + 16);   // [0] hmin, [1] hmax at step start
    __shared__ double red_lo[32], red_hi[32];

    float* dff_home = P.dff + (size_t)e * HW;
    float* dffA = dff_home;
    float* dffB = P.dff_tmp + (size_t)e * HW;
    for (int c = tid; c < HW + 2 * G; c += THREADS) grid[c - G] = P.type_grid[c];
    for (int c = tid; c < HW / 4 + 1; c += THREADS) claim32[c] = 0u;
    for (int c = tid; c < HW; c += THREADS) first[c] = 0xFFFFFFFFu;
    int n = P.n_alive[e];
    const int t0 = P.t_done[e];
    uint32_t* gpos = P.pos + (size_t)e * P.n_max;
    for (int i = tid; i < n; i += THREADS) pos[i] = gpos[i];
    __syncthreads();
    for (int i = tid; i < n; i += THREADS) grid[pos[i]] |= (uint16_t)(i + 1);
    __syncthreads();
    for (int i = tid; i < n; i += THREADS)
        if ((grid[pos[i]] & OCC_MASK) != (uint32_t)(i + 1)) atomicOr(P.err, LEG_ERR_DUP);

    const uint32_t episode = P.episode_base + (uint32_t)e;
    const bool learn = P.learn != 0;
    unsigned long long ped_steps = 0;
    int tl = 0;
    const StencilGeom sgeom = make_stencil_geom(H, W, tid, THREADS);
    for (; tl < P.max_steps && n > 0; ++tl) {
        const uint32_t t = (uint32_t)(t0 + tl);
        ped_steps += (unsigned long long)n;

        // ---- extremes of the H table at step start (rescan when an extreme value moved inwards) --
        if (tid == 0) { misc[1] = learn ? P.hstats->dirty : 0; misc[0] = 0x7fffffff; misc[2] = 0; }
        __syncthreads();
        if (misc[1]) {
            double lo = DINF, hi = -DINF;
            const unsigned int cnt = *P.Ht.count;
            for (unsigned int x = tid; x < cnt; x += THREADS) {
                const double* row = P.Ht.rows + (size_t)P.Ht.order[x] * A;
#pragma unroll
                for (int a = 0; a < A; ++a) { lo = fmin(lo, row[a]); hi = fmax(hi, row[a]); }
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, d));
                hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, d));
            }
            if (lane == 0) { red_lo[warp] = lo; red_hi[warp] = hi; }
            __syncthreads();
            if (tid == 0) {
                for (int w = 1; w < THREADS / 32; ++w) { lo = fmin(lo, red_lo[w]); hi = fmax(hi, red_hi[w]); }
                P.hstats->hmin = lo; P.hstats->hmax = hi; P.hstats->any = cnt > 0; P.hstats->dirty = 0;
            }
            __syncthreads();
        }
        if (tid == 0) {
            dmisc[0] = P.hstats->any ? P.hstats->hmin : DINF;
            dmisc[1] = P.hstats->any ? P.hstats->hmax : -DINF;
        }

        // ================= A0 ===================================================================
        for (int i = tid; i < n; i += THREADS) {
            const int c = (int)pos[i];
            const int r = c / W, col = c - r * W;
            const unsigned long long key = leg_key13<0u>(grid, r, col, H, W, P.block_size, P.nby);      // :174-175
            if (learn) st[i] = leg_find_or_insert(P.V, key, P.err);
            uint32_t valid = 1u << NBR, ex = 0;                       // "stay" is always valid (:212)
#pragma unroll
            for (int k = 0; k < NBR; ++k) {
                const uint32_t g = grid[c + nbr_off<NBR>(k, W)];
                if ((g & OCC_MASK) == 0u) valid |= 1u << k;           // in bounds, map 0 / 3, nobody else there (:180-206)
                if ((g >> TYPE_SHIFT) == TYPE_EXIT) ex |= 1u << k;    // :217-221
            }
            const int first_exit = ex ? __ffs(ex) - 1 : NBR;
            const int slot = leg_find(P.Ht, key);
            if (slot < 0 && first_exit != 0 && learn) atomicMin(&misc[0], i);   // this agent's first lookup inserts the row
            hs[i] = slot < 0 ? LEG_NONE : (uint32_t)slot;
            info[i] = (valid << AI_VALID_SHIFT) | (ex ? AI_EXIT : 0u) | (uint32_t)first_exit;           // act field: first exit slot for now
            nst[i] = (uint32_t)(key & 0xFFFFFFFFu);                  // the key travels to A1 in two halves
            posB[i] = (uint32_t)(key >> 32);
        }
        __syncthreads();

        // ================= A1 ===================================================================
        const int first_new = misc[0];
        for (int i = tid; i < n; i += THREADS) {
            const int c = (int)pos[i];
            uint32_t w = info[i];
            const uint32_t valid = (w >> AI_VALID_SHIFT) & ((1u << A) - 1u);
            const int first_exit = (int)(w & AI_ACT_MASK);
            int act = first_exit;
            if (first_exit > 0) {
                // H row: inserted as zeros by the first lookup (:252-257)
                uint32_t slot = hs[i];
                const bool have = slot != LEG_NONE;
                if (!have && learn) {
                    bool ins = false;
                    slot = leg_find_or_insert(P.Ht, ((unsigned long long)posB[i] << 32) | nst[i], P.err, &ins);
                    hs[i] = slot;
                    if (ins) misc[2] = 1;
                }
                double h[A];
#pragma unroll
                for (int k = 0; k < A; ++k) h[k] = have ? P.Ht.rows[(size_t)slot * A + k] : 0.0;
                double hmin = dmisc[0], hmax = dmisc[1];
                if (first_new <= i || (!learn && !have)) { hmin = fmin(hmin, 0.0); hmax = fmax(hmax, 0.0); }
                if (hmax - hmin > 1e-6) {                               // :288-293
                    const double den = hmax - hmin, rng = P.sff_max - P.sff_min;
#pragma unroll
                    for (int k = 0; k < A; ++k)
                        h[k] = __dadd_rn(__dmul_rn(__ddiv_rn(__dadd_rn(hmax, -h[k]), den), rng), P.sff_min);
                }
                double e_[A], mx = -DINF;
                bool flat = valid != ((1u << A) - 1u);                  // an invalid slot scores -inf (:300) -> isinf (:304) -> flat
#pragma unroll
                for (int k = 0; k < A; ++k) {
                    const int cc = (k == NBR) ? c : c + nbr_off<NBR>(k, W);
                    e_[k] = __dadd_rn(__dmul_rn(-P.kA, h[k]), (double)__fmul_rn(P.kd, dffA[cc]));        // :295-298
                    if (!isfinite(e_[k])) flat = true;
                    mx = fmax(mx, e_[k]);
                }
                double tot = 0.0;
                if (!flat) {
#pragma unroll
                    for (int k = 0; k < A; ++k) { e_[k] = exp(__dadd_rn(e_[k], -mx)); tot += e_[k]; }   // :321
                }
                if (flat || !(isfinite(tot) && tot > 0.0)) {            // flat scores: exp(1 - 1) = 1 per valid slot (:306-312, :321-333)
                    tot = 0.0;
#pragma unroll
                    for (int k = 0; k < A; ++k) { e_[k] = ((valid >> k) & 1u) ? 1.0 : 0.0; tot += e_[k]; }
                }
                const int nv = __popc(valid);
                for (int s = 0; s < NBR && s < first_exit; ++s) {       // one decision per slot before the first exit slot
                    const uint32_t ent = (uint32_t)i * 8u + (uint32_t)s;
                    int slot_c = -1;
                    if (P.epsilon > 0.0) {                              // :329-341
                        const Draw2 d = draw2(P.seed, episode, t, STREAM_EPS, ent);
                        if (d.u0 < P.epsilon) slot_c = (int)__fns(valid, 0, (int)(d.u1 * (double)nv) + 1);
                    }
                    if (slot_c < 0) {
                        const double thresh = draw_u0(P.seed, episode, t, STREAM_MOVE, ent) * tot;      // :343
                        double run = 0.0;
                        slot_c = NBR;
                        bool done = false;
#pragma unroll
                        for (int k = 0; k < NBR; ++k)
                            if (!done) {
                                run += e_[k];
                                if (run > thresh) { slot_c = k; done = true; }
                            }
                    }
                    const uint32_t target = (slot_c == NBR) ? (uint32_t)c : (uint32_t)(c + nbr_off_rt<NBR>(slot_c, W));
                    req[i * 8 + s] = target;
                    atomicAdd(&claim32[target >> 2], 1u << (8 * (target & 3u)));
                    atomicMin(&first[target], ent);
                    act = slot_c;
                }
            }
            if (first_exit < NBR) {                                     // forced exit request for every slot from the first exit on (:223-241)
                const uint32_t target = (uint32_t)(c + nbr_off_rt<NBR>(first_exit, W));
                for (int s = first_exit; s < NBR; ++s) {
                    req[i * 8 + s] = target;
                    atomicAdd(&claim32[target >> 2], 1u << (8 * (target & 3u)));
                    atomicMin(&first[target], (uint32_t)i * 8u + (uint32_t)s);
                }
                act = first_exit;
            }
            info[i] = (w & ~AI_ACT_MASK) | (uint32_t)act;
        }
        __syncthreads();

        // ================= B ====================================================================
        for (int i = tid; i < n; i += THREADS) {
            const int c = (int)pos[i];
            uint32_t best_ord = 0, last_ord = 0, nxt = (uint32_t)c, coll = 0;
            bool any_won = false, any_req = false;
            int wins = 0;
            for (int s = 0; s < NBR; ++s) {
                const uint32_t T = req[i * 8 + s];
                bool dup = false;
                int m = 0;
                for (int j = 0; j < NBR; ++j) {
                    if (req[i * 8 + j] == T) { if (j < s) dup = true; ++m; }
                }
                if (dup) continue;                                     // each distinct cell once
                const int k = (int)claim[T];
                bool won = true;
                uint32_t cc = 0;
                if (k > 1) {
                    int r0 = 0;                                        // entries of the cell's list filed by agents before this one
                    if (T != (uint32_t)c) {
#pragma unroll
                        for (int q = 0; q < NBR; ++q) {
                            const uint32_t o = (grid[(int)T + nbr_off<NBR>(q, W)] & OCC_MASK) - 1u;
                            if (o < (uint32_t)i)
                                for (int j = 0; j < NBR; ++j) r0 += (req[o * 8 + j] == T) ? 1 : 0;
                        }
                    }
                    const int p = (int)(draw2(P.seed, episode, t, STREAM_CONFLICT, T).u1 * (double)k);   // random.choice(agents) (:370)
                    won = p >= r0 && p < r0 + m;
                    cc = (uint32_t)(k - 1);
                }
                const uint32_t ord = first[T];
                if (!any_req || ord > last_ord) { last_ord = ord; coll = cc; any_req = true; }
                if (won) {
                    ++wins;
                    if (!any_won || ord > best_ord) { best_ord = ord; nxt = T; any_won = true; }
                }
            }
            float d = dffA[c];
            for (int q = 0; q < wins; ++q) d = __fadd_rn(d, 1.0f);     // one footprint per grant (:363-365, :373-375)
            dffA[c] = d;
            posB[i] = nxt;
            info[i] = (info[i] & 0xFFFFu) | AI_HASCOLL | (coll << AI_COLL_SHIFT);
        }
        __syncthreads();

        // ================= C ====================================================================
        for (int i = tid; i < n; i += THREADS) {
            const int c = (int)pos[i];
            for (int s = 0; s < NBR; ++s) { const uint32_t T = req[i * 8 + s]; claim[T] = 0; first[T] = 0xFFFFFFFFu; }
            const uint32_t T = posB[i];
            if (T != (uint32_t)c) {
                grid[c] &= (uint16_t)TYPE_BITS;
                if (grid[T] != EXIT_EMPTY) grid[T] |= (uint16_t)(i + 1);
            }
        }
        __syncthreads();

        // ================= T ====================================================================
        if (learn) {
            for (int i = tid; i < n; i += THREADS) {
                uint32_t ns = LEG_NONE;
                if (!(info[i] & AI_EXIT)) {                            // :452-458
                    const int c = (int)posB[i];
                    const int r = c / W, col = c - r * W;
                    ns = leg_find_or_insert(P.V, leg_key13<0u>(grid, r, col, H, W, P.block_size, P.nby), P.err);
                }
                nst[i] = ns;
                if (hs[i] == LEG_NONE) {                               // agents that never looked their row up get it in _update_actor (:520-525)
                    // the state key of time t cannot be rebuilt from the moved grid: it was parked in st's table slot -> V keys
                    bool ins = false;
                    hs[i] = leg_find_or_insert(P.Ht, P.V.keys[st[i]], P.err, &ins);
                    if (ins) misc[2] = 1;
                }
            }
            __syncthreads();
            if (tid == 0) {
                double* V = P.V.rows;
                for (int i = 0; i < n; ++i) {                          // :434-472
                    const uint32_t w = info[i];
                    const double rew = leg_reward(P, (w & AI_EXIT) != 0u, true, (int)((w >> AI_COLL_SHIFT) & 0xFFu));
                    const double v_next = (w & AI_EXIT) ? 0.0 : V[nst[i]];
                    const double v_cur = V[st[i]];
                    const double td = __dadd_rn(__dadd_rn(rew, __dmul_rn(P.gamma, v_next)), -v_cur);
                    V[st[i]] = __dadd_rn(v_cur, __dmul_rn(P.alpha_v, td));
                    tdv[i] = td;
                }
                LegStats* hsx = P.hstats;
                double hmin = hsx->hmin, hmax = hsx->hmax;
                int any = hsx->any, dirty = hsx->dirty;
                if (misc[2]) {                                         // zero rows joined the table in this step
                    if (!any) { any = 1; hmin = 0.0; hmax = 0.0; } else { hmin = fmin(hmin, 0.0); hmax = fmax(hmax, 0.0); }
                }
                double* Hm = P.Ht.rows;
                for (int i = 0; i < n; ++i) {                          // :495-536
                    const uint32_t w = info[i];
                    const uint32_t a = w & AI_ACT_MASK;
                    if ((w >> (AI_VALID_SHIFT + a)) & 1u) {
                        double* cell = Hm + (size_t)hs[i] * A + a;
                        const double old = *cell;
                        const double nw = __dadd_rn(old, __dmul_rn(P.alpha_h, tdv[i]));                 // :534
                        *cell = nw;
                        if (nw < hmin) hmin = nw; else if (old == hmin && nw > old) dirty = 1;
                        if (nw > hmax) hmax = nw; else if (old == hmax && nw < old) dirty = 1;
                    }
                }
                hsx->hmin = hmin; hsx->hmax = hmax; hsx->any = any; hsx->dirty = dirty;
            }
            __syncthreads();
        }

        // ================= K, D =================================================================
        const int n_new = leg_compact<THREADS>(grid, posB, pos, wcnt, n, tid);
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

template <int NBR>
__global__ void __launch_bounds__(256) leg_dff_update_kernel(const float* in, float* out, int H, int W, float c0, float c1, float thr) {
    const size_t off = (size_t)blockIdx.x * H * W;
    dff_decay_diffuse<NBR>(in + off, out + off, H, W, c0, c1, thr, threadIdx.x, make_stencil_geom(H, W, threadIdx.x, 256));
}

}  // namespace ffm

// ---------------------------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------------------------
namespace {

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    const int rc = ffm::set_error(code, fmt, ap);
    va_end(ap);
    return rc;
}

#define CU(call)                                                                               \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return fail(FFM_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

constexpr int MAX_SMEM_OPTIN = 232448;

struct HostTable {
    ffm::LegTable d;          // device pointers
    uint32_t cap;
    int width;
    double default_value;
};

}  // namespace

struct ffm_legacy_s {
    ffm_legacy_config_t cfg;
    int HW, A, nby, threads, smem_bytes;
    bool have_fields, have_positions;
    int64_t launches;
    const void* kernel;
    uint8_t* h_map;
    uint16_t* d_type_grid;
    void* d_score;
    uint32_t* d_pos; int32_t* d_n; int32_t* d_t; unsigned long long* d_ped_steps;
    float* d_dff; float* d_dff_tmp;
    int32_t* d_err;
    ffm::LegStats* d_hstats;
    HostTable V, Ht;
};

namespace {

template <typename S>
const void* pick_ac(int nbr, int threads) {
    if (nbr == 4) return threads == 128 ? (const void*)ffm::ffm_legacy_ac_kernel<S, 4, 128> : (const void*)ffm::ffm_legacy_ac_kernel<S, 4, 256>;
    return threads == 128 ? (const void*)ffm::ffm_legacy_ac_kernel<S, 8, 128> : (const void*)ffm::ffm_legacy_ac_kernel<S, 8, 256>;
}

const void* pick_actor(int nbr, int threads) {
    if (nbr == 4) return threads == 128 ? (const void*)ffm::ffm_legacy_actor_kernel<4, 128> : (const void*)ffm::ffm_legacy_actor_kernel<4, 256>;
    return threads == 128 ? (const void*)ffm::ffm_legacy_actor_kernel<8, 128> : (const void*)ffm::ffm_legacy_actor_kernel<8, 256>;
}

int table_alloc(HostTable& T, uint32_t cap, int width) {
    T.cap = cap; T.width = width; T.default_value = 0.0;
    T.d.mask = cap - 1u;
    CU(cudaMalloc((void**)&T.d.keys, (size_t)cap * 8));
    CU(cudaMalloc((void**)&T.d.rows, (size_t)cap * width * 8));
    CU(cudaMalloc((void**)&T.d.count, 4));
    CU(cudaMalloc((void**)&T.d.order, (size_t)(cap / 2 + 2) * 4));
    CU(cudaMemset(T.d.keys, 0xFF, (size_t)cap * 8));
    CU(cudaMemset(T.d.rows, 0, (size_t)cap * width * 8));
    CU(cudaMemset(T.d.count, 0, 4));
    return FFM_OK;
}
void table_free(HostTable& T) {
    cudaFree(T.d.keys); cudaFree(T.d.rows); cudaFree(T.d.count); cudaFree(T.d.order);
    memset(&T, 0, sizeof(T));
}

int check_flag(ffm_legacy_t h) {
    int32_t flag = 0;
    CU(cudaMemcpy(&flag, h->d_err, 4, cudaMemcpyDeviceToHost));
    if (flag == 0) return FFM_OK;
    CU(cudaMemset(h->d_err, 0, 4));
    if (flag & ffm::LEG_ERR_DUP) return fail(FFM_E_INVALID, "two pedestrians were placed on the same cell");
    if (flag & ffm::LEG_ERR_TABLE_FULL) return fail(FFM_E_UNSUPPORTED, "state table more than half full: raise ffm_legacy_config_t.table_log2_capacity");
    return fail(FFM_E_INVALID, "device validation flag %d", flag);
}

HostTable* which_table(ffm_legacy_t h, int which) {
    if (which == FFM_LEGACY_TABLE_V) return &h->V;
    if (which == FFM_LEGACY_TABLE_H && h->cfg.model == FFM_LEGACY_ACTOR_ONLY) return &h->Ht;
    return nullptr;
}

}  // namespace

extern "C" {

int ffm_legacy_create(const ffm_legacy_config_t* cfg, ffm_legacy_t* out) {
    if (!cfg || !out) return fail(FFM_E_INVALID, "null argument");
    *out = nullptr;
    if (cfg->abi_version != FFM_ABI_VERSION) return fail(FFM_E_INVALID, "abi_version %d != %d", cfg->abi_version, FFM_ABI_VERSION);
    if (cfg->height < 3 || cfg->width < 3) return fail(FFM_E_INVALID, "map must be at least 3x3");
    if ((long long)cfg->height * cfg->width > 65536) return fail(FFM_E_UNSUPPORTED, "the legacy models cover maps of up to 65536 cells");
    if (cfg->neighborhood != FFM_NEUMANN && cfg->neighborhood != FFM_MOORE) return fail(FFM_E_INVALID, "neighborhood must be 4 or 8");
    if (cfg->sff_dtype != FFM_F32 && cfg->sff_dtype != FFM_F64) return fail(FFM_E_INVALID, "sff_dtype must be FFM_F32 or FFM_F64");
    if (cfg->model != FFM_LEGACY_AC && cfg->model != FFM_LEGACY_ACTOR_ONLY) return fail(FFM_E_INVALID, "unknown legacy model %d", cfg->model);
    if (cfg->model == FFM_LEGACY_ACTOR_ONLY && cfg->sff_dtype != FFM_F32)
        return fail(FFM_E_INVALID, "the actor-only model keeps the SFF as float32 (inf -> 0 and astype(float32), ffm_actor_only.py:45-48)");
    if (cfg->learn != FFM_LEARN_NONE && cfg->learn != FFM_LEARN_EXACT) return fail(FFM_E_INVALID, "the legacy models learn sequentially (FFM_LEARN_EXACT) or not at all");
    if (cfg->n_episodes < 1) return fail(FFM_E_INVALID, "n_episodes must be >= 1");
    if (cfg->learn == FFM_LEARN_EXACT && cfg->n_episodes != 1) return fail(FFM_E_INVALID, "FFM_LEARN_EXACT reproduces the reference's sequential table updates and needs n_episodes == 1");
    if (cfg->n_max < 1 || cfg->n_max > ffm::MAX_PEDS) return fail(FFM_E_UNSUPPORTED, "n_max must be in [1, %d]", ffm::MAX_PEDS);
    if (cfg->block_size < 1) return fail(FFM_E_INVALID, "block_size must be >= 1");
    if (cfg->table_log2_capacity != 0 && (cfg->table_log2_capacity < 10 || cfg->table_log2_capacity > 28)) return fail(FFM_E_INVALID, "table_log2_capacity must be 0 (default) or in [10, 28]");
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(FFM_E_INVALID, "device %d not present (%d visible)", cfg->device, ndev);
    CU(cudaSetDevice(cfg->device));
    int cc_major = 0;
    CU(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, cfg->device));
    if (cc_major != 10) return fail(FFM_E_UNSUPPORTED, "libffm_b200 is built for sm_100a only (device is sm_%d*)", cc_major);

    ffm_legacy_s* h = new (std::nothrow) ffm_legacy_s();
    if (!h) return fail(FFM_E_INVALID, "out of host memory");
    memset(h, 0, sizeof(*h));
    h->cfg = *cfg;
    h->HW = cfg->height * cfg->width;
    h->A = cfg->neighborhood + 1;
    h->nby = (cfg->width + cfg->block_size - 1) / cfg->block_size;
    const int HW = h->HW, W = cfg->width, B = cfg->n_episodes, N = cfg->n_max;
    const int ssz = cfg->sff_dtype == FFM_F64 ? 8 : 4;
    h->threads = N <= 128 ? 128 : 256;
    h->smem_bytes = (int)ffm::make_llayout(HW, W, N, cfg->model == FFM_LEGACY_ACTOR_ONLY).total;
    if (h->smem_bytes > MAX_SMEM_OPTIN) { delete h; return fail(FFM_E_UNSUPPORTED, "episode state (%d B) does not fit the shared memory of one SM", h->smem_bytes); }
    if (cfg->model == FFM_LEGACY_ACTOR_ONLY) h->kernel = pick_actor(cfg->neighborhood, h->threads);
    else h->kernel = cfg->sff_dtype == FFM_F64 ? pick_ac<double>(cfg->neighborhood, h->threads) : pick_ac<float>(cfg->neighborhood, h->threads);
    cudaError_t ce = cudaFuncSetAttribute(h->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_bytes);
    if (ce != cudaSuccess) { delete h; return fail(FFM_E_CUDA, "cudaFuncSetAttribute(smem=%d): %s", h->smem_bytes, cudaGetErrorString(ce)); }
#define LALLOC(ptr, bytes)                                                                    \
    do {                                                                                      \
        cudaError_t e_ = cudaMalloc((void**)&(ptr), (bytes));                                 \
        if (e_ != cudaSuccess) { ffm_legacy_destroy(h); return fail(FFM_E_CUDA, "cudaMalloc(%zu) failed: %s", (size_t)(bytes), cudaGetErrorString(e_)); } \
    } while (0)
    h->h_map = (uint8_t*)malloc((size_t)HW);
    LALLOC(h->d_type_grid, (size_t)(HW + 2 * (W + 1)) * 2);
    LALLOC(h->d_score, (size_t)HW * ssz);
    LALLOC(h->d_pos, (size_t)B * N * 4);
    LALLOC(h->d_n, (size_t)B * 4);
    LALLOC(h->d_t, (size_t)B * 4);
    LALLOC(h->d_ped_steps, (size_t)B * 8);
    LALLOC(h->d_dff, (size_t)B * HW * 4 * 2);
    h->d_dff_tmp = h->d_dff + (size_t)B * HW;
    LALLOC(h->d_err, 4);
    LALLOC(h->d_hstats, sizeof(ffm::LegStats));
#undef LALLOC
    cudaMemset(h->d_dff, 0, (size_t)B * HW * 4 * 2);
    cudaMemset(h->d_err, 0, 4);
    cudaMemset(h->d_hstats, 0, sizeof(ffm::LegStats));
    cudaMemset(h->d_n, 0, (size_t)B * 4);
    cudaMemset(h->d_t, 0, (size_t)B * 4);
    cudaMemset(h->d_ped_steps, 0, (size_t)B * 8);
    const uint32_t cap = 1u << (cfg->table_log2_capacity ? cfg->table_log2_capacity : 20);
    int rc = table_alloc(h->V, cap, 1);
    if (!rc && cfg->model == FFM_LEGACY_ACTOR_ONLY) rc = table_alloc(h->Ht, cap, h->A);
    if (rc) { ffm_legacy_destroy(h); return rc; }
    *out = h;
    return FFM_OK;
}

int ffm_legacy_destroy(ffm_legacy_t h) {
    if (!h) return FFM_OK;
    cudaSetDevice(h->cfg.device);
    free(h->h_map);
    cudaFree(h->d_type_grid); cudaFree(h->d_score); cudaFree(h->d_pos); cudaFree(h->d_n); cudaFree(h->d_t);
    cudaFree(h->d_ped_steps); cudaFree(h->d_dff); cudaFree(h->d_err); cudaFree(h->d_hstats);
    table_free(h->V);
    table_free(h->Ht);
    delete h;
    return FFM_OK;
}

int ffm_legacy_set_fields(ffm_legacy_t h, const uint8_t* map, const void* sff) {
    if (!h || !map || !sff) return fail(FFM_E_INVALID, "null argument");
    CU(cudaSetDevice(h->cfg.device));
    const int H = h->cfg.height, W = h->cfg.width, HW = h->HW, G = W + 1;
    std::vector<uint16_t> tg((size_t)HW + 2 * G, (uint16_t)ffm::WALL_CELL);
    for (int c = 0; c < HW; ++c) {
        const uint8_t m = map[c];
        const int r = c / W, col = c - r * W;
        if (m > 3) return fail(FFM_E_INVALID, "map_array holds codes outside {0,1,2,3}");
        if ((r == 0 || r == H - 1 || col == 0 || col == W - 1) && m == FFM_CELL_FREE)
            return fail(FFM_E_INVALID, "map_array has a free cell on its border (the step relies on border walls)");
        uint16_t cell = (uint16_t)ffm::WALL_CELL;
        if (m == FFM_CELL_FREE) cell = 0;
        if (m == FFM_CELL_EXIT) cell = (uint16_t)(ffm::TYPE_EXIT << ffm::TYPE_SHIFT);
        if (m == FFM_CELL_PED) cell = (uint16_t)ffm::LEG_PEDMARK;
        tg[(size_t)c + G] = cell;
    }
    memcpy(h->h_map, map, (size_t)HW);
    CU(cudaMemcpy(h->d_type_grid, tg.data(), tg.size() * 2, cudaMemcpyHostToDevice));
    // score = (-k_S) * sff in the SFF's own dtype (the first product of ffm_ac_core.py:187-190; elementwise, so hoisting is exact)
    if (h->cfg.sff_dtype == FFM_F64) {
        std::vector<double> sc((size_t)HW);
        const double nk = -h->cfg.k_S;
        for (int c = 0; c < HW; ++c) sc[c] = nk * ((const double*)sff)[c];
        CU(cudaMemcpy(h->d_score, sc.data(), (size_t)HW * 8, cudaMemcpyHostToDevice));
    } else {
        std::vector<float> sc((size_t)HW);
        const float nk = (float)(-h->cfg.k_S);
        for (int c = 0; c < HW; ++c) { volatile float v = nk * ((const float*)sff)[c]; sc[c] = v; }
        CU(cudaMemcpy(h->d_score, sc.data(), (size_t)HW * 4, cudaMemcpyHostToDevice));
    }
    h->have_fields = true;
    return FFM_OK;
}

int ffm_legacy_set_positions(ffm_legacy_t h, const int32_t* pos_rc, const int32_t* n) {
    if (!h || !pos_rc || !n) return fail(FFM_E_INVALID, "null argument");
    if (!h->have_fields) return fail(FFM_E_STATE, "ffm_legacy_set_fields first");
    CU(cudaSetDevice(h->cfg.device));
    const int B = h->cfg.n_episodes, N = h->cfg.n_max, H = h->cfg.height, W = h->cfg.width;
    std::vector<uint32_t> pos((size_t)B * N, 0u);
    for (int e = 0; e < B; ++e) {
        if (n[e] < 0 || n[e] > N) return fail(FFM_E_INVALID, "pedestrian count outside [0, n_max]");
        for (int i = 0; i < n[e]; ++i) {
            const int r = pos_rc[((size_t)e * N + i) * 2], c = pos_rc[((size_t)e * N + i) * 2 + 1];
            if (r < 0 || r >= H || c < 0 || c >= W) return fail(FFM_E_INVALID, "pedestrian position outside the map");
            if (h->h_map[r * W + c] != FFM_CELL_FREE) return fail(FFM_E_INVALID, "pedestrian placed on a cell that is not free (map != 0)");
            pos[(size_t)e * N + i] = (uint32_t)(r * W + c);
        }
    }
    CU(cudaMemcpy(h->d_pos, pos.data(), pos.size() * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->d_n, n, (size_t)B * 4, cudaMemcpyHostToDevice));
    CU(cudaMemset(h->d_t, 0, (size_t)B * 4));
    CU(cudaMemset(h->d_ped_steps, 0, (size_t)B * 8));
    h->have_positions = true;
    return FFM_OK;
}

int ffm_legacy_get_positions(ffm_legacy_t h, int32_t* pos_rc, int32_t* n) {
    if (!h || !pos_rc || !n) return fail(FFM_E_INVALID, "null argument");
    CU(cudaSetDevice(h->cfg.device));
    const int B = h->cfg.n_episodes, N = h->cfg.n_max, W = h->cfg.width;
    std::vector<uint32_t> pos((size_t)B * N);
    CU(cudaMemcpy(pos.data(), h->d_pos, pos.size() * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(n, h->d_n, (size_t)B * 4, cudaMemcpyDeviceToHost));
    for (int e = 0; e < B; ++e)
        for (int i = 0; i < N; ++i) {
            const bool live = i < n[e];
            pos_rc[((size_t)e * N + i) * 2] = live ? (int32_t)(pos[(size_t)e * N + i] / W) : -1;
            pos_rc[((size_t)e * N + i) * 2 + 1] = live ? (int32_t)(pos[(size_t)e * N + i] % W) : -1;
        }
    return FFM_OK;
}

int ffm_legacy_set_dff(ffm_legacy_t h, const float* dff) {
    if (!h || !dff) return fail(FFM_E_INVALID, "null argument");
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaMemcpy(h->d_dff, dff, (size_t)h->cfg.n_episodes * h->HW * 4, cudaMemcpyHostToDevice));
    return FFM_OK;
}

int ffm_legacy_get_dff(ffm_legacy_t h, float* dff) {
    if (!h || !dff) return fail(FFM_E_INVALID, "null argument");
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaMemcpy(dff, h->d_dff, (size_t)h->cfg.n_episodes * h->HW * 4, cudaMemcpyDeviceToHost));
    return FFM_OK;
}

int ffm_legacy_zero_dff(ffm_legacy_t h) {
    if (!h) return fail(FFM_E_INVALID, "null argument");
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaMemsetAsync(h->d_dff, 0, (size_t)h->cfg.n_episodes * h->HW * 4, nullptr));
    return FFM_OK;
}

int ffm_legacy_update_dff(ffm_legacy_t h) {
    if (!h) return fail(FFM_E_INVALID, "null argument");
    CU(cudaSetDevice(h->cfg.device));
    const size_t bytes = (size_t)h->cfg.n_episodes * h->HW * 4;
    if (h->cfg.neighborhood == FFM_MOORE)
        ffm::leg_dff_update_kernel<8><<<h->cfg.n_episodes, 256>>>(h->d_dff, h->d_dff_tmp, h->cfg.height, h->cfg.width, h->cfg.dff_c0, h->cfg.dff_c1, h->cfg.dff_threshold);
    else
        ffm::leg_dff_update_kernel<4><<<h->cfg.n_episodes, 256>>>(h->d_dff, h->d_dff_tmp, h->cfg.height, h->cfg.width, h->cfg.dff_c0, h->cfg.dff_c1, h->cfg.dff_threshold);
    CU(cudaGetLastError());
    CU(cudaMemcpy(h->d_dff, h->d_dff_tmp, bytes, cudaMemcpyDeviceToDevice));
    h->launches++;
    return FFM_OK;
}

int ffm_legacy_rollout(ffm_legacy_t h, int32_t max_steps, uint32_t* traj, int32_t* traj_n, int32_t traj_steps) {
    if (!h) return fail(FFM_E_INVALID, "null argument");
    if (!h->have_fields || !h->have_positions) return fail(FFM_E_STATE, "fields and positions must be set before ffm_legacy_rollout");
    if (max_steps < 0) return fail(FFM_E_INVALID, "max_steps < 0");
    if ((traj != nullptr) != (traj_n != nullptr) || (traj && traj_steps < 1)) return fail(FFM_E_INVALID, "traj and traj_n come together, with traj_steps >= 1");
    CU(cudaSetDevice(h->cfg.device));
    const int B = h->cfg.n_episodes, N = h->cfg.n_max;
    ffm::LegacyParams P;
    memset(&P, 0, sizeof(P));
    P.H = h->cfg.height; P.W = h->cfg.width; P.HW = h->HW; P.n_max = N; P.B = B;
    P.max_steps = max_steps; P.learn = h->cfg.learn; P.block_size = h->cfg.block_size; P.nby = h->nby;
    P.type_grid = h->d_type_grid; P.score = h->d_score;
    P.kd = (float)h->cfg.k_D; P.c0 = h->cfg.dff_c0; P.c1 = h->cfg.dff_c1; P.thr = h->cfg.dff_threshold;
    P.kA = h->cfg.k_A; P.gamma = h->cfg.gamma; P.alpha_v = h->cfg.alpha_v; P.alpha_h = h->cfg.alpha_h;
    P.exit_reward = h->cfg.exit_reward; P.step_penalty = h->cfg.step_penalty; P.collision_penalty = h->cfg.collision_penalty;
    P.epsilon = h->cfg.epsilon; P.sff_min = h->cfg.sff_min; P.sff_max = h->cfg.sff_max;
    P.pos = h->d_pos; P.n_alive = h->d_n; P.t_done = h->d_t; P.ped_steps = h->d_ped_steps;
    P.dff = h->d_dff; P.dff_tmp = h->d_dff_tmp;
    P.V = h->V.d; P.Ht = h->Ht.d; P.hstats = h->d_hstats;
    P.seed = h->cfg.seed; P.episode_base = h->cfg.episode_base;
    P.err = h->d_err;
    uint32_t* d_traj = nullptr; int32_t* d_traj_n = nullptr;
    if (traj) {
        CU(cudaMalloc((void**)&d_traj, (size_t)B * traj_steps * N * 4));
        CU(cudaMalloc((void**)&d_traj_n, (size_t)B * traj_steps * 4));
        CU(cudaMemset(d_traj_n, 0, (size_t)B * traj_steps * 4));
        P.traj = d_traj; P.traj_n = d_traj_n; P.traj_steps = traj_steps;
    }
    void* args[] = {&P};
    cudaError_t ce = cudaLaunchKernel(h->kernel, dim3((unsigned)B), dim3((unsigned)h->threads), args, (size_t)h->smem_bytes, nullptr);
    if (ce == cudaSuccess) ce = cudaDeviceSynchronize();
    int rc = FFM_OK;
    if (ce != cudaSuccess) rc = fail(FFM_E_CUDA, "legacy rollout kernel: %s", cudaGetErrorString(ce));
    h->launches++;
    if (!rc && traj) {
        if (cudaMemcpy(traj, d_traj, (size_t)B * traj_steps * N * 4, cudaMemcpyDeviceToHost) != cudaSuccess ||
            cudaMemcpy(traj_n, d_traj_n, (size_t)B * traj_steps * 4, cudaMemcpyDeviceToHost) != cudaSuccess)
            rc = fail(FFM_E_CUDA, "copying the trajectory record failed");
    }
    cudaFree(d_traj); cudaFree(d_traj_n);
    if (rc) return rc;
    return check_flag(h);
}

int ffm_legacy_get_counters(ffm_legacy_t h, int32_t* steps_done, uint64_t* ped_steps) {
    if (!h) return fail(FFM_E_INVALID, "null argument");
    CU(cudaSetDevice(h->cfg.device));
    if (steps_done) CU(cudaMemcpy(steps_done, h->d_t, (size_t)h->cfg.n_episodes * 4, cudaMemcpyDeviceToHost));
    if (ped_steps) CU(cudaMemcpy(ped_steps, h->d_ped_steps, (size_t)h->cfg.n_episodes * 8, cudaMemcpyDeviceToHost));
    return FFM_OK;
}

int ffm_legacy_table_size(ffm_legacy_t h, int32_t which, int64_t* n) {
    if (!h || !n) return fail(FFM_E_INVALID, "null argument");
    HostTable* T = which_table(h, which);
    if (!T) return fail(FFM_E_STATE, "this legacy model has no such table");
    CU(cudaSetDevice(h->cfg.device));
    unsigned int c = 0;
    CU(cudaMemcpy(&c, T->d.count, 4, cudaMemcpyDeviceToHost));
    *n = (int64_t)c;
    return FFM_OK;
}

int ffm_legacy_table_get(ffm_legacy_t h, int32_t which, uint64_t* keys, double* rows, int64_t capacity, int64_t* n) {
    if (!h || !keys || !rows || !n) return fail(FFM_E_INVALID, "null argument");
    HostTable* T = which_table(h, which);
    if (!T) return fail(FFM_E_STATE, "this legacy model has no such table");
    CU(cudaSetDevice(h->cfg.device));
    unsigned int c = 0;
    CU(cudaMemcpy(&c, T->d.count, 4, cudaMemcpyDeviceToHost));
    *n = (int64_t)c;
    if ((int64_t)c > capacity) return fail(FFM_E_INVALID, "table holds %u keys, buffer %lld", c, (long long)capacity);
    if (c == 0) return FFM_OK;
    std::vector<uint32_t> order(c);       // insertion order == the dict's iteration order up to ties inside one step phase
    CU(cudaMemcpy(order.data(), T->d.order, (size_t)c * 4, cudaMemcpyDeviceToHost));
    std::vector<unsigned long long> hk(T->cap);
    std::vector<double> hr((size_t)T->cap * T->width);
    CU(cudaMemcpy(hk.data(), T->d.keys, (size_t)T->cap * 8, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(hr.data(), T->d.rows, (size_t)T->cap * T->width * 8, cudaMemcpyDeviceToHost));
    for (unsigned int i = 0; i < c; ++i) {
        const uint32_t s = order[i];
        keys[i] = hk[s];
        memcpy(rows + (size_t)i * T->width, hr.data() + (size_t)s * T->width, (size_t)T->width * 8);
    }
    return FFM_OK;
}

int ffm_legacy_table_set(ffm_legacy_t h, int32_t which, const uint64_t* keys, const double* rows, int64_t n, double default_value) {
    if (!h || n < 0 || (n > 0 && (!keys || !rows))) return fail(FFM_E_INVALID, "bad argument");
    HostTable* T = which_table(h, which);
    if (!T) return fail(FFM_E_STATE, "this legacy model has no such table");
    if (n > (int64_t)(T->cap / 2)) return fail(FFM_E_UNSUPPORTED, "%lld keys exceed half the table capacity %u: raise table_log2_capacity", (long long)n, T->cap);
    CU(cudaSetDevice(h->cfg.device));
    std::vector<unsigned long long> hk(T->cap, ffm::LEG_EMPTY);
    std::vector<double> hr((size_t)T->cap * T->width, default_value);
    std::vector<uint32_t> order((size_t)n + 1);
    for (int64_t i = 0; i < n; ++i) {
        if (keys[i] == ffm::LEG_EMPTY) return fail(FFM_E_INVALID, "reserved key");
        uint32_t s = ffm::leg_hash(keys[i]) & T->d.mask;
        while (hk[s] != ffm::LEG_EMPTY) {
            if (hk[s] == keys[i]) return fail(FFM_E_INVALID, "duplicate key in the table");
            s = (s + 1u) & T->d.mask;
        }
        hk[s] = keys[i];
        memcpy(hr.data() + (size_t)s * T->width, rows + (size_t)i * T->width, (size_t)T->width * 8);
        order[i] = s;
    }
    const unsigned int c = (unsigned int)n;
    CU(cudaMemcpy(T->d.keys, hk.data(), (size_t)T->cap * 8, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(T->d.rows, hr.data(), (size_t)T->cap * T->width * 8, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(T->d.order, order.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(T->d.count, &c, 4, cudaMemcpyHostToDevice));
    T->default_value = default_value;
    if (which == FFM_LEGACY_TABLE_H) {     // extremes over every value of the table (ffm_actor_only.py:263-276)
        ffm::LegStats stt;
        stt.hmin = INFINITY; stt.hmax = -INFINITY; stt.any = n > 0; stt.dirty = 0;
        for (int64_t i = 0; i < n * T->width; ++i) { stt.hmin = fmin(stt.hmin, rows[i]); stt.hmax = fmax(stt.hmax, rows[i]); }
        CU(cudaMemcpy(h->d_hstats, &stt, sizeof(stt), cudaMemcpyHostToDevice));
    }
    return FFM_OK;
}

int ffm_legacy_set_epsilon(ffm_legacy_t h, double epsilon) {
    if (!h) return fail(FFM_E_INVALID, "null argument");
    h->cfg.epsilon = epsilon;
    return FFM_OK;
}

int ffm_legacy_set_episode_base(ffm_legacy_t h, uint32_t episode_base) {
    if (!h) return fail(FFM_E_INVALID, "null argument");
    h->cfg.episode_base = episode_base;
    return FFM_OK;
}

int64_t ffm_legacy_launch_count(ffm_legacy_t h) { return h ? h->launches : 0; }

}  // extern "C"
