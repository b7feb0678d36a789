/*
 * ffm_b200 -- C ABI of the B200-native Floor-Field-Model hot path.
 *
 * The reference (SoraKurihara/FFM) has no FFI layer: its drivers import the Python classes of
 * model/ffm_*.py directly (main.py:6, run_unified_critic_training.py:16).  This header is the
 * boundary a binding for those classes talks to (ffm_b200/model/*.py binds it with ctypes; see
 * INTEGRATION.md).  Every entry point cites the reference interface it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types cross the ABI
 *   - every function returns FFM_OK (0) or a negative FFM_E_* code and never throws;
 *     ffm_last_error() gives the message for the calling thread
 *   - `space` says where a caller buffer lives: FFM_HOST (pageable or pinned) or FFM_DEVICE
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); calls that take a
 *     stream are asynchronous with respect to the host and ordered on that stream
 *   - one host thread per handle; handles are independent
 *   - cells are (row, col) int32 pairs at the ABI, row-major linear ids (row * width + col) inside
 */
#ifndef FFM_B200_H
#define FFM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FFM_ABI_VERSION 8

enum {
    FFM_OK = 0,
    FFM_E_INVALID = -1,     /* bad argument / configuration (reference: ValueError) */
    FFM_E_CUDA = -2,        /* CUDA runtime failure, message in ffm_last_error() */
    FFM_E_UNSUPPORTED = -3, /* configuration outside what the kernels cover */
    FFM_E_STATE = -4        /* call sequence error (fields / positions not set) */
};

enum { FFM_HOST = 0, FFM_DEVICE = 1 };
enum { FFM_NEUMANN = 4, FFM_MOORE = 8 };       /* get_neighbors(): ffm_core.py:28-34 */
enum { FFM_F32 = 0, FFM_F64 = 1 };             /* dtype of the SFF file (ffm_core.py:17 keeps it) */

/* map cell codes, Create_Map.py:9-19 / ffm_unified.py:283-286 */
enum { FFM_CELL_FREE = 0, FFM_CELL_PED = 1, FFM_CELL_WALL = 2, FFM_CELL_EXIT = 3 };

/* which reference class a handle reproduces */
enum {
    FFM_MODEL_CORE = 0,           /* model/ffm_core.py FloorFieldModel */
    FFM_MODEL_UNIFIED_CRITIC = 1, /* model/ffm_unified.py learning_mode="critic_only" */
    FFM_MODEL_UNIFIED_ACTOR = 2,  /*                       learning_mode="actor_only" */
    FFM_MODEL_UNIFIED_BOTH = 3,   /*                       learning_mode="both" */
    FFM_MODEL_TRAINED = 4,        /* model/ffm_trained_core.py FloorFieldModel (frozen H table) */
    FFM_MODEL_MCQ = 5             /* model/ffm_learning_core.py FloorFieldModel (target-centric Monte-Carlo Q-learning) */
};
/* how the unified model's tables are updated */
enum {
    FFM_LEARN_NONE = 0,    /* tables frozen (evaluation rollouts; any number of episodes) */
    FFM_LEARN_EXACT = 1,   /* the reference's sequential per-agent updates; n_episodes must be 1 */
    FFM_LEARN_BATCHED = 2  /* synchronous batched TD: deltas accumulated against frozen tables, applied
                              by ffm_tables_apply_deltas (after the caller's cross-GPU all-reduce) */
};

typedef struct ffm_sim_s *ffm_sim_t;

/* Static configuration of a batch of B independent episodes on one map.
 * Replaces FloorFieldModel.__init__(map_array, sff_path, N, params)  -- ffm_core.py:7-21. */
typedef struct ffm_config {
    int32_t abi_version;   /* FFM_ABI_VERSION */
    int32_t device;        /* CUDA device ordinal */
    int32_t height, width; /* map shape */
    int32_t neighborhood;  /* FFM_NEUMANN | FFM_MOORE          params["neighborhood"] */
    int32_t sff_dtype;     /* FFM_F32 | FFM_F64                dtype move scores are computed in */
    int32_t n_episodes;    /* B */
    int32_t n_max;         /* capacity: pedestrians per episode (<= 16380) */
    int32_t track_dff;     /* 0: k_D == 0 and the caller never reads .dff -> DFF skipped entirely */
    int32_t reserved0;
    double k_S;            /* params["k_S"]   score = -k_S*sff + k_D*dff   ffm_core.py:77 */
    double k_D;            /* params["k_D"] */
    float dff_c0;          /* float32((1-decay)*(1-diffuse))               ffm_core.py:109 */
    float dff_c1;          /* float32(decay*(1-diffuse)/len(neighbors))    ffm_core.py:113 */
    float dff_threshold;   /* float32(1e-4)                                ffm_core.py:116-117 */
    float reserved1;
    uint64_t seed;         /* Philox key */
    uint32_t episode_base; /* global id of episode 0 of this handle (multi-GPU sharding) */
    uint32_t reserved2;
    /* ---- unified / trained models only (ffm_unified.py:36-53, ffm_trained_core.py:29-36) ---- */
    int32_t model;         /* FFM_MODEL_* */
    int32_t learn;         /* FFM_LEARN_* */
    int32_t block_size;    /* params["block_size"] */
    int32_t reserved3;
    double k_A;            /* params["k_A"] */
    double gamma, alpha_v, alpha_h;
    double exit_reward, step_penalty, collision_penalty;
    double epsilon;        /* params["epsilon"] / set_epsilon() */
    double sff_min, sff_max; /* float(np.min/max(self.sff)) of the inf->0 float32 SFF (ffm_unified.py:425-426) */
    /* ---- FFM_MODEL_MCQ only (ffm_learning_core.py:45-59, :76-77): k_A carries k_Q, alpha_v carries alpha,
     *      step_penalty / collision_penalty / exit_reward are the (positive) costs / reward of the params dict ---- */
    double stop_penalty, timeout_penalty;
    int32_t step_cap;      /* params["max_steps"]: finalize_timeouts() when the step counter reaches it */
    int32_t q_log2_capacity; /* log2 of the slots of the Q hash table, 10..30; 0 = default (21: 2 M slots, 59 MB) */
} ffm_config_t;

/* Optional recorded uniforms that override the keyed Philox streams (parity tests: "both sides
 * consume the same recorded draws").  Layouts are per episode; NULL members fall back to Philox.
 *   move     [B][steps][n_max]          u for np.random.choice of agent idx   ffm_core.py:84
 *   conflict [B][steps][height*width][2] (coin u, winner u) of a target cell  ffm_core.py:95-96
 * `first_step` is the CA step number row 0 corresponds to. */
typedef struct ffm_draws {
    const double *move;
    const double *conflict;
    int32_t steps;
    int32_t first_step;
    int32_t space; /* FFM_DEVICE only */
    int32_t reserved;
} ffm_draws_t;

/* Optional per-step outputs of a rollout (device buffers supplied by the caller).
 *   traj_cells [B][traj_steps][n_max] uint32 linear cell of every pedestrian still inside AFTER
 *              step t, alive-rank order -- what run() appends per step (ffm_core.py:125)
 *   traj_n     [B][traj_steps] int32 number of valid entries of each row */
typedef struct ffm_rollout_out {
    uint32_t *traj_cells;
    int32_t *traj_n;
    int32_t traj_steps;
    int32_t reserved;
    /* Rollout buffer of the unified models (SoA, one column per pedestrian = its index at the start of the
     * launch, rows = steps of this launch, [B][traj_steps][n_max]); any may be NULL:
     *   rec_state  uint32  state id of the agent-step           states[idx]      ffm_unified.py:293-294
     *   rec_action uint8   chosen slot (neighbours.., stay)     actions[idx]     :343,510
     *   rec_reward float32 step_penalty + exit + collisions     reward           :636-648
     *   rec_len    int32 [B][n_max] steps the pedestrian was inside during this launch (its path length)
     * -- the per-agent (state, action, reward) path lists of ffm_learning_core.py:79-81,221-222 in array form,
     * input of ffm_rollout_returns. */
    uint32_t *rec_state;
    uint8_t *rec_action;
    float *rec_reward;
    int32_t *rec_len;
    /* Compact trajectory record -- what run() collects (`buffer.append(copy(positions))`, ffm_core.py:125, main.py:44-52)
     * at 4 bytes per pedestrian-step instead of a dense [T][n_max] slab:
     *   ctraj      int16 [B][ctraj_cap][2]  (row, col) pairs, alive-rank order; the rows of consecutive steps of this launch
     *              back to back, each padded with (-1, -1) to a multiple of 4 entries (16-byte vector stores)
     *   ctraj_off  int32 [B][traj_steps + 1] entry offset of each step's row within the episode's stream (CSR); element
     *              [steps run] = end.  A step whose row would not fit is not recorded and its offset is -1 from there on.
     *   traj_n     (above) the number of valid entries of each row
     * Base model only. */
    int16_t *ctraj;
    int32_t *ctraj_off;
    int64_t ctraj_cap;
} ffm_rollout_out_t;

int ffm_abi_version(void);
const char *ffm_last_error(void);

/* ffm_core.py:7-21 (constructor) / object lifetime */
int ffm_create(const ffm_config_t *cfg, ffm_sim_t *out);
int ffm_destroy(ffm_sim_t sim);

/* map_array (uint8 [H][W], ffm_core.py:16) and SFF ([H][W] float32 or float64 as cfg.sff_dtype,
 * ffm_core.py:17).  Shared by all episodes of the handle. */
int ffm_set_fields(ffm_sim_t sim, const uint8_t *map, const void *sff, int space, void *stream);

/* `.positions = ...` (run_trained_ffm.py:235) / reset() (ffm_unified.py:800-812): (row, col) int32
 * pairs [B][n_max][2] and counts [B]; zeroes the DFF and the step counters. */
int ffm_set_positions(ffm_sim_t sim, const int32_t *pos_rc, const int32_t *n, int space, void *stream);
/* initialize_agents() (ffm_core.py:23-26; radius variant ffm_unified.py:131-171) on the device: for every
 * episode, n[e] distinct free cells (map == 0; with radius >= 0 only cells with |row-exit_row| + |col-exit_col|
 * <= radius, and n clamped to their number, :160-162) drawn uniformly without replacement -- the cells with the
 * n smallest Philox keys (stream PLACE, entity = ordinal of the cell among the eligible cells in row-major
 * order, keyed by the GLOBAL episode id), in key order.  Replaces ffm_set_positions for batched runs; zeroes
 * the DFF and the counters.  n: int32 [B] host array.  Any map size: a histogram of the keys selects the n + few
 * smallest, which are then sorted in one SM's shared memory. */
int ffm_place(ffm_sim_t sim, const int32_t *n, int32_t exit_row, int32_t exit_col, int32_t radius, void *stream);
/* `.positions` read (main.py:44,46) */
int ffm_get_positions(ffm_sim_t sim, int32_t *pos_rc, int32_t *n, int space, void *stream);

/* `.dff` read / assignment (run_trained_ffm.py:236); float32 [B][H][W] */
int ffm_set_dff(ffm_sim_t sim, const float *dff, int space, void *stream);
int ffm_get_dff(ffm_sim_t sim, float *dff, int space, void *stream);
/* update_dff() as a stand-alone call (ffm_core.py:106-117, ffm_unified.py:779-798, ffm_trained_core.py:333-353): one decay +
 * diffusion pass over the DFF of every episode, on the device (inside ffm_rollout the step kernels do it themselves) */
int ffm_update_dff(ffm_sim_t sim, void *stream);

/* step() x max_steps, stopping each episode when nobody is left: the loop of run()
 * (ffm_core.py:119-126, main.py:44-46).  Continues from the current state and step counter. */
int ffm_rollout(ffm_sim_t sim, int32_t max_steps, const ffm_draws_t *draws, const ffm_rollout_out_t *out,
                void *stream);

/* Probe for parity tests ("move probabilities within 1e-6 relative"), base model only: the probability vector
 * every pedestrian of the CURRENT state would sample from in the next step (ffm_core.py:74-83), computed with
 * the rollout kernel's own arithmetic.  probs double [B][n_max][neighborhood+1] in slot order (neighbours in
 * get_neighbors() order, then "stay"; 0 for non-candidates); kind int32 [B][n_max]: 0 no candidate / no
 * request (:63), 1 forced exit (one-hot, :66-72), 2 sampled. */
int ffm_move_probs(ffm_sim_t sim, double *probs, int32_t *kind, int space, void *stream);

/* per-episode counters since the last ffm_set_positions: steps executed (what run() returns,
 * ffm_core.py:126) and pedestrian-steps processed (sum over steps of the alive count).
 * Either pointer may be NULL. */
int ffm_get_counters(ffm_sim_t sim, int32_t *steps, int64_t *ped_steps, int space, void *stream);

/* ---- tables of the unified / trained models ----------------------------------------------------
 * Dense images of the reference's dicts: state id = (bx*nby + by)*256 + rU*64 + rD*16 + rL*4 + rR with
 * bx = row / block_size, by = col / block_size, nby = ceil(width / block_size)
 * (key ((rU,rD,rL,rR),(bx,by)) of _encode_state, ffm_unified.py:188-269).
 *   V      double [S]        self.V            v_seen uint8 [S]  key present (a defaultdict read inserts)
 *   H      double [S][A]     self.H rows       h_seen uint8 [S]  row present;  A = neighbours + 1
 * Any pointer may be NULL (left untouched / not returned). */
int ffm_tables_shape(ffm_sim_t sim, int32_t *n_states, int32_t *n_actions);
int ffm_tables_set(ffm_sim_t sim, const double *V, const uint8_t *v_seen, const double *H, const uint8_t *h_seen,
                   int space, void *stream);   /* set_v_table() :823-830, pretrained_v_path :84-110, h_table_path */
int ffm_tables_get(ffm_sim_t sim, double *V, uint8_t *v_seen, double *H, uint8_t *h_seen, int space,
                   void *stream);              /* get_v_table() :814-821, get_h_table() :847-857 */
/* FFM_LEARN_BATCHED: caller-owned device buffers the rollouts accumulate into and the caller all-reduces (sum)
 * across GPUs -- normally four slices of ONE flat buffer, so that the exchange is a single collective:
 * dV double [S] sum of TD errors, dN double [S] visit counts, dF double [S] > 0 where a key of V was touched
 * (the defaultdict reads of ffm_unified.py:658,661 insert keys), dH double [S][A] sum of alpha_h*delta (may be NULL
 * for critic_only).  ffm_tables_apply_deltas then does, per state visited n times, V += (1-(1-alpha_v)^n) * dV/n
 * (n sequential updates towards the same targets), H += dH, marks the touched keys / visited rows present, zeroes
 * the deltas and refreshes the extremes of H.  Stream-ordered, no host synchronisation. */
int ffm_tables_bind_deltas(ffm_sim_t sim, double *dV, double *dN, double *dF, double *dH);
int ffm_tables_apply_deltas(ffm_sim_t sim, void *stream);
int ffm_set_epsilon(ffm_sim_t sim, double epsilon);   /* set_epsilon() :859-867 */
/* CUDA-graph replays: a captured ffm_rollout freezes its by-value parameters, but epsilon (set_epsilon) and the episode key
 * (reset()) change every round.  Bind a caller-owned DEVICE struct {double epsilon; uint32_t episode_base; uint32_t pad}: the
 * following rollouts read both from it (update it with a stream-ordered copy before each replay); NULL unbinds. */
int ffm_bind_dynamic(ffm_sim_t sim, const void *dyn);
/* global id of episode 0 for the following rollouts (a drop-in object advances it on every reset()) */
int ffm_set_episode_base(ffm_sim_t sim, uint32_t episode_base);

/* ---- Q table of FFM_MODEL_MCQ -----------------------------------------------------------------------
 * self.Q (ffm_learning_core.py:75) is an open-addressing hash table in device memory: 64-bit key = ((tx/3)*nby + ty/3) * 4^9 +
 * sum_i v_i * 4^i over the row-major 3x3 window around the TARGET cell (v = map code with OOB = 2, +1 on occupied free
 * cells, _combined3x3_at_target :115-140), nby = ceil(width/3); rows float32 [5].  Only rows that exist in the dict take a
 * slot (rows are created by _ensure_qvec :289-291, never by the read path :190-191), so the table's size follows the
 * visited states, not the map.  ffm_q_get copies the whole table out: keys uint64 [capacity] (all ones = free slot), rows
 * float32 [capacity][5]; ffm_q_set clears it and inserts n (key, row) pairs (`model.Q = shared_Q`, main_learning.py:81). */
int ffm_q_shape(ffm_sim_t sim, int64_t *capacity);
int ffm_q_get(ffm_sim_t sim, uint64_t *keys, float *rows, int space, void *stream);
int ffm_q_set(ffm_sim_t sim, const uint64_t *keys, const float *rows, int64_t n, int space, void *stream);
/* Coverage pretrain (run_coverage_pretrain_and_training.py:91-166, force_first_step_and_roll): after ffm_set_positions placed
 * ONE agent per episode on its source cell, give every episode its teacher-forced first transition -- target cell T (linear
 * index, -1 = none), the FROM_* action (0..4) and the cap on CA steps after which finalize_timeouts() runs
 * (min(200, max(1, SFF(src) + 10)), :150-162).  Host arrays int32 [B].  Cleared by the next ffm_set_positions. */
int ffm_mcq_set_forced(ffm_sim_t sim, const int32_t *target_cell, const int32_t *from_dir, const int32_t *step_cap, void *stream);
/* After a rollout with FFM_LEARN_BATCHED (table frozen, finish order of every path recorded):
 *   ffm_mcq_backup_ordered   the reference's reverse Monte-Carlo backups (:262-278, :350-355) of the whole batch in episode
 *                            order -- bit-identical to running the episodes one after the other on the shared dict whenever
 *                            the policy did not read Q (beta = 1: coverage pretrain, warm-up episodes)
 *   ffm_mcq_accumulate       returns summed per (row, action) with visit counts into the handle's delta tables
 *   ffm_mcq_export_deltas    multi-GPU exchange by KEY: touched rows -> keys uint64 [capacity], rows float64 [capacity][10]
 *                            (sum G[5], n[5]), *count = rows written (device pointers); clears the local delta tables
 *   ffm_mcq_import_deltas    one rank's exported list added into the local delta tables (import the lists in rank order);
 *                            count = rows to import, or, with count_dev != NULL (a device pointer: the exporter's count that
 *                            travelled with its list), min(*count_dev, count) -- no host synchronisation anywhere in the exchange
 *   ffm_mcq_fold             Q += (1 - (1 - alpha)^n) (sum G / n - Q) per touched entry, delta tables zeroed */
int ffm_mcq_backup_ordered(ffm_sim_t sim, void *stream);
int ffm_mcq_accumulate(ffm_sim_t sim, void *stream);
int ffm_mcq_export_deltas(ffm_sim_t sim, uint64_t *keys, double *rows, int64_t capacity, uint32_t *count, void *stream);
int ffm_mcq_import_deltas(ffm_sim_t sim, const uint64_t *keys, const double *rows, uint32_t count, const uint32_t *count_dev, void *stream);
int ffm_mcq_fold(ffm_sim_t sim, void *stream);
int ffm_set_beta(ffm_sim_t sim, double beta);   /* the beta argument of step(beta), ffm_learning_core.py:145 */
int ffm_mcq_finalize_timeouts(ffm_sim_t sim, void *stream);   /* finalize_timeouts(), ffm_learning_core.py:326-360 */

/* Static-floor-field generation for n_maps maps (uint8 [n_maps][H][W]) -> out [n_maps][H][W] of
 * out_dtype (FFM_F32 | FFM_F64), +inf on non-walkable and unreachable cells.
 *   FFM_SFF_L1 / L2 / LINF   obstacle-blind min-over-exits norm: Create_SFF.py:14-33 (L2 = correctly
 *                            rounded hypot, as in the shipped data/sff/distance_L2.npy),
 *                            create_12x12_map_and_sff.py:36-50
 *   FFM_SFF_BFS4 / BFS8      geodesic distance in 4-/8-connected unit steps (wavefront BFS levels)
 *   FFM_SFF_DIJKSTRA8        geodesic distance with step costs (1, float32(sqrt 2)), float32 sums
 * `rounds` (may be NULL; non-NULL costs one 4-byte read-back + synchronisation) receives the number of tile visits of the
 * geodesic modes' work queue.
 * Host-space calls return with `out` filled; device-space calls are stream-ordered (the geodesic modes never synchronise the
 * host: one launch of persistent CTAs driven by a device-side tile queue). */
enum { FFM_SFF_L1 = 0, FFM_SFF_L2 = 1, FFM_SFF_LINF = 2, FFM_SFF_BFS4 = 3, FFM_SFF_BFS8 = 4, FFM_SFF_DIJKSTRA8 = 5 };
int ffm_sff_generate(const uint8_t *maps, int32_t n_maps, int32_t height, int32_t width, int32_t mode, int32_t out_dtype,
                     void *out, int space, int32_t device, void *stream, int32_t *rounds);

/* Discounted returns of a rollout buffer (device pointers): for every path (b, n) of length len[b][n],
 *   G[b][t][n] = reward[b][t][n] + gamma * G[b][t+1][n],  G beyond the path's end = 0      (float64)
 * -- the reverse scan `G = r + self.gamma * G` of ffm_learning_core.py:262-278, 350-355.  reward / G are
 * [B][T][N] (time-major per episode: consecutive pedestrians are consecutive in memory). */
int ffm_rollout_returns(const float *reward, const int32_t *len, int32_t n_episodes, int32_t steps, int32_t n_max,
                        double gamma, double *returns, int32_t device, void *stream);

/* Roofline denominator of the shared-memory-resident kernels, measured on `device`: conflict-free 16-byte shared
 * loads from every SM (GB/s; nominal 128 B/clk/SM), and the SM clock the driver reports (MHz).  Synchronous. */
int ffm_measure_smem_bandwidth(int32_t device, double *gb_per_s, double *sm_clock_mhz);

/* number of kernels this handle has launched so far (bench.py "gpu_launches") */
int64_t ffm_launch_count(ffm_sim_t sim);

/* dynamic shared memory per CTA and CTAs/SM the rollout kernel of this handle runs with */
int ffm_kernel_info(ffm_sim_t sim, int32_t *smem_bytes, int32_t *threads, int32_t *ctas_per_sm,
                    int32_t *fields_in_smem);

/* thread-block-cluster geometry of the rollout kernel: CTAs per episode (1 = no cluster; the map is split into row
 * bands held in distributed shared memory otherwise), clusters the device keeps resident at once, whether the score
 * field is staged on chip, and which kernel runs (0 = pedestrian-centric, 1 = cell-centric; -1 for the other models) */
int ffm_cluster_info(ffm_sim_t sim, int32_t *cluster, int32_t *max_clusters, int32_t *score_in_smem, int32_t *cell_kernel);

/* ---- Legacy 13-cell models ----------------------------------------------------------------------
 * model/ffm_ac_core.py FloorFieldModel (TD(0) critic over the FFM policy: `step` :111-244, `_encode_state` :62-109,
 * `_update_critic` :246-296) and model/ffm_actor_only.py FloorFieldModelActorOnly (`step` :150-413 -- including the
 * per-neighbour repetition of its decision block, :214-355 --, `_update_critic` :415-474, `_update_actor` :476-540).
 * The state of a pedestrian is the 13 cell values around it plus a coarse block index; the reference keys its dict tables
 * with pickle.dumps((tuple(state_13), (bx, by))).  Here the tables are open-addressing hash tables in HBM keyed by
 *     key = ((bx * nby + by) << 26) | sum_j cell_j << (2 j),   j = 0..12 in state_13 order, nby = ceil(width / block_size)
 * (3x3 block row-major, then the cells two steps up, down, left, right).  A handle of its own: the legacy models share
 * nothing with ffm_sim_t but the draw streams.  All buffers of these calls are HOST buffers; the calls return when done. */
enum {
    FFM_LEGACY_AC = 0,         /* model/ffm_ac_core.py */
    FFM_LEGACY_ACTOR_ONLY = 1  /* model/ffm_actor_only.py */
};
enum { FFM_LEGACY_TABLE_V = 0, FFM_LEGACY_TABLE_H = 1 };
typedef struct ffm_legacy_config {
    int32_t abi_version, device;
    int32_t height, width;
    int32_t neighborhood, sff_dtype;        /* AC: dtype of the SFF file (ffm_ac_core.py:28); actor-only: FFM_F32 (:45-48) */
    int32_t n_episodes, n_max;
    int32_t model, learn;                   /* FFM_LEGACY_*; FFM_LEARN_NONE (frozen tables, any n_episodes) | FFM_LEARN_EXACT (n_episodes == 1) */
    int32_t block_size, table_log2_capacity;/* params["block_size"] (actor-only: 5, ffm_actor_only.py:144); slots per table = 2^k, 0 = 2^20 */
    double k_S, k_D, k_A;
    float dff_c0, dff_c1, dff_threshold, reserved0;
    double gamma, alpha_v, alpha_h, exit_reward, step_penalty, collision_penalty, epsilon;
    double sff_min, sff_max;                /* actor-only: extremes of the inf->0 float32 SFF (:277-278) */
    uint64_t seed;
    uint32_t episode_base, reserved1;
} ffm_legacy_config_t;
typedef struct ffm_legacy_s *ffm_legacy_t;

int ffm_legacy_create(const ffm_legacy_config_t *cfg, ffm_legacy_t *out);         /* __init__: ffm_ac_core.py:9-38, ffm_actor_only.py:21-80 */
int ffm_legacy_destroy(ffm_legacy_t h);
/* map uint8 [H][W] (codes 0..3, no free cell on the border), sff [H][W] of cfg.sff_dtype */
int ffm_legacy_set_fields(ffm_legacy_t h, const uint8_t *map, const void *sff);
/* positions int32 [n_episodes][n_max][2] (row, col), counts int32 [n_episodes]; step counters restart at 0 (reset()) */
int ffm_legacy_set_positions(ffm_legacy_t h, const int32_t *pos_rc, const int32_t *n);
int ffm_legacy_get_positions(ffm_legacy_t h, int32_t *pos_rc, int32_t *n);
int ffm_legacy_set_dff(ffm_legacy_t h, const float *dff);                         /* float32 [n_episodes][H][W] */
int ffm_legacy_get_dff(ffm_legacy_t h, float *dff);
int ffm_legacy_zero_dff(ffm_legacy_t h);                                          /* reset(): self.dff = np.zeros_like(...), ffm_ac_core.py:326 (device-side memset) */
int ffm_legacy_update_dff(ffm_legacy_t h);                                        /* update_dff(): ffm_ac_core.py:298-318 */
/* up to max_steps calls of step() per episode (run(): ffm_ac_core.py:362-390), stopping at evacuation.  traj (may be NULL):
 * uint32 [n_episodes][traj_steps][n_max] linear cells after each step, traj_n int32 [n_episodes][traj_steps] their counts. */
int ffm_legacy_rollout(ffm_legacy_t h, int32_t max_steps, uint32_t *traj, int32_t *traj_n, int32_t traj_steps);
int ffm_legacy_get_counters(ffm_legacy_t h, int32_t *steps_done, uint64_t *ped_steps);   /* [n_episodes] each */
/* the dict tables: number of keys; all (key, row) pairs (row width: 1 for V, neighborhood + 1 for H); replace the table
 * (set_v_table(): ffm_ac_core.py:333-340 -- `default_value` is what an unseen key reads as afterwards, -1.0 there) */
int ffm_legacy_table_size(ffm_legacy_t h, int32_t which, int64_t *n);
int ffm_legacy_table_get(ffm_legacy_t h, int32_t which, uint64_t *keys, double *rows, int64_t capacity, int64_t *n);
int ffm_legacy_table_set(ffm_legacy_t h, int32_t which, const uint64_t *keys, const double *rows, int64_t n, double default_value);
int ffm_legacy_set_epsilon(ffm_legacy_t h, double epsilon);                      /* set_epsilon(): ffm_actor_only.py:578-585 */
int ffm_legacy_set_episode_base(ffm_legacy_t h, uint32_t episode_base);
int64_t ffm_legacy_launch_count(ffm_legacy_t h);

#ifdef __cplusplus
}
#endif
#endif /* FFM_B200_H */
