// C ABI of libffm_b200 (include/ffm_b200.h): handle management, host<->device staging and kernel
// dispatch.  No torch, no C++ types across the boundary, no CPU fallback: every compute entry point
// launches a kernel on the configured device or fails with FFM_E_CUDA.
#include "../../include/ffm_b200.h"

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "ffm_core_kernel.cuh"
#include "ffm_cell_kernel.cuh"
#include "ffm_sff_kernels.cuh"
#include "ffm_unified_kernel.cuh"
#include "ffm_mcq_kernel.cuh"
#include "ffm_internal.h"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                               \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return fail(FFM_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

constexpr int MAX_SMEM_OPTIN = 232448;  // 227 KB per CTA on sm_100

}  // namespace

// error text of the calling thread, for the other translation units (ffm_legacy.cu)
int ffm::set_error(int code, const char* fmt, va_list ap) {
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    return code;
}

struct ffm_sim_s {
    ffm_config_t cfg;
    int HW;
    bool have_fields, have_positions;
    bool fields_in_smem;
    int threads;
    int smem_bytes;
    int ctas_per_sm;
    int64_t launches;
    // device state
    uint8_t* d_map;
    uint16_t* d_type_grid;
    void* d_sff;
    void* d_score;
    uint32_t* d_pos;
    int32_t* d_n;
    int32_t* d_t;
    unsigned long long* d_ped_steps;
    float* d_dff;
    float* d_dff_tmp;
    int32_t* d_pos_rc;   // staging for (row, col) pairs
    int32_t* d_err;      // device-side validation flag
    uint32_t* d_free; int* d_free_count; int32_t* d_n_req;   // placement: eligible cells, their number, requested counts
    int place_er, place_ec, place_radius, place_count;       // what d_free currently holds (-2: nothing)
    const void* kernel;  // selected rollout kernel
    bool cell_kernel;    // base model: the cell-centric kernel (ffm_cell_kernel.cuh); false = round-1 pedestrian-centric kernel
    int cluster;         // CTAs per episode (thread-block cluster, row bands in distributed shared memory); 1 = one CTA
    int max_clusters;    // clusters the device holds at once (cudaOccupancyMaxActiveClusters)
    int RW, RB;          // bitboard words per row, rows per band
    bool wall_in_smem;   // static wall bitboard staged in shared memory
    bool score_in_smem;  // score field staged in shared memory (cluster variants may leave it to L1/L2)
    uint32_t* d_wall_bits;
    // unified / trained models
    int S, A, nby;
    double* d_V; uint8_t* d_vseen; double* d_H; uint8_t* d_hseen;
    double* d_dV; double* d_dN; double* d_dH; double* d_dF;   // borrowed (caller-owned) delta tables
    bool hstats_stale;   // H or its presence flags were replaced by the caller: extremes must be rescanned before the next rollout
    ffm::HStats* d_hstats;
    double* d_blk_lo; double* d_blk_hi; int* d_blk_any;
    double epsilon;
    const void* d_dyn;   // caller-owned device struct {double epsilon; uint32 episode_base; uint32 pad} overriding both (CUDA-graph replays)
    // MC-Q model: the Q dict as an open-addressing hash table
    unsigned long long* d_qkeys; float* d_Q; unsigned int* d_qcount; uint32_t q_cap;
    uint32_t* d_path_state; uint8_t* d_path_code; int32_t* d_path_len; uint16_t* d_path_col;
    int32_t* d_path_shift; uint16_t* d_fin_order; int32_t* d_fin_count;
    int32_t* d_forced;        // [3][B]: target cell, from-direction, step cap of the teacher-forced mini-episodes
    bool have_forced;
    double* d_qG; double* d_qN;   // [q_cap][5] each: returns summed per (row, action) and their visit counts (batched learning)
    double beta;
};

// ---------------------------------------------------------------------------------------------
// small staging kernels
// ---------------------------------------------------------------------------------------------
namespace ffm {

// map codes -> type bits (+ guard band), and score = (-k_S) * sff in the SFF's own dtype
// (the first product of ffm_core.py:77; elementwise, so hoisting it out of the step is exact).
template <typename S>
__global__ void prep_fields_kernel(const uint8_t* map, const S* sff, uint16_t* type_grid, S* score, int H, int W,
                                   int nbr, S neg_ks, int32_t* err) {
    const int HW = H * W, G = W + 1;
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < HW + 2 * G; x += gridDim.x * blockDim.x) {
        const int c = x - G;
        uint32_t cell = WALL_CELL;      // guard band and the non-passable codes: blocked
        if (c >= 0 && c < HW) {
            const uint8_t m = map[c];
            if (m > 3) atomicOr(err, 1);
            const int r = c / W, col = c - r * W;
            // the step never bounds-checks neighbours (nor does ffm_core.py:45-53, which "relies on
            // border walls"): a free cell on the border would let a pedestrian read outside the map
            if ((r == 0 || r == H - 1 || col == 0 || col == W - 1) && m == FFM_CELL_FREE) atomicOr(err, 2);
            if (m == FFM_CELL_EXIT) cell = TYPE_EXIT << TYPE_SHIFT;
            if (m == FFM_CELL_PED) cell = PEDMARK_CELL;   // blocked, and "a pedestrian" to _encode_state (ffm_unified.py:235)
            if (m == FFM_CELL_FREE) {
                bool near = false;   // static flag: some neighbour (of the model's neighbourhood) is an exit
                for (int dr = -1; dr <= 1; ++dr)
                    for (int dc = -1; dc <= 1; ++dc) {
                        if ((dr == 0 && dc == 0) || (nbr == 4 && dr != 0 && dc != 0)) continue;
                        const int rr = r + dr, cc = col + dc;
                        if (rr >= 0 && rr < H && cc >= 0 && cc < W && map[rr * W + cc] == FFM_CELL_EXIT) near = true;
                    }
                cell = (near ? TYPE_NEAR_EXIT : TYPE_FREE) << TYPE_SHIFT;
            }
            score[c] = mul_rn(neg_ks, sff[c]);
        }
        type_grid[x] = (uint16_t)cell;
    }
}

// static bitboard of the cell-centric kernel: row br = r + 1, word bw = col / 32 + 1 (one guard row / word on every
// side), bit set = not passable or outside the map
__global__ void prep_wall_bits_kernel(const uint16_t* type_grid, uint32_t* wall_bits, int H, int W, int RW) {
    const int G = W + 1, total = (H + 2) * RW;
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < total; x += gridDim.x * blockDim.x) {
        const int br = x / RW, bw = x - br * RW;
        const int r = br - 1;
        uint32_t v = 0xffffffffu;
        if (r >= 0 && r < H && bw >= 1 && bw <= RW - 2) {
            v = 0u;
            for (int b = 0; b < 32; ++b) {
                const int col = (bw - 1) * 32 + b;
                const bool blocked = col >= W || ((type_grid[r * W + col + G] >> TYPE_SHIFT) == TYPE_WALL);
                if (blocked) v |= 1u << b;
            }
        }
        wall_bits[x] = v;
    }
}

__global__ void pack_positions_kernel(const int32_t* pos_rc, const int32_t* n, const uint16_t* type_grid, uint32_t* pos,
                                      int B, int n_max, int H, int W, int32_t* err) {
    const int G = W + 1;
    const long long total = (long long)B * n_max;
    for (long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(x / n_max), i = (int)(x - (long long)e * n_max);
        const int ne = n[e];
        if (ne < 0 || ne > n_max) { atomicOr(err, 4); continue; }
        if (i >= ne) continue;
        const int2 rc = reinterpret_cast<const int2*>(pos_rc)[x];
        if (rc.x < 0 || rc.x >= H || rc.y < 0 || rc.y >= W) { atomicOr(err, 8); continue; }
        const int c = rc.x * W + rc.y;
        const uint32_t ty = type_grid[c + G] >> TYPE_SHIFT;
        if (ty != TYPE_FREE && ty != TYPE_NEAR_EXIT) atomicOr(err, 16);   // initialize_agents(): map == 0 cells only
        pos[x] = (uint32_t)c;
    }
}

// ---- initial placement (initialize_agents) ---------------------------------------------------------
// eligible cells of the map in row-major order (np.argwhere(map == 0), optionally radius-limited)
__global__ void collect_free_cells_kernel(const uint16_t* type_grid, int H, int W, int er, int ec, int radius, uint32_t* cells, int* count) {
    // single CTA, ordered compaction by block-wide scan over chunks of blockDim.x cells
    __shared__ int warp_tot[32];
    __shared__ int base;
    const int G = W + 1, HW = H * W, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (int c0 = 0; c0 < HW; c0 += blockDim.x) {
        const int c = c0 + threadIdx.x;
        bool ok = false;
        if (c < HW) {
            const uint32_t ty = type_grid[c + G] >> TYPE_SHIFT;
            ok = (ty == TYPE_FREE || ty == TYPE_NEAR_EXIT);
            if (ok && radius >= 0) { const int r = c / W, col = c - r * W; ok = abs(r - er) + abs(col - ec) <= radius; }
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) warp_tot[warp] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < (int)blockDim.x / 32; ++w) { if (w < warp) before += warp_tot[w]; total += warp_tot[w]; }
        if (ok) cells[base + before + __popc(bal & ((1u << lane) - 1u))] = (uint32_t)c;
        __syncthreads();
        if (threadIdx.x == 0) base += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = base;
}

// one CTA per episode: the n cells with the smallest Philox keys, in key order.
//   1. histogram of the keys' top 12 bits over all F eligible cells -> the bin b in which the n-th smallest key lies
//   2. the cells of bins <= b (n .. n + |bin b| of them) are collected and bitonic-sorted by (key, ordinal) in
//      shared memory -- ties by ordinal make it the stable argsort of the host placement
//   3. the first n become the pedestrians, in key order
// Works for any F (the keys are recomputed instead of stored); `cap` = capacity of the candidate buffer (power of 2).
__global__ void __launch_bounds__(256)
place_kernel(const uint32_t* cells, const int* count_ptr, const int32_t* n_req, uint32_t* pos, int32_t* n_out, int n_max,
             unsigned long long seed, uint32_t episode_base, int cap, int32_t* err) {
    extern __shared__ __align__(16) unsigned char sm[];
    unsigned long long* key = reinterpret_cast<unsigned long long*>(sm);
    uint32_t* ord = reinterpret_cast<uint32_t*>(sm + (size_t)cap * 8);
    uint32_t* hist = reinterpret_cast<uint32_t*>(sm + (size_t)cap * 12);      // 4096 bins
    __shared__ uint32_t part[256];
    __shared__ int s_bin, s_m;
    const int e = blockIdx.x, F = *count_ptr, tid = threadIdx.x, lane = tid & 31;
    const uint32_t episode = episode_base + (uint32_t)e;
    int n = n_req[e];
    n = n < F ? n : F;                                                       // actual_N = min(N, available) (ffm_unified.py:160-162)
    n = n < n_max ? n : n_max;
    auto key_of = [&](int i) {
        const uint4 o = philox4x32_10((uint32_t)i, 0u, episode, STREAM_PLACE, (uint32_t)seed, (uint32_t)(seed >> 32));
        return ((unsigned long long)(o.x >> 5) << 26) | (unsigned long long)(o.y >> 6);     // the 53 bits of u0: same order as the double
    };
    for (int i = tid; i < 4096; i += 256) hist[i] = 0u;
    if (tid == 0) { s_bin = 4095; s_m = 0; }
    __syncthreads();
    for (int i = tid; i < F; i += 256) atomicAdd(&hist[(uint32_t)(key_of(i) >> 41)], 1u);
    __syncthreads();
    // bin of the n-th smallest key: thread t sums bins [16t, 16t+16), then a serial scan over the 256 partial sums
    uint32_t loc = 0;
    for (int k = 0; k < 16; ++k) loc += hist[tid * 16 + k];
    part[tid] = loc;
    __syncthreads();
    if (tid == 0 && n > 0) {
        uint32_t cum = 0;
        int t = 0;
        while (t < 255 && cum + part[t] < (uint32_t)n) { cum += part[t]; ++t; }
        int bb = t * 16;
        while (bb < t * 16 + 15 && cum + hist[bb] < (uint32_t)n) { cum += hist[bb]; ++bb; }
        s_bin = bb;
    }
    __syncthreads();
    const uint32_t bsel = (uint32_t)s_bin;
    for (int base = 0; base < F; base += 256) {
        const int i = base + tid;
        unsigned long long k = 0;
        const bool take = n > 0 && i < F && (uint32_t)((k = key_of(i)) >> 41) <= bsel;
        const uint32_t bal = __ballot_sync(0xffffffffu, take);
        if (bal != 0u) {
            int b0 = 0;
            if (lane == 0) b0 = atomicAdd(&s_m, __popc(bal));
            b0 = __shfl_sync(0xffffffffu, b0, 0);
            const int at = b0 + __popc(bal & ((1u << lane) - 1u));
            if (take && at < cap) { key[at] = k; ord[at] = (uint32_t)i; }
        }
    }
    __syncthreads();
    const int m = s_m;
    if (m > cap) { if (tid == 0) atomicOr(err, 32); return; }                 // candidate buffer too small (practically unreachable)
    int mpad = 2;
    while (mpad < m) mpad <<= 1;
    for (int i = m + tid; i < mpad; i += 256) { key[i] = ~0ULL; ord[i] = 0xFFFFFFFFu; }
    __syncthreads();
    for (int size = 2; size <= mpad; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int x = tid; x < mpad / 2; x += 256) {
                const int i = 2 * x - (x & (stride - 1)), j = i + stride;
                const bool up = (i & size) == 0;
                const unsigned long long ki = key[i], kj = key[j];
                const uint32_t oi = ord[i], oj = ord[j];
                const bool gt = ki > kj || (ki == kj && oi > oj);             // ties by ordinal: a stable argsort
                if (gt == up) { key[i] = kj; key[j] = ki; ord[i] = oj; ord[j] = oi; }
            }
            __syncthreads();
        }
    for (int i = tid; i < n; i += 256) pos[(size_t)e * n_max + i] = cells[ord[i]];
    if (tid == 0) n_out[e] = n;
}

__global__ void unpack_positions_kernel(const uint32_t* pos, const int32_t* n, int32_t* pos_rc, int B, int n_max, int W) {
    const long long total = (long long)B * n_max;
    for (long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(x / n_max), i = (int)(x - (long long)e * n_max);
        int2 rc = make_int2(-1, -1);
        if (i < n[e]) {
            const int c = (int)pos[x];
            rc.x = c / W;
            rc.y = c - rc.x * W;
        }
        reinterpret_cast<int2*>(pos_rc)[x] = rc;
    }
}

// stand-alone update_dff() (ffm_core.py:106-117 / ffm_unified.py:779-798 / ffm_trained_core.py:333-353) of every episode
// of a handle: the same stencil the rollout kernels run per step, one CTA per episode, fields in global memory
template <int NBR>
__global__ void __launch_bounds__(256) dff_update_kernel(const float* in, float* out, int H, int W, float c0, float c1, float thr) {
    const size_t off = (size_t)blockIdx.x * H * W;
    dff_decay_diffuse<NBR>(in + off, out + off, H, W, c0, c1, thr, threadIdx.x, make_stencil_geom(H, W, threadIdx.x, 256));
}

}  // namespace ffm

namespace {

const void* pick_kernel(bool f64, bool small, int nbr, bool dff, bool fs, int threads) {
    return f64 ? ffm::pick_core_kernel_f64(small, nbr, dff, fs, threads) : ffm::pick_core_kernel_f32(small, nbr, dff, fs, threads);
}

int check_device_flag(ffm_sim_t s, cudaStream_t st) {
    int32_t flag = 0;
    CU(cudaMemcpyAsync(&flag, s->d_err, sizeof(flag), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (flag == 0) return FFM_OK;
    CU(cudaMemsetAsync(s->d_err, 0, sizeof(int32_t), st));
    if (flag & 1) return fail(FFM_E_INVALID, "map_array holds codes outside {0,1,2,3}");
    if (flag & 2) return fail(FFM_E_INVALID, "map_array has a free cell on its border (the step relies on border walls)");
    if (flag & 4) return fail(FFM_E_INVALID, "pedestrian count outside [0, n_max]");
    if (flag & 8) return fail(FFM_E_INVALID, "pedestrian position outside the map");
    if (flag & 16) return fail(FFM_E_INVALID, "pedestrian placed on a cell that is not free (map != 0)");
    if (flag & 32) return fail(FFM_E_UNSUPPORTED, "device placement: candidate buffer overflow");
    if (flag & 256) return fail(FFM_E_UNSUPPORTED, "MC-Q exchange: more touched rows than the export list holds (raise export_capacity)");
    if (flag & 128) return fail(FFM_E_INVALID, "two pedestrians were placed on the same cell");
    if (flag & 64) return fail(FFM_E_UNSUPPORTED, "Q hash table more than half full: raise ffm_config_t.q_log2_capacity");
    return fail(FFM_E_INVALID, "device validation flag %d", flag);
}

int copy_in(void* dst, const void* src, size_t bytes, int space, cudaStream_t st) {
    CU(cudaMemcpyAsync(dst, src, bytes, space == FFM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    return FFM_OK;
}
int copy_out(void* dst, const void* src, size_t bytes, int space, cudaStream_t st) {
    CU(cudaMemcpyAsync(dst, src, bytes, space == FFM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    if (space == FFM_HOST) CU(cudaStreamSynchronize(st));
    return FFM_OK;
}

// everything of McqParams that does not depend on the launch
void fill_mcq_params(ffm_sim_t s, ffm::McqParams& M, int max_steps) {
    memset(&M, 0, sizeof(M));
    M.H = s->cfg.height; M.W = s->cfg.width; M.HW = s->HW; M.n_max = s->cfg.n_max; M.B = s->cfg.n_episodes;
    M.max_steps = max_steps < 0 ? 0 : max_steps; M.step_cap = s->cfg.step_cap; M.learn = s->cfg.learn; M.nby = s->nby;
    M.force_finalize = max_steps < 0 ? 1 : 0;
    M.type_grid = s->d_type_grid; M.sff = s->d_sff;
    M.kS = s->cfg.k_S; M.kD = s->cfg.k_D; M.kQ = s->cfg.k_A; M.beta = s->beta; M.alpha = s->cfg.alpha_v; M.gamma = s->cfg.gamma;
    M.rw[ffm::RW_STEP] = -s->cfg.step_penalty; M.rw[ffm::RW_STOP] = -s->cfg.stop_penalty; M.rw[ffm::RW_COLL] = -s->cfg.collision_penalty;
    M.rw[ffm::RW_EXIT] = s->cfg.exit_reward; M.rw[ffm::RW_TIMEOUT] = -s->cfg.timeout_penalty;
    M.c0 = s->cfg.dff_c0; M.c1 = s->cfg.dff_c1; M.thr = s->cfg.dff_threshold;
    M.pos = s->d_pos; M.n_alive = s->d_n; M.t_done = s->d_t; M.ped_steps = s->d_ped_steps;
    M.dff = s->d_dff; M.dff_tmp = s->d_dff_tmp;
    M.qkeys = s->d_qkeys; M.Q = s->d_Q; M.qmask = s->q_cap - 1u; M.q_count = s->d_qcount; M.err = s->d_err;
    M.path_state = s->d_path_state; M.path_code = s->d_path_code; M.path_len = s->d_path_len; M.path_col = s->d_path_col;
    M.path_rows = s->cfg.step_cap + 2; M.path_shift = s->d_path_shift; M.fin_order = s->d_fin_order; M.fin_count = s->d_fin_count;
    if (s->have_forced) {
        const int B = s->cfg.n_episodes;
        M.forced_target = s->d_forced; M.forced_dir = s->d_forced + B; M.ep_cap = s->d_forced + 2 * B;
    }
    M.seed = s->cfg.seed; M.episode_base = s->cfg.episode_base;
}

}  // namespace

extern "C" {

int ffm_abi_version(void) { return FFM_ABI_VERSION; }
const char* ffm_last_error(void) { return g_err; }

int ffm_create(const ffm_config_t* cfg, ffm_sim_t* out) {
    if (!cfg || !out) return fail(FFM_E_INVALID, "null argument");
    *out = nullptr;
    if (cfg->abi_version != FFM_ABI_VERSION) return fail(FFM_E_INVALID, "abi_version %d != %d", cfg->abi_version, FFM_ABI_VERSION);
    if (cfg->height < 3 || cfg->width < 3) return fail(FFM_E_INVALID, "map must be at least 3x3");
    if ((long long)cfg->height * cfg->width > (1 << 24)) return fail(FFM_E_UNSUPPORTED, "map larger than 4096x4096 cells");
    if (cfg->neighborhood != FFM_NEUMANN && cfg->neighborhood != FFM_MOORE) return fail(FFM_E_INVALID, "neighborhood must be 4 or 8");
    if (cfg->sff_dtype != FFM_F32 && cfg->sff_dtype != FFM_F64) return fail(FFM_E_INVALID, "sff_dtype must be FFM_F32 or FFM_F64");
    if (cfg->n_episodes < 1) return fail(FFM_E_INVALID, "n_episodes must be >= 1");
    if (cfg->n_max < 1 || cfg->n_max > ffm::MAX_PEDS) return fail(FFM_E_UNSUPPORTED, "n_max must be in [1, %d]", ffm::MAX_PEDS);
    if (!cfg->track_dff && cfg->k_D != 0.0) return fail(FFM_E_INVALID, "track_dff = 0 requires k_D == 0");
    if (cfg->model < FFM_MODEL_CORE || cfg->model > FFM_MODEL_MCQ) return fail(FFM_E_INVALID, "unknown model %d", cfg->model);
    const bool mcq = cfg->model == FFM_MODEL_MCQ;
    const bool unified = cfg->model != FFM_MODEL_CORE && !mcq;
    if (mcq) {
        if (cfg->neighborhood != FFM_NEUMANN) return fail(FFM_E_INVALID, "the MC-Q model moves on the von Neumann neighbourhood (ffm_learning_core.py:73)");
        if (cfg->learn < FFM_LEARN_NONE || cfg->learn > FFM_LEARN_BATCHED) return fail(FFM_E_INVALID, "unknown learn mode %d", cfg->learn);
        if (cfg->q_log2_capacity != 0 && (cfg->q_log2_capacity < 10 || cfg->q_log2_capacity > 30)) return fail(FFM_E_INVALID, "q_log2_capacity must be 0 (default) or in [10, 30]");
        if (cfg->learn == FFM_LEARN_EXACT && cfg->n_episodes != 1) return fail(FFM_E_INVALID, "FFM_LEARN_EXACT needs n_episodes == 1");
        if (cfg->step_cap < 1) return fail(FFM_E_INVALID, "step_cap (params[\"max_steps\"]) must be >= 1");
        if (!cfg->track_dff) return fail(FFM_E_INVALID, "the MC-Q model always tracks the DFF");
    }
    if (unified) {
        if (cfg->block_size < 1) return fail(FFM_E_INVALID, "block_size must be >= 1");
        if (cfg->learn < FFM_LEARN_NONE || cfg->learn > FFM_LEARN_BATCHED) return fail(FFM_E_INVALID, "unknown learn mode %d", cfg->learn);
        if (cfg->learn == FFM_LEARN_EXACT && cfg->n_episodes != 1)
            return fail(FFM_E_INVALID, "FFM_LEARN_EXACT reproduces the reference's sequential table updates and needs n_episodes == 1");
        if (cfg->model == FFM_MODEL_TRAINED && cfg->learn != FFM_LEARN_NONE) return fail(FFM_E_INVALID, "the trained-actor model does not learn");
        if (!cfg->track_dff) return fail(FFM_E_INVALID, "the unified models always track the DFF");
        if (cfg->model != FFM_MODEL_UNIFIED_CRITIC && cfg->sff_dtype != FFM_F32)
            return fail(FFM_E_INVALID, "the actor / trained modes score with the float32 SFF (inf -> 0 and astype(float32), ffm_unified.py:72-76)");
    }
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(FFM_E_INVALID, "device %d not present (%d visible)", cfg->device, ndev);
    CU(cudaSetDevice(cfg->device));
    int cc_major = 0;
    CU(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, cfg->device));
    if (cc_major != 10) return fail(FFM_E_UNSUPPORTED, "libffm_b200 is built for sm_100a only (device is sm_%d*)", cc_major);

    ffm_sim_s* s = new (std::nothrow) ffm_sim_s();
    if (!s) return fail(FFM_E_INVALID, "out of host memory");
    memset(s, 0, sizeof(*s));
    s->cfg = *cfg;
    s->HW = cfg->height * cfg->width;
    const int HW = s->HW, W = cfg->width, B = cfg->n_episodes, N = cfg->n_max;
    const int ssz = cfg->sff_dtype == FFM_F64 ? 8 : 4;
    const bool dff = cfg->track_dff != 0;

    // kernel variant: fields in shared memory when they fit, and as many threads as there is work per step
    // (rounded to a supported CTA size) without starving co-resident CTAs
    const bool core = !mcq && !unified;
    const int WW = (W + 31) / 32;
    s->RW = WW + 2;
    s->RB = cfg->height;
    s->cluster = 1;
    // Two kernels implement the base model.  The cell-centric one (bitboard movers, ffm_cell_kernel.cuh) wins on maps with
    // enough bitboard words to keep a CTA busy (C2: 5.4e10 vs 4.3e10 ped-steps/s); on small maps (12x12 .. 32x32: a
    // dozen words) the pedestrian-centric one (ffm_core_kernel.cuh) has less fixed work per step (C1: 1.5e10 vs 1.2e10).
    s->cell_kernel = core && cfg->height * WW >= 64;
    if (const char* ev = getenv("FFM_KERNEL")) s->cell_kernel = core && strcmp(ev, "ped") != 0;   // test / tuning override: cell | ped
    const int esz = HW <= 65536 ? 2 : 4;
    auto layout_total = [&](bool fs) -> long long {
        if (mcq) return fs ? (1LL << 30) : (long long)ffm::make_mlayout(HW, W, N).total;
        if (unified) return ffm::make_ulayout(HW, W, N, ssz, fs).total;
        if (s->cell_kernel) return ffm::make_cell_layout(s->RB, W, s->RW, N, ssz, esz, dff, fs).total;
        return ffm::make_layout(HW, W, N, ssz, dff, fs).total;
    };
    s->wall_in_smem = true;
    s->score_in_smem = true;
    if (s->cell_kernel && (long long)HW * W >= (1LL << 32)) s->cell_kernel = false;       // row = umulhi(cell, magic) needs cell * W < 2^32
    // A map whose per-cell state does not fit one SM's shared memory runs as a thread-block cluster: CL CTAs per episode,
    // each holding a band of rows (distributed shared memory).  FFM_CLUSTER=n forces a cluster size (tests, tuning).
    int force_cluster = 0;
    if (const char* ev = getenv("FFM_CLUSTER")) force_cluster = atoi(ev);
    if (s->cell_kernel && (force_cluster > 1 ||
                           ffm::make_cell_layout(s->RB, W, s->RW, N, ssz, esz, dff, false, false).total > (unsigned)MAX_SMEM_OPTIN)) {
        // Measured on BASELINE C3 (256x256, 10 000 pedestrians, DFF on; profiles/r2_c3_kernel_variants.jsonl), ped-steps/s:
        //   2-CTA cluster, 1024 threads, score + DFF in L2 (74 clusters resident, ping-pong working set 38 MB = L2-resident)  2.05e10
        //   2-CTA cluster, 512 threads, the same                                                                              1.98e10
        //   pedestrian-centric kernel, one CTA of 1024 threads per SM, fields in L2 (76 MB working set, 71 GB written back)   1.84e10
        //   4-CTA cluster, everything on chip (33 resident, three cluster barriers per step, 16 warps per SM)                 1.31e10
        //   4- / 8-CTA clusters with the fields in L2, two CTAs per SM                                                        1.27e10 / 1.02e10
        // So: the SMALLEST cluster whose owner grid + claim masks fit, with the fields left in L2; on-chip fields only on
        // request (FFM_FIELDS_SMEM=1).  FFM_KERNEL=ped keeps the pedestrian-centric kernel.
        const int cls[3] = {2, 4, 8};
        const bool want_smem = getenv("FFM_FIELDS_SMEM") != nullptr;
        bool found = false;
        for (int ci = 0; ci < 3 && !found; ++ci) {
            const int cl = cls[ci];
            if (force_cluster > 1 && cl != force_cluster) continue;
            const int rb = (cfg->height + cl - 1) / cl;
            for (int fsi = want_smem ? 1 : 0; fsi >= 0 && !found; --fsi) {
                if (fsi == 1 && getenv("FFM_FIELDS_GLOBAL")) continue;
                for (int si = fsi; si >= 0 && !found; --si) {
                    if (si == 1 && getenv("FFM_SCORE_GLOBAL")) continue;
                    for (int wi = 1; wi >= 0 && !found; --wi) {
                        if (ffm::make_cell_layout(rb, W, s->RW, N, ssz, esz, dff, fsi != 0, wi != 0, si != 0).total > (unsigned)MAX_SMEM_OPTIN) continue;
                        s->cluster = cl; s->RB = rb; s->score_in_smem = si != 0; s->wall_in_smem = wi != 0;
                        s->fields_in_smem = fsi != 0;
                        found = true;
                    }
                }
            }
        }
        if (!found) s->cell_kernel = false;
    }
    auto occupancy = [&](const void* k, int threads, int smem) {
        int occ = 0;
        if (!k) return 0;
        if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, threads, smem) != cudaSuccess) return 0;
        return occ;
    };
    const bool f64 = cfg->sff_dtype == FFM_F64;
    if (s->cell_kernel && s->cluster > 1) {
        // one cluster per episode
        s->threads = (s->cluster == 2 && !s->fields_in_smem) ? 1024 : 512;    // measured (C3): 2.05e10 vs 1.98e10; larger clusters / on-chip fields: 512
        if (const char* ev = getenv("FFM_THREADS")) { const int v = atoi(ev); if (v == 512 || (v == 1024 && s->cluster == 2)) s->threads = v; }   // tuning
        s->smem_bytes = (int)ffm::make_cell_layout(s->RB, W, s->RW, N, ssz, esz, dff, s->fields_in_smem, s->wall_in_smem, s->score_in_smem).total;
        s->kernel = ffm::pick_cell_kernel(f64, HW <= 65536, cfg->neighborhood, dff, s->fields_in_smem, s->threads, s->cluster);
        if (!s->kernel) { delete s; return fail(FFM_E_UNSUPPORTED, "no cluster variant of the rollout kernel for this configuration"); }
    } else if (s->cell_kernel) {
        // Cell-centric kernel: the rollout is bound by its three barriers per step, so what counts is how many CTAs
        // (independent barrier domains) an SM holds, then how many threads each has.  Measured on C2 (one B200):
        // 6 CTAs x 256 threads, fields in shared memory 73.2 ms; the same with fields read through L1 74.4 ms; 5 and 4 CTAs
        // (more registers) 77.8 / 86.1 ms; 10 CTAs x 128 threads, fields through L1, 68.1 ms.  Rule: maximise
        // CTAs/SM x sqrt(threads), with a small preference for shared-memory-resident fields.
        const int chunks = cfg->height * WW;              // 32-cell bitboard chunks per step
        const int tlist[3] = {128, 256, 1024};
        const int force_threads = getenv("FFM_THREADS") ? atoi(getenv("FFM_THREADS")) : 0;
        double best = -1.0;
        for (int ti = 0; ti < 3; ++ti) {
            const int th = tlist[ti];
            if (force_threads ? th != force_threads : ((th == 256 && chunks <= 128) || (th == 1024 && chunks <= 2048))) continue;
            for (int fsi = 1; fsi >= 0; --fsi)
                for (int wi = 1; wi >= 0; --wi) {
                    if (fsi == 1 && getenv("FFM_FIELDS_GLOBAL")) continue;
                    if (fsi == 0 && getenv("FFM_FIELDS_SMEM")) continue;
                    if (wi == 1 && getenv("FFM_WALL_GLOBAL")) continue;
                    const long long tot = ffm::make_cell_layout(s->RB, W, s->RW, N, ssz, esz, dff, fsi != 0, wi != 0).total;
                    if (tot > MAX_SMEM_OPTIN) continue;
                    const void* k = ffm::pick_cell_kernel(f64, HW <= 65536, cfg->neighborhood, dff, fsi != 0, th, 1);
                    const int occ = occupancy(k, th, (int)tot);
                    const double score = occ * sqrt((double)th) * (fsi ? 1.0 : 0.95) * (wi ? 1.0 : 0.99);
                    if (occ > 0 && score > best) {
                        best = score;
                        s->threads = th; s->fields_in_smem = fsi != 0; s->wall_in_smem = wi != 0;
                        s->smem_bytes = (int)tot; s->kernel = k; s->ctas_per_sm = occ;
                    }
                }
        }
        if (best < 0.0) { delete s; return fail(FFM_E_UNSUPPORTED, "no resident configuration of the rollout kernel for this map / capacity"); }
    } else {
        const long long tot_in = layout_total(true), tot_out = layout_total(false);
        if (tot_out > MAX_SMEM_OPTIN) {
            delete s;
            return fail(FFM_E_UNSUPPORTED, "episode state (%lld B) does not fit the 227 KB of shared memory of one SM", tot_out);
        }
        const int work = N > HW / 8 ? N : HW / 8;
        s->threads = work <= 128 ? 128 : (work <= 2048 ? 256 : 1024);
        if (const char* ev = getenv("FFM_THREADS")) {   // tuning override: 128 | 256 | 1024
            const int v = atoi(ev);
            if (v == 128 || v == 256 || v == 1024) s->threads = v;
        }
        if (unified || mcq) s->threads = N <= 128 ? 128 : 256;
        // small crowds on small maps (the 12x12 training configurations): two warps per episode -- more episodes (barrier
        // domains) per SM and fewer idle lanes
        if (unified && N <= 64 && HW <= 1024) s->threads = 64;
        if (unified) if (const char* ev = getenv("FFM_THREADS")) { const int v = atoi(ev); if (v == 64 || v == 128 || v == 256) s->threads = v; }
        // Fields (score, DFF) in shared memory or left in global memory (L1/L2)?  Shared memory is faster per access
        // but costs residency.  Rule: shared unless the global variant keeps at least 1.5x as many CTAs resident.
        auto kernel_for = [&](bool fs) {
            if (mcq) return ffm::pick_mcq_kernel(f64, s->threads);
            if (unified) return ffm::pick_unified_kernel(f64, cfg->neighborhood, fs, s->threads, cfg->model != FFM_MODEL_UNIFIED_CRITIC);
            return pick_kernel(f64, HW <= 65536, cfg->neighborhood, dff, fs, s->threads);
        };
        const int occ_out = occupancy(kernel_for(false), s->threads, (int)tot_out);
        const int occ_in = tot_in <= MAX_SMEM_OPTIN ? occupancy(kernel_for(true), s->threads, (int)tot_in) : 0;
        s->fields_in_smem = occ_in > 0 && 2 * occ_out < 3 * occ_in;
        if (getenv("FFM_FIELDS_GLOBAL")) s->fields_in_smem = false;          // tuning overrides
        if (getenv("FFM_FIELDS_SMEM") && occ_in > 0) s->fields_in_smem = true;
        s->smem_bytes = (int)(s->fields_in_smem ? tot_in : tot_out);
        s->kernel = kernel_for(s->fields_in_smem);
    }
    if (unified) {
        s->A = cfg->neighborhood + 1;
        s->nby = (W + cfg->block_size - 1) / cfg->block_size;
        s->S = ((cfg->height + cfg->block_size - 1) / cfg->block_size) * s->nby * 256;
        s->epsilon = cfg->epsilon;
    }
    cudaError_t ce = cudaFuncSetAttribute(s->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, s->smem_bytes);
    if (ce != cudaSuccess) { delete s; return fail(FFM_E_CUDA, "cudaFuncSetAttribute(smem=%d): %s", s->smem_bytes, cudaGetErrorString(ce)); }
    int occ = 0;
    ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, s->kernel, s->threads, s->smem_bytes);
    if (ce != cudaSuccess || occ < 1) { delete s; return fail(FFM_E_CUDA, "rollout kernel cannot be resident (smem=%d, threads=%d): %s", s->smem_bytes, s->threads, cudaGetErrorString(ce)); }
    s->ctas_per_sm = occ;
    if (s->cluster > 1) {
        cudaLaunchConfig_t lc;
        memset(&lc, 0, sizeof(lc));
        lc.gridDim = dim3((unsigned)s->cluster); lc.blockDim = dim3(s->threads); lc.dynamicSmemBytes = (size_t)s->smem_bytes;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)s->cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        lc.attrs = at; lc.numAttrs = 1;
        int ncl = 0;
        ce = cudaOccupancyMaxActiveClusters(&ncl, s->kernel, &lc);
        if (ce != cudaSuccess || ncl < 1) { delete s; return fail(FFM_E_CUDA, "no cluster of %d CTAs can be resident (smem=%d, threads=%d): %s", s->cluster, s->smem_bytes, s->threads, cudaGetErrorString(ce)); }
        s->max_clusters = ncl;
    }

#define ALLOC(ptr, bytes)                                                                     \
    do {                                                                                      \
        cudaError_t e_ = cudaMalloc((void**)&(ptr), (bytes));                                 \
        if (e_ != cudaSuccess) { ffm_destroy(s); return fail(FFM_E_CUDA, "cudaMalloc(%zu) failed: %s", (size_t)(bytes), cudaGetErrorString(e_)); } \
    } while (0)
    ALLOC(s->d_map, (size_t)HW);
    ALLOC(s->d_type_grid, (size_t)(HW + 2 * (W + 1)) * 2 + 16);   // padded: the rollout kernel bulk-copies it in 16-byte units
    ALLOC(s->d_wall_bits, (size_t)(cfg->height + 2) * s->RW * 4);
    ALLOC(s->d_sff, (size_t)HW * ssz);
    ALLOC(s->d_score, (size_t)HW * ssz);
    ALLOC(s->d_pos, (size_t)B * N * 4);
    ALLOC(s->d_pos_rc, (size_t)B * N * 8);
    ALLOC(s->d_n, (size_t)B * 4);
    ALLOC(s->d_t, (size_t)B * 4);
    ALLOC(s->d_ped_steps, (size_t)B * 8);
    ALLOC(s->d_err, 4);
    ALLOC(s->d_free, (size_t)HW * 4);
    ALLOC(s->d_free_count, 4);
    ALLOC(s->d_n_req, (size_t)B * 4);
    s->place_radius = -2;
    if (dff) {
        // both DFF buffers in one allocation
        const bool two = !s->fields_in_smem || mcq;   // the cluster variants keep the ping-pong buffer on chip too
        ALLOC(s->d_dff, (size_t)B * HW * 4 * (two ? 2 : 1));
        if (two) s->d_dff_tmp = s->d_dff + (size_t)B * HW;
    }
    if (mcq) {
        s->nby = (W + 2) / 3;
        s->q_cap = 1u << (cfg->q_log2_capacity ? cfg->q_log2_capacity : 21);
        const size_t rows = (size_t)cfg->step_cap + 2;
        ALLOC(s->d_qkeys, (size_t)s->q_cap * 8);
        ALLOC(s->d_Q, (size_t)s->q_cap * 5 * 4);
        ALLOC(s->d_qcount, 4);
        ALLOC(s->d_path_state, (size_t)B * rows * N * 4);
        ALLOC(s->d_path_code, (size_t)B * rows * N);
        ALLOC(s->d_path_len, (size_t)B * N * 4);
        ALLOC(s->d_path_col, (size_t)B * N * 2);
        ALLOC(s->d_path_shift, (size_t)B * 4);
        ALLOC(s->d_fin_order, (size_t)B * N * 2);
        ALLOC(s->d_fin_count, (size_t)B * 4);
        ALLOC(s->d_forced, (size_t)B * 3 * 4);
        cudaMemset(s->d_qkeys, 0xFF, (size_t)s->q_cap * 8);
        cudaMemset(s->d_Q, 0, (size_t)s->q_cap * 5 * 4);
        cudaMemset(s->d_qcount, 0, 4);
        cudaMemset(s->d_path_len, 0, (size_t)B * N * 4);
        cudaMemset(s->d_path_shift, 0, (size_t)B * 4);
        cudaMemset(s->d_fin_count, 0, (size_t)B * 4);
        s->beta = 1.0;
    }
    if (unified) {
        ALLOC(s->d_V, (size_t)s->S * 8);
        ALLOC(s->d_vseen, (size_t)s->S);
        ALLOC(s->d_H, (size_t)s->S * s->A * 8);
        ALLOC(s->d_hseen, (size_t)s->S);
        ALLOC(s->d_hstats, sizeof(ffm::HStats));
        ALLOC(s->d_blk_lo, 148 * 8);
        ALLOC(s->d_blk_hi, 148 * 8);
        ALLOC(s->d_blk_any, 148 * 4);
        cudaMemset(s->d_V, 0, (size_t)s->S * 8);
        cudaMemset(s->d_vseen, 0, (size_t)s->S);
        cudaMemset(s->d_H, 0, (size_t)s->S * s->A * 8);
        cudaMemset(s->d_hseen, 0, (size_t)s->S);
        cudaMemset(s->d_hstats, 0, sizeof(ffm::HStats));
    }
#undef ALLOC
    cudaMemset(s->d_err, 0, 4);
    cudaMemset(s->d_n, 0, (size_t)B * 4);
    cudaMemset(s->d_t, 0, (size_t)B * 4);
    cudaMemset(s->d_ped_steps, 0, (size_t)B * 8);
    *out = s;
    return FFM_OK;
}

int ffm_destroy(ffm_sim_t s) {
    if (!s) return FFM_OK;
    cudaSetDevice(s->cfg.device);
    cudaFree(s->d_map); cudaFree(s->d_type_grid); cudaFree(s->d_wall_bits); cudaFree(s->d_sff); cudaFree(s->d_score);
    cudaFree(s->d_pos); cudaFree(s->d_pos_rc); cudaFree(s->d_n); cudaFree(s->d_t);
    cudaFree(s->d_ped_steps); cudaFree(s->d_err); cudaFree(s->d_dff);
    cudaFree(s->d_free); cudaFree(s->d_free_count); cudaFree(s->d_n_req);
    cudaFree(s->d_V); cudaFree(s->d_vseen); cudaFree(s->d_H); cudaFree(s->d_hseen); cudaFree(s->d_hstats);
    cudaFree(s->d_blk_lo); cudaFree(s->d_blk_hi); cudaFree(s->d_blk_any);
    cudaFree(s->d_Q); cudaFree(s->d_qkeys); cudaFree(s->d_qcount); cudaFree(s->d_path_shift); cudaFree(s->d_fin_order); cudaFree(s->d_fin_count);
    cudaFree(s->d_forced); cudaFree(s->d_qG); cudaFree(s->d_qN); cudaFree(s->d_path_state); cudaFree(s->d_path_code); cudaFree(s->d_path_len); cudaFree(s->d_path_col);
    delete s;
    return FFM_OK;
}

int ffm_set_fields(ffm_sim_t s, const uint8_t* map, const void* sff, int space, void* stream) {
    if (!s || !map || !sff) return fail(FFM_E_INVALID, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(s->cfg.device));
    const int HW = s->HW, H = s->cfg.height, W = s->cfg.width;
    const int ssz = s->cfg.sff_dtype == FFM_F64 ? 8 : 4;
    int rc;
    if ((rc = copy_in(s->d_map, map, (size_t)HW, space, st))) return rc;
    if ((rc = copy_in(s->d_sff, sff, (size_t)HW * ssz, space, st))) return rc;
    const int blocks = (HW + 2 * (W + 1) + 255) / 256;
    if (s->cfg.sff_dtype == FFM_F64)
        ffm::prep_fields_kernel<double><<<blocks, 256, 0, st>>>(s->d_map, (const double*)s->d_sff, s->d_type_grid, (double*)s->d_score, H, W, s->cfg.neighborhood, -s->cfg.k_S, s->d_err);
    else
        ffm::prep_fields_kernel<float><<<blocks, 256, 0, st>>>(s->d_map, (const float*)s->d_sff, s->d_type_grid, (float*)s->d_score, H, W, s->cfg.neighborhood, (float)(-s->cfg.k_S), s->d_err);
    CU(cudaGetLastError());
    ffm::prep_wall_bits_kernel<<<((H + 2) * s->RW + 255) / 256, 256, 0, st>>>(s->d_type_grid, s->d_wall_bits, H, W, s->RW);
    CU(cudaGetLastError());
    s->launches += 2;
    if ((rc = check_device_flag(s, st))) return rc;
    s->have_fields = true;
    s->place_radius = -2;
    return FFM_OK;
}

int ffm_place(ffm_sim_t s, const int32_t* n, int32_t exit_row, int32_t exit_col, int32_t radius, void* stream) {
    if (!s || !n) return fail(FFM_E_INVALID, "null argument");
    if (!s->have_fields) return fail(FFM_E_STATE, "ffm_set_fields must precede ffm_place");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(s->cfg.device));
    const int B = s->cfg.n_episodes, N = s->cfg.n_max;
    if (radius < 0) { radius = -1; exit_row = exit_col = 0; }
    if (s->place_radius != radius || s->place_er != exit_row || s->place_ec != exit_col) {
        ffm::collect_free_cells_kernel<<<1, 1024, 0, st>>>(s->d_type_grid, s->cfg.height, s->cfg.width, exit_row, exit_col, radius, s->d_free, s->d_free_count);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(&s->place_count, s->d_free_count, 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        s->place_radius = radius; s->place_er = exit_row; s->place_ec = exit_col;
        s->launches++;
    }
    for (int e = 0; e < B; ++e)
        if (n[e] < 0) return fail(FFM_E_INVALID, "negative pedestrian count");
    if (radius < 0)
        for (int e = 0; e < B; ++e)
            if (n[e] > s->place_count || n[e] > N)    // np.random.choice(len(free), N, replace=False) raises (ffm_core.py:25)
                return fail(FFM_E_INVALID, "Cannot take a larger sample than population when 'replace=False' (%d pedestrians, %d free cells, capacity %d)", n[e], s->place_count, N);
    int nmax_req = 0;
    for (int e = 0; e < B; ++e) nmax_req = n[e] > nmax_req ? n[e] : nmax_req;
    nmax_req = nmax_req < N ? nmax_req : N;
    const int want = (s->place_count < nmax_req + 2048 ? s->place_count : nmax_req + 2048);   // n .. n + |threshold bin| candidates
    int npad = 2;
    while (npad < want) npad <<= 1;
    if (npad > 16384) npad = 16384;
    const size_t smem = (size_t)npad * 12 + 4096 * 4;
    CU(cudaFuncSetAttribute(ffm::place_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CU(cudaMemcpyAsync(s->d_n_req, n, (size_t)B * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(s->d_t, 0, (size_t)B * 4, st));
    CU(cudaMemsetAsync(s->d_ped_steps, 0, (size_t)B * 8, st));
    if (s->d_dff) CU(cudaMemsetAsync(s->d_dff, 0, (size_t)B * s->HW * 4, st));
    ffm::place_kernel<<<B, 256, smem, st>>>(s->d_free, s->d_free_count, s->d_n_req, s->d_pos, s->d_n, N, s->cfg.seed, s->cfg.episode_base, npad, s->d_err);
    CU(cudaGetLastError());
    s->launches++;
    int rcf = check_device_flag(s, st);  // synchronises (n[] is a host buffer of the caller) and surfaces a candidate-buffer overflow
    if (rcf) return rcf;
    s->have_positions = true;
    s->have_forced = false;
    return FFM_OK;
}

int ffm_set_positions(ffm_sim_t s, const int32_t* pos_rc, const int32_t* n, int space, void* stream) {
    if (!s || !pos_rc || !n) return fail(FFM_E_INVALID, "null argument");
    if (!s->have_fields) return fail(FFM_E_STATE, "ffm_set_fields must precede ffm_set_positions");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(s->cfg.device));
    const int B = s->cfg.n_episodes, N = s->cfg.n_max;
    int rc;
    const int32_t* src = pos_rc;
    if (space == FFM_HOST) {
        if ((rc = copy_in(s->d_pos_rc, pos_rc, (size_t)B * N * 8, FFM_HOST, st))) return rc;
        src = s->d_pos_rc;
    }
    if ((rc = copy_in(s->d_n, n, (size_t)B * 4, space, st))) return rc;
    CU(cudaMemsetAsync(s->d_t, 0, (size_t)B * 4, st));
    CU(cudaMemsetAsync(s->d_ped_steps, 0, (size_t)B * 8, st));
    if (s->d_dff) CU(cudaMemsetAsync(s->d_dff, 0, (size_t)B * s->HW * 4, st));
    const long long total = (long long)B * N;
    const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    ffm::pack_positions_kernel<<<blocks, 256, 0, st>>>(src, s->d_n, s->d_type_grid, s->d_pos, B, N, s->cfg.height, s->cfg.width, s->d_err);
    CU(cudaGetLastError());
    s->launches++;
    s->have_positions = true;
    s->have_forced = false;
    return FFM_OK;
}

int ffm_get_positions(ffm_sim_t s, int32_t* pos_rc, int32_t* n, int space, void* stream) {
    if (!s) return fail(FFM_E_INVALID, "null argument");
    if (!s->have_positions) return fail(FFM_E_STATE, "no positions set");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(s->cfg.device));
    const int B = s->cfg.n_episodes, N = s->cfg.n_max;
    int rc;
    if ((rc = check_device_flag(s, st))) return rc;
    if (pos_rc) {
        const long long total = (long long)B * N;
        const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
        int32_t* dst = space == FFM_DEVICE ? pos_rc : s->d_pos_rc;
        ffm::unpack_positions_kernel<<<blocks, 256, 0, st>>>(s->d_pos, s->d_n, dst, B, N, s->cfg.width);
        CU(cudaGetLastError());
        s->launches++;
        if (space == FFM_HOST && (rc = copy_out(pos_rc, s->d_pos_rc, (size_t)B * N * 8, FFM_HOST, st))) return rc;
    }
    if (n && (rc = copy_out(n, s->d_n, (size_t)B * 4, space, st))) return rc;
    return FFM_OK;
}

int ffm_set_dff(ffm_sim_t s, const float* dff, int space, void* stream) {
    if (!s || !dff) return fail(FFM_E_INVALID, "null argument");
    if (!s->d_dff) return fail(FFM_E_STATE, "handle was created with track_dff = 0");
    CU(cudaSetDevice(s->cfg.device));
    return copy_in(s->d_dff, dff, (size_t)s->cfg.n_episodes * s->HW * 4, space, (cudaStream_t)stream);
}

int ffm_get_dff(ffm_sim_t s, float* dff, int space, void* stream) {
    if (!s || !dff) return fail(FFM_E_INVALID, "null argument");
    if (!s->d_dff) return fail(FFM_E_STATE, "handle was created with track_dff = 0");
    CU(cudaSetDevice(s->cfg.device));
    return copy_out(dff, s->d_dff, (size_t)s->cfg.n_episodes * s->HW * 4, space, (cudaStream_t)stream);
}

int ffm_update_dff(ffm_sim_t s, void* stream) {
    if (!s) return fail(FFM_E_INVALID, "null argument");
    if (!s->d_dff) return fail(FFM_E_STATE, "handle was created with track_dff = 0");
    if (s->cfg.model == FFM_MODEL_MCQ) return fail(FFM_E_UNSUPPORTED, "the MC-Q model has no public update_dff()");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(s->cfg.device));
    const size_t bytes = (size_t)s->cfg.n_episodes * s->HW * 4;
    float* tmp = s->d_dff_tmp;
    if (!tmp) CU(cudaMallocAsync((void**)&tmp, bytes, st));
    if (s->cfg.neighborhood == FFM_MOORE)
        ffm::dff_update_kernel<8><<<s->cfg.n_episodes, 256, 0, st>>>(s->d_dff, tmp, s->cfg.height, s->cfg.width, s->cfg.dff_c0, s->cfg.dff_c1, s->cfg.dff_threshold);
    else
        ffm::dff_update_kernel<4><<<s->cfg.n_episodes, 256, 0, st>>>(s->d_dff, tmp, s->cfg.height, s->cfg.width, s->cfg.dff_c0, s->cfg.dff_c1, s->cfg.dff_threshold);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(s->d_dff, tmp, bytes, cudaMemcpyDeviceToDevice, st));
    if (!s->d_dff_tmp) CU(cudaFreeAsync(tmp, st));
    s->launches++;
    return FFM_OK;
}

int ffm_rollout(ffm_sim_t s, int32_t max_steps, const ffm_draws_t* draws, const ffm_rollout_out_t* out, void* stream) {
    if (!s) return fail(FFM_E_INVALID, "null argument");
    if (!s->have_fields || !s->have_positions) return fail(FFM_E_STATE, "fields and positions must be set before ffm_rollout");
    if (max_steps < 0 && !(max_steps == -1 && s->cfg.model == FFM_MODEL_MCQ)) return fail(FFM_E_INVALID, "max_steps < 0");
    if (draws && draws->space != FFM_DEVICE) return fail(FFM_E_INVALID, "recorded draws must live in device memory");
    if (out && out->ctraj && !(s->cfg.model == FFM_MODEL_CORE && s->cell_kernel))
        return fail(FFM_E_UNSUPPORTED, "the compact trajectory record is written by the cell-centric base-model kernel (ffm_cluster_info: cell_kernel == 1)");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(s->cfg.device));
    // the opt-in shared-memory size is per-function state shared by all handles: re-assert ours before launching
    CU(cudaFuncSetAttribute(s->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, s->smem_bytes));
    if (s->cfg.model == FFM_MODEL_MCQ) {
        ffm::McqParams M;
        fill_mcq_params(s, M, max_steps);
        if (draws) { M.move_draws = draws->move; M.conflict_draws = draws->conflict; M.draw_steps = draws->steps; M.draw_first = draws->first_step; }
        if (out && out->traj_cells) {
            if (!out->traj_n) return fail(FFM_E_INVALID, "traj_cells without traj_n");
            M.traj = out->traj_cells; M.traj_n = out->traj_n; M.traj_steps = out->traj_steps;
        }
        void* margs[] = {&M};
        CU(cudaLaunchKernel(s->kernel, dim3(s->cfg.n_episodes), dim3(s->threads), margs, (size_t)s->smem_bytes, st));
        s->launches++;
        return FFM_OK;
    }
    if (s->cfg.model != FFM_MODEL_CORE) {
        if (s->cfg.learn == FFM_LEARN_BATCHED && !s->d_dV) return fail(FFM_E_STATE, "FFM_LEARN_BATCHED needs ffm_tables_bind_deltas first");
        ffm::UnifiedParams U;
        memset(&U, 0, sizeof(U));
        U.H = s->cfg.height; U.W = s->cfg.width; U.HW = s->HW; U.n_max = s->cfg.n_max; U.B = s->cfg.n_episodes;
        U.max_steps = max_steps;
        U.mode = s->cfg.model - 1; U.learn = s->cfg.learn;
        U.block_size = s->cfg.block_size; U.nby = s->nby; U.S = s->S;
        U.magic_w = (uint32_t)(((1ULL << 32) + (uint64_t)U.W - 1) / (uint64_t)U.W);
        U.magic_bs = s->cfg.block_size > 1 ? (uint32_t)(((1ULL << 32) + (uint64_t)s->cfg.block_size - 1) / (uint64_t)s->cfg.block_size) : 0u;
        U.type_grid = s->d_type_grid; U.score = s->d_score;
        U.kd = (float)s->cfg.k_D; U.c0 = s->cfg.dff_c0; U.c1 = s->cfg.dff_c1; U.thr = s->cfg.dff_threshold;
        U.kA = s->cfg.k_A; U.gamma = s->cfg.gamma; U.alpha_v = s->cfg.alpha_v; U.alpha_h = s->cfg.alpha_h;
        U.exit_reward = s->cfg.exit_reward; U.step_penalty = s->cfg.step_penalty; U.collision_penalty = s->cfg.collision_penalty;
        U.epsilon = s->epsilon; U.sff_min = s->cfg.sff_min; U.sff_max = s->cfg.sff_max;
        U.pos = s->d_pos; U.n_alive = s->d_n; U.t_done = s->d_t; U.ped_steps = s->d_ped_steps;
        U.dff = s->d_dff; U.dff_tmp = s->d_dff_tmp;
        U.V = s->d_V; U.v_seen = s->d_vseen; U.Hm = s->d_H; U.h_seen = s->d_hseen; U.dV = s->d_dV; U.dN = s->d_dN; U.dH = s->d_dH; U.dF = s->d_dF;
        if (s->hstats_stale && s->cfg.model != FFM_MODEL_UNIFIED_CRITIC) {
            // the caller replaced H: one small kernel recomputes its extremes before any CTA of the rollout looks at them
            CU(ffm::launch_rescan_hstats(s->d_H, s->d_hseen, s->S, s->A, s->d_hstats, s->d_blk_lo, s->d_blk_hi, s->d_blk_any, 148, st));
            s->launches += 2;
        }
        s->hstats_stale = false;
        U.hstats = s->d_hstats;
        U.seed = s->cfg.seed; U.episode_base = s->cfg.episode_base; U.err = s->d_err;
        U.dyn = reinterpret_cast<const ffm::UnifiedDyn*>(s->d_dyn);
        if (draws) { U.move_draws = draws->move; U.conflict_draws = draws->conflict; U.draw_steps = draws->steps; U.draw_first = draws->first_step; }
        if (out && out->traj_cells) {
            if (!out->traj_n) return fail(FFM_E_INVALID, "traj_cells without traj_n");
            U.traj = out->traj_cells; U.traj_n = out->traj_n; U.traj_steps = out->traj_steps;
        }
        if (out) {
            U.rec_state = out->rec_state; U.rec_action = out->rec_action; U.rec_reward = out->rec_reward; U.rec_len = out->rec_len;
            U.traj_steps = out->traj_steps;
        }
        void* uargs[] = {&U};
        CU(cudaLaunchKernel(s->kernel, dim3(s->cfg.n_episodes), dim3(s->threads), uargs, (size_t)s->smem_bytes, st));
        s->launches++;
        return FFM_OK;
    }
    if (s->cell_kernel) {
        ffm::CellParams C;
        memset(&C, 0, sizeof(C));
        C.H = s->cfg.height; C.W = s->cfg.width; C.HW = s->HW; C.n_max = s->cfg.n_max; C.B = s->cfg.n_episodes;
        C.max_steps = max_steps;
        C.RW = s->RW; C.RB = s->RB; C.wall_in_smem = s->wall_in_smem ? 1 : 0;
        C.L = ffm::make_cell_layout(s->RB, C.W, s->RW, C.n_max, s->cfg.sff_dtype == FFM_F64 ? 8 : 4, s->HW <= 65536 ? 2 : 4, s->d_dff != nullptr, s->fields_in_smem, s->wall_in_smem, s->score_in_smem);
        C.score_in_smem = s->score_in_smem ? 1 : 0;
        C.magic_w = (uint32_t)(((1ULL << 32) + (uint64_t)C.W - 1) / (uint64_t)C.W);
        const uint64_t cpr = (uint64_t)(s->RW - 2);      // chunks (bitboard words) per row; the kernel skips the division when it is 1
        C.magic_cpr = cpr > 1 ? (uint32_t)(((1ULL << 32) + cpr - 1) / cpr) : 0u;
        C.type_grid = s->d_type_grid; C.wall_bits = s->d_wall_bits; C.score = s->d_score;
        C.kd = (float)s->cfg.k_D; C.c0 = s->cfg.dff_c0; C.c1 = s->cfg.dff_c1; C.thr = s->cfg.dff_threshold;
        C.pos = s->d_pos; C.n_alive = s->d_n; C.t_done = s->d_t; C.ped_steps = s->d_ped_steps;
        C.dff = s->d_dff; C.dff_tmp = s->d_dff_tmp;
        C.seed = s->cfg.seed; C.episode_base = s->cfg.episode_base; C.err = s->d_err;
        if (draws) { C.move_draws = draws->move; C.conflict_draws = draws->conflict; C.draw_steps = draws->steps; C.draw_first = draws->first_step; }
        if (out && out->traj_cells) {
            if (!out->traj_n) return fail(FFM_E_INVALID, "traj_cells without traj_n");
            C.traj = out->traj_cells; C.traj_n = out->traj_n; C.traj_steps = out->traj_steps;
        }
        if (out && out->ctraj) {
            if (!out->traj_n || !out->ctraj_off) return fail(FFM_E_INVALID, "ctraj needs ctraj_off and traj_n");
            if (out->ctraj_cap < 4 || (out->ctraj_cap & 3) || out->ctraj_cap > 0x7FFFFFF0LL)
                return fail(FFM_E_INVALID, "ctraj_cap must be a multiple of 4 in [4, 2^31)");
            if (out->traj_steps < 1) return fail(FFM_E_INVALID, "traj_steps < 1");
            C.ctraj = reinterpret_cast<uint32_t*>(out->ctraj); C.ctraj_off = out->ctraj_off; C.ctraj_cap = out->ctraj_cap;
            C.traj_n = out->traj_n; C.traj_steps = out->traj_steps;
        }
#ifdef FFM_PHASE_TIMING
        static unsigned long long* d_dbg = nullptr;
        if (!d_dbg) { cudaMalloc((void**)&d_dbg, 64); cudaMemset(d_dbg, 0, 64); }
        C.dbg = d_dbg;
#endif
        void* cargs[] = {&C};
        if (s->cluster == 1) {
            CU(cudaLaunchKernel(s->kernel, dim3(s->cfg.n_episodes), dim3(s->threads), cargs, (size_t)s->smem_bytes, st));
        } else {
            cudaLaunchConfig_t lc;
            memset(&lc, 0, sizeof(lc));
            lc.gridDim = dim3((unsigned)s->cfg.n_episodes * (unsigned)s->cluster);
            lc.blockDim = dim3(s->threads);
            lc.dynamicSmemBytes = (size_t)s->smem_bytes;
            lc.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = (unsigned)s->cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            lc.attrs = at; lc.numAttrs = 1;
            CU(cudaLaunchKernelExC(&lc, s->kernel, cargs));
        }
        s->launches++;
#ifdef FFM_PHASE_TIMING
        {
            unsigned long long h[8];
            cudaStreamSynchronize(st);
            cudaMemcpy(h, d_dbg, 64, cudaMemcpyDeviceToHost);
            cudaMemset(d_dbg, 0, 64);
            double tot = 0; for (int i = 0; i < 7; ++i) tot += (double)h[i];
            fprintf(stderr, "[phase cycles, share over warps] p1 %.3f dff %.3f wait1 %.3f p2 %.3f wait2 %.3f p3 %.3f wait3 %.3f  (total %.3e warp-cycles)\n",
                    h[0] / tot, h[1] / tot, h[2] / tot, h[3] / tot, h[4] / tot, h[5] / tot, h[6] / tot, tot);
        }
#endif
        return FFM_OK;
    }
    ffm::RolloutParams P;
    memset(&P, 0, sizeof(P));
    P.H = s->cfg.height; P.W = s->cfg.width; P.HW = s->HW; P.n_max = s->cfg.n_max; P.B = s->cfg.n_episodes;
    P.max_steps = max_steps;
    P.type_grid = s->d_type_grid;
    P.score = s->d_score;
    P.kd = (float)s->cfg.k_D;
    P.c0 = s->cfg.dff_c0; P.c1 = s->cfg.dff_c1; P.thr = s->cfg.dff_threshold;
    P.pos = s->d_pos; P.n_alive = s->d_n; P.t_done = s->d_t; P.ped_steps = s->d_ped_steps;
    P.dff = s->d_dff; P.dff_tmp = s->d_dff_tmp;
    P.seed = s->cfg.seed; P.episode_base = s->cfg.episode_base; P.err = s->d_err;
    if (draws) {
        P.move_draws = draws->move; P.conflict_draws = draws->conflict;
        P.draw_steps = draws->steps; P.draw_first = draws->first_step;
    }
    if (out && out->traj_cells) {
        if (!out->traj_n) return fail(FFM_E_INVALID, "traj_cells without traj_n");
        P.traj = out->traj_cells; P.traj_n = out->traj_n; P.traj_steps = out->traj_steps;
    }
    void* args[] = {&P};
    CU(cudaLaunchKernel(s->kernel, dim3(s->cfg.n_episodes), dim3(s->threads), args, (size_t)s->smem_bytes, st));
    s->launches++;
    return FFM_OK;
}

int ffm_move_probs(ffm_sim_t s, double* probs, int32_t* kind, int space, void* stream) {
    if (!s || !probs || !kind) return fail(FFM_E_INVALID, "null argument");
    if (s->cfg.model != FFM_MODEL_CORE) return fail(FFM_E_UNSUPPORTED, "ffm_move_probs covers the base model");
    if (!s->have_fields || !s->have_positions) return fail(FFM_E_STATE, "fields and positions must be set first");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(s->cfg.device));
    const int B = s->cfg.n_episodes, N = s->cfg.n_max, A = s->cfg.neighborhood + 1;
    double* d_probs = probs; int32_t* d_kind = kind;
    if (space == FFM_HOST) {
        CU(cudaMalloc((void**)&d_probs, (size_t)B * N * A * 8));
        CU(cudaMalloc((void**)&d_kind, (size_t)B * N * 4));
    }
    ffm::RolloutParams P;
    memset(&P, 0, sizeof(P));
    P.H = s->cfg.height; P.W = s->cfg.width; P.HW = s->HW; P.n_max = N; P.B = B;
    P.type_grid = s->d_type_grid; P.score = s->d_score; P.kd = (float)s->cfg.k_D;
    P.pos = s->d_pos; P.n_alive = s->d_n; P.dff = s->d_dff;
    const bool f64 = s->cfg.sff_dtype == FFM_F64, dff = s->d_dff != nullptr;
    const void* k = ffm::pick_probs_kernel(f64, s->cfg.neighborhood, dff);
    const size_t smem = (size_t)(s->HW + 2 * (s->cfg.width + 1)) * 2 + 16;
    CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void* args[] = {&P, &d_probs, &d_kind};
    CU(cudaLaunchKernel(k, dim3(B), dim3(256), args, smem, st));
    s->launches++;
    int rc = FFM_OK;
    if (space == FFM_HOST) {
        rc = copy_out(probs, d_probs, (size_t)B * N * A * 8, FFM_HOST, st);
        if (!rc) rc = copy_out(kind, d_kind, (size_t)B * N * 4, FFM_HOST, st);
        cudaFree(d_probs); cudaFree(d_kind);
    }
    return rc;
}

int ffm_get_counters(ffm_sim_t s, int32_t* steps, int64_t* ped_steps, int space, void* stream) {
    if (!s) return fail(FFM_E_INVALID, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(s->cfg.device));
    int rc;
    if (steps && (rc = copy_out(steps, s->d_t, (size_t)s->cfg.n_episodes * 4, space, st))) return rc;
    if (ped_steps && (rc = copy_out(ped_steps, s->d_ped_steps, (size_t)s->cfg.n_episodes * 8, space, st))) return rc;
    if (space == FFM_HOST && (rc = check_device_flag(s, st))) return rc;
    return FFM_OK;
}

int ffm_tables_shape(ffm_sim_t s, int32_t* n_states, int32_t* n_actions) {
    if (!s) return fail(FFM_E_INVALID, "null argument");
    if (s->cfg.model == FFM_MODEL_CORE) return fail(FFM_E_STATE, "the base model has no tables");
    if (n_states) *n_states = s->S;
    if (n_actions) *n_actions = s->A;
    return FFM_OK;
}

int ffm_tables_set(ffm_sim_t s, const double* V, const uint8_t* v_seen, const double* H, const uint8_t* h_seen, int space, void* stream) {
    if (!s) return fail(FFM_E_INVALID, "null argument");
    if (s->cfg.model == FFM_MODEL_CORE) return fail(FFM_E_STATE, "the base model has no tables");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(s->cfg.device));
    int rc;
    if (V && (rc = copy_in(s->d_V, V, (size_t)s->S * 8, space, st))) return rc;
    if (v_seen && (rc = copy_in(s->d_vseen, v_seen, (size_t)s->S, space, st))) return rc;
    if (H && (rc = copy_in(s->d_H, H, (size_t)s->S * s->A * 8, space, st))) return rc;
    if (h_seen && (rc = copy_in(s->d_hseen, h_seen, (size_t)s->S, space, st))) return rc;
    if (H || h_seen) s->hstats_stale = true;   // extremes of the H table are rescanned by the next ffm_rollout
    if (space == FFM_HOST) CU(cudaStreamSynchronize(st));   // host buffers may be reused by the caller; device copies stay stream-ordered
    return FFM_OK;
}

int ffm_tables_get(ffm_sim_t s, double* V, uint8_t* v_seen, double* H, uint8_t* h_seen, int space, void* stream) {
    if (!s) return fail(FFM_E_INVALID, "null argument");
    if (s->cfg.model == FFM_MODEL_CORE) return fail(FFM_E_STATE, "the base model has no tables");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(s->cfg.device));
    int rc;
    if (V && (rc = copy_out(V, s->d_V, (size_t)s->S * 8, space, st))) return rc;
    if (v_seen && (rc = copy_out(v_seen, s->d_vseen, (size_t)s->S, space, st))) return rc;
    if (H && (rc = copy_out(H, s->d_H, (size_t)s->S * s->A * 8, space, st))) return rc;
    if (h_seen && (rc = copy_out(h_seen, s->d_hseen, (size_t)s->S, space, st))) return rc;
    return FFM_OK;
}

int ffm_tables_bind_deltas(ffm_sim_t s, double* dV, double* dN, double* dF, double* dH) {
    if (!s || !dV || !dN || !dF) return fail(FFM_E_INVALID, "null argument");
    if (s->cfg.model == FFM_MODEL_CORE || s->cfg.model == FFM_MODEL_MCQ) return fail(FFM_E_STATE, "only the unified models take V/H delta tables");
    s->d_dV = dV;
    s->d_dN = dN;
    s->d_dF = dF;
    s->d_dH = dH;
    return FFM_OK;
}

int ffm_tables_apply_deltas(ffm_sim_t s, void* stream) {
    if (!s) return fail(FFM_E_INVALID, "null argument");
    if (!s->d_dV) return fail(FFM_E_STATE, "no delta tables bound");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(s->cfg.device));
    const bool has_h = s->cfg.model == FFM_MODEL_UNIFIED_ACTOR || s->cfg.model == FFM_MODEL_UNIFIED_BOTH;
    if (has_h && !s->d_dH) return fail(FFM_E_STATE, "actor learning needs a dH table");
    CU(ffm::launch_apply_deltas(s->d_V, s->d_dV, s->d_dN, s->d_dF, s->cfg.alpha_v, has_h ? s->d_H : nullptr, s->d_dH, s->d_hseen, s->d_vseen,
                                s->S, s->A, s->d_hstats, s->d_blk_lo, s->d_blk_hi, s->d_blk_any, 148, st));
    s->launches += has_h ? 2 : 1;
    if (has_h) s->hstats_stale = false;
    return FFM_OK;
}

int ffm_bind_dynamic(ffm_sim_t s, const void* dyn) {
    if (!s) return fail(FFM_E_INVALID, "null argument");
    if (s->cfg.model == FFM_MODEL_CORE || s->cfg.model == FFM_MODEL_MCQ) return fail(FFM_E_STATE, "only the unified models take device-resident round parameters");
    s->d_dyn = dyn;
    return FFM_OK;
}

int ffm_set_epsilon(ffm_sim_t s, double epsilon) {
    if (!s) return fail(FFM_E_INVALID, "null argument");
    s->epsilon = epsilon < 0.0 ? 0.0 : (epsilon > 1.0 ? 1.0 : epsilon);   // np.clip (ffm_unified.py:867)
    return FFM_OK;
}

int ffm_q_shape(ffm_sim_t s, int64_t* capacity) {
    if (!s || !capacity) return fail(FFM_E_INVALID, "null argument");
    if (s->cfg.model != FFM_MODEL_MCQ) return fail(FFM_E_STATE, "only the MC-Q model has a Q table");
    *capacity = (int64_t)s->q_cap;
    return FFM_OK;
}

int ffm_q_get(ffm_sim_t s, uint64_t* keys, float* rows, int space, void* stream) {
    if (!s) return fail(FFM_E_INVALID, "null argument");
    if (s->cfg.model != FFM_MODEL_MCQ) return fail(FFM_E_STATE, "only the MC-Q model has a Q table");
    CU(cudaSetDevice(s->cfg.device));
    int rc;
    if (space == FFM_HOST && (rc = check_device_flag(s, (cudaStream_t)stream))) return rc;
    if (keys && (rc = copy_out(keys, s->d_qkeys, (size_t)s->q_cap * 8, space, (cudaStream_t)stream))) return rc;
    if (rows && (rc = copy_out(rows, s->d_Q, (size_t)s->q_cap * 20, space, (cudaStream_t)stream))) return rc;
    return FFM_OK;
}

int ffm_q_set(ffm_sim_t s, const uint64_t* keys, const float* rows, int64_t n, int space, void* stream) {
    if (!s || (n > 0 && (!keys || !rows))) return fail(FFM_E_INVALID, "null argument");
    if (s->cfg.model != FFM_MODEL_MCQ) return fail(FFM_E_STATE, "only the MC-Q model has a Q table");
    if (n < 0 || n > (int64_t)(s->q_cap / 2)) return fail(FFM_E_UNSUPPORTED, "%lld rows do not fit a table of %u slots at load factor 1/2: raise q_log2_capacity", (long long)n, s->q_cap);
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(s->cfg.device));
    CU(cudaMemsetAsync(s->d_qkeys, 0xFF, (size_t)s->q_cap * 8, st));
    CU(cudaMemsetAsync(s->d_Q, 0, (size_t)s->q_cap * 20, st));
    CU(cudaMemsetAsync(s->d_qcount, 0, 4, st));
    if (n > 0) {
        const unsigned long long* dk = reinterpret_cast<const unsigned long long*>(keys); const float* dr = rows;
        unsigned long long* tk = nullptr; float* tr = nullptr;
        if (space == FFM_HOST) {
            CU(cudaMallocAsync((void**)&tk, (size_t)n * 8, st));
            CU(cudaMallocAsync((void**)&tr, (size_t)n * 20, st));
            CU(cudaMemcpyAsync(tk, keys, (size_t)n * 8, cudaMemcpyHostToDevice, st));
            CU(cudaMemcpyAsync(tr, rows, (size_t)n * 20, cudaMemcpyHostToDevice, st));
            dk = tk; dr = tr;
        }
        ffm::McqParams M;
        fill_mcq_params(s, M, 0);
        ffm::mcq_insert_rows_kernel<<<(int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184), 256, 0, st>>>(M, dk, dr, (long long)n);
        CU(cudaGetLastError());
        s->launches++;
        if (space == FFM_HOST) { CU(cudaFreeAsync(tk, st)); CU(cudaFreeAsync(tr, st)); CU(cudaStreamSynchronize(st)); }
    }
    return FFM_OK;
}

int ffm_mcq_set_forced(ffm_sim_t s, const int32_t* target_cell, const int32_t* from_dir, const int32_t* step_cap, void* stream) {
    if (!s || !target_cell || !from_dir || !step_cap) return fail(FFM_E_INVALID, "null argument");
    if (s->cfg.model != FFM_MODEL_MCQ) return fail(FFM_E_STATE, "only the MC-Q model takes teacher-forced first transitions");
    if (!s->have_positions) return fail(FFM_E_STATE, "ffm_set_positions (the source cells) must precede ffm_mcq_set_forced");
    const int B = s->cfg.n_episodes;
    for (int e = 0; e < B; ++e) {
        if (target_cell[e] >= s->HW) return fail(FFM_E_INVALID, "forced target outside the map");
        if (target_cell[e] >= 0 && (from_dir[e] < 0 || from_dir[e] > 4)) return fail(FFM_E_INVALID, "from_dir must be one of FROM_UP..FROM_SELF (0..4)");
        if (step_cap[e] < 0) return fail(FFM_E_INVALID, "negative step cap");
    }
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(s->cfg.device));
    CU(cudaMemcpyAsync(s->d_forced, target_cell, (size_t)B * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(s->d_forced + B, from_dir, (size_t)B * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(s->d_forced + 2 * B, step_cap, (size_t)B * 4, cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));
    s->have_forced = true;
    return FFM_OK;
}

int ffm_mcq_backup_ordered(ffm_sim_t s, void* stream) {
    if (!s) return fail(FFM_E_INVALID, "null argument");
    if (s->cfg.model != FFM_MODEL_MCQ || s->cfg.learn != FFM_LEARN_BATCHED) return fail(FFM_E_STATE, "needs an MC-Q handle created with FFM_LEARN_BATCHED");
    CU(cudaSetDevice(s->cfg.device));
    ffm::McqParams M;
    fill_mcq_params(s, M, 0);
    ffm::mcq_backup_ordered_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(M);
    CU(cudaGetLastError());
    s->launches++;
    return FFM_OK;
}

static int mcq_ensure_deltas(ffm_sim_t s, cudaStream_t st) {
    if (s->d_qG) return FFM_OK;
    const size_t bytes = (size_t)s->q_cap * 5 * 8;
    CU(cudaMalloc((void**)&s->d_qG, bytes));
    CU(cudaMalloc((void**)&s->d_qN, bytes));
    CU(cudaMemsetAsync(s->d_qG, 0, bytes, st));
    CU(cudaMemsetAsync(s->d_qN, 0, bytes, st));
    return FFM_OK;
}

int ffm_mcq_accumulate(ffm_sim_t s, void* stream) {
    if (!s) return fail(FFM_E_INVALID, "null argument");
    if (s->cfg.model != FFM_MODEL_MCQ || s->cfg.learn != FFM_LEARN_BATCHED) return fail(FFM_E_STATE, "needs an MC-Q handle created with FFM_LEARN_BATCHED");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(s->cfg.device));
    int rc;
    if ((rc = mcq_ensure_deltas(s, st))) return rc;
    ffm::McqParams M;
    fill_mcq_params(s, M, 0);
    const long long total = (long long)s->cfg.n_episodes * s->cfg.n_max;
    ffm::mcq_accumulate_kernel<<<(int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184), 256, 0, st>>>(M, s->d_qG, s->d_qN);
    CU(cudaGetLastError());
    s->launches++;
    return FFM_OK;
}

int ffm_mcq_fold(ffm_sim_t s, void* stream) {
    if (!s) return fail(FFM_E_INVALID, "null argument");
    if (s->cfg.model != FFM_MODEL_MCQ || !s->d_qG) return fail(FFM_E_STATE, "no accumulated returns (ffm_mcq_accumulate / ffm_mcq_import_deltas first)");
    CU(cudaSetDevice(s->cfg.device));
    ffm::mcq_fold_kernel<<<1184, 256, 0, (cudaStream_t)stream>>>(s->d_Q, s->d_qG, s->d_qN, (size_t)s->q_cap * 5, s->cfg.alpha_v);
    CU(cudaGetLastError());
    s->launches++;
    return FFM_OK;
}

int ffm_mcq_export_deltas(ffm_sim_t s, uint64_t* keys, double* rows, int64_t capacity, uint32_t* count, void* stream) {
    if (!s || !keys || !rows || !count) return fail(FFM_E_INVALID, "null argument");
    if (s->cfg.model != FFM_MODEL_MCQ || !s->d_qG) return fail(FFM_E_STATE, "no accumulated returns (ffm_mcq_accumulate first)");
    if (capacity < 1 || capacity > 0x7fffffffLL) return fail(FFM_E_INVALID, "bad capacity");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(s->cfg.device));
    CU(cudaMemsetAsync(count, 0, 4, st));
    ffm::mcq_export_deltas_kernel<<<1184, 256, 0, st>>>(s->d_qkeys, s->d_qG, s->d_qN, s->q_cap, reinterpret_cast<unsigned long long*>(keys), rows, count, (unsigned int)capacity, s->d_err);
    CU(cudaGetLastError());
    s->launches++;
    return FFM_OK;
}

int ffm_mcq_import_deltas(ffm_sim_t s, const uint64_t* keys, const double* rows, uint32_t count, const uint32_t* count_dev, void* stream) {
    if (!s || (count > 0 && (!keys || !rows))) return fail(FFM_E_INVALID, "null argument");
    if (s->cfg.model != FFM_MODEL_MCQ) return fail(FFM_E_STATE, "only the MC-Q model exchanges return sums");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(s->cfg.device));
    int rc;
    if ((rc = mcq_ensure_deltas(s, st))) return rc;
    if (count == 0) return FFM_OK;
    ffm::McqParams M;
    fill_mcq_params(s, M, 0);
    ffm::mcq_import_deltas_kernel<<<(int)((count + 255u) / 256u < 1184u ? (count + 255u) / 256u : 1184u), 256, 0, st>>>(M, reinterpret_cast<const unsigned long long*>(keys), rows, count, count_dev, s->d_qG, s->d_qN);
    CU(cudaGetLastError());
    s->launches++;
    return FFM_OK;
}

int ffm_mcq_finalize_timeouts(ffm_sim_t s, void* stream) {
    if (!s) return fail(FFM_E_INVALID, "null argument");
    if (s->cfg.model != FFM_MODEL_MCQ) return fail(FFM_E_STATE, "only the MC-Q model finalizes timeouts");
    return ffm_rollout(s, -1, nullptr, nullptr, stream);    // zero steps, then the timeout records + backups
}

int ffm_set_beta(ffm_sim_t s, double beta) {
    if (!s) return fail(FFM_E_INVALID, "null argument");
    s->beta = beta;
    return FFM_OK;
}

int ffm_set_episode_base(ffm_sim_t s, uint32_t episode_base) {
    if (!s) return fail(FFM_E_INVALID, "null argument");
    s->cfg.episode_base = episode_base;
    return FFM_OK;
}

int ffm_sff_generate(const uint8_t* maps, int32_t n_maps, int32_t H, int32_t W, int32_t mode, int32_t out_dtype, void* out,
                     int space, int32_t device, void* stream, int32_t* rounds_out) {
    if (!maps || !out) return fail(FFM_E_INVALID, "null argument");
    if (n_maps < 1 || H < 1 || W < 1) return fail(FFM_E_INVALID, "bad shape");
    if (mode < FFM_SFF_L1 || mode > FFM_SFF_DIJKSTRA8) return fail(FFM_E_INVALID, "unknown SFF mode %d", mode);
    if (out_dtype != FFM_F32 && out_dtype != FFM_F64) return fail(FFM_E_INVALID, "out_dtype must be FFM_F32 or FFM_F64");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(device));
    {   // scratch comes from the stream-ordered pool; keep freed blocks cached so repeated calls do not hit the OS allocator
        static bool pool_ready[64] = {};
        if (device < 64 && !pool_ready[device]) {
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
                unsigned long long keep = ~0ULL;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            pool_ready[device] = true;
        }
    }
    const size_t HW = (size_t)H * W, total = HW * n_maps;
    const size_t osz = out_dtype == FFM_F64 ? 8 : 4;
    uint8_t* d_maps = nullptr; void* d_out = nullptr; float* d_dist = nullptr;
    int32_t* d_exits = nullptr; int32_t* d_counts = nullptr; int* d_queue = nullptr;
    int rc = FFM_OK, rounds = 0;
#define SFF_CU(call)                                                                        \
    do {                                                                                    \
        cudaError_t e_ = (call);                                                            \
        if (e_ != cudaSuccess) { rc = fail(FFM_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); goto done; } \
    } while (0)
    {
        const uint8_t* mp = maps;
        if (space == FFM_HOST) {
            SFF_CU(cudaMallocAsync((void**)&d_maps, total, st));
            SFF_CU(cudaMemcpyAsync(d_maps, maps, total, cudaMemcpyHostToDevice, st));
            mp = d_maps;
            SFF_CU(cudaMallocAsync(&d_out, total * osz, st));
        } else {
            d_out = out;
        }
        const int bx = (int)((HW + 255) / 256 < 148 * 8 ? (HW + 255) / 256 : 148 * 8);
        if (mode <= FFM_SFF_LINF) {
            SFF_CU(cudaMallocAsync((void**)&d_exits, (size_t)n_maps * ffm::SFF_MAX_EXITS * 2 * sizeof(int32_t), st));
            SFF_CU(cudaMallocAsync((void**)&d_counts, (size_t)n_maps * sizeof(int32_t), st));
            SFF_CU(cudaMemsetAsync(d_counts, 0, (size_t)n_maps * sizeof(int32_t), st));
            ffm::sff_collect_exits_kernel<<<dim3(bx, n_maps), 256, 0, st>>>(mp, d_exits, d_counts, H, W);
            if (out_dtype == FFM_F64)
                ffm::sff_norm_min_kernel<double><<<dim3(bx, n_maps), 256, 0, st>>>(mp, d_exits, d_counts, H, W, mode, (double*)d_out);
            else
                ffm::sff_norm_min_kernel<float><<<dim3(bx, n_maps), 256, 0, st>>>(mp, d_exits, d_counts, H, W, mode, (float*)d_out);
            SFF_CU(cudaGetLastError());
            std::vector<int32_t> counts(n_maps);
            SFF_CU(cudaMemcpyAsync(counts.data(), d_counts, (size_t)n_maps * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
            SFF_CU(cudaStreamSynchronize(st));
            for (int i = 0; i < n_maps; ++i)
                if (counts[i] > ffm::SFF_MAX_EXITS) { rc = fail(FFM_E_UNSUPPORTED, "map %d has %d exit cells (max %d)", i, counts[i], ffm::SFF_MAX_EXITS); goto done; }
        } else {
            const int tiles_x = (W + ffm::SFF_TILE - 1) / ffm::SFF_TILE, tiles_y = (H + ffm::SFF_TILE - 1) / ffm::SFF_TILE;
            const size_t ntiles = (size_t)tiles_x * tiles_y * n_maps;
            if (tiles_y > 65535 || n_maps > 65535) { rc = fail(FFM_E_UNSUPPORTED, "too many tiles / maps for one launch"); goto done; }
            if (ntiles > 0x3FFFFFFFULL) { rc = fail(FFM_E_UNSUPPORTED, "too many tiles for one call"); goto done; }
            SFF_CU(cudaMallocAsync((void**)&d_dist, total * sizeof(float), st));
            // queue scratch in ONE allocation: ring [2 * tiles] | flags [tiles] | 4 counters
            const size_t qwords = 3 * ntiles + 32 + ffm::SFF_Q_WORDS;   // ring [2 * tiles] | flags [tiles] | pad to a 128-byte line | counters
            SFF_CU(cudaMallocAsync((void**)&d_queue, qwords * sizeof(int), st));
            ffm::SffQueue q;
            q.ring = d_queue; q.flag = d_queue + 2 * ntiles; q.ctrl = reinterpret_cast<unsigned int*>(d_queue + ((3 * ntiles + 31) & ~(size_t)31));
            q.cap = (unsigned int)(2 * ntiles);
            SFF_CU(cudaMemsetAsync(q.ring, 0xFF, 2 * ntiles * sizeof(int), st));        // -1 = empty slot
            SFF_CU(cudaMemsetAsync(q.flag, 0, (qwords - 2 * ntiles) * sizeof(int), st));
            ffm::sff_relax_init_kernel<<<dim3(bx, n_maps), 256, 0, st>>>(mp, d_dist, q, H, W, tiles_x, tiles_y);
            const float INF = __builtin_huge_valf();
            const float w_axis = 1.0f;
            const float w_diag = mode == FFM_SFF_BFS4 ? INF : (mode == FFM_SFF_BFS8 ? 1.0f : (float)1.4142135623730951);
            // unit-cost modes: one warp per tile visit (sff_bfs_warp_kernel); Dijkstra-8 (float costs): one CTA per visit.
            // FFM_SFF_KERNEL=tile forces the CTA-per-tile kernel for the unit-cost modes too (A/B measurements).
            const bool warp_kernel = mode != FFM_SFF_DIJKSTRA8 && !(getenv("FFM_SFF_KERNEL") && strcmp(getenv("FFM_SFF_KERNEL"), "tile") == 0);
            const void* kq = warp_kernel ? (mode == FFM_SFF_BFS8 ? (const void*)ffm::sff_bfs_warp_kernel<true> : (const void*)ffm::sff_bfs_warp_kernel<false>)
                                         : (const void*)ffm::sff_relax_queue_kernel;
            int per_sm = 0, sms = 0;
            SFF_CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kq, 256, 0));
            SFF_CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
            if (const char* ev = getenv("FFM_SFF_CTAS_PER_SM")) { const int v = atoi(ev); if (v >= 1 && v < per_sm) per_sm = v; }   // tuning
            const size_t resident = (size_t)per_sm * sms;      // persistent CTAs: spinning consumers must all be resident
            const size_t want = warp_kernel ? (ntiles + 7) / 8 : ntiles;
            const int grid = (int)(want < resident ? want : resident);
            if (!warp_kernel) ffm::sff_relax_queue_kernel<<<grid, 256, 0, st>>>(mp, d_dist, q, H, W, tiles_x, tiles_y, w_axis, w_diag);
            else if (mode == FFM_SFF_BFS8) ffm::sff_bfs_warp_kernel<true><<<grid, 256, 0, st>>>(mp, d_dist, q, H, W, tiles_x, tiles_y);
            else ffm::sff_bfs_warp_kernel<false><<<grid, 256, 0, st>>>(mp, d_dist, q, H, W, tiles_x, tiles_y);
            SFF_CU(cudaGetLastError());
            if (rounds_out) {                                  // tile visits, for the caller that asks (one 4-byte read-back)
                unsigned int visits = 0;
                SFF_CU(cudaMemcpyAsync(&visits, q.ctrl + ffm::SFF_Q_VISITS, sizeof(visits), cudaMemcpyDeviceToHost, st));
                SFF_CU(cudaStreamSynchronize(st));
                rounds = (int)visits;
            }
            const int cb = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
            if (out_dtype == FFM_F64) ffm::sff_convert_kernel<double><<<cb, 256, 0, st>>>(d_dist, (double*)d_out, total);
            else ffm::sff_convert_kernel<float><<<cb, 256, 0, st>>>(d_dist, (float*)d_out, total);
            SFF_CU(cudaGetLastError());
        }
        if (space == FFM_HOST) {     // device-space calls stay stream-ordered: no host synchronisation at all in the geodesic modes
            SFF_CU(cudaMemcpyAsync(out, d_out, total * osz, cudaMemcpyDeviceToHost, st));
            SFF_CU(cudaStreamSynchronize(st));
        }
    }
done:
#undef SFF_CU
    if (space == FFM_HOST) { if (d_maps) cudaFreeAsync(d_maps, st); if (d_out) cudaFreeAsync(d_out, st); }
    if (d_dist) cudaFreeAsync(d_dist, st);
    if (d_exits) cudaFreeAsync(d_exits, st);
    if (d_counts) cudaFreeAsync(d_counts, st);
    if (d_queue) cudaFreeAsync(d_queue, st);
    if (rounds_out) *rounds_out = rounds;
    return rc;
}

int ffm_rollout_returns(const float* reward, const int32_t* len, int32_t B, int32_t T, int32_t N, double gamma, double* returns,
                        int32_t device, void* stream) {
    if (!reward || !len || !returns) return fail(FFM_E_INVALID, "null argument");
    if (B < 1 || T < 1 || N < 1) return fail(FFM_E_INVALID, "bad shape");
    CU(cudaSetDevice(device));
    CU(ffm::launch_rollout_returns(reward, len, B, T, N, gamma, returns, (cudaStream_t)stream));
    return FFM_OK;
}

int64_t ffm_launch_count(ffm_sim_t s) { return s ? s->launches : 0; }

int ffm_kernel_info(ffm_sim_t s, int32_t* smem_bytes, int32_t* threads, int32_t* ctas_per_sm, int32_t* fields_in_smem) {
    if (!s) return fail(FFM_E_INVALID, "null argument");
    if (smem_bytes) *smem_bytes = s->smem_bytes;
    if (threads) *threads = s->threads;
    if (ctas_per_sm) *ctas_per_sm = s->ctas_per_sm;
    if (fields_in_smem) *fields_in_smem = s->fields_in_smem ? 1 : 0;
    return FFM_OK;
}

int ffm_cluster_info(ffm_sim_t s, int32_t* cluster, int32_t* max_clusters, int32_t* score_in_smem, int32_t* cell_kernel) {
    if (!s) return fail(FFM_E_INVALID, "null argument");
    if (cluster) *cluster = s->cluster;
    if (max_clusters) *max_clusters = s->cluster > 1 ? s->max_clusters : 0;
    if (score_in_smem) *score_in_smem = (s->fields_in_smem && s->score_in_smem) ? 1 : 0;
    if (cell_kernel) *cell_kernel = s->cfg.model == FFM_MODEL_CORE ? (s->cell_kernel ? 1 : 0) : -1;
    return FFM_OK;
}

}  // extern "C"
