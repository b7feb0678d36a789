// Persistent rollout kernel of the base floor-field CA, cell-centric formulation: one CTA per episode -- or one
// thread-block CLUSTER per episode, the map split into row bands held in distributed shared memory -- all steps
// in-kernel.
//
// Reproduces, per step, FloorFieldModel.step() + update_dff() of the reference (model/ffm_core.py:36-117) and,
// around it, the loop of run() (ffm_core.py:119-126).
//
// The reference walks the pedestrians one by one and rebuilds each one's candidate list with np.isin (:48-60).
// Here the crowd is a BITBOARD (one bit per cell: blocked = wall or occupied) and a step is three block-wide phases
// separated by one barrier each:
//
//   1 (cells -> movers)   every thread owns 16 cells of a bitboard row: the 8 (4) shifted copies of the "free" rows
//                         give, bit-parallel, which occupied cells have a free neighbour at all (:57-63 -- a
//                         pedestrian with no candidate makes no request and draws nothing) and, through a carry-save
//                         adder over the shifted words, HOW MANY free neighbours each has.  Movers are appended to
//                         three work lists by candidate count (2 / 3 / more candidates incl. "stay"), so that phase 2
//                         runs with uniform trip counts.  Blocked pedestrians (55 % of all pedestrian-steps in a
//                         packed crowd) cost nothing beyond these bit operations.
//                         The same phase sweeps the claim masks, refreshes the alive-rank prefix after exits and runs
//                         the DFF decay/diffusion of the PREVIOUS step (:106-117).
//   2 (movers -> requests) per list entry: candidate mask from the bitboard windows, forced exit (:66-72) or
//                         score = -k_S*sff + k_D*dff, exp(score - max), keyed uniform, CDF search (:74-88) -> target;
//                         the request is one atomicOr of the direction bit into the target cell's CLAIM MASK (a byte
//                         per cell: bit k = "the occupant of neighbour k wants this cell"); the requester that finds
//                         the mask empty becomes the cell's resolver.  "Stay" decisions leave their footprint here.
//   3 (resolve + apply)   one thread per requested cell: a lone claimant moves (:91-93); k >= 2 claimants -> coin,
//                         then the floor(u*k)-th claimant in ascending agent index (:94-98), found from the claim
//                         mask (no neighbourhood scan per claimant).  The move is applied at once: owner grid,
//                         bitboard, DFF footprint, exit removal (:101-102).  Concurrent resolvers never touch the
//                         same cells: a pedestrian claims exactly one cell, and a cell has exactly one resolver.
//
// A pedestrian's identity is a stable id (its index at launch); the reference's array index -- the key of its move
// draw -- is the id's alive rank (prefix popcount over the alive bitmap), because the reference compacts the
// position array stably (:101-102).  There are no per-pedestrian arrays at all: the owner grid maps cell -> id.
//
// Cluster mode (CL > 1): CTA b of the cluster owns rows [b*RB, (b+1)*RB) of every per-cell array.  Phase 1 reads
// the bitboard row above / below its band from the neighbour CTA, moves / claims / DFF reads that cross a band
// boundary go through distributed shared memory (cluster.map_shared_rank), the alive bitmap is replicated and exits
// are broadcast; the three barriers become cluster barriers.  This keeps maps that do not fit one SM's shared memory
// (BASELINE C3: 256x256 with DFF) entirely on chip.
#pragma once
#include <cooperative_groups.h>

#include <type_traits>

#include "ffm_core_kernel.cuh"
#include "ffm_dff_stencil.cuh"

namespace ffm {

namespace cg = cooperative_groups;

struct CellLayout {
    uint32_t score, dffA, dffB, grid, cmask, blk, wall, listA, listB, clist, alive, wpre, ctr, bar, total;
    uint32_t cap;   // capacity of each work list (entries)
};

struct CellParams {
    CellLayout L;                // shared-memory layout of one CTA (make_cell_layout, computed by the host)
    int H, W, HW, n_max, B;
    int max_steps;
    int RW;                      // bitboard words per row: ceil(W/32) + 2 guard words
    int RB;                      // rows per band (H when one CTA holds the whole map)
    int wall_in_smem;            // static wall bitboard staged in shared memory (else read through L1 from wall_bits)
    int score_in_smem;           // score field staged in shared memory (else read through L1/L2: it is read-only and shared by all
                                 // episodes of the handle); only meaningful when the fields are shared-memory resident
    uint32_t magic_w;            // ceil(2^32 / W): row = umulhi(cell, magic_w) for cell * W < 2^32
    uint32_t magic_cpr;          // ceil(2^32 / chunks per row), chunks per row = ceil(W/32) (one bitboard word each)
    const uint16_t* type_grid;   // [HW + 2*(W+1)] type bits only, guard band included
    const uint32_t* wall_bits;   // [(H+2) * RW] 1 = not passable / outside the map
    const void* score;           // [HW] S
    float kd, c0, c1, thr;
    uint32_t* pos;               // [B][n_max] linear cells, alive-rank order
    int32_t* n_alive;            // [B]
    int32_t* t_done;             // [B]
    unsigned long long* ped_steps;  // [B]
    float* dff;                  // [B][HW]
    float* dff_tmp;              // [B][HW]   second buffer when the fields stay in global memory
    unsigned long long seed;
    uint32_t episode_base;
    const double* move_draws;    // [B][draw_steps][n_max] or null
    const double* conflict_draws;  // [B][draw_steps][HW][2] or null
    int draw_steps, draw_first;
    uint32_t* traj;              // [B][traj_steps][n_max] or null
    int32_t* traj_n;             // [B][traj_steps]
    int traj_steps;
    // compact trajectory record (what run() collects, ffm_core.py:125 / main.py:44-52): per episode a stream of
    // (row, col) int16 pairs, the rows of consecutive steps back to back, each padded to a multiple of 4 entries
    int32_t* err;                // device validation flag (128: two pedestrians on one cell)
    unsigned long long* dbg;     // FFM_PHASE_TIMING builds: [8] accumulated cycles (work / wait per phase), else unused
    uint32_t* ctraj;             // [B][ctraj_cap]  low half = row, high half = col (an int16 pair in memory)
    int32_t* ctraj_off;          // [B][traj_steps + 1]: entry offset of each step's row (CSR); [steps] = end
    long long ctraj_cap;
};

__host__ __device__ inline CellLayout make_cell_layout(int RB, int W, int RW, int n_max, int sizeof_score, int sizeof_ent, bool dff,
                                                       bool fields_in_smem, bool wall_in_smem = true, bool score_in_smem = true) {
    const uint32_t cells = (uint32_t)RB * W;
    const uint32_t nw = (uint32_t)(n_max + 31) / 32 + 1;
    CellLayout L;
    L.cap = (uint32_t)n_max < cells ? (uint32_t)n_max : cells;
    uint32_t o = 0;
    L.score = o; if (fields_in_smem && score_in_smem) o = align16(o + cells * sizeof_score);
    L.dffA = o;  if (fields_in_smem && dff) o = align16(o + cells * 4u);
    L.dffB = o;  if (fields_in_smem && dff) o = align16(o + cells * 4u);
    L.grid = o;  o = align16(o + (cells + 2u * (W + 1)) * 2u);
    L.cmask = o; o = align16(o + cells);
    L.blk = o;   o = align16(o + (uint32_t)(RB + 2) * RW * 4u);
    L.wall = o;  if (wall_in_smem) o = align16(o + (uint32_t)(RB + 2) * RW * 4u);
    L.listA = o; o = align16(o + L.cap * sizeof_ent);
    L.listB = o; o = align16(o + L.cap * sizeof_ent);
    L.clist = o;   // (the contested-cell list lives in the larger free part of the two work lists)
    L.alive = o; o = align16(o + nw * 4u);
    L.wpre = o;  o = align16(o + nw * 4u);
    L.ctr = o;   o = align16(o + 16u * 4u);
    L.bar = o;   o = align16(o + 8u);
    L.total = o;
    return L;
}

// row / column offset of a neighbour index only known at run time (the tables of nbr_off_rt)
template <int NBR> __device__ __forceinline__ int nbr_dr_rt(int k) { return (int)(((NBR == 8 ? 0xA940u : 0x58u) >> (2 * k)) & 3u) - 1; }
template <int NBR> __device__ __forceinline__ int nbr_dc_rt(int k) { return (int)(((NBR == 8 ? 0x9224u : 0x85u) >> (2 * k)) & 3u) - 1; }

// Candidate masks.  Moore: bit b = 4*(dr+1) + (dc+1) of the 3x3 window (bits 0,1,2, 4,6, 8,9,10 -- ascending bit order
// is the reference's neighbour order, ffm_core.py:32-34), so that the cell offset is linear in the bit index.
// von Neumann: bit k of the reference's list [(-1,0),(1,0),(0,-1),(0,1)] (:30).
template <int NBR> __device__ __forceinline__ int cand_off(int b, int W) { return NBR == 8 ? (b >> 2) * W + (b & 3) - (W + 1) : nbr_off_rt<4>(b, W); }
template <int NBR> __device__ __forceinline__ int cand_dir(int b) { return NBR == 8 ? 3 * (b >> 2) + (b & 3) - (b > 5 ? 1 : 0) : b; }
template <int NBR> __device__ __forceinline__ constexpr int cand_bit(int k) { return NBR == 8 ? k + (k >= 3 ? 1 : 0) + (k >= 4 ? 1 : 0) + (k >= 5 ? 1 : 0) : k; }

// index of the neighbour opposite to k (the direction from the target back to the claimant)
template <int NBR> __device__ __forceinline__ int nbr_opp(int k) { return NBR == 8 ? 7 - k : (k ^ 1); }

// Compare-exchange network: ascending sort of the NBR keys (invalid entries hold 0xFFFFFFFF and end up last).
template <int NBR>
__device__ __forceinline__ void sort_keys(uint32_t (&v)[NBR]) {
#define FFM_CX(a, b) { const uint32_t lo_ = min(v[a], v[b]), hi_ = max(v[a], v[b]); v[a] = lo_; v[b] = hi_; }
    if (NBR == 4) {
        FFM_CX(0, 1) FFM_CX(2, 3) FFM_CX(0, 2) FFM_CX(1, 3) FFM_CX(1, 2)
    } else {
        FFM_CX(0, 1) FFM_CX(2, 3) FFM_CX(4, 5) FFM_CX(6, 7)
        FFM_CX(0, 2) FFM_CX(1, 3) FFM_CX(4, 6) FFM_CX(5, 7)
        FFM_CX(1, 2) FFM_CX(5, 6) FFM_CX(0, 4) FFM_CX(3, 7)
        FFM_CX(1, 5) FFM_CX(2, 6)
        FFM_CX(1, 4) FFM_CX(3, 6)
        FFM_CX(2, 4) FFM_CX(3, 5)
        FFM_CX(3, 4)
    }
#undef FFM_CX
}

// update_dff (ffm_core.py:106-117) over the rows [r0, r1) of a field whose rows are reached through in_row(r)
// (nullptr outside the map: np.pad of the scaled field, :111) -- same arithmetic as dff_decay_diffuse.
template <int NBR, typename RowIn, typename RowOut>
__device__ __forceinline__ void dff_stencil_rows(RowIn in_row, RowOut out_row, int r0, int r1, int W, float c0, float c1, float thr,
                                                 int tid, const StencilGeom& geom) {
    const int cw = geom.cw, bands = geom.bands, rpb = geom.srpb;
    const int band = geom.band, colb = geom.colb;
    if (band >= bands) return;
    const int a0 = r0 + band * rpb, a1 = min(r1, a0 + rpb);
    if (a0 >= a1) return;
    for (int col = colb; col < W; col += cw) {
        const bool hasl = col > 0, hasr = col + 1 < W;
        float up[3], uc[3], un[3], sc = 0.0f;
        auto load_row = [&](int r, float (&u)[3], float& s_centre) {
            float d0 = 0.0f, d1 = 0.0f, d2 = 0.0f;
            const float* row = in_row(r);
            if (row != nullptr) {
                d1 = row[col];
                if (hasl) d0 = row[col - 1];
                if (hasr) d2 = row[col + 1];
            }
            const float s0 = __fmul_rn(c0, d0), s1 = __fmul_rn(c0, d1), s2 = __fmul_rn(c0, d2);   // (:109)
            u[0] = __fmul_rn(c1, s0); u[1] = __fmul_rn(c1, s1); u[2] = __fmul_rn(c1, s2);         // (:113)
            s_centre = s1;
        };
        float dummy;
        load_row(a0 - 1, up, dummy);
        load_row(a0, uc, sc);
        for (int r = a0; r < a1; ++r) {
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
            out_row(r)[col] = acc;
#pragma unroll
            for (int k = 0; k < 3; ++k) { up[k] = uc[k]; uc[k] = un[k]; }
            sc = sn;
        }
    }
}

// atomicAdd on a CTA-local shared-memory word issued by ONE lane: the plain instruction, without the warp-aggregation
// sequence (vote / popc / shuffle) the compiler wraps around atomicAdd() on a possibly uniform address
__device__ __forceinline__ uint32_t smem_atomic_add(uint32_t* p, uint32_t v) {
    uint32_t old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(p)), "r"(v) : "memory");
    return old;
}

// alive-rank prefix: exclusive popcount prefix over the alive words, by one warp (kept out of line: it runs only in
// steps where somebody left, and the hot loop should stay small for the instruction cache)
static __device__ __noinline__ void refresh_alive_prefix(const uint32_t* alive, uint32_t* wpre, int nwords, int lane) {
    uint32_t carry = 0;
#pragma unroll 1
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


// One step's row of the compact trajectory record (what run() collects per step, ffm_core.py:125): (row, col) int16 pairs
// in alive-rank order, appended to the episode's stream at CSR offset ctr[14]; rows are padded with (-1, -1) to a multiple
// of 4 entries.  single_cta: rank-order scatter into the (idle) work lists in shared memory, then 16-byte coalesced
// streaming stores; a cluster's CTAs scatter their own cells straight to global memory.  Kept out of line so that the
// recording code costs the hot loop no registers (the SFF-only variants run at a 40-register cap).
struct RecArgs {             // what record_step_compact needs of CellParams, passed by value (taking the address of the
    uint32_t* ctraj;         // kernel parameter block would copy all of it to the stack)
    int32_t* ctraj_off;
    int32_t* traj_n;
    const uint32_t* wall_bits;
    long long ctraj_cap;
    uint32_t l_grid, l_blk, l_wall, l_alive, l_wpre, l_ctr, l_listA;
    uint32_t magic_cpr;
    int H, W, RW, RB, traj_steps, wall_in_smem, write_n;
};

static __device__ __noinline__ void record_step_compact(const RecArgs P, unsigned char* smem_raw, int e, int tl, int n, int band,
                                                        bool single_cta) {
    struct { uint32_t grid, blk, wall, alive, wpre, ctr, listA; } L = {P.l_grid, P.l_blk, P.l_wall, P.l_alive, P.l_wpre, P.l_ctr, P.l_listA};
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int W = P.W, RW = P.RW, G = W + 1;
    const int r0 = band * P.RB, r1 = min(P.H, r0 + P.RB), RBl = max(0, r1 - r0), lo = r0 * W;
    const uint16_t* grid_l = reinterpret_cast<const uint16_t*>(smem_raw + L.grid) + G;
    const uint32_t* blk_l = reinterpret_cast<const uint32_t*>(smem_raw + L.blk);
    const uint32_t* wall_l = reinterpret_cast<const uint32_t*>(smem_raw + L.wall);
    const uint32_t* alive = reinterpret_cast<const uint32_t*>(smem_raw + L.alive);
    const uint32_t* wpre = reinterpret_cast<const uint32_t*>(smem_raw + L.wpre);
    uint32_t* ctr = reinterpret_cast<uint32_t*>(smem_raw + L.ctr);
    const int cpr = RW - 2, nchunks = RBl * cpr;
    const uint32_t npad = ((uint32_t)n + 3u) & ~3u;
    int coff = (int)ctr[14];
    const bool fits = coff >= 0 && (long long)coff + npad <= P.ctraj_cap;
    if (fits) {
        uint32_t* dst = P.ctraj + (size_t)e * P.ctraj_cap + coff;
        uint32_t* stage = reinterpret_cast<uint32_t*>(smem_raw + L.listA);          // the work lists are idle until the next phase 1
        const bool staged = single_cta && npad * 4u <= L.alive - L.listA;
        uint32_t* out = staged ? stage : dst;
#pragma unroll 1
        for (int ch = tid; ch < nchunks; ch += nthreads) {
            const int lr = cpr == 1 ? ch : (int)__umulhi((uint32_t)ch, P.magic_cpr), j = ch - lr * cpr;
            const int idx = (lr + 1) * RW + 1 + j;
            const uint32_t wv = P.wall_in_smem ? wall_l[idx] : __ldg(P.wall_bits + (size_t)r0 * RW + idx);
            uint32_t occ = blk_l[idx] & ~wv;
            const int cbase = (r0 + lr) * W + 32 * j;
            while (occ) {
                const int b = __ffs(occ) - 1;
                occ &= occ - 1u;
                const uint32_t id = (grid_l[cbase + b - lo] & OCC_MASK) - 1u;
                const uint32_t rank = wpre[id >> 5] + (uint32_t)__popc(alive[id >> 5] & ((1u << (id & 31u)) - 1u));
                out[rank] = (uint32_t)(r0 + lr) | ((uint32_t)(32 * j + b) << 16);
            }
        }
        if ((staged || band == 0) && (uint32_t)tid < npad - (uint32_t)n) out[n + tid] = 0xFFFFFFFFu;
        if (staged) {
            __syncthreads();
            uint4* d4 = reinterpret_cast<uint4*>(dst);
            const uint4* s4 = reinterpret_cast<const uint4*>(stage);
            for (uint32_t x = tid; x < npad / 4u; x += nthreads) __stcs(d4 + x, s4[x]);
        }
        coff += (int)npad;
    } else {
        coff = -1;
    }
    __syncthreads();                                      // the staging area and ctr[14] have been read by everybody
    if (tid == 0) {
        ctr[14] = (uint32_t)coff;
        if (band == 0) {
            P.ctraj_off[(size_t)e * (P.traj_steps + 1) + tl + 1] = coff;
            if (P.write_n) P.traj_n[(size_t)e * P.traj_steps + tl] = n;
        }
    }
}

#ifndef FFM_CELL_MINB256
#define FFM_CELL_MINB256 6      // resident CTAs per SM the 256-thread float32 variant is compiled for (register cap)
#endif
#ifndef FFM_CELL_MINB256_DFF
#define FFM_CELL_MINB256_DFF 4  // the same with the DFF tracked: the stencil's register window needs 64 registers (no spills)
#endif
template <typename S, typename EntT, int NBR, bool DFF, bool FIELDS_IN_SMEM, int THREADS, int CL>
__global__ void __launch_bounds__(THREADS, (CL == 1 && THREADS <= 256 && sizeof(S) == 4) ? ((DFF ? FFM_CELL_MINB256_DFF : FFM_CELL_MINB256) * 256) / THREADS
                                           : ((CL >= 4 && !FIELDS_IN_SMEM && THREADS == 512) ? 2 : 1))   // bands of >= 4-CTA clusters with the fields in L2: two CTAs per SM
ffm_cell_rollout_kernel(const CellParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NW = THREADS / 32;
    constexpr uint32_t EXIT_EMPTY = TYPE_EXIT << TYPE_SHIFT;
    constexpr uint32_t FULL = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = P.W, HW = P.HW, H = P.H, RW = P.RW, G = W + 1;
    const int WW = RW - 2;

    // ---- band geometry ------------------------------------------------------------------------------
    unsigned band = 0;
    int e = blockIdx.x;
    if (CL > 1) {
        band = cg::this_cluster().block_rank();
        e = blockIdx.x / CL;
    }
    const int RB = P.RB;
    const int r0 = (int)band * RB;                       // first row of this band
    const int r1 = min(H, r0 + RB);                      // one past its last row (an empty band has r1 <= r0)
    const int RBl = max(0, r1 - r0);
    const int lo = r0 * W, hi = lo + RBl * W;            // cells owned: [lo, hi)
    const bool top = (CL == 1) || band == 0, bottom = (CL == 1) || r1 >= H;
    // cells addressable in this CTA's arrays: the owned ones plus the guard bands at the map's top / bottom edge
    const int lo_acc = lo - (top ? G : 0);
    const uint32_t span_acc = (uint32_t)(hi - lo_acc + (bottom ? G : 0));

    const CellLayout& L = P.L;
    uint16_t* grid_l = reinterpret_cast<uint16_t*>(smem_raw + L.grid) + G;     // grid_l[c - lo]
    uint32_t* cmask_l = reinterpret_cast<uint32_t*>(smem_raw + L.cmask);       // byte (c - lo) of this array
    uint32_t* blk_l = reinterpret_cast<uint32_t*>(smem_raw + L.blk);           // row (r - r0 + 1), word (col/32 + 1)
    uint32_t* wall_l = reinterpret_cast<uint32_t*>(smem_raw + L.wall);
    EntT* listA = reinterpret_cast<EntT*>(smem_raw + L.listA);
    EntT* listB = reinterpret_cast<EntT*>(smem_raw + L.listB);
    const int cap1 = (int)L.cap - 1;
    uint32_t* alive = reinterpret_cast<uint32_t*>(smem_raw + L.alive);
    uint32_t* wpre = reinterpret_cast<uint32_t*>(smem_raw + L.wpre);
    uint32_t* ctr = reinterpret_cast<uint32_t*>(smem_raw + L.ctr);             // [parity][n2, n3, nbig, n_exit, n_contested, warps done, ..]
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem_raw + L.bar);
    S* score_l = reinterpret_cast<S*>(smem_raw + L.score);
    float* dffA_l = reinterpret_cast<float*>(smem_raw + L.dffA);
    float* dffB_l = reinterpret_cast<float*>(smem_raw + L.dffB);
    const S* score_g = reinterpret_cast<const S*>(P.score);
    float* dff_home = DFF ? P.dff + (size_t)e * HW : nullptr;
    float* dffA_g = dff_home;
    float* dffB_g = DFF ? P.dff_tmp + (size_t)e * HW : nullptr;

    // ---- accessors: a cell of this band is local shared memory, a cell of a neighbouring band is reached through
    //      distributed shared memory (same offset in the neighbour CTA's window) --------------------------------
    auto is_local = [&](int c) -> bool { return CL == 1 || (uint32_t)(c - lo_acc) < span_acc; };
    auto nb_rank = [&](int c) -> unsigned { return c < lo ? band - 1 : band + 1; };
    auto grid_ptr = [&](int c) -> uint16_t* {
        if (is_local(c)) return grid_l + (c - lo);
        const unsigned rk = nb_rank(c);
        return cg::this_cluster().map_shared_rank(grid_l, rk) + (c - (int)rk * RB * W);
    };
    auto cmask_word = [&](int c, uint32_t& shift) -> uint32_t* {          // the 32-bit word holding cell c's claim byte
        uint32_t* base = cmask_l;
        int lc = c - lo;
        if (!is_local(c)) {
            const unsigned rk = nb_rank(c);
            base = cg::this_cluster().map_shared_rank(cmask_l, rk);
            lc = c - (int)rk * RB * W;
        }
        shift = 8u * ((uint32_t)lc & 3u);
        return base + (lc >> 2);
    };
    auto blk_word = [&](int r, int col, uint32_t& bit) -> uint32_t* {     // bitboard word of cell (r, col)
        bit = 1u << (col & 31);
        if (CL == 1 || (r >= r0 && r < r1)) return blk_l + (r - r0 + 1) * RW + 1 + (col >> 5);
        const unsigned rk = r < r0 ? band - 1 : band + 1;
        return cg::this_cluster().map_shared_rank(blk_l, rk) + (r - (int)rk * RB + 1) * RW + 1 + (col >> 5);
    };
    int dpar = 0;   // which physical DFF buffer currently holds the field ("A"); flips once per step in every CTA alike
    auto dff_cur = [&](int c) -> float* {                                 // current DFF value of cell c
        if (!FIELDS_IN_SMEM) return (dpar ? dffB_g : dffA_g) + c;
        float* base = dpar ? dffB_l : dffA_l;
        if (is_local(c) && (CL == 1 || (c >= lo && c < hi))) return base + (c - lo);
        const unsigned rk = nb_rank(c);
        return cg::this_cluster().map_shared_rank(base, rk) + (c - (int)rk * RB * W);
    };
    const bool score_smem = FIELDS_IN_SMEM && P.score_in_smem;
    auto score_at = [&](int c) -> S {
        if (!FIELDS_IN_SMEM) return score_g[c];
        if (!score_smem) return __ldg(score_g + c);
        if (CL == 1 || (c >= lo && c < hi)) return score_l[c - lo];
        const unsigned rk = nb_rank(c);
        return cg::this_cluster().map_shared_rank(score_l, rk)[c - (int)rk * RB * W];
    };
    auto sync_all = [&]() {
        if (CL == 1) __syncthreads(); else cg::this_cluster().sync();
    };

    // ---- prologue: stage the band's fields (TMA bulk copies + mbarrier when 16-byte aligned) -----------------
    const uint32_t cells = (uint32_t)RBl * W;
    const uint32_t grid_elems = cells + 2u * G;
    const uint32_t grid_bytes = align16(grid_elems * 2u), score_bytes = cells * (uint32_t)sizeof(S), dff_bytes = cells * 4u;
    // the band slices start at element lo of their arrays (type_grid: index = cell + G, and local index 0 <-> cell lo - G)
    const bool tma_grid = RBl > 0 && (((size_t)lo * 2u) % 16u == 0u);   // d_type_grid is padded to a multiple of 16 bytes
    const bool tma_fields = FIELDS_IN_SMEM && RBl > 0 && (score_bytes % 16u == 0u || !score_smem) && (dff_bytes % 16u == 0u) &&
                            (((size_t)lo * sizeof(S)) % 16u == 0u) && (((size_t)lo * 4u) % 16u == 0u) && (((size_t)HW * 4u) % 16u == 0u);
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) {
        uint32_t bytes = 0;
        if (tma_grid) bytes += grid_bytes;
        if (tma_fields) bytes += (score_smem ? score_bytes : 0u) + (DFF ? dff_bytes : 0u);
        mbar_arrive_expect_tx(bar, bytes);
        if (tma_grid) bulk_copy_g2s(smem_raw + L.grid, P.type_grid + lo, grid_bytes, bar);
        if (tma_fields) {
            if (score_smem) bulk_copy_g2s(score_l, score_g + lo, score_bytes, bar);
            if (DFF) bulk_copy_g2s(dffA_l, dff_home + lo, dff_bytes, bar);
        }
    }
    if (!tma_grid)
#pragma unroll 1
        for (uint32_t x = tid; x < grid_elems; x += THREADS) grid_l[(int)x - G] = P.type_grid[lo + x];
    if (FIELDS_IN_SMEM && !tma_fields) {
#pragma unroll 1
        for (uint32_t x = tid; x < (score_smem ? cells : 0u); x += THREADS) score_l[x] = score_g[lo + x];
        if (DFF)
#pragma unroll 1
            for (uint32_t x = tid; x < cells; x += THREADS) dffA_l[x] = dff_home[lo + x];
    }
    // bitboards: rows r0-1 .. r1 of the static wall board (the halo rows matter only at the map's edge)
#pragma unroll 1
    for (int x = tid; x < (RB + 2) * RW; x += THREADS) {
        const int lr = x / RW;
        const int gr = r0 + lr;                           // row index in the global board (which has its own +1 offset)
        const uint32_t wv = (gr <= H + 1 && lr <= RBl + 1) ? P.wall_bits[(size_t)gr * RW + (x - lr * RW)] : FULL;
        if (P.wall_in_smem) wall_l[x] = wv;
        blk_l[x] = wv;
    }
    int n = P.n_alive[e];                                 // pedestrians still inside (whole episode)
    const int n_ids = n;                                  // ids 0 .. n_ids-1 (= index at launch)
    const int nwords = (n_ids + 31) >> 5;
    const int t0 = P.t_done[e];
    uint32_t* gpos = P.pos + (size_t)e * P.n_max;
#pragma unroll 1
    for (int w = tid; w <= nwords; w += THREADS) {
        const int b0 = w * 32;
        alive[w] = (b0 + 32 <= n) ? FULL : (b0 < n ? ((1u << (n - b0)) - 1u) : 0u);
        wpre[w] = (uint32_t)(b0 < n ? b0 : n);
    }
    if (tid < 16) ctr[tid] = 0u;
#pragma unroll 1
    for (uint32_t x = tid; x < (cells + 3u) / 4u; x += THREADS) cmask_l[x] = 0u;
    mbar_wait(bar, 0);
    __syncthreads();
#pragma unroll 1
    for (int i = tid; i < n; i += THREADS) {
        const int c = (int)gpos[i];
        if (c >= lo && c < hi) {
            grid_l[c - lo] |= (uint16_t)(i + 1);
            const int r = (int)__umulhi((uint32_t)c, P.magic_w), col = c - r * W;
            const uint32_t prev = atomicOr(&blk_l[(r - r0 + 1) * RW + 1 + (col >> 5)], 1u << (col & 31));
            if ((prev >> (col & 31)) & 1u) atomicOr(P.err, 128);       // the cell already holds a pedestrian (or is a wall)
        }
    }
    sync_all();

    const uint32_t episode = P.episode_base + (uint32_t)e;
    const double* mv_draws = P.move_draws ? P.move_draws + (size_t)e * P.draw_steps * P.n_max : nullptr;
    const double* cf_draws = P.conflict_draws ? P.conflict_draws + (size_t)e * P.draw_steps * HW * 2 : nullptr;
    const int cpr = WW;                                   // 32-cell chunks (one bitboard word) per row
    const int nchunks = RBl * cpr;

    auto refresh_prefix = [&]() { refresh_alive_prefix(alive, wpre, nwords, lane); };
    auto rank_of = [&](uint32_t id) -> uint32_t {
        return wpre[id >> 5] + (uint32_t)__popc(alive[id >> 5] & ((1u << (id & 31u)) - 1u));
    };
    // static wall word at local bitboard index idx (row lr + 1 of this band)
    auto wall_word = [&](int idx) -> uint32_t {
        return P.wall_in_smem ? wall_l[idx] : __ldg(P.wall_bits + (size_t)r0 * RW + idx);
    };
    // pointer to word (j+1) of the bitboard row above / below local row lr (a neighbour CTA's row at a band boundary)
    auto blk_row_word = [&](int lr_abs /* local row index incl. the +1 offset */, int jw) -> const uint32_t* {
        if (CL > 1) {
            if (lr_abs == 0 && !top) return cg::this_cluster().map_shared_rank(blk_l, band - 1) + RB * RW + jw;
            if (lr_abs == RBl + 1 && !bottom) return cg::this_cluster().map_shared_rank(blk_l, band + 1) + 1 * RW + jw;
        }
        return blk_l + lr_abs * RW + jw;
    };
    // 3x3 candidate mask of cell c = (r, col) from the bitboard: bit k set <-> neighbour k passable and unoccupied
    auto cand_mask = [&](int r, int col) -> uint32_t {
        const int lr = r - r0 + 1, cm = col - 1;          // window starts at column col-1 (>= 0: border cells hold no pedestrians)
        const int jw = (cm >> 5) + 1, sh = cm & 31;
        const uint32_t* pu = blk_row_word(lr - 1, jw);
        const uint32_t* pm = blk_l + lr * RW + jw;
        const uint32_t* pd = blk_row_word(lr + 1, jw);
        const uint32_t fu = ~__funnelshift_r(pu[0], pu[1], sh) & 7u;
        const uint32_t fm = ~__funnelshift_r(pm[0], pm[1], sh) & 7u;
        const uint32_t fd = ~__funnelshift_r(pd[0], pd[1], sh) & 7u;
        if (NBR == 8) return fu | ((fm & 5u) << 4) | (fd << 8);           // window bit space (see cand_off)
        return ((fu >> 1) & 1u) | (((fd >> 1) & 1u) << 1) | ((fm & 1u) << 2) | ((fm & 4u) << 1);
    };
    // positions in alive-rank order -> dst[rank] (epilogue / dense trajectory rows)
    auto emit_positions = [&](uint32_t* dst) {
#pragma unroll 1
        for (int ch = tid; ch < nchunks; ch += THREADS) {
            const int lr = cpr == 1 ? ch : (int)__umulhi((uint32_t)ch, P.magic_cpr), j = ch - lr * cpr;
            const int idx = (lr + 1) * RW + 1 + j;
            uint32_t occ = blk_l[idx] & ~wall_word(idx);
            const int cbase = (r0 + lr) * W + 32 * j;
            while (occ) {
                const int b = __ffs(occ) - 1;
                occ &= occ - 1u;
                const int c = cbase + b;
                const uint32_t id = (grid_l[c - lo] & OCC_MASK) - 1u;
                dst[rank_of(id)] = (uint32_t)c;
            }
        }
    };

    const StencilGeom sgeom = make_stencil_geom(RBl, W, tid, THREADS);   // DFF stencil geometry of this band, once (not per step)
    // update_dff (:106-117) of this CTA's rows: current buffer -> the other one.  Width a multiple of 4: the vectorised
    // stencil; rows of a neighbouring band (cluster) come through distributed shared memory.
    auto dff_update = [&]() {
        if (FIELDS_IN_SMEM) {
            const float* inb = dpar ? dffB_l : dffA_l;
            float* outb = dpar ? dffA_l : dffB_l;
            const float* in_up = (CL > 1 && !top) ? cg::this_cluster().map_shared_rank(const_cast<float*>(inb), band - 1) + (RB - 1) * W : nullptr;
            const float* in_dn = (CL > 1 && !bottom) ? cg::this_cluster().map_shared_rank(const_cast<float*>(inb), band + 1) : nullptr;
            auto edge = [&](int r) -> const float* { return r < r0 ? in_up : in_dn; };   // nullptr outside the map
            if ((W & 3) == 0 && W >= 32)
                dff_stencil_v4<NBR>([&](int r) -> const float* { return inb + (r - r0) * W; }, edge,
                                    [&](int r) -> float* { return outb + (r - r0) * W; }, r0, r1, W, P.c0, P.c1, P.thr, tid, sgeom);
            else
                dff_stencil_rows<NBR>([&](int r) -> const float* { return (r >= r0 && r < r1) ? inb + (r - r0) * W : edge(r); },
                                      [&](int r) -> float* { return outb + (r - r0) * W; }, r0, r1, W, P.c0, P.c1, P.thr, tid, sgeom);
        } else {
            const float* inb = dpar ? dffB_g : dffA_g;
            float* outb = dpar ? dffA_g : dffB_g;
            auto edge = [&](int r) -> const float* { return (r < 0 || r >= H) ? nullptr : inb + (size_t)r * W; };
            if ((W & 3) == 0 && W >= 32)
                dff_stencil_v4<NBR>([&](int r) -> const float* { return inb + (size_t)r * W; }, edge,
                                    [&](int r) -> float* { return outb + (size_t)r * W; }, r0, r1, W, P.c0, P.c1, P.thr, tid, sgeom);
            else
                dff_stencil_rows<NBR>(edge, [&](int r) -> float* { return outb + (size_t)r * W; }, r0, r1, W, P.c0, P.c1, P.thr, tid, sgeom);
        }
    };

#ifdef FFM_PHASE_TIMING
    long long tw[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tc = clock64();
#define FFM_TICK(i) { const long long now_ = clock64(); tw[i] += now_ - tc; tc = now_; }
#else
#define FFM_TICK(i)
#endif
    unsigned long long ped_steps = 0;
    // compact trajectory record: entries of this episode's stream written so far (-1: overflowed) live in ctr[14], not in a
    // register of the hot loop
    if (P.ctraj != nullptr && tid == 0 && band == 0) P.ctraj_off[(size_t)e * (P.traj_steps + 1)] = 0;
    bool need_prefix = false;      // somebody left since the prefix was last computed
    bool dff_pending = false;      // the DFF update of the previous step has not run yet
    int tl = 0;
    for (; tl < P.max_steps && n > 0; ++tl) {
        const uint32_t t = (uint32_t)(t0 + tl);
        ped_steps += (unsigned long long)n;
        const int di = (int)t - P.draw_first;
        const bool inj = di >= 0 && di < P.draw_steps;
        uint32_t* cnt = ctr + ((tl & 1) << 3);

        // ================= phase 1: cells -> movers, by candidate count =================================
        {
            uint4* cm4 = reinterpret_cast<uint4*>(cmask_l);                   // claim masks: clean for this step
            const uint32_t n16 = (cells + 15u) / 16u;
            if ((uint32_t)tid < n16) cm4[tid] = make_uint4(0u, 0u, 0u, 0u);
            for (uint32_t x = tid + THREADS; x < n16; x += THREADS) cm4[x] = make_uint4(0u, 0u, 0u, 0u);
        }
        if (need_prefix && warp == NW - 1) refresh_prefix();
        for (int ch0 = 0; ch0 < nchunks; ch0 += THREADS) {
            const int ch = ch0 + tid;
            uint32_t mov = 0, m2 = 0, m3 = 0;
            int cbase = 0;
            if (ch < nchunks) {
                const int lr = cpr == 1 ? ch : (int)__umulhi((uint32_t)ch, P.magic_cpr), j = ch - lr * cpr;
                const int idx = (lr + 1) * RW + 1 + j;
                const uint32_t occw = blk_l[idx] & ~wall_word(idx);            // pedestrians of this word
                if (occw != 0u) {
                    cbase = (r0 + lr) * W + 32 * j;                            // cell of bit 0
                    // free cells of the three rows, and the same shifted by one column either way (bit i <-> column 32j + i;
                    // the bit shifted in comes from the neighbouring word)
                    const uint32_t* pu = blk_row_word(lr, j + 1);
                    const uint32_t* pm = blk_l + idx;
                    const uint32_t* pd = blk_row_word(lr + 2, j + 1);
                    const uint32_t fu = ~pu[0], fm = ~pm[0], fd = ~pd[0];
                    const uint32_t ful = __funnelshift_l(~pu[-1], fu, 1), fur = __funnelshift_r(fu, ~pu[1], 1);   // column - 1 / + 1
                    const uint32_t fml = __funnelshift_l(~pm[-1], fm, 1), fmr = __funnelshift_r(fm, ~pm[1], 1);
                    const uint32_t fdl = __funnelshift_l(~pd[-1], fd, 1), fdr = __funnelshift_r(fd, ~pd[1], 1);
                    // number of free neighbours of every cell, bit-sliced (ones / twos / fours / eights)
                    uint32_t ones, twos, more;
                    if (NBR == 8) {
                        const uint32_t x0 = ful, x1 = fu, x2 = fur, x3 = fml, x4 = fmr, x5 = fdl, x6 = fd, x7 = fdr;
                        const uint32_t sa = x0 ^ x1 ^ x2, ca = (x0 & x1) | (x2 & (x0 ^ x1));
                        const uint32_t sb = x3 ^ x4 ^ x5, cb = (x3 & x4) | (x5 & (x3 ^ x4));
                        const uint32_t sc = x6 ^ x7, cc = x6 & x7;
                        ones = sa ^ sb ^ sc;
                        const uint32_t cd = (sa & sb) | (sc & (sa ^ sb));
                        const uint32_t ts = ca ^ cb ^ cc, tc = (ca & cb) | (cc & (ca ^ cb));
                        twos = ts ^ cd;
                        more = tc | (ts & cd);                                 // fours or eights
                    } else {
                        const uint32_t x0 = fu, x1 = fd, x2 = fml, x3 = fmr;
                        const uint32_t sa = x0 ^ x1 ^ x2, ca = (x0 & x1) | (x2 & (x0 ^ x1));
                        ones = sa ^ x3;
                        const uint32_t c2 = sa & x3;
                        twos = ca ^ c2;
                        more = ca & c2;
                    }
                    mov = occw & (ones | twos | more);                         // at least one candidate (:57-63)
                    m2 = mov & ones & ~(twos | more);                          // exactly one free neighbour (+ stay)
                    m3 = mov & twos & ~(ones | more);                          // exactly two
                }
            }
            if (__ballot_sync(FULL, mov != 0u) == 0u) continue;
            // list offsets: one packed inclusive scan over the warp (10 bits per class), one atomic per class per warp
            const uint32_t mb = mov & ~(m2 | m3);
            const uint32_t mine = (uint32_t)__popc(m2) | ((uint32_t)__popc(m3) << 10) | ((uint32_t)__popc(mb) << 20);
            uint32_t inc = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(FULL, inc, d);
                if (lane >= d) inc += y;
            }
            const uint32_t tot = __shfl_sync(FULL, inc, 31);
            uint32_t b23 = 0, bb = 0;
            if (lane == 0) {     // classes 2 and 3 share one packed counter (16 bits each: at most 16380 pedestrians)
                b23 = smem_atomic_add(&cnt[0], (tot & 0x3FFu) | (((tot >> 10) & 0x3FFu) << 16));
                if (tot >> 20) bb = smem_atomic_add(&cnt[2], tot >> 20);
            }
            b23 = __shfl_sync(FULL, b23, 0);
            bb = __shfl_sync(FULL, bb, 0);
            const uint32_t ex = inc - mine;
            int o2 = (int)((b23 & 0xFFFFu) + (ex & 0x3FFu));
            int o3 = (int)((b23 >> 16) + ((ex >> 10) & 0x3FFu));
            int ob = (int)(bb + (ex >> 20));
            for (uint32_t m = m2; m; m &= m - 1u) listA[o2++] = (EntT)(cbase + __ffs(m) - 1);
            for (uint32_t m = m3; m; m &= m - 1u) listA[cap1 - (o3++)] = (EntT)(cbase + __ffs(m) - 1);
            for (uint32_t m = mb; m; m &= m - 1u) listB[ob++] = (EntT)(cbase + __ffs(m) - 1);
        }
        FFM_TICK(0)
        // DFF decay + diffusion of the previous step (reads the field with that step's footprints) -> other buffer
        if (DFF && dff_pending) dff_update();
        // (global DFF rows written here are read by the neighbour CTAs in the next step: the cluster barrier below is a
        //  release / acquire pair at cluster scope and invalidates L1, so no extra fence is needed -- an explicit
        //  __threadfence() by 1024 threads showed up as 1.4 membar stall cycles per issue)
        FFM_TICK(1)
        sync_all();
        FFM_TICK(2)
        if (DFF && dff_pending) dpar ^= 1;
        dff_pending = false;
        need_prefix = false;
        if (tid == 0) {                                   // the other parity's counters: last read right after the previous
            uint32_t* nx = ctr + (((tl + 1) & 1) << 3);   // step's final barrier, next written in the next step's phase 1
            nx[0] = 0u; nx[1] = 0u; nx[2] = 0u; nx[3] = 0u; nx[4] = 0u; nx[5] = 0u;
        }

        // ================= phase 2: movers -> requests ==================================================
        const int n2 = (int)(cnt[0] & 0xFFFFu), n3 = (int)(cnt[0] >> 16), nb = (int)cnt[2];
        // One mover: candidates in neighbour order then "stay" (the order of the reference's neighbor_coords array,
        // :54,60,64).  The lists are sorted by candidate count, so the trip counts below are warp-uniform.
        auto decide = [&](EntT* slot_ptr) {
            const int c = (int)*slot_ptr;
            const int r = (int)__umulhi((uint32_t)c, P.magic_w), col = c - r * W;
            uint32_t mm = cand_mask(r, col);
            const uint32_t g = grid_l[c - lo];
            const uint32_t id = (g & OCC_MASK) - 1u;
            int kdir = -1;                                                    // neighbour index of the request
            bool footprint = false;
            if ((g >> TYPE_SHIFT) >= TYPE_EXIT) {
                // on an exit, or a free cell next to one: an exit among the candidates forces the request, no draw (:66-72)
                uint32_t exm = 0;
#pragma unroll
                for (int k = 0; k < NBR; ++k)
                    if (((mm >> cand_bit<NBR>(k)) & 1u) && *grid_ptr(c + nbr_off<NBR>(k, W)) == EXIT_EMPTY) exm |= 1u << k;
                if ((g >> TYPE_SHIFT) == TYPE_EXIT && exm == 0u) { footprint = true; mm = 0u; }    // "stay" is the first exit (:64)
                else if (exm != 0u) { kdir = __ffs(exm) - 1; mm = 0u; }
            }
            if (mm != 0u) {
                const int ncand = __popc(mm) + 1;
                int kb[NBR + 1];
                S p[NBR + 1];
                double tot = 0.0;
                if (ncand == 2) {
                    // one free neighbour + "stay" (half of all movers in a packed crowd): the larger score gives exp(0) = 1
                    // and the other one exp(-|difference|) -- the same two values as the general path, one exp
                    kb[0] = __ffs(mm) - 1;
                    const int cc = c + cand_off<NBR>(kb[0], W);
                    S s0 = score_at(cc), s1 = score_at(c);
                    if (DFF) { s0 = add_rn(s0, (S)mul_rn(P.kd, *dff_cur(cc))); s1 = add_rn(s1, (S)mul_rn(P.kd, *dff_cur(c))); }
                    const S d = add_rn(s0, -s1);
                    const S ex = exp_t(d >= (S)0 ? -d : d);                        // NaN (inf - inf) falls through to "no request"
                    p[0] = d >= (S)0 ? (S)1 : ex;
                    p[1] = d >= (S)0 ? ex : (S)1;
                    tot = (double)p[0] + (double)p[1];
                } else {
                    S mx = neg_inf<S>();
                    uint32_t m = mm;
#pragma unroll
                    for (int j = 0; j <= NBR; ++j)
                        if (j < ncand) {
                            int cc = c;
                            kb[j] = 0;
                            if (j < ncand - 1) {
                                kb[j] = __ffs(m) - 1;
                                m &= m - 1u;
                                cc = c + cand_off<NBR>(kb[j], W);
                            }
                            S sc = score_at(cc);                                       // -k_S * sff
                            if (DFF) sc = add_rn(sc, (S)mul_rn(P.kd, *dff_cur(cc)));   // + k_D * dff   (:77)
                            p[j] = sc;
                            mx = max_t(mx, sc);
                        }
#pragma unroll
                    for (int j = 0; j <= NBR; ++j)
                        if (j < ncand) {
                            p[j] = exp_t(add_rn(p[j], -mx));                           // exp(score - max) (:80)
                            tot += (double)p[j];
                        }
                }
                if (isfinite(tot) && tot != 0.0) {                                // (:82)
                    const uint32_t rank = rank_of(id);                            // the reference's array index
                    const double u = (inj && mv_draws) ? mv_draws[(size_t)di * P.n_max + rank]
                                                       : draw_u0(P.seed, episode, t, STREAM_MOVE, rank);
                    const double thresh = u * tot;
                    double run = 0.0;
                    footprint = true;                                             // E_n > u * E_n always: "stay" is last
#pragma unroll
                    for (int j = 0; j < NBR; ++j)
                        if (j < ncand - 1 && footprint) {
                            run += (double)p[j];
                            if (run > thresh) { kdir = cand_dir<NBR>(kb[j]); footprint = false; }
                        }
                }
            }
            EntT out = (EntT)0;                                                   // cell 0 is a border cell: "no request"
            if (kdir >= 0) {
                const uint32_t target = (uint32_t)(c + nbr_off_rt<NBR>(kdir, W));
                uint32_t sh;
                uint32_t* wp = cmask_word((int)target, sh);
                const uint32_t old = atomicOr(wp, (1u << nbr_opp<NBR>(kdir)) << sh);
                if (((old >> sh) & 0xFFu) == 0u) out = (EntT)target;              // first claimant: this entry resolves the cell
            }
            if (DFF && footprint) { float* d = dff_cur(c); *d = __fadd_rn(*d, 1.0f); }   // a granted "stay" (:91-93)
            *slot_ptr = out;
        };
        {
            // warp tasks of 32 entries, the long ones (most candidates) first; warp w takes tasks w, w + NW, ...
            const int tb = (nb + 31) >> 5, t3 = (n3 + 31) >> 5, t2 = (n2 + 31) >> 5;
            for (int task = warp; task < tb + t3 + t2; task += NW) {
                EntT* ptr;
                bool valid;
                if (task < tb) { const int i = task * 32 + lane; valid = i < nb; ptr = &listB[i]; }
                else if (task < tb + t3) { const int i = (task - tb) * 32 + lane; valid = i < n3; ptr = &listA[cap1 - i]; }
                else { const int i = (task - tb - t3) * 32 + lane; valid = i < n2; ptr = &listA[i]; }
                if (valid) decide(ptr);
            }
        }
        FFM_TICK(3)
        sync_all();
        FFM_TICK(4)

        // ================= phase 3: one resolver per requested cell; moves applied at once ===============
        // apply the move of the occupant of T's neighbour `from` into T
        auto apply_move = [&](uint32_t T, int from) -> bool {
            const int q = (int)T + nbr_off_rt<NBR>(from, W);                      // the winner's cell
            uint16_t* gq = grid_ptr(q);
            uint16_t* gT = grid_ptr((int)T);
            const uint32_t vq = *gq, vT = *gT;
            const int rT = (int)__umulhi(T, P.magic_w), cT = (int)T - rT * W;
            const int rq = rT + nbr_dr_rt<NBR>(from), cq = cT + nbr_dc_rt<NBR>(from);
            uint32_t bit;
            uint32_t* bw = blk_word(rq, cq, bit);
            atomicAnd(bw, ~bit);
            *gq = (uint16_t)(vq & TYPE_BITS);
            bool leaves = false;
            if (vT == EXIT_EMPTY) {                                               // reached an exit: removed (:101-102)
                leaves = true;
                const uint32_t id = (vq & OCC_MASK) - 1u;
                if (CL == 1) {
                    atomicAnd(&alive[id >> 5], ~(1u << (id & 31u)));
                } else {
#pragma unroll
                    for (int rk = 0; rk < CL; ++rk)
                        atomicAnd(cg::this_cluster().map_shared_rank(alive, rk) + (id >> 5), ~(1u << (id & 31u)));
                }
            } else {
                *gT = (uint16_t)(vT | (vq & OCC_MASK));
                bw = blk_word(rT, cT, bit);
                atomicOr(bw, bit);
            }
            if (DFF) { float* dq = dff_cur(q); *dq = __fadd_rn(*dq, 1.0f); }      // footprint (:93,98)
            return leaves;
        };
        auto count_exits = [&](bool leaves) {
            const uint32_t bal = __ballot_sync(FULL, leaves);
            if (bal != 0u && lane == 0) {
                if (CL == 1) {
                    smem_atomic_add(&cnt[3], (uint32_t)__popc(bal));
                } else {
#pragma unroll
                    for (int rk = 0; rk < CL; ++rk) atomicAdd(cg::this_cluster().map_shared_rank(cnt, rk) + 3, (uint32_t)__popc(bal));
                }
            }
        };
        const int ntot = n2 + n3 + nb;
        // contested cells (at most ntot / 2) are listed in the larger free part of the two work lists: the middle of
        // listA has cap - n2 - n3 free entries, the tail of listB cap - nb, together at least cap >= ntot
        EntT* clist = ((int)L.cap - n2 - n3 >= (int)L.cap - nb) ? listA + n2 : listB + nb;
        for (int x0 = 0; x0 < ntot; x0 += THREADS) {
            const int x = x0 + tid;
            bool leaves = false, contested = false;
            uint32_t T = 0;
            if (x < ntot) {
                T = (uint32_t)(x < n2 ? listA[x] : (x < n2 + n3 ? listA[cap1 - (x - n2)] : listB[x - n2 - n3]));
                if (T != 0u) {
                    uint32_t sh;
                    const uint32_t cm = (*cmask_word((int)T, sh) >> sh) & 0xFFu;  // bit k: the occupant of neighbour k wants T
                    if (cm & (cm - 1u)) contested = true;                         // two or more claimants: resolved below
                    else leaves = apply_move(T, __ffs(cm) - 1);                   // lone claimant: moves (:91-93)
                }
            }
            count_exits(leaves);
            // contested cells are collected and resolved below
            const uint32_t cb = __ballot_sync(FULL, contested);
            if (cb != 0u) {
                uint32_t base = 0;
                if (lane == 0) base = smem_atomic_add(&cnt[4], (uint32_t)__popc(cb));
                base = __shfl_sync(FULL, base, 0);
                if (contested) clist[base + __popc(cb & lanemask_lt())] = (EntT)T;
            }
        }
        // k >= 2 claimants of cell clist[i]: coin, then the floor(u*k)-th claimant in ascending agent index (:94-98)
        auto resolve_contested = [&](int i, int ncont) {
            bool leaves = false;
            if (i < ncont) {
                const uint32_t T = (uint32_t)((volatile EntT*)clist)[i];
                uint32_t sh;
                const uint32_t cm = (*cmask_word((int)T, sh) >> sh) & 0xFFu;
                const int k = __popc(cm);
                Draw2 d;
                if (inj && cf_draws) {
                    d.u0 = cf_draws[((size_t)di * HW + T) * 2];
                    d.u1 = cf_draws[((size_t)di * HW + T) * 2 + 1];
                } else {
                    d = draw2(P.seed, episode, t, STREAM_CONFLICT, T);
                }
                if (d.u0 < 0.5) {                                         // coin (:95): somebody moves
                    // agents[int(u * k)] in ascending agent index (:96): sort (id, direction) keys
                    uint32_t key[NBR];
#pragma unroll
                    for (int q = 0; q < NBR; ++q) {
                        key[q] = FULL;
                        if ((cm >> q) & 1u) key[q] = (((uint32_t)*grid_ptr((int)T + nbr_off<NBR>(q, W)) & OCC_MASK) << 3) | (uint32_t)q;
                    }
                    sort_keys<NBR>(key);
                    const int w = (int)(d.u1 * (double)k);
                    uint32_t sel = key[0];
#pragma unroll
                    for (int q = 1; q < NBR; ++q) if (q == w) sel = key[q];
                    leaves = apply_move(T, (int)(sel & 7u));
                }
            }
            count_exits(leaves);
        };
        if (THREADS >= 512) {
            // a large CTA: every warp takes its share of the contested cells once the list is complete
            __syncthreads();
            const int ncont = (int)cnt[4];
            for (int i0 = warp * 32; i0 < ncont; i0 += THREADS) resolve_contested(i0 + lane, ncont);
        } else {
            // contested cells are rare per warp: the last warp to finish the loop above resolves them with full lanes
            __syncwarp();
            uint32_t done = 0;
            if (lane == 0) { __threadfence_block(); done = smem_atomic_add(&cnt[5], 1u); __threadfence_block(); }
            done = __shfl_sync(FULL, done, 0);
            if (done == NW - 1) {
                const int ncont = (int)*((volatile uint32_t*)&cnt[4]);
                for (int i0 = 0; i0 < ncont; i0 += 32) resolve_contested(i0 + lane, ncont);
            }
        }
        FFM_TICK(5)
        sync_all();
        FFM_TICK(6)

        const int n_exit = (int)cnt[3];
        if (n_exit > 0) { n -= n_exit; need_prefix = true; }
        dff_pending = DFF;

        // trajectory row: positions after this step, alive-rank order (ffm_core.py:125)
        if (P.traj != nullptr && tl < P.traj_steps) {
            if (need_prefix) {
                if (warp == 0) refresh_prefix();
                __syncthreads();
                need_prefix = false;
            }
            emit_positions(P.traj + ((size_t)e * P.traj_steps + tl) * P.n_max);
            if (tid == 0 && band == 0) P.traj_n[(size_t)e * P.traj_steps + tl] = n;
        }
        // compact record: (row, col) int16 pairs of this step's row appended to the episode's stream (CSR offsets)
        if (P.ctraj != nullptr && tl < P.traj_steps) {
            if (need_prefix) {
                if (warp == 0) refresh_prefix();
                __syncthreads();
                need_prefix = false;
            }
            RecArgs R;
            R.ctraj = P.ctraj; R.ctraj_off = P.ctraj_off; R.traj_n = P.traj_n; R.wall_bits = P.wall_bits; R.ctraj_cap = P.ctraj_cap;
            R.l_grid = L.grid; R.l_blk = L.blk; R.l_wall = L.wall; R.l_alive = L.alive; R.l_wpre = L.wpre; R.l_ctr = L.ctr; R.l_listA = L.listA;
            R.magic_cpr = P.magic_cpr; R.H = H; R.W = W; R.RW = RW; R.RB = RB; R.traj_steps = P.traj_steps;
            R.wall_in_smem = P.wall_in_smem; R.write_n = P.traj == nullptr;
            record_step_compact(R, smem_raw, e, tl, n, (int)band, CL == 1);
        }
    }

    // ---- the last step's DFF update ------------------------------------------------------------------------
    if (DFF && dff_pending) {
        dff_update();
        dpar ^= 1;
    }
    if (need_prefix && warp == 0) refresh_prefix();
    sync_all();

    // ---- epilogue: state back to HBM, alive-rank order ---------------------------------------------------------
    emit_positions(gpos);
    if (DFF) {
        if (FIELDS_IN_SMEM) {
            float* cur = dpar ? dffB_l : dffA_l;
            if (tma_fields) {               // shared -> global bulk store of the band's final DFF
                fence_proxy_async_smem();
                __syncthreads();
                if (tid == 0) { bulk_copy_s2g(dff_home + lo, cur, dff_bytes); bulk_commit_wait_all(); }
            } else {
                for (uint32_t x = tid; x < cells; x += THREADS) dff_home[lo + x] = cur[x];
            }
        } else if (dpar) {
            for (int c = lo + tid; c < hi; c += THREADS) dffA_g[c] = dffB_g[c];
        }
    }
#ifdef FFM_PHASE_TIMING
    if (P.dbg != nullptr && lane == 0)
        for (int i = 0; i < 8; ++i) atomicAdd(P.dbg + i, (unsigned long long)tw[i]);
#endif
    if (tid == 0 && band == 0) {
        P.n_alive[e] = n;
        P.t_done[e] = t0 + tl;
        P.ped_steps[e] += ped_steps;
    }
    if (CL > 1) cg::this_cluster().sync();     // nobody leaves while a neighbour may still read its shared memory
}

}  // namespace ffm
