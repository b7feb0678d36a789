/*
 * CPU restatement in C of the reference's base CA step -- TEST INFRASTRUCTURE / CPU BASELINE ONLY
 * (see oracle/__init__.py; nothing under ffm_b200/ links or calls this).
 *
 * Follows /root/reference/model/ffm_core.py function by function with an occupancy grid in place of
 * the per-agent np.delete + np.isin (ffm_core.py:48-60); arithmetic mirrors NumPy's:
 *   score = -k_S*sff + k_D*dff            separate multiply and add, dtype = promote(sff, float32)  (:77)
 *   probs = exp(score - max); probs /= probs.sum()   NumPy pairwise-sum order for n <= 9          (:78-83)
 *   np.random.choice(n, p)                p -> float64, cumsum, / last, searchsorted(u, 'right')   (:84)
 *   conflicts                             coin u0 < 0.5, winner agents[int(u1*k)]                  (:90-98)
 *   update_dff                            float32 scalars, pad-then-shift accumulation, threshold  (:106-117)
 * Draws: Philox4x32-10 keyed (entity, step, episode, stream), the protocol of oracle/philox.py.
 * Pinned by tests/test_c_oracle.py against oracle/ffm_numpy.py (itself pinned to the reference).
 * Build: make -C oracle/c   (gcc -O2 -ffp-contract=off: no FMA contraction).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { STREAM_MOVE = 0, STREAM_CONFLICT = 1 };

static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
static double u53(uint32_t a, uint32_t b) {
    return (double)(((uint64_t)(a >> 5) << 26) | (uint64_t)(b >> 6)) * (1.0 / 9007199254740992.0);
}
static void draw2(uint64_t seed, uint32_t episode, uint32_t step, uint32_t stream, uint32_t entity, double* u0, double* u1) {
    uint32_t c[4] = {entity, step, episode, stream};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    *u0 = u53(c[0], c[1]);
    *u1 = u53(c[2], c[3]);
}

typedef struct {
    const uint8_t* map; const void* sff; int sff_f64; int H, W, nbr;
    double k_S, k_D; float c0, c1, thr;
    const int32_t* pos_rc; const int32_t* n; int B, n_max;
    uint64_t seed; uint32_t episode_base; int max_steps, track_dff;
    int32_t* steps; int64_t* ped_steps; double* min_margin; int32_t* final_pos; int32_t* final_n; float* final_dff;
    int32_t* traj; int32_t* traj_n; int traj_steps;
    double guard; double* move_out; int move_out_steps;   /* guarded recording of the move draws */
    int next; pthread_mutex_t mu;
} job_t;

static const int DR8[8] = {-1, -1, -1, 0, 0, 1, 1, 1}, DC8[8] = {-1, 0, 1, -1, 1, -1, 0, 1};   /* ffm_core.py:32-34 */
static const int DR4[4] = {-1, 1, 0, 0}, DC4[4] = {0, 0, -1, 1};                                /* ffm_core.py:30 */

/* NumPy add.reduce of n <= 9 contiguous floats (pairwise_sum, umath/loops_utils.h.src) */
static float np_sum_f32(const float* a, int n) {
    if (n < 8) { float r = 0.f; for (int i = 0; i < n; ++i) r += a[i]; return r; }
    float r = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
    for (int i = 8; i < n; ++i) r += a[i];
    return r;
}
static double np_sum_f64(const double* a, int n) {
    if (n < 8) { double r = 0.; for (int i = 0; i < n; ++i) r += a[i]; return r; }
    double r = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
    for (int i = 8; i < n; ++i) r += a[i];
    return r;
}

static void run_episode(job_t* J, int e) {
    const int H = J->H, W = J->W, HW = H * W, nbr = J->nbr;
    const int* DR = nbr == 8 ? DR8 : DR4; const int* DC = nbr == 8 ? DC8 : DC4;
    int n = J->n[e];
    int32_t* pos = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    int32_t* nxt = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    int32_t* tgt = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    int32_t* rnk = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));   /* rank among the claimants of tgt, ascending agent index */
    uint8_t* occ = (uint8_t*)calloc((size_t)HW, 1);
    int32_t* head = (int32_t*)malloc(sizeof(int32_t) * (size_t)HW);    /* per-cell claimant count */
    float* dff = (float*)calloc((size_t)HW, sizeof(float));
    float* scaled = (float*)malloc(sizeof(float) * (size_t)HW);
    float* newd = (float*)malloc(sizeof(float) * (size_t)HW);
    for (int i = 0; i < n; ++i) pos[i] = J->pos_rc[((size_t)e * J->n_max + i) * 2] * W + J->pos_rc[((size_t)e * J->n_max + i) * 2 + 1];
    memset(head, 0, sizeof(int32_t) * (size_t)HW);
    const uint32_t episode = J->episode_base + (uint32_t)e;
    const float kSf = (float)(-J->k_S), kDf = (float)J->k_D;
    const float* sff32 = (const float*)J->sff; const double* sff64 = (const double*)J->sff;
    double min_margin = INFINITY; int64_t ped_steps = 0; int t = 0;
    const int do_dff = J->track_dff;

    for (; t < J->max_steps && n > 0; ++t) {
        ped_steps += n;
        for (int i = 0; i < n; ++i) occ[pos[i]] = 1;
        for (int i = 0; i < n; ++i) {                                       /* ffm_core.py:40 */
            const int c = pos[i], r = c / W, col = c % W;
            int cand[9], nc = 0;
            for (int k = 0; k < nbr; ++k) {
                const int cc = (r + DR[k]) * W + (col + DC[k]);
                const uint8_t m = J->map[cc];
                if ((m == 0 || m == 3) && !occ[cc]) cand[nc++] = cc;        /* :52-60 */
            }
            tgt[i] = -1;
            if (nc == 0) continue;                                          /* :63 */
            cand[nc++] = c;                                                 /* stay last (:64) */
            int forced = -1;
            for (int j = 0; j < nc; ++j) if (J->map[cand[j]] == 3) { forced = cand[j]; break; }   /* :66-68 */
            if (forced >= 0) { tgt[i] = forced; rnk[i] = head[forced]++; continue; }
            double cdf[9]; int ok = 0;
            if (!J->sff_f64) {
                float s[9], mx = -INFINITY;
                for (int j = 0; j < nc; ++j) {
                    float a = kSf * sff32[cand[j]];
                    float b = kDf * dff[cand[j]];
                    s[j] = a + b;                                           /* :77 */
                    if (s[j] > mx) mx = s[j];
                }
                for (int j = 0; j < nc; ++j) s[j] = expf(s[j] - mx);        /* :80 */
                float sum = np_sum_f32(s, nc);                              /* :81 */
                if (isfinite(sum) && sum != 0.f) {                          /* :82 */
                    double run = 0.;
                    for (int j = 0; j < nc; ++j) { s[j] = s[j] / sum; run += (double)s[j]; cdf[j] = run; }   /* :83, choice() */
                    ok = 1;
                }
            } else {
                double s[9], mx = -INFINITY;
                for (int j = 0; j < nc; ++j) {
                    double a = (-J->k_S) * sff64[cand[j]];
                    float b = kDf * dff[cand[j]];
                    s[j] = a + (double)b;
                    if (s[j] > mx) mx = s[j];
                }
                for (int j = 0; j < nc; ++j) s[j] = exp(s[j] - mx);
                double sum = np_sum_f64(s, nc);
                if (isfinite(sum) && sum != 0.) {
                    double run = 0.;
                    for (int j = 0; j < nc; ++j) { s[j] = s[j] / sum; run += s[j]; cdf[j] = run; }
                    ok = 1;
                }
            }
            if (!ok) continue;
            const double last = cdf[nc - 1];
            double u0, u1; draw2(J->seed, episode, (uint32_t)t, STREAM_MOVE, (uint32_t)i, &u0, &u1);
            for (int j = 0; j < nc; ++j) cdf[j] /= last;
            if (J->guard > 0.0) {
                /* recorder of SURVEY.md 8(c): a draw within `guard` of a CDF boundary is re-drawn (next
                 * attempt = same key, entity + attempt * 65536) so that the recorded stream is insensitive
                 * to last-ulp differences of exp between implementations */
                for (uint32_t attempt = 1;; ++attempt) {
                    double mg = INFINITY;
                    for (int j = 0; j < nc; ++j) { const double d = fabs(cdf[j] - u0); if (d < mg) mg = d; }
                    if (mg >= J->guard) break;
                    draw2(J->seed, episode, (uint32_t)t, STREAM_MOVE, (uint32_t)i + attempt * 65536u, &u0, &u1);
                }
            }
            if (J->move_out && t < J->move_out_steps) J->move_out[((size_t)e * J->move_out_steps + t) * J->n_max + i] = u0;
            int idx = 0;
            for (int j = 0; j < nc; ++j) {
                if (cdf[j] <= u0) idx++;                                    /* searchsorted(..., 'right') */
                const double mg = fabs(cdf[j] - u0);
                if (mg < min_margin) min_margin = mg;
            }
            if (idx >= nc) idx = nc - 1;
            tgt[i] = cand[idx];
            rnk[i] = head[cand[idx]]++;
        }
        /* conflicts (:90-98): outcome per target cell is independent of iteration order */
        for (int i = 0; i < n; ++i) nxt[i] = pos[i];
        for (int i = 0; i < n; ++i) {
            const int T = tgt[i];
            if (T < 0) continue;
            const int k = head[T];
            int moved = 0;
            if (k == 1) moved = 1;
            else {
                const int r = rnk[i];
                double u0, u1; draw2(J->seed, episode, (uint32_t)t, STREAM_CONFLICT, (uint32_t)T, &u0, &u1);
                moved = (u0 < 0.5) && ((int)(u1 * (double)k) == r);
            }
            if (moved) { nxt[i] = T; dff[pos[i]] += 1.0f; }
        }
        for (int i = 0; i < n; ++i) { occ[pos[i]] = 0; if (tgt[i] >= 0) head[tgt[i]] = 0; }
        int m = 0;
        for (int i = 0; i < n; ++i) if (J->map[nxt[i]] != 3) pos[m++] = nxt[i];   /* :101-102, stable */
        n = m;
        if (do_dff) {                                                       /* :106-117 */
            for (int c = 0; c < HW; ++c) scaled[c] = J->c0 * dff[c];
            for (int r = 0; r < H; ++r)
                for (int col = 0; col < W; ++col) {
                    float acc = scaled[r * W + col];
                    for (int k = 0; k < nbr; ++k) {
                        const int rr = r + DR[k], cc = col + DC[k];
                        const float v = (rr >= 0 && rr < H && cc >= 0 && cc < W) ? scaled[rr * W + cc] : 0.0f;
                        const float term = J->c1 * v;
                        acc = acc + term;
                    }
                    if (acc < J->thr) acc = 0.0f;
                    newd[r * W + col] = acc;
                }
            float* tmp = dff; dff = newd; newd = tmp;
        } else {
            memset(dff, 0, sizeof(float) * (size_t)HW);
        }
        if (J->traj && t < J->traj_steps) {
            int32_t* row = J->traj + ((size_t)e * J->traj_steps + t) * J->n_max;
            for (int i = 0; i < n; ++i) row[i] = pos[i];
            J->traj_n[(size_t)e * J->traj_steps + t] = n;
        }
    }
    J->steps[e] = t;
    J->ped_steps[e] = ped_steps;
    if (J->min_margin) J->min_margin[e] = min_margin;
    if (J->final_n) J->final_n[e] = n;
    if (J->final_pos) for (int i = 0; i < n; ++i) J->final_pos[(size_t)e * J->n_max + i] = pos[i];
    if (J->final_dff) memcpy(J->final_dff + (size_t)e * HW, dff, sizeof(float) * (size_t)HW);
    free(pos); free(nxt); free(tgt); free(rnk); free(occ); free(head); free(dff); free(scaled); free(newd);
}

static void* worker(void* arg) {
    job_t* J = (job_t*)arg;
    for (;;) {
        pthread_mutex_lock(&J->mu);
        const int e = J->next++;
        pthread_mutex_unlock(&J->mu);
        if (e >= J->B) break;
        run_episode(J, e);
    }
    return NULL;
}

int ffm_oracle_core_run(const uint8_t* map, const void* sff, int sff_f64, int H, int W, int nbr, double k_S, double k_D,
                        float c0, float c1, float thr, const int32_t* pos_rc, const int32_t* n, int B, int n_max,
                        uint64_t seed, uint32_t episode_base, int max_steps, int track_dff, int threads,
                        int32_t* steps, int64_t* ped_steps, double* min_margin, int32_t* final_pos, int32_t* final_n,
                        float* final_dff, int32_t* traj, int32_t* traj_n, int traj_steps,
                        double guard, double* move_out, int move_out_steps) {
    job_t J;
    memset(&J, 0, sizeof(J));
    J.map = map; J.sff = sff; J.sff_f64 = sff_f64; J.H = H; J.W = W; J.nbr = nbr; J.k_S = k_S; J.k_D = k_D;
    J.c0 = c0; J.c1 = c1; J.thr = thr; J.pos_rc = pos_rc; J.n = n; J.B = B; J.n_max = n_max; J.seed = seed;
    J.episode_base = episode_base; J.max_steps = max_steps; J.track_dff = track_dff;
    J.steps = steps; J.ped_steps = ped_steps; J.min_margin = min_margin; J.final_pos = final_pos; J.final_n = final_n;
    J.final_dff = final_dff; J.traj = traj; J.traj_n = traj_n; J.traj_steps = traj_steps;
    J.guard = guard; J.move_out = move_out; J.move_out_steps = move_out_steps;
    pthread_mutex_init(&J.mu, NULL);
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t th[256];
    for (int i = 0; i < threads; ++i) pthread_create(&th[i], NULL, worker, &J);
    for (int i = 0; i < threads; ++i) pthread_join(th[i], NULL);
    pthread_mutex_destroy(&J.mu);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Geodesic static floor field: textbook Dijkstra (binary heap, lazy deletion) from all exit cells
 * over walkable cells (map 0 or 3), float32 path sums d[v] = fl32(d[u] + w), step costs
 * (w_axis, w_diag); w_diag < 0 disables diagonal steps.  Unit costs give BFS levels.  This is the
 * oracle of the GPU relaxation kernel (ffm_b200/csrc/ffm_sff_kernels.cuh); on obstacle-free rooms it
 * coincides with the reference's Create_SFF.py L1 (4-connected) / Linf (8-connected) fields.
 * ---------------------------------------------------------------------------------------------- */
typedef struct { float d; int32_t c; } heap_item;

static void heap_push(heap_item* h, int* n, heap_item it) {
    int i = (*n)++;
    while (i > 0) {
        int p = (i - 1) / 2;
        if (h[p].d <= it.d) break;
        h[i] = h[p]; i = p;
    }
    h[i] = it;
}
static heap_item heap_pop(heap_item* h, int* n) {
    heap_item top = h[0], last = h[--(*n)];
    int i = 0;
    for (;;) {
        int l = 2 * i + 1, r = l + 1, m = i;
        float md = last.d;
        if (l < *n && h[l].d < md) { m = l; md = h[l].d; }
        if (r < *n && h[r].d < md) { m = r; }
        if (m == i) break;
        h[i] = h[m]; i = m;
    }
    h[i] = last;
    return top;
}

int ffm_oracle_geodesic(const uint8_t* map, int H, int W, float w_axis, float w_diag, float* out) {
    const size_t HW = (size_t)H * W;
    size_t cap = HW * 9 + 16;
    heap_item* heap = (heap_item*)malloc(sizeof(heap_item) * cap);
    if (!heap) return -1;
    int hn = 0;
    for (size_t c = 0; c < HW; ++c) {
        out[c] = INFINITY;
        if (map[c] == 3) { out[c] = 0.0f; heap_item it = {0.0f, (int32_t)c}; heap_push(heap, &hn, it); }
    }
    while (hn > 0) {
        heap_item it = heap_pop(heap, &hn);
        if (it.d > out[it.c]) continue;
        const int r = it.c / W, col = it.c % W;
        for (int dr = -1; dr <= 1; ++dr)
            for (int dc = -1; dc <= 1; ++dc) {
                if (dr == 0 && dc == 0) continue;
                const int diag = (dr != 0 && dc != 0);
                if (diag && w_diag < 0.0f) continue;
                const int rr = r + dr, cc = col + dc;
                if (rr < 0 || rr >= H || cc < 0 || cc >= W) continue;
                const uint8_t m = map[(size_t)rr * W + cc];
                if (!(m == 0 || m == 3)) continue;
                const float w = diag ? w_diag : w_axis;
                const float nd = it.d + w;
                if (nd < out[(size_t)rr * W + cc]) {
                    out[(size_t)rr * W + cc] = nd;
                    heap_item ni = {nd, (int32_t)(rr * W + cc)};
                    heap_push(heap, &hn, ni);
                }
            }
    }
    free(heap);
    return 0;
}
