// update_dff (model/ffm_core.py:106-117 = ffm_unified.py:779-798 = ffm_trained_core.py:333-353) for maps whose
// width is a multiple of 4: the vectorised form of dff_decay_diffuse / dff_stencil_rows, same arithmetic bit for bit.
//
//   new = c0 * dff                      (:109; c0 = f32((1-decay)(1-diffuse)))
//   new += c1 * shift_k(pad(new))       (:111-114; one rounded product and one rounded add per neighbour, list order)
//   new[new < 1e-4] = 0                 (:116-117)
//
// A thread owns FOUR adjacent columns and walks down a strip of rows with a 3-row register window: per row one
// 16-byte load, the two halo columns come from the neighbouring lanes by shuffle (a real load only at a warp edge, an
// exact 0 at the map edge: np.pad), the 6 scaled values s = c0*d and the 6 products u = c1*s are formed once, and the
// four outputs are four independent chains of 8 (4) rounded adds.  ~15 instructions per cell instead of ~90 in the
// scalar walk (ncu, profiles/r2g_*), one shared-memory wavefront per 8 cells.
#pragma once
#include "ffm_device.cuh"

namespace ffm {

// Per-thread geometry of the walk: depends on the band height, the width and the CTA size only, so a persistent kernel computes it
// ONCE before its step loop (the integer divisions below were 6 % of the unified kernel's instructions when redone every step).
struct StencilGeom {
    int G4, rpb, strip, gfirst, gstep;   // vectorised walk (dff_stencil_v4)
    bool strip_in_range;
    int cw, bands, srpb, band, colb;     // scalar walk (dff_stencil_rows / dff_decay_diffuse)
};
__device__ __forceinline__ StencilGeom make_stencil_geom(int Hb, int W, int tid, int nthreads) {
    StencilGeom g;
    g.G4 = W >> 2;
    const int G4 = g.G4 > 0 ? g.G4 : 1;
    const int strips = G4 < nthreads ? nthreads / G4 : 1;
    g.rpb = (max(Hb, 1) + strips - 1) / strips;
    g.strip = G4 < nthreads ? tid / G4 : 0;
    g.gfirst = G4 < nthreads ? tid - g.strip * G4 : tid;
    g.gstep = G4 < nthreads ? G4 : nthreads;
    g.strip_in_range = g.strip < strips;
    g.cw = W < nthreads ? W : nthreads;
    g.bands = W < nthreads ? nthreads / W : 1;
    g.srpb = (max(Hb, 1) + g.bands - 1) / g.bands;
    g.band = tid / g.cw;
    g.colb = tid - g.band * g.cw;
    return g;
}

// in_band(r)  -> row r of the input for r0 <= r < r1 (the caller's own rows: plain shared / global pointer)
// in_edge(r)  -> row r outside [r0, r1): the neighbouring band's row (distributed shared memory) or nullptr outside the map
// out_band(r) -> row r of the output, r0 <= r < r1
// All rows are 16-byte aligned, W % 4 == 0.  Every thread of the CTA must call (warp shuffles inside).
template <int NBR, typename InBand, typename InEdge, typename OutBand>
__device__ __forceinline__ void dff_stencil_v4(InBand in_band, InEdge in_edge, OutBand out_band, int r0, int r1, int W, float c0,
                                               float c1, float thr, int tid, const StencilGeom& geom) {
    const int G4 = geom.G4;                                   // column groups per row
    const int Hb = r1 - r0;
    if (Hb <= 0) return;
    const int lane = tid & 31;
    const int rpb = geom.rpb;                                 // rows per strip (uniform trip count: shuffles stay converged)
    const int strip = geom.strip, gfirst = geom.gfirst;
    const int gstep = geom.gstep;                             // column passes when a row has more groups than the CTA threads
    const int a0 = r0 + strip * rpb;
    const int a1 = min(r1, a0 + rpb);
    const bool strip_ok = geom.strip_in_range && a0 < r1;
    for (int gb = 0; gb < G4; gb += gstep) {
        const int g = gb + gfirst;
        const bool act = strip_ok && g < G4;
        const bool edge_l = g == 0, edge_r = g == G4 - 1;
        const int c4 = g << 2;

        // one row -> u[0..5] = c1 * (c0 * d) for columns c4-1 .. c4+4, s[0..3] = c0 * d for the thread's own columns
        auto load_row = [&](int r, float (&u)[6], float (&s)[4]) {
            float4 d = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            const float* row = nullptr;
            if (act && r <= a1) {                             // rows a0-1 .. a1 feed this strip; in_edge only ever sees r0-1 and r1
                row = (r >= r0 && r < r1) ? in_band(r) : in_edge(r);
                if (row != nullptr) d = *reinterpret_cast<const float4*>(row + c4);
            }
            float dl = __shfl_up_sync(0xffffffffu, d.w, 1);
            float dr = __shfl_down_sync(0xffffffffu, d.x, 1);
            if (edge_l) dl = 0.0f;
            else if (lane == 0) dl = row != nullptr ? row[c4 - 1] : 0.0f;
            if (edge_r) dr = 0.0f;
            else if (lane == 31) dr = row != nullptr ? row[c4 + 4] : 0.0f;
            const float sl = __fmul_rn(c0, dl), sr = __fmul_rn(c0, dr);                            // (:109)
            s[0] = __fmul_rn(c0, d.x); s[1] = __fmul_rn(c0, d.y); s[2] = __fmul_rn(c0, d.z); s[3] = __fmul_rn(c0, d.w);
            u[0] = __fmul_rn(c1, sl);                                                              // (:113)
            u[1] = __fmul_rn(c1, s[0]); u[2] = __fmul_rn(c1, s[1]); u[3] = __fmul_rn(c1, s[2]); u[4] = __fmul_rn(c1, s[3]);
            u[5] = __fmul_rn(c1, sr);
        };
        // output row r from the window (up, cur, next); then the next row is loaded into `un`
        auto row_step = [&](int r, const float (&up)[6], const float (&uc)[6], float (&un)[6], const float (&sc)[4], float (&sn)[4]) {
            load_row(r + 1, un, sn);
            float o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float acc = sc[k];
                if (NBR == 8) {   // (-1,-1) (-1,0) (-1,1) (0,-1) (0,1) (1,-1) (1,0) (1,1)
                    acc = __fadd_rn(acc, up[k]); acc = __fadd_rn(acc, up[k + 1]); acc = __fadd_rn(acc, up[k + 2]);
                    acc = __fadd_rn(acc, uc[k]); acc = __fadd_rn(acc, uc[k + 2]);
                    acc = __fadd_rn(acc, un[k]); acc = __fadd_rn(acc, un[k + 1]); acc = __fadd_rn(acc, un[k + 2]);
                } else {          // (-1,0) (1,0) (0,-1) (0,1)
                    acc = __fadd_rn(acc, up[k + 1]); acc = __fadd_rn(acc, un[k + 1]);
                    acc = __fadd_rn(acc, uc[k]); acc = __fadd_rn(acc, uc[k + 2]);
                }
                o[k] = acc < thr ? 0.0f : acc;                                                     // (:116-117)
            }
            if (act && r < a1) *reinterpret_cast<float4*>(out_band(r) + c4) = make_float4(o[0], o[1], o[2], o[3]);
        };
        float A[6], B[6], C[6], sA[4], sB[4], sC[4];
        load_row(a0 - 1, A, sA);
        load_row(a0, B, sB);
        const int aend = a0 + rpb;
#pragma unroll 1
        for (int r = a0; r < aend; r += 3) {                  // the window rotates by renaming (no register moves)
            row_step(r, A, B, C, sB, sC);
            if (r + 1 < aend) row_step(r + 1, B, C, A, sC, sA);
            if (r + 2 < aend) row_step(r + 2, C, A, B, sA, sB);
        }
    }
}

}  // namespace ffm
