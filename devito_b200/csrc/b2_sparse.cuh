// Sparse source injection / receiver interpolation kernels (internal interface).
#pragma once
#include "b2_common.cuh"

namespace b2 {

struct SparseDev {
    bool present = false;
    DevArray data, gp, w[3];
    int npoint_total = 0;   // second extent of data
    int nt = 0;
    int p_m = 0, p_M = -1;
    int r = 1;
    int ndim = 3;
};

// Field geometry the sparse kernels need (internal 3-dim convention, see IsoPlan)
struct FieldGeom {
    long long sx = 0, sy = 0;
    size_t slot_elems = 0;
    int so = 0;
    int ndim = 3;
    int lo[3] = {0, 0, 0};      // x_m, y_m, z_m
    int hi[3] = {0, 0, 0};      // x_M, y_M, z_M
    // x-slab decomposition: a neighbour owns the cells beyond x_m / x_M, so injection must not
    // touch them (the owner injects and its boundary planes are then exchanged)
    bool nb_lo = false, nb_hi = false;
    // streamed time loop (b2_api_iso.cu): [lo[0], hi[0]] is one x-range of a skewed sweep; interpolation then
    // samples only the cells of that range and ADDS its partial sum to the trace
    bool restrict_x = false;
    // ... or a range of dim 1 (the streamed loop under x-slab decomposition skews along y): [lo[1], hi[1]] is the
    // range, cut_lo1 / cut_hi1 say whether its ends are interior cuts (no reach beyond them)
    bool restrict_y = false, cut_lo1 = false, cut_hi1 = false;
    // fused halo step: cells injected into the first / last `pw` owned planes are mirrored into the
    // neighbour's halo copy (the sweep kernel stored those planes there before the injection)
    float *peer_lo = nullptr, *peer_hi = nullptr;
    long long off_lo = 0, off_hi = 0;        // element offset of "my plane 0" in the neighbour's array (slot included)
    int nown = 0, pw = 0;
};

int sparse_stage_in(const b2_sparse *s, int ndim, SparseDev &out, bool copy_data_in);
int sparse_stage_out(SparseDev &s, bool copy_data_back);

// u[cell] += w*w*w * src[time][p] * scale(cell) for up to two fields (TTI injects into u and v)
// scale(cell) = dt^2 * vp^2 | dt^2 * vp[cell]^2 | dt^2 / m[cell]
int launch_inject(const SparseDev &s, const FieldGeom &g, float *f0, float *f1, int time,
                  int param_kind, const float *param, float scalar_scale, float dt2);

// rec[time][p] = sum w*w*w * (f0[cell] (+ f1[cell]))
int launch_interp(const SparseDev &s, const FieldGeom &g, const float *f0, const float *f1, int time);

}  // namespace b2
