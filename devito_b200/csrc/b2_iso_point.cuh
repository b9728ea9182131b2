// One-point forms of the generic (one thread per grid point) isotropic kernels.
//
// They are `__host__ __device__` so that the SAME code that the __global__ wrappers in b2_iso.cu run
// on the GPU can be looped over on the CPU by tests/support/emu_points.cu and compared with the
// oracle without a GPU (index arithmetic, mirrored taps, OT4 composition). The tiled TMA kernels are
// not covered by this; they are compared with the oracle on the GPU.
#pragma once
#include "b2_common.cuh"

namespace b2 {

struct IsoGK {
    const float *__restrict__ u0;
    const float *__restrict__ um;
    float *__restrict__ u1;
    const float *__restrict__ damp;
    const float *__restrict__ param;
    long long sx, sy;
    int n0, n1, n2;          // extents of this launch (dim0 count is xcount)
    int o0, o1, o2;          // array index of first point (o0 already includes xlo)
    int r0, r1, r2;          // radius per dim
    int param_kind;
    float m_dt2, inv_dt, inv_dt2;
    float w[3][B2_MAX_RADIUS + 1];
    // OT4 (reference kernel='OT4', examples/seismic/acoustic/operators.py:50-68):
    // lap(u) is replaced by lap(u) + dt^2/12 * lap(W), W = lap(u)/m tabulated by ot4_w_point
    float *__restrict__ W;
    float ot4c;              // dt^2 / 12
    float vp2;               // scalar vp^2 (1/m) when param_kind == SCALAR
    // Born (reference `Born` operator, acoustic/operators.py:235-277): the linearised field's new
    // time level and the model perturbation dm (own strides / halo width)
    float *__restrict__ U1;
    const float *__restrict__ dm;
    long long dsx, dsy;
    int dg0, dg1, dg2;       // index in dm of the first iterated point (grid index + dm's halo width)
};

#define B2_HD __host__ __device__ __forceinline__

B2_HD long long iso_index(const IsoGK &k, int x, int y, int z) {
    return (long long)(k.o0 + x) * k.sx + (long long)(k.o1 + y) * k.sy + (k.o2 + z);
}

// plain star Laplacian of `f` at idx (weights include 1/h^2)
B2_HD float iso_star(const IsoGK &k, const float *__restrict__ f, long long idx) {
    const float c = f[idx];
    float acc = (k.w[0][0] + k.w[1][0] + k.w[2][0]) * c;
    for (int i = 1; i <= k.r0; ++i)
        acc += k.w[0][i] * (f[idx - i * k.sx] + f[idx + i * k.sx]);
    for (int i = 1; i <= k.r1; ++i)
        acc += k.w[1][i] * (f[idx - i * k.sy] + f[idx + i * k.sy]);
    for (int i = 1; i <= k.r2; ++i)
        acc += k.w[2][i] * (f[idx - i] + f[idx + i]);
    return acc;
}

// u+ = ( m/dt^2 (2u - u-) + damp/dt u + H ) / ( m/dt^2 + damp/dt )
B2_HD float iso_advance(const IsoGK &k, long long idx, float c, float H) {
    float m_dt2 = k.m_dt2;
    if (k.param_kind == B2_PARAM_VP) {
        const float v = k.param[idx];
        m_dt2 = k.inv_dt2 / (v * v);
    } else if (k.param_kind == B2_PARAM_M) {
        m_dt2 = k.param[idx] * k.inv_dt2;
    }
    const float d = k.damp ? k.damp[idx] * k.inv_dt : 0.f;
    const float num = m_dt2 * (2.f * c - k.um[idx]) + d * c + H;
    return num / (m_dt2 + d);
}

B2_HD void iso_point(const IsoGK &k, int x, int y, int z) {
    const long long idx = iso_index(k, x, y, z);
    float H = iso_star(k, k.u0, idx);
    if (k.W) H += k.ot4c * iso_star(k, k.W, idx);
    k.u1[idx] = iso_advance(k, idx, k.u0[idx], H);
}

// OT4 first pass, run over the iteration box grown by the radius: W = lap(u) / m at this point
B2_HD void ot4_w_point(const IsoGK &k, int x, int y, int z) {
    const long long idx = iso_index(k, x, y, z);
    float minv = k.vp2;
    if (k.param_kind == B2_PARAM_VP) {
        const float v = k.param[idx];
        minv = v * v;
    } else if (k.param_kind == B2_PARAM_M) {
        minv = 1.0f / k.param[idx];
    }
    k.W[idx] = iso_star(k, k.u0, idx) * minv;
}

// Born source: eqn2 of the reference's `Born` operator is the plain update of U with
// q = -dm * u.dt2 added to the numerator, u.dt2 = (u[t+1] - 2 u[t] + u[t-1]) / dt^2 taken AFTER the
// source was injected into u[t+1]. The plain update of U has already run (same kernels as u), so
// this adds q / (m/dt^2 + damp/dt). Here u0/um/u1 are the time levels of u.
B2_HD void born_src_point(const IsoGK &k, int x, int y, int z) {
    const long long idx = iso_index(k, x, y, z);
    const long long j = (long long)(k.dg0 + x) * k.dsx + (long long)(k.dg1 + y) * k.dsy + (k.dg2 + z);
    float m_dt2 = k.m_dt2;
    if (k.param_kind == B2_PARAM_VP) {
        const float v = k.param[idx];
        m_dt2 = k.inv_dt2 / (v * v);
    } else if (k.param_kind == B2_PARAM_M) {
        m_dt2 = k.param[idx] * k.inv_dt2;
    }
    const float d = k.damp ? k.damp[idx] * k.inv_dt : 0.f;
    const float q = -k.dm[j] * ((k.u1[idx] - 2.f * k.u0[idx] + k.um[idx]) * k.inv_dt2);
    k.U1[idx] += q / (m_dt2 + d);
}

// Snapshot (time-subsampled saving, reference examples/seismic/tutorials/08_snapshotting.ipynb:455-505:
// `Eq(usave, u)` on a ConditionalDimension -> `if (time % factor == 0) usave[time/factor][..] = u[t0][..]`):
// copy of the iteration box between two arrays with different halo widths.
struct SnapK {
    const float *__restrict__ src;
    float *__restrict__ dst;
    long long sx, sy, dsx, dsy;      // strides of the wavefield slot / of one snapshot
    int n0, n1, n2;
    int o0, o1, o2;                  // index of the first point in the wavefield slot
    int d0, d1, d2;                  // index of the first point in the snapshot
};

B2_HD void snapshot_point(const SnapK &k, int x, int y, int z) {
    k.dst[(long long)(k.d0 + x) * k.dsx + (long long)(k.d1 + y) * k.dsy + (k.d2 + z)] =
        k.src[(long long)(k.o0 + x) * k.sx + (long long)(k.o1 + y) * k.sy + (k.o2 + z)];
}

// Generic constant-coefficient explicit update (b2_linear_forward): out[p] = sum_k coef_k * in_k[p + d_k]
struct LinK {
    float *__restrict__ out;
    const float *__restrict__ lvl[4];    // base pointers of the distinct time levels read
    long long sx, sy;
    int n0, n1, n2, o0, o1, o2;
    int ntaps;
    int sel[B2_MAX_TAPS];                // which level each tap reads
    long long delta[B2_MAX_TAPS];        // linear index offset of each tap
    float coef[B2_MAX_TAPS];
};

B2_HD void linear_point(const LinK &k, int x, int y, int z) {
    const long long idx = (long long)(k.o0 + x) * k.sx + (long long)(k.o1 + y) * k.sy + (k.o2 + z);
    float acc = 0.f;
    for (int i = 0; i < k.ntaps; ++i) acc += k.coef[i] * k.lvl[k.sel[i]][idx + k.delta[i]];
    k.out[idx] = acc;
}

// Free surface on the low side of the last dimension (reference `freesurface`,
// examples/seismic/acoustic/operators.py:5-47; generated form `r1[z]*u[t0][..][4 + abs(z - 1)]`,
// `u[t2][x][y][4] = 0`): rows z <= radius redone with the vertical taps that reach z - k <= 0
// replaced by sign(z - k) * u[|z - k|] (so a tap landing on z = 0 contributes 0); row 0 cleared.
B2_HD void iso_fs_point(const IsoGK &k, int x, int y, int z) {
    const long long idx = iso_index(k, x, y, z);
    if (z == 0) { k.u1[idx] = 0.f; return; }
    const float c = k.u0[idx];
    float acc = (k.w[0][0] + k.w[1][0] + k.w[2][0]) * c;
    for (int i = 1; i <= k.r0; ++i)
        acc += k.w[0][i] * (k.u0[idx - i * k.sx] + k.u0[idx + i * k.sx]);
    for (int i = 1; i <= k.r1; ++i)
        acc += k.w[1][i] * (k.u0[idx - i * k.sy] + k.u0[idx + i * k.sy]);
    for (int i = 1; i <= k.r2; ++i) {
        float lo;
        if (z - i > 0) lo = k.u0[idx - i];
        else if (z - i < 0) lo = -k.u0[idx - z + (i - z)];
        else lo = 0.f;
        acc += k.w[2][i] * (lo + k.u0[idx + i]);
    }
    k.u1[idx] = iso_advance(k, idx, c, acc);
}

}  // namespace b2
