/*
 * oracle.c — CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may call this.  The product path (devito_b200 + libb200stencil.so) never does.
 *
 * Parity is PINNED: this restatement is checked (tests/test_oracle_golden.py) against
 *   (a) golden vectors produced by running the reference itself (devito @ 436199c, CPU
 *       OpenMP backend, gcc) in the build container — see oracle/make_golden.py, fixtures
 *       in tests/golden/*.npz;
 *   (b) the reference's own known-answer test norm(rec) = 490.56 +- 1e-2
 *       (tests/test_gpu_openacc.py:205-251).
 *
 * What it restates (all citations relative to the reference tree):
 *   iso update      examples/seismic/acoustic/operators.py:71-107 (iso_stencil), :50-68
 *                   (laplacian); generated form `Forward` section0
 *   TTI update      examples/seismic/tti/operators.py:65-104 (Gzz_centered), :146-183
 *                   (Gh_centered), :186-247 (kernel_centered), :12-39 (second_order_stencil)
 *   injection       devito/operations/interpolators.py:553-624 (_inject), guards :284-311
 *   interpolation   devito/operations/interpolators.py:510-551 (_interpolate)
 *   time loop       slot rotation t0 = time%T, t1 = (time+1)%T, t2 = (time-1)%T
 *                   (devito/ir/clusters/algorithms.py:321-427; printed in
 *                   examples/seismic/tutorials/08_snapshotting.ipynb:473)
 * The arithmetic is written in the operation order of the C code the reference generates
 * (float32 throughout), so that with DEVITO_SAFE_MATH=1 the two agree to rounding.
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#define IDX3(x, y, z) ((size_t)(x) * sx + (size_t)(y) * sy + (size_t)(z))

typedef struct {
    float *data;       /* (nt, npoint) */
    const int *gp;     /* (npoint, ndim) */
    const float *w[3]; /* (npoint, 2r) */
    int nt, npoint, p_m, p_M, r;
} osparse;

/* sparse ops on one field slot; ndim 2 or 3; lo/hi inclusive iteration bounds */
static void inject(const osparse *s, int ndim, float *f0, float *f1, size_t sx, size_t sy, int so,
                   const int *lo, const int *hi, int time, int param_kind, const float *param,
                   float vp, float dt) {
    if (!s || time < 0 || time >= s->nt) return;
    const int n = 2 * s->r, r = s->r;
    for (int p = s->p_m; p <= s->p_M; ++p) {
        const float sv = s->data[(size_t)time * s->npoint + p];
        if (ndim == 3) {
            for (int a = 0; a < n; ++a)
                for (int b = 0; b < n; ++b)
                    for (int c = 0; c < n; ++c) {
                        const int cx = s->gp[p * 3 + 0] + a - r + 1;
                        const int cy = s->gp[p * 3 + 1] + b - r + 1;
                        const int cz = s->gp[p * 3 + 2] + c - r + 1;
                        if (cx < lo[0] - r || cx > hi[0] + r || cy < lo[1] - r || cy > hi[1] + r ||
                            cz < lo[2] - r || cz > hi[2] + r)
                            continue;
                        const size_t i = IDX3(cx + so, cy + so, cz + so);
                        float scale;
                        if (param_kind == 0) scale = (vp * vp) * (dt * dt);
                        else if (param_kind == 1) scale = (param[i] * param[i]) * (dt * dt);
                        else scale = (dt * dt) / param[i];
                        const float val = scale * s->w[0][p * n + a] * s->w[1][p * n + b] *
                                          s->w[2][p * n + c] * sv;
                        f0[i] += val;
                        if (f1) f1[i] += val;
                    }
        } else {
            for (int a = 0; a < n; ++a)
                for (int b = 0; b < n; ++b) {
                    const int cx = s->gp[p * 2 + 0] + a - r + 1;
                    const int cy = s->gp[p * 2 + 1] + b - r + 1;
                    if (cx < lo[0] - r || cx > hi[0] + r || cy < lo[1] - r || cy > hi[1] + r) continue;
                    const size_t i = (size_t)(cx + so) * sy + (size_t)(cy + so);
                    float scale;
                    if (param_kind == 0) scale = (vp * vp) * (dt * dt);
                    else if (param_kind == 1) scale = (param[i] * param[i]) * (dt * dt);
                    else scale = (dt * dt) / param[i];
                    f0[i] += scale * s->w[0][p * n + a] * s->w[1][p * n + b] * sv;
                }
        }
    }
}

static void interp(const osparse *s, int ndim, const float *f0, const float *f1, size_t sx, size_t sy,
                   int so, const int *lo, const int *hi, int time) {
    if (!s || time < 0 || time >= s->nt) return;
    const int n = 2 * s->r, r = s->r;
#pragma omp parallel for schedule(static)
    for (int p = s->p_m; p <= s->p_M; ++p) {
        float sum = 0.0f;
        if (ndim == 3) {
            for (int a = 0; a < n; ++a)
                for (int b = 0; b < n; ++b)
                    for (int c = 0; c < n; ++c) {
                        const int cx = s->gp[p * 3 + 0] + a - r + 1;
                        const int cy = s->gp[p * 3 + 1] + b - r + 1;
                        const int cz = s->gp[p * 3 + 2] + c - r + 1;
                        if (cx < lo[0] - r || cx > hi[0] + r || cy < lo[1] - r || cy > hi[1] + r ||
                            cz < lo[2] - r || cz > hi[2] + r)
                            continue;
                        const size_t i = IDX3(cx + so, cy + so, cz + so);
                        float v = f0[i];
                        if (f1) v += f1[i];
                        sum += s->w[0][p * n + a] * s->w[1][p * n + b] * s->w[2][p * n + c] * v;
                    }
        } else {
            for (int a = 0; a < n; ++a)
                for (int b = 0; b < n; ++b) {
                    const int cx = s->gp[p * 2 + 0] + a - r + 1;
                    const int cy = s->gp[p * 2 + 1] + b - r + 1;
                    if (cx < lo[0] - r || cx > hi[0] + r || cy < lo[1] - r || cy > hi[1] + r) continue;
                    const size_t i = (size_t)(cx + so) * sy + (size_t)(cy + so);
                    sum += s->w[0][p * n + a] * s->w[1][p * n + b] * f0[i];
                }
        }
        s->data[(size_t)time * s->npoint + p] = sum;
    }
}

/* ------------------------------------------------------------------------------------------
 * isotropic acoustic forward.  u: (tsize, [ax,] ay, az) with halo `so`;  w[d][0..radius].
 * `alloc` = allocated extents per space dim; lo/hi = inclusive bounds per space dim.
 * ---------------------------------------------------------------------------------------- */
int oracle_iso_forward(int ndim, float *u, int tsize, const int *alloc, int so, int radius,
                       const float *wx, const float *wy, const float *wz, const float *damp,
                       int param_kind, const float *param, float vp, float dt, const int *lo,
                       const int *hi, int time_m, int time_M, osparse *src, osparse *rec,
                       int rec_toff, int adjoint, float *grad, const int *galloc, int ghalo,
                       const float *usave, int free_surface, int ot4) {
    /* ot4 != 0: the reference's 4th-order-in-time kernel (acoustic/operators.py:50-68, kernel='OT4'):
     * H = lap(u) + dt^2/12 * lap( lap(u) / m ), with the inner Laplacian and 1/m sampled at the shifted
     * point (devito/finite_differences/differentiable.py:442-449 `biharmonic`). 3-D, radius <= so/2. */
    /* free_surface != 0: the reference's `freesurface` (acoustic/operators.py:5-47) on the LAST
     * dimension: taps of the vertical derivative that fall above the surface, z - k < 0, read
     * sign(z-k) * u[|z-k|] (antisymmetric mirror; a tap landing exactly on z = 0 contributes 0), and
     * after the update u[t+1][.., z=0] = 0 (before injection). Halo cells above the surface are never
     * read by the stencil. Requires the iteration to start at z = 0. */
    /* grad/usave != NULL: imaging condition of the reference's `Gradient` operator
     * (acoustic/operators.py:222): after each step grad -= usave[time] * u.dt2 (3-D only).
     * adjoint != 0: the reference's `Adjoint` operator (acoustic/operators.py:153-187) — the same
     * update with the roles of t+1 / t-1 exchanged, time running from time_M down to time_m */
    const int R = radius;
    size_t sx, sy, slot;
    if (ndim == 3) {
        sy = (size_t)alloc[2];
        sx = (size_t)alloc[1] * alloc[2];
        slot = (size_t)alloc[0] * sx;
    } else {
        sy = (size_t)alloc[1];
        sx = 0;
        slot = (size_t)alloc[0] * sy;
    }
    const float r2 = 1.0f / (dt * dt);
    const float r3 = 1.0f / dt;
    const float r1s = 1.0f / (vp * vp);
    const int dir = adjoint ? -1 : 1;
    float *W = NULL;                      /* OT4: lap(u)/m on the iteration box grown by R */
    if (ot4) {
        if (ndim != 3 || free_surface || 2 * R > so) return 2;
        W = (float *)calloc(slot, sizeof(float));
        if (!W) return 3;
    }
    const float ot4c = dt * dt / 12.0f;
    for (int time = adjoint ? time_M : time_m; adjoint ? time >= time_m : time <= time_M; time += dir) {
        const int t0 = ((time % tsize) + tsize) % tsize;
        const int t1 = (((time + dir) % tsize) + tsize) % tsize;
        const int t2 = (((time - dir) % tsize) + tsize) % tsize;
        const float *u0 = u + (size_t)t0 * slot;
        const float *um = u + (size_t)t2 * slot;
        float *u1 = u + (size_t)t1 * slot;
        if (ot4) {
#pragma omp parallel for collapse(2) schedule(static)
            for (int x = lo[0] - R; x <= hi[0] + R; ++x)
                for (int y = lo[1] - R; y <= hi[1] + R; ++y)
                    for (int z = lo[2] - R; z <= hi[2] + R; ++z) {
                        const size_t i = IDX3(x + so, y + so, z + so);
                        float l = (wx[0] + wy[0] + wz[0]) * u0[i];
                        for (int k = 1; k <= R; ++k)
                            l += wx[k] * (u0[i - k * sx] + u0[i + k * sx]) +
                                 wy[k] * (u0[i - k * sy] + u0[i + k * sy]) + wz[k] * (u0[i - k] + u0[i + k]);
                        float minv = vp * vp;
                        if (param_kind == 1) minv = param[i] * param[i];
                        else if (param_kind == 2) minv = 1.0f / param[i];
                        W[i] = l * minv;
                    }
        }
        if (ndim == 3) {
#pragma omp parallel for collapse(2) schedule(static)
            for (int x = lo[0]; x <= hi[0]; ++x)
                for (int y = lo[1]; y <= hi[1]; ++y)
                    for (int z = lo[2]; z <= hi[2]; ++z) {
                        const size_t i = IDX3(x + so, y + so, z + so);
                        float lap = (wx[0] + wy[0] + wz[0]) * u0[i];
                        if (ot4) {
                            float bl = (wx[0] + wy[0] + wz[0]) * W[i];
                            for (int k = 1; k <= R; ++k)
                                bl += wx[k] * (W[i - k * sx] + W[i + k * sx]) +
                                      wy[k] * (W[i - k * sy] + W[i + k * sy]) + wz[k] * (W[i - k] + W[i + k]);
                            lap += ot4c * bl;
                        }
                        for (int k = 1; k <= R; ++k) {
                            float zlo = u0[i - k];
                            if (free_surface && z - k <= 0) zlo = (z - k < 0) ? -u0[i - z + (k - z)] : 0.0f;
                            lap += wx[k] * (u0[i - k * sx] + u0[i + k * sx]) +
                                   wy[k] * (u0[i - k * sy] + u0[i + k * sy]) +
                                   wz[k] * (zlo + u0[i + k]);
                        }
                        float r1 = r1s;
                        if (param_kind == 1) r1 = 1.0f / (param[i] * param[i]);
                        else if (param_kind == 2) r1 = param[i];
                        const float d = damp ? damp[i] : 0.0f;
                        u1[i] = (-r1 * (-2.0f * r2 * u0[i] + r2 * um[i]) + r3 * d * u0[i] + lap) /
                                (r1 * r2 + r3 * d);
                    }
        } else {
#pragma omp parallel for schedule(static)
            for (int x = lo[0]; x <= hi[0]; ++x)
                for (int y = lo[1]; y <= hi[1]; ++y) {
                    const size_t i = (size_t)(x + so) * sy + (size_t)(y + so);
                    float lap = (wx[0] + wy[0]) * u0[i];
                    for (int k = 1; k <= R; ++k) {
                        float zlo = u0[i - k];
                        if (free_surface && y - k <= 0) zlo = (y - k < 0) ? -u0[i - y + (k - y)] : 0.0f;
                        lap += wx[k] * (u0[i - k * sy] + u0[i + k * sy]) + wy[k] * (zlo + u0[i + k]);
                    }
                    float r1 = r1s;
                    if (param_kind == 1) r1 = 1.0f / (param[i] * param[i]);
                    else if (param_kind == 2) r1 = param[i];
                    const float d = damp ? damp[i] : 0.0f;
                    u1[i] = (-r1 * (-2.0f * r2 * u0[i] + r2 * um[i]) + r3 * d * u0[i] + lap) /
                            (r1 * r2 + r3 * d);
                }
        }
        if (free_surface) {
            if (lo[ndim - 1] != 0) return 1;
            if (ndim == 3) {
                for (int x = lo[0]; x <= hi[0]; ++x)
                    for (int y = lo[1]; y <= hi[1]; ++y) u1[IDX3(x + so, y + so, so)] = 0.0f;
            } else {
                for (int x = lo[0]; x <= hi[0]; ++x) u1[(size_t)(x + so) * sy + (size_t)so] = 0.0f;
            }
        }
        inject(src, ndim, u1, NULL, sx, sy, so, lo, hi, time, param_kind, param, vp, dt);
        interp(rec, ndim, rec_toff ? u1 : u0, NULL, sx, sy, so, lo, hi, time);
        if (grad && usave && ndim == 3) {
            const float *us = usave + (size_t)time * slot;
            const size_t gsy = (size_t)galloc[2], gsx = (size_t)galloc[1] * galloc[2];
#pragma omp parallel for collapse(2) schedule(static)
            for (int x = lo[0]; x <= hi[0]; ++x)
                for (int y = lo[1]; y <= hi[1]; ++y)
                    for (int z = lo[2]; z <= hi[2]; ++z) {
                        const size_t i = IDX3(x + so, y + so, z + so);
                        const size_t g = (size_t)(x + ghalo) * gsx + (size_t)(y + ghalo) * gsy + (z + ghalo);
                        grad[g] += -us[i] * (r2 * u1[i] - 2.0f * r2 * u0[i] + r2 * um[i]);
                    }
        }
    }
    free(W);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Linearised (Born) modelling: the reference's `Born` operator, examples/seismic/acoustic/
 * operators.py:235-277. Per time step, in the order of its equation list:
 *     u[t+1]  = update(u)                       (eqn1)
 *     u[t+1] += inject(src[time])               (source)
 *     U[t+1]  = update(U) with the extra source q = -dm * u.dt2 in the numerator, where
 *               u.dt2 = (u[t+1] - 2 u[t] + u[t-1]) / dt^2 uses the u[t+1] just computed (eqn2)
 *     rec[time] = interpolate(U[t])             (receivers)
 * 3-D, OT2, no free surface. `dm` has its own halo width (the reference builds it with
 * space_order 0).
 * ---------------------------------------------------------------------------------------- */
int oracle_born_forward(float *u, float *U, int tsize, const int *alloc, int so, int radius,
                        const float *wx, const float *wy, const float *wz, const float *damp,
                        int param_kind, const float *param, float vp, float dt, const int *lo,
                        const int *hi, int time_m, int time_M, osparse *src, osparse *rec,
                        const float *dm, const int *dmalloc, int dmhalo) {
    const int R = radius;
    const size_t sy = (size_t)alloc[2], sx = (size_t)alloc[1] * alloc[2];
    const size_t slot = (size_t)alloc[0] * sx;
    const size_t dsy = (size_t)dmalloc[2], dsx = (size_t)dmalloc[1] * dmalloc[2];
    const float r2 = 1.0f / (dt * dt);
    const float r3 = 1.0f / dt;
    const float r1s = 1.0f / (vp * vp);
    for (int time = time_m; time <= time_M; ++time) {
        const int t0 = ((time % tsize) + tsize) % tsize;
        const int t1 = (((time + 1) % tsize) + tsize) % tsize;
        const int t2 = (((time - 1) % tsize) + tsize) % tsize;
        for (int pass = 0; pass < 2; ++pass) {
            float *f = pass == 0 ? u : U;
            const float *f0 = f + (size_t)t0 * slot, *fm = f + (size_t)t2 * slot;
            float *f1 = f + (size_t)t1 * slot;
            const float *u0 = u + (size_t)t0 * slot, *um = u + (size_t)t2 * slot, *u1 = u + (size_t)t1 * slot;
#pragma omp parallel for collapse(2) schedule(static)
            for (int x = lo[0]; x <= hi[0]; ++x)
                for (int y = lo[1]; y <= hi[1]; ++y)
                    for (int z = lo[2]; z <= hi[2]; ++z) {
                        const size_t i = IDX3(x + so, y + so, z + so);
                        float lap = (wx[0] + wy[0] + wz[0]) * f0[i];
                        for (int k = 1; k <= R; ++k)
                            lap += wx[k] * (f0[i - k * sx] + f0[i + k * sx]) +
                                   wy[k] * (f0[i - k * sy] + f0[i + k * sy]) +
                                   wz[k] * (f0[i - k] + f0[i + k]);
                        float r1 = r1s;
                        if (param_kind == 1) r1 = 1.0f / (param[i] * param[i]);
                        else if (param_kind == 2) r1 = param[i];
                        const float d = damp ? damp[i] : 0.0f;
                        float q = 0.0f;
                        if (pass == 1) {
                            const size_t j = (size_t)(x + dmhalo) * dsx + (size_t)(y + dmhalo) * dsy + (z + dmhalo);
                            q = -dm[j] * (r2 * u1[i] - 2.0f * r2 * u0[i] + r2 * um[i]);
                        }
                        f1[i] = (-r1 * (-2.0f * r2 * f0[i] + r2 * fm[i]) + r3 * d * f0[i] + lap + q) /
                                (r1 * r2 + r3 * d);
                    }
            if (pass == 0)
                inject(src, 3, f1, NULL, sx, sy, so, lo, hi, time, param_kind, param, vp, dt);
        }
        interp(rec, 3, U + (size_t)t0 * slot, NULL, sx, sy, so, lo, hi, time);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * TTI centred forward, scalar parameters, 3-D.  w2[d][0..R] second-derivative weights,
 * w1[d][0..R-1] half-node first-derivative weights (offsets -R/2+1..R/2 about x+h/2).
 * Follows the generated `ForwardTTI`: first the rotated first derivatives Gz(u), Gz(v) on
 * the box extended by [-R/2, R/2-1] (reference: CIRE temporaries r12..r17), then the update.
 * ---------------------------------------------------------------------------------------- */
int oracle_tti_forward(float *u, float *v, int tsize, const int *alloc, int so, int radius,
                       const float *w2x, const float *w2y, const float *w2z, const float *w1x,
                       const float *w1y, const float *w1z, const float *damp, float vp,
                       float epsilon, float delta, float theta, float phi, float dt, const int *lo,
                       const int *hi, int time_m, int time_M, osparse *src, osparse *rec,
                       int rec_toff, const float *vp_a, const float *eps_a, const float *delta_a,
                       const float *theta_a, const float *phi_a) {
    /* *_a: array-valued parameters (NULL -> scalar), preset `layers-tti`. Rotation factors are
     * sampled at the Gz point inside Gz and at the shifted point in the outer derivative, exactly
     * like the reference's generated code (time-invariant tables r2..r5). */
    const int R = radius, h = radius / 2;
    const size_t sy = (size_t)alloc[2], sx = (size_t)alloc[1] * alloc[2];
    const size_t slot = (size_t)alloc[0] * sx;
    float *gzu = (float *)calloc(slot, sizeof(float));
    float *gzv = (float *)calloc(slot, sizeof(float));
    if (!gzu || !gzv) return 202;
    const int arrays = vp_a || eps_a || delta_a || theta_a || phi_a;
    float *tcx = NULL, *tcy = NULL, *tcz = NULL, *tsd = NULL;
    if (arrays) {
        tcx = (float *)malloc(slot * sizeof(float));
        tcy = (float *)malloc(slot * sizeof(float));
        tcz = (float *)malloc(slot * sizeof(float));
        tsd = (float *)malloc(slot * sizeof(float));
        for (size_t i = 0; i < slot; ++i) {
            const float th = theta_a ? theta_a[i] : theta, ph = phi_a ? phi_a[i] : phi;
            const float s_ = sinf(th);
            tcz[i] = cosf(th);
            tcy[i] = s_ * sinf(ph);
            tcx[i] = s_ * cosf(ph);
            tsd[i] = sqrtf(2 * (delta_a ? delta_a[i] : delta) + 1);
        }
    }
    const float r18s = sqrtf(2 * delta + 1);
    const float ct = cosf(theta), st = sinf(theta), sp = sinf(phi), cp = cosf(phi);
    const float czs = ct, cys = sp * st, cxs = st * cp;
    const float e2s = 2 * epsilon + 1;
    const float r9s = 1.0f / (vp * vp), r10 = 1.0f / (dt * dt), r11 = 1.0f / dt;
    for (int time = time_m; time <= time_M; ++time) {
        const int t0 = ((time % tsize) + tsize) % tsize;
        const int t1 = (((time + 1) % tsize) + tsize) % tsize;
        const int t2 = (((time - 1) % tsize) + tsize) % tsize;
        const float *u0 = u + (size_t)t0 * slot, *v0 = v + (size_t)t0 * slot;
        const float *um = u + (size_t)t2 * slot, *vm = v + (size_t)t2 * slot;
        float *u1 = u + (size_t)t1 * slot, *v1 = v + (size_t)t1 * slot;
#pragma omp parallel for collapse(2) schedule(static)
        for (int x = lo[0] - h; x <= hi[0] + h - 1; ++x)
            for (int y = lo[1] - h; y <= hi[1] + h - 1; ++y)
                for (int z = lo[2] - h; z <= hi[2] + h - 1; ++z) {
                    const size_t i = IDX3(x + so, y + so, z + so);
                    float dxu = 0, dyu = 0, dzu = 0, dxv = 0, dyv = 0, dzv = 0;
                    for (int j = 0; j < R; ++j) {
                        const ptrdiff_t o = j - h + 1;
                        dxu += w1x[j] * u0[i + o * (ptrdiff_t)sx];
                        dyu += w1y[j] * u0[i + o * (ptrdiff_t)sy];
                        dzu += w1z[j] * u0[i + o];
                        dxv += w1x[j] * v0[i + o * (ptrdiff_t)sx];
                        dyv += w1y[j] * v0[i + o * (ptrdiff_t)sy];
                        dzv += w1z[j] * v0[i + o];
                    }
                    const float cx = arrays ? tcx[i] : cxs, cy = arrays ? tcy[i] : cys, cz = arrays ? tcz[i] : czs;
                    gzu[i] = cz * dzu + cy * dyu + cx * dxu;
                    gzv[i] = cz * dzv + cy * dyv + cx * dxv;
                }
#pragma omp parallel for collapse(2) schedule(static)
        for (int x = lo[0]; x <= hi[0]; ++x)
            for (int y = lo[1]; y <= hi[1]; ++y)
                for (int z = lo[2]; z <= hi[2]; ++z) {
                    const size_t i = IDX3(x + so, y + so, z + so);
                    float lap = (w2x[0] + w2y[0] + w2z[0]) * u0[i];
                    for (int k = 1; k <= R; ++k)
                        lap += w2x[k] * (u0[i - k * sx] + u0[i + k * sx]) +
                               w2y[k] * (u0[i - k * sy] + u0[i + k * sy]) +
                               w2z[k] * (u0[i - k] + u0[i + k]);
                    /* outer half-node derivatives of (r12,r13,r14) and (r15,r16,r17) */
                    float H0 = 0, Hz = 0;
                    const float r18 = arrays ? tsd[i] : r18s;
                    const float e2 = eps_a ? 2 * eps_a[i] + 1 : e2s;
                    const float r9 = vp_a ? 1.0f / (vp_a[i] * vp_a[i]) : r9s;
                    for (int j = 0; j < R; ++j) {
                        const ptrdiff_t o = j - h;
                        const size_t iz = i + o, iy = i + o * (ptrdiff_t)sy, ix = i + o * (ptrdiff_t)sx;
                        const float cz = arrays ? tcz[iz] : czs, cy = arrays ? tcy[iy] : cys,
                                    cx = arrays ? tcx[ix] : cxs;
                        H0 += w1z[j] * (gzv[iz] * (r18 * cz) - gzu[iz] * e2 * cz) +
                              w1x[j] * (gzv[ix] * (r18 * cx) - gzu[ix] * e2 * cx) +
                              w1y[j] * (gzv[iy] * (r18 * cy) - gzu[iy] * e2 * cy);
                        Hz += w1z[j] * (gzv[iz] * cz - gzu[iz] * (r18 * cz)) +
                              w1y[j] * (gzv[iy] * cy - gzu[iy] * (r18 * cy)) +
                              w1x[j] * (gzv[ix] * cx - gzu[ix] * (r18 * cx));
                    }
                    const float d = damp ? damp[i] : 0.0f;
                    const float r26 = 1.0f / (r10 * r9 + r11 * d);
                    u1[i] = r26 * (e2 * lap + r10 * r9 * (2.0f * u0[i] - um[i]) + r11 * d * u0[i] + H0);
                    v1[i] = r26 * (r11 * d * v0[i] + r18 * lap - r9 * (-2.0f * r10 * v0[i] + r10 * vm[i]) + Hz);
                }
        inject(src, 3, u1, v1, sx, sy, so, lo, hi, time, vp_a ? 1 : 0, vp_a, vp, dt);
        interp(rec, 3, rec_toff ? u1 : u0, rec_toff ? v1 : v0, sx, sy, so, lo, hi, time);
    }
    free(gzu);
    free(gzv);
    free(tcx); free(tcy); free(tcz); free(tsd);
    return 0;
}
