"""Shared problem set-up for the parity tests: rebuilds the inputs of a golden fixture with the
ORACLE's host-side restatements (oracle/oracle.py), independent of the product package."""
import os

import numpy as np

from oracle import oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + '.npz')))


def rel_linf(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def domain(arr, so):
    sl = (slice(None),) + tuple(slice(so, s - so) for s in arr.shape[1:])
    return arr[sl]


def iso_problem(n, nbl, so, tn, f0=0.010, vp=1.5, h=10.0, interpolation='linear', r=None,
                nrec_grid=True, vp_array=None):
    """Inputs of `demo_model('constant-isotropic') + setup_geometry` restated with the oracle."""
    N = n + 2 * nbl
    spacing = (np.float32(h),) * 3
    origin = tuple(np.float32(-nbl * h) for _ in range(3))
    vp_max = float(vp if vp_array is None else vp_array.max())
    dt = float(O.critical_dt(so, 3, h, vp_max))
    nt, tvals = O.time_axis(0.0, tn, dt)
    damp = O.damp_field((N, N, N), nbl, spacing, so)
    dom = tuple((n - 1) * h for _ in range(3))
    src_c = np.array([[dom[0] * .5, dom[1] * .5, h]])
    rx = np.linspace(0, dom[0], n)
    ry = np.linspace(0, dom[1], n)
    rec_c = np.empty((n * n, 3))
    rec_c[:, 0] = np.repeat(rx, n)
    rec_c[:, 1] = np.tile(ry, n)
    rec_c[:, 2] = 2 * h
    rr = r or (1 if interpolation == 'linear' else 4)
    sgp, sw = O.tabulate(src_c.astype(np.float32), origin, spacing, rr, interpolation)
    rgp, rw = O.tabulate(rec_c.astype(np.float32), origin, spacing, rr, interpolation)
    src = np.zeros((nt, 1), dtype=np.float32)
    src[:, 0] = O.ricker(f0, tvals)
    rec = np.zeros((nt, n * n), dtype=np.float32)
    w = [O.fd2_weights(so, h)] * 3
    u = np.zeros((3, N + 2 * so, N + 2 * so, N + 2 * so), dtype=np.float32)
    return dict(N=N, so=so, dt=dt, nt=nt, damp=damp, w=w, u=u, vp=vp,
                src=dict(data=src, gp=sgp, w=sw, r=rr), rec=dict(data=rec, gp=rgp, w=rw, r=rr),
                src_coords=src_c, rec_coords=rec_c, origin=origin, spacing=spacing)


def fs_problem(n, nbl, so, tn, h=10.0, nlayers=3, interpolation='linear', f0=0.010):
    """Inputs of `demo_model('layers-isotropic', fs=True) + setup_geometry` restated with the oracle:
    free surface on top of z (no absorbing layer there, origin_z unchanged;
    examples/seismic/model.py:113-131, :166-172), layered velocity edge-padded into the layers."""
    shape = (n, n, n)
    N = (n + 2 * nbl, n + 2 * nbl, n + nbl)
    spacing = (np.float32(h),) * 3
    origin = (np.float32(-nbl * h), np.float32(-nbl * h), np.float32(0.0))
    vp_phys = O.layered_vp(shape, nlayers)
    vp_dom = np.pad(vp_phys, ((nbl, nbl), (nbl, nbl), (0, nbl)), mode='edge')
    dt = float(O.critical_dt(so, 3, h, float(vp_phys.max())))
    nt, tvals = O.time_axis(0.0, tn, dt)
    damp = O.damp_field(N, nbl, spacing, so, fs=True)
    dom = tuple((n - 1) * h for _ in range(3))
    src_c = np.array([[dom[0] * .5, dom[1] * .5, h]])
    rx = np.linspace(0, dom[0], n)
    ry = np.linspace(0, dom[1], n)
    rec_c = np.empty((n * n, 3))
    rec_c[:, 0] = np.repeat(rx, n)
    rec_c[:, 1] = np.tile(ry, n)
    rec_c[:, 2] = 2 * h
    rr = 1 if interpolation == 'linear' else 4
    sgp, sw = O.tabulate(src_c.astype(np.float32), origin, spacing, rr, interpolation)
    rgp, rw = O.tabulate(rec_c.astype(np.float32), origin, spacing, rr, interpolation)
    src = np.zeros((nt, 1), dtype=np.float32)
    src[:, 0] = O.ricker(f0, tvals)
    rec = np.zeros((nt, n * n), dtype=np.float32)
    w = [O.fd2_weights(so, h)] * 3
    u = np.zeros((3,) + tuple(s + 2 * so for s in N), dtype=np.float32)
    return dict(N=N, so=so, dt=dt, nt=nt, damp=damp, w=w, u=u, vp_dom=vp_dom,
                vp=np.pad(vp_dom, so, mode='edge').astype(np.float32),
                src=dict(data=src, gp=sgp, w=sw, r=rr), rec=dict(data=rec, gp=rgp, w=rw, r=rr),
                src_coords=src_c, rec_coords=rec_c, origin=origin, spacing=spacing)


def varying_tti_parameters(n):
    """vp, epsilon, delta, theta, phi of the fixture `tti3d_so8_varying` on the physical n^3 grid (the formulas of
    oracle/make_golden.py::tti_varying): every parameter varies along every axis."""
    gx, gy, gz = np.meshgrid(*[np.linspace(0., 1., n, dtype=np.float32)] * 3, indexing='ij')
    return dict(vp=(1.5 + 0.6 * gx + 0.5 * gy + 0.9 * gz).astype(np.float32),
                epsilon=(0.25 * gx * gz + 0.05).astype(np.float32), delta=(0.12 * gy + 0.02).astype(np.float32),
                theta=(0.2 + 0.9 * gx * gy + 0.3 * gz).astype(np.float32),
                phi=(0.1 + 0.8 * gy * gz - 0.4 * gx).astype(np.float32))
