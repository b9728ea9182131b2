"""Generate golden vectors by running the REFERENCE itself (devito @ /root/reference) with its
own CPU OpenMP backend and gcc, in the build container.

  PYTHONPATH=oracle/refshim:/root/reference DEVITO_LANGUAGE=openmp DEVITO_ARCH=gcc \
  DEVITO_SAFE_MATH=1 DEVITO_LOGGING=ERROR python oracle/make_golden.py

`oracle/refshim` holds stand-ins for three pure-Python dependencies of the reference that are
not installable offline (cgen, codepy, anytree); the numerical path (SymPy lowering -> C ->
gcc) is the reference's own.  The fixtures land in tests/golden/*.npz; the reference is not
needed to run the tests.
"""
import os
import sys

import numpy as np

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')


def save(name, **arrays):
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB",
          {k: (v.shape if hasattr(v, 'shape') and v.ndim else v) for k, v in arrays.items() if np.size(v) < 8})


def kat2d():
    """tests/test_gpu_openacc.py:205-251 (norm(rec) = 490.56 +- 1e-2)."""
    from devito import Grid, TimeFunction, Function, Eq, Operator, solve, norm
    from examples.seismic import TimeAxis, RickerSource, Receiver
    shape, extent = (101, 101), (1000, 1000)
    v = np.empty(shape, dtype=np.float32)
    v[:, :51] = 1.5
    v[:, 51:] = 2.5
    grid = Grid(shape=shape, extent=extent, origin=(0., 0.))
    dt = 1.6
    time_range = TimeAxis(start=0., stop=1000., step=dt)
    src = RickerSource(name='src', grid=grid, f0=0.010, npoint=1, time_range=time_range)
    src.coordinates.data[0, :] = np.array(extent) * .5
    src.coordinates.data[0, -1] = 20.
    rec = Receiver(name='rec', grid=grid, npoint=101, time_range=time_range)
    rec.coordinates.data[:, 0] = np.linspace(0, extent[0], num=101)
    rec.coordinates.data[:, 1] = 20.
    u = TimeFunction(name="u", grid=grid, time_order=2, space_order=2)
    m = Function(name='m', grid=grid)
    m.data[:] = 1. / (v * v)
    stencil = Eq(u.forward, solve(m * u.dt2 - u.laplace, u.forward))
    src_term = src.inject(field=u.forward, expr=src * dt ** 2 / m)
    rec_term = rec.interpolate(expr=u.forward)
    op = Operator([stencil] + src_term + rec_term)
    op(time=time_range.num - 1, dt=dt)
    save('kat2d_so2', norm_rec=np.float32(norm(rec)), rec=np.array(rec.data[::5, ::4]),
         u_last=np.array(u.data[(time_range.num) % 3, ::2, ::2]), nt=time_range.num,
         src=np.array(src.data[:, 0]))


def acoustic(name, so, n, nbl, tn, preset='constant-isotropic', interpolation='linear', kernel='OT2', **kw):
    from devito import norm
    from examples.seismic import demo_model, setup_geometry
    from examples.seismic.acoustic import AcousticWaveSolver
    model = demo_model(preset, spacing=(10., 10., 10.), shape=(n, n, n), nbl=nbl, space_order=so,
                       dtype=np.float32, **kw)
    geometry = setup_geometry(model, tn, interpolation=interpolation)
    solver = AcousticWaveSolver(model, geometry, space_order=so, kernel=kernel)
    rec, u, _ = solver.forward()
    extra = {}
    if not model.vp.is_Constant:
        extra['vp'] = np.array(model.vp.data)
    save(name, so=so, n=n, nbl=nbl, tn=tn, dt=np.float32(model.critical_dt), nt=geometry.nt,
         damp=np.array(model.damp.data), src=np.array(geometry.src.data),
         src_coords=np.array(geometry.src.coordinates.data),
         rec_coords=np.array(geometry.rec.coordinates.data), rec=np.array(rec.data),
         u=np.array(u.data), norm_rec=np.float32(norm(rec)), norm_u=np.float32(norm(u)),
         dt_run=np.float32(solver.dt), **extra)


def kat3d_fs(interpolation, name):
    """examples/seismic/acoustic/acoustic_example.py:80-87: `run(fs=True, dtype=float32)` -> norm(rec)
    = 369.955 (linear) / 402.216 (sinc), rtol 1e-3."""
    from devito import norm
    from examples.seismic.acoustic.acoustic_example import acoustic_setup
    solver = acoustic_setup(shape=(50, 50, 50), spacing=(20., 20., 20.), nbl=40, tn=1000., space_order=4,
                            kernel='OT2', fs=True, preset='layers-isotropic', dtype=np.float32,
                            interpolation=interpolation)
    rec, u, _ = solver.forward()
    g = solver.geometry
    save(name, norm_rec=np.float32(norm(rec)), rec=np.array(rec.data[::4, ::7]), nt=g.nt,
         dt=np.float32(solver.model.critical_dt), grid_shape=np.array(solver.model.grid.shape),
         u_last=np.array(u.data[(g.nt - 1) % 3, ::3, ::3, ::3]))


def adjoint_variant(name, so, n, nbl, tn, kernel='OT2', **kw):
    """Forward then adjoint for the free-surface / OT4 variants of the acoustic operators."""
    from examples.seismic import demo_model, setup_geometry
    from examples.seismic.acoustic import AcousticWaveSolver
    model = demo_model('layers-isotropic', spacing=(10., 10., 10.), shape=(n, n, n), nbl=nbl, space_order=so,
                       dtype=np.float32, nlayers=2, **kw)
    geometry = setup_geometry(model, tn)
    solver = AcousticWaveSolver(model, geometry, space_order=so, kernel=kernel)
    rec, u, _ = solver.forward()
    srca, v, _ = solver.adjoint(rec)
    save(name, so=so, n=n, nbl=nbl, tn=tn, dt=np.float32(model.critical_dt), dt_run=np.float32(solver.dt),
         nt=geometry.nt, vp=np.array(model.vp.data), rec=np.array(rec.data), srca=np.array(srca.data),
         v=np.array(v.data))


def adjoint(name, so, n, nbl, tn):
    """Forward then adjoint (acoustic/wavesolver.py:118-156): receiver data back-propagated."""
    from devito import norm
    from examples.seismic import demo_model, setup_geometry
    from examples.seismic.acoustic import AcousticWaveSolver
    model = demo_model('constant-isotropic', spacing=(10., 10., 10.), shape=(n, n, n), nbl=nbl,
                       space_order=so, dtype=np.float32)
    geometry = setup_geometry(model, tn)
    solver = AcousticWaveSolver(model, geometry, space_order=so)
    rec, u, _ = solver.forward()
    srca, v, _ = solver.adjoint(rec)
    save(name, so=so, n=n, nbl=nbl, tn=tn, dt=np.float32(model.critical_dt), nt=geometry.nt,
         rec=np.array(rec.data), srca=np.array(srca.data), v=np.array(v.data),
         src=np.array(geometry.src.data), norm_v=np.float32(norm(v)))


def born(name, so, n, nbl, tn):
    """Linearised modelling `solver.jacobian(dm)` (acoustic/wavesolver.py:216-254) with a smooth blob
    as model perturbation."""
    from examples.seismic import demo_model, setup_geometry
    from examples.seismic.acoustic import AcousticWaveSolver
    model = demo_model('layers-isotropic', spacing=(10., 10., 10.), shape=(n, n, n), nbl=nbl, space_order=so,
                       dtype=np.float32, nlayers=2)
    geometry = setup_geometry(model, tn)
    solver = AcousticWaveSolver(model, geometry, space_order=so)
    N = n + 2 * nbl
    ax = np.arange(N, dtype=np.float64)
    X, Y, Z = np.meshgrid(ax, ax, ax, indexing='ij')
    c = (N - 1) / 2.0
    dm = (0.05 * np.exp(-((X - c) ** 2 + (Y - c) ** 2 + (Z - c - 2) ** 2) / (2 * 4.0 ** 2))).astype(np.float32)
    rec, u, U, _ = solver.jacobian(dm)
    save(name, so=so, n=n, nbl=nbl, tn=tn, dt=np.float32(model.critical_dt), nt=geometry.nt,
         vp=np.array(model.vp.data), dm=dm, rec=np.array(rec.data), u=np.array(u.data), U=np.array(U.data))


def snapshots(name, so, n, nbl, tn, factor):
    """Time-subsampled saving, examples/seismic/tutorials/08_snapshotting.ipynb:455-505: `Eq(usave, u)`
    with usave on ConditionalDimension(factor)."""
    from devito import ConditionalDimension, Eq, Operator, TimeFunction, solve
    from examples.seismic import demo_model, setup_geometry
    model = demo_model('constant-isotropic', spacing=(10., 10., 10.), shape=(n, n, n), nbl=nbl, space_order=so,
                       dtype=np.float32, bcs='damp')
    geometry = setup_geometry(model, tn)
    nt = geometry.nt
    nsnaps = (nt + factor - 1) // factor
    t_sub = ConditionalDimension('t_sub', parent=model.grid.time_dim, factor=factor)
    usave = TimeFunction(name='usave', grid=model.grid, time_order=2, space_order=2, save=nsnaps, time_dim=t_sub)
    u = TimeFunction(name='u', grid=model.grid, time_order=2, space_order=so)
    pde = model.m * u.dt2 - u.laplace + model.damp * u.dt
    stencil = Eq(u.forward, solve(pde, u.forward))
    src, rec = geometry.src, geometry.rec
    dt = model.critical_dt
    op = Operator([stencil] + src.inject(field=u.forward, expr=src * dt ** 2 / model.m) + [Eq(usave, u)] +
                  rec.interpolate(expr=u), subs=model.spacing_map)
    op(time=nt - 2, dt=dt)
    save(name, so=so, n=n, nbl=nbl, tn=tn, dt=np.float32(dt), nt=nt, factor=factor, nsnaps=nsnaps,
         usave=np.array(usave.data), rec=np.array(rec.data), u=np.array(u.data))


def gradient(name, so, n, nbl, tn):
    """Forward with the saved wavefield, then the Gradient operator (acoustic/operators.py:190-232,
    wavesolver.py:158-230): adjoint propagation of the data + imaging condition."""
    from devito import norm
    from examples.seismic import demo_model, setup_geometry
    from examples.seismic.acoustic import AcousticWaveSolver
    model = demo_model('constant-isotropic', spacing=(10., 10., 10.), shape=(n, n, n), nbl=nbl,
                       space_order=so, dtype=np.float32)
    geometry = setup_geometry(model, tn)
    solver = AcousticWaveSolver(model, geometry, space_order=so)
    rec, u, _ = solver.forward(save=True)
    grad, _ = solver.jacobian_adjoint(rec, u)
    save(name, so=so, n=n, nbl=nbl, tn=tn, dt=np.float32(model.critical_dt), nt=geometry.nt,
         rec=np.array(rec.data), grad=np.array(grad.data), norm_grad=np.float32(norm(grad)),
         u_last=np.array(u.data[geometry.nt - 1]))


def tti(name, so, n, nbl, tn, preset='constant-tti', **kw):
    from devito import norm
    from examples.seismic import demo_model, setup_geometry
    from examples.seismic.tti import AnisotropicWaveSolver
    model = demo_model(preset, spacing=(10., 10., 10.), shape=(n, n, n), nbl=nbl,
                       space_order=so, dtype=np.float32, **kw)
    geometry = setup_geometry(model, tn)
    solver = AnisotropicWaveSolver(model, geometry, space_order=so)
    rec, u, v, _ = solver.forward()
    save(name, so=so, n=n, nbl=nbl, tn=tn, dt=np.float32(model.critical_dt), nt=geometry.nt,
         damp=np.array(model.damp.data), src=np.array(geometry.src.data),
         src_coords=np.array(geometry.src.coordinates.data),
         rec_coords=np.array(geometry.rec.coordinates.data), rec=np.array(rec.data),
         u=np.array(u.data), v=np.array(v.data), norm_rec=np.float32(norm(rec)),
         norm_u=np.float32(norm(u)), norm_v=np.float32(norm(v)),
         epsilon=np.array(model.epsilon.data, dtype=np.float32), delta=np.array(model.delta.data, dtype=np.float32),
         theta=np.array(model.theta.data, dtype=np.float32), phi=np.array(model.phi.data, dtype=np.float32),
         vp=np.array(model.vp.data, dtype=np.float32))


def tti_varying(name, so, n, nbl, tn):
    """TTI with vp, epsilon, delta, theta, phi varying along EVERY axis (the `layers-tti` preset varies along z only):
    pins where the reference samples the rotation factors (at the Gz point inside Gz, at the shifted point in the
    outer derivative) in all three directions."""
    from devito import norm
    from examples.seismic import SeismicModel, setup_geometry
    from examples.seismic.tti import AnisotropicWaveSolver
    shape = (n, n, n)
    gx, gy, gz = np.meshgrid(*[np.linspace(0., 1., m, dtype=np.float32) for m in shape], indexing='ij')
    par = dict(vp=(1.5 + 0.6 * gx + 0.5 * gy + 0.9 * gz).astype(np.float32),
               epsilon=(0.25 * gx * gz + 0.05).astype(np.float32), delta=(0.12 * gy + 0.02).astype(np.float32),
               theta=(0.2 + 0.9 * gx * gy + 0.3 * gz).astype(np.float32),
               phi=(0.1 + 0.8 * gy * gz - 0.4 * gx).astype(np.float32))
    model = SeismicModel(space_order=so, origin=(0., 0., 0.), shape=shape, dtype=np.float32, spacing=(10., 10., 10.),
                         nbl=nbl, bcs="damp", **par)
    geometry = setup_geometry(model, tn)
    solver = AnisotropicWaveSolver(model, geometry, space_order=so)
    rec, u, v, _ = solver.forward()
    # the parameter arrays follow the formulas above (tests/helpers.py::varying_tti_parameters restates them); the model
    # pads them with edge values over the absorbing layers, which the fixture records for vp only as a cross-check
    save(name, so=so, n=n, nbl=nbl, tn=tn, dt=np.float32(model.critical_dt), nt=geometry.nt,
         src_coords=np.array(geometry.src.coordinates.data), rec=np.array(rec.data),
         slot=(geometry.nt - 1) % 3, u_last=np.array(u.data[(geometry.nt - 1) % 3]),
         v_last=np.array(v.data[(geometry.nt - 1) % 3]), norm_rec=np.float32(norm(rec)),
         norm_u=np.float32(norm(u)), norm_v=np.float32(norm(v)),
         vp_line=np.array(model.vp.data[:, n // 2 + nbl, n // 3 + nbl], dtype=np.float32),
         theta_line=np.array(model.theta.data[n // 3 + nbl, :, n // 2 + nbl], dtype=np.float32))


def coefficients():
    """FD weights and CFL numbers straight from the reference's machinery."""
    from devito import Grid, TimeFunction
    from examples.seismic import demo_model
    out = {}
    for so in (4, 8, 12, 16):
        model = demo_model('constant-isotropic', spacing=(10., 10., 10.), shape=(8, 8, 8), nbl=2,
                           space_order=so, dtype=np.float32)
        out[f'dt_iso_so{so}'] = np.float32(model.critical_dt)
        mt = demo_model('constant-tti', spacing=(10., 10., 10.), shape=(8, 8, 8), nbl=2,
                        space_order=so, dtype=np.float32)
        out[f'dt_tti_so{so}'] = np.float32(mt.critical_dt)
    save('coefficients', **out)


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ['kat2d', 'fs', 'ot4', 'born', 'snap', 'adjvar', 'iso8', 'iso12', 'iso4layers', 'iso8sinc', 'adj8', 'grad8', 'tti8', 'tti4', 'tti4layers', 'tti8varying', 'coef']
    if 'kat2d' in which:
        kat2d()
    if 'fs' in which:
        acoustic('iso3d_so4_fs', so=4, n=20, nbl=8, tn=150.0, preset='layers-isotropic', nlayers=3, fs=True)
        acoustic('iso3d_so8_fs_sinc', so=8, n=20, nbl=8, tn=120.0, preset='layers-isotropic', nlayers=2,
                 interpolation='sinc', fs=True)
        kat3d_fs('linear', 'kat3d_fs_linear')
        kat3d_fs('sinc', 'kat3d_fs_sinc')
    if 'ot4' in which:
        acoustic('iso3d_so8_ot4', so=8, n=20, nbl=8, tn=150.0, kernel='OT4')
        acoustic('iso3d_so4_ot4_layers', so=4, n=20, nbl=8, tn=150.0, kernel='OT4',
                 preset='layers-isotropic', nlayers=3)
    if 'adjvar' in which:
        adjoint_variant('adj3d_so4_fs', so=4, n=20, nbl=8, tn=120.0, fs=True)
        adjoint_variant('adj3d_so4_ot4', so=4, n=20, nbl=8, tn=120.0, kernel='OT4')
    if 'snap' in which:
        snapshots('snap3d_so4', so=4, n=20, nbl=8, tn=150.0, factor=4)
    if 'born' in which:
        born('born3d_so8', so=8, n=20, nbl=8, tn=150.0)
    if 'iso8' in which:
        acoustic('iso3d_so8', so=8, n=20, nbl=8, tn=150.0)
    if 'iso12' in which:
        acoustic('iso3d_so12', so=12, n=20, nbl=8, tn=150.0)
    if 'iso4layers' in which:
        acoustic('iso3d_so4_layers', so=4, n=20, nbl=8, tn=150.0, preset='layers-isotropic', nlayers=3)
    if 'iso8sinc' in which:
        acoustic('iso3d_so8_sinc', so=8, n=20, nbl=8, tn=100.0, interpolation='sinc')
    if 'adj8' in which:
        adjoint('adj3d_so8', so=8, n=20, nbl=8, tn=150.0)
    if 'grad8' in which:
        gradient('grad3d_so8', so=8, n=20, nbl=8, tn=150.0)
    if 'tti8' in which:
        tti('tti3d_so8', so=8, n=20, nbl=8, tn=150.0)
    if 'tti8varying' in which:
        tti_varying('tti3d_so8_varying', so=8, n=24, nbl=8, tn=90.0)
    if 'tti4layers' in which:
        tti('tti3d_so4_layers', so=4, n=20, nbl=8, tn=100.0, preset='layers-tti', nlayers=3)
    if 'tti4' in which:
        tti('tti3d_so4', so=4, n=20, nbl=8, tn=120.0)
    if 'coef' in which:
        coefficients()
