"""Free surface: reference `freesurface`, examples/seismic/acoustic/operators.py:5-47,
models built with `fs=True`. CPU tests pin the host side (model geometry, damping without a top
layer, operator recognition) against the reference goldens; the GPU tests compare the CUDA path
(`b2_iso_args.free_surface`, kernel `k_iso_fs_fix`) with the goldens and with the reference's
known-answer norms of acoustic_example.py:80-87 (369.955 linear / 402.216 sinc, rtol 1e-3).
The oracle side of the same goldens is tests/test_oracle_golden.py."""
import numpy as np
import pytest

from helpers import load_golden, rel_linf


def _solver(so=4, n=20, nbl=8, tn=150.0, h=10.0, nlayers=3, interpolation='linear'):
    from devito_b200.seismic import AcousticWaveSolver, demo_model, setup_geometry
    model = demo_model('layers-isotropic', shape=(n,) * 3, spacing=(h,) * 3, nbl=nbl, space_order=so,
                       nlayers=nlayers, fs=True)
    geometry = setup_geometry(model, tn, interpolation=interpolation)
    return model, geometry, AcousticWaveSolver(model, geometry, space_order=so)


@pytest.mark.parametrize('name,so,nlayers,interp', [('iso3d_so4_fs', 4, 3, 'linear'),
                                                    ('iso3d_so8_fs_sinc', 8, 2, 'sinc')])
def test_free_surface_model_matches_reference(name, so, nlayers, interp):
    g = load_golden(name)
    model, geometry, solver = _solver(so=so, n=int(g['n']), nbl=int(g['nbl']), tn=float(g['tn']),
                                      nlayers=nlayers, interpolation=interp)
    n, nbl = int(g['n']), int(g['nbl'])
    assert model.grid.shape == (n + 2 * nbl, n + 2 * nbl, n + nbl)          # no layer above the surface
    assert tuple(float(o) for o in model.grid.origin) == (-10.0 * nbl, -10.0 * nbl, 0.0)
    assert model.padsizes == [(nbl, nbl), (nbl, nbl), (0, nbl)]
    assert np.float32(model.critical_dt) == g['dt'] and geometry.nt == int(g['nt'])
    assert rel_linf(model.damp.data, g['damp']) < 1e-6
    assert np.array_equal(np.asarray(model.vp.data), g['vp'])
    np.testing.assert_allclose(geometry.src.coordinates.data, g['src_coords'], rtol=1e-6)
    np.testing.assert_allclose(geometry.rec.coordinates.data, g['rec_coords'], rtol=1e-6)
    for op in (solver.op_fwd(), solver.op_adj()):
        assert op.backend == 'cuda-sm100a' and op._plan['free_surface']
        assert 'free_surface' in str(op)


def test_free_surface_needs_the_matching_update(tmp_path, monkeypatch):
    from devito_b200 import Eq, FreeSurface, Operator, TimeFunction, solve
    from devito_b200.exceptions import InvalidOperator, InvalidArgument
    monkeypatch.setenv('DEVITO_B200_JITDIR', str(tmp_path))
    model, geometry, solver = _solver()
    u = TimeFunction(name='u', grid=model.grid, time_order=2, space_order=4)
    pde = model.m * u.dt2 - u.laplace + model.damp * u.dt
    upd = Eq(u.forward, solve(pde, u.forward), subdomain=model.grid.subdomains['physdomain'])
    fsd = model.grid.subdomains['fsdomain']
    op = Operator([upd, FreeSurface(upd, fsd)], subs=model.spacing_map)
    assert op._plan['free_surface']
    assert 'a.free_surface = 1;' in op.cinterface(force=True)[0]
    # the update restricted to physdomain WITHOUT its free-surface rows does not tile the grid
    assert Operator([upd], subs=model.spacing_map).backend == 'numpy-interpreter'
    # a free surface next to an update on the whole grid is not the reference's scheme: refused
    whole = Eq(u.forward, solve(pde, u.forward))
    with pytest.raises(InvalidOperator):
        Operator([whole, FreeSurface(whole, fsd)], subs=model.spacing_map)
    # the surface row must be part of the iteration
    with pytest.raises(InvalidArgument):
        op.arguments(dt=1.0, z_m=1)


# ---------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize('kernel', [1, 0])
@pytest.mark.parametrize('name,so,nlayers,interp', [('iso3d_so4_fs', 4, 3, 'linear'),
                                                    ('iso3d_so8_fs_sinc', 8, 2, 'sinc')])
def test_free_surface_vs_reference_golden(name, so, nlayers, interp, kernel):
    g = load_golden(name)
    model, geometry, solver = _solver(so=so, n=int(g['n']), nbl=int(g['nbl']), tn=float(g['tn']),
                                      nlayers=nlayers, interpolation=interp)
    rec, u, _ = solver.forward(kernel=kernel)
    assert rel_linf(u.data, g['u']) < 1e-5
    assert rel_linf(rec.data, g['rec']) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize('name,interp,kat', [('kat3d_fs_linear', 'linear', 369.955),
                                             ('kat3d_fs_sinc', 'sinc', 402.216)])
def test_free_surface_known_answer_norms(name, interp, kat):
    """examples/seismic/acoustic/acoustic_example.py:80-87 `run(fs=True, dtype=float32)`."""
    from devito_b200 import norm
    g = load_golden(name)
    model, geometry, solver = _solver(so=4, n=50, nbl=40, tn=1000.0, h=20.0, nlayers=3, interpolation=interp)
    assert model.grid.shape == tuple(int(v) for v in g['grid_shape']) and geometry.nt == int(g['nt'])
    rec, u, _ = solver.forward()
    assert np.isclose(norm(rec), kat, rtol=1e-3, atol=0)
    assert rel_linf(np.asarray(rec.data)[::4, ::7], g['rec']) < 1e-4
    assert rel_linf(np.asarray(u.data)[(geometry.nt - 1) % 3, ::3, ::3, ::3], g['u_last']) < 1e-4


@pytest.mark.gpu
def test_free_surface_adjoint_vs_reference_golden():
    g = load_golden('adj3d_so4_fs')
    model, geometry, solver = _solver(so=4, n=int(g['n']), nbl=int(g['nbl']), tn=float(g['tn']), nlayers=2)
    rec, u, _ = solver.forward()
    assert rel_linf(rec.data, g['rec']) < 1e-5
    srca, v, _ = solver.adjoint(rec)
    assert rel_linf(v.data, g['v']) < 1e-4                 # non-deterministic atomics in the reference
    assert rel_linf(srca.data, g['srca']) < 1e-4
