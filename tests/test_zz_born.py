"""Linearised (Born) modelling — the reference's `Born` operator / `solver.jacobian(dm)`
(examples/seismic/acoustic/operators.py:235-277, wavesolver.py:216-254). CPU: recognition from the
reference's formulation; GPU: `b2_iso_args.born_U/born_dm` against the reference golden. The oracle
side of the golden is tests/test_oracle_golden.py; the kernel's point code runs on the CPU in
tests/test_zz_emulation.py."""
import numpy as np
import pytest

from helpers import load_golden, rel_linf


def _solver(g, so=8):
    from devito_b200.seismic import AcousticWaveSolver, demo_model, setup_geometry
    n, nbl = int(g['n']), int(g['nbl'])
    model = demo_model('layers-isotropic', shape=(n,) * 3, spacing=(10.,) * 3, nbl=nbl, space_order=so, nlayers=2)
    geometry = setup_geometry(model, float(g['tn']))
    return model, geometry, AcousticWaveSolver(model, geometry, space_order=so)


def test_born_is_recognised():
    g = load_golden('born3d_so8')
    model, geometry, solver = _solver(g)
    assert np.array_equal(np.asarray(model.vp.data), g['vp']) and geometry.nt == int(g['nt'])
    op = solver.op_born()
    assert op.backend == 'cuda-sm100a', op._why_not
    p = op._plan
    assert (p['u'].name, p['born_U'].name, p['born_dm'].name) == ('u', 'U', 'dm')
    assert p['src'].name == 'src' and p['rec'].name == 'rec' and p['rec_toff'] == 0
    # `dm` may be passed as a bare array of the grid's shape, like the reference does
    args = op.arguments(dm=np.ascontiguousarray(g['dm']), dt=float(model.critical_dt))
    assert args['born_dm'].storage.host.shape == model.grid.shape
    assert 'a.born_dm = (struct b2_dataobj *)dm_vec;' in str(op)


def test_born_lookalikes_are_refused():
    from devito_b200 import Eq, Function, Operator, TimeFunction, solve
    g = load_golden('born3d_so8')
    model, geometry, solver = _solver(g)
    grid = model.grid
    u = TimeFunction(name='u', grid=grid, time_order=2, space_order=8)
    U = TimeFunction(name='U', grid=grid, time_order=2, space_order=8)
    dm = Function(name='dm', grid=grid, space_order=0)

    def update(f, q=0):
        pde = model.m * f.dt2 - f.laplace - q + model.damp * f.dt
        return Eq(f.forward, solve(pde, f.forward))
    good = Operator([update(u), update(U, q=-dm * u.dt2)], subs=model.spacing_map)
    assert good.backend == 'cuda-sm100a' and good._plan['born_U'] is U
    # a first time derivative, or the wrong sign, is not the Born source
    for q in (-dm * u.dt, dm * u.dt2, -dm * u.dt2 * 2):
        op = Operator([update(u), update(U, q=q)], subs=model.spacing_map)
        assert op.backend == 'numpy-interpreter', q


@pytest.mark.gpu
def test_born_vs_reference_golden():
    g = load_golden('born3d_so8')
    model, geometry, solver = _solver(g)
    rec, u, U, _ = solver.jacobian(np.ascontiguousarray(g['dm']))
    assert rel_linf(u.data, g['u']) < 1e-5
    # U is driven by a second time difference (cancellation): 1e-4, like the oracle-vs-reference test
    assert rel_linf(U.data, g['U']) < 1e-4
    assert rel_linf(rec.data, g['rec']) < 1e-4
    # Function input gives the same result as the bare array
    from devito_b200 import Function
    dmf = Function(name='dm', grid=model.grid, space_order=0)
    dmf.data[:] = g['dm']
    rec2, _, U2, _ = solver.jacobian(dmf)
    assert np.array_equal(np.asarray(U2.data), np.asarray(U.data))
