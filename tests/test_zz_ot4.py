"""kernel='OT4' (SURVEY §8f rank 2): the reference's 4th-order-in-time acoustic scheme,
examples/seismic/acoustic/operators.py:50-68 — H = lap(u) + dt^2/12 * lap(lap(u)/m), stepped with
dt = 1.73 * critical_dt (acoustic/wavesolver.py:39-44). CPU: recognition from the reference's own
formulation; GPU: `b2_iso_args.ot4` (two-pass generic kernels) against the reference goldens. The
oracle side of the goldens is tests/test_oracle_golden.py, the kernels' point code runs on the CPU in
tests/test_zz_emulation.py."""
import numpy as np
import pytest

from helpers import load_golden, rel_linf

CASES = [('iso3d_so8_ot4', 'constant-isotropic', 8), ('iso3d_so4_ot4_layers', 'layers-isotropic', 4)]


def _solver(g, preset, so):
    from devito_b200.seismic import AcousticWaveSolver, demo_model, setup_geometry
    n, nbl = int(g['n']), int(g['nbl'])
    model = demo_model(preset, shape=(n,) * 3, spacing=(10.,) * 3, nbl=nbl, space_order=so, nlayers=3)
    geometry = setup_geometry(model, float(g['tn']))
    return model, geometry, AcousticWaveSolver(model, geometry, space_order=so, kernel='OT4')


@pytest.mark.parametrize('name,preset,so', CASES)
def test_ot4_is_recognised(name, preset, so):
    g = load_golden(name)
    model, geometry, solver = _solver(g, preset, so)
    assert np.float32(solver.dt) == g['dt_run'] and geometry.nt == int(g['nt'])
    for op in (solver.op_fwd(), solver.op_adj()):
        assert op.backend == 'cuda-sm100a', op._why_not
        assert op._plan['ot4'] and op._plan['R'] == so // 2
        assert op._plan['m_role'][0] == ('vp_c' if preset.startswith('constant') else 'vp_f')
        assert 'a.ot4 = 1;' in str(op)


def test_ot4_lookalikes_are_refused():
    """A biharmonic term with the wrong factor, or with 1/m sampled at the output point instead of the
    shifted point, is not the reference's scheme."""
    from devito_b200 import Eq, Operator, TimeFunction, solve
    from devito_b200.seismic import demo_model
    model = demo_model('layers-isotropic', shape=(12,) * 3, spacing=(10.,) * 3, nbl=4, space_order=4)
    model._initialize_bcs(bcs="damp")
    u = TimeFunction(name='u', grid=model.grid, time_order=2, space_order=4)
    s = model.grid.time_dim.spacing
    m = model.m

    def build(H):
        pde = m * u.dt2 - H + model.damp * u.dt
        return Operator([Eq(u.forward, solve(pde, u.forward))], subs=model.spacing_map)
    good = build(u.laplace + s ** 2 / 12 * u.biharmonic(1 / m))
    assert good.backend == 'cuda-sm100a' and good._plan['ot4']
    assert build(u.laplace + s ** 2 / 10 * u.biharmonic(1 / m)).backend == 'numpy-interpreter'
    assert build(u.laplace + s ** 2 / 12 * (1 / m) * u.biharmonic(1)).backend == 'numpy-interpreter'


@pytest.mark.gpu
@pytest.mark.parametrize('name,preset,so', CASES)
def test_ot4_vs_reference_golden(name, preset, so):
    g = load_golden(name)
    model, geometry, solver = _solver(g, preset, so)
    rec, u, _ = solver.forward()
    assert rel_linf(u.data, g['u']) < 1e-5
    assert rel_linf(rec.data, g['rec']) < 1e-5


@pytest.mark.gpu
def test_ot4_adjoint_vs_reference_golden():
    from devito_b200.seismic import AcousticWaveSolver, demo_model, setup_geometry
    g = load_golden('adj3d_so4_ot4')
    n, nbl = int(g['n']), int(g['nbl'])
    model = demo_model('layers-isotropic', shape=(n,) * 3, spacing=(10.,) * 3, nbl=nbl, space_order=4, nlayers=2)
    geometry = setup_geometry(model, float(g['tn']))
    solver = AcousticWaveSolver(model, geometry, space_order=4, kernel='OT4')
    assert np.float32(solver.dt) == g['dt_run']
    rec, u, _ = solver.forward()
    assert rel_linf(rec.data, g['rec']) < 1e-5
    srca, v, _ = solver.adjoint(rec)
    assert rel_linf(v.data, g['v']) < 1e-4
    assert rel_linf(srca.data, g['srca']) < 1e-4
