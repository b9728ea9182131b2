"""Time-subsampled snapshots: `Eq(usave, u)` with `usave` saved on a
ConditionalDimension — examples/seismic/tutorials/08_snapshotting.ipynb:455-505. CPU: the oracle stepped
from Python reproduces the reference golden, the operator is recognised; GPU:
`b2_iso_args.snap/snap_factor` against the golden."""
import numpy as np
import pytest

from oracle import oracle as O
from helpers import load_golden, rel_linf, domain, iso_problem


def test_snapshot_semantics_oracle_vs_reference():
    """`if (time % factor == 0) usave[time / factor] = u[time % 3]` — the pre-update time level."""
    g = load_golden('snap3d_so4')
    so, factor, nsnaps = 4, int(g['factor']), int(g['nsnaps'])
    p = iso_problem(int(g['n']), int(g['nbl']), so, float(g['tn']))
    assert p['nt'] == int(g['nt']) and np.float32(p['dt']) == g['dt']
    usave = np.zeros((nsnaps,) + g['usave'].shape[1:], dtype=np.float32)
    for time in range(1, p['nt'] - 1):                    # time = 1 .. nt-2, one oracle call per step
        O.iso_forward(p['u'], so, p['w'], p['dt'], time, time, damp=p['damp'], vp=1.5, src=p['src'], rec=p['rec'])
        if time % factor == 0:
            usave[time // factor] = domain(p['u'], so)[time % 3]
    assert rel_linf(usave, g['usave']) < 1e-5
    assert not g['usave'][0].any()                         # time = 0 is never visited (time_m = 1)
    assert rel_linf(domain(p['u'], so), g['u']) < 1e-5
    assert rel_linf(p['rec']['data'], g['rec']) < 1e-5


def _operator(g):
    from devito_b200 import ConditionalDimension, Eq, Operator, TimeFunction, solve
    from devito_b200.seismic import demo_model, setup_geometry
    n, nbl, so, factor = int(g['n']), int(g['nbl']), 4, int(g['factor'])
    model = demo_model('constant-isotropic', shape=(n,) * 3, spacing=(10.,) * 3, nbl=nbl, space_order=so, bcs='damp')
    geometry = setup_geometry(model, float(g['tn']))
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
    return op, model, geometry, u, usave, rec, dt


def test_snapshot_operator_is_recognised():
    from devito_b200.exceptions import InvalidArgument
    g = load_golden('snap3d_so4')
    op, model, geometry, u, usave, rec, dt = _operator(g)
    assert op.backend == 'cuda-sm100a', op._why_not
    p = op._plan
    assert p['snap'] is usave and p['snap_factor'] == int(g['factor']) and p['snap_toff'] == 0
    assert usave.shape[0] == int(g['nsnaps']) and geometry.nt == int(g['nt'])
    args = op.arguments(time=geometry.nt - 2, dt=dt)
    assert (args['time_m'], args['time_M']) == (1, geometry.nt - 2)
    with pytest.raises(InvalidArgument):                   # one snapshot too many
        op.arguments(time=int(g['nsnaps']) * int(g['factor']), dt=dt)
    assert 'a.snap_factor = 4;' in str(op)


@pytest.mark.gpu
def test_snapshots_vs_reference_golden():
    g = load_golden('snap3d_so4')
    op, model, geometry, u, usave, rec, dt = _operator(g)
    op(time=geometry.nt - 2, dt=dt)
    assert rel_linf(usave.data, g['usave']) < 1e-5
    assert rel_linf(u.data, g['u']) < 1e-5
    assert rel_linf(rec.data, g['rec']) < 1e-5


@pytest.mark.gpu
def test_snapshots_streamed_to_host_match_the_resident_run():
    """Host-resident snapshots are drained box by box on a copy stream while the stencil runs
    (SnapStreamer in b2_api_iso.cu); same values as the device-resident run, halo untouched."""
    g = load_golden('snap3d_so4')
    op, model, geometry, u, usave, rec, dt = _operator(g)
    usave.data_with_halo[:] = -3.0                         # marker: the halo and unvisited slots must survive
    op(time=geometry.nt - 2, dt=dt, resident=False)
    got = np.array(usave.data_with_halo)
    assert rel_linf(np.asarray(usave.data)[1:], g['usave'][1:]) < 1e-5
    h = usave.space_order
    assert np.all(got[:, :h] == -3.0) and np.all(got[:, :, :, -h:] == -3.0)
    assert np.all(got[0] == -3.0)                          # time = 0 is never visited
