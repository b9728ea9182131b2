"""BASELINE config 1 on the GPU: the 2-D diffusion example (examples/cfd/example_diffusion.py) through
`b2_linear_forward`, against the reference's NumPy twin (:61-83) and its own assertions (:162-166)."""
import numpy as np
import pytest

from helpers import rel_linf

pytestmark = [pytest.mark.gpu, pytest.mark.pending]


def _twin(init, a, dt, hx, hy, nt):
    ref = init.astype(np.float64)
    for _ in range(nt):
        new = ref.copy()
        new[1:-1, 1:-1] = ref[1:-1, 1:-1] + a * dt * (
            (ref[2:, 1:-1] - 2 * ref[1:-1, 1:-1] + ref[:-2, 1:-1]) / hx ** 2 +
            (ref[1:-1, 2:] - 2 * ref[1:-1, 1:-1] + ref[1:-1, :-2]) / hy ** 2)
        ref = new
    return ref


def test_diffusion_matches_numpy_twin():
    from devito_b200 import Eq, Grid, Operator, TimeFunction, solve
    n, nt, a = 512, 50, 0.5                                # BASELINE config 1: 512^2, space_order 2
    g = Grid(shape=(n, n), extent=(2., 2.))
    u = TimeFunction(name='u', grid=g, time_order=1, space_order=2)
    init = np.zeros((n, n), dtype=np.float32)
    init[n // 4:n // 2, n // 4:n // 2] = 1.0
    u.data[0] = init
    u.data[1] = init
    hx, hy = g.spacing
    dt = 0.2 * hx * hy / a
    op = Operator([Eq(u.forward, solve(Eq(u.dt, a * u.laplace), u.forward), subdomain=g.interior)])
    assert op.backend == 'cuda-sm100a'
    summary = op(time_M=nt - 1, dt=dt)
    assert rel_linf(u.data[nt % 2], _twin(init, a, float(np.float32(dt)), hx, hy, nt)) < 1e-5
    assert summary['section0'].time > 0


def test_reference_example_assertions():
    """examples/cfd/example_diffusion.py:162-166 `test_diffusion2d`: ring initial condition, 1000 steps."""
    from devito_b200 import Eq, Grid, Operator, TimeFunction, solve
    spacing, timesteps, a = 0.01, 1000, 0.5
    nx = ny = int(2 / spacing)
    xx, yy = np.meshgrid(np.linspace(0., 2., nx, dtype=np.float32), np.linspace(0., 2., ny, dtype=np.float32))
    ui = np.zeros((nx, ny), dtype=np.float32)
    r = (xx - 1.) ** 2. + (yy - 1.) ** 2.
    ui[np.logical_and(.05 <= r, r <= .1)] = 1.
    dx2 = dy2 = spacing ** 2
    dt = dx2 * dy2 / (2 * a * (dx2 + dy2))
    grid = Grid(shape=(nx, ny))
    u = TimeFunction(name='u', grid=grid, time_order=1, space_order=2)
    u.data[0, :] = ui[:]
    op = Operator(Eq(u.forward, solve(Eq(u.dt, a * (u.dx2 + u.dy2)), u.forward)))
    op.apply(u=u, t=timesteps, dt=dt)
    out = np.asarray(u.data[1, :])
    assert out.max() < 2.4
    assert np.linalg.norm(out, ord=2) < 13
