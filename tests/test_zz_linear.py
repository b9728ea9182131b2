"""BASELINE config 1 on the GPU: the 2-D diffusion example (examples/cfd/example_diffusion.py) through
`b2_linear_forward`, against the reference's NumPy twin (:61-83) and its own assertions (:162-166)."""
import numpy as np
import pytest

from helpers import rel_linf

pytestmark = [pytest.mark.gpu]


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
    nx = ny = int(1 / spacing)                          # ring_initial (:24-32): 100 x 100 on the unit square
    xx, yy = np.meshgrid(np.linspace(0., 1., nx, dtype=np.float32), np.linspace(0., 1., ny, dtype=np.float32))
    ui = np.zeros((nx, ny), dtype=np.float32)
    r = (xx - .5) ** 2. + (yy - .5) ** 2.
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
    # the reference itself (gcc/OpenMP, run in the build container) gives max 0.2370588, norm 12.809233
    assert abs(float(out.max()) - 0.2370588) < 1e-4 * 0.2370588
    assert abs(float(np.linalg.norm(out, ord=2)) - 12.809233) < 1e-4 * 12.809233


def test_backward_update_runs_backward_in_time():
    """An explicit update of `u.backward` is stepped from time_M down to time_m."""
    from devito_b200 import Eq, Grid, Operator, TimeFunction
    n, nt = 48, 8
    g = Grid(shape=(n, n), extent=(1., 1.))
    u = TimeFunction(name='u', grid=g, time_order=1, space_order=2)
    rng = np.random.default_rng(5)
    last = rng.standard_normal((n, n)).astype(np.float32)
    u.data[(nt - 1) % 2] = last
    u.data[nt % 2] = 7.0                                    # garbage a forward-running loop would pick up
    x, y = g.dimensions
    eq = Eq(u.backward, 0.5 * u + 0.125 * (u.subs(x, x + 1) + u.subs(x, x - 1) + u.subs(y, y + 1) + u.subs(y, y - 1)),
            subdomain=g.interior)
    op = Operator([eq])
    assert op.backend == 'cuda-sm100a' and op._plan['wshift'] == -1
    op(time_m=1, time_M=nt - 1)
    buf = np.zeros((2, n, n))                               # NumPy twin on the same two rotating slots
    buf[(nt - 1) % 2], buf[nt % 2] = last, 7.0
    for t in range(nt - 1, 0, -1):
        a, b = buf[t % 2], buf[(t - 1) % 2]
        b[1:-1, 1:-1] = 0.5 * a[1:-1, 1:-1] + 0.125 * (a[2:, 1:-1] + a[:-2, 1:-1] + a[1:-1, 2:] + a[1:-1, :-2])
    assert rel_linf(u.data[0], buf[0]) < 1e-5
