"""Parity at BASELINE.json's full size (iso so=8, 1024^3) through a size-independent property:
LOCALITY. After K explicit steps a point source has influenced only the cells within K*radius of
its support, so the wavefield in a window around each source must equal a run of the CPU oracle on
a small grid that holds just that window (same local damping values, same source signature). Five
sources are placed where the full-size launch configuration is most exposed — the grid centre, an
x-chunk boundary of the sweep kernel, a corner of the absorbing layer, and the far y / x-z edges —
and everything outside their windows must still be exactly zero.

The CPU twin runs the same check with the oracle on both sides (a 200^3 "big" grid), which
validates the window arithmetic of this test itself without a GPU.
"""
import numpy as np
import pytest

from oracle import oracle as O
from helpers import rel_linf

SO, R, H, VP = 8, 4, 10.0, 1.5
K = 6                      # time steps: support radius K*R - R + 2 = 22 cells  <  WIN/2 - R
WIN = 64                   # window edge; the source's base cell sits at window index 31
BASE = 31


def _signature(nt, npoint):
    """A different, O(1) time series per source (first K+2 samples matter)."""
    t = np.arange(nt, dtype=np.float64)[:, None]
    p = np.arange(npoint, dtype=np.float64)[None, :]
    return (np.cos(0.9 * t + 0.7 * p) * (1.0 + 0.25 * p)).astype(np.float32)


def _coords(bases, origin):
    """Physical coordinates putting source `i` half-way between cells bases[i] and bases[i]+1."""
    return (np.asarray(origin, dtype=np.float64)[None, :] + (np.asarray(bases) + 0.5) * H).astype(np.float32)


def _small_run(damp_window, sig, dt):
    """Oracle on the WIN^3 window: zero initial state, no absorbing layer of its own."""
    n = WIN
    u = np.zeros((3, n + 2 * SO, n + 2 * SO, n + 2 * SO), dtype=np.float32)
    damp = np.zeros(u.shape[1:], dtype=np.float32)
    damp[SO:-SO, SO:-SO, SO:-SO] = damp_window
    c = np.full((1, 3), (BASE + 0.5) * H, dtype=np.float32)
    gp, ws = O.tabulate(c, (0., 0., 0.), (np.float32(H),) * 3, 1, 'linear')
    src = dict(data=np.ascontiguousarray(sig.reshape(-1, 1)), gp=gp, w=ws, r=1)
    w = [O.fd2_weights(SO, H)] * 3
    O.iso_forward(u, SO, w, dt, 1, K, damp=damp, vp=VP, src=src)
    return u


def _check_windows(window_of, damp_window_of, count_nonzero_total, bases, sig, dt):
    """window_of(lo) -> (3, WIN, WIN, WIN) block of the big run's DOMAIN starting at cell `lo`."""
    nz = 0
    for i, b in enumerate(bases):
        lo = tuple(int(x) - BASE for x in b)
        small = _small_run(damp_window_of(lo), sig[:, i], dt)[:, SO:-SO, SO:-SO, SO:-SO]
        big = window_of(lo)
        assert np.abs(small).max() > 1e-3
        err = rel_linf(big, small)
        assert err < 1e-5, f"source {i} at {b}: window differs from the oracle, rel L-inf {err:.3e}"
        # the wave has not reached the rim of the window
        rim = np.ones(big.shape[1:], dtype=bool)
        rim[R:-R, R:-R, R:-R] = False
        assert not big[:, rim].any()
        nz += int(np.count_nonzero(big))
    assert count_nonzero_total() == nz, "the wavefield is non-zero outside the source windows"


def test_window_locality_oracle_twin():
    """CPU: oracle on a 200^3 grid vs oracle on the windows — validates the test's own arithmetic."""
    n, nbl = 160, 20
    N = n + 2 * nbl
    spacing = (np.float32(H),) * 3
    origin = tuple(np.float32(-nbl * H) for _ in range(3))
    dt = float(O.critical_dt(SO, 3, H, VP))
    bases = [(36, 36, 36), (150, 120, 100), (100, N - 1 - 36, 40)]
    nt = K + 3
    sig = _signature(nt, len(bases))
    gp, ws = O.tabulate(_coords(bases, origin), origin, spacing, 1, 'linear')
    assert [tuple(g) for g in gp] == bases
    damp = O.damp_field((N, N, N), nbl, spacing, SO)
    u = np.zeros((3, N + 2 * SO, N + 2 * SO, N + 2 * SO), dtype=np.float32)
    O.iso_forward(u, SO, [O.fd2_weights(SO, H)] * 3, dt, 1, K, damp=damp, vp=VP,
                  src=dict(data=sig, gp=gp, w=ws, r=1))

    def window_of(lo):
        return u[:, SO + lo[0]:SO + lo[0] + WIN, SO + lo[1]:SO + lo[1] + WIN, SO + lo[2]:SO + lo[2] + WIN]

    def damp_window_of(lo):
        return damp[SO + lo[0]:SO + lo[0] + WIN, SO + lo[1]:SO + lo[1] + WIN, SO + lo[2]:SO + lo[2] + WIN]

    _check_windows(window_of, damp_window_of, lambda: int(np.count_nonzero(u)), bases, sig, dt)


@pytest.mark.gpu
def test_window_locality_full_size_1024():
    """B200: the headline configuration (iso so=8, 1024^3 incl. the 40-cell absorbing layers)."""
    import torch
    from devito_b200 import TimeFunction
    from devito_b200.seismic import AcousticWaveSolver, PointSource, demo_model, setup_geometry
    n, nbl = 944, 40
    N = n + 2 * nbl
    model = demo_model('constant-isotropic', shape=(n,) * 3, spacing=(H,) * 3, nbl=nbl, space_order=SO)
    geometry = setup_geometry(model, tn=30.)
    solver = AcousticWaveSolver(model, geometry, space_order=SO)
    dt = float(model.critical_dt)
    assert dt == float(O.critical_dt(SO, 3, H, VP))
    origin = tuple(float(o) for o in model.grid.origin)
    assert origin == (-nbl * H,) * 3
    bases = [(511, 511, 511),            # grid centre
             (255, 300, 700),            # support straddles the x-chunk boundary at plane 256
             (36, 36, 36),               # inside the absorbing corner (damp != 0)
             (767, N - 1 - 36, 200),     # far y edge, x-chunk boundary at 768
             (N - 1 - 36, 500, N - 1 - 36)]
    coords = _coords(bases, origin)
    src = PointSource(name='src', grid=model.grid, time_range=geometry.time_axis, coordinates=coords)
    assert src.nt >= K + 2
    sig = _signature(src.nt, len(bases))
    src.data[:] = sig
    gp, _ = src.tabulate()
    assert [tuple(int(v) for v in g) for g in gp] == bases
    u = TimeFunction(name='u', grid=model.grid, time_order=2, space_order=SO)
    solver.forward(src=src, u=u, time_m=1, time_M=K)
    dev = u.storage.dev
    assert isinstance(dev, torch.Tensor) and dev.is_cuda and tuple(dev.shape) == (3,) + (N + 2 * SO,) * 3
    damp = model.damp.data_ro_domain

    def window_of(lo):
        blk = dev[:, SO + lo[0]:SO + lo[0] + WIN, SO + lo[1]:SO + lo[1] + WIN, SO + lo[2]:SO + lo[2] + WIN]
        return blk.cpu().numpy()

    def damp_window_of(lo):
        return np.asarray(damp[lo[0]:lo[0] + WIN, lo[1]:lo[1] + WIN, lo[2]:lo[2] + WIN])

    _check_windows(window_of, damp_window_of, lambda: int(torch.count_nonzero(dev).item()), bases, sig, dt)
