"""Acquisition geometry (mirror of examples/seismic/utils.py:14-209)."""
import numpy as np

from ..sparse import _default_radius
from .source import TimeAxis, PointSource, Receiver, RickerSource, GaborSource, WaveletSource

__all__ = ['AcquisitionGeometry', 'setup_geometry', 'setup_rec_coords']

sources = {'Wavelet': WaveletSource, 'Ricker': RickerSource, 'Gabor': GaborSource}


def setup_rec_coords(model):
    """Receiver line (2-D) / shape[0] x shape[1] grid (3-D) at depth origin_z + 2 h_z
    (utils.py:33-53)."""
    nrecx = model.shape[0]
    recx = np.linspace(model.origin[0], model.domain_size[0], nrecx)
    if model.dim == 1:
        return recx.reshape((nrecx, 1))
    if model.dim == 2:
        rec = np.empty((nrecx, 2))
        rec[:, 0] = recx
        rec[:, -1] = model.origin[-1] + 2 * model.spacing[-1]
        return rec
    nrecy = model.shape[1]
    recy = np.linspace(model.origin[1], model.domain_size[1], nrecy)
    rec = np.empty((nrecx * nrecy, 3))
    rec[:, 0] = np.repeat(recx, nrecy)
    rec[:, 1] = np.tile(recy, nrecx)
    rec[:, -1] = model.origin[-1] + 2 * model.spacing[-1]
    return rec


def setup_geometry(model, tn, f0=0.010, interpolation='linear', **kwargs):
    """One Ricker source at the domain centre, depth origin_z + h_z (utils.py:14-30)."""
    src = np.empty((1, model.dim))
    if model.dim > 1:
        src[0, :] = np.array(model.domain_size) * .5
        src[0, -1] = model.origin[-1] + model.spacing[-1]
    else:
        src[0, 0] = 2 * model.spacing[0]
    rec = kwargs.pop('rec_coordinates', None)
    if rec is None:
        rec = setup_rec_coords(model)
    r = kwargs.get('r', _default_radius[interpolation])
    return AcquisitionGeometry(model, rec, src, t0=0.0, tn=tn, src_type='Ricker', f0=f0,
                               interpolation=interpolation, r=r)


class AcquisitionGeometry:
    """Source/receiver positions plus the time axis of one shot (examples/seismic/utils.py:56-209).
    `.src` / `.rec` build fresh sparse functions on every access, like the reference."""

    def __init__(self, model, rec_positions, src_positions, t0, tn, **kwargs):
        ndim = model.dim
        self.rec_positions = np.asarray(rec_positions, dtype=np.float64).reshape(-1, ndim)
        self.src_positions = np.asarray(src_positions, dtype=np.float64).reshape(-1, ndim)
        self._src_type = kwargs.get('src_type')
        if self._src_type is not None and self._src_type not in sources:
            raise ValueError(f"unknown source type {self._src_type!r}")
        self._f0, self._a, self._t0w = kwargs.get('f0'), kwargs.get('a'), kwargs.get('t0w')
        if self._src_type is not None and self._f0 is None:
            raise ValueError(f"Peak frequency must be provided in KHz for source of type {self._src_type}")
        self._model, self._grid = model, model.grid
        self._dt = model.critical_dt
        self._t0, self._tn = t0, tn
        self._interpolation = kwargs.get('interpolation', 'linear')
        self._r = kwargs.get('r', _default_radius[self._interpolation])

    def resample(self, dt):
        self._dt = dt
        return self

    # read-only views -------------------------------------------------------------------------------
    grid = property(lambda self: self._grid)
    src_type = property(lambda self: self._src_type)
    f0 = property(lambda self: self._f0)
    t0 = property(lambda self: self._t0)
    tn = property(lambda self: self._tn)
    dt = property(lambda self: self._dt)
    r = property(lambda self: self._r)
    interpolation = property(lambda self: self._interpolation)
    dtype = property(lambda self: self._grid.dtype)
    nrec = property(lambda self: self.rec_positions.shape[0])
    nsrc = property(lambda self: self.src_positions.shape[0])

    @property
    def time_axis(self):
        return TimeAxis(start=self._t0, stop=self._tn, step=self._dt)

    @property
    def nt(self):
        return self.time_axis.num

    # sparse functions ------------------------------------------------------------------------------
    def _common(self):
        return dict(grid=self._grid, time_range=self.time_axis, interpolation=self._interpolation, r=self._r)

    def new_rec(self, name='rec', coordinates=None):
        pos = self.rec_positions if coordinates is None else coordinates
        return Receiver(name=name, npoint=self.nrec, coordinates=pos, **self._common())

    def new_src(self, name='src', src_type='self', coordinates=None):
        pos = self.src_positions if coordinates is None else coordinates
        if self._src_type is None or src_type is None:
            return PointSource(name=name, npoint=self.nsrc, coordinates=pos, **self._common())
        return sources[self._src_type](name=name, npoint=self.nsrc, coordinates=pos, f0=self._f0,
                                       t0=self._t0w, a=self._a, **self._common())

    rec = property(lambda self: self.new_rec())
    src = property(lambda self: self.new_src())
