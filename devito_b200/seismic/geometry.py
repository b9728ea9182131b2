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
    def __init__(self, model, rec_positions, src_positions, t0, tn, **kwargs):
        self.src_positions = np.reshape(src_positions, (-1, model.dim))
        self.rec_positions = np.reshape(rec_positions, (-1, model.dim))
        self._nrec = self.rec_positions.shape[0]
        self._nsrc = self.src_positions.shape[0]
        self._src_type = kwargs.get('src_type')
        assert self._src_type in sources or self._src_type is None
        self._f0 = kwargs.get('f0')
        self._a = kwargs.get('a')
        self._t0w = kwargs.get('t0w')
        self._grid = model.grid
        self._model = model
        self._dt = model.critical_dt
        self._t0, self._tn = t0, tn
        self._interpolation = kwargs.get('interpolation', 'linear')
        self._r = kwargs.get('r', _default_radius[self._interpolation])

    def resample(self, dt):
        self._dt = dt
        return self

    @property
    def time_axis(self): return TimeAxis(start=self.t0, stop=self.tn, step=self.dt)
    @property
    def src_type(self): return self._src_type
    @property
    def grid(self): return self._grid
    @property
    def f0(self): return self._f0
    @property
    def tn(self): return self._tn
    @property
    def t0(self): return self._t0
    @property
    def dt(self): return self._dt
    @property
    def nt(self): return self.time_axis.num
    @property
    def nrec(self): return self._nrec
    @property
    def nsrc(self): return self._nsrc
    @property
    def dtype(self): return self.grid.dtype
    @property
    def r(self): return self._r
    @property
    def interpolation(self): return self._interpolation

    @property
    def rec(self): return self.new_rec()

    def new_rec(self, name='rec', coordinates=None):
        coords = coordinates if coordinates is not None else self.rec_positions
        return Receiver(name=name, grid=self.grid, time_range=self.time_axis, npoint=self.nrec,
                        interpolation=self.interpolation, r=self._r, coordinates=coords)

    @property
    def src(self): return self.new_src()

    def new_src(self, name='src', src_type='self', coordinates=None):
        coords = coordinates if coordinates is not None else self.src_positions
        if self.src_type is None or src_type is None:
            return PointSource(name=name, grid=self.grid, time_range=self.time_axis, npoint=self.nsrc,
                               coordinates=coords, interpolation=self.interpolation, r=self._r)
        return sources[self.src_type](name=name, grid=self.grid, f0=self.f0, time_range=self.time_axis,
                                      npoint=self.nsrc, coordinates=coords, t0=self._t0w, a=self._a,
                                      interpolation=self.interpolation, r=self._r)
