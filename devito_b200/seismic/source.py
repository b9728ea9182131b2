"""Time axis, point sources and receivers.

Host-side mirror of the reference's seismic sources (examples/seismic/source.py:24-289): same
constructor keywords (`name, grid, time_range, npoint, coordinates, data, f0, a, t0`), same
attributes (`.time_values`, `.time_range`, `.wavelet`), same numerical definitions — written from
scratch on this package's `SparseTimeFunction`.
"""
import math

import numpy as np

from ..sparse import SparseTimeFunction

__all__ = ['TimeAxis', 'PointSource', 'Receiver', 'Shot', 'WaveletSource', 'RickerSource',
           'GaborSource']


class TimeAxis:
    """Uniform time axis defined by any three of (start, step, num, stop).

    Reference rule (source.py:24-63): when `num` is the missing one it is
    `ceil((stop - start + step) / step)` and `stop` is then moved to `start + step*(num-1)`;
    the sample values are `linspace(start, stop, num)`."""

    __slots__ = ('start', 'stop', 'step', 'num', '_values')

    def __init__(self, start=None, step=None, num=None, stop=None):
        given = {'start': start, 'step': step, 'num': num, 'stop': stop}
        missing = [k for k, v in given.items() if v is None]
        if len(missing) != 1:
            raise ValueError("exactly three of start, step, num and stop must be given")
        what = missing[0]
        if what == 'start':
            start = stop - step * (num - 1)
        elif what == 'step':
            step = (stop - start) / (num - 1)
        elif what == 'num':
            num = int(math.ceil((stop - start + step) / step))
            stop = start + step * (num - 1)
        else:
            stop = start + step * (num - 1)
        if not isinstance(num, (int, np.integer)):
            raise TypeError("num must be an integer")
        self.start, self.stop, self.step, self.num = float(start), float(stop), float(step), int(num)
        self._values = None

    def __repr__(self):
        return f"TimeAxis: start={self.start:g}, stop={self.stop:g}, step={self.step:g}, num={self.num:g}"

    __str__ = __repr__

    def _rebuild(self):
        return TimeAxis(start=self.start, stop=self.stop, num=self.num)

    @property
    def time_values(self):
        if self._values is None:
            self._values = np.linspace(self.start, self.stop, self.num)
        return self._values


class PointSource(SparseTimeFunction):
    """A set of `npoint` off-grid points, each carrying one time series sampled on `time_range`."""

    @classmethod
    def __args_setup__(cls, *args, **kwargs):
        axis = kwargs.get('time_range')
        if axis is None:
            raise TypeError("PointSource needs `time_range`")
        kwargs['nt'] = axis.num
        if kwargs.get('npoint', kwargs.get('npoint_global')) is None:
            coords = kwargs.get('coordinates', kwargs.get('coordinates_data'))
            if coords is None:
                raise TypeError("Need either `npoint` or `coordinates`")
            kwargs['npoint'] = int(np.shape(coords)[0])
        return args, kwargs

    def __init_finalize__(self, *args, **kwargs):
        axis = kwargs.pop('time_range')
        initial = kwargs.pop('data', None)
        kwargs.setdefault('time_order', 2)
        super().__init_finalize__(*args, **kwargs)
        self._time_range = axis._rebuild()
        if initial is not None:
            self.data[:] = initial

    @property
    def time_range(self):
        return self._time_range

    @property
    def time_values(self):
        return self._time_range.time_values


# the reference uses one class for sources, receivers and shot records
Receiver = PointSource
Shot = PointSource


class WaveletSource(PointSource):
    """A PointSource whose traces are initialised with an analytic wavelet of peak frequency `f0`
    (kHz), amplitude `a` and delay `t0` (ms)."""

    @classmethod
    def __args_setup__(cls, *args, **kwargs):
        kwargs.setdefault('npoint', 1)
        return super().__args_setup__(*args, **kwargs)

    def __init_finalize__(self, *args, **kwargs):
        super().__init_finalize__(*args, **kwargs)
        self.f0, self.a, self.t0 = kwargs.get('f0'), kwargs.get('a'), kwargs.get('t0')
        if not self.alias:
            trace = self.wavelet
            for p in range(self.npoint):
                self.data[:, p] = trace

    @property
    def wavelet(self):
        raise NotImplementedError("subclasses define the wavelet")


class RickerSource(WaveletSource):
    """Ricker wavelet (1 - 2 r^2) exp(-r^2), r = pi f0 (t - t0), default t0 = 1/f0
    (source.py:284-289)."""

    @property
    def wavelet(self):
        delay = self.t0 if self.t0 else 1.0 / self.f0
        amp = self.a if self.a else 1.0
        r2 = (np.pi * self.f0 * (self.time_values - delay)) ** 2
        return amp * (1.0 - 2.0 * r2) * np.exp(-r2)


class GaborSource(WaveletSource):
    """Gabor wavelet exp(-2 s^2) cos(2 pi s), s = (t - t0) f0 / 2, default t0 = 3/f0."""

    @property
    def wavelet(self):
        half = 0.5 * self.f0
        delay = self.t0 if self.t0 else 1.5 / half
        amp = self.a if self.a else 1.0
        s = (self.time_values - delay) * half
        return amp * np.exp(-2.0 * s * s) * np.cos(2.0 * np.pi * s)
