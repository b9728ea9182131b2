"""Sources, receivers and the time axis (mirror of examples/seismic/source.py:24-289)."""
from functools import cached_property

import numpy as np

from ..sparse import SparseTimeFunction

__all__ = ['TimeAxis', 'PointSource', 'Receiver', 'Shot', 'WaveletSource', 'RickerSource',
           'GaborSource']


class TimeAxis:
    """start/step/num/stop with exactly three given (source.py:24-86): when `num` is derived,
    num = ceil((stop - start + step)/step) and stop is re-derived."""

    def __init__(self, start=None, step=None, num=None, stop=None):
        try:
            if start is None:
                start = step * (1 - num) + stop
            elif step is None:
                step = (stop - start) / (num - 1)
            elif num is None:
                num = int(np.ceil((stop - start + step) / step))
                stop = step * (num - 1) + start
            elif stop is None:
                stop = step * (num - 1) + start
            else:
                raise ValueError
        except Exception:
            raise ValueError("Three of args start, step, num and stop may be set") from None
        if not isinstance(num, int):
            raise TypeError("input argument must be of type int")
        self.start, self.stop, self.step, self.num = float(start), float(stop), float(step), int(num)

    def __str__(self):
        return f'TimeAxis: start={self.start:g}, stop={self.stop:g}, step={self.step:g}, num={self.num:g}'

    def _rebuild(self):
        return TimeAxis(start=self.start, stop=self.stop, num=self.num)

    @cached_property
    def time_values(self):
        return np.linspace(self.start, self.stop, self.num)


class PointSource(SparseTimeFunction):
    """A set of sparse points carrying a time series each (source.py:90-186)."""

    @classmethod
    def __args_setup__(cls, *args, **kwargs):
        kwargs['nt'] = kwargs['time_range'].num
        npoint = kwargs.get('npoint', kwargs.get('npoint_global'))
        if npoint is None:
            coordinates = kwargs.get('coordinates', kwargs.get('coordinates_data'))
            if coordinates is None:
                raise TypeError("Need either `npoint` or `coordinates`")
            kwargs['npoint'] = np.asarray(coordinates).shape[0]
        return args, kwargs

    def __init_finalize__(self, *args, **kwargs):
        time_range = kwargs.pop('time_range')
        data = kwargs.pop('data', None)
        kwargs.setdefault('time_order', 2)
        super().__init_finalize__(*args, **kwargs)
        self._time_range = time_range._rebuild()
        if data is not None:
            self.data[:] = data

    @cached_property
    def time_values(self):
        return self._time_range.time_values

    @property
    def time_range(self):
        return self._time_range


Receiver = PointSource
Shot = PointSource


class WaveletSource(PointSource):
    @classmethod
    def __args_setup__(cls, *args, **kwargs):
        kwargs.setdefault('npoint', 1)
        return super().__args_setup__(*args, **kwargs)

    def __init_finalize__(self, *args, **kwargs):
        super().__init_finalize__(*args, **kwargs)
        self.f0 = kwargs.get('f0')
        self.a = kwargs.get('a')
        self.t0 = kwargs.get('t0')
        if not self.alias:
            for p in range(kwargs['npoint']):
                self.data[:, p] = self.wavelet

    @property
    def wavelet(self):
        raise NotImplementedError


class RickerSource(WaveletSource):
    """r = pi f0 (t - t0); (1 - 2 r^2) exp(-r^2), t0 = 1/f0 (source.py:284-289)."""

    @property
    def wavelet(self):
        t0 = self.t0 or 1 / self.f0
        a = self.a or 1
        r = np.pi * self.f0 * (self.time_values - t0)
        return a * (1 - 2. * r ** 2) * np.exp(-r ** 2)


class GaborSource(WaveletSource):
    @property
    def wavelet(self):
        agauss = 0.5 * self.f0
        tcut = self.t0 or 1.5 / agauss
        s = (self.time_values - tcut) * agauss
        a = self.a or 1
        return a * np.exp(-2 * s ** 2) * np.cos(2 * np.pi * s)
