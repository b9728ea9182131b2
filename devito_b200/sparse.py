"""Sparse (off-grid) functions: sources and receivers.

Mirrors `SparseTimeFunction` (devito/types/sparse.py:1039) with its `inject` / `interpolate`
front-ends (:1122-1178) and the host-side fp64 tabulation of base cell indices and per-dim
weights (devito/operations/interpolators.py:674-718: `_cell_indices`, `_linear_weights`,
`_sinc_weights`; Kaiser table :862-864).  Under slab decomposition points are routed to every
rank whose owned-plus-support range contains them (reference: `_dist_scatter`,
devito/types/sparse.py:608-666 with support from :303-318) — here all ranks hold the full
point list and the kernels' bound guards do the selection, with receivers owned by exactly
one rank for the final gather.
"""
import numpy as np

from .symbolics import Index, as_expr
from .types import DiscreteFunction, DefaultDimension

__all__ = ['SparseFunction', 'SparseTimeFunction', 'Injection', 'Interpolation',
           '_default_radius']

_default_radius = {'linear': 1, 'sinc': 4, 'nearest': 0}

# Kaiser window parameter b(r), Hicks 2002 table 1 (interpolators.py:862-864)
_b_table = {2: 2.94, 3: 4.53, 4: 4.14, 5: 5.26, 6: 6.40, 7: 7.51, 8: 8.56, 9: 9.56, 10: 10.64}


def as_fp64_decimal(v):
    """fp64 value of the shortest decimal that round-trips `v`
    (devito/tools/dtypes_lowering.py:22-29)."""
    return np.float64(np.format_float_positional(v, unique=True, trim='0'))


class _Coordinates:
    """Stand-in for the reference's coordinates SubFunction: exposes `.data`."""

    def __init__(self, npoint, ndim, dtype):
        self._data = np.zeros((npoint, ndim), dtype=dtype)

    @property
    def data(self):
        return self._data

    @data.setter
    def data(self, v):
        self._data[:] = v

    @property
    def shape(self):
        return self._data.shape


class _SparseOp:
    """Common base of Injection/Interpolation: behaves like a list element so that
    `eqns + src_term + rec_term` works (interpolators.py:120-124)."""

    def __add__(self, other):
        return [self] + list(other)

    def __radd__(self, other):
        return list(other) + [self]


def _check_radius(sfunction, *exprs):
    """devito/operations/interpolators.py:28-37 `check_radius`: the support of a point reaches `r` cells
    beyond the domain, so every dense function involved needs at least `r` halo points."""
    r = sfunction.r
    orders = {n.function.space_order for e in exprs for n in as_expr(e).preorder()
              if getattr(n, 'is_Access', False) and not getattr(n.function, 'is_SparseFunction', False)}
    so = min(orders or {r})
    if so < r:
        raise ValueError(f"Space order {so} too small for interpolation r {r}")


class Injection(_SparseOp):
    """`field[cells] += weights * expr` for every sparse point (interpolators.py:127-189)."""

    def __init__(self, sfunction, field, expr, implicit_dims=None):
        self.sfunction = sfunction
        self.fields = tuple(field) if isinstance(field, (tuple, list)) else (field,)
        self.exprs = tuple(expr) if isinstance(expr, (tuple, list)) else (expr,) * len(self.fields)
        self.exprs = tuple(as_expr(e) for e in self.exprs)
        _check_radius(sfunction, *self.fields, *self.exprs)
        self.implicit_dims = implicit_dims

    def __repr__(self):
        return f"Injection({self.sfunction.name} -> {', '.join(map(repr, self.fields))})"


class Interpolation(_SparseOp):
    """`sfunction[time, p] = sum weights * expr[cells]` (interpolators.py:127-189)."""

    def __init__(self, sfunction, expr, increment=False, implicit_dims=None):
        self.sfunction = sfunction
        self.expr = as_expr(expr)
        _check_radius(sfunction, self.expr)
        self.increment = increment
        self.implicit_dims = implicit_dims

    def __repr__(self):
        return f"Interpolation({self.expr!r} -> {self.sfunction.name})"


class SparseFunction(DiscreteFunction):
    is_SparseFunction = True

    __rkwargs__ = ('name', 'npoint', 'grid', 'coordinates', 'space_order', 'dtype',
                   'interpolation', 'r')

    def __init_finalize__(self, *args, **kwargs):
        self._name = kwargs['name']
        self._alias = kwargs.get('alias', False)
        self._grid = kwargs['grid']
        self._space_order = int(kwargs.get('space_order', 0))
        npoint = kwargs.get('npoint', kwargs.get('npoint_global'))
        coords = kwargs.get('coordinates', kwargs.get('coordinates_data'))
        if npoint is None:
            if coords is None:
                raise TypeError("Need either `npoint` or `coordinates`")
            npoint = np.asarray(coords).shape[0]
        self.npoint = int(npoint)
        dtype = kwargs.get('dtype')
        self._dtype = np.dtype(dtype if dtype is not None else self._grid.dtype).type
        self._sparse_dim = kwargs.get('dimension') or DefaultDimension(f'p_{self._name}')
        self.interpolation = kwargs.get('interpolation', 'linear')
        r = kwargs.get('r')
        self._radius = r or _default_radius[self.interpolation]
        if self.interpolation == 'sinc' and not (2 <= self._radius <= 10):
            raise ValueError("'sinc' interpolator requires 2 <= r <= 10")
        if self.interpolation == 'linear' and self._radius != 1:
            self._radius = 1
        self._coordinates = _Coordinates(self.npoint, self._grid.dim, self._dtype)
        if coords is not None:
            self._coordinates.data[:] = np.asarray(coords)
        self._setup_dims_shape(kwargs)
        self._halo = tuple((0, 0) for _ in self._dimensions)
        self._storage = None
        self._indices = tuple(Index(d, 0) for d in self._dimensions)

    def _setup_dims_shape(self, kwargs):
        self._dimensions = (self._sparse_dim,)
        self._shape = (self.npoint,)

    @property
    def r(self):
        return self._radius

    radius = r

    @property
    def coordinates(self):
        return self._coordinates

    @property
    def coordinates_data(self):
        return self._coordinates.data

    # -- symbolic front-ends -------------------------------------------------------------------
    def inject(self, field, expr, implicit_dims=None):
        """devito/types/sparse.py:1150-1178"""
        return [Injection(self, field, expr, implicit_dims=implicit_dims)]

    def interpolate(self, expr, increment=False, self_subs=None, implicit_dims=None):
        """devito/types/sparse.py:1122-1148"""
        return [Interpolation(self, expr, increment=increment, implicit_dims=implicit_dims)]

    # -- host-side tabulation (fp64), interpolators.py:674-718 ---------------------------------
    def _positions_fp64(self, origin=None):
        grid = self._grid
        spacing = np.array([as_fp64_decimal(h) for h in grid.spacing])
        origin = np.array([as_fp64_decimal(o) for o in (origin if origin is not None else grid.origin)])
        c64 = np.asarray(self._coordinates.data, dtype=np.float64)
        return (c64 - origin) / spacing

    def tabulate(self, origin=None):
        """Returns (gp int32 (npoint, ndim), [w_d f32 (npoint, 2r)] per dim). `gp` is relative
        to `origin` (default: the global grid origin)."""
        pos = self._positions_fp64(origin)
        gp = np.floor(pos).astype(np.int32)
        frac = pos - np.floor(pos)
        r = self._radius
        ws = []
        for j in range(self._grid.dim):
            if self.interpolation == 'sinc':
                from scipy.special import i0
                b = _b_table[r]
                b0 = i0(b)
                data = np.zeros((self.npoint, 2 * r), dtype=self._dtype)
                for ri in range(2 * r):
                    rpos = ri - r + 1 - frac[:, j]
                    data[:, ri] = i0(b * np.sqrt(1 - (rpos / r) ** 2)) / b0 * np.sinc(rpos)
            elif self.interpolation == 'linear':
                data = np.empty((self.npoint, 2), dtype=self._dtype)
                data[:, 0] = 1.0 - frac[:, j]
                data[:, 1] = frac[:, j]
            else:
                raise NotImplementedError(f"interpolation={self.interpolation!r}")
            ws.append(np.ascontiguousarray(data))
        return np.ascontiguousarray(gp), ws


class SparseTimeFunction(SparseFunction):
    is_SparseTimeFunction = True

    __rkwargs__ = tuple(SparseFunction.__rkwargs__) + ('nt', 'time_order')

    def __init_finalize__(self, *args, **kwargs):
        self.nt = int(kwargs['nt'])
        self.time_order = int(kwargs.get('time_order', 1))
        self.time_dim = kwargs.get('time_dim') or kwargs['grid'].time_dim
        super().__init_finalize__(*args, **kwargs)

    def _setup_dims_shape(self, kwargs):
        self._dimensions = (self.time_dim, self._sparse_dim)
        self._shape = (self.nt, self.npoint)
