"""Reference-side binding: the Operator backend a devito maintainer would add as `devito/core/b200.py`.

    import devito                       # the REAL reference (devito @ 436199c)
    import devito_b200.refplugin        # registers (Blackwell, 'advanced', 'cuda')
    configuration['platform'], configuration['language'] = 'blackwell', 'cuda'
    # ... examples/seismic run unchanged: the wave propagators execute in libb200stencil.so

The open-source reference has no operator for `(Blackwell, 'advanced', 'cuda')`
(devito/operator/registry.py:33-57 raises "Cannot compile an Operator for ..."); this module fills
that slot with `B200CudaOperator`:

* `_build` lowers the expressions with the reference's own CPU pipeline (so `op.parameters`,
  `op.arguments()` and argument checking stay the reference's, devito/operator/operator.py:218-315,
  :608-757), and — next to it — restates the user's equations in `devito_b200`'s expression tree
  and hands them to that package's pattern recogniser. Operators that are not one of the
  propagation schemes of libb200stencil (the set-up operators of examples/seismic: `initdamp`,
  `smooth`, `norm` ...) simply stay CPU operators and JIT-compile with gcc like before.
* `apply` of a recognised operator calls `self.arguments(**kwargs)` — the reference builds its own
  `struct dataobj` for every array (devito/types/dense.py:737-777) and its own coordinate tables
  `src_gp/src_wx/...` (devito/operations/interpolators.py:674-718) — and passes THOSE STRUCTS to
  `b2_iso_forward` / `b2_tti_forward` (include/b200stencil.h). The library stages host arrays in and
  out inside the call, like the reference's device path does around its kernel
  (devito/passes/iet/definitions.py:636-671).

Nothing here is imported by `devito_b200` itself; importing this module needs the reference
(`/root/reference`, or its unmodified install under `baseline/_ref`) on `sys.path`.
"""
import ctypes
import os
import time as _time
from fractions import Fraction

import numpy as np
import sympy

import devito
from devito import configuration
from devito.arch.archinfo import Blackwell, get_platform
from devito.core.cpu import Cpu64AdvOmpOperator
from devito.data.allocators import default_allocator
from devito.operator.registry import operator_registry
from devito.operations.interpolators import Injection as _RefInjection, Interpolation as _RefInterpolation
from devito.types.basic import AbstractFunction
from devito.types import Eq as _RefEq

import devito_b200 as dv
from devito_b200 import _lib as L_
from devito_b200.equation import FreeSurface
from devito_b200.symbolics import Access, Index, Number, Call, _np_funcs

__all__ = ['B200CudaOperator', 'register', 'activate']


class _Unconvertible(Exception):
    pass


# ---------------------------------------------------------------------------------------------
# the user's (reference) expressions restated in devito_b200's tree — for RECOGNITION only;
# no data is copied, the shadow objects never allocate
# ---------------------------------------------------------------------------------------------
class _Shadow:
    def __init__(self):
        self.grids, self.fns, self.consts, self.sparse = {}, {}, {}, {}

    def grid(self, g):
        if id(g) in self.grids:
            return self.grids[id(g)]
        subs = []
        sd = g.subdomains
        if 'physdomain' in sd or 'fsdomain' in sd:
            from devito_b200.seismic.model import PhysicalDomain, FSDomain
            fs = 'fsdomain' in sd
            so = int(getattr(sd['fsdomain'], 'size', 0)) if fs else int(getattr(sd.get('physdomain'), 'so', 0))
            subs.append(PhysicalDomain(so, fs=fs))
            if fs:
                subs.append(FSDomain(so))
        sg = dv.Grid(shape=tuple(g.shape), extent=tuple(float(e) for e in g.extent),
                     origin=tuple(float(o) for o in g.origin), dtype=g.dtype, subdomains=subs)
        if [d.name for d in g.dimensions] != [d.name for d in sg.dimensions]:
            raise _Unconvertible("grid dimensions are not named x, y, z")
        self.grids[id(g)] = sg
        return sg

    def function(self, f):
        if id(f) in self.fns:
            return self.fns[id(f)]
        if f.is_SparseTimeFunction:
            sg = self.grid(f.grid)
            kind = type(f.interpolator).__name__
            interp = {'LinearInterpolator': 'linear', 'SincInterpolator': 'sinc'}.get(kind)
            if interp is None:
                raise _Unconvertible(f"sparse function with {kind}")
            s = dv.SparseTimeFunction(name=f.name, grid=sg, npoint=int(f.npoint), nt=int(f.nt),
                                      interpolation=interp, r=int(f.r))
        elif f.is_TimeFunction:
            sg = self.grid(f.grid)
            s = dv.TimeFunction(name=f.name, grid=sg, time_order=int(f.time_order), space_order=int(f.space_order),
                                save=int(f.save) if isinstance(f.save, (int, np.integer)) else None)
            if f.save is not None and not isinstance(f.save, (int, np.integer)):
                raise _Unconvertible("buffered / conditional saving")
        elif f.is_Function:
            s = dv.Function(name=f.name, grid=self.grid(f.grid), space_order=int(f.space_order))
        else:
            raise _Unconvertible(f"function type {type(f).__name__}")
        if f.is_DiscreteFunction and getattr(f, 'staggered', None) not in (None, ()) and \
                any(v != 0 for v in getattr(f.staggered, '_getters', {}).values() if isinstance(v, int)):
            raise _Unconvertible("staggered function")
        self.fns[id(f)] = s
        return s

    def constant(self, c):
        if id(c) not in self.consts:
            self.consts[id(c)] = dv.Constant(name=c.name, value=float(c.data))
        return self.consts[id(c)]

    # -- expressions -------------------------------------------------------------------------------
    def symbol(self, s, grid):
        for sp in grid.spacing_symbols:
            if sp.name == s.name:
                return sp
        if s.name == grid.stepping_dim.spacing.name:
            return grid.stepping_dim.spacing
        raise _Unconvertible(f"free symbol {s}")

    def access(self, e):
        f = e.function
        sf = self.function(f)
        if all(i == d for i, d in zip(e.indices, f.dimensions)):
            return sf
        idx = []
        for i, d, sd in zip(e.indices, f.dimensions, sf.dimensions):
            off = sympy.sympify(i) - d
            if off == 0:
                idx.append(Index(sd, 0))
                continue
            if d.is_Time or d.is_Stepping:
                sp = d.spacing
            else:
                sp = d.spacing
            k = sympy.nsimplify(sympy.simplify(off / sp), rational=True)
            if not k.is_Rational:
                raise _Unconvertible(f"index {i} of {f.name}")
            idx.append(Index(sd, Fraction(int(k.p), int(k.q))))
        return Access(sf, idx)

    def expr(self, e, grid):
        e = sympy.sympify(e)
        if isinstance(e, AbstractFunction):
            return self.access(e)
        if getattr(e, 'is_Constant', False) and hasattr(e, 'data') and not e.is_Number:
            return self.constant(e)
        if e.is_Number:
            if e.is_Integer:
                return Number(int(e))
            return Number(float(e))
        if e.is_Symbol:
            return self.symbol(e, grid)
        if e.is_Add:
            out = self.expr(e.args[0], grid)
            for a in e.args[1:]:
                out = out + self.expr(a, grid)
            return out
        if e.is_Mul:
            out = self.expr(e.args[0], grid)
            for a in e.args[1:]:
                out = out * self.expr(a, grid)
            return out
        if e.is_Pow:
            return self.expr(e.base, grid) ** self.expr(e.exp, grid)
        name = type(e).__name__
        if name in _np_funcs and len(e.args) == 1:
            return Call(name, self.expr(e.args[0], grid))
        raise _Unconvertible(f"expression node {name}")


def _restate(expressions, kwargs):
    """devito_b200.Operator over the restated equations, or None when they are not on its fast path."""
    sh = _Shadow()
    items, eqs = [], []
    grid = None
    for e in expressions:
        for f in (e.lhs.function,) if isinstance(e, _RefEq) and hasattr(e.lhs, 'function') else ():
            if getattr(f, 'grid', None) is not None:
                grid = grid or f.grid
    if grid is None:
        return None
    sg = sh.grid(grid)
    fs_rows = []
    for e in expressions:
        if isinstance(e, _RefInjection):
            sf = sh.function(e.interpolator.sfunction)
            if isinstance(e.field, (tuple, list, sympy.Tuple)):      # TTI: one source into u and v
                field = [sh.expr(f, sg) for f in e.field]
                ex = ([sh.expr(x, sg) for x in e.expr] if isinstance(e.expr, (tuple, list, sympy.Tuple))
                      else sh.expr(e.expr, sg))
            else:
                field, ex = sh.expr(e.field, sg), sh.expr(e.expr, sg)
            items += sf.inject(field=field, expr=ex)
        elif isinstance(e, _RefInterpolation):
            sf = sh.function(e.interpolator.sfunction)
            items += sf.interpolate(expr=sh.expr(e.expr, sg), increment=bool(e.increment))
        elif isinstance(e, _RefEq):
            sd = e.subdomain
            name = getattr(sd, 'name', None)
            if name == 'fsdomain':
                fs_rows.append(e)           # examples/seismic/acoustic/operators.py:5-47 `freesurface`
                continue
            ssd = sg.subdomains[name] if name in ('physdomain', 'interior') else None
            if sd is not None and ssd is None and name not in (None, 'domain'):
                raise _Unconvertible(f"subdomain {name}")
            cls = dv.Inc if type(e).__name__ == 'Inc' else dv.Eq
            q = cls(sh.expr(e.lhs, sg), sh.expr(e.rhs.evaluate, sg), subdomain=ssd)
            eqs.append((e, q))
            items.append(q)
        else:
            raise _Unconvertible(f"{type(e).__name__} in the operator")
    if fs_rows:
        # the reference's free surface: [mirrored copy of the update on fsdomain, `u.forward[z=0] = 0`]
        if len(fs_rows) != 2 or len(eqs) != 1:
            raise _Unconvertible("free-surface equations next to more than one update")
        mirrored, zero = fs_rows
        ref_update, q = eqs[0]
        if mirrored.lhs != ref_update.lhs or zero.rhs != 0 or zero.lhs.function is not ref_update.lhs.function:
            raise _Unconvertible("unexpected free-surface equations")
        if not mirrored.rhs.has(devito.sign):
            raise _Unconvertible("free-surface equation without mirrored taps")
        items.insert(items.index(q) + 1, FreeSurface(q, sg.subdomains['fsdomain']))
    subs = {}
    for k, v in (kwargs.get('subs') or {}).items():
        subs[sh.symbol(k, sg)] = float(v)
    op = dv.Operator(items, subs=subs, name=kwargs.get('name', 'Kernel'))
    if op.backend != 'cuda-sm100a' or op._plan.get('kind') not in ('iso', 'tti'):
        return None
    return op


# ---------------------------------------------------------------------------------------------
# the Operator
# ---------------------------------------------------------------------------------------------
def _addr(cobj):
    """Address of the struct behind whatever ctypes handle `op.arguments()` holds (byref / pointer)."""
    if hasattr(cobj, '_obj'):
        return ctypes.addressof(cobj._obj)
    if hasattr(cobj, 'contents'):
        return ctypes.addressof(cobj.contents)
    return ctypes.addressof(cobj)


class B200CudaOperator(Cpu64AdvOmpOperator):
    """`(Blackwell, 'advanced', 'cuda')`: wave propagators run in libb200stencil.so, everything else
    stays on the reference's CPU path."""

    @classmethod
    def _normalize_kwargs(cls, **kwargs):
        # the CPU lowering (argument machinery, set-up operators) is OpenMP C on the host cores
        kwargs = dict(kwargs)
        kwargs['options'] = dict(kwargs['options'], openmp=True)
        kwargs['b200_target'] = (kwargs['platform'], kwargs['language'])
        kwargs['platform'] = get_platform()
        kwargs['language'] = 'openmp'
        kwargs['compiler'] = configuration['compiler'].__new_with__(platform=kwargs['platform'], language='openmp',
                                                                    mpi=configuration['mpi'])
        kwargs['allocator'] = default_allocator(f"{kwargs['compiler'].__class__.__name__}.openmp.{kwargs['platform']}")
        return super()._normalize_kwargs(**kwargs)

    @classmethod
    def _build(cls, expressions, **kwargs):
        kwargs.pop('b200_target', None)
        op = super()._build(expressions, **kwargs)
        try:
            op._b200 = _restate(expressions, kwargs)
            op._b200_why = None
        except (_Unconvertible, ValueError, TypeError, KeyError, AttributeError) as e:
            op._b200, op._b200_why = None, f"{type(e).__name__}: {e}"
        return op

    @property
    def backend(self):
        return 'cuda-sm100a' if getattr(self, '_b200', None) is not None else 'reference-cpu'

    def arguments(self, **kwargs):
        # thread-count defaults ask `configuration['platform']` — the GPU here — for its "cores"
        # (devito/types/parallel.py:66-70): give the host's instead
        names = {p.name for p in self.parameters}
        host = get_platform()
        for n in ('nthreads', 'nthreads_nonaffine'):
            if n in names:
                kwargs.setdefault(n, int(os.environ.get('OMP_NUM_THREADS', host.cores_physical)))
        return super().arguments(**kwargs)

    def apply(self, **kwargs):
        shadow = getattr(self, '_b200', None)
        if shadow is None:
            return super().apply(**kwargs)
        kwargs.pop('autotune', None)
        # the reference's own argument processing: defaults, overrides, checks, struct dataobj
        with self._profiler.timer_on('arguments-preprocess'):
            args = self.arguments(**kwargs)
        p = shadow._plan
        ov = {'resident': False, 'dt': float(args['dt']), 'time_m': int(args['time_m']), 'time_M': int(args['time_M'])}
        for d in p['grid'].dimensions:
            ov[d.min_name], ov[d.max_name] = int(args[d.min_name]), int(args[d.max_name])
        hold = []

        def dataobj(name):
            fo = L_.ForeignDataobj(_addr(args[name]), keep=args[name])
            hold.append(fo)
            return fo

        fields = [p['u']] + ([p['v']] if p['kind'] == 'tti' else [])
        for f in fields + [p.get(k) for k in ('damp', 'usave', 'grad', 'born_U', 'born_dm', 'snap')]:
            if f is not None:
                ov[f.name] = dataobj(f.name)
        if p['kind'] == 'iso':
            kind, obj = p['m_role']
            if obj is not None:
                ov[obj.name] = dataobj(obj.name) if kind.endswith('_f') else float(args[obj.name])
        else:
            for n, c in p['consts'].items():
                ov[n] = dataobj(n) if hasattr(c, 'space_order') else float(args[n])
        for key in ('src', 'rec'):
            sf = p.get(key)
            if sf is None:
                continue
            nd = p['grid'].dim
            # weight tables: `<name>_wx..` (linear, interpolators.py:674-718), `wsincrp_<name>x..` (sinc, :721-789)
            ws = [dataobj(f"{sf.name}_w{'xyz'[i]}" if f"{sf.name}_w{'xyz'[i]}" in args else f"wsincrp_{sf.name}{'xyz'[i]}")
                  for i in range(nd)]
            ov[sf.name] = L_.ForeignSparse(sf.name, dataobj(sf.name), dataobj(f"{sf.name}_gp"), ws,
                                           p_m=int(args[f'p_{sf.name}_m']), p_M=int(args[f'p_{sf.name}_M']), r=sf.r)
        t0 = _time.perf_counter()
        summary = shadow.apply(**ov)
        self._b200_last = {'wall': _time.perf_counter() - t0, 'arguments': sorted(ov)}
        return summary


def register():
    """`operator_registry.add(...)` exactly like devito/core/__init__.py:37-83 does for the stock backends."""
    operator_registry.add(B200CudaOperator, Blackwell, 'advanced', 'cuda')
    operator_registry.add(B200CudaOperator, Blackwell, 'noop', 'cuda')
    operator_registry.add(B200CudaOperator, Blackwell, 'advanced-fsg', 'cuda')


def activate():
    """What `DEVITO_PLATFORM=blackwell DEVITO_LANGUAGE=cuda` does."""
    register()
    configuration['platform'] = 'blackwell'
    configuration['language'] = 'cuda'


register()
