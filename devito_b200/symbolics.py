"""A small expression layer: just enough symbolic machinery for the wave-propagation DSL.

The reference builds its DSL on SymPy subclasses (devito/types/basic.py,
devito/finite_differences/differentiable.py) and lowers expressions through a compiler.
Here expressions are a light-weight immutable tree (numbers, symbols, dimension symbols,
function accesses with affine indices, +, *, **, elementary calls and lazy finite-difference
`Derivative` nodes).  The tree is consumed two ways only:
  * the pattern recogniser in `operator.py` extracts the linear stencil of a time update and
    matches it against the hand-written CUDA kernels;
  * the NumPy interpreter in `interpreter.py` evaluates everything else (set-up operators).

Finite-difference semantics follow the reference exactly (file:line cited at each rule):
index sets from `generate_indices` (devito/finite_differences/tools.py:244-308), Taylor
weights from `sympy.finite_diff_weights` rounded to 9 significant digits
(devito/finite_differences/finite_difference.py:27, 185-187).
"""
import math
import numbers
from fractions import Fraction
from functools import lru_cache

import numpy as np

__all__ = ['div', 'grad', 'Expr', 'Number', 'Symbol', 'Add', 'Mul', 'Pow', 'Call', 'Access', 'Index',
           'Derivative', 'as_expr', 'sin', 'cos', 'sqrt', 'Abs', 'sign', 'exp', 'floor', 'INT',
           'fd_weights', 'fd_offsets', 'retrieve_functions', 'retrieve_derivatives',
           'linear_terms', 'NonLinear']

_PRECISION = 9   # devito/finite_differences/finite_difference.py:27


def as_expr(obj):
    if isinstance(obj, Expr):
        return obj
    if isinstance(obj, bool):
        return Number(int(obj))
    if isinstance(obj, (numbers.Integral, np.integer)):
        return Number(int(obj))
    if isinstance(obj, Fraction):
        return Number(obj)
    if isinstance(obj, (numbers.Real, np.floating)):
        return Number(float(obj))
    try:                                    # sympy numbers (e.g. Rational from user code)
        import sympy
        if isinstance(obj, sympy.Rational):
            return Number(Fraction(int(obj.p), int(obj.q)))
        if isinstance(obj, sympy.Number):
            return Number(float(obj))
    except ImportError:                     # pragma: no cover
        pass
    raise TypeError(f"cannot convert {type(obj).__name__} to an expression")


class Expr:
    """Base class of all expression nodes."""
    __array_ufunc__ = None          # numpy scalars defer to our reflected operators
    is_Number = False
    is_Symbol = False
    is_Dimension = False
    is_Access = False
    is_Derivative = False
    is_Constant = False
    args = ()

    # -- structural identity -----------------------------------------------------------------
    def _key(self):
        return (type(self).__name__,) + tuple(a._key() if isinstance(a, Expr) else a for a in self.args)

    def __eq__(self, other):
        if self is other:
            return True
        if not isinstance(other, Expr):
            try:
                other = as_expr(other)
            except TypeError:
                return False
        return self._key() == other._key()

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash(self._key())

    def __bool__(self):
        return True

    # -- arithmetic ---------------------------------------------------------------------------
    def __add__(self, o): return Add.make(self, o)
    def __radd__(self, o): return Add.make(o, self)
    def __sub__(self, o): return Add.make(self, Mul.make(-1, o))
    def __rsub__(self, o): return Add.make(o, Mul.make(-1, self))
    def __mul__(self, o): return Mul.make(self, o)
    def __rmul__(self, o): return Mul.make(o, self)
    def __truediv__(self, o): return Mul.make(self, Pow.make(o, -1))
    def __rtruediv__(self, o): return Mul.make(o, Pow.make(self, -1))
    def __pow__(self, o): return Pow.make(self, o)
    def __rpow__(self, o): return Pow.make(o, self)
    def __neg__(self): return Mul.make(-1, self)
    def __pos__(self): return self
    def __abs__(self): return Abs(self)

    # -- traversal ---------------------------------------------------------------------------
    def _rebuild(self, *args):
        return type(self)(*args)

    def map_leaves(self, fn):
        """Rebuild the tree bottom-up with `fn` applied to every leaf / Access node."""
        if not self.args:
            return fn(self)
        return self._rebuild(*[a.map_leaves(fn) if isinstance(a, Expr) else a for a in self.args])

    def preorder(self):
        yield self
        for a in self.args:
            if isinstance(a, Expr):
                yield from a.preorder()

    @property
    def free_symbols(self):
        return {e for e in self.preorder() if e.is_Symbol}

    # -- substitution ------------------------------------------------------------------------
    def subs(self, *args):
        if len(args) == 2:
            mapping = {args[0]: args[1]}
        else:
            mapping = dict(args[0])
        mapping = {_sub_key(k): v for k, v in mapping.items()}
        return _substitute(self, mapping)

    xreplace = subs

    def _subs(self, old, new):
        return self.subs({old: new})

    # -- finite differences ------------------------------------------------------------------
    @property
    def evaluate(self):
        """Expand all Derivative nodes into explicit weighted sums."""
        if not self.args:
            return self
        return self._rebuild(*[a.evaluate if isinstance(a, Expr) else a for a in self.args])

    def _shift(self, dim, k):
        """Translate every function access by k grid points along `dim`."""
        if not self.args:
            return self
        return self._rebuild(*[a._shift(dim, k) if isinstance(a, Expr) else a for a in self.args])

    @property
    def _space_dims(self):
        for e in self.preorder():
            if e.is_Access and e.function.grid is not None:
                return e.function.grid.dimensions
        return ()

    @property
    def _time_dim(self):
        for e in self.preorder():
            if e.is_Access and getattr(e.function, 'time_dim', None) is not None:
                return e.function.time_dim
        return None

    @property
    def _default_fd_order(self):
        orders = [e.function.space_order for e in self.preorder()
                  if e.is_Access and getattr(e.function, 'space_order', None) is not None]
        return min(orders) if orders else 1

    def _dim_by_name(self, name):
        for d in self._space_dims:
            if d.name == name:
                return d
        return None

    def __getattr__(self, name):
        # derivative shortcuts `dx`, `dy2`, `dxl`, ... (devito/finite_differences/tools.py:83-142)
        if name.startswith('__') or name.startswith('_'):
            raise AttributeError(name)
        if name.startswith('d') and len(name) >= 2:
            body = name[1:]
            side = None
            if body and body[-1] in 'lrc' and len(body) > 1 and self._dim_by_name(body[:-1]) is not None:
                side = {'l': -1, 'r': 1, 'c': 0}[body[-1]]
                body = body[:-1]
            order = 1
            if body and body[-1].isdigit():
                order = int(body[-1])
                body = body[:-1]
            d = self._dim_by_name(body)
            if d is not None:
                return Derivative(self, (d, order), side=side)
            t = self._time_dim
            if t is not None and body == 't':
                return Derivative(self, (t, order), side=side)
        raise AttributeError(f"{type(self).__name__!r} object has no attribute {name!r}")

    @property
    def laplace(self):
        """Sum of second derivatives over the space dimensions
        (devito/finite_differences/differentiable.py:334)."""
        out = Number(0)
        for d in self._space_dims:
            out = out + Derivative(self, (d, 2))
        return out

    def laplacian(self, **kwargs):
        return self.laplace

    def biharmonic(self, weight=1):
        """Weighted biharmonic operator `laplace(weight * laplace(self))`
        (devito/finite_differences/differentiable.py:442-449)."""
        inner = self.laplace * weight
        out = Number(0)
        for d in self._space_dims:
            out = out + Derivative(inner, (d, 2))
        return out


def _sub_key(k):
    try:
        return as_expr(k)
    except TypeError:
        return k


def _substitute(expr, mapping):
    if expr in mapping:
        return as_expr(mapping[expr])
    if expr.is_Access:
        return expr._subs_indices(mapping)
    if expr.is_Derivative:
        return expr._rebuild_with(_substitute(expr.expr, mapping))
    if not expr.args:
        return expr
    return expr._rebuild(*[_substitute(a, mapping) if isinstance(a, Expr) else a for a in expr.args])


class Number(Expr):
    is_Number = True

    def __init__(self, value):
        if isinstance(value, Number):
            value = value.value
        if isinstance(value, float) and value.is_integer() and abs(value) < 2**53:
            pass
        self.value = value

    def _key(self):
        v = self.value
        return ('N', float(v))

    def __repr__(self):
        return repr(self.value) if not isinstance(self.value, Fraction) else f"{self.value}"

    def __float__(self):
        return float(self.value)

    def __int__(self):
        return int(self.value)

    def __bool__(self):
        return self.value != 0

    def __lt__(self, o): return float(self) < float(o)
    def __le__(self, o): return float(self) <= float(o)
    def __gt__(self, o): return float(self) > float(o)
    def __ge__(self, o): return float(self) >= float(o)


class Symbol(Expr):
    """A scalar runtime symbol (spacings h_x, dt, bounds x_m ...)."""
    is_Symbol = True

    def __init__(self, name, dtype=np.float32, is_const=True):
        self.name = name
        self.dtype = dtype

    def _key(self):
        return ('S', type(self).__name__, self.name)

    def __repr__(self):
        return self.name


def _num(v):
    return v.value if isinstance(v, Number) else None


class Add(Expr):
    def __init__(self, *args):
        self.args = tuple(args)

    @staticmethod
    def make(*terms):
        flat = []
        const = 0
        for t in terms:
            t = as_expr(t)
            if isinstance(t, Add):
                items = t.args
            else:
                items = (t,)
            for i in items:
                if i.is_Number:
                    const = const + i.value
                else:
                    flat.append(i)
        if const != 0 or not flat:
            flat.append(Number(const))
        if len(flat) == 1:
            return flat[0]
        return Add(*flat)

    def __repr__(self):
        return "(" + " + ".join(map(repr, self.args)) + ")"


class Mul(Expr):
    def __init__(self, *args):
        self.args = tuple(args)

    @staticmethod
    def make(*factors):
        flat = []
        const = 1
        for f in factors:
            f = as_expr(f)
            items = f.args if isinstance(f, Mul) else (f,)
            for i in items:
                if i.is_Number:
                    const = const * i.value
                else:
                    flat.append(i)
        if const == 0:
            return Number(0)
        if not flat:
            return Number(const)
        if const != 1:
            flat.insert(0, Number(const))
        if len(flat) == 1:
            return flat[0]
        return Mul(*flat)

    def __repr__(self):
        return "*".join(map(repr, self.args))


class Pow(Expr):
    def __init__(self, base, exponent):
        self.args = (base, exponent)

    @property
    def base(self):
        return self.args[0]

    @property
    def exponent(self):
        return self.args[1]

    @staticmethod
    def make(base, exponent):
        base, exponent = as_expr(base), as_expr(exponent)
        if exponent.is_Number:
            e = exponent.value
            if e == 1:
                return base
            if e == 0:
                return Number(1)
            if base.is_Number:
                b = base.value
                if isinstance(e, int) and isinstance(b, (int, Fraction)):
                    return Number(Fraction(b) ** e if e < 0 else b ** e)
                return Number(float(b) ** float(e))
            if isinstance(base, Pow) and base.exponent.is_Number and isinstance(e, int):
                return Pow.make(base.base, base.exponent.value * e)
        return Pow(base, exponent)

    def __repr__(self):
        return f"{self.base!r}**{self.exponent!r}"


_np_funcs = {'sin': np.sin, 'cos': np.cos, 'sqrt': np.sqrt, 'Abs': np.abs, 'sign': np.sign,
             'exp': np.exp, 'floor': np.floor, 'INT': np.trunc, 'tan': np.tan, 'log': np.log}
_py_funcs = {'sin': math.sin, 'cos': math.cos, 'sqrt': math.sqrt, 'Abs': abs,
             'sign': lambda v: (v > 0) - (v < 0), 'exp': math.exp, 'floor': math.floor,
             'INT': math.trunc, 'tan': math.tan, 'log': math.log}


class Call(Expr):
    """Elementary function application (devito/finite_differences/elementary.py)."""

    def __init__(self, name, arg):
        self.name = name
        self.args = (name, arg)

    @property
    def arg(self):
        return self.args[1]

    def _key(self):
        return ('C', self.name, self.arg._key())

    def _rebuild(self, name, arg):
        return _call(name, arg)

    def __repr__(self):
        return f"{self.name}({self.arg!r})"


def _call(name, arg):
    arg = as_expr(arg)
    if arg.is_Number:
        return as_expr(_py_funcs[name](float(arg.value) if not isinstance(arg.value, int) else arg.value))
    return Call(name, arg)


def sin(x): return _call('sin', x)
def cos(x): return _call('cos', x)
def sqrt(x): return _call('sqrt', x)
def Abs(x): return _call('Abs', x)
def sign(x): return _call('sign', x)
def exp(x): return _call('exp', x)
def floor(x): return _call('floor', x)
def INT(x): return _call('INT', x)


# ---------------------------------------------------------------------------------------------
# function accesses
# ---------------------------------------------------------------------------------------------
class Index:
    """One index of a function access: `base + shift*base.spacing`, or an absolute integer."""
    __slots__ = ('base', 'shift', 'absolute')

    def __init__(self, base, shift=0, absolute=None):
        self.base = base
        self.shift = Fraction(shift)
        self.absolute = absolute

    def _key(self):
        return ('I', self.base.name if self.base is not None else None, self.shift, self.absolute)

    def __repr__(self):
        if self.absolute is not None:
            return str(self.absolute)
        if self.shift == 0:
            return self.base.name
        return f"{self.base.name}{'+' if self.shift > 0 else '-'}{abs(self.shift)}"

    @staticmethod
    def parse(expr, dim):
        """Turn an index expression into an `Index` for axis `dim`."""
        expr = as_expr(expr)
        if expr.is_Number:
            return Index(dim, 0, absolute=int(expr.value))
        if expr.is_Dimension:
            return Index(expr, 0)
        if isinstance(expr, Add):
            base, shift = None, Fraction(0)
            for t in expr.args:
                if t.is_Dimension:
                    if base is not None:
                        raise ValueError(f"unsupported index {expr!r}")
                    base = t
                else:
                    c = _spacing_multiple(t, None)
                    if c is None:
                        raise ValueError(f"unsupported index {expr!r}")
                    shift += c[1]
                    sp = c[0]
                    if base is not None and sp is not None and sp is not base.spacing and sp != base.spacing:
                        raise ValueError(f"index {expr!r} mixes spacings")
            if base is None:
                raise ValueError(f"unsupported index {expr!r}")
            return Index(base, shift)
        raise ValueError(f"unsupported index {expr!r}")


def _spacing_multiple(term, _):
    """term == c * spacing  ->  (spacing_symbol, Fraction c); integers -> (None, c)."""
    if term.is_Number:
        return (None, Fraction(term.value).limit_denominator(1 << 20))
    if term.is_Symbol:
        return (term, Fraction(1))
    if isinstance(term, Mul) and len(term.args) == 2 and term.args[0].is_Number and term.args[1].is_Symbol:
        return (term.args[1], Fraction(term.args[0].value).limit_denominator(1 << 20))
    return None


class Access(Expr):
    """`f[indices]`; a discrete function object is itself the access at its own dimensions."""
    is_Access = True

    def __init__(self, function, indices):
        self.function = function
        self._indices = tuple(indices)

    @property
    def index_objs(self):
        return self._indices

    def _key(self):
        return ('A', self.function.name, id(self.function)) + tuple(i._key() for i in self._indices)

    def __repr__(self):
        return f"{self.function.name}[{', '.join(map(repr, self._indices))}]"

    def map_leaves(self, fn):
        return fn(self)

    # devito-style conveniences used by user code
    @property
    def name(self): return self.function.name
    @property
    def grid(self): return self.function.grid
    @property
    def space_order(self): return self.function.space_order
    @property
    def dimensions(self): return self.function.dimensions
    @property
    def indices(self):
        return tuple((i.base + i.shift * i.base.spacing) if i.absolute is None else Number(i.absolute)
                     for i in self._indices)

    def _shift(self, dim, k):
        new = []
        changed = False
        for i in self._indices:
            if i.absolute is None and (i.base is dim or getattr(i.base, 'root', i.base) is dim):
                new.append(Index(i.base, i.shift + k))
                changed = True
            else:
                new.append(i)
        return Access(self.function, new) if changed else self

    def _subs_indices(self, mapping):
        new = []
        changed = False
        for i in self._indices:
            if i.absolute is None and i.base in mapping:
                tgt = Index.parse(as_expr(mapping[i.base]), i.base)
                if tgt.absolute is None:
                    tgt = Index(tgt.base, tgt.shift + i.shift)
                new.append(tgt)
                changed = True
            else:
                new.append(i)
        return Access(self.function, new) if changed else self

    @property
    def forward(self):
        t = self.function.time_dim
        return self._shift(t, 1)

    @property
    def backward(self):
        t = self.function.time_dim
        return self._shift(t, -1)

    @property
    def dt(self):
        f = self.function
        return Derivative(self, (f.time_dim, 1), fd_order=f.time_order)

    @property
    def dt2(self):
        f = self.function
        return Derivative(self, (f.time_dim, 2), fd_order=f.time_order)


# ---------------------------------------------------------------------------------------------
# finite differences
# ---------------------------------------------------------------------------------------------
def fd_offsets(fd_order, mid, is_time=False, side=0):
    """Stencil offsets, reference rule devito/finite_differences/tools.py:289-302."""
    r = Fraction(fd_order, 2)
    mid = Fraction(mid)
    o_min = math.ceil(mid - r) + side
    o_max = math.floor(mid + r) + side
    if o_max == o_min:
        o_max += 1          # time dims / non-staggered functions (tools.py:298-300)
    return list(range(o_min, o_max + 1))


@lru_cache(maxsize=None)
def _fd_weights_cached(deriv_order, offsets, mid):
    from sympy import finite_diff_weights, Rational
    w = finite_diff_weights(deriv_order, [Rational(o) for o in offsets], Rational(mid.numerator, mid.denominator))
    return tuple(float(c.evalf(_PRECISION)) for c in w[-1][-1])


def fd_weights(deriv_order, offsets, mid=0):
    """Taylor weights for unit spacing, rounded to 9 significant digits like the reference
    (devito/finite_differences/tools.py:231-236, finite_difference.py:185-187)."""
    return _fd_weights_cached(int(deriv_order), tuple(int(o) for o in offsets), Fraction(mid))


class Derivative(Expr):
    """Lazy finite-difference derivative (devito/finite_differences/derivative.py).

    `u.dx(fd_order=4, x0=x + x.spacing/2)` re-parameterises; `.T` transposes (adjoint);
    `.evaluate` expands into a weighted sum of shifted copies of the operand."""
    is_Derivative = True

    def __init__(self, expr, *dims, fd_order=None, x0=None, side=None, transpose=False):
        self.expr = as_expr(expr)
        self.dims = tuple((d, int(o)) for d, o in dims)
        self.fd_order = fd_order
        self.x0 = dict(x0 or {})          # {dim: Fraction shift in units of spacing}
        self.side = side
        self.transpose = transpose
        self.args = (self.expr,)

    def _key(self):
        return ('D', self.expr._key(), tuple((d.name, o) for d, o in self.dims), self.fd_order,
                tuple(sorted((d.name, s) for d, s in self.x0.items())), self.side, self.transpose)

    def __repr__(self):
        return f"Derivative({self.expr!r}, {', '.join(f'({d.name},{o})' for d, o in self.dims)})"

    def _rebuild_with(self, expr):
        return Derivative(expr, *self.dims, fd_order=self.fd_order, x0=self.x0, side=self.side,
                          transpose=self.transpose)

    def _rebuild(self, expr):
        return self._rebuild_with(expr)

    def map_leaves(self, fn):
        return self._rebuild_with(self.expr.map_leaves(fn))

    def _shift(self, dim, k):
        return self._rebuild_with(self.expr._shift(dim, k))

    def __call__(self, x0=None, fd_order=None, side=None, **kwargs):
        new_x0 = dict(self.x0)
        if x0 is not None:
            if not isinstance(x0, dict):
                x0 = {self.dims[0][0]: x0}
            for d, v in x0.items():
                idx = Index.parse(as_expr(v), d)
                if idx.absolute is not None or (idx.base is not d and idx.base != d):
                    raise ValueError(f"unsupported x0={v!r} for dimension {d.name}")
                new_x0[d] = idx.shift
        return Derivative(self.expr, *self.dims, fd_order=fd_order if fd_order is not None else self.fd_order,
                          x0=new_x0, side=side if side is not None else self.side, transpose=self.transpose)

    @property
    def T(self):
        return Derivative(self.expr, *self.dims, fd_order=self.fd_order, x0=self.x0, side=self.side,
                          transpose=not self.transpose)

    @property
    def evaluate(self):
        expr = self.expr.evaluate
        for d, order in self.dims:
            expr = self._expand_one(expr, d, order)
        return expr

    def _expand_one(self, expr, dim, deriv_order):
        fd_order = self.fd_order
        if fd_order is None:
            fd_order = expr._default_fd_order if not dim.is_Time else 2
        if isinstance(fd_order, dict):
            fd_order = fd_order[dim]
        # first derivative with 2nd-order FD -> 1st order (finite_difference.py:137-140)
        if deriv_order == 1 and fd_order == 2 and self.side is None:
            fd_order = 1
        mid = self.x0.get(dim, Fraction(0))
        if deriv_order == 0 and mid == 0:
            return expr
        offsets = fd_offsets(fd_order, mid, is_time=dim.is_Time, side=self.side or 0)
        weights = fd_weights(deriv_order, offsets, mid)
        if self.transpose:
            # reference: weights reversed, index set reversed AND mirrored about the reference
            # point (finite_difference.py:196-198, tools.py:180-193) == weight w_k now multiplies
            # the sample at -o_k
            offsets = [-o for o in offsets]
        scale = Pow.make(dim.spacing, -deriv_order)
        terms = []
        for o, w in zip(offsets, weights):
            if w == 0.0:
                continue
            terms.append(Mul.make(Number(w), scale, expr._shift(dim, o)))
        return Add.make(*terms) if terms else Number(0)


# ---------------------------------------------------------------------------------------------
# queries
# ---------------------------------------------------------------------------------------------
def retrieve_functions(expr):
    exprs = expr if isinstance(expr, (list, tuple, set)) else [expr]
    out = []
    for e in exprs:
        for n in as_expr(e).preorder():
            if n.is_Access and n not in out:
                out.append(n)
    return out


def retrieve_derivatives(expr):
    return [n for n in as_expr(expr).preorder() if n.is_Derivative]


class NonLinear(Exception):
    pass


def linear_terms(expr, is_unknown):
    """Decompose `expr = sum_k coef_k * access_k + rest` where `access_k` are the accesses for
    which `is_unknown(access)` holds. Returns ({access: coef_expr}, rest). Raises NonLinear
    if an unknown appears non-linearly."""
    expr = as_expr(expr)

    def has_unknown(e):
        return any(n.is_Access and is_unknown(n) for n in e.preorder())

    def rec(e):
        if e.is_Access and is_unknown(e):
            return {e: Number(1)}, Number(0)
        if not has_unknown(e):
            return {}, e
        if isinstance(e, Add):
            terms, rest = {}, Number(0)
            for a in e.args:
                t, r = rec(a)
                for k, v in t.items():
                    terms[k] = terms[k] + v if k in terms else v
                rest = rest + r
            return terms, rest
        if isinstance(e, Mul):
            with_u = [a for a in e.args if has_unknown(a)]
            if len(with_u) != 1:
                raise NonLinear(repr(e))
            others = [a for a in e.args if a is not with_u[0]]
            t, r = rec(with_u[0])
            scale = Mul.make(*others) if others else Number(1)
            return {k: scale * v for k, v in t.items()}, scale * r
        if e.is_Derivative:
            return rec(e.evaluate)
        raise NonLinear(repr(e))

    return rec(expr)


def div(expr, shift=None, order=None, method='FD', **kwargs):
    """Sum of first derivatives along the space dimensions (devito/finite_differences/operators.py)."""
    expr = as_expr(expr)
    out = Number(0)
    for d in expr._space_dims:
        out = out + Derivative(expr, (d, 1), fd_order=order)
    return out


def grad(expr, shift=None, order=None, method='FD', **kwargs):
    """Tuple of first derivatives along the space dimensions."""
    expr = as_expr(expr)
    return tuple(Derivative(expr, (d, 1), fd_order=order) for d in expr._space_dims)
