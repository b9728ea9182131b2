"""Equations and the linear solve used to derive explicit time updates.

`Eq`/`Inc` mirror devito/types/equation.py; `solve` mirrors devito/operations/solve.py:19-78:
the expression must be linear in the target, and the result is `-rest / coefficient`."""
from .symbolics import as_expr, linear_terms, NonLinear

__all__ = ['Eq', 'Inc', 'FreeSurface', 'solve']


class Eq:
    is_Increment = False

    def __init__(self, lhs, rhs=0, subdomain=None, coefficients=None, implicit_dims=None, **kwargs):
        self.lhs = as_expr(lhs)
        self.rhs = as_expr(rhs)
        self.subdomain = subdomain
        self.implicit_dims = implicit_dims

    @property
    def args(self):
        return (self.lhs, self.rhs)

    def func(self, lhs, rhs, subdomain=None, **kwargs):
        return type(self)(lhs, rhs, subdomain=subdomain if subdomain is not None else self.subdomain,
                          implicit_dims=self.implicit_dims)

    @property
    def evaluate(self):
        return self.func(self.lhs, self.rhs.evaluate)

    def subs(self, mapping):
        return self.func(self.lhs.subs(mapping), self.rhs.subs(mapping))

    xreplace = subs

    def __repr__(self):
        return f"{type(self).__name__}({self.lhs!r}, {self.rhs!r})"

    # list-like concatenation convenience (`[eq] + src_term`)
    def __add__(self, other):
        return [self] + list(other)

    def __radd__(self, other):
        return list(other) + [self]


class Inc(Eq):
    is_Increment = True


class FreeSurface:
    """Free-surface boundary condition for the explicit update `eq` on the low end of the grid's last
    dimension — what the reference's `freesurface(model, eq)` builds symbolically
    (examples/seismic/acoustic/operators.py:5-47): on `subdomain` (the top `space_order` rows) the same
    update with every vertical tap that falls above the surface replaced by the antisymmetric mirror
    `sign(z - k) * u[|z - k|]`, followed by `u.forward[z = 0] = 0`.

    It is a first-class object here (the reference rewrites sub-expressions with `sign`/`INT(abs())`
    indices); the CUDA path implements it natively (`b2_iso_args.free_surface`)."""
    is_Increment = False

    def __init__(self, eq, subdomain):
        if not isinstance(eq, Eq):
            raise TypeError("FreeSurface wraps the time-update Eq it mirrors")
        self.eq = eq
        self.subdomain = subdomain

    @property
    def field(self):
        return self.eq.lhs.function

    def __repr__(self):
        return f"FreeSurface({self.eq.lhs!r})"


def solve(eq, target, **kwargs):
    """Algebraically solve `eq == 0` (or an `Eq`) for `target`; `eq` must be linear in it."""
    if isinstance(eq, Eq):
        eq = eq.lhs - eq.rhs
    expr = as_expr(eq).evaluate
    target = as_expr(target)
    try:
        terms, rest = linear_terms(expr, lambda a: a == target)
    except NonLinear as e:
        raise ValueError(f"solve: expression is not linear in {target!r}: {e}") from None
    if not terms:
        raise ValueError(f"solve: {target!r} does not appear in the expression")
    coef = list(terms.values())[0]
    return (-1 * rest) / coef
