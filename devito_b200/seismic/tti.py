"""TTI (centred kernel) forward modelling — mirror of examples/seismic/tti/operators.py:12-247,
431-480 and tti/wavesolver.py (forward only)."""
from .. import Eq, Operator, TimeFunction, solve, cos, sin, sqrt
from ..tools import memoized_meth

__all__ = ['kernel_centered', 'ForwardOperator', 'AnisotropicWaveSolver']


def trig_func(model):
    theta = getattr(model, 'theta', 0)
    phi = getattr(model, 'phi', 0)
    return cos(theta), sin(theta), cos(phi), sin(phi)


def Gzz_centered(model, field):
    """Rotated second derivative along the symmetry axis (operators.py:65-104)."""
    costheta, sintheta, cosphi, sinphi = trig_func(model)
    order1 = field.space_order // 2
    x, y, z = field.grid.dimensions
    dx, dy, dz = x.spacing / 2, y.spacing / 2, z.spacing / 2
    Gz = (sintheta * cosphi * field.dx(fd_order=order1, x0=x + dx) +
          sintheta * sinphi * field.dy(fd_order=order1, x0=y + dy) +
          costheta * field.dz(fd_order=order1, x0=z + dz))
    Gzz = (Gz * costheta).dz(fd_order=order1, x0=z - dz)
    if sintheta != 0:
        Gzz += (Gz * sintheta * cosphi).dx(fd_order=order1, x0=x - dx)
    if sinphi != 0:
        Gzz += (Gz * sintheta * sinphi).dy(fd_order=order1, x0=y - dy)
    return Gzz


def kernel_centered(model, u, v):
    """operators.py:186-247 + :12-39"""
    delta, epsilon = sqrt(1 + 2 * model.delta), 1 + 2 * model.epsilon
    Gxx = u.laplace - Gzz_centered(model, u)
    Gzz = Gzz_centered(model, v)
    H0 = epsilon * Gxx + delta * Gzz
    Hz = delta * Gxx + Gzz
    m, damp = model.m, model.damp
    stencilp = solve(m * u.dt2 - H0 + damp * u.dt, u.forward)
    stencilr = solve(m * v.dt2 - Hz + damp * v.dt, v.forward)
    sd = model.grid.subdomains['physdomain']
    return [Eq(u.forward, stencilp, subdomain=sd), Eq(v.forward, stencilr, subdomain=sd)]


def ForwardOperator(model, geometry, space_order=4, save=False, kernel='centered', **kwargs):
    """operators.py:431-480"""
    if kernel != 'centered' or model.dim != 3:
        raise NotImplementedError("only the 3-D centred TTI kernel is on this backend's path")
    dt = model.grid.time_dim.spacing
    m = model.m
    u = TimeFunction(name='u', grid=model.grid, save=geometry.nt if save else None, time_order=2,
                     space_order=space_order)
    v = TimeFunction(name='v', grid=model.grid, save=geometry.nt if save else None, time_order=2,
                     space_order=space_order)
    src, rec = geometry.src, geometry.rec
    stencils = kernel_centered(model, u, v)
    stencils += src.inject(field=(u.forward, v.forward), expr=src * dt ** 2 / m)
    stencils += rec.interpolate(expr=u + v)
    return Operator(stencils, subs=model.spacing_map, name='ForwardTTI', **kwargs)


class AnisotropicWaveSolver:
    def __init__(self, model, geometry, space_order=4, kernel='centered', **kwargs):
        self.model = model
        self.model._initialize_bcs(bcs="damp")
        self.geometry = geometry
        self.kernel = kernel
        if space_order % 2 != 0:
            raise ValueError("space_order must be even")
        self.space_order = space_order
        self._kwargs = kwargs

    @property
    def dt(self):
        return self.model.critical_dt

    @memoized_meth
    def op_fwd(self, save=False):
        return ForwardOperator(self.model, save=save, geometry=self.geometry,
                               space_order=self.space_order, kernel=self.kernel, **self._kwargs)

    def forward(self, src=None, rec=None, u=None, v=None, model=None, save=False, **kwargs):
        src = src or self.geometry.src
        rec = rec or self.geometry.rec
        mk = lambda n: TimeFunction(name=n, grid=self.model.grid, save=self.geometry.nt if save else None,
                                    time_order=2, space_order=self.space_order)
        u = u or mk('u')
        v = v or mk('v')
        model = model or self.model
        kwargs.update(model.physical_params(**kwargs))
        summary = self.op_fwd(save).apply(src=src, rec=rec, u=u, v=v, dt=kwargs.pop('dt', self.dt), **kwargs)
        return rec, u, v, summary
