"""Isotropic acoustic modelling (mirror of examples/seismic/acoustic/operators.py:50-187 and
wavesolver.py:11-156): forward and adjoint operators (gradient/Born are SURVEY §8f)."""
from .. import Eq, FreeSurface, Inc, Function, Operator, TimeFunction, solve
from ..tools import memoized_meth

__all__ = ['laplacian', 'iso_stencil', 'ForwardOperator', 'AdjointOperator', 'GradientOperator',
           'BornOperator', 'AcousticWaveSolver']


def laplacian(field, model, kernel):
    """Spatial operator (operators.py:50-68): for the 4th-order-in-time scheme the 4th time derivative
    is traded for a double Laplacian, H = laplace + s**2/12 * laplace(1/m * laplace)."""
    if kernel not in ('OT2', 'OT4'):
        raise ValueError("Unrecognized kernel")
    s = model.grid.time_dim.spacing
    biharmonic = field.biharmonic(1 / model.m) if kernel == 'OT4' else 0
    return field.laplace + s ** 2 / 12 * biharmonic


def iso_stencil(field, model, kernel='OT2', **kwargs):
    """u.dt2 * m - H(u) + damp * u.dt = 0 solved for u.forward (operators.py:71-107)."""
    forward = kwargs.get('forward', True)
    unext = field.forward if forward else field.backward
    udt = field.dt if forward else field.dt.T
    lap = laplacian(field, model, kernel)
    eq_time = solve(model.m * field.dt2 - lap - kwargs.get('q', 0) + model.damp * udt, unext)
    update = Eq(unext, eq_time, subdomain=model.grid.subdomains['physdomain'])
    if model.fs:
        # operators.py:105-106 `freesurface(model, Eq(unext, eq_time))`: the top rows get the same
        # update with vertical taps mirrored about the surface, and the surface row is cleared
        return [update, FreeSurface(update, model.grid.subdomains['fsdomain'])]
    return [update]


def ForwardOperator(model, geometry, space_order=4, save=False, kernel='OT2', **kwargs):
    """operators.py:113-150"""
    m = model.m
    u = TimeFunction(name='u', grid=model.grid, save=geometry.nt if save else None, time_order=2,
                     space_order=space_order)
    src, rec = geometry.src, geometry.rec
    s = model.grid.stepping_dim.spacing
    eqn = iso_stencil(u, model, kernel)
    src_term = src.inject(field=u.forward, expr=src * s ** 2 / m)
    rec_term = rec.interpolate(expr=u)
    return Operator(eqn + src_term + rec_term, subs=model.spacing_map, name='Forward', **kwargs)


def AdjointOperator(model, geometry, space_order=4, kernel='OT2', save=None, **kwargs):
    """operators.py:153-187: the update solved for v.backward, receivers injected, adjoint source
    sampled at the source position."""
    m = model.m
    v = TimeFunction(name='v', grid=model.grid, save=None, time_order=2, space_order=space_order)
    srca = geometry.new_src(name='srca', src_type=None)
    rec = geometry.rec
    s = model.grid.stepping_dim.spacing
    eqn = iso_stencil(v, model, kernel, forward=False)
    receivers = rec.inject(field=v.backward, expr=rec * s ** 2 / m)
    source_a = srca.interpolate(expr=v)
    return Operator(eqn + receivers + source_a, subs=model.spacing_map, name='Adjoint', **kwargs)


def GradientOperator(model, geometry, space_order=4, save=True, kernel='OT2', **kwargs):
    """operators.py:190-232: adjoint propagation of the data + imaging condition grad -= u * v.dt2."""
    if kernel != 'OT2':
        raise NotImplementedError("the OT4 imaging condition (extra biharmonic term, operators.py:225-226) "
                                  "is not on this backend's path yet")
    m = model.m
    grad = Function(name='grad', grid=model.grid)
    u = TimeFunction(name='u', grid=model.grid, save=geometry.nt if save else None, time_order=2,
                     space_order=space_order)
    v = TimeFunction(name='v', grid=model.grid, save=None, time_order=2, space_order=space_order)
    rec = geometry.rec
    s = model.grid.stepping_dim.spacing
    eqn = iso_stencil(v, model, kernel, forward=False)
    gradient_update = Inc(grad, -u * v.dt2)
    receivers = rec.inject(field=v.backward, expr=rec * s ** 2 / m)
    return Operator(eqn + receivers + [gradient_update], subs=model.spacing_map, name='Gradient', **kwargs)


def BornOperator(model, geometry, space_order=4, kernel='OT2', **kwargs):
    """operators.py:235-277: background field u driven by the source, linearised field U driven by
    `-dm * u.dt2`, receivers sampling U."""
    kwargs.pop('save', None)
    m = model.m
    src, rec = geometry.src, geometry.rec
    u = TimeFunction(name='u', grid=model.grid, save=None, time_order=2, space_order=space_order)
    U = TimeFunction(name='U', grid=model.grid, save=None, time_order=2, space_order=space_order)
    dm = Function(name='dm', grid=model.grid, space_order=0)
    s = model.grid.stepping_dim.spacing
    eqn1 = iso_stencil(u, model, kernel)
    eqn2 = iso_stencil(U, model, kernel, q=-dm * u.dt2)
    source = src.inject(field=u.forward, expr=src * s ** 2 / m)
    receivers = rec.interpolate(expr=U)
    return Operator(eqn1 + source + eqn2 + receivers, subs=model.spacing_map, name='Born', **kwargs)


class AcousticWaveSolver:
    """wavesolver.py:11-120 (`forward` only)."""

    def __init__(self, model, geometry, kernel='OT2', space_order=4, **kwargs):
        self.model = model
        self.model._initialize_bcs(bcs="damp")
        self.geometry = geometry
        self.space_order = space_order
        self.kernel = kernel
        self._kwargs = kwargs

    @property
    def dt(self):
        # the 4th-order scheme is stable with a sqrt(3) = 1.73 larger step (wavesolver.py:39-44)
        if self.kernel == 'OT4':
            return self.model.dtype(1.73 * self.model.critical_dt)
        return self.model.critical_dt

    @memoized_meth
    def op_fwd(self, save=None):
        return ForwardOperator(self.model, save=save, geometry=self.geometry, kernel=self.kernel,
                               space_order=self.space_order, **self._kwargs)

    @memoized_meth
    def op_adj(self):
        return AdjointOperator(self.model, save=None, geometry=self.geometry, kernel=self.kernel,
                               space_order=self.space_order, **self._kwargs)

    def adjoint(self, rec, srca=None, v=None, model=None, **kwargs):
        """wavesolver.py:118-156"""
        srca = srca or self.geometry.new_src(name='srca', src_type=None)
        v = v or TimeFunction(name='v', grid=self.model.grid, time_order=2, space_order=self.space_order)
        model = model or self.model
        kwargs.update(model.physical_params(**kwargs))
        summary = self.op_adj().apply(srca=srca, rec=rec, v=v, dt=kwargs.pop('dt', self.dt), **kwargs)
        return srca, v, summary

    @memoized_meth
    def op_grad(self, save=True):
        return GradientOperator(self.model, save=save, geometry=self.geometry, kernel=self.kernel,
                                space_order=self.space_order, **self._kwargs)

    def jacobian_adjoint(self, rec, u, src=None, v=None, grad=None, model=None, **kwargs):
        """wavesolver.py:158-230 (without checkpointing): gradient of the data misfit w.r.t. m."""
        dt = kwargs.pop('dt', self.dt)
        grad = grad or Function(name='grad', grid=self.model.grid)
        v = v or TimeFunction(name='v', grid=self.model.grid, time_order=2, space_order=self.space_order)
        model = model or self.model
        kwargs.update(model.physical_params(**kwargs))
        summary = self.op_grad().apply(rec=rec, grad=grad, v=v, u=u, dt=dt, **kwargs)
        return grad, summary

    jacobian_adjoint.__name__ = 'jacobian_adjoint'
    gradient = jacobian_adjoint

    @memoized_meth
    def op_born(self):
        return BornOperator(self.model, save=None, geometry=self.geometry, kernel=self.kernel,
                            space_order=self.space_order, **self._kwargs)

    def jacobian(self, dmin, src=None, rec=None, u=None, U=None, model=None, **kwargs):
        """wavesolver.py:216-254: linearised (Born) modelling for the model perturbation `dmin`
        (a Function, or an array of the grid's shape)."""
        src = src or self.geometry.src
        rec = rec or self.geometry.rec
        mk = lambda n: TimeFunction(name=n, grid=self.model.grid, time_order=2, space_order=self.space_order)
        u = u or mk('u')
        U = U or mk('U')
        model = model or self.model
        kwargs.update(model.physical_params(**kwargs))
        summary = self.op_born().apply(dm=dmin, u=u, U=U, src=src, rec=rec, dt=kwargs.pop('dt', self.dt),
                                       **kwargs)
        return rec, u, U, summary

    born = jacobian

    def forward(self, src=None, rec=None, u=None, model=None, save=None, **kwargs):
        src = src or self.geometry.src
        rec = rec or self.geometry.rec
        u = u or TimeFunction(name='u', grid=self.model.grid, save=self.geometry.nt if save else None,
                              time_order=2, space_order=self.space_order)
        model = model or self.model
        kwargs.update(model.physical_params(**kwargs))
        summary = self.op_fwd(save).apply(src=src, rec=rec, u=u, dt=kwargs.pop('dt', self.dt), **kwargs)
        return rec, u, summary
