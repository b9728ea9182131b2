"""Built-in reductions and initialisers.

The reference implements these as mini-Operators (devito/builtins/arithmetic.py:11-220,
devito/builtins/initializers.py:12-330); they are set-up code, not the hot path, so here
they are NumPy one-liners on the host copy (with a `torch.distributed` all-reduce under
slab decomposition, replacing the reference's `MPI_Allreduce`, devito/mpi/routines.py:1415-1429).
"""
import numpy as np

from .types import Function, Constant
from . import distributed

__all__ = ['norm', 'sumall', 'sum', 'inner', 'mmin', 'mmax', 'assign', 'smooth', 'gaussian_smooth',
           'initialize_function']


def _domain(f):
    return f.data_ro_domain


def _allreduce(val, op='sum'):
    w = distributed.world
    if not (w.initialized and w.size > 1):
        return val
    import torch
    import torch.distributed as dist
    t = torch.tensor([val], dtype=torch.float64)
    if dist.get_backend() == 'nccl':
        t = t.cuda()
    dist.all_reduce(t, op={'sum': dist.ReduceOp.SUM, 'max': dist.ReduceOp.MAX,
                           'min': dist.ReduceOp.MIN}[op])
    return float(t.item())


def _is_sparse(f):
    return getattr(f, 'is_SparseFunction', False)


def norm(f, order=2):
    """(sum |f|^order)^(1/order) over the domain, accumulated in float64, cast back to f.dtype
    (devito/builtins/arithmetic.py:11-41)."""
    d = np.abs(np.asarray(_domain(f), dtype=np.float64)) ** order
    s = float(d.sum())
    if not _is_sparse(f):
        s = _allreduce(s)
    return f.dtype(s ** (1.0 / order))


def sumall(f):
    s = float(np.asarray(_domain(f), dtype=np.float64).sum())
    if not _is_sparse(f):
        s = _allreduce(s)
    return f.dtype(s)


sum = sumall


def inner(f, g):
    s = float((np.asarray(_domain(f), dtype=np.float64) * np.asarray(_domain(g), dtype=np.float64)).sum())
    if not _is_sparse(f):
        s = _allreduce(s)
    return f.dtype(s)


def mmin(f):
    if isinstance(f, Constant):
        return f.data
    if np.isscalar(f):
        return f
    return f.dtype(_allreduce(float(np.min(_domain(f))), 'min'))


def mmax(f):
    if isinstance(f, Constant):
        return f.data
    if np.isscalar(f):
        return f
    return f.dtype(_allreduce(float(np.max(_domain(f))), 'max'))


def assign(f, rhs=0, options=None, name='assign', **kwargs):
    if isinstance(rhs, Function):
        f.data[:] = rhs.data_ro_domain
    else:
        f.data[:] = rhs


def smooth(f, g, axis=None):
    """Simple 3-point running average of g into f along `axis` dims (builtins/initializers.py:50)."""
    src = np.asarray(g.data_ro_domain if isinstance(g, Function) else g, dtype=np.float64)
    out = src.copy()
    dims = range(src.ndim) if axis is None else [f.dimensions.index(a) for a in (axis if isinstance(axis, (list, tuple)) else [axis])]
    for ax in dims:
        pad = np.pad(out, [(1, 1) if i == ax else (0, 0) for i in range(out.ndim)], mode='edge')
        sl = lambda o: tuple(slice(o, o + out.shape[i]) if i == ax else slice(None) for i in range(out.ndim))
        out = (pad[sl(0)] + pad[sl(1)] + pad[sl(2)]) / 3.0
    f.data[:] = out.astype(f.dtype)


def gaussian_smooth(f, sigma=1, truncate=4.0, mode='reflect'):
    """Gaussian smoothing, same contract as scipy.ndimage.gaussian_filter
    (devito/builtins/initializers.py:88-217)."""
    from scipy.ndimage import gaussian_filter
    if isinstance(f, Function):
        f.data[:] = gaussian_filter(np.asarray(f.data_ro_domain), sigma=sigma, truncate=truncate, mode=mode)
        return f
    return gaussian_filter(f, sigma=sigma, truncate=truncate, mode=mode)


def initialize_function(function, data, nbl, mapper=None, mode='constant', name=None, **kwargs):
    """Write `data` into the interior of `function` and fill the `nbl` layers plus the outer halo
    by edge replication (`mode='constant'`) or reflection
    (devito/builtins/initializers.py:220-330, builtins/utils.py:61-115)."""
    nd = function.ndim
    if isinstance(nbl, int):
        nbl = tuple((nbl, nbl) for _ in range(nd))
    nbl = tuple(tuple(int(v) for v in p) for p in nbl)
    data = np.asarray(data)
    dist = function.grid.distributor if function.grid is not None else None
    if dist is not None and dist.is_parallel:
        # `data` is the global physical array: pad globally, then cut this rank's x-slab (+halo)
        full = np.pad(data, nbl, mode='edge' if mode == 'constant' else mode)
        so = function.space_order
        full = np.pad(full, so, mode='edge')
        lo, hi = dist.x_range
        function.data_with_halo[...] = full[lo:hi + 2 * so].astype(function.dtype)
        return
    padded = np.pad(data, nbl, mode='edge' if mode == 'constant' else mode)
    if padded.shape != function.shape:
        raise ValueError(f"initialize_function: padded data shape {padded.shape} != {function.shape}")
    so = function.space_order
    full = np.pad(padded, so, mode='edge') if so else padded
    function.data_with_halo[...] = full.astype(function.dtype)
