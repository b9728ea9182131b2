"""Slab domain decomposition along x, one process per GPU.

Replaces the reference's MPI `Distributor` (devito/mpi/distributed.py:316-485) for the one
topology this backend shards on: `topology=('*', 1, 1)` / `DEVITO_TOPOLOGY=x`
(devito/mpi/distributed.py:956-961).  Splitting follows `np.array_split` like the reference
(:379-382).  x is the slowest-varying axis of the (t, x, y, z) layout, so every exchanged
face is one contiguous block and travels by NCCL send/recv without packing (b2_halo.cu).
Process plumbing is `torch.distributed` (rank/world from the launcher env); the data path
uses the library's own NCCL communicator.
"""
import os

import numpy as np

from .parameters import configuration

__all__ = ['Distributor', 'init_distributed', 'world']


class _World:
    def __init__(self):
        self.rank = 0
        self.size = 1
        self.initialized = False
        self.halo_ctx = None       # opaque b2_halo_ctx*
        self.device = None


world = _World()


def init_distributed(backend=None):
    """Initialise one-process-per-GPU execution from the torchrun environment
    (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT)."""
    import torch
    import torch.distributed as dist
    if world.initialized:
        return world
    ws = int(os.environ.get('WORLD_SIZE', '1'))
    if ws <= 1:
        world.initialized = True
        return world
    if not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if torch.cuda.is_available():
            torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
        dist.init_process_group(backend=backend)
    world.rank = dist.get_rank()
    world.size = dist.get_world_size()
    world.initialized = True
    configuration['mpi'] = True
    if torch.cuda.is_available():
        # one process per GPU: live on the socket the GPU hangs off (what `mpirun --bind-to` does for the reference)
        from .numa import bind_to_gpu
        world.numa = bind_to_gpu(int(os.environ.get('LOCAL_RANK', '0')))
    return world


def finalize_distributed():
    import torch.distributed as dist
    if world.halo_ctx is not None:
        from ._lib import lib
        for m in _p2p['mapped']:
            lib().b2_ipc_close(m)
        _p2p.update(ready=False, disabled=False, flags=None, mapped=[])
        lib().b2_halo_destroy(world.halo_ctx)
        world.halo_ctx = None
    if dist.is_initialized():
        dist.destroy_process_group()
    world.rank, world.size, world.initialized = 0, 1, False
    configuration['mpi'] = False


def halo_context(deviceid):
    """Create (once) the NCCL communicator the C library uses for halo exchange."""
    if world.size <= 1:
        return None
    if world.halo_ctx is not None:
        return world.halo_ctx
    import ctypes
    import torch
    import torch.distributed as dist
    from ._lib import lib, nccl_library_path
    L = lib()
    path = nccl_library_path()
    buf = ctypes.create_string_buffer(128)
    if world.rank == 0:
        rc = L.b2_nccl_unique_id(path, buf)
        if rc:
            raise RuntimeError(f"b2_nccl_unique_id failed: {L.b2_last_error().decode()}")
    payload = [bytes(buf.raw)]
    dist.broadcast_object_list(payload, src=0)
    ctx = L.b2_halo_create(path, payload[0], world.rank, world.size, deviceid)
    if not ctx:
        raise RuntimeError(f"b2_halo_create failed: {L.b2_last_error().decode()}")
    world.halo_ctx = ctx
    return ctx


# ---------------------------------------------------------------------------------------------
# peer-memory halo path: CUDA IPC handle exchange (host plumbing only; the data path is in
# csrc/b2_halo.cu)
# ---------------------------------------------------------------------------------------------
_p2p = {'ready': False, 'disabled': False, 'flags': None, 'mapped': []}


def p2p_enabled():
    return os.environ.get('B2_HALO', 'p2p').lower() != 'nccl' and not _p2p['disabled']


def _gather(obj):
    import torch.distributed as dist
    out = [None] * world.size
    dist.all_gather_object(out, obj)
    return out


def _p2p_setup(deviceid):
    """Collective: every rank allocates its flag buffer and maps its neighbours'."""
    import ctypes
    from ._lib import lib
    if _p2p['ready'] or not p2p_enabled():
        return _p2p['ready']
    L = lib()
    ctx = halo_context(deviceid)
    flags = L.b2_malloc_device(64, deviceid)
    L.b2_memset_device(flags, 0, 64, deviceid)
    L.b2_synchronize(deviceid)
    buf = ctypes.create_string_buffer(64)
    ok = L.b2_ipc_get_handle(flags, buf) == 0
    handles = _gather(bytes(buf.raw) if ok else None)
    if any(h is None for h in handles):
        _p2p['disabled'] = True
        return False
    left = world.rank - 1 if world.rank > 0 else None
    right = world.rank + 1 if world.rank < world.size - 1 else None
    lp = L.b2_ipc_open(handles[left]) if left is not None else None
    rp = L.b2_ipc_open(handles[right]) if right is not None else None
    good = (left is None or lp) and (right is None or rp)
    if not all(_gather(bool(good))):
        _p2p['disabled'] = True
        return False
    # we signal slot [1] ("from right") of the left neighbour and slot [0] of the right neighbour
    fl = ctypes.c_void_p(lp + 4) if lp else None
    fr = ctypes.c_void_p(rp) if rp else None
    if L.b2_halo_p2p_setup(ctx, flags, fl, fr):
        raise RuntimeError(L.b2_last_error().decode())
    _p2p.update(ready=True, flags=flags)
    _p2p['mapped'] += [m for m in (lp, rp) if m]
    return True


def register_field(storage, grid, deviceid):
    """Collective: export this rank's device allocation of a wavefield to its x-neighbours and map
    theirs, so that boundary planes can be stored straight into the neighbour's halo."""
    import ctypes
    from ._lib import lib
    if storage.p2p_registered or not _p2p_setup(deviceid):
        return storage.p2p_registered
    L = lib()
    ctx = halo_context(deviceid)
    buf = ctypes.create_string_buffer(64)
    ok = L.b2_ipc_get_handle(storage.dev.data_ptr(), buf) == 0
    info = _gather((bytes(buf.raw) if ok else None, int(grid.shape[0])))
    if any(h is None for h, _ in info):
        return False
    left = world.rank - 1 if world.rank > 0 else None
    right = world.rank + 1 if world.rank < world.size - 1 else None
    lp = L.b2_ipc_open(info[left][0]) if left is not None else None
    rp = L.b2_ipc_open(info[right][0]) if right is not None else None
    good = (left is None or lp) and (right is None or rp)
    if not all(_gather(bool(good))):
        return False
    nl = info[left][1] if left is not None else 0
    nr = info[right][1] if right is not None else 0
    if L.b2_halo_p2p_register(ctx, storage.dev.data_ptr(), lp, rp, nl, nr):
        raise RuntimeError(L.b2_last_error().decode())
    _p2p['mapped'] += [m for m in (lp, rp) if m]
    storage.p2p_registered = True
    return True


class Distributor:
    """Per-Grid decomposition info (local shape, global offsets, neighbours)."""

    def __init__(self, glb_shape, dimensions, topology=None):
        self.glb_shape = tuple(glb_shape)
        self.dimensions = tuple(dimensions)
        nprocs = world.size if (configuration['mpi'] and world.initialized) else 1
        if topology is not None:
            topo = tuple(nprocs if t == '*' else int(t) for t in topology)
            if int(np.prod(topo)) != nprocs or any(t != 1 for t in topo[1:]):
                raise ValueError(f"only x-slab topologies ('*', 1, ...) are supported, got {topology}")
        self.nprocs = nprocs
        self.myrank = world.rank if nprocs > 1 else 0
        self.topology = (nprocs,) + (1,) * (len(glb_shape) - 1)
        # np.array_split semantics (devito/mpi/distributed.py:379-382)
        parts = np.array_split(np.arange(self.glb_shape[0]), nprocs)
        self._x_ranges = [(int(p[0]), int(p[-1]) + 1) if len(p) else (0, 0) for p in parts]
        lo, hi = self._x_ranges[self.myrank]
        self.shape = (hi - lo,) + self.glb_shape[1:]
        self.offsets = (lo,) + (0,) * (len(glb_shape) - 1)

    @property
    def is_parallel(self):
        return self.nprocs > 1

    @property
    def x_range(self):
        return self._x_ranges[self.myrank]

    def x_range_of(self, rank):
        return self._x_ranges[rank]

    @property
    def neighbors(self):
        left = self.myrank - 1 if self.myrank > 0 else None
        right = self.myrank + 1 if self.myrank < self.nprocs - 1 else None
        return left, right

    @property
    def is_boundary_left(self):
        return self.myrank == 0

    @property
    def is_boundary_right(self):
        return self.myrank == self.nprocs - 1
