"""Worker for the world_size-2 tests (launched by torch.distributed.run). Mode 'host' needs no
GPU (gloo): decomposition, parameter slabs, reductions. Mode 'gpu' runs the decomposed
propagator with NCCL halo exchange and compares with the single-GPU result."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import devito_b200 as dv  # noqa: E402
from devito_b200.seismic import (SeismicModel, demo_model, setup_geometry, AcousticWaveSolver,  # noqa: E402
                                 AnisotropicWaveSolver, damp_profile)


def host_mode():
    import torch.distributed as dist
    w = dv.init_distributed(backend='gloo')
    assert w.size == 2
    nbl, so, n = 6, 8, (21, 12, 12)
    vp = np.linspace(1.5, 3.0, int(np.prod(n)), dtype=np.float32).reshape(n)
    model = SeismicModel(origin=(0., 0., 0.), spacing=(10., 10., 10.), shape=n, space_order=so, vp=vp,
                         nbl=nbl, bcs="damp", topology=('*', 1, 1))
    d = model.grid.distributor
    N = n[0] + 2 * nbl
    parts = np.array_split(np.arange(N), 2)                 # devito/mpi/distributed.py:379-382
    assert d.x_range == (int(parts[w.rank][0]), int(parts[w.rank][-1]) + 1)
    assert model.grid.shape == (len(parts[w.rank]), n[1] + 2 * nbl, n[2] + 2 * nbl)
    assert model.grid.shape_global == (N, n[1] + 2 * nbl, n[2] + 2 * nbl)
    # damping profile and padded parameter follow GLOBAL indices
    full = damp_profile(model.grid.shape_global, [(nbl, nbl)] * 3, model.grid.spacing)
    lo, hi = d.x_range
    assert np.allclose(np.asarray(model.damp.data), full[lo:hi], rtol=1e-6, atol=1e-9)
    # ... and equals the serial DSL-built profile (initialize_damp through the interpreter)
    from devito_b200 import configuration
    serial = damp_profile(model.grid.shape_global, [(nbl, nbl)] * 3, model.grid.spacing, x_range=(lo, hi))
    assert np.array_equal(np.asarray(model.damp.data), serial)
    vp_full = np.pad(np.pad(vp, nbl, mode='edge'), so, mode='edge')
    assert np.array_equal(np.asarray(model.vp.data_with_halo), vp_full[lo:hi + 2 * so])
    # reductions agree with the serial values on every rank
    assert np.isclose(dv.mmax(model.vp), vp.max())
    serial_norm = np.sqrt(np.sum(full.astype(np.float64) ** 2))
    assert np.isclose(float(dv.norm(model.damp)), serial_norm, rtol=1e-6)
    assert model.critical_dt > 0
    # sparse points are expressed relative to the local slab
    geometry = setup_geometry(model, 30.0)
    op = AcousticWaveSolver(model, geometry, space_order=so).op_fwd()
    assert op.backend == 'cuda-sm100a'
    # free-surface model under decomposition: no layer above z, slab-local profile, operator recognised
    fsm = SeismicModel(origin=(0., 0., 0.), spacing=(10., 10., 10.), shape=n, space_order=so, vp=vp,
                       nbl=nbl, bcs="damp", fs=True, topology=('*', 1, 1))
    assert fsm.grid.shape == (len(parts[w.rank]), n[1] + 2 * nbl, n[2] + nbl)
    assert tuple(float(o) for o in fsm.grid.origin) == (-10. * nbl, -10. * nbl, 0.)
    fs_full = damp_profile(fsm.grid.shape_global, fsm.padsizes, fsm.grid.spacing)
    assert not fs_full[nbl:-nbl, nbl:-nbl, 0].any() and fs_full[nbl:-nbl, nbl:-nbl, -1].all()
    assert np.array_equal(np.asarray(fsm.damp.data), fs_full[lo:hi])
    fs_op = AcousticWaveSolver(fsm, setup_geometry(fsm, 30.0), space_order=so).op_fwd()
    assert fs_op.backend == 'cuda-sm100a' and fs_op._plan['free_surface']
    # array-valued TTI parameters: every rank holds its slab of each table's source array with valid halo planes
    # (the factor tables of k_tti_fused<.., ARR> are built from them, halo included), operator recognised
    th = (0.2 + 0.5 * vp / 3.0).astype(np.float32)
    tm = SeismicModel(origin=(0., 0., 0.), spacing=(10., 10., 10.), shape=n, space_order=so, vp=vp, nbl=nbl, bcs="damp",
                      epsilon=(0.1 * (vp - 1.5)).astype(np.float32), delta=(0.05 * (vp - 1.5)).astype(np.float32),
                      theta=th, phi=(0.5 * th).astype(np.float32), topology=('*', 1, 1))
    th_full = np.pad(np.pad(th, nbl, mode='edge'), so, mode='edge')
    assert np.array_equal(np.asarray(tm.theta.data_with_halo), th_full[lo:hi + 2 * so])
    t_op = AnisotropicWaveSolver(tm, setup_geometry(tm, 30.0), space_order=so).op_fwd()
    assert t_op.backend == 'cuda-sm100a'
    dist.barrier()
    if w.rank == 0:
        print('DIST-HOST-OK')


def gpu_mode(kind):
    import torch
    import torch.distributed as dist
    from dist_cases import run_case
    w = dv.init_distributed()
    local = int(os.environ.get('LOCAL_RANK', '0'))
    dv.configuration['deviceid'] = local
    res, model = run_case(kind, w.size, topology=('*', 1, 1))
    for name, arr in res.items():
        if name == 'rec':
            if w.rank == 0:
                np.save(f'/tmp/b2_dist_{kind}_rec.npy', np.asarray(arr))
        else:
            np.save(f'/tmp/b2_dist_{kind}_{name}_{w.rank}.npy', np.asarray(arr))
    dist.barrier()
    from devito_b200.distributed import finalize_distributed
    finalize_distributed()
    if w.rank == 0:
        print('DIST-GPU-DONE')


if __name__ == '__main__':
    if sys.argv[1] == 'host':
        host_mode()
    else:
        gpu_mode(sys.argv[2])
