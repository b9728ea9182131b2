#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (driver contract).

Metric (BASELINE.json): GPts/s of the 3-D isotropic acoustic forward operator, space_order=8,
on a 1024^3 grid (absorbing layers included: GPts/s counts every grid point the stencil
updates, devito/operator/profiling.py:355-366), fp32, synthetic constant-velocity model,
one Ricker source, receivers sampled every step.

A "step" = one `Operator.apply` of the Forward operator over NT time steps.
  value : whole-job GPts/s with fields resident in HBM when the timed region starts
  e2e   : the same call made with HOST buffers through the C ABI (H2D of u/damp/src and D2H of
          u/rec inside the timed region — the reference's per-apply copy semantics)
  roofline : the stencil kernel's algorithmic bytes (16 B/point) / its mean launch duration
             (CUDA events on the library stream), against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline : the reference's CPU implementation timed on this box's host cores

`--impl reference` times the reference's own CPU code path (oracle/_ref = C code emitted by the
reference's code generator for this operator, compiled here with its flags; else the oracle
port) on a bounded sample of the same workload.

Multi-GPU (torchrun, one rank per GPU): x-slab decomposition, weak scaling — every rank owns a
1024-plane slab of a (N*1024) x 1024 x 1024 grid; boundary planes are stored into the neighbour
GPU's halo over NVLink (CUDA-IPC peer memory + device flags), or exchanged by NCCL send/recv
overlapped with the interior update (B2_HALO=nccl, and the first step of every apply).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_ALG = 16.0    # algorithmic bytes / point / step: u[t] r + u[t-1] r + damp r + u[t+1] w


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=4)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--grid', type=int, default=int(os.environ.get('B2_BENCH_GRID', 1024)),
                    help='grid points per dimension incl. absorbing layers')
    ap.add_argument('--nt', type=int, default=int(os.environ.get('B2_BENCH_NT', 256)),
                    help='time steps per Operator.apply')
    ap.add_argument('--space-order', type=int, default=8)
    ap.add_argument('--workload', default='iso', choices=['iso', 'tti'],
                    help="iso: the headline metric; tti: BASELINE config 4 (not the driver's line)")
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured'
    return 6650.0, 'fallback'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.index = index
        self.samples = []
        self._stop = False
        self._t = None

    def _run(self):
        while not self._stop:
            try:
                out = subprocess.run(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                      '-i', str(self.index)], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(',')]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop = True
        self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith('active') for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons}


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference's implementation of the path on the host cores
# ---------------------------------------------------------------------------------------------
def usable_cores():
    """Host threads this process may actually use: affinity mask capped by the cgroup CPU quota."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    try:
        with open('/sys/fs/cgroup/cpu.max') as f:
            q, per = f.read().split()
        if q != 'max':
            n = max(1, min(n, int(float(q) / float(per) + 0.5)))
    except Exception:
        pass
    return n


_cpu_problem_cache = {}


def _cpu_problem(so, grid_n, nts):
    key = (so, grid_n, nts)
    if key in _cpu_problem_cache:
        return _cpu_problem_cache[key]
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from oracle import oracle as O
    from helpers import iso_problem
    nbl = 40 if grid_n >= 160 else 8
    n = grid_n - 2 * nbl
    p = iso_problem(n, nbl, so, tn=1.0)            # geometry; the time range is overridden
    src = dict(p['src'], data=np.ascontiguousarray(np.resize(p['src']['data'], (nts, 1)).astype(np.float32)))
    rec_c = p['rec_coords'][:: max(1, len(p['rec_coords']) // 512)][:512]
    rgp, rw = O.tabulate(rec_c.astype(np.float32), p['origin'], p['spacing'])
    rec = dict(data=np.zeros((nts, len(rec_c)), dtype=np.float32), gp=rgp, w=rw, r=1)
    _cpu_problem_cache[key] = (p, src, rec)
    return _cpu_problem_cache[key]


def _cpu_run_once(so, grid_n, nt, threads):
    from oracle import oracle as O
    from oracle import refrun
    p, src, rec = _cpu_problem(so, grid_n, 512)
    ref = refrun.load_forward(so)
    t0 = time.perf_counter()
    if ref is not None:
        kind = 'reference'
        refrun.run_forward(ref, p['u'], p['damp'], 1.5, p['dt'], 1, nt, src, rec, so, threads)
    else:
        kind = 'port'
        os.environ['OMP_NUM_THREADS'] = str(threads)
        O.iso_forward(p['u'], so, p['w'], p['dt'], 1, nt, damp=p['damp'], vp=1.5, src=src, rec=rec, fast=True)
    return time.perf_counter() - t0, kind


_cpu_threads_choice = {}


def cpu_reference_run(so, grid_n, nt, threads=None, budget_s=12.0):
    """Time the reference's CPU code path on a bounded sample of the workload: the same operator
    (iso acoustic, same space order, source + 512 receivers) on a grid_n^3 grid. The thread count
    is the best of a short probe over {all usable cores, 1/2, 1/4, ...} (the reference's own
    advice is one thread per physical core, benchmarks/user/README.md:22-64); the number of time
    steps is then sized to ~budget_s seconds. Returns (GPts/s, kind, sample, cores, seconds)."""
    cores = usable_cores()
    if threads is None:
        if so not in _cpu_threads_choice:
            cands = sorted({max(1, cores), max(1, cores // 2), max(1, cores // 4), min(cores, 16), min(cores, 8)},
                           reverse=True)
            best = None
            _cpu_run_once(so, grid_n, 1, cands[0])                      # build / first touch
            for c in cands:
                el, _ = _cpu_run_once(so, grid_n, 2, c)
                if best is None or el < best[0]:
                    best = (el, c)
            _cpu_threads_choice[so] = best[1], best[0] / 2
        threads, per_step = _cpu_threads_choice[so]
        nt = int(max(4, min(500, budget_s / max(per_step, 1e-4))))
    el, kind = _cpu_run_once(so, grid_n, nt, threads)
    pts = float(grid_n) ** 3 * nt
    sample = (f"iso so={so} {grid_n}^3 x {nt} steps (same operator, smaller grid), {threads} of "
              f"{cores} usable host threads")
    return pts / el / 1e9, kind, sample, threads, el


def run_reference_arm(a):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    vals = []
    budget = max(3.0, min(15.0, 150.0 / max(1, a.warmup + a.steps)))   # whole arm within minutes
    for i in range(a.warmup + a.steps):
        gp, kind, sample, cores, el = cpu_reference_run(a.space_order, 384, 8, budget_s=budget)
        if i >= a.warmup:
            vals.append((gp, el))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([e for _, e in vals])) * 1e3
    line = {"impl": "reference", "metric": "GPts/s (3D isotropic acoustic forward, so=%d, 1024^3 per GPU)" % a.space_order,
            "value": value, "unit": "GPts/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"3D isotropic acoustic so={a.space_order}; the reference's CPU (OpenMP) "
                                   f"path on a bounded sample: {sample}",
                       "l2": "inputs larger than the CPU caches"},
            "cpu_baseline": {"value": value, "unit": "GPts/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "GPts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------
def main():
    a = parse()
    if a.impl == 'reference':
        run_reference_arm(a)
        return
    import torch
    import torch.distributed as dist
    import devito_b200 as dv
    from devito_b200 import _lib
    from devito_b200.seismic import (SeismicModel, AcquisitionGeometry, AcousticWaveSolver,
                                     AnisotropicWaveSolver, TimeAxis)

    world = dv.init_distributed()
    rank, nranks = world.rank, world.size
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dv.configuration['deviceid'] = local
    dev_ = torch.device('cuda', local)
    L = _lib.lib()

    so, G, NT, nbl = a.space_order, a.grid, a.nt, 40
    n = G - 2 * nbl
    # global grid: (nranks*G) x G x G, slab-decomposed along x (weak scaling)
    shape = (nranks * G - 2 * nbl, n, n)
    tti = a.workload == 'tti'
    extra = dict(epsilon=.3, delta=.2, theta=.7, phi=.35) if tti else {}
    model = SeismicModel(origin=(0., 0., 0.), spacing=(10., 10., 10.), shape=shape, space_order=so,
                         vp=1.5, nbl=nbl, bcs="damp", topology=('*', 1, 1) if nranks > 1 else None,
                         **extra)
    dt = model.critical_dt
    tn = float(dt) * (NT + 1)
    src_c = np.array([[model.domain_size[0] * .5, model.domain_size[1] * .5, 10.0]])
    rx = np.linspace(0, model.domain_size[0], 32)
    ry = np.linspace(0, model.domain_size[1], 16)
    rec_c = np.array([[x, y, 20.0] for x in rx for y in ry])           # 512 receivers
    geometry = AcquisitionGeometry(model, rec_c, src_c, t0=0.0, tn=tn, src_type='Ricker', f0=0.010)
    solver = (AnisotropicWaveSolver if tti else AcousticWaveSolver)(model, geometry, space_order=so)
    nt_steps = geometry.nt - 2                                          # time = 1 .. nt-2
    u = dv.TimeFunction(name='u', grid=model.grid, time_order=2, space_order=so)
    v = dv.TimeFunction(name='v', grid=model.grid, time_order=2, space_order=so) if tti else None
    fkw = dict(v=v) if tti else {}
    b_alg = 28.0 if tti else B_ALG
    src, rec = geometry.src, geometry.rec
    pts_step = float(nranks) * G * G * G * nt_steps                     # points per apply (whole job)

    def barrier():
        torch.cuda.synchronize()
        if nranks > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # The library enqueues on torch's current stream so that CUDA events recorded on that stream
    # bracket exactly its work (torch.cuda.Event only sees torch's current stream).
    bench_stream = torch.cuda.Stream(device=dev_)
    torch.cuda.set_stream(bench_stream)
    L.b2_set_stream(ctypes_voidp(bench_stream.cuda_stream))

    def timed(fn, reps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        dev = e0.elapsed_time(e1) * 1e-3
        barrier()
        # device time between the two events (includes host gaps between applies); max over ranks
        t = torch.tensor([dev, wall], dtype=torch.float64, device=dev_)
        if nranks > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0].item())

    resident = lambda: solver.forward(src=src, rec=rec, u=u, **fkw)
    for _ in range(a.warmup):
        resident()
    launches0 = L.b2_launch_count()
    L.b2_kernel_timing_enable(1)
    L.b2_kernel_timing_reset()
    with ClockSampler(local) as clk:
        t_res = timed(resident, a.steps)
    nl = ctypes_int()
    import ctypes as _ct
    k_ms = L.b2_kernel_timing_ms(_ct.byref(nl))
    L.b2_kernel_timing_enable(0)
    launches = int(L.b2_launch_count() - launches0)
    value = pts_step * a.steps / t_res / 1e9
    clocks = clk.summary()

    peak, peak_kind = peaks()
    pts_launch = float(G) ** 3                  # interior launch of one rank ~ the whole slab
    if nranks > 1:
        pts_launch = None
    roof = None
    traffic = None
    try:   # dram__bytes_read+write per launch from the committed `ncu --set full` capture
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
            traffic = json.load(f).get(f"{a.workload}_so{so}_{G}")
    except Exception:
        traffic = None
    if nl.value and k_ms > 0 and nranks == 1:
        ach = b_alg * pts_launch / (k_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_kind})",
                "kernel": "k_tti_*" if tti else "k_iso_tma", "launch_ms": k_ms, "launches_timed": int(nl.value)}

    e2e = None
    if not a.no_e2e:
        hostcall = lambda: solver.forward(src=src, rec=rec, u=u, resident=False, **fkw)
        _ = u.data_with_halo          # materialise (pinned) host copies outside the timed region
        _ = model.damp.data_with_halo
        for _ in range(min(a.warmup, 1) or 1):
            hostcall()
        t_e2e = timed(hostcall, a.steps)
        h2d = u.storage.host_ro.nbytes + model.damp.storage.host_ro.nbytes + src.data.nbytes
        d2h = u.storage.host_ro.nbytes + rec.data.nbytes
        e2e = {"value": pts_step * a.steps / t_e2e / 1e9, "unit": "GPts/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": t_e2e / a.steps * 1e3}

    cpu = None
    if rank == 0 and nranks == 1 and not a.no_cpu:
        try:
            gp, kind, sample, cores, el = cpu_reference_run(so, 384, 8, budget_s=10.0)
            cpu = {"value": gp, "unit": "GPts/s", "cores": cores, "kind": kind, "sample": sample}
        except Exception as e:                                           # never hide the GPU number
            cpu = {"value": None, "unit": "GPts/s", "cores": None, "kind": "port", "sample": f"failed: {e}"}

    if rank == 0:
        line = {"metric": "GPts/s (3D %s forward, so=%d, %d^3 per GPU)" % ("TTI" if tti else "isotropic acoustic", so, G),
                "value": value, "unit": "GPts/s", "n_gpus": nranks, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": t_res / a.steps * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"3D {'TTI centred' if tti else 'isotropic acoustic'} so={so}, grid {nranks * G}x{G}x{G} "
                                       f"(nbl=40 included), {nt_steps} time steps per apply, 1 Ricker source, "
                                       f"512 receivers, constant vp=1.5",
                           "decomposition": f"x-slabs over {nranks} GPU(s)" if nranks > 1 else "single GPU",
                           "halo": (None if nranks == 1 else
                                    "peer-memory stores over NVLink (CUDA IPC + device flags)"
                                    if getattr(u.storage, 'p2p_registered', False) else "NCCL send/recv"),
                           "l2": "inputs (18 GB/GPU) larger than L2; no flush needed",
                           "time_steps_per_apply": nt_steps, "dt": float(dt)},
                "clocks": clocks, "gpu_launches": launches,
                "roofline": roof, "cpu_baseline": cpu, "e2e": e2e}
        print(json.dumps(line), flush=True)
    if nranks > 1:
        from devito_b200.distributed import finalize_distributed
        finalize_distributed()


def ctypes_int():
    import ctypes
    return ctypes.c_int(0)


def ctypes_voidp(v):
    import ctypes
    return ctypes.c_void_p(v)


if __name__ == '__main__':
    main()
