"""Worker for tests/test_zy_refplugin.py — runs in its own interpreter because here `import devito` is the
REAL reference (installed unmodified under baseline/_ref, or /root/reference in the build container), made
importable next to devito_b200 by the launcher's PYTHONPATH.

  ffi : no GPU. `(Blackwell, 'advanced', 'cuda')` is selected, the reference's examples build their
        operators, and `b2_iso_forward` / `b2_tti_forward` are replaced by recording doubles that check
        the structs the plugin passes: they must be the reference's own `struct dataobj` (same data
        pointers as the reference's arrays, allocated extents, coordinate tables).
  gpu : the reference's examples/seismic/{acoustic,tti} example functions run unchanged on the GPU;
        norms against the reference's own known-answer values and its CPU run (tests/golden).
"""
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

import devito                                             # noqa: E402  (the reference)
import devito_b200.refplugin as rp                        # noqa: E402
from devito_b200 import _lib as L_                        # noqa: E402

assert 'devito_b200' not in devito.__file__
rp.activate()

from devito import configuration, norm                    # noqa: E402
from examples.seismic import demo_model, setup_geometry   # noqa: E402
from examples.seismic.acoustic import AcousticWaveSolver  # noqa: E402
from examples.seismic.tti import AnisotropicWaveSolver    # noqa: E402


def ffi():
    calls = []

    class Double:
        """Stands in for libb200stencil.so: records what crosses the C ABI."""
        def b2_iso_forward(self, ref):
            a = ref._obj
            u = a.u.contents
            calls.append(dict(kind='iso', ndim=a.ndim, so=a.space_order, R=a.radius, u_data=u.data,
                              u_size=[u.size[i] for i in range(4)], u_dmap=u.dmap, dt=a.dt, vp=a.vp,
                              param_kind=a.param_kind, param_data=a.param.contents.data if a.param else None,
                              damp_data=a.damp.contents.data, time=(a.time_m, a.time_M),
                              box=(a.x_m, a.x_M, a.y_m, a.y_M, a.z_m, a.z_M), fs=a.free_surface, adjoint=a.adjoint,
                              src=dict(data=a.src.contents.data.contents.data, gp=a.src.contents.gp.contents.data,
                                       w0=a.src.contents.w[0].contents.data, r=a.src.contents.r,
                                       p=(a.src.contents.p_m, a.src.contents.p_M)),
                              rec=dict(data=a.rec.contents.data.contents.data, gp=a.rec.contents.gp.contents.data,
                                       nrec=a.rec.contents.data.contents.size[1], r=a.rec.contents.r)))
            return 0

        def b2_tti_forward(self, ref):
            a = ref._obj
            calls.append(dict(kind='tti', so=a.space_order, R=a.radius, u_data=a.u.contents.data,
                              v_data=a.v.contents.data, scal=(a.vp, a.epsilon, a.delta, a.theta, a.phi),
                              arrays=[bool(x) for x in (a.vp_arr, a.epsilon_arr, a.delta_arr, a.theta_arr, a.phi_arr)]))
            return 0

        def b2_system_forward(self, ref):
            """NumPy emulation of the tap-table executor (csrc/b2_api_system.cu) on the structs that cross the ABI."""
            a = ref._obj
            nd, h = a.ndim, a.halo

            def arr(dobj, ndim):
                shape = tuple(dobj.size[i] for i in range(ndim))
                n = int(np.prod(shape))
                return np.ctypeslib.as_array((ctypes.c_float * n).from_address(dobj.data)).reshape(shape)
            F = [arr(a.fields[i].contents, nd + 1) for i in range(a.nfields)]
            C = [arr(a.coefs[i].contents, nd) for i in range(a.ncoefs)]
            lo = [a.x_m, a.y_m, a.z_m][:nd]
            hi = [a.x_M, a.y_M, a.z_M][:nd]
            box = tuple(slice(l + h, u + h + 1) for l, u in zip(lo, hi))
            cbox = tuple(slice(l, u + 1) for l, u in zip(lo, hi))
            ntaps = 0
            for time in range(a.time_m, a.time_M + 1):
                for s in range(a.nstages):
                    st = a.stages[s]
                    acc = np.zeros([u - l + 1 for l, u in zip(lo, hi)], dtype=np.float32)
                    for i in range(st.ntaps):
                        t = st.taps[i]
                        src = F[t.field][(time + t.tshift) % F[t.field].shape[0]]
                        sl = tuple(slice(l + h + t.off[d], u + h + 1 + t.off[d]) for d, (l, u) in enumerate(zip(lo, hi)))
                        c = np.float32(t.coef) * (C[t.cfield][cbox] if t.cfield >= 0 else np.float32(1.0))
                        acc += c * src[sl]
                        ntaps += 1
                    out = F[st.out_field][(time + st.out_tshift) % F[st.out_field].shape[0]]
                    out[box] = acc
                for i in range(a.ninject):
                    q = a.inject[i]
                    sp = q.s.contents
                    data = arr(sp.data.contents, 2)
                    gp = np.ctypeslib.as_array((ctypes.c_int * (data.shape[1] * nd)).from_address(sp.gp.contents.data)).reshape(-1, nd)
                    ws = [arr(sp.w[d].contents, 2) for d in range(nd)]
                    r = sp.r
                    for p in range(sp.p_m, sp.p_M + 1):
                        for off in np.ndindex(*([2 * r] * nd)):
                            cell = [gp[p, d] + off[d] - r + 1 for d in range(nd)]
                            if any(c < l - r or c > u + r for c, l, u in zip(cell, lo, hi)):
                                continue
                            w = np.prod([ws[d][p, off[d]] for d in range(nd)])
                            sc = q.scale
                            if q.param_kind:
                                pv = arr(q.param.contents, nd)[tuple(c + h for c in cell)]
                                sc = q.scale * pv * pv if q.param_kind == 1 else q.scale / pv
                            for j in range(q.nfields):
                                f = F[q.fields[j]]
                                f[(time + q.tshift) % f.shape[0]][tuple(c + h for c in cell)] += np.float32(w * data[time, p] * sc)
                for i in range(a.ninterp):
                    q = a.interp[i]
                    sp = q.s.contents
                    data = arr(sp.data.contents, 2)
                    gp = np.ctypeslib.as_array((ctypes.c_int * (data.shape[1] * nd)).from_address(sp.gp.contents.data)).reshape(-1, nd)
                    ws = [arr(sp.w[d].contents, 2) for d in range(nd)]
                    r = sp.r
                    f = F[q.field][(time + q.tshift) % F[q.field].shape[0]]
                    if time >= data.shape[0]:
                        continue
                    for p in range(sp.p_m, sp.p_M + 1):
                        tot = 0.0
                        for off in np.ndindex(*([2 * r] * nd)):
                            cell = [gp[p, d] + off[d] - r + 1 for d in range(nd)]
                            if any(c < l - r or c > u + r for c, l, u in zip(cell, lo, hi)):
                                continue
                            tot += np.prod([ws[d][p, off[d]] for d in range(nd)]) * f[tuple(c + h for c in cell)]
                        data[time, p] = tot
            calls.append(dict(kind='system', nfields=a.nfields, ncoefs=a.ncoefs, nstages=a.nstages, ntaps=ntaps))
            return 0

        def b2_last_error(self):
            return b''

    L_.lib = lambda: Double()
    # no device here: the tabulated coefficient arrays stay host arrays for the emulation
    rp._coef_resident = lambda arr: L_.make_dataobj(host=np.ascontiguousarray(arr, dtype=np.float32))
    kw = dict(shape=(20, 20, 20), nbl=6, spacing=(20., 20., 20.), dtype=np.float32)
    # layered velocity (array parameter), free surface
    model = demo_model('layers-isotropic', space_order=4, fs=True, **kw)
    geometry = setup_geometry(model, 100.0)
    solver = AcousticWaveSolver(model, geometry, space_order=4)
    op = solver.op_fwd()
    assert type(op).__name__ == 'B200CudaOperator' and op.backend == 'cuda-sm100a', op._b200_why
    u = devito.TimeFunction(name='u', grid=model.grid, time_order=2, space_order=4)
    rec = geometry.rec
    solver.forward(u=u, rec=rec)
    c = calls[-1]
    assert c['kind'] == 'iso' and c['fs'] == 1 and c['so'] == 4 and c['R'] == 2 and c['ndim'] == 3
    assert c['u_data'] == u._data.ctypes.data and c['u_dmap'] is None          # the reference's own array
    assert c['u_size'] == list(u._data.shape) == [3, 40, 40, 34]
    assert c['damp_data'] == model.damp._data.ctypes.data
    assert c['param_kind'] == 1 and c['param_data'] == model.vp._data.ctypes.data
    assert c['rec']['data'] == rec._data.ctypes.data and c['rec']['nrec'] == rec.npoint
    assert c['src']['r'] == 1 and c['src']['p'] == (0, 0)
    assert abs(c['dt'] - float(model.critical_dt)) < 1e-6
    assert c['time'] == (1, geometry.nt - 2) and c['box'] == (0, 31, 0, 31, 0, 25)
    # scalar velocity: the Constant travels as a number; runtime overrides go through the reference's arguments()
    m2 = demo_model('constant-isotropic', space_order=8, **kw)
    s2 = AcousticWaveSolver(m2, setup_geometry(m2, 100.0, interpolation='sinc'), space_order=8)
    s2.forward()
    assert calls[-1]['param_kind'] == 0 and abs(calls[-1]['vp'] - 1.5) < 1e-6 and calls[-1]['src']['r'] == 4
    s2.forward(vp=2.0)
    assert abs(calls[-1]['vp'] - 2.0) < 1e-6
    s2.forward(vp=devito.Constant(name='v', value=2.5, dtype=np.float32))
    assert abs(calls[-1]['vp'] - 2.5) < 1e-6
    s2.adjoint(rec=s2.geometry.rec)
    assert calls[-1]['adjoint'] == 1
    # TTI
    m3 = demo_model('constant-tti', space_order=8, **kw)
    s3 = AnisotropicWaveSolver(m3, setup_geometry(m3, 100.0), space_order=8)
    s3.forward()
    c = calls[-1]
    assert c['kind'] == 'tti' and c['R'] == 4 and not any(c['arrays'])
    assert np.allclose(c['scal'], (1.5, .3, .2, .7, .35), atol=1e-6)
    m4 = demo_model('layers-tti', space_order=4, **kw)
    AnisotropicWaveSolver(m4, setup_geometry(m4, 100.0), space_order=4).forward()
    assert calls[-1]['kind'] == 'tti' and all(calls[-1]['arrays'])
    # first-order systems on staggered grids (elastic, viscoelastic, staggered TTI): the tap tables and coefficient
    # arrays the plugin derives from the reference's evaluated equations, executed by the NumPy emulation of
    # `b2_system_forward` above, must reproduce the reference's own CPU run of the same Operator object
    from examples.seismic.elastic import ElasticWaveSolver
    from examples.seismic.viscoelastic import ViscoelasticWaveSolver

    def fields_of(out):
        res = {}
        for o in out:
            if hasattr(o, 'values') and not hasattr(o, 'data'):       # TensorTimeFunction
                res.update({f.name: np.array(f.data) for f in o.values()})
            elif hasattr(o, '__iter__') and not hasattr(o, 'data'):   # VectorTimeFunction
                res.update({f.name: np.array(f.data) for f in o})
            elif hasattr(o, 'data'):
                res[o.name] = np.array(o.data)
        return res

    errs = {}
    cases = [('elastic-3d', ElasticWaveSolver, 'layers-elastic', (14, 12, 13), {}, 10, 10),
             ('elastic-2d', ElasticWaveSolver, 'layers-elastic', (24, 21), {}, 6, 6),
             ('viscoelastic-2d', ViscoelasticWaveSolver, 'layers-viscoelastic', (22, 20), {}, 9, 9),
             ('tti-staggered-3d', AnisotropicWaveSolver, 'layers-tti', (13, 12, 14), dict(kernel='staggered'), None, None),
             ('tti-staggered-2d', AnisotropicWaveSolver, 'layers-tti', (22, 20), dict(kernel='staggered'), None, None)]
    for tag, cls, preset, shape, skw, nstages, nfields in cases:
        nd = len(shape)
        me = demo_model(preset, space_order=4, shape=shape, nbl=4, spacing=(10.,) * nd, dtype=np.float32)
        se = cls(me, setup_geometry(me, 22.0), space_order=4, **skw)
        ope = se.op_fwd()
        assert ope.backend == 'cuda-sm100a' and ope._b200_sys is not None, (tag, ope._b200_why)
        got = fields_of(se.forward()[:-1])
        assert calls[-1]['kind'] == 'system', tag
        if nstages is not None:
            assert calls[-1]['nstages'] == nstages and calls[-1]['nfields'] == nfields, (tag, calls[-1])
        with rp.reference_cpu():                           # the same Operator objects on the reference's CPU path
            ref = fields_of(se.forward()[:-1])
        assert set(ref) == set(got) and len(ref) >= 3, (tag, sorted(ref))
        assert any(np.any(got[n] != ref[n]) for n in ref), tag          # two different code paths did run
        for name in ref:
            scale = max(float(np.abs(ref[name]).max()), 1e-30)
            err = float(np.abs(got[name] - ref[name]).max()) / scale
            # staggered TTI: sin/cos of the angle fields are tabulated with NumPy here and by the C library's sinf/cosf
            # in the reference's generated code (last-bit differences, amplified by the short run on a tiny grid)
            assert err < (5e-4 if 'tti' in tag else 5e-5), (tag, name, err)
            assert float(np.abs(ref[name]).max()) > 0, (tag, name)
            errs[(tag, name)] = err
    # the tabulated coefficient arrays of a system are cached across applies and re-tabulated when a parameter changes
    ntab = []
    real_eval = rp._eval_coef
    rp._eval_coef = lambda *a, **k: (ntab.append(1), real_eval(*a, **k))[1]
    try:
        se.forward()
        assert not ntab, "same parameters: the cached coefficient arrays must be reused"
        me.vp.data[:] = me.vp.data * 1.01
        se.forward()
        assert ntab, "a changed parameter array must trigger a new tabulation"
    finally:
        rp._eval_coef = real_eval
    # the set-up operators stayed on the reference's CPU path (and ran: the damping profile is there)
    assert float(np.max(model.damp.data)) > 0
    print('REFPLUGIN-FFI-OK', len(calls), {k: float(f'{v:.2e}') for k, v in errs.items() if v > 2e-5})


def gpu():
    from examples.seismic.acoustic.acoustic_example import run as arun
    from examples.seismic.tti.tti_example import run as trun
    with open(os.path.join(ROOT, 'tests', 'golden', 'refplugin_norms.json')) as f:
        gold = json.load(f)
    L = L_.lib()
    out = {}
    n0 = L.b2_launch_count()
    # the reference's own known answers (acoustic_example.py:80-87)
    for interp, kat in (('linear', 369.955), ('sinc', 402.216)):
        _, _, _, [rec, u] = arun(fs=True, dtype=np.float32, interpolation=interp)
        got = float(norm(rec))
        out[f'iso_fs_{interp}'] = got
        assert np.isclose(got, kat, rtol=1e-3, atol=0), (interp, got, kat)
        assert np.isclose(got, gold[f'iso_fs_{interp}']['norm_rec'], rtol=2e-4), (interp, got)
    for tag, kw in [('iso_layers_so4', dict(fs=False)),
                    ('iso_const_so8', dict(fs=False, preset='constant-isotropic', space_order=8, nbl=20)),
                    ('iso_ot4_so8', dict(fs=False, space_order=8, kernel='OT4', nbl=20))]:
        _, _, _, [rec, u] = arun(dtype=np.float32, **kw)
        got = float(norm(rec)), float(np.linalg.norm(np.asarray(u, dtype=np.float64)))
        out[tag] = got
        assert np.isclose(got[0], gold[tag]['norm_rec'], rtol=2e-4), (tag, got, gold[tag])
        assert np.isclose(got[1], gold[tag]['norm_u'], rtol=1e-3 if 'ot4' in tag else 2e-4), (tag, got, gold[tag])
    for tag, kw in [('tti_layers_so4', dict()), ('tti_const_so8', dict(preset='constant-tti', space_order=8))]:
        _, _, _, [rec, u, v] = trun(dtype=np.float32, **kw)
        got = float(norm(rec)), float(norm(u)), float(norm(v))
        out[tag] = got
        assert np.isclose(got[0], gold[tag]['norm_rec'], rtol=1e-3), (tag, got, gold[tag])
        assert np.isclose(got[1], gold[tag]['norm_u'], rtol=1e-3), (tag, got, gold[tag])
        assert np.isclose(got[2], gold[tag]['norm_v'], rtol=1e-3), (tag, got, gold[tag])
    # first-order systems on staggered grids through b2_system_forward: the reference's own examples against their
    # known answers (elastic_example.py:44-45, viscoelastic_example.py:45-46), and GPU vs the reference's CPU run of
    # the same Operator object (3-D elastic, staggered TTI 2-D / 3-D)
    from examples.seismic.elastic.elastic_example import run as erun
    from examples.seismic.viscoelastic.viscoelastic_example import run as verun
    from examples.seismic.elastic import ElasticWaveSolver
    from examples.seismic.viscoelastic import ViscoelasticWaveSolver
    n1 = L.b2_launch_count()
    _, _, _, [rec1, rec2, v, tau] = erun(dtype=np.float32)
    out['elastic_2d'] = (float(norm(rec1)), float(norm(rec2)))
    assert np.isclose(out['elastic_2d'][0], 19.9367, atol=1e-3, rtol=0), out['elastic_2d']
    assert np.isclose(out['elastic_2d'][1], 0.6689, atol=1e-3, rtol=0), out['elastic_2d']
    _, _, _, [rec1, rec2, v, tau] = verun(dtype=np.float32)
    out['viscoelastic_2d'] = (float(norm(rec1)), float(norm(rec2)))
    assert np.isclose(out['viscoelastic_2d'][0], 12.62339, atol=1e-3, rtol=0), out['viscoelastic_2d']
    assert np.isclose(out['viscoelastic_2d'][1], 0.330103, atol=1e-3, rtol=0), out['viscoelastic_2d']
    assert int(L.b2_launch_count() - n1) > 1000

    def fields_of(res):
        d = {}
        for o in res:
            if hasattr(o, 'values') and not hasattr(o, 'data'):
                d.update({f.name: np.array(f.data) for f in o.values()})
            elif hasattr(o, '__iter__') and not hasattr(o, 'data'):
                d.update({f.name: np.array(f.data) for f in o})
            elif hasattr(o, 'data'):
                d[o.name] = np.array(o.data)
        return d
    for tag, cls, preset, shape, skw in [('elastic-3d', ElasticWaveSolver, 'layers-elastic', (40, 36, 38), {}),
                                         ('viscoelastic-3d', ViscoelasticWaveSolver, 'layers-viscoelastic', (30, 28, 26), {}),
                                         ('tti-staggered-3d', AnisotropicWaveSolver, 'layers-tti', (36, 40, 34), dict(kernel='staggered')),
                                         ('tti-staggered-2d', AnisotropicWaveSolver, 'layers-tti', (80, 90), dict(kernel='staggered'))]:
        nd = len(shape)
        me = demo_model(preset, space_order=8 if 'tti' in tag else 4, shape=shape, nbl=10, spacing=(10.,) * nd, dtype=np.float32)
        se = cls(me, setup_geometry(me, 120.0), space_order=8 if 'tti' in tag else 4, **skw)
        ope = se.op_fwd()
        assert ope.backend == 'cuda-sm100a' and ope._b200_sys is not None, (tag, ope._b200_why)
        nb = L.b2_launch_count()
        got = fields_of(se.forward()[:-1])
        assert L.b2_launch_count() > nb
        with rp.reference_cpu():                           # the same Operator objects on the reference's CPU path
            ref = fields_of(se.forward()[:-1])
        worst = 0.0
        for name in ref:
            scale = max(float(np.abs(ref[name]).max()), 1e-30)
            worst = max(worst, float(np.abs(got[name] - ref[name]).max()) / scale)
        out[tag] = worst
        assert worst < 1e-4, (tag, worst)
    launches = int(L.b2_launch_count() - n0)
    assert launches > 1000, launches                  # the propagators ran in libb200stencil.so
    print('REFPLUGIN-GPU-OK', json.dumps({'launches': launches, 'norms': out}))


if __name__ == '__main__':
    {'ffi': ffi, 'gpu': gpu}[sys.argv[1]]()
