import numpy as np, sys, os
import devito
import devito_b200.refplugin as rp
from devito_b200 import _lib as L_
rp.activate()
from devito import norm
from examples.seismic import demo_model, setup_geometry
from examples.seismic.elastic import ElasticWaveSolver
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
for shape, nbl, tn in [((24,21),4,22.0), ((50,50),40,200.0), ((50,50),40,1000.0)]:
    me = demo_model('layers-elastic', space_order=4, shape=shape, nbl=nbl, spacing=(20.,)*2, dtype=np.float32)
    se = ElasticWaveSolver(me, setup_geometry(me, tn), space_order=4)
    ope = se.op_fwd()
    got = fields_of(se.forward()[:-1])
    ope._b200_sys = None
    ref = fields_of(se.forward()[:-1])
    print(shape, tn, {n: float(np.abs(got[n]-ref[n]).max()/max(np.abs(ref[n]).max(),1e-30)) for n in ref})
    r2g, r2r = got['rec2'], ref['rec2']
    print('  rec2 norms', np.linalg.norm(r2g), np.linalg.norm(r2r), 'first bad row', next((i for i in range(r2r.shape[0]) if np.abs(r2g[i]-r2r[i]).max() > 1e-4*np.abs(r2r).max()), None), r2r.shape)
    e = np.abs(r2g - r2r); i,j = np.unravel_index(e.argmax(), e.shape); print('  worst at', i, j, r2g[i,j], r2r[i,j])
