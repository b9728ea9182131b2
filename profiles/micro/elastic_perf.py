"""Throughput of the generic system executor on the reference's 3-D elastic example (through the plugin) next to the
reference's own CPU run of the same Operator objects. Usage (GPU box):
  PYTHONPATH=$ROOT:$ROOT/oracle/refshim:$ROOT/baseline/_ref DEVITO_ARCH=gcc DEVITO_LOGGING=ERROR python profiles/micro/elastic_perf.py"""
import time, sys
import numpy as np
import devito
import devito_b200.refplugin as rp
rp.activate()
from examples.seismic import demo_model, setup_geometry
from examples.seismic.elastic import ElasticWaveSolver
n, so, nbl = int(sys.argv[1]) if len(sys.argv) > 1 else 200, 8, 20
model = demo_model('layers-elastic', space_order=so, shape=(n, n, n), nbl=nbl, spacing=(10.,) * 3, dtype=np.float32)
solver = ElasticWaveSolver(model, setup_geometry(model, 150.0), space_order=so)
nt = solver.geometry.nt
N = (n + 2 * nbl) ** 3
solver.forward()                                           # warm-up: builds, tabulates, stages
t0 = time.perf_counter(); out = solver.forward(); t1 = time.perf_counter()
summ = out[-1]
print(f"GPU: grid {(n+2*nbl)}^3, {nt} steps: wall {t1-t0:.3f} s (host-staged, coefficient tabulation on the host included), "
      f"stages {summ['section0'].time:.3f} s -> {N*(nt-1)/summ['section0'].time/1e9:.2f} GPts/s (grid points x steps; 9 fields per point)")
with rp.reference_cpu():
    solver.forward()
    t0 = time.perf_counter(); out = solver.forward(); t1 = time.perf_counter()
print(f"CPU (reference, OpenMP): wall {t1-t0:.3f} s -> {N*(nt-1)/(t1-t0)/1e9:.3f} GPts/s")
