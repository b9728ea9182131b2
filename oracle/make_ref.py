"""Emit the REFERENCE's own generated C for the hot-path operators into oracle/_ref/.

Run in the build container (needs /root/reference + oracle/refshim):
    python oracle/make_ref.py
For each operator the reference's compiler pipeline (devito Operator -> C string,
devito/operator/operator.py:832-835) is run unmodified with its CPU OpenMP backend; the
resulting translation unit is written to oracle/_ref/<name>.c with a JSON sidecar listing the
parameter order (op.parameters). These are OUTPUTS of the reference, not its sources; the
directory is git-ignored but travels to the GPU box, where oracle/refrun.py compiles them with
the reference's own flags (-O3 -march=native -ffast-math -fopenmp) and times them as the
`kind: "reference"` CPU baseline.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, '_ref')

CHILD = r'''
import json, os, sys
import numpy as np
from examples.seismic import demo_model, setup_geometry
from examples.seismic.acoustic import AcousticWaveSolver
from examples.seismic.tti import AnisotropicWaveSolver
out = sys.argv[1]
def emit(name, op):
    with open(os.path.join(out, name + '.c'), 'w') as f:
        f.write(str(op))
    params = []
    for p in op.parameters:
        params.append({'name': p.name, 'kind': type(p).__name__,
                       'is_fn': bool(getattr(p, 'is_AbstractFunction', False) or getattr(p, 'is_DiscreteFunction', False))})
    with open(os.path.join(out, name + '.json'), 'w') as f:
        json.dump({'name': op.name, 'parameters': params}, f, indent=1)
    print('emitted', name)
for so in (8, 12):
    m = demo_model('constant-isotropic', spacing=(10., 10., 10.), shape=(16, 16, 16), nbl=4, space_order=so, dtype=np.float32)
    g = setup_geometry(m, 20.0)
    emit(f'forward_iso_so{so}', AcousticWaveSolver(m, g, space_order=so).op_fwd())
m = demo_model('constant-tti', spacing=(10., 10., 10.), shape=(16, 16, 16), nbl=4, space_order=8, dtype=np.float32)
g = setup_geometry(m, 20.0)
emit('forward_tti_so8', AnisotropicWaveSolver(m, g, space_order=8).op_fwd())
'''


def main():
    os.makedirs(OUT, exist_ok=True)
    env = dict(os.environ)
    env['PYTHONPATH'] = os.pathsep.join([os.path.join(HERE, 'refshim'), '/root/reference'])
    env.update(DEVITO_LANGUAGE='openmp', DEVITO_ARCH='gcc', DEVITO_LOGGING='ERROR')
    env.pop('CC', None)
    subprocess.run([sys.executable, '-c', CHILD, OUT], check=True, env=env, cwd='/tmp')


if __name__ == '__main__':
    main()
