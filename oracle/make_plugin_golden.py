"""Golden numbers for tests/test_zy_refplugin.py: the REFERENCE's own examples run with its stock CPU
backend (gcc/OpenMP) in the build container.

  PYTHONPATH=oracle/refshim:/root/reference DEVITO_LANGUAGE=openmp DEVITO_ARCH=gcc \
  DEVITO_LOGGING=ERROR python oracle/make_plugin_golden.py

The GPU test runs the very same example functions with `(Blackwell, 'advanced', 'cuda')` selected
(devito_b200/refplugin.py) and compares norm(rec) / norm(u). The reference's own known-answer values
(acoustic_example.py:80-87: 369.955 linear / 402.216 sinc with a free surface) are checked as well.
"""
import json
import os

import numpy as np

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden', 'refplugin_norms.json')


def main():
    from devito import norm
    from examples.seismic.acoustic.acoustic_example import run as arun
    from examples.seismic.tti.tti_example import run as trun
    out = {}
    for tag, kw in [('iso_fs_linear', dict(fs=True, interpolation='linear')),
                    ('iso_fs_sinc', dict(fs=True, interpolation='sinc')),
                    ('iso_layers_so4', dict(fs=False)),
                    ('iso_const_so8', dict(fs=False, preset='constant-isotropic', space_order=8, nbl=20)),
                    ('iso_ot4_so8', dict(fs=False, space_order=8, kernel='OT4', nbl=20))]:
        _, _, _, [rec, u] = arun(dtype=np.float32, **kw)
        out[tag] = {'norm_rec': float(norm(rec)), 'norm_u': float(np.linalg.norm(np.asarray(u, dtype=np.float64)))}
        print(tag, out[tag])
    for tag, kw in [('tti_layers_so4', dict()), ('tti_const_so8', dict(preset='constant-tti', space_order=8))]:
        _, _, _, [rec, u, v] = trun(dtype=np.float32, **kw)
        out[tag] = {'norm_rec': float(norm(rec)), 'norm_u': float(norm(u)), 'norm_v': float(norm(v))}
        print(tag, out[tag])
    with open(OUT, 'w') as f:
        json.dump(out, f, indent=1)


if __name__ == '__main__':
    main()
