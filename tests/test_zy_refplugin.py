"""The reference-side binding, EXECUTED: the real devito (unmodified; baseline/_ref on the GPU box,
/root/reference in the build container) selects `(Blackwell, 'advanced', 'cuda')`, which
devito_b200/refplugin.py registers (devito/operator/registry.py:33-57), and the reference's own
examples/seismic run unchanged with their wave propagators executing in libb200stencil.so.
Runs in a sub-process (here `import devito` must be the reference, not this package's alias)."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _reference_path():
    for cand in (os.path.join(ROOT, 'baseline', '_ref'), '/root/reference'):
        if os.path.isdir(os.path.join(cand, 'devito')) and os.path.isdir(os.path.join(cand, 'examples')):
            return cand
    return None


def _run(mode, timeout):
    ref = _reference_path()
    if ref is None:
        pytest.skip("the reference is not available (baseline/_ref: pip install --no-deps --target baseline/_ref "
                    "/root/reference)")
    env = dict(os.environ)
    env['PYTHONPATH'] = os.pathsep.join([ROOT, os.path.join(ROOT, 'oracle', 'refshim'), ref])
    env.update(DEVITO_ARCH='gcc', DEVITO_LOGGING='ERROR', OMP_NUM_THREADS='4')
    env.pop('DEVITO_PLATFORM', None)
    env.pop('DEVITO_LANGUAGE', None)
    r = subprocess.run([sys.executable, os.path.join(HERE, 'refplugin_worker.py'), mode], capture_output=True,
                       text=True, timeout=timeout, env=env, cwd='/tmp')
    return r


def test_reference_operator_marshals_its_own_structs_into_the_c_abi():
    r = _run('ffi', 900)
    assert 'REFPLUGIN-FFI-OK' in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.gpu
def test_reference_examples_run_unchanged_on_the_gpu():
    r = _run('gpu', 1500)
    assert 'REFPLUGIN-GPU-OK' in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
    print(r.stdout[-600:])
