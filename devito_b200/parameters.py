"""Global configuration, mirroring the keys the reference exposes
(devito/parameters.py:21-139, keys defined devito/__init__.py:70-179; env vars
``DEVITO_*`` devito/parameters.py:142-159).  Only the keys that influence this backend are
interpreted; the others are accepted and stored so that user scripts keep working."""
import os
from functools import wraps

__all__ = ['configuration', 'switchconfig']

_defaults = {
    'platform': 'b200', 'language': 'cuda', 'compiler': 'nvcc', 'mpi': False,
    'topology': None, 'log-level': 'INFO', 'opt': 'advanced', 'opt-options': {},
    'profiling': 'basic', 'autotuning': 'off', 'develop-mode': False, 'safe-math': False,
    'deviceid': -1, 'ignore-unknowns': False, 'first-touch': False, 'jit-backdoor': False,
    'autopadding': False, 'errctl': 'basic',
}
_env = {'DEVITO_PLATFORM': 'platform', 'DEVITO_LANGUAGE': 'language', 'DEVITO_ARCH': 'compiler',
        'DEVITO_MPI': 'mpi', 'DEVITO_TOPOLOGY': 'topology', 'DEVITO_LOGGING': 'log-level',
        'DEVITO_OPT': 'opt', 'DEVITO_PROFILING': 'profiling', 'DEVITO_DEVICEID': 'deviceid',
        'DEVITO_IGNORE_UNKNOWN_PARAMS': 'ignore-unknowns', 'DEVITO_ERRCTL': 'errctl'}


class Parameters(dict):
    def __init__(self):
        super().__init__(_defaults)
        for var, key in _env.items():
            if var in os.environ:
                val = os.environ[var]
                if key in ('mpi', 'ignore-unknowns'):
                    val = val not in ('0', '', 'False', 'false')
                elif key == 'deviceid':
                    val = int(val)
                self[key] = val


configuration = Parameters()


class switchconfig:
    """Decorator / context manager to temporarily change configuration values
    (devito/parameters.py:262-285)."""

    def __init__(self, condition=True, **params):
        self.params = {k.replace('_', '-'): v for k, v in params.items()} if condition else {}
        self.previous = {}

    def __enter__(self):
        self.previous = {k: configuration.get(k) for k in self.params}
        configuration.update(self.params)
        return self

    def __exit__(self, *exc):
        configuration.update(self.previous)
        return False

    def __call__(self, func):
        @wraps(func)
        def wrapper(*args, **kwargs):
            with self:
                return func(*args, **kwargs)
        return wrapper
