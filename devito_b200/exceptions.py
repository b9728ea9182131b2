"""Exception types with the reference's names (devito/exceptions.py)."""


class DevitoError(Exception):
    pass


class InvalidArgument(DevitoError, ValueError):
    """Raised by the pre-FFI argument checks (devito/types/dense.py:934-957,
    devito/operator/operator.py:588-592)."""


class InvalidOperator(DevitoError):
    pass


class ExecutionError(DevitoError, RuntimeError):
    """Non-zero return code of the C-ABI call (devito/operator/operator.py:734-772)."""


class BackendUnavailable(DevitoError, RuntimeError):
    """The CUDA extension / a GPU is missing. The hot path never falls back to the CPU."""
