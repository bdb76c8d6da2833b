"""teamoflow_b200 -- B200-native (sm_100a) matrix-factorization hot path behind TeAMOFlow's
``teamoflow.mf`` plugin surface.  Python host code -> ctypes C ABI (``include/tmf.h``) ->
hand-written CUDA (``teamoflow_b200/csrc``).  No TensorFlow, no CPU fallback."""
# same version string as the reference package it stands in for (src/teamoflow/__init__.py:3)
__version__ = '0.0.2'

from . import mf  # noqa: E402,F401
