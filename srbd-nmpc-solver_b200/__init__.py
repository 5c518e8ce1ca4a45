"""srbd-nmpc-solver_b200 — B200-native SRBD-NMPC hot path (linearize -> assemble -> OCP-QP IPM).

The product is the CUDA library `libsrbd_b200.so` behind the C-ABI of include/srbd_b200.h; this
package holds its sources (csrc/), the C++ host facades mirroring the reference's classes (host/)
and a thin ctypes binding used by the tests and bench.py.  Importing the package does not load the
library; the first call does, and fails loudly if it was not built (no CPU fallback).
"""
from . import capi  # noqa: F401
from . import sharding  # noqa: F401
from . import workload  # noqa: F401
from .binding import Context, default_ipm_args, default_model_params  # noqa: F401
