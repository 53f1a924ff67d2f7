"""midagma_b200 -- B200-native (sm_100a) implementation of DAGMA's inner
optimisation hot path behind the dagma / midagma Python API.

Host code here mirrors the reference's classes (same names, arguments, defaults,
return values and failure behaviour); all arithmetic runs in hand-written CUDA
kernels in ``lib/libdagma_b200.so`` reached through the C ABI of
``include/dagma_b200.h``.  There is no CPU fallback.
"""
from .linear import DagmaLinear, fit_batch, minimize_batch  # noqa: F401

__all__ = ["DagmaLinear", "fit_batch", "minimize_batch"]
