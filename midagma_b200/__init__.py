"""midagma_b200 -- B200-native (sm_100a) implementation of DAGMA's inner
optimisation hot path behind the dagma / midagma Python API.

Host code here mirrors the reference's classes (same names, arguments, defaults,
return values and failure behaviour); all arithmetic runs in hand-written CUDA
kernels in ``lib/libdagma_b200.so`` reached through the C ABI of
``include/dagma_b200.h``.  There is no CPU fallback.
"""
import os as _os

# Mid-d batches run several problems side by side, one persistent kernel and one stream each (linear._batch_lanes): with
# the driver's default of 8 hardware work queues, streams beyond that share a queue and their long kernels serialise
# (measured: 8 lanes at d = 128, 102 k -> 145 k problem-iterations/s).  Only effective when set before the CUDA context
# exists; an explicit setting of the user wins.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from .linear import DagmaLinear, fit_batch, minimize_batch  # noqa: F401

__all__ = ["DagmaLinear", "fit_batch", "minimize_batch"]
