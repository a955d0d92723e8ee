"""B200-native bundle-adjustment hot path of ajingu/RealSenseCalibration.

The product is ``csrc/`` (CUDA kernels + the ``ba_cuda_*`` C ABI, ``include/ba_cuda.h``)
and ``host/`` (the reference's BALProblem / BAManager / ReprojectionCheck classes in C++
on top of that ABI).  The Python modules here are thin ctypes plumbing for tests and
``bench.py``: they never compute on the CPU and raise if the CUDA library is missing.
"""
from .abi import Options, Iteration, Summary  # noqa: F401

__all__ = ["Options", "Iteration", "Summary"]
