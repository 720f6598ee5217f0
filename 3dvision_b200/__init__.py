"""3dvision_b200 — B200-native (sm_100a) registration hot path of stojicnnnn/3DVision.

The package name starts with a digit, so import it with::

    import importlib
    b3d = importlib.import_module("3dvision_b200")

Contents: ``registration`` (mirror of the reference's registration.hpp /
gpu_registration.hpp interface), ``_capi`` (ctypes binding of include/b3d.h),
``dist`` (multi-GPU bootstrap + protocol mirror), ``synthetic`` (workloads),
``csrc/`` (CUDA kernels + C-ABI), ``shim/`` (C++ drop-in header).
"""
from . import _capi, synthetic  # noqa: F401
from ._capi import B3DError, Context, Pool, cuda_available  # noqa: F401
from .registration import (FPFHFeatures, GPURegistration, PointCloud,  # noqa: F401
                           Registration, RegistrationResult)

__all__ = ["B3DError", "Context", "Pool", "cuda_available", "FPFHFeatures", "GPURegistration", "PointCloud",
           "Registration", "RegistrationResult", "synthetic"]
