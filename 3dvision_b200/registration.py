"""Host-side mirror of the reference's registration interface for the hot path.

Same names, argument meaning, defaults and error behaviour as
``include/registration.hpp:10-60`` and ``include/gpu_registration.hpp:8-19`` of
stojicnnnn/3DVision, so a caller of ``Registration::ransacRegistration`` /
``Registration::icpRefine`` / ``GPURegistration::icpRefine`` can switch over
(the C++ drop-in shim is ``shim/registration.hpp``; this module is the
Python face used by the tests and bench).  All compute happens in libb3d.so
through the C-ABI of ``include/b3d.h``; nothing here computes on the CPU.

``voxelDownsample``, ``estimateNormals`` and ``computeFPFH`` are
provided too; ``loadReferenceModel`` (file I/O) is out of scope and absent.
"""
from __future__ import annotations

import threading
from dataclasses import dataclass, field

import numpy as np

from . import _capi


@dataclass
class PointCloud:
    """registration.hpp:10-19 — three packed xyz arrays (float32, shape (n,3))."""
    points: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), np.float32))
    normals: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), np.float32))
    colors: np.ndarray = field(default_factory=lambda: np.zeros((0, 3), np.float32))

    def size(self) -> int:
        return int(np.asarray(self.points).reshape(-1, 3).shape[0])

    def empty(self) -> bool:
        return self.size() == 0

    def hasNormals(self) -> bool:
        return np.asarray(self.normals).reshape(-1, 3).shape[0] == self.size()

    def hasColors(self) -> bool:
        return np.asarray(self.colors).reshape(-1, 3).shape[0] == self.size()


@dataclass
class FPFHFeatures:
    """registration.hpp:21-24 — descriptors, float32 (n,33), L1-normalised rows."""
    descriptors: np.ndarray = field(default_factory=lambda: np.zeros((0, 33), np.float32))

    def size(self) -> int:
        return int(np.asarray(self.descriptors).reshape(-1, 33).shape[0])


@dataclass
class RegistrationResult:
    """registration.hpp:26-30 — identity / 0 / 0 by default."""
    transformation: np.ndarray = field(default_factory=lambda: np.eye(4, dtype=np.float32))
    fitness: float = 0.0
    rmse: float = 0.0


_tls = threading.local()


def _context(device: int = 0) -> _capi.Context:
    """thread_local context: the orchestrator calls from a pool of workers
    (src/pipeline.cpp:321-327), each of which gets its own stream + workspace."""
    ctxs = getattr(_tls, "ctxs", None)
    if ctxs is None:
        ctxs = _tls.ctxs = {}
    if device not in ctxs:
        ctxs[device] = _capi.Context(device)
    return ctxs[device]


class Registration:
    """Static interface of registration.hpp:32-60 (hot-path members only)."""

    device = 0

    @staticmethod
    def voxelDownsample(cloud: PointCloud, voxel_size: float) -> PointCloud:
        """registration.hpp:34 / registration.cpp:29-60. Points (and colors) averaged per voxel, emitted in the
        reference's unordered_map iteration order; normals are dropped, as there."""
        pts, col = _context(Registration.device).voxel_downsample(cloud.points, float(voxel_size),
                                                                  cloud.colors if cloud.hasColors() and cloud.size() else None)
        return PointCloud(points=pts, colors=col if col is not None else np.zeros((0, 3), np.float32))

    @staticmethod
    def estimateNormals(cloud: PointCloud, k: int = 30) -> None:
        """registration.hpp:36 / registration.cpp:105-130. In place, like the reference."""
        cloud.normals = _context(Registration.device).estimate_normals(cloud.points, int(k))

    @staticmethod
    def computeFPFH(cloud: PointCloud, radius: float) -> FPFHFeatures:
        """registration.hpp:38 / registration.cpp:133-201."""
        return FPFHFeatures(_context(Registration.device).compute_fpfh(cloud.points, cloud.normals, float(radius)))

    @staticmethod
    def ransacRegistration(source: PointCloud, target: PointCloud,
                           source_features: FPFHFeatures, target_features: FPFHFeatures,
                           voxel_size: float, max_iterations: int = 100000,
                           confidence: float = 0.999) -> RegistrationResult:
        """registration.hpp:40-48 / registration.cpp:204-295."""
        T, fit, rmse, _ = _context(Registration.device).ransac(
            source.points, target.points, source_features.descriptors, target_features.descriptors,
            float(voxel_size), int(max_iterations), float(confidence))
        return RegistrationResult(T, fit, rmse)

    @staticmethod
    def icpRefine(source: PointCloud, target: PointCloud, initial_transform,
                  distance_threshold: float, max_iterations: int = 200,
                  point_to_plane: bool = True) -> RegistrationResult:
        """registration.hpp:50-57 / registration.cpp:297-414."""
        normals = target.normals if target.hasNormals() and target.size() > 0 else None
        T, fit, rmse, _ = _context(Registration.device).icp(
            source.points, target.points, normals, initial_transform,
            float(distance_threshold), int(max_iterations), bool(point_to_plane))
        return RegistrationResult(T, fit, rmse)


class GPURegistration:
    """gpu_registration.hpp:8-19."""

    @staticmethod
    def icpRefine(source: PointCloud, target: PointCloud, initial_transform,
                  distance_threshold: float, max_iterations: int = 200) -> RegistrationResult:
        """gpu_registration.hpp:10-16. Raises RuntimeError when CUDA is unavailable
        (gpu_impl.cpp:258 throws std::runtime_error), so a caller's
        ``except Exception`` fallback keeps working (pipeline.cpp:108-121)."""
        if not GPURegistration.isCudaAvailable():
            raise RuntimeError("CUDA not available")
        return Registration.icpRefine(source, target, initial_transform, distance_threshold, max_iterations, True)

    @staticmethod
    def isCudaAvailable() -> bool:
        """gpu_registration.hpp:18."""
        return _capi.cuda_available()
