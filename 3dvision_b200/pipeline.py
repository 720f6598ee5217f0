"""Batched multi-object registration — the hot part of the reference's orchestrator.

``Pipeline::run`` (src/pipeline.cpp:311-338) pushes one ``processInstance`` task per detected
object into a pool of ``num_threads`` workers (default 8, pipeline.cpp:16, 321-327); each task
runs ``Registration::ransacRegistration`` and then ``GPURegistration::icpRefine`` (or the CPU
``Registration::icpRefine``) on its own clouds (pipeline.cpp:97-129).  This module mirrors that
unit of work and its pool for SURVEY.md §8(e) "batched multi-object": every worker thread owns
one ``b3d_ctx`` (stream + workspace, ``registration._context``), so instances overlap on one GPU
exactly as the reference's pool would drive them, and ``dist.sharded_batch`` deals instances
round-robin to the ranks of a torch.distributed group (instance i -> rank i mod G, no data-path
collective; one 18-float-per-instance gather of the poses at the end).

Mask/deprojection/downsampling/feature stages of ``processInstance`` (pipeline.cpp:42-95) are
outside the hot path — an ``Instance`` starts where they end.
"""
from __future__ import annotations

import threading
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass

import numpy as np

from .registration import (FPFHFeatures, GPURegistration, PointCloud, Registration, RegistrationResult, _context)


@dataclass
class Instance:
    """Inputs of one ``processInstance`` call at the point the hot path starts (pipeline.cpp:97)."""
    source: PointCloud                  # scene instance cloud, downsampled (source_down)
    target: PointCloud                  # reference model, with normals (model_down)
    source_features: FPFHFeatures
    target_features: FPFHFeatures
    voxel_size: float
    ransac_iterations: int = 100000     # registration.hpp:46
    confidence: float = 0.999           # registration.hpp:47
    icp_distance_factor: float = 0.4    # pipeline_config.hpp:28; threshold = voxel_size * factor (pipeline.cpp:104)
    icp_iterations: int = 200           # registration.hpp:55
    use_point_to_plane: bool = True     # pipeline_config.hpp:31 (CPU entry only; the GPU entry is always plane)
    use_gpu: bool = True                # pipeline.cpp:107 (config.use_gpu)


def process_instance(inst: Instance) -> tuple[RegistrationResult, RegistrationResult]:
    """pipeline.cpp:97-129: coarse RANSAC pose, then ICP refinement. Returns (coarse, refined)."""
    coarse = Registration.ransacRegistration(inst.source, inst.target, inst.source_features, inst.target_features,
                                             inst.voxel_size, inst.ransac_iterations, inst.confidence)
    thr = inst.voxel_size * inst.icp_distance_factor
    if inst.use_gpu and GPURegistration.isCudaAvailable():
        fine = GPURegistration.icpRefine(inst.source, inst.target, coarse.transformation, thr, inst.icp_iterations)
    else:   # same library entry (there is no CPU path here); raises if the device is missing
        fine = Registration.icpRefine(inst.source, inst.target, coarse.transformation, thr, inst.icp_iterations,
                                      inst.use_point_to_plane)
    return coarse, fine


def world_pose(refined_transformation, camera_extrinsics=None):
    """pipeline.cpp:136-137: T_camera_object = refined^-1; T_world_object = extrinsics * T_camera_object.
    Runs in libb3d.so (``b3d_world_poses``): Eigen's SSE 4x4 inverse kernel and packet product order, on the calling
    thread's context like every other stage."""
    return _context().world_poses(np.asarray(refined_transformation, np.float32).reshape(1, 4, 4), camera_extrinsics)[0]


def world_poses(refined_transformations, camera_extrinsics=None):
    """world_pose for a batch of refined transforms in one call: (n,4,4) -> (n,4,4)."""
    return _context().world_poses(refined_transformations, camera_extrinsics)


def filter_duplicates(waypoints, min_distance: float):
    """Pipeline::filterDuplicates (pipeline.cpp:153-180): greedy de-duplication of the per-instance poses by the distance
    between their translations; of two poses closer than min_distance the one nearer the origin is kept, in the slot of
    the first.  Runs in libb3d.so (``b3d_filter_duplicates``)."""
    waypoints = list(waypoints)
    if not waypoints:
        return []
    return _context().filter_duplicates(np.stack([np.asarray(w, np.float32).reshape(4, 4) for w in waypoints]), min_distance)


_pools: dict[int, ThreadPoolExecutor] = {}
_pools_lock = threading.Lock()


def _pool(num_threads: int) -> ThreadPoolExecutor:
    """Workers persist across batches so their thread-local contexts (and device workspaces) do too,
    like the reference's ThreadPool member (include/thread_pool.hpp:16-34)."""
    with _pools_lock:
        if num_threads not in _pools:
            _pools[num_threads] = ThreadPoolExecutor(max_workers=num_threads, thread_name_prefix="b3d-worker")
        return _pools[num_threads]


def register_batch(instances, num_threads: int = 8):
    """Register every instance; results in input order. Exceptions propagate like a failed future
    (pipeline.cpp:330-336 collects ``future.get()`` in order)."""
    instances = list(instances)
    if not instances:
        return []
    if num_threads <= 1:
        return [process_instance(i) for i in instances]
    return list(_pool(num_threads).map(process_instance, instances))
