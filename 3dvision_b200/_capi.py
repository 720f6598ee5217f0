"""ctypes binding of include/b3d.h (libb3d.so, sm_100a CUDA kernels).

There is no CPU implementation behind this module: if the library is missing it
raises, and without a B200-class device every compute call raises ``B3DError``
(status ``B3D_ERR_NO_DEVICE``), which mirrors the ``std::runtime_error`` the
reference's GPU entry point throws (src/gpu_impl.cpp:258).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb3d.so")

B3D_OK = 0
B3D_ERR_NO_DEVICE = -1
B3D_ERR_CUDA = -2
B3D_ERR_INVALID = -3
B3D_ERR_ALLOC = -4
B3D_ERR_STATE = -5
NO_MATCH = 0xFFFFFFFF

# every symbol include/b3d.h declares (tests check the library exports exactly these)
SYMBOLS = [
    "b3d_device_count", "b3d_pool_create", "b3d_pool_register", "b3d_pool_destroy", "b3d_comm_unique_id", "b3d_comm_init", "b3d_comm_attach", "b3d_comm_destroy",
    "b3d_ransac_sharded", "b3d_ransac_sharded_resident", "b3d_register_scene_sharded",
    "b3d_cuda_available", "b3d_ctx_create", "b3d_ctx_destroy", "b3d_ctx_set_stream", "b3d_strerror", "b3d_last_error",
    "b3d_ransac", "b3d_icp",
    "b3d_set_clouds", "b3d_set_features", "b3d_set_match_mode", "b3d_match_features", "b3d_get_correspondences", "b3d_get_correspondences_dev", "b3d_set_correspondences",
    "b3d_correspondences_devptr", "b3d_set_score_mode", "b3d_ransac_prepare", "b3d_ransac_score", "b3d_ransac_reduce", "b3d_ransac_reduce3", "b3d_ransac_finish", "b3d_set_finish_mode",
    "b3d_ransac_counts", "b3d_ransac_hypotheses", "b3d_set_icp_mode", "b3d_icp_run", "b3d_icp_nearest",
    "b3d_kernel_launches", "b3d_stage_ms", "b3d_measure_fp32_rate", "b3d_score_recounts", "b3d_icp_exact_sum_stats", "b3d_sequential_sum", "b3d_icp_exact_sum_dump", "b3d_euler_rotations",
    "b3d_prepare_model", "b3d_register_scene", "b3d_register_scene_device", "b3d_depth_to_cloud", "b3d_register_depth", "b3d_world_poses", "b3d_filter_duplicates", "b3d_voxel_downsample", "b3d_estimate_normals", "b3d_compute_fpfh",
]


class SceneResult(C.Structure):
    """b3d_scene_result (include/b3d.h)."""
    _fields_ = [("coarse_T", C.c_float * 16), ("coarse_fitness", C.c_float), ("coarse_rmse", C.c_float),
                ("coarse_best_iteration", C.c_int32), ("T", C.c_float * 16), ("fitness", C.c_float), ("rmse", C.c_float),
                ("icp_iterations", C.c_int32), ("n_source_points", C.c_uint32)]


class Instance(C.Structure):
    """b3d_instance (include/b3d.h)."""
    _fields_ = [("src_xyz", C.c_void_p), ("n_src", C.c_size_t), ("tgt_xyz", C.c_void_p), ("tgt_normals", C.c_void_p), ("n_tgt", C.c_size_t),
                ("src_desc", C.c_void_p), ("tgt_desc", C.c_void_p), ("voxel_size", C.c_float), ("ransac_max_iterations", C.c_int),
                ("ransac_confidence", C.c_float), ("icp_distance_threshold", C.c_float), ("icp_max_iterations", C.c_int), ("point_to_plane", C.c_int)]


class InstanceResult(C.Structure):
    """b3d_instance_result (include/b3d.h)."""
    _fields_ = [("coarse_T", C.c_float * 16), ("coarse_fitness", C.c_float), ("coarse_rmse", C.c_float), ("T", C.c_float * 16),
                ("fitness", C.c_float), ("rmse", C.c_float), ("icp_iterations", C.c_int32), ("status", C.c_int32)]


class B3DError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"b3d status {status}: {msg}")
        self.status = status


_lib = None


def lib():
    """Load libb3d.so. Fails loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing — build it with __graft_entry__.build() "
                              "(nvcc, sm_100a). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        _declare(L)
        _lib = L
    return _lib


_vp = C.c_void_p
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int32)
_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int32)


def _declare(L):
    L.b3d_cuda_available.restype = C.c_int
    L.b3d_ctx_create.argtypes = [C.c_int, C.POINTER(_vp)]
    L.b3d_ctx_destroy.argtypes = [_vp]
    L.b3d_ctx_destroy.restype = None
    L.b3d_ctx_set_stream.argtypes = [_vp, _vp]
    L.b3d_strerror.argtypes = [C.c_int]
    L.b3d_strerror.restype = C.c_char_p
    L.b3d_last_error.argtypes = [_vp]
    L.b3d_last_error.restype = C.c_char_p
    L.b3d_ransac.argtypes = [_vp, _vp, C.c_size_t, _vp, C.c_size_t, _vp, _vp, C.c_float, C.c_int, C.c_float,
                             _f32p, _f32p, _f32p, _i32p]
    L.b3d_icp.argtypes = [_vp, _vp, C.c_size_t, _vp, _vp, C.c_size_t, _f32p, C.c_float, C.c_int, C.c_int,
                          _f32p, _f32p, _f32p, _i32p]
    L.b3d_set_clouds.argtypes = [_vp, _vp, C.c_size_t, _vp, _vp, C.c_size_t, C.c_int]
    L.b3d_set_features.argtypes = [_vp, _vp, _vp, C.c_int]
    L.b3d_match_features.argtypes = [_vp, C.c_size_t, C.c_size_t]
    L.b3d_set_match_mode.argtypes = [_vp, C.c_int]
    L.b3d_set_score_mode.argtypes = [_vp, C.c_int]
    L.b3d_set_icp_mode.argtypes = [_vp, C.c_int]
    L.b3d_get_correspondences.argtypes = [_vp, _vp]
    L.b3d_set_correspondences.argtypes = [_vp, _vp, C.c_int]
    L.b3d_get_correspondences_dev.argtypes = [_vp, _vp]
    L.b3d_correspondences_devptr.argtypes = [_vp, C.POINTER(_vp)]
    L.b3d_ransac_prepare.argtypes = [_vp, C.c_float, C.c_int, C.c_float]
    L.b3d_ransac_score.argtypes = [_vp, C.c_int, C.c_int]
    L.b3d_ransac_reduce.argtypes = [_vp, C.c_int, C.c_int, _vp, _vp]
    L.b3d_ransac_reduce3.argtypes = [_vp, C.c_int, C.c_int, _vp]
    L.b3d_set_finish_mode.argtypes = [_vp, C.c_int]
    L.b3d_ransac_finish.argtypes = [_vp, _vp, _f32p, _f32p, _f32p, _i32p]
    L.b3d_ransac_counts.argtypes = [_vp, C.c_int, C.c_int, _vp]
    L.b3d_ransac_hypotheses.argtypes = [_vp, C.c_int, C.c_int, _vp]
    L.b3d_icp_run.argtypes = [_vp, _f32p, C.c_float, C.c_int, C.c_int, C.c_int, _f32p, _f32p, _f32p, _i32p]
    L.b3d_icp_nearest.argtypes = [_vp, _f32p, C.c_float, _vp, _vp]
    L.b3d_kernel_launches.argtypes = [_vp]
    L.b3d_kernel_launches.restype = C.c_uint64
    L.b3d_stage_ms.argtypes = [_vp, C.c_int]
    L.b3d_stage_ms.restype = C.c_float
    L.b3d_measure_fp32_rate.argtypes = [_vp, C.POINTER(C.c_double)]
    L.b3d_score_recounts.argtypes = [_vp, C.POINTER(C.c_uint64)]
    L.b3d_icp_exact_sum_stats.argtypes = [_vp, C.POINTER(C.c_uint32)]
    L.b3d_icp_exact_sum_dump.argtypes = [_vp, _fp, C.c_size_t, C.POINTER(C.c_size_t), _fp]
    L.b3d_euler_rotations.argtypes = [_vp, _fp, C.c_size_t, _fp]
    L.b3d_sequential_sum.argtypes = [_vp, _fp, C.c_size_t, C.POINTER(C.c_float), C.POINTER(C.c_uint32)]
    L.b3d_voxel_downsample.argtypes = [_vp, _vp, C.c_size_t, _vp, C.c_float, _vp, _vp, C.c_size_t, C.POINTER(C.c_size_t)]
    L.b3d_depth_to_cloud.argtypes = [_vp, _vp, C.c_int, C.c_int, _vp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                     C.c_float, _vp, _vp, _vp, C.c_size_t, C.POINTER(C.c_size_t)]
    L.b3d_register_depth.argtypes = [_vp, _vp, C.c_int, C.c_int, _vp, C.c_int, C.c_int] + [C.c_float] * 6 + [C.c_float, C.c_int, C.c_float, C.c_int,
                                     C.c_float, C.c_float, C.c_int, C.c_int, C.POINTER(SceneResult)]
    L.b3d_world_poses.argtypes = [_vp, _vp, C.c_size_t, _vp, _vp]
    L.b3d_pool_create.argtypes = [C.c_int, C.POINTER(C.c_int), C.c_int, C.POINTER(_vp)]
    L.b3d_pool_register.argtypes = [_vp, C.POINTER(Instance), C.c_size_t, C.POINTER(InstanceResult)]
    L.b3d_pool_destroy.argtypes = [_vp]
    L.b3d_pool_destroy.restype = None
    L.b3d_comm_unique_id.argtypes = [_vp]
    L.b3d_comm_init.argtypes = [_vp, _vp, C.c_int, C.c_int]
    L.b3d_comm_attach.argtypes = [_vp, _vp, C.c_int, C.c_int]
    L.b3d_comm_destroy.argtypes = [_vp]
    L.b3d_ransac_sharded.argtypes = [_vp, _vp, C.c_size_t, _vp, C.c_size_t, _vp, _vp, C.c_float, C.c_int, C.c_float, _vp, _fp, _fp, _ip]
    L.b3d_ransac_sharded_resident.argtypes = [_vp, C.c_float, C.c_int, C.c_float, C.c_int, _vp, _fp, _fp, _ip]
    L.b3d_register_scene_sharded.argtypes = [_vp, _vp, C.c_size_t, C.c_float, C.c_int, C.c_float, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int,
                                             C.POINTER(SceneResult)]
    L.b3d_filter_duplicates.argtypes = [_vp, _vp, C.c_size_t, C.c_float, _vp, C.POINTER(C.c_size_t)]
    L.b3d_register_scene_device.argtypes = [_vp, _vp, C.c_size_t, C.c_float, C.c_int, C.c_float, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int,
                                            C.POINTER(SceneResult)]
    L.b3d_prepare_model.argtypes = [_vp, _vp, C.c_size_t, C.c_float, C.c_int, C.c_float, C.POINTER(C.c_size_t)]
    L.b3d_register_scene.argtypes = [_vp, _vp, C.c_size_t, C.c_float, C.c_int, C.c_float, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int,
                                     C.POINTER(SceneResult)]
    L.b3d_estimate_normals.argtypes = [_vp, _vp, C.c_size_t, C.c_int, _vp]
    L.b3d_compute_fpfh.argtypes = [_vp, _vp, _vp, C.c_size_t, C.c_float, _vp]


def cuda_available() -> bool:
    return bool(lib().b3d_cuda_available())


def _ptr(a):
    """Host numpy array / raw integer address / None -> c_void_p."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return C.c_void_p(a.ctypes.data)


def _T_colmajor(T) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(T, np.float32).reshape(4, 4).T).reshape(16)


def _T_from_colmajor(buf) -> np.ndarray:
    return np.asarray(buf, np.float32).reshape(4, 4).T.copy()


class Pool:
    """b3d_pool: the orchestrator's worker pool (pipeline.cpp:321-327) in C — n_workers host threads, one context each, dealt over
    `devices`; register() runs ransacRegistration + icpRefine for every instance and returns [(coarse, refined)] in input order."""

    def __init__(self, n_workers: int = 8, devices=(0,)):
        self._L = lib()
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        h = _vp()
        rc = self._L.b3d_pool_create(int(n_workers), devs, len(devices), C.byref(h))
        if rc != B3D_OK:
            raise B3DError(rc, self._L.b3d_strerror(rc).decode())
        self._h = h

    def register(self, instances):
        """instances: iterable of dicts / objects with source, target, target_normals, source_desc, target_desc, voxel_size and
        optional ransac_iterations, confidence, icp_threshold, icp_iterations, point_to_plane."""
        instances = list(instances)
        n = len(instances)
        if n == 0:
            return []
        keep, items = [], (Instance * n)()
        for it, inst in zip(items, instances):
            g = (lambda k, d=None: inst.get(k, d)) if isinstance(inst, dict) else (lambda k, d=None: getattr(inst, k, d))
            src = _as_f32(g("source"), 3); tgt = _as_f32(g("target"), 3)
            nrm = _as_f32(g("target_normals"), 3) if g("target_normals") is not None else None
            sd = _as_f32(g("source_desc"), 33); td = _as_f32(g("target_desc"), 33)
            keep.append((src, tgt, nrm, sd, td))
            voxel = float(g("voxel_size"))
            thr = g("icp_threshold")
            it.src_xyz = src.ctypes.data; it.n_src = src.shape[0]; it.tgt_xyz = tgt.ctypes.data; it.n_tgt = tgt.shape[0]
            it.tgt_normals = nrm.ctypes.data if nrm is not None and nrm.shape[0] == tgt.shape[0] else None
            it.src_desc = sd.ctypes.data; it.tgt_desc = td.ctypes.data
            it.voxel_size = voxel; it.ransac_max_iterations = int(g("ransac_iterations", 100000)); it.ransac_confidence = float(g("confidence", 0.999))
            it.icp_distance_threshold = float(voxel * 0.4 if thr is None else thr); it.icp_max_iterations = int(g("icp_iterations", 200))
            it.point_to_plane = int(bool(g("point_to_plane", True)))
        res = (InstanceResult * n)()
        rc = self._L.b3d_pool_register(self._h, items, n, res)
        if rc != B3D_OK:
            raise B3DError(rc, self._L.b3d_strerror(rc).decode())
        return [((_T_from_colmajor(np.array(r.coarse_T, np.float32)), r.coarse_fitness, r.coarse_rmse),
                 (_T_from_colmajor(np.array(r.T, np.float32)), r.fitness, r.rmse, r.icp_iterations)) for r in res]

    def close(self):
        if self._h:
            self._L.b3d_pool_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class Context:
    """One b3d_ctx: a stream plus persistent device workspace. Not thread-safe; use one per thread."""

    def __init__(self, device: int = 0):
        self._h = _vp()
        self._L = lib()
        rc = self._L.b3d_ctx_create(device, C.byref(self._h))
        if rc != B3D_OK:
            raise B3DError(rc, self._L.b3d_strerror(rc).decode())
        self.device = device

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.b3d_ctx_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc):
        if rc != B3D_OK:
            detail = self._L.b3d_last_error(self._h).decode() or self._L.b3d_strerror(rc).decode()
            raise B3DError(rc, detail)

    # ---- plumbing ----
    def set_stream(self, cuda_stream: int | None):
        """Run on an external cudaStream_t. ``0`` (torch's default stream handle) selects the legacy default
        stream (cudaStreamLegacy == 0x1); ``None`` restores the context's own non-blocking stream."""
        if cuda_stream is None:
            arg = None
        else:
            arg = _vp(cuda_stream if cuda_stream != 0 else 1)
        self._check(self._L.b3d_ctx_set_stream(self._h, arg))

    @property
    def kernel_launches(self) -> int:
        return int(self._L.b3d_kernel_launches(self._h))

    def stage_ms(self, stage: int) -> float:
        return float(self._L.b3d_stage_ms(self._h, stage))

    def score_recounts(self) -> int:
        v = C.c_uint64()
        self._check(self._L.b3d_score_recounts(self._h, C.byref(v)))
        return int(v.value)

    def icp_exact_sum_stats(self) -> np.ndarray:
        """(32, 4) uint32: per running sum of the last ICP call: walk rounds, blocks added term by term, SM cycles / 16."""
        out = np.zeros(128, np.uint32)
        self._check(self._L.b3d_icp_exact_sum_stats(self._h, out.ctypes.data_as(C.POINTER(C.c_uint32))))
        return out.reshape(32, 4)

    def icp_exact_sum_dump(self, n_sums: int = 28):
        """(terms (n_sums, stride) float32, sums (32,) float32) of the last iteration of the last reference-order ICP call."""
        stride = C.c_size_t(); sums = np.zeros(32, np.float32)
        self._check(self._L.b3d_icp_exact_sum_dump(self._h, None, 0, C.byref(stride), sums.ctypes.data_as(_fp)))
        terms = np.zeros((n_sums, stride.value), np.float32)
        self._check(self._L.b3d_icp_exact_sum_dump(self._h, terms.ctypes.data_as(_fp), terms.size, C.byref(stride), sums.ctypes.data_as(_fp)))
        return terms, sums

    def euler_rotations(self, angles_xyz) -> np.ndarray:
        """(n, 3) angles -> (n, 3, 3) Rx(a) Ry(b) Rz(g) as the ICP update forms it (registration.cpp:369-371), on the device."""
        a = np.ascontiguousarray(angles_xyz, np.float32).reshape(-1, 3)
        out = np.empty((a.shape[0], 3, 3), np.float32)
        self._check(self._L.b3d_euler_rotations(self._h, a.ctypes.data_as(_fp) if a.size else None, a.shape[0],
                                                out.ctypes.data_as(_fp) if a.size else None))
        return out

    def sequential_sum(self, terms: np.ndarray):
        """fp32 `s = 0; for x in terms: s += x` through the device's exact-sum passes; returns (sum, stats[3])."""
        t = np.ascontiguousarray(terms, np.float32).reshape(-1)
        out = C.c_float()
        stats = np.zeros(3, np.uint32)
        self._check(self._L.b3d_sequential_sum(self._h, t.ctypes.data_as(_fp) if t.size else None, t.size, C.byref(out),
                                               stats.ctypes.data_as(C.POINTER(C.c_uint32))))
        return np.float32(out.value), stats

    def measure_fp32_rate(self) -> float:
        """Sustained un-fused FMUL+FADD lane-ops/s on this device (scoring-kernel roofline)."""
        v = C.c_double()
        self._check(self._L.b3d_measure_fp32_rate(self._h, C.byref(v)))
        return v.value

    # ---- whole path ----
    def ransac(self, src, tgt, src_desc, tgt_desc, voxel_size, max_iterations=100000, confidence=0.999):
        src = _as_f32(src, 3); tgt = _as_f32(tgt, 3); sd = _as_f32(src_desc, 33); td = _as_f32(tgt_desc, 33)
        if sd.shape[0] != src.shape[0] or td.shape[0] != tgt.shape[0]:
            raise ValueError("descriptor rows must match cloud sizes")
        T = np.empty(16, np.float32); fit = C.c_float(); rm = C.c_float(); best = C.c_int32()
        self._n_src, self._n_tgt, self._H = src.shape[0], tgt.shape[0], max_iterations
        self._check(self._L.b3d_ransac(self._h, _ptr(src), src.shape[0], _ptr(tgt), tgt.shape[0], _ptr(sd), _ptr(td),
                                       voxel_size, max_iterations, confidence,
                                       T.ctypes.data_as(_f32p), C.byref(fit), C.byref(rm), C.byref(best)))
        return _T_from_colmajor(T), fit.value, rm.value, best.value

    # ---- multi-GPU (include/b3d.h "multi-GPU"): this context as one rank of an NCCL group -------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        rc = lib().b3d_comm_unique_id(buf)
        if rc != B3D_OK:
            raise B3DError(rc, lib().b3d_strerror(rc).decode())
        return buf.raw

    def comm_init(self, unique_id: bytes | None, rank: int, world: int):
        """ncclCommInitRank on this context's device (collective); world == 1 needs no id."""
        self._check(self._L.b3d_comm_init(self._h, C.c_char_p(unique_id) if unique_id is not None else None, int(rank), int(world)))
        self.rank, self.world = int(rank), int(world)

    def comm_destroy(self):
        self._check(self._L.b3d_comm_destroy(self._h))
        self.rank, self.world = 0, 1

    def ransac_sharded(self, src, tgt, src_desc, tgt_desc, voxel_size, max_iterations=100000, confidence=0.999):
        """b3d_ransac_sharded: collective; host buffers in, each rank uploads only its rows of the source descriptors."""
        src = _as_f32(src, 3); tgt = _as_f32(tgt, 3); sd = _as_f32(src_desc, 33); td = _as_f32(tgt_desc, 33)
        T = np.empty(16, np.float32); fit = C.c_float(); rm = C.c_float(); best = C.c_int32()
        self._n_src, self._n_tgt, self._H = src.shape[0], tgt.shape[0], max_iterations
        self._check(self._L.b3d_ransac_sharded(self._h, _ptr(src), src.shape[0], _ptr(tgt), tgt.shape[0], _ptr(sd), _ptr(td), voxel_size,
                                               max_iterations, confidence, _ptr(T), C.byref(fit), C.byref(rm), C.byref(best)))
        return _T_from_colmajor(T), fit.value, rm.value, best.value

    def ransac_sharded_resident(self, voxel_size, max_iterations=100000, confidence=0.999, match=True):
        T = np.empty(16, np.float32); fit = C.c_float(); rm = C.c_float(); best = C.c_int32()
        self._H = max_iterations
        self._check(self._L.b3d_ransac_sharded_resident(self._h, voxel_size, max_iterations, confidence, int(bool(match)), _ptr(T),
                                                        C.byref(fit), C.byref(rm), C.byref(best)))
        return _T_from_colmajor(T), fit.value, rm.value, best.value

    def register_scene_sharded(self, scene_xyz, voxel_size, normals_k=30, fpfh_radius=None, ransac_max_iterations=100000, confidence=0.999,
                               icp_threshold=None, icp_max_iterations=200, point_to_plane=True):
        xyz = _as_f32(scene_xyz, 3)
        radius = voxel_size * 5.0 if fpfh_radius is None else fpfh_radius
        thr = voxel_size * 0.4 if icp_threshold is None else icp_threshold
        r = SceneResult()
        self._check(self._L.b3d_register_scene_sharded(self._h, _ptr(xyz), xyz.shape[0], voxel_size, int(normals_k), radius,
                                                       int(ransac_max_iterations), confidence, thr, int(icp_max_iterations),
                                                       int(bool(point_to_plane)), C.byref(r)))
        self._n_src = r.n_source_points; self._H = int(ransac_max_iterations)
        return {"coarse": (_T_from_colmajor(np.array(r.coarse_T, np.float32)), r.coarse_fitness, r.coarse_rmse, r.coarse_best_iteration),
                "refined": (_T_from_colmajor(np.array(r.T, np.float32)), r.fitness, r.rmse, r.icp_iterations),
                "n_source_points": int(r.n_source_points)}

    def icp(self, src, tgt, tgt_normals, T0, distance_threshold, max_iterations=200, point_to_plane=True):
        src = _as_f32(src, 3); tgt = _as_f32(tgt, 3)
        nrm = _as_f32(tgt_normals, 3) if tgt_normals is not None else None
        if nrm is not None and nrm.shape[0] != tgt.shape[0]:
            nrm = None                      # PointCloud::hasNormals() false, registration.hpp:17
        T0c = _T_colmajor(T0)
        T = np.empty(16, np.float32); fit = C.c_float(); rm = C.c_float(); it = C.c_int32()
        self._n_src, self._n_tgt = src.shape[0], tgt.shape[0]
        self._check(self._L.b3d_icp(self._h, _ptr(src), src.shape[0], _ptr(tgt), _ptr(nrm), tgt.shape[0],
                                    T0c.ctypes.data_as(_f32p), distance_threshold, max_iterations, int(bool(point_to_plane)),
                                    T.ctypes.data_as(_f32p), C.byref(fit), C.byref(rm), C.byref(it)))
        return _T_from_colmajor(T), fit.value, rm.value, it.value

    # ---- staged ----
    def set_clouds(self, src, tgt, tgt_normals=None):
        src = _as_f32(src, 3); tgt = _as_f32(tgt, 3)
        nrm = _as_f32(tgt_normals, 3) if tgt_normals is not None else None
        self._keep = (src, tgt, nrm)
        self._n_src, self._n_tgt = src.shape[0], tgt.shape[0]
        self._check(self._L.b3d_set_clouds(self._h, _ptr(src), src.shape[0], _ptr(tgt), _ptr(nrm), tgt.shape[0], 0))

    def set_clouds_device(self, src_ptr: int, n_src: int, tgt_ptr: int, nrm_ptr: int | None, n_tgt: int):
        self._n_src, self._n_tgt = n_src, n_tgt
        self._check(self._L.b3d_set_clouds(self._h, _vp(src_ptr), n_src, _vp(tgt_ptr), _vp(nrm_ptr) if nrm_ptr else None, n_tgt, 1))

    def set_features(self, src_desc, tgt_desc):
        sd = _as_f32(src_desc, 33); td = _as_f32(tgt_desc, 33)
        self._keepf = (sd, td)
        self._check(self._L.b3d_set_features(self._h, _ptr(sd), _ptr(td), 0))

    def set_features_device(self, sd_ptr: int, td_ptr: int):
        self._check(self._L.b3d_set_features(self._h, _vp(sd_ptr), _vp(td_ptr), 1))

    def set_match_mode(self, mode: int):
        """0 auto, 1 exact CUDA-core kernel, 2 tcgen05 screen + exact re-score (identical results)."""
        self._check(self._L.b3d_set_match_mode(self._h, mode))

    def match_features(self, row0=0, row1=None):
        self._check(self._L.b3d_match_features(self._h, row0, self._n_src if row1 is None else row1))

    def get_correspondences(self) -> np.ndarray:
        out = np.empty(self._n_src, np.uint32)
        self._check(self._L.b3d_get_correspondences(self._h, _ptr(out)))
        return out

    def set_correspondences(self, corr):
        corr = np.ascontiguousarray(corr, np.uint32)
        if corr.shape[0] != self._n_src:
            raise ValueError("correspondences must have one entry per source point")
        self._check(self._L.b3d_set_correspondences(self._h, _ptr(corr), 0))

    def get_correspondences_device(self, dst_devptr: int):
        """Stream-ordered D2D copy of correspondences[n_src] (uint32) into caller-owned device memory."""
        self._check(self._L.b3d_get_correspondences_dev(self._h, _vp(dst_devptr)))

    def set_correspondences_device(self, src_devptr: int):
        self._check(self._L.b3d_set_correspondences(self._h, _vp(src_devptr), 1))

    def mark_correspondences_set(self):
        """After writing into correspondences_devptr() directly (all-gather between ranks)."""
        self._check(self._L.b3d_set_correspondences(self._h, None, 1))

    def correspondences_devptr(self) -> int:
        p = _vp()
        self._check(self._L.b3d_correspondences_devptr(self._h, C.byref(p)))
        return int(p.value)

    def set_score_mode(self, mode: int):
        """0 packed FMA screen + exact band re-count (default), 1 un-fused arithmetic everywhere, 2 scalar FMA screen
        (identical counts in all three); 3 bail-out: identical winner/result, hopeless hypotheses dropped part-way."""
        self._check(self._L.b3d_set_score_mode(self._h, mode))

    def ransac_prepare(self, voxel_size, max_iterations, confidence):
        self._H = max_iterations
        self._check(self._L.b3d_ransac_prepare(self._h, voxel_size, max_iterations, confidence))

    def ransac_score(self, h0=0, h1=None):
        self._check(self._L.b3d_ransac_score(self._h, h0, self._H if h1 is None else h1))

    def ransac_reduce(self, h0, h1, keys_devptr: int, limit_devptr: int | None = None):
        self._check(self._L.b3d_ransac_reduce(self._h, h0, h1, _vp(limit_devptr) if limit_devptr else None, _vp(keys_devptr)))

    def set_finish_mode(self, mode: int):
        """RANSAC rmse sum: 0 parallel exact (default), 1 one-chain kernel; same bits."""
        self._check(self._L.b3d_set_finish_mode(self._h, int(mode)))

    def ransac_reduce3(self, h0, h1, keys3_devptr: int):
        self._check(self._L.b3d_ransac_reduce3(self._h, h0, h1, _vp(keys3_devptr)))

    def ransac_finish(self, keys_devptr: int):
        T = np.empty(16, np.float32); fit = C.c_float(); rm = C.c_float(); best = C.c_int32()
        self._check(self._L.b3d_ransac_finish(self._h, _vp(keys_devptr), T.ctypes.data_as(_f32p), C.byref(fit), C.byref(rm), C.byref(best)))
        return _T_from_colmajor(T), fit.value, rm.value, best.value

    def ransac_counts(self, h0=0, h1=None) -> np.ndarray:
        h1 = self._H if h1 is None else h1
        out = np.empty(h1 - h0, np.int32)
        self._check(self._L.b3d_ransac_counts(self._h, h0, h1, _ptr(out)))
        return out

    def ransac_hypotheses(self, h0=0, h1=None) -> np.ndarray:
        h1 = self._H if h1 is None else h1
        out = np.zeros((h1 - h0, 12), np.float32)
        self._check(self._L.b3d_ransac_hypotheses(self._h, h0, h1, _ptr(out)))
        return out

    def set_icp_mode(self, mode: int):
        """0 (default): sums in the reference's order, exact and parallel; 1: fp64 tree sums (fast, opt-in);
        3: reference order through one dependent add chain (cross-check)."""
        self._check(self._L.b3d_set_icp_mode(self._h, mode))

    def icp_run(self, T0, distance_threshold, max_iterations=200, point_to_plane=True, stop_on_convergence=True):
        T0c = _T_colmajor(T0)
        T = np.empty(16, np.float32); fit = C.c_float(); rm = C.c_float(); it = C.c_int32()
        self._check(self._L.b3d_icp_run(self._h, T0c.ctypes.data_as(_f32p), distance_threshold, max_iterations,
                                        int(bool(point_to_plane)), int(bool(stop_on_convergence)),
                                        T.ctypes.data_as(_f32p), C.byref(fit), C.byref(rm), C.byref(it)))
        return _T_from_colmajor(T), fit.value, rm.value, it.value

    # ---- whole registration, resident (pipeline.cpp:86-129 as one call)
    def prepare_model(self, model_xyz, voxel_size, normals_k=30, fpfh_radius=None) -> int:
        xyz = _as_f32(model_xyz, 3)
        m = C.c_size_t()
        radius = voxel_size * 5.0 if fpfh_radius is None else fpfh_radius       # pipeline.cpp:95
        self._check(self._L.b3d_prepare_model(self._h, _ptr(xyz), xyz.shape[0], voxel_size, int(normals_k), radius, C.byref(m)))
        self._n_tgt = m.value
        return m.value

    def register_scene(self, scene_xyz, voxel_size, normals_k=30, fpfh_radius=None, ransac_max_iterations=100000, confidence=0.999,
                       icp_threshold=None, icp_max_iterations=200, point_to_plane=True):
        """-> dict(coarse=(T, fitness, rmse, best_iteration), refined=(T, fitness, rmse, iterations), n_source_points)."""
        xyz = _as_f32(scene_xyz, 3)
        radius = voxel_size * 5.0 if fpfh_radius is None else fpfh_radius
        thr = voxel_size * 0.4 if icp_threshold is None else icp_threshold      # pipeline_config.hpp:28
        r = SceneResult()
        self._check(self._L.b3d_register_scene(self._h, _ptr(xyz), xyz.shape[0], voxel_size, int(normals_k), radius, int(ransac_max_iterations),
                                               confidence, thr, int(icp_max_iterations), int(bool(point_to_plane)), C.byref(r)))
        self._n_src = r.n_source_points; self._H = int(ransac_max_iterations)
        return {"coarse": (_T_from_colmajor(np.array(r.coarse_T, np.float32)), r.coarse_fitness, r.coarse_rmse, r.coarse_best_iteration),
                "refined": (_T_from_colmajor(np.array(r.T, np.float32)), r.fitness, r.rmse, r.icp_iterations),
                "n_source_points": int(r.n_source_points)}

    def register_scene_device(self, scene_devptr: int, n: int, voxel_size, normals_k=30, fpfh_radius=None, ransac_max_iterations=100000,
                              confidence=0.999, icp_threshold=None, icp_max_iterations=200, point_to_plane=True):
        """register_scene for packed xyz floats already resident on this context's device (raw device address)."""
        radius = voxel_size * 5.0 if fpfh_radius is None else fpfh_radius
        thr = voxel_size * 0.4 if icp_threshold is None else icp_threshold
        r = SceneResult()
        self._check(self._L.b3d_register_scene_device(self._h, _ptr(int(scene_devptr)), int(n), voxel_size, int(normals_k), radius,
                                                      int(ransac_max_iterations), confidence, thr, int(icp_max_iterations),
                                                      int(bool(point_to_plane)), C.byref(r)))
        self._n_src = r.n_source_points; self._H = int(ransac_max_iterations)
        return {"coarse": (_T_from_colmajor(np.array(r.coarse_T, np.float32)), r.coarse_fitness, r.coarse_rmse, r.coarse_best_iteration),
                "refined": (_T_from_colmajor(np.array(r.T, np.float32)), r.fitness, r.rmse, r.icp_iterations),
                "n_source_points": int(r.n_source_points)}

    def world_poses(self, refined_T, extrinsics=None) -> np.ndarray:
        """pipeline.cpp:136-137 for a batch: (n,4,4) refined transforms -> extrinsics @ inverse (extrinsics None: the inverse)."""
        T = np.asarray(refined_T, np.float32).reshape(-1, 4, 4)
        flat = np.ascontiguousarray(np.stack([_T_colmajor(t) for t in T])) if len(T) else np.zeros((0, 16), np.float32)
        ext = _T_colmajor(extrinsics) if extrinsics is not None else None
        out = np.empty_like(flat)
        self._check(self._L.b3d_world_poses(self._h, _ptr(flat), len(T), _ptr(ext), _ptr(out)))
        return np.stack([_T_from_colmajor(o) for o in out]) if len(T) else np.zeros((0, 4, 4), np.float32)

    def filter_duplicates(self, waypoints, min_distance: float) -> list:
        """Pipeline::filterDuplicates, pipeline.cpp:153-180."""
        T = np.asarray(waypoints, np.float32).reshape(-1, 4, 4)
        if not len(T):
            return []
        flat = np.ascontiguousarray(np.stack([_T_colmajor(t) for t in T])); out = np.empty_like(flat); n = C.c_size_t()
        self._check(self._L.b3d_filter_duplicates(self._h, _ptr(flat), len(T), float(min_distance), _ptr(out), C.byref(n)))
        return [_T_from_colmajor(out[i]) for i in range(n.value)]

    def depth_to_cloud(self, depth, mask, scale_to_meters, clipping_max, fx, fy, cx, cy, bgr=None):
        """pipeline.cpp:38-84. depth uint16 (h,w); mask uint8 (any size: read through the reference's nearest-neighbour
        resize, pipeline.cpp:39-41) or None; bgr uint8 (h,w,3) or None -> (xyz, rgb or None)."""
        depth = np.ascontiguousarray(depth, np.uint16); h, w = depth.shape
        mask = np.ascontiguousarray(mask, np.uint8) if mask is not None else None
        mh, mw = mask.shape if mask is not None else (0, 0)
        bgr = np.ascontiguousarray(bgr, np.uint8) if bgr is not None else None
        xyz = np.empty((h * w, 3), np.float32); rgb = np.empty((h * w, 3), np.float32) if bgr is not None else None
        n = C.c_size_t()
        self._check(self._L.b3d_depth_to_cloud(self._h, _ptr(depth), w, h, _ptr(mask), mw, mh, scale_to_meters, clipping_max, fx, fy, cx, cy, _ptr(bgr),
                                               _ptr(xyz), _ptr(rgb), h * w, C.byref(n)))
        return xyz[:n.value].copy(), (rgb[:n.value].copy() if rgb is not None else None)

    def register_depth(self, depth, mask, scale_to_meters, clipping_max, fx, fy, cx, cy, voxel_size, normals_k=30, fpfh_radius=None,
                       ransac_max_iterations=100000, confidence=0.999, icp_threshold=None, icp_max_iterations=200, point_to_plane=True):
        depth = np.ascontiguousarray(depth, np.uint16); h, w = depth.shape
        mask = np.ascontiguousarray(mask, np.uint8) if mask is not None else None
        mh, mw = mask.shape if mask is not None else (0, 0)
        radius = voxel_size * 5.0 if fpfh_radius is None else fpfh_radius
        thr = voxel_size * 0.4 if icp_threshold is None else icp_threshold
        r = SceneResult()
        self._check(self._L.b3d_register_depth(self._h, _ptr(depth), w, h, _ptr(mask), mw, mh, scale_to_meters, clipping_max, fx, fy, cx, cy, voxel_size,
                                               int(normals_k), radius, int(ransac_max_iterations), confidence, thr, int(icp_max_iterations),
                                               int(bool(point_to_plane)), C.byref(r)))
        self._n_src = r.n_source_points; self._H = int(ransac_max_iterations)
        return {"coarse": (_T_from_colmajor(np.array(r.coarse_T, np.float32)), r.coarse_fitness, r.coarse_rmse, r.coarse_best_iteration),
                "refined": (_T_from_colmajor(np.array(r.T, np.float32)), r.fitness, r.rmse, r.icp_iterations),
                "n_source_points": int(r.n_source_points)}

    # ---- stages feeding the hot path (registration.cpp:29-60, 105-130, 133-201)
    def voxel_downsample(self, xyz, voxel_size, colors=None):
        xyz = _as_f32(xyz, 3)
        n = xyz.shape[0]
        col = _as_f32(colors, 3) if colors is not None and np.asarray(colors).size else None
        out = np.empty((max(n, 1), 3), np.float32)
        out_col = np.empty((max(n, 1), 3), np.float32) if col is not None else None
        m = C.c_size_t()
        self._check(self._L.b3d_voxel_downsample(self._h, _ptr(xyz), n, _ptr(col) if col is not None else None, voxel_size,
                                                 _ptr(out), _ptr(out_col) if out_col is not None else None, n, C.byref(m)))
        return out[:m.value].copy(), (out_col[:m.value].copy() if out_col is not None else None)

    def estimate_normals(self, xyz, k=30):
        xyz = _as_f32(xyz, 3)
        out = np.empty_like(xyz)
        self._check(self._L.b3d_estimate_normals(self._h, _ptr(xyz), xyz.shape[0], int(k), _ptr(out)))
        return out

    def compute_fpfh(self, xyz, normals, radius):
        xyz = _as_f32(xyz, 3); nrm = _as_f32(normals, 3)
        if nrm.shape[0] != xyz.shape[0]:
            raise ValueError("compute_fpfh: one normal per point required")
        out = np.empty((xyz.shape[0], 33), np.float32)
        self._check(self._L.b3d_compute_fpfh(self._h, _ptr(xyz), _ptr(nrm), xyz.shape[0], radius, _ptr(out)))
        return out

    def icp_nearest(self, T, distance_threshold):
        Tc = _T_colmajor(T)
        idx = np.empty(self._n_src, np.uint32); d2 = np.empty(self._n_src, np.float32)
        self._check(self._L.b3d_icp_nearest(self._h, Tc.ctypes.data_as(_f32p), distance_threshold, _ptr(idx), _ptr(d2)))
        return idx, d2


def _as_f32(a, cols):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.size == 0:
        return a.reshape(0, cols)
    return a.reshape(-1, cols)
