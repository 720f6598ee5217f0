"""ctypes front-end of the CPU parity oracle (TEST INFRASTRUCTURE ONLY).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
(``3dvision_b200``) never does.  Functions mirror the reference entry points of
``/root/reference/src/registration.cpp`` (file:line cited per function).

Parity status: *unpinned* — see the header of ``registration_oracle.cpp``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile oracle/_build/liboracle.so with the committed Makefile."""
    srcs = [os.path.join(_HERE, f) for f in ("registration_oracle.cpp", "pipeline_inputs.cpp", "pose_post.cpp", "Makefile")]
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs if os.path.exists(s))
    if stale:
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _declare(_lib)
    return _lib


class use_library:
    """Context manager: route every call of this module to another build of the restatement (the variant libraries of
    `make -C oracle variants`, used by oracle/sensitivity.py)."""

    def __init__(self, path: str):
        self.path = path

    def __enter__(self):
        global _lib
        self.prev = _lib
        _lib = C.CDLL(self.path)
        _declare(_lib)
        return _lib

    def __exit__(self, *exc):
        global _lib
        _lib = self.prev


_f32p = C.POINTER(C.c_float)
_u32p = C.POINTER(C.c_uint32)
_i32p = C.POINTER(C.c_int32)
_u64p = C.POINTER(C.c_uint64)


def _declare(L):
    L.orc_mt19937_raw.argtypes = [C.c_uint32, C.c_size_t, _u32p]
    L.orc_uniform_indices_std.argtypes = [C.c_uint32, C.c_uint64, C.c_size_t, _u64p]
    L.orc_uniform_indices_lemire.argtypes = [C.c_uint32, C.c_uint64, C.c_size_t, _u64p]
    L.orc_uniform_indices_lemire.restype = C.c_size_t
    L.orc_svd3.argtypes = [_f32p] * 4
    L.orc_ldlt6_solve.argtypes = [_f32p] * 3
    L.orc_kabsch3.argtypes = [_f32p] * 4
    L.orc_euler_xyz.argtypes = [C.c_float, C.c_float, C.c_float, _f32p]
    L.orc_match_features.argtypes = [_f32p, C.c_size_t, C.c_size_t, _f32p, C.c_size_t, _u32p]
    L.orc_ransac.argtypes = [_f32p, C.c_size_t, _f32p, C.c_size_t, _u32p, C.c_float, C.c_int, C.c_float,
                             C.c_int, C.c_int, _f32p, _f32p, _f32p, _i32p, _i32p, _i32p]
    L.orc_ransac.restype = C.c_int
    L.orc_ransac_hypothesis.argtypes = [_f32p, C.c_size_t, _f32p, _u32p, C.c_int, _f32p, _f32p, _u64p]
    L.orc_ransac_hypothesis.restype = C.c_int
    L.orc_ransac_registration.argtypes = [_f32p, C.c_size_t, _f32p, C.c_size_t, _f32p, _f32p,
                                          C.c_float, C.c_int, C.c_float, _f32p, _f32p, _f32p]
    L.orc_icp.argtypes = [_f32p, C.c_size_t, _f32p, _f32p, C.c_size_t, _f32p, C.c_float, C.c_int, C.c_int, C.c_int,
                          _f32p, _f32p, _f32p, _i32p, _u32p, _f32p, _i32p]
    L.orc_demo_scene_points.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, _f32p]
    L.orc_demo_scene_points.restype = C.c_size_t
    L.orc_demo_model_points.argtypes = [_f32p]
    L.orc_demo_model_points.restype = C.c_size_t
    L.orc_voxel_downsample.argtypes = [_f32p, C.c_size_t, C.c_float, _f32p]
    L.orc_voxel_downsample.restype = C.c_size_t
    L.orc_estimate_normals.argtypes = [_f32p, C.c_size_t, C.c_int, _f32p]
    L.orc_compute_fpfh.argtypes = [_f32p, _f32p, C.c_size_t, C.c_float, _f32p]
    L.orc_mat4_inverse.argtypes = [_f32p, _f32p]
    L.orc_world_pose.argtypes = [_f32p, _f32p, _f32p]
    L.orc_filter_duplicates.argtypes = [_f32p, C.c_size_t, C.c_float, _f32p]
    L.orc_filter_duplicates.restype = C.c_size_t
    L.orc_resize_mask_nearest.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None:
        a = a.reshape(shape)
    return a


def _p(a, t=_f32p):
    return a.ctypes.data_as(t) if a is not None else None


@dataclass
class Result:
    """RegistrationResult (include/registration.hpp:26-30). transformation is 4x4 (row, col)."""
    transformation: np.ndarray
    fitness: float
    rmse: float
    extra: dict


def _T_from_colmajor(buf):
    return np.asarray(buf, dtype=np.float32).reshape(4, 4).T.copy()


def _T_to_colmajor(T):
    return np.ascontiguousarray(np.asarray(T, dtype=np.float32).reshape(4, 4).T).reshape(16)


# ---- RNG KATs ------------------------------------------------
def mt19937_raw(seed: int, count: int) -> np.ndarray:
    out = np.empty(count, np.uint32)
    lib().orc_mt19937_raw(seed, count, _p(out, _u32p))
    return out


def uniform_indices_std(seed: int, n: int, count: int) -> np.ndarray:
    out = np.empty(count, np.uint64)
    lib().orc_uniform_indices_std(seed, n, count, _p(out, _u64p))
    return out


def uniform_indices_lemire(seed: int, n: int, count: int):
    out = np.empty(count, np.uint64)
    used = lib().orc_uniform_indices_lemire(seed, n, count, _p(out, _u64p))
    return out, int(used)


# ---- linear algebra taps -----------------------------------------------------------
def svd3(M):
    M = _f32(M, (3, 3)); U = np.empty((3, 3), np.float32); V = np.empty((3, 3), np.float32); S = np.empty(3, np.float32)
    lib().orc_svd3(_p(M), _p(U), _p(V), _p(S))
    return U, S, V


def ldlt6_solve(A, b):
    A = _f32(A, (6, 6)); b = _f32(b, (6,)); x = np.empty(6, np.float32)
    lib().orc_ldlt6_solve(_p(A), _p(b), _p(x))
    return x


def kabsch3(s, q):
    s = _f32(s, (3, 3)); q = _f32(q, (3, 3)); R = np.empty((3, 3), np.float32); t = np.empty(3, np.float32)
    lib().orc_kabsch3(_p(s), _p(q), _p(R), _p(t))
    return R, t


def euler_xyz(a, b, g):
    R = np.empty((3, 3), np.float32)
    lib().orc_euler_xyz(a, b, g, _p(R))
    return R


# ---- hot path ------------------------------------------------------------------------
def match_features(src_desc, tgt_desc, row0=0, row1=None) -> np.ndarray:
    """registration.cpp:216-232 — returns correspondences[row0:row1] (uint32)."""
    sd = _f32(src_desc).reshape(-1, 33); td = _f32(tgt_desc).reshape(-1, 33)
    if row1 is None:
        row1 = sd.shape[0]
    corr = np.zeros(sd.shape[0], np.uint32)
    lib().orc_match_features(_p(sd), row0, row1, _p(td), td.shape[0], _p(corr, _u32p))
    return corr[row0:row1]


def ransac(src, tgt, corr, voxel_size, max_iterations=100000, confidence=0.999,
           iter_lo=0, iter_hi=None, want_counts=False) -> Result:
    """registration.cpp:234-295 given correspondences."""
    src = _f32(src).reshape(-1, 3); tgt = _f32(tgt).reshape(-1, 3)
    corr = np.ascontiguousarray(corr, np.uint32)
    if iter_hi is None:
        iter_hi = max_iterations
    T = np.empty(16, np.float32); fit = C.c_float(); rm = C.c_float()
    counts = np.empty(max_iterations, np.int32) if want_counts else None
    best = C.c_int32(); run = C.c_int32()
    lib().orc_ransac(_p(src), src.shape[0], _p(tgt), tgt.shape[0], _p(corr, _u32p), voxel_size, max_iterations,
                     confidence, iter_lo, iter_hi, _p(T), C.byref(fit), C.byref(rm),
                     _p(counts, _i32p) if want_counts else None, C.byref(best), C.byref(run))
    return Result(_T_from_colmajor(T), fit.value, rm.value,
                  {"counts": counts, "best_iter": best.value, "iters_run": run.value})


def ransac_hypothesis(src, tgt, corr, it):
    src = _f32(src).reshape(-1, 3); tgt = _f32(tgt).reshape(-1, 3); corr = np.ascontiguousarray(corr, np.uint32)
    R = np.empty((3, 3), np.float32); t = np.empty(3, np.float32); tri = np.empty(3, np.uint64)
    ok = lib().orc_ransac_hypothesis(_p(src), src.shape[0], _p(tgt), _p(corr, _u32p), it, _p(R), _p(t), _p(tri, _u64p))
    return bool(ok), R, t, tri


def ransac_registration(src, tgt, src_desc, tgt_desc, voxel_size, max_iterations=100000, confidence=0.999) -> Result:
    """Registration::ransacRegistration, registration.cpp:204-295."""
    src = _f32(src).reshape(-1, 3); tgt = _f32(tgt).reshape(-1, 3)
    sd = _f32(src_desc).reshape(-1, 33); td = _f32(tgt_desc).reshape(-1, 33)
    T = np.empty(16, np.float32); fit = C.c_float(); rm = C.c_float()
    lib().orc_ransac_registration(_p(src), src.shape[0], _p(tgt), tgt.shape[0], _p(sd), _p(td), voxel_size,
                                  max_iterations, confidence, _p(T), C.byref(fit), C.byref(rm))
    return Result(_T_from_colmajor(T), fit.value, rm.value, {})


def icp(src, tgt, tgt_normals, T0, distance_threshold, max_iterations=200, point_to_plane=True,
        stop_on_convergence=True, want_nn0=False) -> Result:
    """Registration::icpRefine, registration.cpp:297-414."""
    src = _f32(src).reshape(-1, 3); tgt = _f32(tgt).reshape(-1, 3)
    nrm = _f32(tgt_normals).reshape(-1, 3) if tgt_normals is not None else None
    if nrm is not None and nrm.shape[0] != tgt.shape[0]:
        nrm = None   # hasNormals() false (registration.hpp:17)
    T0c = _T_to_colmajor(T0)
    T = np.empty(16, np.float32); fit = C.c_float(); rm = C.c_float(); run = C.c_int32()
    nn = np.empty(src.shape[0], np.uint32) if want_nn0 else None
    d2 = np.empty(src.shape[0], np.float32) if want_nn0 else None
    ncorr = np.full(max(max_iterations, 1), -1, np.int32)
    lib().orc_icp(_p(src), src.shape[0], _p(tgt), _p(nrm), tgt.shape[0], _p(T0c), distance_threshold, max_iterations,
                  int(bool(point_to_plane)), int(bool(stop_on_convergence)), _p(T), C.byref(fit), C.byref(rm),
                  C.byref(run), _p(nn, _u32p) if want_nn0 else None, _p(d2) if want_nn0 else None, _p(ncorr, _i32p))
    return Result(_T_from_colmajor(T), fit.value, rm.value,
                  {"iters_run": run.value, "nn_idx0": nn, "nn_d2_0": d2, "ncorr": ncorr})


# ---- input builders for config 0 (demo scene), pipeline.cpp / registration.cpp:15-201 ----
def demo_scene_points(w=1280, h=720, scale=1000.0, clip=1.5) -> np.ndarray:
    out = np.empty((w * h, 3), np.float32)
    n = lib().orc_demo_scene_points(w, h, scale, clip, _p(out))
    return out[:n].copy()


def demo_model_points() -> np.ndarray:
    out = np.empty((41 * 41, 3), np.float32)
    n = lib().orc_demo_model_points(_p(out))
    return out[:n].copy()


def voxel_downsample(xyz, voxel) -> np.ndarray:
    xyz = _f32(xyz).reshape(-1, 3); out = np.empty_like(xyz)
    n = lib().orc_voxel_downsample(_p(xyz), xyz.shape[0], voxel, _p(out))
    return out[:n].copy()


def estimate_normals(xyz, k=30) -> np.ndarray:
    xyz = _f32(xyz).reshape(-1, 3); out = np.empty_like(xyz)
    lib().orc_estimate_normals(_p(xyz), xyz.shape[0], k, _p(out))
    return out


def compute_fpfh(xyz, normals, radius) -> np.ndarray:
    xyz = _f32(xyz).reshape(-1, 3); nrm = _f32(normals).reshape(-1, 3)
    out = np.empty((xyz.shape[0], 33), np.float32)
    lib().orc_compute_fpfh(_p(xyz), _p(nrm), xyz.shape[0], radius, _p(out))
    return out


def depth_to_cloud(depth, mask, scale_to_meters, clipping_max, fx, fy, cx, cy, bgr=None):
    """pipeline.cpp:38-84, CPU branch. depth uint16 (h,w); mask uint8 (h,w) or None; bgr uint8 (h,w,3) or None."""
    depth = np.ascontiguousarray(depth, np.uint16); h, w = depth.shape
    mask = np.ascontiguousarray(mask, np.uint8) if mask is not None else None
    bgr = np.ascontiguousarray(bgr, np.uint8) if bgr is not None else None
    xyz = np.empty((h * w, 3), np.float32); rgb = np.empty((h * w, 3), np.float32) if bgr is not None else None
    f = lib().orc_depth_to_cloud
    f.restype = C.c_size_t
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                  C.c_void_p, C.c_void_p, C.c_void_p]
    n = f(depth.ctypes.data, w, h, mask.ctypes.data if mask is not None else None, scale_to_meters, clipping_max, fx, fy, cx, cy,
          bgr.ctypes.data if bgr is not None else None, xyz.ctypes.data, rgb.ctypes.data if rgb is not None else None)
    return xyz[:n].copy(), (rgb[:n].copy() if rgb is not None else None)


# ---- pose post-processing and mask resize around the hot path (pipeline.cpp:38-41, 136-137, 153-180) ----
def mat4_inverse(M) -> np.ndarray:
    """Eigen::Matrix4f::inverse() (SSE 2x2-block kernel restated)."""
    out = np.empty(16, np.float32)
    lib().orc_mat4_inverse(_p(_T_to_colmajor(M)), _p(out))
    return _T_from_colmajor(out)


def world_pose(refined_T, extrinsics=None) -> np.ndarray:
    """pipeline.cpp:136-137: extrinsics * refined.transformation.inverse() (extrinsics None -> the inverse alone)."""
    out = np.empty(16, np.float32)
    ext = _T_to_colmajor(extrinsics) if extrinsics is not None else None
    lib().orc_world_pose(_p(ext), _p(_T_to_colmajor(refined_T)), _p(out))
    return _T_from_colmajor(out)


def filter_duplicates(waypoints, min_distance: float) -> list:
    """Pipeline::filterDuplicates, pipeline.cpp:153-180."""
    wps = [np.asarray(w, np.float32).reshape(4, 4) for w in waypoints]
    if not wps:
        return []
    flat = np.ascontiguousarray(np.stack([_T_to_colmajor(w) for w in wps]))
    out = np.empty_like(flat)
    n = lib().orc_filter_duplicates(_p(flat), len(wps), min_distance, _p(out))
    return [_T_from_colmajor(out[i]) for i in range(n)]


def resize_mask_nearest(mask, dst_w: int, dst_h: int) -> np.ndarray:
    """cv::resize(mask, ..., depth.size(), 0, 0, INTER_NEAREST), pipeline.cpp:39-41."""
    mask = np.ascontiguousarray(mask, np.uint8); sh, sw = mask.shape
    out = np.empty((dst_h, dst_w), np.uint8)
    lib().orc_resize_mask_nearest(mask.ctypes.data, sw, sh, dst_w, dst_h, out.ctypes.data)
    return out
