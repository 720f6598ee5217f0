#!/usr/bin/env python
"""Diff the restated oracle against the REAL reference (oracle/_ref/libref.so, built by `make -C oracle ref EIGEN_DIR=...`
from the unmodified /root/reference/src/registration.cpp) on every golden / seeded case.  Test infrastructure only.

  python oracle/diff_ref.py          exit 0 and "PINNED" if every compared output is bit-identical;
                                     exit 1 with a per-case table otherwise; exit 2 if libref.so does not exist
                                     (this image: no Eigen -> `make ref` cannot run).
The day an Eigen checkout is available this turns "parity unpinned" into a measured statement in one command.
"""
import ctypes as C
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

syn = importlib.import_module("3dvision_b200.synthetic")
REF = os.path.join(HERE, "_ref", "libref.so")
f32p = C.POINTER(C.c_float)


def p(a):
    return a.ctypes.data_as(f32p) if a is not None else None


class Ref:
    def __init__(self, path):
        L = self.L = C.CDLL(path)
        L.ref_ransac_registration.argtypes = [f32p, C.c_size_t, f32p, C.c_size_t, f32p, f32p, C.c_float, C.c_int, C.c_float, f32p, f32p, f32p]
        L.ref_icp.argtypes = [f32p, C.c_size_t, f32p, f32p, C.c_size_t, f32p, C.c_float, C.c_int, C.c_int, f32p, f32p, f32p]
        L.ref_voxel_downsample.argtypes = [f32p, C.c_size_t, C.c_float, f32p]; L.ref_voxel_downsample.restype = C.c_size_t
        L.ref_estimate_normals.argtypes = [f32p, C.c_size_t, C.c_int, f32p]
        L.ref_compute_fpfh.argtypes = [f32p, f32p, C.c_size_t, C.c_float, f32p]

    def ransac_registration(self, src, tgt, sd, td, voxel, H, conf):
        T = np.empty(16, np.float32); fit = C.c_float(); rm = C.c_float()
        self.L.ref_ransac_registration(p(src), len(src), p(tgt), len(tgt), p(sd), p(td), voxel, H, conf, p(T), C.byref(fit), C.byref(rm))
        return T.reshape(4, 4).T.copy(), fit.value, rm.value

    def icp(self, src, tgt, nrm, T0, thr, iters, plane):
        T0c = np.ascontiguousarray(np.asarray(T0, np.float32).T).reshape(16)
        T = np.empty(16, np.float32); fit = C.c_float(); rm = C.c_float()
        self.L.ref_icp(p(src), len(src), p(tgt), p(nrm), len(tgt), p(T0c), thr, iters, int(plane), p(T), C.byref(fit), C.byref(rm))
        return T.reshape(4, 4).T.copy(), fit.value, rm.value

    def voxel_downsample(self, xyz, voxel):
        out = np.empty_like(xyz); n = self.L.ref_voxel_downsample(p(xyz), len(xyz), voxel, p(out)); return out[:n].copy()

    def estimate_normals(self, xyz, k):
        out = np.empty_like(xyz); self.L.ref_estimate_normals(p(xyz), len(xyz), k, p(out)); return out

    def compute_fpfh(self, xyz, nrm, radius):
        out = np.empty((len(xyz), 33), np.float32); self.L.ref_compute_fpfh(p(xyz), p(nrm), len(xyz), radius, p(out)); return out


def same(a, b):
    a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def main():
    if not os.path.exists(REF):
        print(f"{REF} not built: run `make -C oracle ref EIGEN_DIR=/path/to/eigen3` (needs Eigen >= 3.3 headers)")
        return 2
    R = Ref(REF)
    rows = []
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    # front end: down-sample (container order), normals, FPFH
    raw = f32(syn.torus(6000, np.random.default_rng(97))[0])
    pts_o, pts_r = O.voxel_downsample(raw, 0.006), R.voxel_downsample(raw, 0.006)
    rows.append(("voxelDownsample torus", same(pts_o, pts_r)))
    nrm_o, nrm_r = O.estimate_normals(pts_o, 30), R.estimate_normals(pts_o, 30)
    rows.append(("estimateNormals torus", same(nrm_o, nrm_r)))
    rows.append(("computeFPFH torus", same(O.compute_fpfh(pts_o, nrm_o, 0.03), R.compute_fpfh(pts_o, nrm_o, 0.03))))
    # hot path on the seeded cases the GPU parity tests use
    for n_src, n_tgt, seed, H in ((3000, 2500, 7, 4000), (2977, 1842, 94, 4000), (20000, 3000, 95, 2000)):
        c = syn.ransac_case(n_src=n_src, n_tgt=n_tgt, seed=seed, max_iterations=H)
        o = O.ransac_registration(c.source, c.target, c.source_desc, c.target_desc, c.voxel_size, H, 0.999)
        Tr, fr, rr = R.ransac_registration(f32(c.source), f32(c.target), f32(c.source_desc), f32(c.target_desc), c.voxel_size, H, 0.999)
        rows.append((f"ransacRegistration {n_src}x{n_tgt} H={H}", same(o.transformation, Tr) and o.fitness == fr and o.rmse == rr))
        T0 = c.T_true.copy(); T0[:3, 3] += np.float32(2e-4)
        for plane, thr, iters in ((True, c.voxel_size * 0.4, 30), (False, 0.004, 15)):
            oi = O.icp(c.source, c.target, c.target_normals, T0, thr, iters, plane)
            Ti, fi, ri = R.icp(f32(c.source), f32(c.target), f32(c.target_normals), T0, thr, iters, plane)
            rows.append((f"icpRefine {'plane' if plane else 'point'} {n_src}x{n_tgt} thr={thr:.4g}",
                         same(oi.transformation, Ti) and oi.fitness == fi and oi.rmse == ri))
    # configs[0] front to back
    voxel = 0.001
    src = O.voxel_downsample(O.demo_scene_points(), voxel); tgt = O.voxel_downsample(O.demo_model_points(), voxel)
    rows.append(("configs[0] voxelDownsample", same(src, R.voxel_downsample(f32(O.demo_scene_points()), voxel))))
    tn = O.estimate_normals(tgt, 30); rows.append(("configs[0] model normals", same(tn, R.estimate_normals(tgt, 30))))
    tf = O.compute_fpfh(tgt, tn, voxel * 5.0); rows.append(("configs[0] model FPFH", same(tf, R.compute_fpfh(tgt, tn, voxel * 5.0))))
    bad = [n for n, ok in rows if not ok]
    for n, ok in rows:
        print(f"{'same bits' if ok else 'DIFFERS  '}  {n}")
    print("PINNED: the restatement equals the reference on every case" if not bad else f"{len(bad)} of {len(rows)} cases differ")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
