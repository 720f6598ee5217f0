"""Randomised soak of the stages that feed the hot path (voxelDownsample, estimateNormals, computeFPFH, depth -> cloud)
against the CPU oracle, through the C-ABI: random clouds incl. degenerate shapes, random voxel / k / radius; and of the pose
post-processing behind it (world poses through Eigen's 4x4 inverse, filterDuplicates).
Every output must equal the oracle's bit for bit.
usage: python scripts/fuzz_features.py [cases] [seed0]"""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
b3d = importlib.import_module("3dvision_b200._capi")
syn = importlib.import_module("3dvision_b200.synthetic")
from oracle import oracle  # noqa: E402  (test infrastructure: the checker)


def same_bits(a, b):
    a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def once(ctx, seed):
    xyz, voxel, k, radius = syn.random_cloud(seed)
    bad = []
    try:
        down, _ = ctx.voxel_downsample(xyz, voxel)
    except b3d.B3DError as e:                                     # |coordinate / voxel| >= 2^20 is refused, not mis-computed
        return [], f"seed {seed}: refused ({e})"
    want = oracle.voxel_downsample(xyz, voxel)
    if not same_bits(down, want):
        bad.append("voxel_downsample")
    pts = want if want.shape[0] >= 1 else xyz
    kk = min(k, 128)
    nrm_want = oracle.estimate_normals(pts, kk)
    if not same_bits(ctx.estimate_normals(pts, kk), nrm_want):
        bad.append("estimate_normals")
    f_want = oracle.compute_fpfh(pts, nrm_want, radius)
    if not same_bits(ctx.compute_fpfh(pts, nrm_want, radius), f_want):
        bad.append("compute_fpfh")
    # the raw cloud too (un-voxelised: duplicates, dense clusters)
    if xyz.shape[0] <= 2500:
        nr = oracle.estimate_normals(xyz, kk)
        if not same_bits(ctx.estimate_normals(xyz, kk), nr):
            bad.append("estimate_normals(raw)")
        if not same_bits(ctx.compute_fpfh(xyz, nr, radius), oracle.compute_fpfh(xyz, nr, radius)):
            bad.append("compute_fpfh(raw)")
    return bad, f"seed {seed}: n {xyz.shape[0]} -> {want.shape[0]} voxel {voxel:.3e} k {k} radius {radius:.3e}"


def depth_once(ctx, seed):
    rng = np.random.default_rng(seed)
    h, w = int(rng.integers(1, 200)), int(rng.integers(1, 300))
    depth = rng.integers(0, 4000, (h, w)).astype(np.uint16)
    depth[rng.random((h, w)) < 0.2] = 0
    mask = None
    if rng.random() < 0.7:
        mh, mw = (h, w) if rng.random() < 0.5 else (int(rng.integers(1, 120)), int(rng.integers(1, 160)))
        mask = (rng.random((mh, mw)) < 0.6).astype(np.uint8) * np.uint8(rng.choice([1, 255]))
    bgr = rng.integers(0, 256, (h, w, 3)).astype(np.uint8) if rng.random() < 0.5 else None
    scale = float(rng.choice([1000.0, 0.001, 4000.0, 1.0]))
    clip = float(rng.uniform(0.2, 5.0))
    fx, fy = float(rng.uniform(100, 1200)), float(rng.uniform(100, 1200))
    cx, cy = float(rng.uniform(0, w)), float(rng.uniform(0, h))
    mask_ref = mask if mask is None or mask.shape == (h, w) else oracle.resize_mask_nearest(mask, w, h)      # pipeline.cpp:39-41
    want = oracle.depth_to_cloud(depth, mask_ref, scale, clip, fx, fy, cx, cy, bgr)
    got = ctx.depth_to_cloud(depth, mask, scale, clip, fx, fy, cx, cy, bgr)
    ok = same_bits(got[0], want[0]) and ((bgr is None) or same_bits(got[1], want[1]))
    return ([] if ok else ["depth_to_cloud"]), f"depth seed {seed}: {h}x{w} mask {None if mask is None else mask.shape} scale {scale} clip {clip:.2f}"


def pose_once(ctx, seed):
    """T_world_object = extrinsics * refined^-1 (Eigen's SSE 4x4 inverse + packet product) and Pipeline::filterDuplicates."""
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 40))
    Ts = []
    for _ in range(n):
        kind = int(rng.integers(0, 5))
        if kind <= 1:
            T = syn.rigid(rng.standard_normal(3), float(rng.uniform(-180, 180)), rng.uniform(-2, 2, 3))
        elif kind == 2:
            T = rng.standard_normal((4, 4)) * 10.0 ** rng.uniform(-3, 3)                       # general
        elif kind == 3:
            T = rng.standard_normal((4, 4)); T[2] = T[0] * rng.uniform(-2, 2) + T[1] * 1e-7 * rng.standard_normal()   # nearly singular
        else:
            T = np.diag(10.0 ** rng.uniform(-6, 6, 4)) @ syn.rigid(rng.standard_normal(3), float(rng.uniform(-180, 180)), rng.uniform(-2, 2, 3))
        Ts.append(np.asarray(T, np.float32))
    Ts = np.stack(Ts)
    ext = None if rng.random() < 0.3 else np.asarray(syn.rigid(rng.standard_normal(3), float(rng.uniform(-180, 180)), rng.uniform(-2, 2, 3)), np.float32)
    bad = []
    with np.errstate(all="ignore"):
        got = ctx.world_poses(Ts, ext)
        for T, g in zip(Ts, got):
            w = oracle.world_pose(T, ext)
            if not (np.array_equal(g.view(np.uint32), w.view(np.uint32)) or (np.isnan(g) == np.isnan(w)).all() and np.array_equal(g[~np.isnan(g)], w[~np.isnan(w)])):
                bad.append("world_pose"); break
    centres = rng.uniform(-0.5, 0.5, (int(rng.integers(1, 12)), 3))
    wps = []
    for _ in range(int(rng.integers(0, 80))):
        W = np.asarray(syn.rigid(rng.standard_normal(3), float(rng.uniform(-180, 180)), [0, 0, 0]), np.float32)
        W[:3, 3] = centres[rng.integers(0, len(centres))] + rng.normal(0, 10.0 ** rng.uniform(-4, -1), 3)
        wps.append(W)
    min_d = float(rng.choice([0.0, 10.0 ** rng.uniform(-4, 0.5)]))
    want = oracle.filter_duplicates(wps, min_d); gotw = ctx.filter_duplicates(wps, min_d)
    if len(want) != len(gotw) or not all(np.array_equal(a, b) for a, b in zip(gotw, want)):
        bad.append("filter_duplicates")
    return bad, f"pose seed {seed}: {n} poses, {len(wps)} waypoints, min_distance {min_d:.3e}"


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    n_bad = 0
    t0 = time.time()
    with b3d.Context(0) as ctx:
        for s in range(seed0, seed0 + cases):
            for fn in (once, depth_once, pose_once):
                bad, what = fn(ctx, s)
                if bad:
                    n_bad += 1
                    print("MISMATCH", bad, what, flush=True)
    print(f"{cases} clouds + {cases} depth images + {cases} pose sets, {n_bad} mismatching cases, {time.time() - t0:.0f} s")
    sys.exit(1 if n_bad else 0)


if __name__ == "__main__":
    main()
