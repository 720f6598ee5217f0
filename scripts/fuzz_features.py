"""Randomised soak of the stages that feed the hot path (voxelDownsample, estimateNormals, computeFPFH, depth -> cloud)
against the CPU oracle, through the C-ABI: random clouds incl. degenerate shapes, random voxel / k / radius.
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


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    n_bad = 0
    t0 = time.time()
    with b3d.Context(0) as ctx:
        for s in range(seed0, seed0 + cases):
            for fn in (once, depth_once):
                bad, what = fn(ctx, s)
                if bad:
                    n_bad += 1
                    print("MISMATCH", bad, what, flush=True)
    print(f"{cases} clouds + {cases} depth images, {n_bad} mismatching cases, {time.time() - t0:.0f} s")
    sys.exit(1 if n_bad else 0)


if __name__ == "__main__":
    main()
