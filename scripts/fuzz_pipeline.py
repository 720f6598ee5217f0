"""Randomised soak of the one-call resident registration (b3d_prepare_model + b3d_register_scene: voxelDownsample ->
estimateNormals -> computeFPFH -> ransacRegistration -> icpRefine on the device, src/pipeline.cpp:86-129) against the same
chain run stage by stage through the CPU oracle.  Random model / scene clouds (surface, noisy, partial, sometimes
degenerate), random voxel, k, FPFH radius, hypothesis count, confidence, ICP threshold / metric / iteration cap.
usage: python scripts/fuzz_pipeline.py [cases] [seed0]"""
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


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def once(seed):
    c = syn.random_scene_case(seed)
    tgt = oracle.voxel_downsample(c["model"], c["voxel"])
    src = oracle.voxel_downsample(c["scene"], c["voxel"])
    if tgt.shape[0] == 0 or src.shape[0] == 0:
        return True, f"seed {seed}: empty after down-sampling"
    tn = oracle.estimate_normals(tgt, c["k"]); tf = oracle.compute_fpfh(tgt, tn, c["radius"])
    sn = oracle.estimate_normals(src, c["k"]); sf = oracle.compute_fpfh(src, sn, c["radius"])
    coarse = oracle.ransac_registration(src, tgt, sf, tf, c["voxel"], c["H"], c["conf"])
    fine = oracle.icp(src, tgt, tn, coarse.transformation, c["icp_thr"], c["icp_iters"], c["plane"])
    with b3d.Context(0) as ctx:
        m = ctx.prepare_model(c["model"], c["voxel"], c["k"], c["radius"])
        out = ctx.register_scene(c["scene"], c["voxel"], c["k"], c["radius"], c["H"], c["conf"], c["icp_thr"], c["icp_iters"], c["plane"])
    T0, f0, r0, _ = out["coarse"]; T1, f1, r1, it = out["refined"]
    ok = m == tgt.shape[0] and out["n_source_points"] == src.shape[0] \
        and np.array_equal(bits(T0), bits(coarse.transformation)) and bits(f0) == bits(coarse.fitness) and bits(r0) == bits(coarse.rmse) \
        and np.array_equal(bits(T1), bits(fine.transformation)) and bits(f1) == bits(fine.fitness) and bits(r1) == bits(fine.rmse) \
        and it == fine.extra["iters_run"]
    return ok, (f"seed {seed}: model {c['model'].shape[0]}->{tgt.shape[0]} scene {c['scene'].shape[0]}->{src.shape[0]} voxel {c['voxel']:.3e} k {c['k']} "
                f"radius {c['radius']:.3e} H {c['H']} conf {c['conf']} icp thr {c['icp_thr']:.2e} x{c['icp_iters']} plane {c['plane']}: "
                f"coarse fit {f0} vs {coarse.fitness}; fine fit {f1} rmse {r1} it {it} vs {fine.fitness} {fine.rmse} {fine.extra['iters_run']}")


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    bad = 0
    t0 = time.time()
    for s in range(seed0, seed0 + cases):
        ok, what = once(s)
        if not ok:
            bad += 1
            print("MISMATCH", what, flush=True)
    print(f"{cases} scenes, {bad} mismatches, {time.time() - t0:.0f} s")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
