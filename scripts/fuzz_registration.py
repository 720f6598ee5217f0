"""Randomised soak of the whole hot path against the CPU oracle, through the C-ABI: small random ICP cases (both error
metrics, random sizes / noise / thresholds / start poses / iteration caps) and small random ransacRegistration cases
(random sizes, inlier ratios, hypothesis counts, confidences).  Every output must equal the oracle's bit for bit.
usage: python scripts/fuzz_registration.py [icp_cases] [ransac_cases] [seed0]"""
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


def same(a, b):
    a, b = np.float32(a), np.float32(b)
    return a.view(np.uint32) == b.view(np.uint32) or (np.isnan(a) and np.isnan(b))


def icp_once(ctx, seed):
    c, plane = syn.random_icp_case(seed)
    ref = oracle.icp(c.source, c.target, c.target_normals, c.T_init, c.threshold, c.iterations, plane)
    T, fit, rmse, n = ctx.icp(c.source, c.target, c.target_normals, c.T_init, c.threshold, c.iterations, plane)
    ok = n == ref.extra["iters_run"] and np.array_equal(T.view(np.uint32), ref.transformation.view(np.uint32)) and same(fit, ref.fitness) and same(rmse, ref.rmse)
    return ok, f"icp seed {seed}: model {c.target.shape[0]} scene {c.source.shape[0]} thr {c.threshold:.2e} iters {c.iterations} plane {plane}: " \
               f"got it {n} fit {fit} rmse {rmse}; ref it {ref.extra['iters_run']} fit {ref.fitness} rmse {ref.rmse}"


def ransac_once(ctx, seed):
    c, conf = syn.random_ransac_case(seed)
    H = c.max_iterations
    ref = oracle.ransac_registration(c.source, c.target, c.source_desc, c.target_desc, c.voxel_size, H, conf)
    T, fit, rmse = ctx.ransac(c.source, c.target, c.source_desc, c.target_desc, c.voxel_size, H, conf)[:3]
    ok = np.array_equal(T.view(np.uint32), ref.transformation.view(np.uint32)) and same(fit, ref.fitness) and same(rmse, ref.rmse)
    return ok, f"ransac seed {seed}: src {c.source.shape[0]} tgt {c.target.shape[0]} H {H} conf {conf} voxel {c.voxel_size:.2e}: " \
               f"got fit {fit} rmse {rmse}; ref fit {ref.fitness} rmse {ref.rmse}"


def main():
    n_icp = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    n_ransac = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    seed0 = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    bad = 0
    t0 = time.time()
    with b3d.Context(0) as ctx:
        for fn, count in ((icp_once, n_icp), (ransac_once, n_ransac)):
            for s in range(seed0, seed0 + count):
                ok, what = fn(ctx, s)
                if not ok:
                    bad += 1
                    print("MISMATCH", what, flush=True)
    print(f"{n_icp} icp + {n_ransac} ransac cases, {bad} mismatches, {time.time() - t0:.0f} s")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
