"""Randomised soak of the C worker pool (b3d_pool: host threads in C, one context each, sharing one device): random batches
of instances with very different sizes run concurrently must give, instance by instance, the bits of the same instance run
alone on a fresh context (ransacRegistration + icpRefine).  Looks for races on process-wide state (cached device probe,
shared-memory opt-in flags, the stream-ordered allocator) rather than for arithmetic.
usage: python scripts/fuzz_pool.py [batches] [seed0]"""
import importlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
b3d = importlib.import_module("3dvision_b200._capi")
syn = importlib.import_module("3dvision_b200.synthetic")


def main():
    batches = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    bad = 0
    t0 = time.time()
    n_inst = 0
    for s in range(seed0, seed0 + batches):
        rng = np.random.default_rng(s)
        insts = []
        for j in range(int(rng.integers(1, 24))):
            c = syn.ransac_case(n_src=int(rng.integers(50, 30_000)), n_tgt=int(rng.integers(50, 10_000)), seed=1000 * s + j,
                                inlier_frac=float(rng.uniform(0.2, 0.95)), voxel=float(10.0 ** rng.uniform(-3.2, -2.2)),
                                noise=float(10.0 ** rng.uniform(-4.5, -3.0)), max_iterations=int(rng.choice([200, 3000, 30_000])))
            insts.append(dict(source=c.source, target=c.target, target_normals=c.target_normals, source_desc=c.source_desc, target_desc=c.target_desc,
                              voxel_size=c.voxel_size, ransac_iterations=c.max_iterations, confidence=float(rng.choice([0.5, 0.999, 2.0])),
                              icp_threshold=float(c.voxel_size * rng.choice([0.4, 1.0, 3.0])), icp_iterations=int(rng.choice([3, 30, 200])),
                              point_to_plane=bool(rng.random() < 0.7)))
        with b3d.Pool(int(rng.integers(1, 13)), devices=(0,)) as pool:
            got = pool.register(insts)
            again = pool.register(insts[::-1])[::-1]                 # same pool, other assignment of instances to workers
        n_inst += len(insts)
        for j, (inst, g, g2) in enumerate(zip(insts, got, again)):
            with b3d.Context(0) as ctx:
                T0, f0, r0 = ctx.ransac(inst["source"], inst["target"], inst["source_desc"], inst["target_desc"], inst["voxel_size"],
                                        inst["ransac_iterations"], inst["confidence"])[:3]
                T1, f1, r1, it = ctx.icp(inst["source"], inst["target"], inst["target_normals"], T0, inst["icp_threshold"], inst["icp_iterations"],
                                         inst["point_to_plane"])
            for tag, (gc, gf) in (("first", g), ("reversed", g2)):
                ok = np.array_equal(gc[0].view(np.uint32), T0.view(np.uint32)) and np.float32(gc[1]) == np.float32(f0) and np.float32(gc[2]) == np.float32(r0) \
                    and np.array_equal(gf[0].view(np.uint32), T1.view(np.uint32)) and np.float32(gf[1]) == np.float32(f1) and np.float32(gf[2]) == np.float32(r1) and gf[3] == it
                if not ok:
                    bad += 1
                    print(f"MISMATCH batch {s} instance {j} ({tag}): {inst['source'].shape[0]} x {inst['target'].shape[0]}", flush=True)
    print(f"{batches} batches, {n_inst} instances x 2 runs, {bad} mismatches, {time.time() - t0:.0f} s")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
