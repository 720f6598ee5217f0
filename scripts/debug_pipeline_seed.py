"""Per-iteration comparison of the ICP stage of one fuzz_pipeline seed against the oracle (debug aid)."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
b3d = importlib.import_module("3dvision_b200._capi")
syn = importlib.import_module("3dvision_b200.synthetic")
from oracle import oracle

np.set_printoptions(precision=9, linewidth=200)
for seed in [int(a) for a in sys.argv[1:]]:
    c = syn.random_scene_case(seed)
    tgt = oracle.voxel_downsample(c["model"], c["voxel"]); src = oracle.voxel_downsample(c["scene"], c["voxel"])
    tn = oracle.estimate_normals(tgt, c["k"]); tf = oracle.compute_fpfh(tgt, tn, c["radius"])
    sn = oracle.estimate_normals(src, c["k"]); sf = oracle.compute_fpfh(src, sn, c["radius"])
    coarse = oracle.ransac_registration(src, tgt, sf, tf, c["voxel"], c["H"], c["conf"])
    full = oracle.icp(src, tgt, tn, coarse.transformation, c["icp_thr"], c["icp_iters"], c["plane"])
    print("seed", seed, "iters_run", full.extra["iters_run"], "ncorr", full.extra.get("ncorr"))
    with b3d.Context(0) as ctx:
        for k in range(1, full.extra["iters_run"] + 1):
            ref = oracle.icp(src, tgt, tn, coarse.transformation, c["icp_thr"], k, c["plane"])
            line = [f"k={k}"]
            for mode in (0, 3):
                ctx.set_icp_mode(mode)
                T, fit, rmse, n = ctx.icp(src, tgt, tn, coarse.transformation, c["icp_thr"], k, c["plane"])
                sameT = np.array_equal(T.view(np.uint32), ref.transformation.view(np.uint32))
                line.append(f"mode{mode}: T {'==' if sameT else 'DIFF %.3e' % np.abs(T - ref.transformation).max()} fit {fit == ref.fitness} rmse {rmse!r} vs {ref.rmse!r} n {n} vs {ref.extra['iters_run']}")
            ctx.set_icp_mode(0)
            print("  ", " | ".join(line))
