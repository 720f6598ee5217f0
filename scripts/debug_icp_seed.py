"""Per-iteration comparison of one fuzz_registration ICP seed against the oracle in every ICP mode (debug aid)."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
b3d = importlib.import_module("3dvision_b200._capi")
syn = importlib.import_module("3dvision_b200.synthetic")
from oracle import oracle

def case_of(seed):
    c, plane = syn.random_icp_case(seed)
    return c, c.threshold, c.iterations, plane

with b3d.Context(0) as ctx:
    for seed in [int(a) for a in sys.argv[1:]]:
        c, thr, iters, plane = case_of(seed)
        full = oracle.icp(c.source, c.target, c.target_normals, c.T_init, thr, iters, plane)
        print("seed", seed, "iters_run", full.extra["iters_run"], "ncorr", full.extra.get("ncorr"))
        for k in range(1, full.extra["iters_run"] + 1):
            ref = oracle.icp(c.source, c.target, c.target_normals, c.T_init, thr, k, plane)
            line = [f"k={k}"]
            for mode in (0, 3, 1):
                ctx.set_icp_mode(mode)
                T, fit, rmse, n = ctx.icp(c.source, c.target, c.target_normals, c.T_init, thr, k, plane)
                same = np.array_equal(T.view(np.uint32), ref.transformation.view(np.uint32))
                line.append(f"mode{mode}:{'==' if same else 'DIFF %.3e' % np.abs(T - ref.transformation).max()}")
            ctx.set_icp_mode(0)
            print("  ", " ".join(line))
            if "DIFF" in line[1]:
                T, *_ = ctx.icp(c.source, c.target, c.target_normals, c.T_init, thr, k, plane)
                np.set_printoptions(precision=9, linewidth=200)
                print("device\n", T, "\noracle\n", ref.transformation)
                break
