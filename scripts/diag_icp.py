"""Diagnostic: GPU ICP vs oracle, per iteration count and mode."""
import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
b3d = importlib.import_module("3dvision_b200"); syn = b3d.synthetic
from oracle import oracle as O
ctx = b3d.Context(0)
case = syn.icp_case(n_model=6000, n_scene=9000, seed=21)
for plane in (True, False):
    for iters in (1, 2, 3, 5, 10, 30):
        ref = O.icp(case.source, case.target, case.target_normals, case.T_init, case.threshold, iters, plane)
        T, fit, rmse, it = ctx.icp(case.source, case.target, case.target_normals, case.T_init, case.threshold, iters, plane)
        print(f"plane={plane} iters={iters} run gpu/ref={it}/{ref.extra['iters_run']} fit {fit==ref.fitness} drmse={abs(rmse-ref.rmse):.2e} "
              f"rot={syn.rotation_error(T, ref.transformation):.2e} trans={syn.translation_error(T, ref.transformation):.2e}")
