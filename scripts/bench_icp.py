"""Quick ICP timing on configs[1] (300k x 100k, point-to-plane, 50 iterations) and a C1-scale case."""
import importlib, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
b3d = importlib.import_module("3dvision_b200"); syn = b3d.synthetic
ctx = b3d.Context(0)
for name, kw, thr_scale in [("C2 300k x 100k thr 5mm", {}, 1.0), ("C2 thr 1mm", {}, 0.2), ("30k x 5k", dict(n_model=5000, n_scene=30000), 1.0)]:
    ic = syn.icp_case(**kw)
    thr = ic.threshold * thr_scale
    ctx.set_clouds(ic.source, ic.target, ic.target_normals)
    ctx.set_icp_mode(0)
    ctx.icp_run(ic.T_init, thr, 12, True, False)      # long enough to build (and allocate) the lazily built second level
    for plane, mode in ((True, 0), (True, 1), (True, 3), (False, 0), (False, 1), (False, 3)):
        ctx.set_icp_mode(mode)
        t0 = time.perf_counter()
        T, fit, rmse, it = ctx.icp_run(ic.T_init, thr, 50, plane, False)
        wall = time.perf_counter() - t0
        print(f"{name:26s} plane={plane} mode={mode} iters={it} build={ctx.stage_ms(4)*1e3:6.1f}+{max(ctx.stage_ms(6),0)*1e3:5.1f} us  loop={ctx.stage_ms(5):8.3f} ms  per-iter={ctx.stage_ms(5)/50*1e3:7.1f} us "
              f"wall={wall*1e3:7.2f} ms fit={fit:.4f} rot_err={syn.rotation_error(T, ic.T_true):.2e}")
