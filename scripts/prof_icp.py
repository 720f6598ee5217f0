"""One ICP configuration for profiling: python scripts/prof_icp.py [mode] [plane 0/1] [iters] [thr_scale] [n_model n_scene]"""
import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
b3d = importlib.import_module("3dvision_b200"); syn = b3d.synthetic
a = sys.argv[1:]
mode = int(a[0]) if len(a) > 0 else 0
plane = bool(int(a[1])) if len(a) > 1 else True
iters = int(a[2]) if len(a) > 2 else 12
thr_scale = float(a[3]) if len(a) > 3 else 1.0
kw = dict(n_model=int(a[4]), n_scene=int(a[5])) if len(a) > 5 else {}
ctx = b3d.Context(0)
ic = syn.icp_case(**kw)
ctx.set_clouds(ic.source, ic.target, ic.target_normals)
ctx.set_icp_mode(mode)
T, fit, rmse, it = ctx.icp_run(ic.T_init, ic.threshold * thr_scale, iters, plane, False)
print(f"mode={mode} plane={plane} iters={it} loop={ctx.stage_ms(5):.3f} ms per-iter={ctx.stage_ms(5)/max(it,1)*1e3:.1f} us fit={fit:.4f}")
if mode in (0, 2):
    st = ctx.icp_exact_sum_stats()
    import numpy as np
    nblk = -(-ic.source.shape[0] // 32)
    print("per iteration and sum: walk rounds / term-by-term blocks (of %d) / chain us @1.965GHz:" % nblk)
    print(" ".join(f"{int(r)//it}/{int(q)//it}/{int(cy)*16/1965/it:.0f}" for r, q, cy, _ in st if cy))
