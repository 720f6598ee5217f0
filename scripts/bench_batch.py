"""Where a configs[3] instance spends its time, and how the worker pool scales on one GPU."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
b3d = importlib.import_module("3dvision_b200")
syn = b3d.synthetic
pipe = importlib.import_module("3dvision_b200.pipeline")
reg = importlib.import_module("3dvision_b200.registration")

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
cases = syn.batch_cases(n)
insts = [pipe.Instance(reg.PointCloud(c.source), reg.PointCloud(c.target, c.target_normals), reg.FPFHFeatures(c.source_desc),
                       reg.FPFHFeatures(c.target_desc), c.voxel_size) for c in cases]
ctx = reg._context(0)
for rep in range(2):
    for c in cases[:3]:
        t0 = time.perf_counter()
        T0, f0, r0, _ = ctx.ransac(c.source, c.target, c.source_desc, c.target_desc, c.voxel_size, 100000, 0.999)
        t1 = time.perf_counter()
        st = [ctx.stage_ms(s) for s in range(4)]
        T, fit, rmse, it = ctx.icp(c.source, c.target, c.target_normals, T0, c.voxel_size * 0.4, 200, True)
        t2 = time.perf_counter()
        print(f"n_tgt={c.target.shape[0]} ransac {1e3*(t1-t0):.2f} ms (match {st[0]:.2f} prep {st[1]:.2f} score {st[2]:.2f} fin {st[3]:.2f}) "
              f"icp {1e3*(t2-t1):.2f} ms ({it} it; grid {ctx.stage_ms(4):.2f} bin {ctx.stage_ms(6):.2f} iters {ctx.stage_ms(5):.2f}) fit {f0:.3f}->{fit:.3f}")
for threads in (1, 2, 4, 8, 16):
    pipe.register_batch(insts, threads)
    t0 = time.perf_counter()
    pipe.register_batch(insts, threads)
    dt = time.perf_counter() - t0
    print(f"threads={threads}: {1e3*dt/n:.2f} ms/instance, {n/dt:.1f} reg/s")
