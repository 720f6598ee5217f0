"""How a rank's share of configs[3] behaves: n instances through b3d_pool with w workers on one GPU (python scripts/bench_batch8.py)."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
b3d = importlib.import_module("3dvision_b200"); syn = b3d.synthetic
cases = syn.batch_cases(64)
insts = [dict(source=c.source, target=c.target, target_normals=c.target_normals, source_desc=c.source_desc, target_desc=c.target_desc, voxel_size=c.voxel_size) for c in cases]
for n, w in ((1, 1), (8, 1), (8, 4), (8, 8), (16, 8), (64, 8), (64, 16)):
    with b3d.Pool(w, devices=(0,)) as pool:
        pool.register(insts[:n])
        tt = []
        for _ in range(3):
            t0 = time.perf_counter(); pool.register(insts[:n]); tt.append(time.perf_counter() - t0)
        t = float(np.median(tt))
        print(f"instances={n:3d} workers={w:2d}  batch {1e3*t:7.2f} ms  per instance {1e3*t/n:6.2f} ms  {n/t:7.1f} reg/s")
