"""Stage times of the feature front end (voxelDownsample -> estimateNormals -> computeFPFH) on a synthetic scene."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
b3d = importlib.import_module("3dvision_b200")
syn = b3d.synthetic

n_raw = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
voxel = float(sys.argv[2]) if len(sys.argv) > 2 else 0.003
rng = np.random.default_rng(7)
raw, _ = syn.torus(n_raw, rng)
raw = (raw + rng.normal(0, 0.0003, raw.shape)).astype(np.float32)
ctx = b3d.Context(0)
for rep in range(3):
    t0 = time.perf_counter(); pts, _ = ctx.voxel_downsample(raw, voxel); t1 = time.perf_counter()
    nrm = ctx.estimate_normals(pts, 30); t2 = time.perf_counter()
    desc = ctx.compute_fpfh(pts, nrm, voxel * 5.0); t3 = time.perf_counter()
    print(f"rep {rep}: raw {n_raw} -> {pts.shape[0]} pts | downsample {1e3*(t1-t0):.2f} ms (device {ctx.stage_ms(7):.2f}) | "
          f"normals {1e3*(t2-t1):.2f} ms (device {ctx.stage_ms(8):.2f}) | fpfh {1e3*(t3-t2):.2f} ms (device {ctx.stage_ms(9):.2f})")
