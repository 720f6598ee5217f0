"""One whole-pipeline registration (bench.py's `also.pipeline` workload) for profiling: python scripts/prof_pipeline.py [reps]"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
b3d = importlib.import_module("3dvision_b200"); syn = b3d.synthetic
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n_raw, voxel, H = 1_000_000, 0.0037, 100_000
rng = np.random.default_rng(1234 + 5)
model_raw = syn.rough_torus(n_raw, rng)
T_true = syn.rigid([0.2, 0.9, -0.3], 25.0, [0.05, -0.03, 0.08])
scene_raw = (syn.apply(np.linalg.inv(T_true), syn.rough_torus(n_raw, rng)) + rng.normal(0, 0.0003, (n_raw, 3))).astype(np.float32)
c = b3d.Context(0)
c.prepare_model(model_raw, voxel)
names = ["match", "ransac_prepare", "score", "select_finish", "icp_grid", "icp_iterations", "icp_binning", "voxel_downsample", "normals", "fpfh"]
for _ in range(reps):
    t0 = time.perf_counter()
    out = c.register_scene(scene_raw, voxel, ransac_max_iterations=H)
    ms = 1e3 * (time.perf_counter() - t0)
print(f"pipeline {ms:.2f} ms  icp_iterations={out['refined'][3]}  " + " ".join(f"{n}={c.stage_ms(i):.3f}" for i, n in enumerate(names)))
