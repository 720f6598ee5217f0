"""Whole registration front to back through the reference-shaped interface: raw scene -> voxelDownsample -> estimateNormals ->
computeFPFH -> ransacRegistration -> icpRefine, against a model prepared once (as Pipeline::run does)."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
b3d = importlib.import_module("3dvision_b200")
syn = b3d.synthetic
reg = importlib.import_module("3dvision_b200.registration")
R = reg.Registration

n_raw = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
voxel = float(sys.argv[2]) if len(sys.argv) > 2 else 0.003
H = int(sys.argv[3]) if len(sys.argv) > 3 else 100_000
rng = np.random.default_rng(1239)
model_raw = syn.rough_torus(n_raw, rng)
scene_on_model = syn.rough_torus(n_raw, rng)
T_true = syn.rigid([0.2, 0.9, -0.3], 25.0, [0.05, -0.03, 0.08])            # scene -> model
scene_raw = (syn.apply(np.linalg.inv(T_true), scene_on_model) + rng.normal(0, 0.0003, (n_raw, 3))).astype(np.float32)

model = R.voxelDownsample(reg.PointCloud(model_raw), voxel)
R.estimateNormals(model, 30)
model_f = R.computeFPFH(model, voxel * 5.0)
ctx = reg._context(0)
ctx.set_score_mode(3)
for rep in range(4):
    t = [time.perf_counter()]
    src = R.voxelDownsample(reg.PointCloud(scene_raw), voxel); t.append(time.perf_counter())
    R.estimateNormals(src, 30); t.append(time.perf_counter())
    src_f = R.computeFPFH(src, voxel * 5.0); t.append(time.perf_counter())
    coarse = R.ransacRegistration(src, model, src_f, model_f, voxel, H, 0.999); t.append(time.perf_counter())
    st = [ctx.stage_ms(s) for s in range(4)]
    fine = R.icpRefine(src, model, coarse.transformation, voxel * 0.4, 200, True); t.append(time.perf_counter())
    d = [1e3 * (b - a) for a, b in zip(t[:-1], t[1:])]
    print(f"rep {rep}: {n_raw} raw -> {src.size()} src vs {model.size()} model | down {d[0]:.2f} normals {d[1]:.2f} fpfh {d[2]:.2f} "
          f"ransac {d[3]:.2f} (match {st[0]:.2f} prep {st[1]:.2f} score {st[2]:.2f} fin {st[3]:.2f}) icp {d[4]:.2f} | total {sum(d):.2f} ms | "
          f"ransac fit {coarse.fitness:.3f} icp fit {fine.fitness:.3f} rot err {syn.rotation_error(fine.transformation, T_true):.2e} "
          f"trans err {syn.translation_error(fine.transformation, T_true):.2e}")

c2 = b3d.Context(0)
c2.set_score_mode(3)
c2.prepare_model(model_raw, voxel)
for rep in range(4):
    t0 = time.perf_counter()
    out = c2.register_scene(scene_raw, voxel, ransac_max_iterations=H)
    dt = 1e3 * (time.perf_counter() - t0)
    st = {n: c2.stage_ms(i) for i, n in enumerate(["match", "prep", "score", "fin", "grid", "icp", "bin", "down", "normals", "fpfh"])}
    T = out["refined"][0]
    print(f"fused rep {rep}: {dt:.2f} ms | " + " ".join(f"{k} {v:.2f}" for k, v in st.items()) +
          f" | icp fit {out['refined'][1]:.3f} rot err {syn.rotation_error(T, T_true):.2e} trans err {syn.translation_error(T, T_true):.2e}")
