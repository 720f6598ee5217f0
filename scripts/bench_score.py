"""Scoring-kernel timing on configs[2] (100k pairs) for a given hypothesis count; prints recount share."""
import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
b3d = importlib.import_module("3dvision_b200"); syn = b3d.synthetic
H = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
case = syn.ransac_case(max_iterations=H)
ctx = b3d.Context(0)
ctx.set_clouds(case.source, case.target); ctx.set_features(case.source_desc, case.target_desc)
ctx.match_features()
ctx.ransac_prepare(case.voxel_size, H, 2.0)
res = {}
for mode in (0, 4, 2, 1, 3, 3, 0):
    ctx.set_score_mode(mode)
    ctx.ransac_score()
    counts = ctx.ransac_counts()
    ms = ctx.stage_ms(2)
    rec = ctx.score_recounts()
    groups = H * (100352 // 32)
    print(f"mode={mode} score={ms:8.3f} ms  {H/ms/1e3:8.2f} M hyp/s  recounted groups={rec} ({100.0*rec/groups:.3f}% of hypothesis-groups)")
    res[mode] = counts
print("winner (mode 3 vs 0): max count", res[3].max(), res[0].max(), "argmax", int(res[3].argmax()), int(res[0].argmax()), "pruned", float((res[3] == -4).mean()))
print("counts identical:", np.array_equal(res[0], res[1]), np.array_equal(res[0], res[4]), " max inliers", res[0].max(), " good hyps (>10% inl):", (res[0] > 10000).mean())
