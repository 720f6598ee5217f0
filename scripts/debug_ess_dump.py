"""Dump the exact-sum terms of iteration k of one fuzz_pipeline seed's ICP and re-add them on the host (debug aid)."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
b3d = importlib.import_module("3dvision_b200._capi")
syn = importlib.import_module("3dvision_b200.synthetic")
from oracle import oracle

seed, k = int(sys.argv[1]), int(sys.argv[2])
c = syn.random_scene_case(seed)
tgt = oracle.voxel_downsample(c["model"], c["voxel"]); src = oracle.voxel_downsample(c["scene"], c["voxel"])
tn = oracle.estimate_normals(tgt, c["k"]); tf = oracle.compute_fpfh(tgt, tn, c["radius"])
sn = oracle.estimate_normals(src, c["k"]); sf = oracle.compute_fpfh(src, sn, c["radius"])
coarse = oracle.ransac_registration(src, tgt, sf, tf, c["voxel"], c["H"], c["conf"])
with b3d.Context(0) as ctx:
    ctx.icp(src, tgt, tn, coarse.transformation, c["icp_thr"], k, c["plane"])
    terms, sums = ctx.icp_exact_sum_dump(28)
    n = src.shape[0]
    print("n", n, "stride", terms.shape[1])
    for v in range(28):
        x = terms[v, :n]
        want = np.add.accumulate(x, dtype=np.float32)[-1]
        again, st = ctx.sequential_sum(x)
        flag = "" if (want.view(np.uint32) == sums[v].view(np.uint32)) else "   <-- ICP chain differs"
        flag2 = "" if (want.view(np.uint32) == np.float32(again).view(np.uint32)) else "   <-- sequential_sum differs"
        print(v, repr(want), repr(sums[v]), repr(again), int((x != 0).sum()), flag, flag2)
    np.save("gpurun_out/ess_terms_seed%d_k%d.npy" % (seed, k), terms[:, :n])
