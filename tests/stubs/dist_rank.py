"""One rank of a 2+ GPU run of the sharded C-ABI path (spawned by tests/test_gpu_dist_nccl.py; no torch involved):
   python dist_rank.py <rank> <world> <workdir>
Rank 0 writes the ncclUniqueId to <workdir>/id.bin; every rank creates a context on device <rank>, joins the communicator
(b3d_comm_init), runs b3d_ransac_sharded (two confidences) and b3d_register_scene_sharded, and saves what it got."""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
b3d = importlib.import_module("3dvision_b200")
syn = b3d.synthetic


def main():
    rank, world, work = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    idf = os.path.join(work, "id.bin")
    ctx = b3d.Context(rank)
    if rank == 0:
        uid = ctx.comm_unique_id()
        with open(idf + ".tmp", "wb") as f:
            f.write(uid)
        os.replace(idf + ".tmp", idf)
    else:
        t0 = time.time()
        while not os.path.exists(idf):
            if time.time() - t0 > 120:
                raise SystemExit("no unique id from rank 0")
            time.sleep(0.05)
        uid = open(idf, "rb").read()
    ctx.comm_init(uid, rank, world)
    out = {}
    c = syn.ransac_case(n_src=20_001, n_tgt=15_000, seed=321, max_iterations=30_000)     # odd row count: a ragged last chunk
    for name, conf in (("full", 2.0), ("exit", 0.30)):
        T, fit, rmse, best = ctx.ransac_sharded(c.source, c.target, c.source_desc, c.target_desc, c.voxel_size, 30_000, conf)
        out[name + "_T"] = T; out[name + "_r"] = np.float32([fit, rmse]); out[name + "_id"] = np.int32([best])
    rng = np.random.default_rng(55)
    model = syn.rough_torus(60_000, rng)
    Tt = syn.rigid([0.2, 0.9, -0.3], 25.0, [0.05, -0.03, 0.08])
    scene = (syn.apply(np.linalg.inv(Tt), syn.rough_torus(60_000, rng)) + rng.normal(0, 0.0003, (60_000, 3))).astype(np.float32)
    ctx.prepare_model(model, 0.008)
    r = ctx.register_scene_sharded(scene, 0.008, ransac_max_iterations=20_000)
    out["scene_coarse"] = r["coarse"][0]; out["scene_T"] = r["refined"][0]
    out["scene_r"] = np.float32([r["coarse"][1], r["coarse"][2], r["refined"][1], r["refined"][2]]); out["scene_id"] = np.int32([r["coarse"][3], r["refined"][3]])
    np.savez(os.path.join(work, f"rank{rank}.npz"), **out)
    ctx.comm_destroy()
    ctx.close()


if __name__ == "__main__":
    main()
