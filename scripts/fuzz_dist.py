"""Randomised soak of the sharded C-ABI path over real NCCL ranks (one process per GPU, no torch): every rank runs
b3d_ransac_sharded on random cases (sizes, inlier ratios, hypothesis counts, confidences -> exits in any rank's shard,
ragged last chunks, more ranks than rows) and compares it bit for bit with the plain one-GPU b3d_ransac of a second,
communicator-less context on its own device.
usage: python scripts/fuzz_dist.py <world> [cases] [seed0]        (spawns the ranks itself)"""
import importlib
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)


def rank_main(rank, world, work, cases, seed0):
    b3d = importlib.import_module("3dvision_b200._capi")
    syn = importlib.import_module("3dvision_b200.synthetic")
    idf = os.path.join(work, "id.bin")
    ctx = b3d.Context(rank)
    lone = b3d.Context(rank)
    if rank == 0:
        with open(idf + ".tmp", "wb") as f:
            f.write(ctx.comm_unique_id())
        os.replace(idf + ".tmp", idf)
    else:
        t0 = time.time()
        while not os.path.exists(idf):
            if time.time() - t0 > 120:
                raise SystemExit("no unique id from rank 0")
            time.sleep(0.05)
    ctx.comm_init(open(idf, "rb").read(), rank, world)
    bad = 0
    for s in range(seed0, seed0 + cases):
        rng = np.random.default_rng(s)
        n_src, n_tgt = int(rng.integers(1, 30_000)), int(rng.integers(1, 12_000))
        if rng.random() < 0.1:
            n_src = int(rng.integers(1, 3 * world))                       # fewer rows than ranks
        H = int(rng.choice([1, 3, world - 1 if world > 1 else 1, 100, 5_000, 40_000]))
        conf = float(rng.choice([0.02, 0.3, 0.999, 2.0]))
        c = syn.ransac_case(n_src=n_src, n_tgt=n_tgt, seed=s, inlier_frac=float(rng.uniform(0.02, 0.98)),
                            voxel=float(10.0 ** rng.uniform(-3.3, -2.0)), noise=float(10.0 ** rng.uniform(-4.5, -3.0)), max_iterations=H)
        a = ctx.ransac_sharded(c.source, c.target, c.source_desc, c.target_desc, c.voxel_size, H, conf)
        b = lone.ransac(c.source, c.target, c.source_desc, c.target_desc, c.voxel_size, H, conf)
        same = np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32)) and np.float32(a[1]).tobytes() == np.float32(b[1]).tobytes() \
            and np.float32(a[2]).tobytes() == np.float32(b[2]).tobytes() and a[3] == b[3]
        if not same:
            bad += 1
            print(f"MISMATCH rank {rank} seed {s}: src {n_src} tgt {n_tgt} H {H} conf {conf}: sharded {a[1:]} lone {b[1:]}", flush=True)
    print(f"rank {rank}/{world}: {cases} cases, {bad} mismatches", flush=True)
    ctx.comm_destroy(); ctx.close(); lone.close()
    sys.exit(1 if bad else 0)


def main():
    if sys.argv[1] == "--rank":
        rank_main(int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], int(sys.argv[5]), int(sys.argv[6]))
        return
    world = int(sys.argv[1])
    cases = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    seed0 = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    with tempfile.TemporaryDirectory() as work:
        procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--rank", str(r), str(world), work, str(cases), str(seed0)]) for r in range(world)]
        rc = 0
        for p in procs:
            try:
                rc |= p.wait(timeout=1500)
            except subprocess.TimeoutExpired:
                p.kill(); rc |= 1
    sys.exit(rc)


if __name__ == "__main__":
    main()
