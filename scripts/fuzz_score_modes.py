"""Soak of the RANSAC scoring kernels against each other at sizes the CPU oracle is too slow for: the packed-FMA screen
with exact re-count (modes 0 / 4), the scalar screen (2) and the bail-out scorer (3) against mode 1, which evaluates
every pair with the reference's un-fused arithmetic (src/registration.cpp:270-279).  Per-hypothesis inlier counts must be
identical (modes 0, 2, 4); the result (pose, fitness, rmse, winning iteration) must be identical in every mode.
usage: python scripts/fuzz_score_modes.py [cases] [seed0]"""
import importlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
b3d = importlib.import_module("3dvision_b200._capi")
syn = importlib.import_module("3dvision_b200.synthetic")


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    import torch
    bad = 0
    t0 = time.time()
    with b3d.Context(0) as ctx:
        keys = torch.zeros(4, dtype=torch.int64, device="cuda")
        for s in range(seed0, seed0 + cases):
            rng = np.random.default_rng(s)
            n_src, n_tgt = int(rng.integers(500, 60_000)), int(rng.integers(500, 20_000))
            H = int(rng.integers(100, 60_000))
            conf = float(rng.choice([0.3, 0.999, 2.0]))
            c = syn.ransac_case(n_src=n_src, n_tgt=n_tgt, seed=s, inlier_frac=float(rng.uniform(0.02, 0.98)),
                                voxel=float(10.0 ** rng.uniform(-3.5, -1.5)), noise=float(10.0 ** rng.uniform(-4.5, -2.5)), max_iterations=H)
            corr = np.where(c.true_match >= 0, c.true_match, rng.integers(0, n_tgt, n_src)).astype(np.uint32)
            ctx.set_clouds(c.source, c.target); ctx.set_correspondences(corr)
            ref_counts = ref_res = None
            for mode in (1, 0, 2, 4, 3):
                ctx.set_score_mode(mode)
                ctx.ransac_prepare(c.voxel_size, H, conf)
                ctx.ransac_score(0, H)
                ctx.ransac_reduce3(0, H, keys.data_ptr())
                torch.cuda.synchronize()
                from_keys = importlib.import_module("3dvision_b200.dist").resolve_keys([tuple(int(v) for v in keys[:3].cpu().numpy())])
                fin = torch.tensor([from_keys, 0], dtype=torch.int64, device="cuda")
                T, fit, rmse, hid = ctx.ransac_finish(fin.data_ptr())
                res = (T.view(np.uint32).tobytes(), np.float32(fit).tobytes(), np.float32(rmse).tobytes(), hid)
                counts = ctx.ransac_counts(0, H) if mode != 3 else None
                if mode == 1:
                    ref_counts, ref_res = counts, res
                    continue
                if res != ref_res or (counts is not None and not np.array_equal(np.where(ref_counts == -2, counts, ref_counts), counts)):
                    bad += 1
                    print(f"MISMATCH seed {s} mode {mode}: src {n_src} H {H} conf {conf} voxel {c.voxel_size:.2e}: winner {hid} vs {ref_res[3]}", flush=True)
            ctx.set_score_mode(0)
    print(f"{cases} cases x 4 modes, {bad} mismatches, {time.time() - t0:.0f} s")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
