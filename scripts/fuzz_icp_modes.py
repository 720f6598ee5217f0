"""Soak of the default ICP mode (parallel exact sums) against the one-chain replay (mode 3: one thread adds every term in
the reference's order) at sizes the CPU oracle is too slow for: random 20k-300k-point scenes, both error metrics.
The two modes share the search and the solve and differ only in how the sums are formed, so every output must be equal.
usage: python scripts/fuzz_icp_modes.py [cases] [seed0]"""
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
    bad = 0
    t0 = time.time()
    with b3d.Context(0) as ctx:
        for s in range(seed0, seed0 + cases):
            rng = np.random.default_rng(s)
            n_model, n_scene = int(rng.integers(5_000, 100_000)), int(rng.integers(20_000, 300_000))
            thr = float(10.0 ** rng.uniform(-3.5, -2.0))
            iters = int(rng.choice([2, 5, 15, 40]))
            plane = bool(rng.integers(0, 2))
            c = syn.icp_case(n_model=n_model, n_scene=n_scene, seed=s, noise=float(10.0 ** rng.uniform(-4.5, -2.8)), threshold=thr,
                             iterations=iters, init_angle_deg=float(rng.uniform(0.0, 3.0)), init_shift=float(rng.uniform(0.0, 0.004)))
            out = []
            for mode in (0, 3):
                ctx.set_icp_mode(mode)
                out.append(ctx.icp(c.source, c.target, c.target_normals, c.T_init, thr, iters, plane))
            ctx.set_icp_mode(0)
            a, b = out
            same = a[3] == b[3] and np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32)) and \
                np.float32(a[1]).view(np.uint32) == np.float32(b[1]).view(np.uint32) and np.float32(a[2]).view(np.uint32) == np.float32(b[2]).view(np.uint32)
            if not same:
                bad += 1
                print(f"MISMATCH seed {s}: model {n_model} scene {n_scene} thr {thr:.2e} iters {iters} plane {plane}: {a[1:]} vs {b[1:]}", flush=True)
    print(f"{cases} cases, {bad} mismatches, {time.time() - t0:.0f} s")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
