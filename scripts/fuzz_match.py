"""Randomised soak of the tensor-core matcher (screen GEMM + exact re-score, csrc/b3d_match_tc.cu) and the CUDA-core matcher
against the oracle's nearest-descriptor search (src/registration.cpp:216-232): random sizes and descriptor populations
(histograms of random sparsity, near-duplicates, exact duplicates -> lowest index must win, scaled outliers, zero rows).
usage: python scripts/fuzz_match.py [cases] [seed0] [large_cases]"""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
b3d = importlib.import_module("3dvision_b200._capi")
syn = importlib.import_module("3dvision_b200.synthetic")
from oracle import oracle  # noqa: E402  (test infrastructure: the checker)


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    bad = 0
    t0 = time.time()
    with b3d.Context(0) as ctx:
        for s in range(seed0, seed0 + cases):
            sd, td = syn.random_descriptors(s)
            want = oracle.match_features(sd, td)
            pts_s = np.zeros((sd.shape[0], 3), np.float32); pts_t = np.zeros((td.shape[0], 3), np.float32)
            ctx.set_clouds(pts_s, pts_t); ctx.set_features(sd, td)
            for mode in (2, 1):
                ctx.set_match_mode(mode)
                ctx.match_features()
                got = ctx.get_correspondences()
                if not np.array_equal(got, want):
                    bad += 1
                    w = np.flatnonzero(got != want)
                    print(f"MISMATCH seed {s} mode {mode}: {sd.shape[0]} x {td.shape[0]}, {w.size} rows, first row {w[0]}: got {got[w[0]]} want {want[w[0]]}", flush=True)
            ctx.set_match_mode(0)
        # large sets, beyond what the CPU oracle finishes quickly: the two device matchers (independent kernels) against each other
        n_large = int(sys.argv[3]) if len(sys.argv) > 3 else 0
        for s in range(seed0, seed0 + n_large):
            rng = np.random.default_rng(10_000_000 + s)
            ns, nt = int(rng.integers(3000, 60_000)), int(rng.integers(3000, 60_000))
            td = syn.histograms(nt, rng, sparsity=float(rng.uniform(0.0, 0.9)))
            sd = syn.histograms(ns, rng, sparsity=float(rng.uniform(0.0, 0.9)))
            near = rng.random(ns) < rng.uniform(0.0, 1.0)
            sd[near] = np.abs(td[rng.integers(0, nt, int(near.sum()))] + rng.normal(0, 10.0 ** rng.uniform(-7, -2), (int(near.sum()), 33))).astype(np.float32)
            if rng.random() < 0.5:
                td[rng.integers(0, nt, nt // 20)] = td[rng.integers(0, nt, nt // 20)]          # exact duplicates: lowest index wins
            if rng.random() < 0.3:
                td[rng.integers(0, nt, 20)] *= np.float32(rng.uniform(3.0, 300.0))
            ctx.set_clouds(np.zeros((ns, 3), np.float32), np.zeros((nt, 3), np.float32)); ctx.set_features(sd, td)
            out = []
            for mode in (2, 1):
                ctx.set_match_mode(mode); ctx.match_features(); out.append(ctx.get_correspondences())
            ctx.set_match_mode(0)
            if not np.array_equal(out[0], out[1]):
                bad += 1
                print(f"MISMATCH large seed {s}: {ns} x {nt}: {int((out[0] != out[1]).sum())} rows", flush=True)
    print(f"{cases} cases x 2 matchers vs oracle + {n_large} large cases matcher vs matcher, {bad} mismatches, {time.time() - t0:.0f} s")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
