"""Randomised soak of b3d_sequential_sum (the device's exact-sum passes) against in-order fp32 addition.
usage: python scripts/fuzz_exact_sums.py [cases] [seed0]"""
import importlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
b3d = importlib.import_module("3dvision_b200._capi")
syn = importlib.import_module("3dvision_b200.synthetic")


def seq(x):
    with np.errstate(all="ignore"):
        return np.add.accumulate(x, dtype=np.float32)[-1] if x.size else np.float32(0)


def case(rng, seed=0):
    return syn.sparse_mixed_terms(rng) if seed % 2 else syn.adversarial_terms(rng)


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    bad = 0
    t0 = time.time()
    total = 0
    with b3d.Context(0) as ctx:
        for s in range(seed0, seed0 + cases):
            x = case(np.random.default_rng(s), s)
            got, _ = ctx.sequential_sum(x)
            ref = seq(x)
            total += x.size
            if not (np.float32(got).view(np.uint32) == np.float32(ref).view(np.uint32) or (np.isnan(got) and np.isnan(ref))):
                bad += 1
                print("MISMATCH seed", s, "n", x.size, got, ref, flush=True)
    print(f"{cases} cases, {total} terms, {bad} mismatches, {time.time() - t0:.1f} s")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
