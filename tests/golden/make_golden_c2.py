#!/usr/bin/env python
"""tests/golden/c2_icp_fullsize.json: the CPU oracle on BASELINE.json configs[1] AT FULL SIZE (300 000-point scene against
the 100 000-point model, point-to-plane) for the first iterations — brute-force nearest neighbour, 3e10 pair evaluations
per iteration, minutes on one core, which is why it is a committed fixture and not a test-time computation.
Run from the repo root:  python tests/golden/make_golden_c2.py   (inputs are regenerated from the seed; only digests,
transforms, fitness and rmse are stored).  Frozen ORACLE answers: pins the CUDA path at full size to the restatement."""
import hashlib
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

syn = importlib.import_module("3dvision_b200.synthetic")


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    c = syn.icp_case()
    out = {"config": "configs[1] synthetic 100k-point model vs 300k-point noisy scene, point-to-plane",
           "n_scene": int(c.source.shape[0]), "n_model": int(c.target.shape[0]), "threshold": float(c.threshold),
           "sha256": {"source": digest(c.source), "target": digest(c.target), "normals": digest(c.target_normals),
                      "T_init": digest(c.T_init)},
           "after": {}}
    for iters in (1, 2):
        t0 = time.time()
        r = O.icp(c.source, c.target, c.target_normals, c.T_init, c.threshold, iters, True)
        out["after"][str(iters)] = {"T": [float(x) for x in np.asarray(r.transformation, np.float32).reshape(-1)],
                                    "fitness": float(r.fitness), "rmse": float(r.rmse), "iterations": int(r.extra["iters_run"]),
                                    "oracle_seconds": round(time.time() - t0, 1)}
        print(iters, out["after"][str(iters)], flush=True)
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "c2_icp_fullsize.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
