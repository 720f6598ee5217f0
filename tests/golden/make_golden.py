#!/usr/bin/env python
"""Generates tests/golden/*.json from the CPU oracle (run from the repo root; takes ~3 min).

The reference ships no golden vectors for this path, and it cannot be compiled
or imported here (C++/Eigen, no Eigen in the image), so these fixtures freeze the ORACLE's
answers — they pin the oracle against regressions and give the CUDA tests fixed targets, they
do not pin the oracle against the real reference ("parity unpinned", oracle header).

Inputs are never stored: every case is regenerated from a seed (3dvision_b200/synthetic.py)
or from the deterministic demo-scene builders (oracle/pipeline_inputs.cpp).  Only digests and
the small results are committed.
"""
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
OUT = os.path.dirname(os.path.abspath(__file__))


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def T_list(T):
    return [float(x) for x in np.asarray(T, np.float32).reshape(-1)]


def demo_inputs():
    """BASELINE.json configs[0]: procedural box scene, voxel 0.001 (pipeline.cpp:211-257, 275-294, 92-95)."""
    voxel = 0.001
    scene = O.demo_scene_points()
    src = O.voxel_downsample(scene, voxel)
    src_n = O.estimate_normals(src, 30)
    src_f = O.compute_fpfh(src, src_n, voxel * 5.0)
    model = O.demo_model_points()
    tgt = O.voxel_downsample(model, voxel)
    tgt_n = O.estimate_normals(tgt, 30)
    tgt_f = O.compute_fpfh(tgt, tgt_n, voxel * 5.0)
    return dict(voxel=voxel, src=src, src_f=src_f, tgt=tgt, tgt_n=tgt_n, tgt_f=tgt_f, n_raw=int(scene.shape[0]))


def make_demo():
    t0 = time.time()
    d = demo_inputs()
    corr = O.match_features(d["src_f"], d["tgt_f"])
    H = 100000
    r = O.ransac(d["src"], d["tgt"], corr, d["voxel"], H, 0.999, want_counts=True)
    icp_thr = np.float32(d["voxel"]) * np.float32(0.4)
    icp = O.icp(d["src"], d["tgt"], d["tgt_n"], r.transformation, float(icp_thr), 200, True)
    head = O.ransac(d["src"], d["tgt"], corr, d["voxel"], 2000, 2.0, want_counts=True)
    g = {
        "config": "configs[0] demo procedural box scene, CPU path",
        "n_raw_points": d["n_raw"], "n_src": int(d["src"].shape[0]), "n_tgt": int(d["tgt"].shape[0]),
        "voxel": d["voxel"], "ransac_max_iterations": H, "confidence": 0.999, "icp_threshold": float(icp_thr),
        "sha256": {"src": digest(d["src"]), "tgt": digest(d["tgt"]), "src_fpfh": digest(d["src_f"]),
                   "tgt_fpfh": digest(d["tgt_f"]), "tgt_normals": digest(d["tgt_n"]),
                   "correspondences": digest(corr), "counts": digest(r.extra["counts"]),
                   "counts_first_2000_no_exit": digest(head.extra["counts"])},
        "ransac": {"T": T_list(r.transformation), "fitness": r.fitness, "rmse": r.rmse,
                   "best_iteration": r.extra["best_iter"], "iterations_run": r.extra["iters_run"]},
        "icp": {"T": T_list(icp.transformation), "fitness": icp.fitness, "rmse": icp.rmse,
                "iterations": icp.extra["iters_run"], "ncorr_first": [int(x) for x in icp.extra["ncorr"][:8]]},
        "seconds": round(time.time() - t0, 1),
    }
    json.dump(g, open(os.path.join(OUT, "demo_scene.json"), "w"), indent=1)
    print("demo_scene.json", g["ransac"], g["icp"]["fitness"], g["icp"]["iterations"], g["seconds"], "s")


def make_seeded():
    cases = []
    for (n_src, n_tgt, H, seed, conf) in [(3000, 2500, 4000, 7, 0.999), (257, 2500, 1500, 257, 2.0), (40, 300, 600, 40, 2.0)]:
        c = syn.ransac_case(n_src=n_src, n_tgt=n_tgt, seed=seed, max_iterations=H)
        corr = O.match_features(c.source_desc, c.target_desc)
        r = O.ransac(c.source, c.target, corr, c.voxel_size, H, conf, want_counts=True)
        cases.append({"kind": "ransac", "n_src": n_src, "n_tgt": n_tgt, "H": H, "seed": seed, "confidence": conf,
                      "voxel": c.voxel_size,
                      "sha256": {"source": digest(c.source), "source_desc": digest(c.source_desc),
                                 "correspondences": digest(corr), "counts": digest(r.extra["counts"])},
                      "T": T_list(r.transformation), "fitness": r.fitness, "rmse": r.rmse,
                      "best_iteration": r.extra["best_iter"], "iterations_run": r.extra["iters_run"]})
    for (n_model, n_scene, seed, plane, iters) in [(6000, 9000, 21, True, 30), (6000, 9000, 21, False, 30), (1500, 2000, 5, True, 12)]:
        c = syn.icp_case(n_model=n_model, n_scene=n_scene, seed=seed)
        r = O.icp(c.source, c.target, c.target_normals, c.T_init, c.threshold, iters, plane, want_nn0=True)
        cases.append({"kind": "icp", "n_model": n_model, "n_scene": n_scene, "seed": seed, "point_to_plane": plane,
                      "max_iterations": iters, "threshold": c.threshold,
                      "sha256": {"source": digest(c.source), "nn_idx0": digest(r.extra["nn_idx0"]), "nn_d2_0": digest(r.extra["nn_d2_0"])},
                      "T": T_list(r.transformation), "fitness": r.fitness, "rmse": r.rmse, "iterations": r.extra["iters_run"],
                      "ncorr": [int(x) for x in r.extra["ncorr"][:r.extra["iters_run"] + 1]]})
    json.dump({"cases": cases}, open(os.path.join(OUT, "seeded_cases.json"), "w"), indent=1)
    print("seeded_cases.json", len(cases), "cases")


if __name__ == "__main__":
    make_seeded()
    make_demo()
