#!/usr/bin/env python
"""Sensitivity of the oracle's outputs to the evaluation orders it ASSUMES of Eigen 3.4.

The reference cannot be compiled here (no Eigen), so every "bit-exact" claim is against the restatement.  This script
measures how much rides on each assumption: it rebuilds the restatement with ONE assumption swapped for the other
plausible reading (`make -C oracle variants`, ORC_* macros) and counts, over configs[0] and the seeded golden cases, how
many correspondence indices, per-hypothesis inlier counts, winner ids, normals, descriptors and final transforms change.
Run from the repo root:  python oracle/sensitivity.py [--hyp 20000]     (test infrastructure; CPU only; ~10 min).
The table it prints is the one in DESIGN.md §2.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

syn = importlib.import_module("3dvision_b200.synthetic")

VARIANTS = {
    "redux_left": "3-term redux (dot, norm, R*p, mean) as (a0+a1)+a2 instead of a0+(a1+a2)",
    "quat_scalar": "Quaternion product in the generic scalar order instead of the SSE quat_product order",
    "ldlt_linear": "LDLT triangular solves summed left to right instead of halving / SSE predux order",
    "mat4_pairwise": "4x4 delta*T coefficients as (a0b0+a1b1)+(a2b2+a3b3) instead of the sequential packet order",
}


def run_all(hyp):
    """Everything the parity tests compare, for whichever library is active."""
    out = {}
    voxel = 0.001
    scene = O.demo_scene_points()
    src = O.voxel_downsample(scene, voxel); src_n = O.estimate_normals(src, 30); src_f = O.compute_fpfh(src, src_n, voxel * 5.0)
    tgt = O.voxel_downsample(O.demo_model_points(), voxel); tgt_n = O.estimate_normals(tgt, 30); tgt_f = O.compute_fpfh(tgt, tgt_n, voxel * 5.0)
    corr = O.match_features(src_f, tgt_f)
    r = O.ransac(src, tgt, corr, voxel, hyp, 2.0, want_counts=True)
    icp = O.icp(src, tgt, tgt_n, r.transformation, float(np.float32(voxel) * np.float32(0.4)), 200, True)
    out["demo"] = dict(normals=np.concatenate([src_n, tgt_n]), fpfh=np.concatenate([src_f, tgt_f]), corr=corr, counts=r.extra["counts"],
                       winner=r.extra["best_iter"], T=r.transformation, icp_T=icp.transformation, icp_it=icp.extra["iters_run"])
    # a curved surface: normals / descriptors that are not degenerate
    pts = O.voxel_downsample(syn.torus(6000, np.random.default_rng(97))[0], 0.006)
    nrm = O.estimate_normals(pts, 30)
    out["torus"] = dict(normals=nrm, fpfh=O.compute_fpfh(pts, nrm, 0.03))
    for i, (n_src, n_tgt, seed) in enumerate(((3000, 2500, 7), (2977, 1842, 94), (20000, 3000, 95))):
        c = syn.ransac_case(n_src=n_src, n_tgt=n_tgt, seed=seed, max_iterations=4000)
        corr = O.match_features(c.source_desc, c.target_desc)
        r = O.ransac(c.source, c.target, corr, c.voxel_size, 4000, 2.0, want_counts=True)
        T0 = c.T_true.copy(); T0[:3, 3] += np.float32(2e-4)
        icp = O.icp(c.source, c.target, c.target_normals, T0, c.voxel_size * 0.4, 30, True)          # the noise-floor threshold
        p2p = O.icp(c.source, c.target, None, T0, 0.004, 15, False)
        out[f"seeded{i}"] = dict(corr=corr, counts=r.extra["counts"], winner=r.extra["best_iter"], T=r.transformation,
                                 icp_T=icp.transformation, icp_it=icp.extra["iters_run"], p2p_T=p2p.transformation, p2p_it=p2p.extra["iters_run"])
    ic = syn.icp_case(n_model=3000, n_scene=4000, seed=21)
    icp = O.icp(ic.source, ic.target, ic.target_normals, ic.T_init, ic.threshold, 30, True)
    out["icp_small"] = dict(icp_T=icp.transformation, icp_it=icp.extra["iters_run"])
    return out


def compare(base, var):
    """Totals over all cases: (changed, out of) per quantity, and the largest transform deviations."""
    tot = {k: [0, 0] for k in ("normals", "fpfh", "corr", "counts", "winner", "icp_it", "p2p_it")}
    dev = {"T": 0.0, "icp_T_rot": 0.0, "icp_T_trans": 0.0, "p2p_T_rot": 0.0, "p2p_T_trans": 0.0}
    for case in base:
        b, v = base[case], var[case]
        for k in tot:
            if k not in b:
                continue
            if k in ("winner", "icp_it", "p2p_it"):
                tot[k][0] += int(b[k] != v[k]); tot[k][1] += 1
            elif k in ("normals", "fpfh"):
                rows = (b[k].view(np.uint32) != v[k].view(np.uint32)).any(axis=1)
                tot[k][0] += int(rows.sum()); tot[k][1] += rows.size
            else:
                tot[k][0] += int((b[k] != v[k]).sum()); tot[k][1] += b[k].size
        if "T" in b:
            dev["T"] = max(dev["T"], float(np.abs(b["T"].astype(np.float64) - v["T"]).max()))
        for key in ("icp_T", "p2p_T"):
            if key in b:
                dev[key + "_rot"] = max(dev[key + "_rot"], syn.rotation_error(b[key], v[key]))
                dev[key + "_trans"] = max(dev[key + "_trans"], syn.translation_error(b[key], v[key]))
    return tot, dev


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hyp", type=int, default=20000, help="hypotheses scored on configs[0] (100000 in the reference's default)")
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    subprocess.run(["make", "-C", HERE, "all", "variants"], check=True, stdout=subprocess.DEVNULL)
    base = run_all(args.hyp)
    rows = []
    for name, what in VARIANTS.items():
        with O.use_library(os.path.join(HERE, "_build", f"liboracle_{name}.so")):
            var = run_all(args.hyp)
        tot, dev = compare(base, var)
        rows.append((name, what, tot, dev))
    hdr = ("| variant | normals rows | FPFH rows | correspondence indices | inlier counts | winner ids | RANSAC T max abs | "
           "ICP plane iters / rot / trans | ICP point iters / rot / trans |")
    print(hdr); print("|" + "---|" * 9)
    for name, what, t, d in rows:
        f = lambda k: f"{t[k][0]} / {t[k][1]}"
        print(f"| `{name}`: {what} | {f('normals')} | {f('fpfh')} | {f('corr')} | {f('counts')} | {f('winner')} | {d['T']:.1e} | "
              f"{f('icp_it')} / {d['icp_T_rot']:.1e} / {d['icp_T_trans']:.1e} | {f('p2p_it')} / {d['p2p_T_rot']:.1e} / {d['p2p_T_trans']:.1e} |")
    if args.json:
        json.dump([{"variant": n, "what": w, "changed": t, "deviation": d} for n, w, t, d in rows], open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
