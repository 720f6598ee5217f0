"""Committed golden fixtures (tests/golden/*.json, made by tests/golden/make_golden.py).

CPU part: the oracle still reproduces every frozen answer (regression pin of the oracle).
GPU part: the CUDA path reproduces the same frozen answers through the C-ABI.
The fixtures freeze the ORACLE, not the real reference ("parity unpinned": the reference has no
golden vectors and cannot be built here)."""
import hashlib
import importlib
import json
import os

import numpy as np
import pytest

syn = importlib.import_module("3dvision_b200.synthetic")
HERE = os.path.dirname(os.path.abspath(__file__))
SEEDED = json.load(open(os.path.join(HERE, "golden", "seeded_cases.json")))["cases"]
DEMO = json.load(open(os.path.join(HERE, "golden", "demo_scene.json")))


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def T_of(g):
    return np.asarray(g, np.float32).reshape(4, 4)


RANSAC_CASES = [c for c in SEEDED if c["kind"] == "ransac"]
ICP_CASES = [c for c in SEEDED if c["kind"] == "icp"]


def ransac_inputs(g):
    c = syn.ransac_case(n_src=g["n_src"], n_tgt=g["n_tgt"], seed=g["seed"], max_iterations=g["H"])
    assert digest(c.source) == g["sha256"]["source"] and digest(c.source_desc) == g["sha256"]["source_desc"], \
        "synthetic generator drifted: regenerate the goldens"
    return c


def icp_inputs(g):
    c = syn.icp_case(n_model=g["n_model"], n_scene=g["n_scene"], seed=g["seed"])
    assert digest(c.source) == g["sha256"]["source"], "synthetic generator drifted: regenerate the goldens"
    return c


# ------------------------------------------------------------------------------------ CPU: oracle vs golden
@pytest.mark.parametrize("g", RANSAC_CASES, ids=lambda g: f"ransac-{g['n_src']}x{g['H']}")
def test_oracle_ransac_golden(oracle, g):
    c = ransac_inputs(g)
    corr = oracle.match_features(c.source_desc, c.target_desc)
    assert digest(corr) == g["sha256"]["correspondences"]
    r = oracle.ransac(c.source, c.target, corr, g["voxel"], g["H"], g["confidence"], want_counts=True)
    assert digest(r.extra["counts"]) == g["sha256"]["counts"]
    assert r.extra["best_iter"] == g["best_iteration"] and r.extra["iters_run"] == g["iterations_run"]
    assert np.array_equal(r.transformation, T_of(g["T"])) and r.fitness == np.float32(g["fitness"]) and r.rmse == np.float32(g["rmse"])


@pytest.mark.parametrize("g", ICP_CASES, ids=lambda g: f"icp-{g['n_scene']}-{'plane' if g['point_to_plane'] else 'point'}")
def test_oracle_icp_golden(oracle, g):
    c = icp_inputs(g)
    r = oracle.icp(c.source, c.target, c.target_normals, c.T_init, g["threshold"], g["max_iterations"], g["point_to_plane"], want_nn0=True)
    assert digest(r.extra["nn_idx0"]) == g["sha256"]["nn_idx0"] and digest(r.extra["nn_d2_0"]) == g["sha256"]["nn_d2_0"]
    assert r.extra["iters_run"] == g["iterations"]
    assert np.array_equal(r.transformation, T_of(g["T"])) and r.fitness == np.float32(g["fitness"])


@pytest.fixture(scope="module")
def demo(oracle):
    """BASELINE.json configs[0] inputs, rebuilt by the oracle's restatement of the pipeline's pre-stages (~10 s)."""
    voxel = DEMO["voxel"]
    src = oracle.voxel_downsample(oracle.demo_scene_points(), voxel)
    src_f = oracle.compute_fpfh(src, oracle.estimate_normals(src, 30), voxel * 5.0)
    tgt = oracle.voxel_downsample(oracle.demo_model_points(), voxel)
    tgt_n = oracle.estimate_normals(tgt, 30)
    tgt_f = oracle.compute_fpfh(tgt, tgt_n, voxel * 5.0)
    return dict(voxel=voxel, src=src, src_f=src_f, tgt=tgt, tgt_n=tgt_n, tgt_f=tgt_f)


def test_demo_scene_inputs_golden(oracle, demo):
    """Sizes computed in SURVEY.md Appendix E: 40 401 masked pixels -> 32 129 voxels; 1 600 model points."""
    assert demo["src"].shape[0] == DEMO["n_src"] == 32129 and demo["tgt"].shape[0] == DEMO["n_tgt"] == 1600
    for key, name in [("src", "src"), ("tgt", "tgt"), ("src_f", "src_fpfh"), ("tgt_f", "tgt_fpfh"), ("tgt_n", "tgt_normals")]:
        assert digest(demo[key]) == DEMO["sha256"][name], name
    corr = oracle.match_features(demo["src_f"], demo["tgt_f"])
    assert digest(corr) == DEMO["sha256"]["correspondences"]
    head = oracle.ransac(demo["src"], demo["tgt"], corr, demo["voxel"], 2000, 2.0, want_counts=True)
    assert digest(head.extra["counts"]) == DEMO["sha256"]["counts_first_2000_no_exit"]


# ------------------------------------------------------------------------------------ GPU: CUDA path vs golden
@pytest.mark.gpu
@pytest.mark.parametrize("g", RANSAC_CASES, ids=lambda g: f"ransac-{g['n_src']}x{g['H']}")
def test_cuda_ransac_golden(ctx, g):
    c = ransac_inputs(g)
    ctx.set_clouds(c.source, c.target); ctx.set_features(c.source_desc, c.target_desc)
    ctx.match_features()
    assert digest(ctx.get_correspondences()) == g["sha256"]["correspondences"]
    T, fit, rmse, best = ctx.ransac(c.source, c.target, c.source_desc, c.target_desc, g["voxel"], g["H"], g["confidence"])
    assert best == g["best_iteration"]
    assert np.array_equal(T, T_of(g["T"])) and fit == np.float32(g["fitness"]) and rmse == np.float32(g["rmse"])
    if g["iterations_run"] == g["H"]:                     # no early exit: every hypothesis has a reference count
        ctx.ransac_prepare(g["voxel"], g["H"], g["confidence"]); ctx.ransac_score()
        assert digest(ctx.ransac_counts()) == g["sha256"]["counts"]


@pytest.mark.gpu
@pytest.mark.parametrize("g", [c for c in ICP_CASES if c["point_to_plane"]], ids=lambda g: f"icp-{g['n_scene']}")
def test_cuda_icp_golden(ctx, g):
    c = icp_inputs(g)
    T, fit, rmse, iters = ctx.icp(c.source, c.target, c.target_normals, c.T_init, g["threshold"], g["max_iterations"], True)
    assert iters == g["iterations"] and fit == np.float32(g["fitness"])
    assert syn.rotation_error(T, T_of(g["T"])) < 1e-5 and syn.translation_error(T, T_of(g["T"])) < 1e-6
    assert np.array_equal(T, T_of(g["T"])) and rmse == np.float32(g["rmse"])      # default mode adds in the reference's order


@pytest.mark.gpu
def test_cuda_demo_scene_golden(ctx, demo):
    """configs[0] end to end on the GPU: degenerate planar scene (two distinct descriptors, every row ties)."""
    d = demo
    T, fit, rmse, best = ctx.ransac(d["src"], d["tgt"], d["src_f"], d["tgt_f"], d["voxel"], DEMO["ransac_max_iterations"], DEMO["confidence"])
    assert digest(ctx.get_correspondences()) == DEMO["sha256"]["correspondences"]
    assert best == DEMO["ransac"]["best_iteration"]
    assert np.array_equal(T, T_of(DEMO["ransac"]["T"]))
    assert fit == np.float32(DEMO["ransac"]["fitness"]) and rmse == np.float32(DEMO["ransac"]["rmse"])
    ctx.set_clouds(d["src"], d["tgt"]); ctx.set_correspondences(ctx.get_correspondences())
    ctx.ransac_prepare(d["voxel"], DEMO["ransac_max_iterations"], DEMO["confidence"]); ctx.ransac_score()
    assert digest(ctx.ransac_counts()) == DEMO["sha256"]["counts"]
    Ti, fi, ri, iters = ctx.icp(d["src"], d["tgt"], d["tgt_n"], T, DEMO["icp_threshold"], 200, True)
    # the orchestrator's own ICP call: threshold 0.4 * voxel (pipeline.cpp:104), default mode
    assert iters == DEMO["icp"]["iterations"] and fi == np.float32(DEMO["icp"]["fitness"])
    assert syn.rotation_error(Ti, T_of(DEMO["icp"]["T"])) < 1e-5 and syn.translation_error(Ti, T_of(DEMO["icp"]["T"])) < 1e-6
    assert np.array_equal(Ti, T_of(DEMO["icp"]["T"])) and ri == np.float32(DEMO["icp"]["rmse"])
