"""BASELINE.json's full-size configurations on the GPU.  The CPU oracle cannot run these sizes in a unit
test, so parity is checked (a) bit-exactly on bounded slices the oracle can afford (rows, hypothesis-id
ranges, source-point prefixes — every figure is independent per row / hypothesis / point), (b) between
independent GPU kernels that share no screening arithmetic, and (c) through size-independent properties
(determinism, permutation invariance, known ground truth)."""
import hashlib
import importlib
import json
import os
import threading

import numpy as np
import pytest

syn = importlib.import_module("3dvision_b200.synthetic")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c3():
    return syn.ransac_case()          # configs[2]: 100k x 100k descriptors, H = 1M


def test_c3_matching_full_size(ctx, oracle, c3):
    ctx.set_clouds(c3.source, c3.target); ctx.set_features(c3.source_desc, c3.target_desc)
    ctx.set_match_mode(2); ctx.match_features()
    corr = ctx.get_correspondences().copy()
    ctx.match_features()
    assert np.array_equal(corr, ctx.get_correspondences())                 # deterministic / idempotent
    rows = np.random.default_rng(0).choice(c3.source.shape[0], 160, replace=False)
    for r in rows:                                                          # oracle on a bounded slice of rows
        assert corr[r] == oracle.match_features(c3.source_desc, c3.target_desc, int(r), int(r) + 1)[0]
    ctx.set_match_mode(1)
    ctx.set_correspondences(np.zeros(c3.source.shape[0], np.uint32))
    ctx.match_features(70_000, 73_000)                                      # exact CUDA-core kernel on a row range
    assert np.array_equal(ctx.get_correspondences()[70_000:73_000], corr[70_000:73_000])
    ctx.set_match_mode(0)
    inl = c3.true_match >= 0
    assert (corr[inl] == c3.true_match[inl]).mean() > 0.995


def test_c3_scoring_full_size(ctx, oracle, c3):
    """1M hypotheses: RNG stream / Lemire compaction / Kabsch / counts checked against the oracle on id ranges at the
    start, middle and very end of the stream; packed-FFMA2 screen vs un-fused kernel on 60k hypotheses."""
    corr = np.where(c3.true_match >= 0, c3.true_match, 0).astype(np.uint32)
    H = c3.max_iterations
    ctx.set_clouds(c3.source, c3.target); ctx.set_correspondences(corr)
    ctx.ransac_prepare(c3.voxel_size, H, 2.0)
    for lo in (0, 500_000, H - 48):
        hi = lo + 48
        ctx.set_score_mode(0); ctx.ransac_score(lo, hi)
        got = ctx.ransac_counts(lo, hi)
        ref = oracle.ransac(c3.source, c3.target, corr, c3.voxel_size, H, 2.0, iter_lo=lo, iter_hi=hi, want_counts=True)
        assert np.array_equal(got, ref.extra["counts"][lo:hi]), f"counts differ in [{lo},{hi})"
    ctx.set_score_mode(0); ctx.ransac_score(200_000, 260_000); a = ctx.ransac_counts(200_000, 260_000)
    assert ctx.score_recounts() > 0                                          # the band path was exercised
    ctx.set_score_mode(2); ctx.ransac_score(200_000, 260_000); b = ctx.ransac_counts(200_000, 260_000)
    ctx.set_score_mode(1); ctx.ransac_score(200_000, 260_000); c = ctx.ransac_counts(200_000, 260_000)
    ctx.set_score_mode(0)
    assert np.array_equal(a, c) and np.array_equal(b, c)
    assert c.max() > 0.5 * c3.source.shape[0]                                # good hypotheses exist in that range


def test_c3_near_threshold_pairs_stress_the_band(ctx, oracle):
    """Noise comparable to the inlier threshold puts many pairs inside the screening band; counts must stay exact."""
    c = syn.ransac_case(n_src=20_000, n_tgt=15_000, seed=55, noise=0.0009, max_iterations=3000)
    corr = np.where(c.true_match >= 0, c.true_match, 0).astype(np.uint32)
    ctx.set_clouds(c.source, c.target); ctx.set_correspondences(corr)
    ctx.ransac_prepare(c.voxel_size, 3000, 2.0); ctx.ransac_score()
    ref = oracle.ransac(c.source, c.target, corr, c.voxel_size, 3000, 2.0, want_counts=True)
    assert np.array_equal(ctx.ransac_counts(), ref.extra["counts"])
    assert ctx.score_recounts() > 1000


@pytest.fixture(scope="module")
def c2():
    return syn.icp_case()             # configs[1]: 300k scene vs 100k model


def test_c2_icp_full_size(ctx, oracle, c2):
    ctx.set_clouds(c2.source, c2.target, c2.target_normals)
    idx, d2 = ctx.icp_nearest(c2.T_init, c2.threshold)
    n = 2500                                                                # oracle brute force on a prefix of the source
    ref = oracle.icp(c2.source[:n], c2.target, c2.target_normals, c2.T_init, c2.threshold, 1, True, want_nn0=True)
    kept = np.sqrt(ref.extra["nn_d2_0"]) <= np.float32(c2.threshold)
    assert kept.sum() > 2000
    assert np.array_equal(idx[:n][kept], ref.extra["nn_idx0"][kept]) and np.array_equal(d2[:n][kept], ref.extra["nn_d2_0"][kept])
    assert (idx[:n][~kept] == 0xFFFFFFFF).all()
    # default mode = the reference's summation order: deterministic, and equal to the one-chain replay (mode 3, an
    # independent kernel with one dependent add per matched point) after every one of the 50 iterations' worth of updates
    T, fit, rmse, it = ctx.icp_run(c2.T_init, c2.threshold, c2.iterations, True, False)
    T2, fit2, rmse2, _ = ctx.icp_run(c2.T_init, c2.threshold, c2.iterations, True, False)
    assert np.array_equal(T, T2) and fit == fit2 and rmse == rmse2
    assert it == c2.iterations and fit > 0.99
    assert syn.rotation_error(T, c2.T_true) < 3e-4 and syn.translation_error(T, c2.T_true) < 5e-5
    ctx.set_icp_mode(3)
    try:
        Tl, fitl, rmsel, _ = ctx.icp_run(c2.T_init, c2.threshold, 12, True, False)
    finally:
        ctx.set_icp_mode(0)
    Td, fitd, rmsed, _ = ctx.icp_run(c2.T_init, c2.threshold, 12, True, False)
    assert np.array_equal(Td, Tl) and fitd == fitl and rmsed == rmsel
    # the production call (convergence break on) stops where the fixed-length run says it should and returns that state
    Tp, fitp, rmsep, itp = ctx.icp_run(c2.T_init, c2.threshold, 200, True, True)
    assert 2 <= itp < 200 and syn.rotation_error(Tp, c2.T_true) < 3e-4
    Tq, _, rmseq, itq = ctx.icp_run(c2.T_init, c2.threshold, itp, True, False)
    assert itq == itp and np.array_equal(Tp, Tq) and rmsep == rmseq


def test_c2_icp_full_size_against_the_committed_oracle_iterations(ctx, c2):
    """tests/golden/c2_icp_fullsize.json holds the CPU oracle's transform / fitness / rmse after 1 and 2 iterations of
    configs[1] at FULL size (3e10 brute-force pair evaluations each — generated offline by make_golden_c2.py)."""
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "c2_icp_fullsize.json")))
    dig = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    assert dig(c2.source) == g["sha256"]["source"] and dig(c2.target) == g["sha256"]["target"]
    assert dig(c2.target_normals) == g["sha256"]["normals"] and dig(c2.T_init) == g["sha256"]["T_init"]
    ctx.set_clouds(c2.source, c2.target, c2.target_normals)
    for iters, want in sorted(g["after"].items()):
        T, fit, rmse, it = ctx.icp_run(c2.T_init, c2.threshold, int(iters), True, False)
        Tw = np.asarray(want["T"], np.float32).reshape(4, 4)
        assert it == want["iterations"] and fit == np.float32(want["fitness"])
        assert syn.rotation_error(T, Tw) < 1e-5 and syn.translation_error(T, Tw) < 1e-6
        assert np.array_equal(T, Tw) and rmse == np.float32(want["rmse"])


def test_c2_icp_fast_mode_is_order_free(ctx, c2):
    """b3d_set_icp_mode(1), fp64 tree sums: permuting the source does not move the result beyond fp64 rounding."""
    ctx.set_icp_mode(1)
    try:
        ctx.set_clouds(c2.source, c2.target, c2.target_normals)
        T, fit, rmse, it = ctx.icp_run(c2.T_init, c2.threshold, c2.iterations, True, False)
        assert it == c2.iterations and fit > 0.99 and syn.rotation_error(T, c2.T_true) < 3e-4
        perm = np.random.default_rng(1).permutation(c2.source.shape[0])
        ctx.set_clouds(c2.source[perm], c2.target, c2.target_normals)
        T3, fit3, _, _ = ctx.icp_run(c2.T_init, c2.threshold, c2.iterations, True, False)
        assert fit3 == fit and syn.rotation_error(T3, T) < 1e-6 and syn.translation_error(T3, T) < 1e-7
    finally:
        ctx.set_icp_mode(0)


def test_c2_target_order_only_relabels_indices(ctx, c2):
    ctx.set_clouds(c2.source, c2.target, c2.target_normals)
    idx, d2 = ctx.icp_nearest(c2.T_init, c2.threshold)
    tperm = np.random.default_rng(2).permutation(c2.target.shape[0])
    ctx.set_clouds(c2.source, c2.target[tperm], c2.target_normals[tperm])
    idx_p, d2_p = ctx.icp_nearest(c2.T_init, c2.threshold)
    assert np.array_equal(d2_p, d2)
    m = idx != 0xFFFFFFFF
    assert np.array_equal(c2.target[tperm][idx_p[m]], c2.target[idx[m]])


def test_c4_batched_instances_from_a_thread_pool(b3d, oracle):
    """configs[3] in miniature: independent instances registered concurrently from pool threads, each thread with its
    own context (the orchestrator's contract, pipeline.cpp:321-327). Every instance must equal the oracle."""
    n_inst, results, errors = 16, {}, []

    def work(i):
        try:
            c = syn.ransac_case(n_src=1200 + 37 * i, n_tgt=900 + 11 * i, seed=1000 + i, max_iterations=1200)
            src = b3d.PointCloud(points=c.source); tgt = b3d.PointCloud(points=c.target, normals=c.target_normals)
            coarse = b3d.Registration.ransacRegistration(src, tgt, b3d.FPFHFeatures(c.source_desc), b3d.FPFHFeatures(c.target_desc),
                                                         c.voxel_size, c.max_iterations)
            fine = b3d.GPURegistration.icpRefine(src, tgt, coarse.transformation, 0.004, 10)
            results[i] = (c, coarse, fine)
        except Exception as e:            # noqa: BLE001
            errors.append((i, repr(e)))

    threads = [threading.Thread(target=work, args=(i,)) for i in range(n_inst)]
    for k in range(0, n_inst, 8):
        for t in threads[k:k + 8]:
            t.start()
        for t in threads[k:k + 8]:
            t.join()
    assert not errors, errors
    for i, (c, coarse, fine) in results.items():
        ref = oracle.ransac_registration(c.source, c.target, c.source_desc, c.target_desc, c.voxel_size, c.max_iterations, 0.999)
        assert np.array_equal(coarse.transformation, ref.transformation) and coarse.fitness == ref.fitness and coarse.rmse == ref.rmse, i
        refi = oracle.icp(c.source, c.target, c.target_normals, ref.transformation, 0.004, 10, True)
        assert fine.fitness == refi.fitness, i
        assert syn.rotation_error(fine.transformation, refi.transformation) < 1e-5
        assert syn.translation_error(fine.transformation, refi.transformation) < 1e-6


def test_c5_large_source_cloud(ctx, oracle):
    """configs[4] scale on the hot path: 2M-point source (30 % outliers) against a 200k-point model."""
    rng = np.random.default_rng(77)
    model, nrm = syn.torus(200_000, rng, R=0.3, r=0.1)
    n = 2_000_000
    pick = rng.integers(0, model.shape[0], n)
    T_true = syn.rigid([0.5, 0.1, 0.8], 12.0, [0.02, 0.01, -0.03])
    src = syn.apply(np.linalg.inv(T_true), model[pick] + rng.normal(0, 2e-4, (n, 3)).astype(np.float32))
    out = rng.random(n) < 0.3
    src[out] = rng.uniform(src.min(0), src.max(0), (int(out.sum()), 3)).astype(np.float32)
    T0 = (syn.rigid([0.1, 0.9, 0.2], 0.3, [0.0008, -0.0005, 0.0006]) @ T_true).astype(np.float32)
    ctx.set_clouds(src, model, nrm)
    idx, d2 = ctx.icp_nearest(T0, 0.003)
    T, fit, rmse, it = ctx.icp_run(T0, 0.003, 12, True, True)
    assert 0.6 < fit < 0.8 and it >= 2
    assert syn.rotation_error(T, T_true) < 5e-4 and syn.translation_error(T, T_true) < 1e-4
    k = 1500                                                                 # oracle on a prefix: exact NN + exact inlier counts
    ref = oracle.icp(src[:k], model, nrm, T0, 0.003, 1, True, want_nn0=True)
    kept = np.sqrt(ref.extra["nn_d2_0"]) <= np.float32(0.003)
    assert np.array_equal(idx[:k][kept], ref.extra["nn_idx0"][kept]) and (idx[:k][~kept] == 0xFFFFFFFF).all()
    corr = np.where(out, 0, pick).astype(np.uint32)
    ctx.set_correspondences(corr)
    ctx.ransac_prepare(0.001, 600, 2.0); ctx.ransac_score()
    got = ctx.ransac_counts()
    refr = oracle.ransac(src, model, corr, 0.001, 600, 2.0, iter_lo=0, iter_hi=24, want_counts=True)
    assert np.array_equal(got[:24], refr.extra["counts"][:24])


def _staged_result(ctx, mode, voxel, H, confidence):
    import torch
    ctx.set_score_mode(mode)
    try:
        ctx.ransac_prepare(voxel, H, confidence)
        ctx.ransac_score()
        keys = torch.zeros(2, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        ctx.ransac_reduce(0, H, keys.data_ptr())
        return ctx.ransac_finish(keys.data_ptr()), ctx.ransac_counts()
    finally:
        ctx.set_score_mode(0)


@pytest.mark.parametrize("inlier_frac,confidence,H", [(0.7, 0.999, 20_000), (0.7, 0.30, 20_000), (0.15, 0.999, 12_000), (0.02, 0.999, 9_000)])
def test_bailout_scoring_keeps_the_exact_winner(ctx, oracle, inlier_frac, confidence, H):
    """score mode 3 drops hypotheses that provably cannot reach the best full count; the winner, its transform,
    fitness, rmse and the early-exit behaviour must be identical to full scoring (and so to the oracle)."""
    c = syn.ransac_case(n_src=30_000, n_tgt=20_000, seed=91, inlier_frac=inlier_frac, max_iterations=H)
    corr = np.where(c.true_match >= 0, c.true_match, 0).astype(np.uint32)
    ctx.set_clouds(c.source, c.target); ctx.set_correspondences(corr)
    (T0, f0, r0, b0), counts_full = _staged_result(ctx, 0, c.voxel_size, H, confidence)
    (T3, f3, r3, b3), counts_bail = _staged_result(ctx, 3, c.voxel_size, H, confidence)
    assert b3 == b0 and np.array_equal(T3, T0) and f3 == f0 and r3 == r0
    kept = counts_bail != -4
    assert np.array_equal(counts_bail[kept], counts_full[kept])             # survivors carry exact full counts
    assert counts_full[~kept].max(initial=-1) < counts_full.max()            # nothing that could win or tie was dropped
    ref = oracle.ransac(c.source, c.target, corr, c.voxel_size, H, confidence)
    assert b3 == ref.extra["best_iter"] and np.array_equal(T3, ref.transformation) and f3 == ref.fitness and r3 == ref.rmse
    if inlier_frac >= 0.5:
        assert (~kept).mean() > 0.3                                          # and it actually pruned


def test_c5_stress_depth_scene_front_to_back(b3d, oracle):
    """configs[4]: a ~10M-pixel depth-derived scene with 30 % outlier pixels through the whole resident path
    (deprojection -> down-sampling -> normals -> FPFH -> RANSAC -> ICP).  The oracle's O(N^2) stages cannot run here, so:
    (a) the deprojected cloud is compared with the oracle bit for bit (O(N)), (b) the device replay of the reference's
    unordered_map order is compared with the real container at ~0.5M voxels, (c) the one-call path must equal the staged
    path exactly and be deterministic, and ICP on the resident clouds must hold the known pose."""
    rng = np.random.default_rng(1234 + 4)
    h, w, f = 2560, 4096, 3000.0
    uu, vv = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32))
    zz = (0.9 + 0.10 * np.sin(uu / 310.0) * np.cos(vv / 270.0) + 0.05 * np.cos((uu + 2 * vv) / 190.0)
          + 0.03 * np.sin(uu / 67.0 + 1.0) * np.sin(vv / 83.0) + 0.012 * np.cos(uu / 23.0) * np.cos(vv / 29.0))
    depth = np.round(zz * 1000.0).astype(np.uint16)
    outl = rng.random((h, w)) < 0.30
    depth[outl] = rng.integers(300, 1500, int(outl.sum())).astype(np.uint16)          # 30 % outlier pixels
    voxel = 0.004
    args = (1000.0, 1.5, f, f, w / 2.0, h / 2.0)
    c = b3d.Context(0)
    try:
        cloud, _ = c.depth_to_cloud(depth, None, *args)
        want, _ = oracle.depth_to_cloud(depth, None, *args)
        assert cloud.shape[0] > 10_000_000 and np.array_equal(cloud.view(np.uint32), want.view(np.uint32))
        down_dev, _ = c.voxel_downsample(cloud, voxel)
        down_host = oracle.voxel_downsample(cloud, voxel)                            # the real std::unordered_map, on the CPU
        assert down_dev.shape[0] > 300_000 and np.array_equal(down_dev.view(np.uint32), down_host.view(np.uint32))
        # model = the clean surface seen from another pose
        clean = np.round(zz * 1000.0).astype(np.uint16)
        surf, _ = c.depth_to_cloud(clean[::3, ::3].copy(), None, 1000.0, 1.5, f / 3, f / 3, w / 6.0, h / 6.0)
        T_true = syn.rigid([0.3, 0.2, 0.9], 15.0, [0.04, -0.02, 0.05])                 # scene -> model
        model = syn.apply(T_true, surf)
        assert c.prepare_model(model, voxel) > 50_000
        a = c.register_depth(depth, None, *args, voxel, ransac_max_iterations=20000, icp_max_iterations=30)
        b = c.register_scene(cloud, voxel, ransac_max_iterations=20000, icp_max_iterations=30)
        a2 = c.register_depth(depth, None, *args, voxel, ransac_max_iterations=20000, icp_max_iterations=30)
        assert a["n_source_points"] == b["n_source_points"] == down_dev.shape[0]
        for x, y in ((a, b), (a, a2)):
            assert np.array_equal(x["coarse"][0], y["coarse"][0]) and np.array_equal(x["refined"][0], y["refined"][0])
            assert x["refined"][1:] == y["refined"][1:]
        # RANSAC on a smooth height field under 30 % scattered outliers is not expected to find the pose with 20 000
        # hypotheses; what must hold at this size is the geometry of the resident clouds: ICP started at the true pose stays there
        assert 0.0 <= a["refined"][1] <= 1.0 and np.isfinite(a["refined"][0]).all()
        T, fit, rmse, iters = c.icp_run(T_true.astype(np.float32), voxel * 1.5, 30, True, True)
        assert fit > 0.05 and syn.rotation_error(T, T_true) < 2e-3 and syn.translation_error(T, T_true) < 1e-3
    finally:
        c.close()
