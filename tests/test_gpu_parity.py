"""Parity of the CUDA path (through the C-ABI) against the CPU oracle on seeded inputs.

Bars (north_star): bit-exact correspondence indices, inlier counts and NN indices;
final transforms within 1e-5 (rotation, Frobenius) and 1e-6 m (translation).
Caveat carried by every parity claim: the oracle restates Eigen 3.4 arithmetic and is
itself "parity unpinned" (oracle/registration_oracle.cpp header).
"""
import importlib

import numpy as np
import pytest

syn = importlib.import_module("3dvision_b200.synthetic")

pytestmark = pytest.mark.gpu

ROT_TOL = 1e-5      # Frobenius norm of the rotation-block difference
TRANS_TOL = 1e-6    # metres
# Point-to-point with the fp64 tree sums (b3d_set_icp_mode 1) only.  The reference accumulates the centroids and
# the cross-covariance sequentially in fp32 (registration.cpp:374-386); over a few thousand points that sum carries
# ~1e-6 m of rounding noise per iteration *in the reference itself*, and point-to-point ICP converges slowly along
# the surface, so the noise is amplified to ~1e-4 by the time the |d rmse| < 1e-6 stop fires.  The DEFAULT
# path (both error metrics) produces the reference's sequential sums bit for bit (b3d_ess.cuh) and is bit-identical
# to the oracle (tested below); mode 3 is the older one-chain replay, kept as an independent cross-check.
P2P_ROT_TOL = 5e-4
P2P_TRANS_TOL = 2e-4


def small_ransac_case(n_src=3000, n_tgt=2500, H=4000, seed=7, **kw):
    return syn.ransac_case(n_src=n_src, n_tgt=n_tgt, seed=seed, max_iterations=H, **kw)


# ------------------------------------------------------------------ feature matching
@pytest.mark.parametrize("n_src,n_tgt", [(1, 1), (5, 3), (64, 64), (65, 63), (700, 1000), (3000, 2500)])
def test_match_indices_bit_exact(ctx, oracle, n_src, n_tgt):
    rng = np.random.default_rng(n_src * 1000 + n_tgt)
    sd = syn.histograms(n_src, rng); td = syn.histograms(n_tgt, rng)
    pts_s = rng.random((n_src, 3), dtype=np.float32); pts_t = rng.random((n_tgt, 3), dtype=np.float32)
    ctx.set_clouds(pts_s, pts_t)
    ctx.set_features(sd, td)
    ctx.match_features()
    got = ctx.get_correspondences()
    want = oracle.match_features(sd, td)
    assert np.array_equal(got, want)


def test_match_ties_pick_lowest_index(ctx, oracle):
    """Planar scenes give bit-identical descriptors: the lowest target index must win."""
    rng = np.random.default_rng(3)
    base = syn.histograms(7, rng)
    td = base[rng.integers(0, 7, 500)]              # many exact duplicates
    sd = base[rng.integers(0, 7, 300)]
    pts = rng.random((500, 3), dtype=np.float32)
    ctx.set_clouds(pts[:300], pts)
    ctx.set_features(sd, td)
    ctx.match_features()
    got = ctx.get_correspondences()
    want = oracle.match_features(sd, td)
    assert np.array_equal(got, want)
    for i in range(300):                            # independent check of the tie rule
        same = np.flatnonzero((td == sd[i]).all(1))
        assert got[i] == same[0]


def test_match_row_range(ctx, oracle):
    rng = np.random.default_rng(11)
    sd = syn.histograms(500, rng); td = syn.histograms(400, rng)
    pts = rng.random((500, 3), dtype=np.float32)
    ctx.set_clouds(pts, pts[:400]); ctx.set_features(sd, td)
    ctx.match_features(0, 500)
    full = ctx.get_correspondences().copy()
    ctx.set_correspondences(np.zeros(500, np.uint32))
    ctx.match_features(130, 387)
    part = ctx.get_correspondences()
    assert np.array_equal(part[130:387], full[130:387])
    assert not part[:130].any() and not part[387:].any()


# ------------------------------------------------------------------ RANSAC
def test_rng_triples_and_hypotheses_match_oracle(ctx, oracle):
    case = small_ransac_case()
    corr = oracle.match_features(case.source_desc, case.target_desc)
    ctx.set_clouds(case.source, case.target)
    ctx.set_correspondences(corr)
    ctx.ransac_prepare(case.voxel_size, case.max_iterations, 2.0)
    hyp = ctx.ransac_hypotheses()
    for it in list(range(0, 40)) + [999, 2500, case.max_iterations - 1]:
        ok, R, t, _ = oracle.ransac_hypothesis(case.source, case.target, corr, it)
        if ok:
            assert np.array_equal(hyp[it, :9].reshape(3, 3), R), f"R differs at hypothesis {it}"
            assert np.array_equal(hyp[it, 9:], t), f"t differs at hypothesis {it}"


@pytest.mark.parametrize("n_src,H", [(3000, 4000), (257, 1500), (40, 600)])
def test_inlier_counts_bit_exact(ctx, oracle, n_src, H):
    """Per-hypothesis inlier counts == oracle for every hypothesis, degenerate triples included
    (small n_src makes i0==i1 collisions and Lemire rejections likely)."""
    case = small_ransac_case(n_src=n_src, n_tgt=2500, H=H, seed=n_src)
    corr = oracle.match_features(case.source_desc, case.target_desc)
    ctx.set_clouds(case.source, case.target)
    ctx.set_correspondences(corr)
    ctx.ransac_prepare(case.voxel_size, H, 2.0)
    ctx.ransac_score()
    got = ctx.ransac_counts()
    ref = oracle.ransac(case.source, case.target, corr, case.voxel_size, H, 2.0, want_counts=True)
    want = ref.extra["counts"]
    assert np.array_equal(got, want)
    assert (want == -1).sum() == (got == -1).sum()


def test_ransac_end_to_end_matches_oracle(ctx, oracle):
    case = small_ransac_case(n_src=3000, n_tgt=2500, H=4000)
    T, fit, rmse, best = ctx.ransac(case.source, case.target, case.source_desc, case.target_desc,
                                    case.voxel_size, case.max_iterations, 0.999)
    ref = oracle.ransac_registration(case.source, case.target, case.source_desc, case.target_desc,
                                     case.voxel_size, case.max_iterations, 0.999)
    corr = oracle.match_features(case.source_desc, case.target_desc)
    ref2 = oracle.ransac(case.source, case.target, corr, case.voxel_size, case.max_iterations, 0.999)
    assert best == ref2.extra["best_iter"]
    assert np.array_equal(T, ref.transformation)          # same hypothesis, same arithmetic => same bits
    assert fit == ref.fitness
    assert rmse == ref.rmse                               # sequential fp32 sum reproduced on device
    assert syn.rotation_error(T, case.T_true) < 0.05      # and it is the right pose


def test_ransac_early_exit_and_first_wins(ctx, oracle):
    """Noise-free rigid motion: the first non-degenerate triple already has fitness 1.0 > confidence,
    so the reference breaks there (registration.cpp:290); later, equally good hypotheses must not win."""
    rng = np.random.default_rng(5)
    tgt = rng.uniform(-0.2, 0.2, (800, 3)).astype(np.float32)
    T = syn.rigid([0.1, 0.7, 0.3], 33.0, [0.05, 0.02, -0.04])
    src = syn.apply(np.linalg.inv(T), tgt)
    corr = np.arange(800, dtype=np.uint32)
    ctx.set_clouds(src, tgt); ctx.set_correspondences(corr)
    ctx.ransac_prepare(0.001, 500, 0.9)
    ctx.ransac_score()
    import torch
    keys = torch.zeros(2, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    ctx.ransac_reduce(0, 500, keys.data_ptr())
    Tg, fit, rmse, best = ctx.ransac_finish(keys.data_ptr())
    ref = oracle.ransac(src, tgt, corr, 0.001, 500, 0.9, want_counts=True)
    assert best == ref.extra["best_iter"]
    assert ref.extra["iters_run"] == best + 1             # the oracle stopped right there
    assert np.array_equal(Tg, ref.transformation) and fit == ref.fitness and rmse == ref.rmse


def test_ransac_no_inliers_returns_identity(ctx, oracle):
    rng = np.random.default_rng(9)
    src = rng.uniform(-1, 1, (200, 3)).astype(np.float32)
    tgt = rng.uniform(50, 60, (200, 3)).astype(np.float32)
    sd = syn.histograms(200, rng); td = syn.histograms(200, rng)
    T, fit, rmse, best = ctx.ransac(src, tgt, sd, td, 1e-6, 300, 0.999)
    ref = oracle.ransac_registration(src, tgt, sd, td, 1e-6, 300, 0.999)
    assert np.array_equal(T, ref.transformation) and fit == ref.fitness and rmse == ref.rmse
    if ref.fitness == 0.0:
        assert best == -1 and np.array_equal(T, np.eye(4, dtype=np.float32))


def test_ransac_degenerate_inputs(ctx):
    """max_iterations = 0 and single-point clouds never throw (registration.hpp:27-29 defaults)."""
    rng = np.random.default_rng(1)
    sd = syn.histograms(4, rng)
    pts = rng.random((4, 3), dtype=np.float32)
    T, fit, rmse, best = ctx.ransac(pts, pts, sd, sd, 0.001, 0, 0.999)
    assert np.array_equal(T, np.eye(4, dtype=np.float32)) and fit == 0.0 and rmse == 0.0 and best == -1
    T, fit, rmse, best = ctx.ransac(pts[:1], pts[:1], sd[:1], sd[:1], 0.001, 50, 0.999)
    assert fit == 0.0 and best == -1                     # every triple is degenerate with one point


# ------------------------------------------------------------------ ICP
def icp_small(seed=21, n_model=6000, n_scene=9000, **kw):
    return syn.icp_case(n_model=n_model, n_scene=n_scene, seed=seed, **kw)


def test_icp_nearest_bit_exact(ctx, oracle):
    """Iteration-0 NN: same index and same d2 bits wherever the reference keeps the match."""
    case = icp_small()
    ref = oracle.icp(case.source, case.target, case.target_normals, case.T_init, case.threshold, 1, True, want_nn0=True)
    ctx.set_clouds(case.source, case.target, case.target_normals)
    idx, d2 = ctx.icp_nearest(case.T_init, case.threshold)
    ref_idx, ref_d2 = ref.extra["nn_idx0"], ref.extra["nn_d2_0"]
    kept = np.sqrt(ref_d2) <= np.float32(case.threshold)
    assert kept.sum() > 1000
    assert np.array_equal(idx[kept], ref_idx[kept])
    assert np.array_equal(d2[kept], ref_d2[kept])
    assert (idx[~kept] == 0xFFFFFFFF).all()
    assert int(kept.sum()) == ref.extra["ncorr"][0]


def test_icp_nearest_ties_resolve_to_lowest_index(ctx, oracle):
    rng = np.random.default_rng(2)
    base = rng.uniform(-0.05, 0.05, (300, 3)).astype(np.float32)
    tgt = np.concatenate([base, base, base])[rng.permutation(900)]       # every point three times
    src = (base + rng.normal(0, 1e-4, base.shape)).astype(np.float32)
    ctx.set_clouds(src, tgt)
    idx, d2 = ctx.icp_nearest(np.eye(4, dtype=np.float32), 0.01)
    ref = oracle.icp(src, tgt, None, np.eye(4, dtype=np.float32), 0.01, 1, False, want_nn0=True)
    assert np.array_equal(idx, ref.extra["nn_idx0"])
    assert np.array_equal(d2, ref.extra["nn_d2_0"])


def test_icp_point_to_plane_default_is_bit_identical_and_converges(ctx, oracle):
    """The path b3d_icp / the shim run (no mode set): reference-order sums, so not just within 1e-5 / 1e-6 m but equal."""
    case = icp_small()
    ref = oracle.icp(case.source, case.target, case.target_normals, case.T_init, case.threshold, 30, True)
    T, fit, rmse, iters = ctx.icp(case.source, case.target, case.target_normals, case.T_init, case.threshold, 30, True)
    assert iters == ref.extra["iters_run"]
    assert np.array_equal(T, ref.transformation) and fit == ref.fitness and rmse == ref.rmse
    assert syn.rotation_error(T, case.T_true) < 1e-2        # and it converged to the real pose (sparse model)


def test_icp_point_to_plane_fast_mode_within_tolerance(ctx, oracle):
    """b3d_set_icp_mode(1): fp64 tree sums (order-free).  Holds the north-star tolerance on a well-conditioned threshold."""
    case = icp_small()
    ref = oracle.icp(case.source, case.target, case.target_normals, case.T_init, case.threshold, 30, True)
    ctx.set_icp_mode(1)
    try:
        T, fit, rmse, iters = ctx.icp(case.source, case.target, case.target_normals, case.T_init, case.threshold, 30, True)
    finally:
        ctx.set_icp_mode(0)
    assert iters == ref.extra["iters_run"] and fit == ref.fitness and abs(rmse - ref.rmse) < 1e-7
    assert syn.rotation_error(T, ref.transformation) < ROT_TOL
    assert syn.translation_error(T, ref.transformation) < TRANS_TOL


@pytest.mark.parametrize("mode", [0, 3])
@pytest.mark.parametrize("iters", [1, 2, 5, 30])
def test_icp_point_to_point_is_bit_identical(ctx, oracle, iters, mode):
    """Point-to-point adds in the reference's order: transform, fitness and rmse match the oracle bit for bit at every
    iteration count (so the convergence break fires at the same iteration too).  mode 0 = default (parallel exact sums),
    mode 3 = one-chain replay."""
    case = icp_small()
    ref = oracle.icp(case.source, case.target, case.target_normals, case.T_init, case.threshold, iters, False)
    ctx.set_icp_mode(mode)
    try:
        T, fit, rmse, it = ctx.icp(case.source, case.target, case.target_normals, case.T_init, case.threshold, iters, False)
    finally:
        ctx.set_icp_mode(0)
    assert it == ref.extra["iters_run"]
    assert np.array_equal(T, ref.transformation) and fit == ref.fitness and rmse == ref.rmse


@pytest.mark.parametrize("mode", [0, 3])
@pytest.mark.parametrize("iters", [1, 2, 7, 40])
def test_icp_point_to_plane_reference_order_is_bit_identical(ctx, oracle, iters, mode):
    """ATA / ATb / total_error added one matched point at a time (registration.cpp:343-354): default mode and mode 3."""
    case = icp_small()
    ref = oracle.icp(case.source, case.target, case.target_normals, case.T_init, case.threshold, iters, True)
    ctx.set_icp_mode(mode)
    try:
        T, fit, rmse, n = ctx.icp(case.source, case.target, case.target_normals, case.T_init, case.threshold, iters, True)
    finally:
        ctx.set_icp_mode(0)
    assert n == ref.extra["iters_run"]
    assert np.array_equal(T, ref.transformation) and np.float32(fit) == np.float32(ref.fitness) and np.float32(rmse) == np.float32(ref.rmse)


@pytest.mark.parametrize("n_src,n_tgt,seed", [(2977, 1842, 94), (20000, 3000, 95), (40000, 10000, 96)])
def test_icp_default_mode_at_the_orchestrators_threshold(ctx, oracle, n_src, n_tgt, seed):
    """Threshold 0.4*voxel ~ sensor noise (what Pipeline::processInstance passes, pipeline.cpp:104): few matches, the matched
    set flips with the last bit of the pose, and only the reference's own summation order reproduces its result.  This runs
    the DEFAULT mode — the one b3d_icp, the C++ shim and b3d_register_scene use.  Also the binned (n_src >= 16384) order."""
    c = syn.ransac_case(n_src=n_src, n_tgt=n_tgt, seed=seed, max_iterations=10)
    T0 = c.T_true.copy(); T0[:3, 3] += np.float32(2e-4)
    ref = oracle.icp(c.source, c.target, c.target_normals, T0, c.voxel_size * 0.4, 30, True)
    T, fit, rmse, n = ctx.icp(c.source, c.target, c.target_normals, T0, c.voxel_size * 0.4, 30, True)
    assert n == ref.extra["iters_run"]
    assert syn.rotation_error(T, ref.transformation) < ROT_TOL and syn.translation_error(T, ref.transformation) < TRANS_TOL
    assert np.array_equal(T, ref.transformation)             # in fact equal
    assert np.float32(fit) == np.float32(ref.fitness) and np.float32(rmse) == np.float32(ref.rmse)


def test_icp_fast_mode_is_opt_in_and_drifts_at_the_noise_floor(ctx, oracle):
    """Documents why mode 1 is not the default: same case as above, fp64 tree sums — still a valid registration, but the
    trajectory leaves the reference's (only a loose bound holds)."""
    c = syn.ransac_case(n_src=20000, n_tgt=3000, seed=95, max_iterations=10)
    T0 = c.T_true.copy(); T0[:3, 3] += np.float32(2e-4)
    ref = oracle.icp(c.source, c.target, c.target_normals, T0, c.voxel_size * 0.4, 30, True)
    ctx.set_icp_mode(1)
    try:
        T, fit, rmse, n = ctx.icp(c.source, c.target, c.target_normals, T0, c.voxel_size * 0.4, 30, True)
    finally:
        ctx.set_icp_mode(0)
    assert syn.rotation_error(T, ref.transformation) < 5e-3 and syn.translation_error(T, ref.transformation) < 1e-3


def test_icp_point_to_point_large_binned_source_is_bit_identical(ctx, oracle):
    """n_src >= 16384 takes the cell-binned query order; results are still written and summed in source order."""
    case = syn.icp_case(n_model=3000, n_scene=20000, seed=61)
    ref = oracle.icp(case.source, case.target, None, case.T_init, case.threshold, 6, False)
    T, fit, rmse, it = ctx.icp(case.source, case.target, None, case.T_init, case.threshold, 6, False)
    assert it == ref.extra["iters_run"] and np.array_equal(T, ref.transformation) and fit == ref.fitness and rmse == ref.rmse


def test_icp_point_to_point_fast_mode_within_documented_tolerance(ctx, oracle):
    case = icp_small()
    ref = oracle.icp(case.source, case.target, case.target_normals, case.T_init, case.threshold, 30, False)
    ctx.set_icp_mode(1)
    try:
        T, fit, rmse, iters = ctx.icp(case.source, case.target, case.target_normals, case.T_init, case.threshold, 30, False)
    finally:
        ctx.set_icp_mode(0)
    assert iters == ref.extra["iters_run"] and abs(fit - ref.fitness) < 2e-3 and abs(rmse - ref.rmse) < 1e-6
    assert syn.rotation_error(T, ref.transformation) < P2P_ROT_TOL and syn.translation_error(T, ref.transformation) < P2P_TRANS_TOL


def test_icp_without_normals_falls_back_to_point_to_point(ctx, oracle):
    case = icp_small(seed=33, n_model=3000, n_scene=4000)
    ref = oracle.icp(case.source, case.target, None, case.T_init, case.threshold, 10, True)
    T, fit, rmse, iters = ctx.icp(case.source, case.target, None, case.T_init, case.threshold, 10, True)
    assert iters == ref.extra["iters_run"] and fit == ref.fitness
    assert np.array_equal(T, ref.transformation) and rmse == ref.rmse


def test_icp_too_few_correspondences_keeps_initial(ctx, oracle):
    """n_corr < 3 at iteration 0 => break; result = {initial_transform, 0, 0} (registration.cpp:309-311, 361)."""
    rng = np.random.default_rng(4)
    src = rng.uniform(-1, 1, (500, 3)).astype(np.float32)
    tgt = rng.uniform(10, 11, (400, 3)).astype(np.float32)
    T0 = syn.rigid([0, 0, 1], 10.0, [0.1, 0.2, 0.3]).astype(np.float32)
    T, fit, rmse, iters = ctx.icp(src, tgt, None, T0, 0.01, 20, True)
    assert np.array_equal(T, T0) and fit == 0.0 and rmse == 0.0 and iters == 0


def test_icp_identical_clouds_converges_at_iteration_one(ctx, oracle):
    rng = np.random.default_rng(6)
    pts, nrm = syn.torus(4000, rng)
    ref = oracle.icp(pts, pts, nrm, np.eye(4, dtype=np.float32), 0.002, 50, True)
    T, fit, rmse, iters = ctx.icp(pts, pts, nrm, np.eye(4, dtype=np.float32), 0.002, 50, True)
    assert ref.extra["iters_run"] == 2 and iters == 2      # converged flag needs iter > 0
    assert fit == 1.0 and rmse == 0.0
    assert syn.rotation_error(T, np.eye(4)) < 1e-6 and syn.translation_error(T, np.eye(4)) < 1e-7


# ------------------------------------------------------------------ reference-facing interface
def test_registration_interface_mirrors_reference(b3d, oracle):
    case = small_ransac_case(n_src=1500, n_tgt=1200, H=1500)
    src = b3d.PointCloud(points=case.source); tgt = b3d.PointCloud(points=case.target)
    res = b3d.Registration.ransacRegistration(src, tgt, b3d.FPFHFeatures(case.source_desc), b3d.FPFHFeatures(case.target_desc),
                                              case.voxel_size, case.max_iterations)
    ref = oracle.ransac_registration(case.source, case.target, case.source_desc, case.target_desc, case.voxel_size,
                                     case.max_iterations, 0.999)
    assert np.array_equal(res.transformation, ref.transformation) and res.fitness == ref.fitness
    assert b3d.GPURegistration.isCudaAvailable()
    ic = icp_small(seed=44, n_model=2000, n_scene=2500)
    tgt = b3d.PointCloud(points=ic.target, normals=ic.target_normals)
    r1 = b3d.GPURegistration.icpRefine(b3d.PointCloud(points=ic.source), tgt, ic.T_init, ic.threshold, 15)
    r2 = b3d.Registration.icpRefine(b3d.PointCloud(points=ic.source), tgt, ic.T_init, ic.threshold, 15, True)
    assert np.array_equal(r1.transformation, r2.transformation)


# ------------------------------------------------------------------ staged-call state machine
def test_new_clouds_invalidate_old_features(b3d, ctx):
    """set_clouds(A) -> set_features -> set_clouds(larger B): the descriptors were sized for A (or are a caller's
    device pointer); matching them against B would read out of bounds, so it must be refused."""
    rng = np.random.default_rng(5)
    a_s, a_t = rng.random((50, 3), dtype=np.float32), rng.random((40, 3), dtype=np.float32)
    ctx.set_clouds(a_s, a_t)
    ctx.set_features(syn.histograms(50, rng), syn.histograms(40, rng))
    ctx.match_features()
    ctx.set_clouds(rng.random((5000, 3), dtype=np.float32), rng.random((4000, 3), dtype=np.float32))
    with pytest.raises(b3d.B3DError) as e:
        ctx.match_features()
    assert e.value.status == b3d._capi.B3D_ERR_STATE


def test_host_features_drop_the_resident_model(b3d, ctx):
    """b3d_set_features (host path) overwrites the buffer b3d_prepare_model left the model's descriptors in; a later
    register_scene must not silently match against them."""
    rng = np.random.default_rng(6)
    model, _ = syn.torus(3000, rng)
    ctx.prepare_model(model, 0.01, 30, 0.05)
    pts = rng.random((60, 3), dtype=np.float32)
    ctx.set_clouds(pts, pts)
    ctx.set_features(syn.histograms(60, rng), syn.histograms(60, rng))
    with pytest.raises(b3d.B3DError) as e:
        ctx.register_scene(model, 0.01, 30, 0.05, 100, 0.999, 0.004, 5, True)
    assert e.value.status == b3d._capi.B3D_ERR_STATE


def test_ransac_rmse_parallel_exact_sum_equals_the_one_chain_kernel(ctx, oracle):
    """finish mode 0 (default, csrc/b3d_ess.cuh) and mode 1 (one dependent add per inlier): same bits, equal to the oracle."""
    c = syn.ransac_case(n_src=40_000, n_tgt=20_000, seed=71, max_iterations=500, inlier_frac=0.5)
    res = []
    for mode in (0, 1):
        ctx.set_finish_mode(mode)
        try:
            res.append(ctx.ransac(c.source, c.target, c.source_desc, c.target_desc, c.voxel_size, 500, 0.999))
        finally:
            ctx.set_finish_mode(0)
    assert np.array_equal(res[0][0], res[1][0]) and res[0][1:] == res[1][1:]
    ref = oracle.ransac_registration(c.source, c.target, c.source_desc, c.target_desc, c.voxel_size, 500, 0.999)
    assert np.array_equal(res[0][0], ref.transformation) and res[0][1] == ref.fitness and res[0][2] == ref.rmse


def test_ransac_chunked_scoring_stops_where_the_reference_breaks(ctx, oracle):
    """registration.cpp:290: a clean scene exceeds the confidence after a few iterations and the loop breaks.  Scoring runs in
    chunks of hypothesis ids with a device-side exit flag: later chunks do no work, ids behind the exit read -2 (never ran),
    and the result is the reference's."""
    H = 700_000                                                                # 1.4e10 pair evaluations: two chunks
    c = syn.ransac_case(n_src=20_000, n_tgt=15_000, seed=77, inlier_frac=1.0, noise=0.00005, max_iterations=H)
    corr = np.where(c.true_match >= 0, c.true_match, 0).astype(np.uint32)
    ref = oracle.ransac(c.source, c.target, corr, c.voxel_size, H, 0.5, want_counts=True)
    assert 0 < ref.extra["iters_run"] < 2000                                  # the reference stopped early
    ctx.set_clouds(c.source, c.target); ctx.set_correspondences(corr)
    ctx.ransac_prepare(c.voxel_size, H, 0.5)
    import time
    ctx.ransac_score(); ctx.ransac_counts(0, 8)                               # warm-up + sync
    t0 = time.perf_counter(); ctx.ransac_score(); got = ctx.ransac_counts(); t_exit = time.perf_counter() - t0
    assert np.array_equal(got, ref.extra["counts"])                           # -2 behind the break, like the reference's loop
    ctx.ransac_prepare(c.voxel_size, H, 2.0)
    ctx.ransac_score(); ctx.ransac_counts(0, 8)
    t0 = time.perf_counter(); ctx.ransac_score(); ctx.ransac_counts(0, 8); t_full = time.perf_counter() - t0
    assert t_exit < 0.75 * t_full                                             # the chunks behind the exit did no work
    T, fit, rmse, best = ctx.ransac(c.source, c.target, c.source_desc, c.target_desc, c.voxel_size, H, 0.5)
    r2 = oracle.ransac_registration(c.source, c.target, c.source_desc, c.target_desc, c.voxel_size, H, 0.5)
    assert np.array_equal(T, r2.transformation) and fit == r2.fitness and rmse == r2.rmse
