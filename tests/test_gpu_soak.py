"""A slice of the randomised soaks (scripts/fuzz_registration.py) as regression tests: random small ICP and
ransacRegistration cases through the C-ABI, every output equal to the oracle's bit for bit.  The first ICP seeds are the
ones whose singular 6x6 systems (a handful of correspondences) returned large update angles and so exposed CUDA's sinf /
cosf against glibc's (csrc/b3d_libm.cuh)."""
import importlib

import numpy as np
import pytest

syn = importlib.import_module("3dvision_b200.synthetic")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(b3d):
    with b3d.Context(0) as c:
        yield c


def _bits(a):
    return np.asarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("seed", [80, 436, 592] + list(range(1000, 1040)))
def test_random_icp_case_is_bit_identical(ctx, oracle, seed):
    c, plane = syn.random_icp_case(seed)
    ref = oracle.icp(c.source, c.target, c.target_normals, c.T_init, c.threshold, c.iterations, plane)
    T, fit, rmse, n = ctx.icp(c.source, c.target, c.target_normals, c.T_init, c.threshold, c.iterations, plane)
    assert n == ref.extra["iters_run"]
    assert np.array_equal(_bits(T), _bits(ref.transformation), ) and _bits(fit) == _bits(ref.fitness) and _bits(rmse) == _bits(ref.rmse)


@pytest.mark.parametrize("seed", range(2000, 2040))
def test_random_ransac_case_is_bit_identical(ctx, oracle, seed):
    c, conf = syn.random_ransac_case(seed)
    H = c.max_iterations
    ref = oracle.ransac_registration(c.source, c.target, c.source_desc, c.target_desc, c.voxel_size, H, conf)
    T, fit, rmse = ctx.ransac(c.source, c.target, c.source_desc, c.target_desc, c.voxel_size, H, conf)[:3]
    assert np.array_equal(_bits(T), _bits(ref.transformation)) and _bits(fit) == _bits(ref.fitness) and _bits(rmse) == _bits(ref.rmse)


def _same(a, b):
    a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("seed", range(3000, 3040))
def test_random_cloud_front_end_is_bit_identical(ctx, oracle, b3d, seed):
    """scripts/fuzz_features.py: voxelDownsample, estimateNormals, computeFPFH on random (often degenerate) clouds."""
    xyz, voxel, k, radius = syn.random_cloud(seed)
    try:
        down, _ = ctx.voxel_downsample(xyz, voxel)
    except b3d.B3DError:
        pytest.skip("|coordinate / voxel| >= 2^20 is refused")
    want = oracle.voxel_downsample(xyz, voxel)
    assert _same(down, want)
    nrm = oracle.estimate_normals(want, k)
    assert _same(ctx.estimate_normals(want, k), nrm)
    assert _same(ctx.compute_fpfh(want, nrm, radius), oracle.compute_fpfh(want, nrm, radius))


@pytest.mark.parametrize("seed", range(4000, 4040))
def test_random_descriptor_sets_match_identically_on_both_matchers(ctx, oracle, seed):
    """scripts/fuzz_match.py: tensor-core screen + exact re-score (mode 2) and the CUDA-core kernel (mode 1)."""
    sd, td = syn.random_descriptors(seed)
    want = oracle.match_features(sd, td)
    ctx.set_clouds(np.zeros((sd.shape[0], 3), np.float32), np.zeros((td.shape[0], 3), np.float32))
    ctx.set_features(sd, td)
    try:
        for mode in (2, 1):
            ctx.set_match_mode(mode)
            ctx.match_features()
            assert np.array_equal(ctx.get_correspondences(), want)
    finally:
        ctx.set_match_mode(0)


@pytest.mark.parametrize("seed", [856] + list(range(5000, 5030)))
def test_random_scene_through_the_one_call_pipeline_is_bit_identical(b3d, oracle, seed):
    """scripts/fuzz_pipeline.py: b3d_prepare_model + b3d_register_scene against the oracle's stage-by-stage chain.  Seed 856
    is the case whose 11th ICP iteration had a sum on a float rounding tie (tests/test_ess_host.py)."""
    c = syn.random_scene_case(seed)
    tgt = oracle.voxel_downsample(c["model"], c["voxel"]); src = oracle.voxel_downsample(c["scene"], c["voxel"])
    if tgt.shape[0] == 0 or src.shape[0] == 0:
        pytest.skip("empty after down-sampling")
    tn = oracle.estimate_normals(tgt, c["k"]); tf = oracle.compute_fpfh(tgt, tn, c["radius"])
    sn = oracle.estimate_normals(src, c["k"]); sf = oracle.compute_fpfh(src, sn, c["radius"])
    coarse = oracle.ransac_registration(src, tgt, sf, tf, c["voxel"], c["H"], c["conf"])
    fine = oracle.icp(src, tgt, tn, coarse.transformation, c["icp_thr"], c["icp_iters"], c["plane"])
    with b3d.Context(0) as cx:
        assert cx.prepare_model(c["model"], c["voxel"], c["k"], c["radius"]) == tgt.shape[0]
        out = cx.register_scene(c["scene"], c["voxel"], c["k"], c["radius"], c["H"], c["conf"], c["icp_thr"], c["icp_iters"], c["plane"])
    T0, f0, r0, _ = out["coarse"]; T1, f1, r1, it = out["refined"]
    assert out["n_source_points"] == src.shape[0]
    assert np.array_equal(_bits(T0), _bits(coarse.transformation)) and _bits(f0) == _bits(coarse.fitness) and _bits(r0) == _bits(coarse.rmse)
    assert it == fine.extra["iters_run"]
    assert np.array_equal(_bits(T1), _bits(fine.transformation)) and _bits(f1) == _bits(fine.fitness) and _bits(r1) == _bits(fine.rmse)
