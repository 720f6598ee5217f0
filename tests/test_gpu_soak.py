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
