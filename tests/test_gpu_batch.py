"""Batched multi-object registration (SURVEY.md §8e, configs[3]) through the reference-shaped interface:
a pool of workers, one b3d context each (pipeline.cpp:321-327), every instance = ransacRegistration + icpRefine.
Each instance's result must be what a lone call gives (bit-identical) and match the oracle at the usual bars."""
import importlib

import numpy as np
import pytest

syn = importlib.import_module("3dvision_b200.synthetic")
pipe = importlib.import_module("3dvision_b200.pipeline")
reg = importlib.import_module("3dvision_b200.registration")

pytestmark = pytest.mark.gpu

ROT_TOL, TRANS_TOL = 1e-5, 1e-6
# ICP threshold = FACTOR * voxel: the orchestrator's own 0.4 (pipeline_config.hpp:28, pipeline.cpp:104).  It puts the threshold
# at the synthetic noise floor, where the iteration is chaotic in the last bit of the sums; the default mode adds in the
# reference's order, so the batch must still equal the oracle bit for bit.
FACTOR = 0.4


def _instance(c, H):
    return pipe.Instance(reg.PointCloud(c.source), reg.PointCloud(c.target, c.target_normals),
                         reg.FPFHFeatures(c.source_desc), reg.FPFHFeatures(c.target_desc), c.voxel_size,
                         ransac_iterations=H, icp_iterations=30, icp_distance_factor=FACTOR)


def test_batch_equals_lone_calls_and_oracle(b3d, oracle):
    if not b3d.cuda_available():
        pytest.fail("no CUDA device (no CPU fallback exists)")
    H = 3000
    cases = [syn.ransac_case(n_src=1500 + 211 * i, n_tgt=900 + 157 * i, seed=90 + i, max_iterations=H,
                             inlier_frac=0.55 + 0.04 * i) for i in range(7)]
    insts = [_instance(c, H) for c in cases]
    lone = [pipe.process_instance(i) for i in insts]                 # calling thread's context, one at a time
    for threads in (3, 8):
        batch = pipe.register_batch(insts, num_threads=threads)
        assert len(batch) == len(insts)
        for (c0, f0), (c1, f1) in zip(lone, batch):
            assert np.array_equal(c0.transformation, c1.transformation) and c0.fitness == c1.fitness and c0.rmse == c1.rmse
            assert np.array_equal(f0.transformation, f1.transformation) and f0.fitness == f1.fitness and f0.rmse == f1.rmse
    for c, (coarse, fine) in zip(cases, lone):
        want = oracle.ransac_registration(c.source, c.target, c.source_desc, c.target_desc, c.voxel_size, H, 0.999)
        assert np.array_equal(coarse.transformation, want.transformation)
        assert np.float32(coarse.fitness) == np.float32(want.fitness) and np.float32(coarse.rmse) == np.float32(want.rmse)
        wf = oracle.icp(c.source, c.target, c.target_normals, want.transformation, c.voxel_size * FACTOR, 30, True)
        assert syn.rotation_error(fine.transformation, wf.transformation) < ROT_TOL
        assert syn.translation_error(fine.transformation, wf.transformation) < TRANS_TOL
        assert np.array_equal(fine.transformation, wf.transformation)
        assert np.float32(fine.fitness) == np.float32(wf.fitness) and np.float32(fine.rmse) == np.float32(wf.rmse)
    # the same batch through b3d_pool, the worker pool behind the C-ABI (host threads in C, one context each)
    with b3d.Pool(5, devices=(0,)) as pool:
        for _ in range(2):                                            # a pool is reusable
            got = pool.register([dict(source=c.source, target=c.target, target_normals=c.target_normals, source_desc=c.source_desc,
                                      target_desc=c.target_desc, voxel_size=c.voxel_size, ransac_iterations=H, icp_iterations=30) for c in cases])
            for (c0, f0), (gc, gf) in zip(lone, got):
                assert np.array_equal(c0.transformation, gc[0]) and c0.fitness == gc[1] and c0.rmse == gc[2]
                assert np.array_equal(f0.transformation, gf[0]) and f0.fitness == gf[1] and f0.rmse == gf[2]
        assert pool.register([]) == []


def test_empty_batch_and_single_thread(b3d):
    assert pipe.register_batch([], num_threads=4) == []
    c = syn.ransac_case(n_src=400, n_tgt=300, seed=5, max_iterations=500)
    (coarse, fine), = pipe.register_batch([_instance(c, 500)], num_threads=1)
    assert coarse.transformation.shape == (4, 4) and fine.transformation.shape == (4, 4)
