"""The multi-GPU path behind the C-ABI for real: one PROCESS per GPU, NCCL called from C (csrc/b3d_dist.cu), no torch.
Needs at least two sm_100 devices (`gpurun --gpus 2 -- python -m pytest tests/test_gpu_dist_nccl.py -m gpu`); on a
one-GPU box the test skips (NCCL refuses two ranks on one device) and tests/test_gpu_dist_emulated.py covers the logic."""
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

syn = importlib.import_module("3dvision_b200.synthetic")
pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_sharded_calls_over_nccl_equal_the_single_gpu_result(tmp_path, b3d, oracle):
    n_dev = b3d._capi.lib().b3d_device_count()
    if n_dev < 2:
        pytest.skip("needs two GPUs")
    world = 2 if n_dev < 4 else 4
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "stubs", "dist_rank.py"), str(r), str(world), str(tmp_path)]) for r in range(world)]
    assert all(p.wait(timeout=600) == 0 for p in procs)
    got = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    for r in range(1, world):                                                # identical on every rank
        for k in got[0].files:
            assert np.array_equal(got[0][k], got[r][k]), (r, k)
    # == the one-GPU call == the oracle
    c = syn.ransac_case(n_src=20_001, n_tgt=15_000, seed=321, max_iterations=30_000)
    with b3d.Context(0) as ctx:
        for name, conf in (("full", 2.0), ("exit", 0.30)):
            T, fit, rmse, best = ctx.ransac(c.source, c.target, c.source_desc, c.target_desc, c.voxel_size, 30_000, conf)
            assert np.array_equal(got[0][name + "_T"], T) and got[0][name + "_id"][0] == best
            assert got[0][name + "_r"][0] == np.float32(fit) and got[0][name + "_r"][1] == np.float32(rmse)
        rng = np.random.default_rng(55)
        model = syn.rough_torus(60_000, rng)
        Tt = syn.rigid([0.2, 0.9, -0.3], 25.0, [0.05, -0.03, 0.08])
        scene = (syn.apply(np.linalg.inv(Tt), syn.rough_torus(60_000, rng)) + rng.normal(0, 0.0003, (60_000, 3))).astype(np.float32)
        ctx.prepare_model(model, 0.008)
        r = ctx.register_scene(scene, 0.008, ransac_max_iterations=20_000)
        assert np.array_equal(got[0]["scene_coarse"], r["coarse"][0]) and np.array_equal(got[0]["scene_T"], r["refined"][0])
        assert got[0]["scene_id"][0] == r["coarse"][3] and got[0]["scene_id"][1] == r["refined"][3]
    ref = oracle.ransac(c.source, c.target, oracle.match_features(c.source_desc, c.target_desc), c.voxel_size, 30_000, 0.30)
    assert got[0]["exit_id"][0] == ref.extra["best_iter"] and np.array_equal(got[0]["exit_T"], ref.transformation)
