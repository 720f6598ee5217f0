"""The sharded RANSAC protocol on ONE GPU through the staged C-ABI: ranks are emulated one after the other (each with its
own context), the two collectives are replaced by their definitions (concatenation of the gathered slices / keys).
Checks that any world size, with and without an early exit, gives the sequential reference's winner — the same logic
csrc/b3d_dist.cu runs over NCCL (tests/test_gpu_dist_nccl.py exercises that one when two GPUs are visible)."""
import importlib

import numpy as np
import pytest

bdist = importlib.import_module("3dvision_b200.dist")
syn = importlib.import_module("3dvision_b200.synthetic")
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world,confidence", [(1, 2.0), (3, 2.0), (8, 2.0), (8, 0.30), (5, 0.05)])
def test_emulated_ranks_agree_with_the_oracle(b3d, oracle, world, confidence):
    import torch
    H = 6000
    c = syn.ransac_case(n_src=5000, n_tgt=4000, seed=123, max_iterations=H)
    n = c.source.shape[0]
    ranks = []
    for g in range(world):
        ctx = b3d.Context(0)
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        ctx.set_clouds(c.source, c.target); ctx.set_features(c.source_desc, c.target_desc)
        ranks.append(ctx)
    # matching: contiguous row chunks, all-gather
    corr = np.zeros(n, np.uint32)
    for g, ctx in enumerate(ranks):
        r0, r1 = bdist.shard_range(n, g, world)
        if r1 > r0:
            ctx.match_features(r0, r1)
            corr[r0:r1] = ctx.get_correspondences()[r0:r1]
    assert np.array_equal(corr, oracle.match_features(c.source_desc, c.target_desc))
    # scoring shards + the three keys per rank
    keys = torch.zeros((world, 3), dtype=torch.int64, device="cuda")
    for g, ctx in enumerate(ranks):
        ctx.set_correspondences(corr)
        ctx.ransac_prepare(c.voxel_size, H, confidence)
        h0, h1 = bdist.shard_range(H, g, world)
        ctx.ransac_score(h0, h1)
        ctx.ransac_reduce3(h0, h1, keys[g].data_ptr())
    torch.cuda.synchronize()
    winner = bdist.resolve_keys([tuple(int(v) for v in row) for row in keys.cpu().numpy()])
    ref = oracle.ransac_registration(c.source, c.target, c.source_desc, c.target_desc, c.voxel_size, H, confidence)
    ref_id = oracle.ransac(c.source, c.target, corr, c.voxel_size, H, confidence).extra["best_iter"]
    final = torch.tensor([winner, 0], dtype=torch.int64, device="cuda")
    for ctx in ranks:                                                        # every rank rebuilds the same result from the id
        T, fit, rmse, hid = ctx.ransac_finish(final.data_ptr())
        assert hid == ref_id
        assert np.array_equal(T, ref.transformation) and fit == ref.fitness and rmse == ref.rmse
    for ctx in ranks:
        ctx.close()


def test_one_rank_group_equals_the_plain_call(b3d, oracle):
    """world == 1: b3d_ransac_sharded needs no NCCL and must equal b3d_ransac bit for bit."""
    c = syn.ransac_case(n_src=3000, n_tgt=2500, seed=7, max_iterations=3000)
    with b3d.Context(0) as ctx:
        ctx.comm_init(None, 0, 1)
        a = ctx.ransac_sharded(c.source, c.target, c.source_desc, c.target_desc, c.voxel_size, 3000, 0.999)
        b = ctx.ransac(c.source, c.target, c.source_desc, c.target_desc, c.voxel_size, 3000, 0.999)
        assert np.array_equal(a[0], b[0]) and a[1:] == b[1:]
        ref = oracle.ransac_registration(c.source, c.target, c.source_desc, c.target_desc, c.voxel_size, 3000, 0.999)
        assert np.array_equal(a[0], ref.transformation) and a[1] == ref.fitness and a[2] == ref.rmse
