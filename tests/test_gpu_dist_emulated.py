"""The NCCL sharding path on ONE GPU: ranks are emulated sequentially (each with its own context and the torch
default stream handoff the real path uses), the collectives are replaced by their definitions (SUM of disjoint
slices, MAX of keys).  Checks that any world size gives the sequential reference's winner."""
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
    ranks = []
    for g in range(world):
        ctx = b3d.Context(0)
        ctx.set_clouds(c.source, c.target); ctx.set_features(c.source_desc, c.target_desc)
        ranks.append(bdist.CudaBackend(ctx, c.source.shape[0]))            # binds the context to torch's current stream
    # matching: disjoint row slices, all-reduce(SUM)
    total = torch.zeros(c.source.shape[0], dtype=torch.int32, device="cuda")
    for g, be in enumerate(ranks):
        r0, r1 = bdist.shard_range(c.source.shape[0], g, world)
        total += be.match_rows(r0, r1)
    assert np.array_equal(total.cpu().numpy().astype(np.uint32), oracle.match_features(c.source_desc, c.target_desc))
    for be in ranks:
        be._corr.copy_(total); be.correspondences_ready()
        be.prepare(c.voxel_size, H, confidence)
    # scoring shards + the two MAX reductions
    keys = []
    for g, be in enumerate(ranks):
        h0, h1 = bdist.shard_range(H, g, world)
        be.score(h0, h1)
        keys.append(be.reduce(h0, h1, with_limit=False).clone())
    exit_key = torch.stack([k[1] for k in keys]).max()
    best = []
    for g, be in enumerate(ranks):
        h0, h1 = bdist.shard_range(H, g, world)
        be.keys[1] = exit_key
        best.append(be.reduce(h0, h1, with_limit=True)[0].clone())
    winner = torch.stack(best).max()
    ref = oracle.ransac_registration(c.source, c.target, c.source_desc, c.target_desc, c.voxel_size, H, confidence)
    corr = oracle.match_features(c.source_desc, c.target_desc)
    ref_id = oracle.ransac(c.source, c.target, corr, c.voxel_size, H, confidence).extra["best_iter"]
    for be in ranks:                                                         # every rank rebuilds the same result from the id
        be.keys[0] = winner
        T, fit, rmse, hid = be.finish()
        assert hid == ref_id
        assert np.array_equal(T, ref.transformation) and fit == ref.fitness and rmse == ref.rmse
    for be in ranks:
        be.ctx.close()
