"""world_size-2 gloo tests of the hypothesis/row sharding protocol (3dvision_b200/dist.py) on CPU.

The CUDA context is replaced by a stand-in that answers each backend call from the oracle, so the
code under test is the protocol the NCCL path runs behind the C-ABI (csrc/b3d_dist.cu), mirrored line for line in
dist.py: all-gather of the correspondence slices, one all-gather of three packed keys per rank, and the local
resolution of the reference's sequential best / early-exit rule."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

bdist = importlib.import_module("3dvision_b200.dist")
syn = importlib.import_module("3dvision_b200.synthetic")


class OracleBackend:
    def __init__(self, case):
        from oracle import oracle as O
        self.O = O; self.case = case; self.n_src = case.source.shape[0]

    def match_rows(self, r0, r1):
        return self.O.match_features(self.case.source_desc, self.case.target_desc, r0, r1)

    def set_correspondences(self, corr):
        self.corr = np.asarray(corr, np.uint32)

    def prepare(self, voxel, H, confidence):
        self.voxel, self.H, self.conf = voxel, H, confidence

    def score(self, h0, h1):
        r = self.O.ransac(self.case.source, self.case.target, self.corr, self.voxel, self.H, 2.0, iter_lo=h0, iter_hi=h1, want_counts=True)
        self.counts = r.extra["counts"]

    def keys3(self, h0, h1):
        """ransac_reduce3_impl: (best over ids up to this range's own first exit, that exit key, best over all ids)."""
        n = np.float32(self.n_src)
        exit_key, overall = 0, 0
        for h in range(h0, h1):
            c = int(self.counts[h])
            if c <= 0:
                continue
            fit = np.float32(c) / n
            if fit > np.float32(self.conf):
                exit_key = max(exit_key, bdist.pack_exit_key(h))
            overall = max(overall, bdist.pack_best_key(fit, h))
        limit = 0xFFFFFFFF - exit_key if exit_key else 0xFFFFFFFF
        up_to = 0
        for h in range(h0, h1):
            c = int(self.counts[h])
            if c > 0 and h <= limit:
                up_to = max(up_to, bdist.pack_best_key(np.float32(c) / n, h))
        return up_to, exit_key, overall

    def finish(self, best_key):
        fit, hid = bdist.unpack_best_key(int(best_key))
        if hid < 0:
            return np.eye(4, dtype=np.float32), 0.0, 0.0, -1
        ok, R, t, _ = self.O.ransac_hypothesis(self.case.source, self.case.target, self.corr, hid)
        T = np.eye(4, dtype=np.float32); T[:3, :3] = R; T[:3, 3] = t
        return T, fit, None, hid


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, confidence, H, seed):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        case = syn.ransac_case(n_src=600, n_tgt=500, seed=seed, max_iterations=H)
        T, fit, _, best = bdist.sharded_ransac(OracleBackend(case), case.voxel_size, H, confidence)
        ref = O.ransac_registration(case.source, case.target, case.source_desc, case.target_desc, case.voxel_size, H, confidence)
        corr = O.match_features(case.source_desc, case.target_desc)
        ref2 = O.ransac(case.source, case.target, corr, case.voxel_size, H, confidence)
        assert best == ref2.extra["best_iter"], (rank, best, ref2.extra["best_iter"])
        assert np.array_equal(T, ref.transformation) and np.float32(fit) == np.float32(ref.fitness)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("confidence,H,seed", [(2.0, 700, 1),      # no early exit: plain (max fitness, min id)
                                                (0.30, 700, 1),     # exit id inside one rank's shard restricts the other's
                                                (0.05, 700, 2),     # exits almost immediately (rank 0's shard)
                                                (0.999, 3, 3)])     # fewer hypotheses than would fill both shards evenly
def test_sharded_ransac_matches_sequential_reference(confidence, H, seed):
    mp.spawn(_worker, args=(2, _free_port(), confidence, H, seed), nprocs=2, join=True)


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 100, 1_000_000):
        for world in (1, 2, 3, 8):
            r = [bdist.shard_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            chunk = -(-total // world)
            assert all(b - a <= chunk for a, b in r) and all(b - a == chunk for a, b in r if b < total)    # equal chunks: one all-gather


def test_resolve_keys_follows_the_sequential_rule():
    k, e = bdist.pack_best_key, bdist.pack_exit_key
    # no exit anywhere: plain maximum (highest fitness, then lowest id)
    assert bdist.resolve_keys([(k(0.2, 5), 0, k(0.2, 5)), (k(0.4, 130), 0, k(0.4, 130)), (k(0.4, 250), 0, k(0.4, 250))]) == k(0.4, 130)
    # rank 1 exits at id 140: rank 0 counts in full, rank 1 only up to 140, rank 2 (which holds a better pose) never ran
    keys = [(k(0.3, 7), 0, k(0.3, 7)), (k(0.6, 140), e(140), k(0.9, 180)), (k(0.95, 210), e(205), k(0.95, 210))]
    assert bdist.resolve_keys(keys) == k(0.6, 140)
    # an earlier rank's best survives if the exit hypothesis is not better than it (cannot happen with one confidence, still the rule)
    assert bdist.resolve_keys([(k(0.7, 3), 0, k(0.7, 3)), (k(0.65, 120), e(120), k(0.65, 120))]) == k(0.7, 3)
    assert bdist.resolve_keys([(0, 0, 0), (0, 0, 0)]) == 0


def test_key_packing_orders_like_the_sequential_rule():
    """MAX over keys == strict '>' on fitness with the earliest id winning (registration.cpp:284)."""
    k = bdist.pack_best_key
    assert k(0.5, 10) > k(0.25, 3) and k(0.5, 3) > k(0.5, 10)
    assert bdist.unpack_best_key(k(0.7001799941062927, 21264)) == (float(np.float32(0.7001799941062927)), 21264)
    assert bdist.unpack_best_key(0) == (0.0, -1)
    assert bdist.pack_exit_key(5) > bdist.pack_exit_key(9)          # MAX picks the earliest exit


def _batch_worker(rank, world, port, n_inst):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        cases = [syn.ransac_case(n_src=300, n_tgt=200 + 10 * i, seed=50 + i, max_iterations=200) for i in range(n_inst)]
        seen = []

        def run_one(c):
            seen.append(c)
            r = O.ransac_registration(c.source, c.target, c.source_desc, c.target_desc, c.voxel_size, 200, 0.999)
            f = O.icp(c.source, c.target, c.target_normals, r.transformation, c.voxel_size * 0.4, 5, True)
            return f.transformation, f.fitness, f.rmse

        out = bdist.sharded_batch(cases, run_one)
        assert len(seen) == len(range(rank, n_inst, world))                 # instance i -> rank i mod G, nothing else
        assert all(c is cases[i] for c, i in zip(seen, range(rank, n_inst, world)))
        for c, (T, fit, rmse) in zip(cases, out):                           # every rank ends with every pose, bit-identical
            T1, f1, r1 = run_one(c)
            assert np.array_equal(T, T1) and np.float32(fit) == np.float32(f1) and np.float32(rmse) == np.float32(r1)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_inst", [5, 1, 0])      # odd count, fewer instances than ranks, empty batch
def test_sharded_batch_deals_instances_round_robin_and_gathers_poses(n_inst):
    mp.spawn(_batch_worker, args=(2, _free_port(), n_inst), nprocs=2, join=True)
