"""Hypothesis / row sharding of ransacRegistration over torch.distributed (one process per GPU).

The reference has no multi-GPU code; this is the §8(e) design:
  * descriptor matching: source rows split across ranks, index slices combined with one
    all-reduce (each rank contributes zeros outside its slice);
  * hypotheses: rank g scores ids [g*H/G, (g+1)*H/G) against the replicated pair array;
  * selection: two 8-byte MAX all-reduces reproduce the sequential rule of
    registration.cpp:284-290 exactly —
      keys[1] = 0xFFFFFFFF - (first id with fitness > confidence)   (early exit)
      keys[0] = (fitness_bits << 32) | (0xFFFFFFFF - id), restricted to ids <= that exit id
    so the winner is (max fitness, min id) among the iterations the reference would have run;
  * the winner's transform / fitness / rmse are recomputed on every rank from its id.
Single-cloud ICP does not shard ("replicas only").

The protocol is written against a small backend interface so the same code runs over NCCL
with the CUDA context and over gloo with a CPU stand-in in the tests.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int):
    """Contiguous, balanced [lo, hi) of `total` items for `rank` of `world`."""
    return (total * rank) // world, (total * (rank + 1)) // world


def pack_best_key(fitness: float, hyp_id: int) -> int:
    """(fitness_bits << 32) | (0xFFFFFFFF - id): MAX picks the highest fitness, then the lowest id."""
    bits = int(np.float32(fitness).view(np.uint32))
    return (bits << 32) | (0xFFFFFFFF - int(hyp_id))


def unpack_best_key(key: int):
    if key == 0:
        return 0.0, -1
    fitness = float(np.uint32(key >> 32).view(np.float32))
    return fitness, 0xFFFFFFFF - (key & 0xFFFFFFFF)


def pack_exit_key(hyp_id: int) -> int:
    return 0xFFFFFFFF - int(hyp_id)


class CudaBackend:
    """Adapter from a b3d Context (clouds + features already resident) to the protocol."""

    def __init__(self, ctx, n_src: int):
        self.ctx = ctx
        self.n_src = n_src
        self.keys = torch.zeros(2, dtype=torch.int64, device="cuda")
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        self._corr = torch.zeros(max(n_src, 1), dtype=torch.int32, device="cuda")     # exchange buffer (uint32 bits)

    def match_rows(self, r0, r1):
        self.ctx.match_features(r0, r1)
        self.ctx.get_correspondences_device(self._corr.data_ptr())                    # stream-ordered D2D
        self._corr[:r0].zero_()
        self._corr[r1:].zero_()
        return self._corr

    def correspondences_ready(self):
        self.ctx.set_correspondences_device(self._corr.data_ptr())

    def prepare(self, voxel, H, confidence):
        self.ctx.ransac_prepare(voxel, H, confidence)

    def score(self, h0, h1):
        self.ctx.ransac_score(h0, h1)

    def reduce(self, h0, h1, with_limit):
        k = self.keys
        self.ctx.ransac_reduce(h0, h1, k.data_ptr(), k[1:].data_ptr() if with_limit else None)
        return k

    def finish(self):
        return self.ctx.ransac_finish(self.keys.data_ptr())


def sharded_ransac(backend, voxel_size: float, max_iterations: int, confidence: float, group=None,
                   match: bool = True):
    """Run ransacRegistration with rows and hypotheses sharded over `group`. Returns
    (T 4x4, fitness, rmse, best_iteration) — identical on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if match:
        r0, r1 = shard_range(backend.n_src, rank, world)
        corr = backend.match_rows(r0, r1)
        if world > 1:
            dist.all_reduce(corr, op=dist.ReduceOp.SUM, group=group)     # disjoint slices, zeros elsewhere
        backend.correspondences_ready()
    backend.prepare(voxel_size, max_iterations, confidence)
    h0, h1 = shard_range(max_iterations, rank, world)
    backend.score(h0, h1)
    keys = backend.reduce(h0, h1, with_limit=False)
    if world > 1:
        dist.all_reduce(keys[1:2], op=dist.ReduceOp.MAX, group=group)    # global first-exit id
        keys = backend.reduce(h0, h1, with_limit=True)                   # best among ids <= exit id
        dist.all_reduce(keys[0:1], op=dist.ReduceOp.MAX, group=group)
    return backend.finish()


def sharded_batch(instances, run_one, group=None, device="cpu"):
    """Batched multi-object registration over `group` (SURVEY.md §8e, configs[3]): instance i is owned by
    rank i mod G and processed there by `run_one(instance) -> (T 4x4, fitness, rmse)`; there is no
    data-path collective.  The poses are then gathered with one SUM all-reduce of an (n,18) fp32 table
    whose rows are zero everywhere but on the owner (disjoint rows, so SUM is a gather and the result is
    bit-identical to the owner's).  `run_many`, if the callable has it, is used instead so a rank can
    overlap its instances on its own worker pool.  Returns a list of (T, fitness, rmse), same on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    instances = list(instances)
    n = len(instances)
    mine = list(range(rank, n, world))
    many = getattr(run_one, "run_many", None)
    local = many([instances[i] for i in mine]) if many else [run_one(instances[i]) for i in mine]
    table = np.zeros((n, 18), np.float32)
    for i, (T, fit, rmse) in zip(mine, local):
        table[i, :16] = np.asarray(T, np.float32).reshape(16)
        table[i, 16], table[i, 17] = fit, rmse
    t = torch.from_numpy(table).to(device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    table = t.cpu().numpy()
    return [(table[i, :16].reshape(4, 4).copy(), float(table[i, 16]), float(table[i, 17])) for i in range(n)]
