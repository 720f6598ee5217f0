"""Multi-GPU helpers (one process per GPU).

The sharded hot path itself lives behind the C-ABI (csrc/b3d_dist.cu: ``b3d_ransac_sharded``,
``b3d_register_scene_sharded``; NCCL called from C, no Python on the data path).  What is here:

  * ``init_comm``: hands rank 0's ``ncclUniqueId`` to the other ranks of a torch.distributed group and calls
    ``b3d_comm_init`` — torch.distributed is only the bootstrap (a C++ host would use MPI, a file or its own pool);
  * ``sharded_ransac`` + ``resolve_keys``: a line-for-line mirror of the C protocol over a small backend interface, so
    the selection logic can be tested at world_size 2 over gloo on a CPU box with an oracle-backed stand-in
    (tests/test_dist_gloo.py).  The protocol:
      - descriptor matching: source rows in contiguous chunks of ceil(n/G), index slices all-gathered;
      - hypotheses: rank g scores ids [g*ceil(H/G), (g+1)*ceil(H/G)) against the replicated pair array;
      - selection: ONE all-gather of three keys per rank — the rank's first id with fitness > confidence, its best key
        over all its ids, its best key over ids up to that exit — from which every rank resolves the sequential rule of
        registration.cpp:284-290 locally (ranges are contiguous and ordered by rank);
      - the winner's transform / fitness / rmse are recomputed on every rank from its id;
  * ``sharded_batch``: batched multi-object registration, instance i -> rank i mod G, no data-path collective.
Single-cloud ICP does not shard ("replicas only").
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int):
    """Contiguous [lo, hi) of `total` items for `rank` of `world`: chunks of ceil(total / world), as b3d_dist.cu deals
    rows and hypothesis ids (equal chunks let the index slices be gathered with one ncclAllGather)."""
    chunk = (total + world - 1) // world
    lo = min(chunk * rank, total)
    return lo, min(lo + chunk, total)


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


def resolve_keys(all_keys) -> int:
    """resolve_keys_kernel of b3d_dist.cu: all_keys[r] = (best up to r's own exit, r's first exit key or 0, best over all of
    r's ids), ranks in id order.  The reference stops at the first exit: ranks before it count in full, the exit rank up
    to its exit, later ranks not at all."""
    best = 0
    for up_to_exit, exit_key, overall in all_keys:
        if exit_key:
            return max(best, up_to_exit)
        best = max(best, overall)
    return best


def init_comm(ctx, group=None):
    """Make `ctx` a rank of an NCCL communicator spanning `group` (b3d_comm_init); collective."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world == 1:
        ctx.comm_init(None, 0, 1)
        return
    box = [ctx.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    ctx.comm_init(box[0], rank, world)


def sharded_ransac(backend, voxel_size: float, max_iterations: int, confidence: float, group=None):
    """Mirror of ransac_sharded_resident_impl (b3d_dist.cu) over a backend with match_rows / set_correspondences /
    prepare / score / keys3 / finish.  Returns (T 4x4, fitness, rmse, best_iteration) — identical on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = backend.n_src
    chunk = (n + world - 1) // world
    r0, r1 = shard_range(n, rank, world)
    mine = torch.zeros(max(chunk, 1), dtype=torch.int64)
    if r1 > r0:
        mine[:r1 - r0] = torch.from_numpy(np.asarray(backend.match_rows(r0, r1), np.int64))
    if world > 1:
        parts = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine, group=group)                          # ncclAllGather of the (padded) index slices
        corr = torch.cat(parts)[:n]
    else:
        corr = mine[:n]
    backend.set_correspondences(corr.numpy().astype(np.uint32))
    backend.prepare(voxel_size, max_iterations, confidence)
    h0, h1 = shard_range(max_iterations, rank, world)
    backend.score(h0, h1)
    k3 = torch.tensor([int(k) for k in backend.keys3(h0, h1)], dtype=torch.int64)      # keys fit in 63 bits (fitness <= 1.0)
    if world > 1:
        parts = [torch.zeros_like(k3) for _ in range(world)]
        dist.all_gather(parts, k3, group=group)                            # the one selection collective: 24 bytes per rank
        all_keys = [tuple(int(v) for v in p) for p in parts]
    else:
        all_keys = [tuple(int(v) for v in k3)]
    return backend.finish(resolve_keys(all_keys))


def sharded_batch(instances, run_one, group=None, device="cpu"):
    """Batched multi-object registration over `group` (SURVEY.md §8e, configs[3]): instance i is owned by
    rank i mod G and processed there by `run_one(instance) -> (T 4x4, fitness, rmse)`; there is no
    data-path collective.  The poses are then gathered with one all-gather of an (n,18) fp32 table, each
    rank's rows taken from their owner (bit-identical to the owner's, -0.0 included).  `run_many`, if the callable has it, is used instead so a rank can
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
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t, group=group)                             # a gather, not a SUM: bit patterns (-0.0 included) survive
        for r, part in enumerate(parts):
            rows = list(range(r, n, world))
            if rows:
                t[rows] = part[rows]
    table = t.cpu().numpy()
    return [(table[i, :16].reshape(4, 4).copy(), float(table[i, 16]), float(table[i, 17])) for i in range(n)]
