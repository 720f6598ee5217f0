"""configs[3] per rank under torchrun: where does a rank's batch time go? (own share through b3d_pool vs the whole sharded_batch)"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
b3d = importlib.import_module("3dvision_b200"); syn = b3d.synthetic; bdist = importlib.import_module("3dvision_b200.dist")
cases = syn.batch_cases(64)
insts = [dict(source=c.source, target=c.target, target_normals=c.target_normals, source_desc=c.source_desc, target_desc=c.target_desc, voxel_size=c.voxel_size) for c in cases]
mine = insts[rank::world]
for workers in (8, 4):
    pool = b3d.Pool(workers, devices=(local,))
    pool.register(mine)
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    tt = []
    for _ in range(3):
        if world > 1: dist.barrier()
        t0 = time.perf_counter(); pool.register(mine); tt.append(time.perf_counter() - t0)
    own = float(np.median(tt))
    def run_one(i): return pool.register([i])[0][1][:3]
    run_one.run_many = lambda lst: [f[:3] for _, f in pool.register(lst)]
    bdist.sharded_batch(insts, run_one, device="cuda")
    tt = []
    for _ in range(3):
        if world > 1: dist.barrier()
        t0 = time.perf_counter(); bdist.sharded_batch(insts, run_one, device="cuda"); torch.cuda.synchronize(); tt.append(time.perf_counter() - t0)
    full = float(np.median(tt))
    t = torch.tensor([own, full], dtype=torch.float64, device="cuda")
    if world > 1:
        parts = [torch.zeros_like(t) for _ in range(world)]; dist.all_gather(parts, t)
    else:
        parts = [t]
    if rank == 0:
        print(f"workers={workers} cpus={os.cpu_count()} own-share ms per rank:", [round(1e3 * float(p[0]), 1) for p in parts], " sharded_batch ms per rank:", [round(1e3 * float(p[1]), 1) for p in parts], flush=True)
    pool.close()
if world > 1: dist.destroy_process_group()
