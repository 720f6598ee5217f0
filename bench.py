#!/usr/bin/env python
"""bench.py — registration hot-path benchmark (contract in the task statement, section 4).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Headline: RANSAC hypotheses/s of one whole Registration::ransacRegistration call
(FPFH matching + hypothesis generation + inlier scoring + selection) on BASELINE.json
configs[2]: Ns = Nt = 100 000 descriptors / correspondences, H = 1 000 000 hypotheses.
N > 1 shards source rows and hypothesis ids across ranks behind the C-ABI (b3d_ransac_sharded: NCCL called from C,
one all-gather of index slices + one of three keys per rank; strong scaling).  `value` is device time with inputs
resident in HBM; `e2e` goes through the reference-facing C-ABI call with pinned HOST buffers (H2D + D2H inside the
timed region; each rank uploads only its rows of the source descriptors).  `also` carries the other figures
BASELINE.json's metric names: ICP iterations/s on configs[1] (300k x 100k point-to-plane, 50 iterations, in the default
reference-order mode and in the opt-in fast mode), full registration ms, the whole pipeline from a raw 1M-point scene
(sharded over the N GPUs), configs[0] in full on CPU and GPU, configs[3] batched, configs[4] stress.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

OPS_PER_PAIR = 28          # un-fused fp32 ops per (hypothesis, correspondence)
N_SRC = N_TGT = 100_000
N_HYP = 1_000_000
WORKLOAD = "configs[2]: FPFH match + RANSAC, Ns=Nt=100k descriptors/correspondences, H=1M hypotheses"
CONFIDENCE = 2.0           # never exits early: all H hypotheses are scored (throughput setting)
ISSUE_PEAK = 148 * 4 * 32 * 1.965e9      # lane-instructions/s: 148 SMs x 4 schedulers x 32 lanes x 1.965 GHz (nominal)
SCORE_LANE_INSTR_PER_PAIR = 13.4         # ncu: 4.18e10 warp instructions for 1e11 pairs (profiles/r1_final_ncu_score_kp2.csv)
SCORE_FMA_PIPE_PCT = 69.7                # ncu sm__inst_executed_pipe_fma / fmaheavy utilisation of the same capture


def bench_config():
    """The `config` object — identical in both arms (ours and --impl reference)."""
    return {"workload": WORKLOAD, "n_src": N_SRC, "n_tgt": N_TGT, "hypotheses": N_HYP, "confidence": CONFIDENCE,
            "l2": "GPU arm: flushed (256 MiB write) between timed steps; CPU arm: not applicable"}


# ----------------------------------------------------------------------------- helpers
class ClockSampler:
    """nvidia-smi clock / throttle sampling during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush(); self.tmp.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.tmp.read().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        self.tmp.close()
        try:
            os.unlink(self.tmp.name)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def pinned(a: np.ndarray):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    return t, t.numpy()


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ----------------------------------------------------------------------------- CPU baseline (oracle)
def cpu_sample(case, corr_cache: dict, threads: int, match_rows: int, hyps: int):
    """Time the oracle on a bounded slice of the workload; every figure is exactly linear in the
    sliced dimension. Returns per-thread seconds for (match slice, ransac slice)."""
    from oracle import oracle as O
    if "corr" not in corr_cache:
        corr_cache["corr"] = np.where(case.true_match >= 0, case.true_match, 0).astype(np.uint32)
    corr = corr_cache["corr"]
    times = [None] * threads

    def work(i):
        t0 = time.perf_counter()
        O.match_features(case.source_desc, case.target_desc, 0, match_rows)
        t1 = time.perf_counter()
        O.ransac(case.source, case.target, corr, case.voxel_size, hyps, CONFIDENCE)
        t2 = time.perf_counter()
        times[i] = (t1 - t0, t2 - t1)

    ths = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    return max(t[0] for t in times), max(t[1] for t in times)


def cpu_rate(case, threads, match_rows, hyps, cache):
    tm, tr = cpu_sample(case, cache, threads, match_rows, hyps)
    full_s = tm * (N_SRC / match_rows) + tr * (N_HYP / hyps)        # one full ransacRegistration on one core
    return threads * N_HYP / full_s, tm, tr, full_s


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the real registration.cpp needs
    Eigen and cannot be compiled here) on the host cores, on a bounded sample per step."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    from oracle import oracle as O
    O.lib()
    syn = importlib.import_module("3dvision_b200.synthetic")
    case = syn.ransac_case(n_src=N_SRC, n_tgt=N_TGT, max_iterations=N_HYP)
    threads = max(1, min(os.cpu_count() or 1, 8))      # reference parallelism = its 8-worker pool over instances
    rows, hyps = 512, 2048
    cache = {}
    for _ in range(args.warmup):
        cpu_rate(case, threads, 16, 32, cache)
    vals, wall = [], 0.0
    for _ in range(args.steps):
        t0 = time.perf_counter()
        v, tm, tr, full_s = cpu_rate(case, threads, rows, hyps, cache)
        wall += time.perf_counter() - t0
        vals.append(v)
    value = float(np.mean(vals))
    sample = (f"per step and per thread: first {rows} source rows x all {N_TGT} targets (matching) + first {hyps} hypotheses "
              f"x all {N_SRC} correspondences (RANSAC), extrapolated linearly to the full job; {threads} independent "
              f"single-threaded registrations in parallel (the reference's only parallelism, pipeline.cpp:321-327)")
    line = {
        "impl": "reference", "metric": "ransac_hyp_per_s", "value": value, "unit": "hyp/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(),
        "cpu_baseline": {"value": value, "unit": "hyp/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "hyp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "extrapolated_full_step_ms": 1e3 * threads * N_HYP / value,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- ours
def run_ours(args):
    # Only the JSON line may reach stdout: NCCL / torch print banners ("NCCL version ...") to fd 1.
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line: dict):
        os.write(real_stdout, (json.dumps(line) + "\n").encode())

    import torch
    import torch.distributed as dist
    rank, local_rank, world = dist_env()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        torch.cuda.set_device(0)
    dev = torch.cuda.current_device()
    b3d = importlib.import_module("3dvision_b200")
    bdist = importlib.import_module("3dvision_b200.dist")
    syn = b3d.synthetic
    if not b3d.cuda_available():
        raise SystemExit("bench.py: no sm_100 CUDA device and no CPU fallback")

    case = syn.ransac_case(n_src=N_SRC, n_tgt=N_TGT, max_iterations=N_HYP)
    ctx = b3d.Context(dev)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)

    # resident inputs (value) ---------------------------------------------------------
    d_src = torch.from_numpy(case.source).cuda(); d_tgt = torch.from_numpy(case.target).cuda()
    d_sd = torch.from_numpy(case.source_desc).cuda(); d_td = torch.from_numpy(case.target_desc).cuda()
    ctx.set_clouds_device(d_src.data_ptr(), N_SRC, d_tgt.data_ptr(), None, N_TGT)
    ctx.set_features_device(d_sd.data_ptr(), d_td.data_ptr())
    bdist.init_comm(ctx)                  # ncclCommInitRank inside libb3d.so; torch.distributed only carries the unique id
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")       # > 126 MB L2
    confidence = CONFIDENCE

    def step_resident():                  # b3d_ransac_sharded_resident: match rows + hypothesis ids sharded, NCCL called from C
        return ctx.ransac_sharded_resident(case.voxel_size, N_HYP, confidence, match=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(dev)
    if rank == 0:
        sampler.start()
    launches0 = ctx.kernel_launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    stage_acc = np.zeros(4)
    result = None
    for i in range(args.steps):
        flush.zero_()
        barrier()
        ev[i][0].record()
        result = step_resident()
        ev[i][1].record()
        torch.cuda.synchronize()
        stage_acc += [ctx.stage_ms(s) for s in range(4)]
    barrier()
    launches = ctx.kernel_launches - launches0
    clocks = sampler.stop() if rank == 0 else {}
    my_ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([my_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = N_HYP * args.steps / (total_ms * 1e-3)

    # end to end through the C-ABI with pinned host buffers -------------------------------
    keep = [pinned(case.source), pinned(case.target), pinned(case.source_desc), pinned(case.target_desc)]
    h_src, h_tgt, h_sd, h_td = (k[1] for k in keep)
    r0, r1 = bdist.shard_range(N_SRC, rank, world)
    h2d = h_src.nbytes + h_tgt.nbytes + (r1 - r0) * 33 * 4 + h_td.nbytes        # per rank: only its rows of the source descriptors
    d2h = 80

    def step_e2e():                       # b3d_ransac_sharded: the reference-facing call, host buffers in, pose out
        return ctx.ransac_sharded(h_src, h_tgt, h_sd, h_td, case.voxel_size, N_HYP, confidence)

    step_e2e()
    barrier()
    e2e_s = 0.0
    for _ in range(args.steps):
        flush.zero_()
        barrier()
        t0 = time.perf_counter()
        result_e2e = step_e2e()
        torch.cuda.synchronize()
        e2e_s += time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = N_HYP * args.steps / float(te.item())
    assert np.array_equal(result[0], result_e2e[0]), "resident and host-buffer paths disagree"
    # restore resident pointers for anything that follows
    ctx.set_clouds_device(d_src.data_ptr(), N_SRC, d_tgt.data_ptr(), None, N_TGT)
    ctx.set_features_device(d_sd.data_ptr(), d_td.data_ptr())

    batched = batched_registration(b3d, bdist, syn, dev, rank, world, barrier, flush)
    pipeline = whole_pipeline(b3d, bdist, syn, flush, rank, world, barrier)          # collective: every rank takes part
    stress = stress_scene(b3d, bdist, syn, flush, rank, world, barrier)

    if rank != 0:
        ctx.close()
        if world > 1:
            dist.destroy_process_group()
        return

    # roofline of the dominant kernel (inlier scoring) -----------------------------------
    h0, h1 = bdist.shard_range(N_HYP, rank, world)
    score_ms = stage_acc[2] / args.steps
    achieved = OPS_PER_PAIR * (h1 - h0) * float(N_SRC) / (score_ms * 1e-3) / 1e12
    peak = ctx.measure_fp32_rate() / 1e12
    roofline = {
        "bound": "fp32", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
        # dram__bytes_read.sum + dram__bytes_write.sum of one full-size launch (ncu --set full, profiles/r1_final_ncu_score_kp2.csv:
        # 55.6 MB + 2.2 MB); the algorithmic bytes are 48 MB of hypotheses + 3.2 MB of pairs + 4 MB of counts = 55.2 MB
        "traffic": (57.8e6 if (h1 - h0) == N_HYP and N_SRC == 100_000 else None), "traffic_unit": "bytes/launch",
        "kernel": "score_screen2_kernel<2>", "kernel_ms": score_ms,
        # how much is left: issued lane-instructions against the scheduler issue peak, and the FMA pipe's share (ncu capture)
        "issue": {"lane_instr_per_pair": SCORE_LANE_INSTR_PER_PAIR,
                  "issued_tera_lane_instr_per_s": SCORE_LANE_INSTR_PER_PAIR * (h1 - h0) * float(N_SRC) / (score_ms * 1e-3) / 1e12,
                  "issue_peak_tera_lane_instr_per_s": ISSUE_PEAK / 1e12,
                  "frac_of_issue_peak": SCORE_LANE_INSTR_PER_PAIR * (h1 - h0) * float(N_SRC) / (score_ms * 1e-3) / ISSUE_PEAK,
                  "fma_pipe_pct_ncu": SCORE_FMA_PIPE_PCT,
                  "source": "instructions per pair and FMA-pipe % from the ncu --set full capture under profiles/; rate from this run's CUDA-event time"},
        "note": ("SURVEY.md §8(d): RANSAC scoring is FP32 CUDA-core issue bound (not HBM, not tensor). achieved = 28 un-fused "
                 "fp32 ops (the reference's arithmetic) x hypotheses x correspondences per launch / CUDA-event kernel time; peak = "
                 "un-fused FMUL+FADD issue rate measured live on this GPU by b3d_measure_fp32_rate (MEASURED_PEAKS.json has no fp32 "
                 "figure). frac > 1 is expected: the kernel screens with packed FFMA2 (15 fused ops instead of 27 un-fused per pair, four hypotheses per thread) "
                 "and re-counts only pairs inside the proven error band with the reference arithmetic, so it retires the "
                 "reference's algorithmic work with fewer issued instructions; counts stay bit-identical (tests)."),
    }
    bf16_peak = 1386.4
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        roofline["hbm_peak_gbs_measured"] = peaks.get("hbm_gbs")
        bf16_peak = float(peaks.get("bf16_tflops_sustained") or bf16_peak)
    except Exception:
        pass
    # second kernel of the step: descriptor matching on the tensor cores (tcgen05 screen GEMM + exact re-score)
    match_ms = stage_acc[0] / args.steps
    m_rows = bdist.shard_range(N_SRC, rank, world)
    match_flops = 2.0 * 33.0 * float(m_rows[1] - m_rows[0]) * float(N_TGT)
    match_roofline = {"bound": "tensor", "achieved": match_flops / (match_ms * 1e-3) / 1e12, "peak": bf16_peak, "unit": "TFLOP/s",
                      "frac": match_flops / (match_ms * 1e-3) / 1e12 / bf16_peak, "kernel": "match_tc_kernel (+ packing, seeding, finalize)",
                      "stage_ms": match_ms,
                      "note": "useful flops 2*33*rows*Nt (K = 33) over the matching stage's CUDA-event time against the sustained bf16 figure of "
                              "MEASURED_PEAKS.json; the GEMM issues K' = 112 (hi/lo bf16 split, 3.4x) and its epilogue is TMEM-read bound "
                              "(every accumulator goes through tcgen05.ld once), DESIGN.md 4.1"}

    line = {
        "metric": "ransac_hyp_per_s", "value": value, "unit": "hyp/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(),
        "sharding": f"source rows and hypothesis ids in {world} contiguous chunks (b3d_ransac_sharded, NCCL from C)",
        "clocks": {"sm_mhz": clocks.get("sm_mhz"), "sm_max_mhz": clocks.get("sm_max_mhz"), "reasons": clocks.get("reasons", [])},
        "e2e": {"value": e2e_value, "unit": "hyp/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "match_roofline": match_roofline,
        "stages_ms": {"match": stage_acc[0] / args.steps, "prepare": stage_acc[1] / args.steps,
                      "score": stage_acc[2] / args.steps, "select_finish": stage_acc[3] / args.steps},
        "result": {"fitness": result[1], "rmse": result[2], "best_iteration": result[3]},
    }

    line["also"] = {"batched": batched, "pipeline": pipeline, "stress": stress}
    if world == 1:
        line["also"].update(secondary(ctx, b3d, syn, case, flush))
        # same call with bail-out scoring (b3d_set_score_mode 3): identical winner / transform / fitness / rmse,
        # hypotheses that provably cannot reach the best full count are dropped part-way
        ctx.set_clouds_device(d_src.data_ptr(), N_SRC, d_tgt.data_ptr(), None, N_TGT)     # secondary() re-used the context
        ctx.set_features_device(d_sd.data_ptr(), d_td.data_ptr())
        ctx.set_score_mode(3)
        for _ in range(2):
            rb = step_resident()
        torch.cuda.synchronize()
        tb = []
        for _ in range(max(args.steps, 3)):
            flush.zero_(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); rb = step_resident(); e1.record(); torch.cuda.synchronize()
            tb.append(e0.elapsed_time(e1))
        ctx.set_score_mode(0)
        assert np.array_equal(rb[0], result[0]) and rb[1:] == result[1:], "bail-out scoring changed the result"
        line["also"]["bailout"] = {"hyp_per_s": N_HYP / (float(np.median(tb)) * 1e-3), "ms_per_step": float(np.median(tb)),
                                   "note": "exact bail-out test (score mode 3): same winner, transform, fitness and rmse as the headline run; "
                                           "not every hypothesis is scored on every correspondence, so it is reported beside, not as, the headline"}
        cache = {}
        t0 = time.perf_counter()
        v, tm, tr, full_s = cpu_rate(case, 1, 512, 2048, cache)
        line["cpu_baseline"] = {
            "value": v, "unit": "hyp/s", "cores": 1, "kind": "port",
            "sample": (f"oracle (CPU restatement of registration.cpp:204-295), one thread: first 512 source rows x {N_TGT} targets "
                       f"({tm:.2f} s) + first 2048 hypotheses x {N_SRC} correspondences ({tr:.2f} s), extrapolated linearly to the "
                       f"full job ({full_s:.0f} s); host has {os.cpu_count()} cores"),
        }
    emit(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def pipeline_cpu_estimate(scene_raw, voxel, n_src, n_model, H, sample=3000):
    """The CPU path's time for the same scene, from the oracle port on one host core: voxelDownsample in full (O(N)), normals and
    FPFH on a `sample`-point subset scaled quadratically (both are O(N^2) scans, registration.cpp:63-102), matching and scoring
    from the per-pair rates of the headline's cpu_baseline run."""
    from oracle import oracle as O
    t0 = time.perf_counter(); down = O.voxel_downsample(scene_raw, voxel); t_down = time.perf_counter() - t0
    sub = np.ascontiguousarray(down[:sample])
    t0 = time.perf_counter(); nrm = O.estimate_normals(sub, 30); t_n = time.perf_counter() - t0
    t0 = time.perf_counter(); O.compute_fpfh(sub, nrm, voxel * 5.0); t_f = time.perf_counter() - t0
    scale = (float(n_src) / float(sub.shape[0])) ** 2
    rng = np.random.default_rng(0)
    a = rng.random((256, 33), dtype=np.float32); b = rng.random((20000, 33), dtype=np.float32)
    t0 = time.perf_counter(); O.match_features(a, b); t_m = (time.perf_counter() - t0) / (256 * 20000)
    src = rng.random((20000, 3), dtype=np.float32); corr = np.arange(20000, dtype=np.uint32)
    t0 = time.perf_counter(); O.ransac(src, src, corr, voxel, 512, 2.0); t_s = (time.perf_counter() - t0) / (512 * 20000)
    parts = {"voxel_downsample_s": t_down, "normals_s": t_n * scale, "fpfh_s": t_f * scale,
             "matching_s": t_m * float(n_src) * float(n_model), "ransac_scoring_s": t_s * float(H) * float(n_src)}
    return {"seconds": float(sum(parts.values())), "cores": 1, "kind": "port",
            "sample": f"voxelDownsample in full; estimateNormals / computeFPFH on {sub.shape[0]} points x ({n_src}/{sub.shape[0]})^2; matching and "
                      "scoring from measured per-pair rates; ICP omitted", **{k: float(v) for k, v in parts.items()}}


def batched_registration(b3d, bdist, syn, dev, rank, world, barrier, flush, n_instances=64, threads=8, reps=3):
    """configs[3]: 64 object instances, each ransacRegistration(H=100000, conf 0.999) + icpRefine(<=200 it), dealt
    round-robin to the ranks (instance i -> rank i mod N), each rank driving its share through b3d_pool — the
    orchestrator's worker pool (pipeline.cpp:321-327) behind the C-ABI: `threads` host threads with one context/stream
    each, pulling instances off a shared counter.  Host buffers in, poses out; wall clock, max over ranks."""
    import torch
    import torch.distributed as dist
    cases = syn.batch_cases(n_instances)
    insts = [dict(source=c.source, target=c.target, target_normals=c.target_normals, source_desc=c.source_desc, target_desc=c.target_desc,
                  voxel_size=c.voxel_size) for c in cases]
    # worker threads are host threads: 8 ranks x 8 workers on a 32-core box oversubscribe the CPU and slow every launch
    # (measured at N = 8: 25.4 ms per batch with 8 workers per rank, 21.6 ms with 4), so a rank takes its share of the cores
    threads = max(2, min(threads, (os.cpu_count() or threads) // max(world, 1)))
    pool = b3d.Pool(threads, devices=(dev,))

    def run_one(inst):
        (_, _, _), (T, fit, rmse, _) = pool.register([inst])[0]
        return T, fit, rmse
    run_one.run_many = lambda lst: [(f[0], f[1], f[2]) for _, f in pool.register(lst)]

    try:
        out = bdist.sharded_batch(insts, run_one, device="cuda")            # warm-up: contexts, workspaces, RNG caches
        secs = []
        for _ in range(reps):
            flush.zero_()
            barrier()
            t0 = time.perf_counter()
            out = bdist.sharded_batch(insts, run_one, device="cuda")
            torch.cuda.synchronize()
            t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            secs.append(float(t.item()))
    finally:
        pool.close()
    s = float(np.median(secs))
    res = {"workload": f"configs[3]: {n_instances} instances (30k scene points vs 2k-10k model points each), ransacRegistration(H=100000, "
                       f"conf 0.999) + icpRefine(thr 0.4*voxel, <=200 it), instance i -> rank i mod {world}, b3d_pool with {threads} worker contexts per rank",
           "registrations_per_s": n_instances / s, "ms_per_batch": 1e3 * s, "n_gpus": world,
           "min_icp_fitness": float(min(o[1] for o in out)),
           "max_rot_err_vs_truth": float(max(syn.rotation_error(o[0], c.T_true) for o, c in zip(out, cases))),
           "max_trans_err_vs_truth": float(max(syn.translation_error(o[0], c.T_true) for o, c in zip(out, cases)))}
    if rank == 0 and world == 1:
        # the reference's CPU pool on a bounded sample: the oracle port on 2 of the 64 instances with a 2 000-hypothesis budget, one
        # thread; scoring is linear in the hypothesis count and dominates, ICP (brute-force NN) is timed in full for 2 iterations
        from oracle import oracle as O
        t_r = t_i = 0.0
        for c in cases[:2]:
            t0 = time.perf_counter(); r = O.ransac_registration(c.source, c.target, c.source_desc, c.target_desc, c.voxel_size, 2000, 2.0); t_r += time.perf_counter() - t0
            t0 = time.perf_counter(); O.icp(c.source, c.target, c.target_normals, r.transformation, c.voxel_size * 0.4, 2, True, stop_on_convergence=False); t_i += time.perf_counter() - t0
        per_inst = (t_r / 2) * (100_000 / 2000) + (t_i / 2) / 2 * 5          # H = 100 000 hypotheses; ~5 ICP iterations to converge
        cores = max(1, min(os.cpu_count() or 1, 8))
        res["cpu_baseline"] = {"value": cores / per_inst, "unit": "registrations/s", "cores": cores, "kind": "port",
                               "sample": f"oracle port, one thread, 2 of the {n_instances} instances: matching + 2 000 hypotheses ({t_r / 2:.2f} s each, scaled x50 to "
                                         f"H = 100 000) + 2 ICP iterations ({t_i / 2:.2f} s, scaled to 5); {per_inst:.0f} s per instance, x {cores} pool threads "
                                         "(the reference's only parallelism)"}
    return res


def icp_cpu_baseline(ic, n_sample=30_000):
    """The reference's CPU ICP (oracle port, one thread) on a bounded sample of configs[1]: ONE point-to-plane iteration of the
    first n_sample source points against all 100k targets; an iteration is a brute-force scan of Ns x Nt pairs
    (registration.cpp:325-338), so its time is linear in Ns."""
    from oracle import oracle as O
    t0 = time.perf_counter()
    O.icp(ic.source[:n_sample], ic.target, ic.target_normals, ic.T_init, ic.threshold, 1, True)
    t = time.perf_counter() - t0
    full = t * ic.source.shape[0] / n_sample
    return {"value": 1.0 / full, "unit": "iterations/s", "cores": 1, "kind": "port",
            "sample": f"one iteration on the first {n_sample} of {ic.source.shape[0]} source points x all {ic.target.shape[0]} targets "
                      f"({t:.1f} s), scaled linearly in the source count to {full:.0f} s per full iteration"}


def demo_scene_full(ctx, b3d):
    """BASELINE.json configs[0]: the demo procedural box scene (pipeline.cpp:211-257, 275-282; voxel 0.001, H = 100 000, conf 0.999,
    ICP thr 0.4*voxel <= 200 it), run IN FULL on the CPU (oracle port, one thread) and on the GPU, end to end from the raw points."""
    from oracle import oracle as O
    voxel = 0.001
    scene, model = O.demo_scene_points(), O.demo_model_points()
    t0 = time.perf_counter()
    src = O.voxel_downsample(scene, voxel); sn = O.estimate_normals(src, 30); sf = O.compute_fpfh(src, sn, voxel * 5.0)
    tgt = O.voxel_downsample(model, voxel); tn = O.estimate_normals(tgt, 30); tf = O.compute_fpfh(tgt, tn, voxel * 5.0)
    t_front = time.perf_counter() - t0
    r = O.ransac_registration(src, tgt, sf, tf, voxel, 100_000, 0.999)
    f = O.icp(src, tgt, tn, r.transformation, float(np.float32(voxel) * np.float32(0.4)), 200, True)
    cpu_s = time.perf_counter() - t0
    import torch
    c = b3d.Context(torch.cuda.current_device())
    try:
        tt = []
        for _ in range(4):
            t0 = time.perf_counter()
            c.prepare_model(model, voxel)
            g = c.register_scene(scene, voxel)
            tt.append(time.perf_counter() - t0)
        same = bool(np.array_equal(g["coarse"][0], r.transformation) and np.array_equal(g["refined"][0], f.transformation)
                    and g["refined"][1] == f.fitness and g["refined"][2] == f.rmse)
    finally:
        c.close()
    return {"workload": "configs[0]: demo procedural box scene, 40 401 raw points -> 32 129 source points vs 1 600 model points; voxelDownsample + "
                        "estimateNormals + computeFPFH (both clouds) + ransacRegistration(H=100000, conf 0.999) + icpRefine(0.4*voxel, <=200 it)",
            "gpu_ms": 1e3 * float(np.median(tt[1:])), "gpu_equals_cpu_bits": same,
            "cpu_baseline": {"value": 1e3 * cpu_s, "unit": "ms", "cores": 1, "kind": "port",
                             "sample": f"the whole configuration, not a sample ({t_front:.1f} s of it in the O(N^2) feature stages)"}}


def secondary(ctx, b3d, syn, case, flush):
    """N=1 only: the other two figures of BASELINE.json's metric."""
    import torch
    out = {}
    # ICP iterations/s, configs[1]: 300k scene vs 100k model, point-to-plane, exactly 50 iterations
    ic = syn.icp_case()
    ctx.set_clouds(ic.source, ic.target, ic.target_normals)
    icp_cpu = icp_cpu_baseline(ic)
    for mode, key in ((0, "icp"), (1, "icp_fast_mode")):
        ctx.set_icp_mode(mode)
        ctx.icp_run(ic.T_init, ic.threshold, 12, True, False)      # warm-up long enough to allocate the second level
        torch.cuda.synchronize()
        reps, ms = 3, 0.0
        for _ in range(reps):
            flush.zero_(); torch.cuda.synchronize()
            T, fit, rmse, iters = ctx.icp_run(ic.T_init, ic.threshold, ic.iterations, True, False)
            ms += ctx.stage_ms(5)
        ms /= reps
        out[key] = {"workload": "configs[1]: 300k-point scene vs 100k-point model, point-to-plane, 50 iterations (no convergence break)",
                    "mode": ("b3d_set_icp_mode(0), the default: sums in the reference's order (parallel exact summation), bit-identical to the CPU path"
                             if mode == 0 else "b3d_set_icp_mode(1), opt-in: fp64 tree sums, order-free, tolerance-level parity"),
                    "iters_per_s": ic.iterations / (ms * 1e-3), "ms_per_iteration": ms / ic.iterations,
                    "grid_build_ms": ctx.stage_ms(4), "fitness": fit, "rmse": rmse,
                    "rot_err_vs_truth": syn.rotation_error(T, ic.T_true), "trans_err_vs_truth": syn.translation_error(T, ic.T_true),
                    "cpu_baseline": icp_cpu}
        # HBM roofline of the iteration: compulsory bytes 16 Ns + 32 Nt + 16 slots over the measured time
        comp = 16.0 * ic.source.shape[0] + 32.0 * ic.target.shape[0] + 16.0 * 262144
        out[key]["roofline"] = {"bound": "hbm", "achieved": comp / (ms / ic.iterations * 1e-3) / 1e9, "peak": 6552.0, "unit": "GB/s",
                                "frac": comp / (ms / ic.iterations * 1e-3) / 1e9 / 6552.0,
                                "note": "latency / divergence bound at this size (9-12 MB per iteration), as SURVEY.md 8d predicted"}
    ctx.set_icp_mode(0)
    out["demo_scene"] = demo_scene_full(ctx, b3d)
    # full registration with the reference's default budget: H = 100 000, confidence 0.999, ICP <= 200 iterations
    tgt_normals = case.target_normals
    keep = [pinned(case.source), pinned(case.target), pinned(case.source_desc), pinned(case.target_desc), pinned(tgt_normals)]
    h_src, h_tgt, h_sd, h_td, h_n = (k[1] for k in keep)

    def full():
        T0, f0, r0, _ = ctx.ransac(h_src, h_tgt, h_sd, h_td, case.voxel_size, 100_000, 0.999)
        return ctx.icp(h_src, h_tgt, h_n, T0, case.voxel_size * 0.4, 200, True), (f0, r0)

    def timed():
        full(); full()
        torch.cuda.synchronize()
        tt = []
        for _ in range(5):
            flush.zero_(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            res = full()
            tt.append(time.perf_counter() - t0)
        return tt, res

    ctx.set_score_mode(3)
    t_bail, res_bail = timed()
    ctx.set_score_mode(0)
    t, ((T, fit, rmse, iters), (f0, r0)) = timed()
    assert np.array_equal(res_bail[0][0], T), "bail-out scoring changed the registration result"
    out["registration"] = {"workload": "1M-point scene -> 100k source points vs 100k model: ransacRegistration(H=100000, conf 0.999) + "
                                       "icpRefine(thr 0.4*voxel, <=200 it, point-to-plane, default reference-order mode), host buffers in, pose out",
                           "ms": 1e3 * float(np.median(t)), "ms_with_bailout_scoring": 1e3 * float(np.median(t_bail)), "ransac_fitness": f0, "icp_fitness": fit, "icp_iterations": iters,
                           "rot_err_vs_truth": syn.rotation_error(T, case.T_true),
                           "trans_err_vs_truth": syn.translation_error(T, case.T_true)}
    return out


def whole_pipeline(b3d, bdist, syn, flush, rank, world, barrier, n_raw=1_000_000, voxel=0.0037, H=100_000):
    """BASELINE.json's 'end-to-end registration ms': raw 1M-point scene -> voxelDownsample -> estimateNormals -> computeFPFH ->
    ransacRegistration(H=100000, conf 0.999) -> icpRefine, against a ~100k-point model prepared once (as Pipeline::run prepares
    the reference model).  Real FPFH descriptors of a rough torus (no synthetic histograms), host buffer in, pose out.
    One b3d_register_scene_sharded call per scene on every rank (world 1: identical to b3d_register_scene): the front end runs
    on every rank, matching rows and hypothesis ids are split over the ranks, the refinement is replicated."""
    import torch
    import torch.distributed as dist
    rng = np.random.default_rng(1234 + 5)
    model_raw = syn.rough_torus(n_raw, rng)
    T_true = syn.rigid([0.2, 0.9, -0.3], 25.0, [0.05, -0.03, 0.08])
    scene_raw = (syn.apply(np.linalg.inv(T_true), syn.rough_torus(n_raw, rng)) + rng.normal(0, 0.0003, (n_raw, 3))).astype(np.float32)
    keep, h_scene = pinned(scene_raw)
    c = b3d.Context(torch.cuda.current_device())
    try:
        bdist.init_comm(c)
        n_model = c.prepare_model(model_raw, voxel)
        res = {}
        for mode, key in ((0, "ms"), (3, "ms_with_bailout_scoring")):
            c.set_score_mode(mode)
            c.register_scene_sharded(h_scene, voxel, ransac_max_iterations=H); c.register_scene_sharded(h_scene, voxel, ransac_max_iterations=H)
            tt = []
            for _ in range(5):
                flush.zero_()
                barrier()
                t0 = time.perf_counter()
                out = c.register_scene_sharded(h_scene, voxel, ransac_max_iterations=H)
                t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
                if world > 1:
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                tt.append(float(t.item()))
            res[key] = 1e3 * float(np.median(tt))
            res["result_" + key] = out
        a, b = res.pop("result_ms"), res.pop("result_ms_with_bailout_scoring")
        assert np.array_equal(a["refined"][0], b["refined"][0]), "bail-out scoring changed the pipeline result"
        names = ["match", "ransac_prepare", "score", "select_finish", "icp_grid", "icp_iterations", "icp_binning", "voxel_downsample",
                 "normals", "fpfh"]
        T = a["refined"][0]
        if rank == 0 and world == 1:
            res["cpu_port_estimate"] = pipeline_cpu_estimate(scene_raw, voxel, a["n_source_points"], n_model, H)
        res.update({"workload": f"raw {n_raw}-point scene -> {a['n_source_points']} points vs {n_model}-point model (voxel {voxel}); one "
                                "b3d_register_scene_sharded call = voxelDownsample + estimateNormals(30) + computeFPFH(5*voxel) + ransacRegistration"
                                f"(H={H}, conf 0.999) + icpRefine(0.4*voxel, <=200 it, default reference-order mode); real FPFH descriptors; "
                                f"pinned host buffer in, pose out; {world} GPU(s), max over ranks",
                    "n_gpus": world,
                    "stages_ms_bailout_run": {n: c.stage_ms(i) for i, n in enumerate(names)},
                    "h2d_bytes": int(scene_raw.nbytes), "ransac_fitness": a["coarse"][1], "icp_fitness": a["refined"][1],
                    "icp_iterations": a["refined"][3], "rot_err_vs_truth": syn.rotation_error(T, T_true),
                    "trans_err_vs_truth": syn.translation_error(T, T_true)})
        c.comm_destroy()
        return res
    finally:
        c.close()


def stress_scene(b3d, bdist, syn, flush, rank, world, barrier, H=100_000, voxel=0.004):
    """BASELINE.json configs[4]: a 10M-point depth-derived scene (2560 x 4096 synthetic depth map through the pinhole model of
    pipeline.cpp:68-83) with 30 % uniformly random outlier pixels; full FPFH + RANSAC + ICP against a ~110k-point model of the
    clean surface seen from another pose; sharded over the ranks like the pipeline figure.  Also the HBM roofline of the two
    grid stages at this scale: building the voxel-hash grid over the down-sampled cloud and one nearest-neighbour pass."""
    import torch
    import torch.distributed as dist
    rng = np.random.default_rng(1234 + 4)
    h, w, f = 2560, 4096, 3000.0
    uu, vv = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32))
    zz = (0.9 + 0.10 * np.sin(uu / 310.0) * np.cos(vv / 270.0) + 0.05 * np.cos((uu + 2 * vv) / 190.0)
          + 0.03 * np.sin(uu / 67.0 + 1.0) * np.sin(vv / 83.0) + 0.012 * np.cos(uu / 23.0) * np.cos(vv / 29.0))
    depth = np.round(zz * 1000.0).astype(np.uint16)
    clean = depth.copy()
    outl = rng.random((h, w)) < 0.30
    depth[outl] = rng.integers(300, 1500, int(outl.sum())).astype(np.uint16)
    args = (1000.0, 1.5, f, f, w / 2.0, h / 2.0)
    c = b3d.Context(torch.cuda.current_device())
    try:
        bdist.init_comm(c)
        cloud, _ = c.depth_to_cloud(depth, None, *args)
        surf, _ = c.depth_to_cloud(clean[::3, ::3].copy(), None, 1000.0, 1.5, f / 3, f / 3, w / 6.0, h / 6.0)
        T_true = syn.rigid([0.3, 0.2, 0.9], 15.0, [0.04, -0.02, 0.05])
        n_model = c.prepare_model(syn.apply(T_true, surf), voxel)
        keep, h_cloud = pinned(cloud)
        c.register_scene_sharded(h_cloud, voxel, ransac_max_iterations=H, icp_max_iterations=30)
        tt = []
        for _ in range(3):
            flush.zero_()
            barrier()
            t0 = time.perf_counter()
            out = c.register_scene_sharded(h_cloud, voxel, ransac_max_iterations=H, icp_max_iterations=30)
            t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            tt.append(float(t.item()))
        names = ["match", "ransac_prepare", "score", "select_finish", "icp_grid", "icp_iterations", "icp_binning", "voxel_downsample", "normals", "fpfh"]
        res = {"workload": f"configs[4]: {cloud.shape[0]}-point depth-derived scene, 30 % outlier pixels -> {out['n_source_points']} points (voxel {voxel}) vs "
                           f"{n_model}-point model; one b3d_register_scene_sharded call (H={H}, conf 0.999, ICP <= 30 it), pinned host cloud in, "
                           f"pose out; {world} GPU(s), max over ranks",
               "n_gpus": world, "ms": 1e3 * float(np.median(tt)), "h2d_bytes": int(cloud.nbytes),
               "stages_ms": {n: c.stage_ms(i) for i, n in enumerate(names)},
               "ransac_fitness": out["coarse"][1], "icp_fitness": out["refined"][1], "icp_iterations": out["refined"][3]}
        if rank == 0:
            # grid stages at this scale, resident clouds: source = the down-sampled scene (~0.5M points) against itself as target
            down, _ = c.voxel_downsample(cloud, voxel)
            n = down.shape[0]
            c.set_clouds(down, down)
            I = np.eye(4, dtype=np.float32)
            c.icp_nearest(I, voxel * 1.5); c.icp_nearest(I, voxel * 1.5)
            g_ms = c.stage_ms(4)
            c.set_icp_mode(1)
            c.icp_run(I, voxel * 1.5, 3, False, False)
            c.icp_run(I, voxel * 1.5, 1, False, False)
            it_ms = c.stage_ms(5)
            c.set_icp_mode(0)
            res["grid_roofline"] = {
                "points": int(n),
                "grid_build": {"bound": "hbm", "achieved": 40.0 * n / (g_ms * 1e-3) / 1e9, "peak": 6552.0, "unit": "GB/s", "frac": 40.0 * n / (g_ms * 1e-3) / 1e9 / 6552.0,
                               "ms": g_ms, "algorithmic_bytes_per_point": 40},
                "nn_iteration": {"bound": "hbm", "achieved": (16.0 * n + 32.0 * n) / (it_ms * 1e-3) / 1e9, "peak": 6552.0, "unit": "GB/s",
                                 "frac": (16.0 * n + 32.0 * n) / (it_ms * 1e-3) / 1e9 / 6552.0, "ms": it_ms,
                                 "algorithmic_bytes": "16 B per source point + 32 B per target point (first touch)"},
                "note": "both stages are latency / atomics bound, far from the HBM roofline: the grid build is seven small dependent launches "
                        "(bounds, clear, insert with warp-aggregated atomics, 3-kernel scan, scatter); the search is divergent hash probing"}
        c.comm_destroy()
        return res
    finally:
        c.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
