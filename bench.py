#!/usr/bin/env python
"""bench.py -- 64^3 DC chunks/sec (density -> Hermite -> active voxels -> QEF -> mesh).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the whole hot path over one batch: BASELINE.json configs[1], the
clipmap LOD0 ring of 8x8x8 = 512 chunks of 64^3 default noise terrain.  With N GPUs every rank
meshes its own 512-chunk ring (rank r is shifted by 8r chunks in x), no data-path collective:
weak scaling, value = N*512*K chunks / max-over-ranks device time.

  value   : kernels + the per-batch count read-back, results left resident in HBM
  e2e     : the same pass through the host-facing C ABI call (lvn_meshgen_generate_batch): chunk
            list uploaded, every mesh / seam arena copied back into pinned host memory, all
            inside the timed region
  roofline: the dominant kernel (Hermite, FP32-bound), algorithmic flops of SURVEY.md 8(d)
            over its CUDA-event time, against the FP32 peak measured in the same run
  cpu_baseline / --impl reference: the reference's own kernels (leven/cl/*.cl) compiled for the
            host cores through oracle/ref_shim (kind "reference"); the C restatement
            (oracle/lvn_oracle.c, kind "port") where that library was not built
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 93923590          # leven/default.cfg:8
V = 64
SIZE = 256               # LOD0 chunk: 64 voxels * LEAF_SIZE_SCALE
CY0 = 9                  # floor(h(0,0) / 64): h(0,0) = 598.49 voxels for this seed (checked below)
RING = 8                 # 8 x 8 x 8 chunks
FLOP_PER_DENSITY = 1217
METRIC = "64^3 DC chunks/sec (density->QEF->mesh)"


def ring_chunks(rank):
    h = RING // 2
    return np.array([[(cx + RING * rank) * SIZE, (CY0 + dy) * SIZE, cz * SIZE, SIZE]
                     for dy in range(-h, h) for cz in range(-h, h) for cx in range(-h, h)], np.int32)


def workload_name():
    return (f"configs[1]: clipmap LOD0 ring, {RING}x{RING}x{RING}={RING ** 3} chunks of 64^3 default noise terrain "
            f"(seed {SEED}) per GPU, one batch per step")


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)), "MEASURED_PEAKS.json"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None
        self.marks = []

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
            time.sleep(0.5)      # let nvidia-smi attach before the timed region starts
        except Exception:
            self.proc = None

    def mark(self):
        """wall-clock bounds of the timed regions: only samples taken inside count"""
        self.marks.append(time.time())

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        import datetime
        sm, mx, reasons = [], [], set()
        lo, hi = (min(self.marks), max(self.marks)) if len(self.marks) >= 2 else (0.0, float("inf"))
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                clk, cmax = float(f[1]), float(f[2])
            except ValueError:
                continue
            if ts < lo - 0.02 or ts > hi + 0.02:    # taken before / after the timed regions
                continue
            sm.append(clk); mx.append(cmax)
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def reference_sample(ms, n):
    """n chunks spread over the workload (every len/n-th chunk of the ring, offset so that x, y and z all vary)"""
    n = max(1, min(int(n), len(ms)))
    idx = (np.arange(n) * len(ms)) // n + (np.arange(n) % 8)
    return ms[np.minimum(idx, len(ms) - 1)]


def time_reference_kernels(image, sample, steps, warmup):
    """oracle/_ref: the reference's own OpenCL C kernels compiled for the host (oracle/ref_shim),
    driven through the reference host sequence, OpenMP over the work-items of every NDRange."""
    from oracle import ref as R
    threads = R.set_num_threads(os.cpu_count() or 1)      # torchrun exports OMP_NUM_THREADS=1
    rw = R.RefWorld(image, default_material=0)
    for _ in range(warmup):
        rw.generate_chunk_mesh([int(v) for v in sample[0][:3]], int(sample[0][3]))
    non_empty = 0
    t0 = time.perf_counter()
    for _ in range(steps):
        non_empty = 0
        for c in sample:
            r = rw.generate_chunk_mesh([int(v) for v in c[:3]], int(c[3]))
            non_empty += int(r["numNodes"] > 0)
    dt = time.perf_counter() - t0
    return len(sample) * steps / dt, dt, threads, non_empty


def run_reference(args, rank):
    """reference arm: the reference's CPU-side pipeline on all host threads, same workload; rank 0
    only.  oracle/_ref (the reference's kernel text compiled for the host: kind "reference") when
    it was built, else the oracle port (kind "port").  Each step is a bounded sample of the
    512-chunk ring, sized so that the whole run stays within about two minutes."""
    if rank != 0:
        return
    from oracle import oracle as O
    from oracle import ref as R
    ms = ring_chunks(0)
    if R.available():
        image = O.noise_image(SEED)
        _, dt1, _, _ = time_reference_kernels(image, ms[256:257], 1, 1)        # one chunk, warm
        per_step = max(1, min(32, int(120.0 / (max(args.steps, 1) * max(dt1, 1e-3)))))
        sample = reference_sample(ms, per_step)
        value, dt, threads, non_empty = time_reference_kernels(image, sample, args.steps, min(args.warmup, 1))
        kind = "reference"
        what = (f"{len(sample)} of the workload's {len(ms)} chunks per step (evenly spread, {non_empty} contain surface), "
                "leven/cl kernels compiled for the host (oracle/_ref), OpenMP over the work-items of each NDRange")
    else:
        O.set_num_threads(os.cpu_count() or 1)       # torchrun exports OMP_NUM_THREADS=1
        world = O.World(seed=SEED, default_material=0, voxels_per_chunk=V)
        threads = 1
        for _ in range(args.warmup if args.warmup < 2 else 1):      # one warm pass is enough on the CPU
            _, threads = world.batch_counts(ms[:64])
        t0 = time.perf_counter()
        for _ in range(args.steps):
            counts, threads = world.batch_counts(ms)
        dt = time.perf_counter() - t0
        value = len(ms) * args.steps / dt
        kind = "port"
        what = (f"all {len(ms)} chunks of the workload per step, OpenMP over chunks; "
                f"{int((counts[:, 1] > 0).sum())} contain surface (oracle/_ref not built)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "chunks/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name()},
        "cpu_baseline": {"value": value, "unit": "chunks/s", "cores": int(threads), "kind": kind, "sample": what},
        "e2e": {"value": value, "unit": "chunks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import leven_b200.compute as lc

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world_size > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world_size == 1:
            return float(x)
        t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world_size == 1:
            return float(x)
        t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    assert lc.Compute_SetDevice(local_rank) == 0
    # one process per GPU: keep this process and the pinned arenas it allocates below on the GPU's
    # socket (LVN_NUMA_BIND=0 leaves the scheduler's placement alone)
    numa = (-1, 0, False)
    if os.environ.get("LVN_NUMA_BIND", "1") != "0":
        numa = lc.Compute_BindHostNuma(local_rank)
    rc = lc.Compute_Initialise(SEED, 0, 2)
    assert rc == 0, f"Compute_Initialise: {lc.GetCLErrorString(rc)} {lc.last_cuda_error()}"
    ctx = lc.Compute_MeshGenContext.create(V)
    assert ctx.privateCtx_
    stream = torch.cuda.current_stream()
    ctx.setStream(stream.cuda_stream)        # so that torch.cuda.Event brackets the kernels
    ms = ring_chunks(rank)
    nchunks = len(ms)

    # sizing pass (also validates CY0: the origin stack's surface chunk is non-empty)
    rc, res, view = ctx.generateBatchDevice(ms)
    assert rc == 0, f"generateBatchDevice: {lc.GetCLErrorString(rc)} {lc.last_cuda_error()}"
    if rank == 0:
        origin = [i for i, m in enumerate(ms) if m[0] == 0 and m[1] == CY0 * SIZE and m[2] == 0][0]
        assert res[origin]["numVertices"] > 0, "CY0 does not name the surface chunk above the origin"
    totV, totT, totS = int(view.totalVertices), int(view.totalTriangles), int(view.totalSeamNodes)

    def pinned(n, dtype):
        t = torch.empty(max(n, 1) * dtype.itemsize, dtype=torch.uint8, pin_memory=True)
        return t, t.numpy().view(dtype)
    keepV, hostV = pinned(totV + 1024, lc.MeshVertex)
    keepT, hostT = pinned(totT + 1024, lc.MeshTriangle)
    keepS, hostS = pinned(totS + 1024, lc.SeamNodeInfo)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def timed_region(step_fn, steps):
        """K steps; L2 flushed before each; CUDA events on the launching stream around each step"""
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        t0 = time.perf_counter()
        for a, b in evs:
            flush.zero_()
            a.record(stream)
            step_fn()
            b.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        dev_ms = sum(a.elapsed_time(b) for a, b in evs)
        return dev_ms, wall

    def step_device():
        rc, _, _ = ctx.generateBatchDevice(ms)
        assert rc == 0

    def step_e2e():
        rc, _ = ctx.generateBatch(ms, hostV, hostT, hostS)
        assert rc == 0

    for _ in range(args.warmup):
        flush.zero_()
        step_device()
    fp32_peak = lc.MeasureFP32Peak()

    # ---- timed: device-resident (lanes pipelined over the context's streams) ----
    ctx.getStats(reset=True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.mark()
    dev_ms, wall = timed_region(step_device, args.steps)
    run_stats = ctx.getStats(reset=True)
    pipe = ctx.getPipeline()
    # ---- timed: end to end through the host-facing call ----
    for _ in range(2):
        step_e2e()
    e2e_ms, e2e_wall = timed_region(step_e2e, args.steps)
    pipe_e2e = ctx.getPipeline()
    # the link itself on this box: the step's download as one device -> pinned-host copy (outside the
    # timed regions; boxes of one pool differ here, and the end-to-end step is bound by it)
    link_bytes = (totV * lc.MeshVertex.itemsize + totT * lc.MeshTriangle.itemsize + totS * lc.SeamNodeInfo.itemsize)
    link_host = torch.empty(link_bytes, dtype=torch.uint8, pin_memory=True)
    link_ms = []
    barrier()          # every rank copies at the same time: the ranks of one box share its host side
    for _ in range(7):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        link_host.copy_(flush[:link_bytes], non_blocking=True)
        b.record(stream)
        torch.cuda.synchronize()
        link_ms.append(a.elapsed_time(b))
    link_ms = float(np.median(link_ms[2:]))
    sampler.mark()
    clocks = sampler.stop()
    # ---- per-kernel durations: the same step with one lane on one stream, so that every kernel
    #      runs alone between its two CUDA events (in the timed regions above kernels of
    #      different lanes overlap and an event pair would also time its neighbours) ----
    ctx.setProfiling(True)
    ctx.getStats(reset=True)
    prof_steps = max(1, min(args.steps, 50))
    prof_ms, _ = timed_region(step_device, prof_steps)
    stats = ctx.getStats(reset=True)
    ctx.setProfiling(False)

    dev_ms_max = max_over_ranks(dev_ms)
    e2e_ms_max = max_over_ranks(e2e_ms)
    total_chunks = sum_over_ranks(nchunks) * args.steps
    # every rank's own figures (rank order): end-to-end ms per step, device-resident ms per step, link GB/s
    mine = torch.tensor([e2e_ms / args.steps, dev_ms / args.steps, link_bytes / link_ms / 1e6], device=dev, dtype=torch.float64)
    per_rank = [mine.clone() for _ in range(world_size)]
    if world_size > 1:
        dist.all_gather(per_rank, mine)
    per_rank = [[round(float(v), 4) for v in t.tolist()] for t in per_rank]
    value = total_chunks / (dev_ms_max * 1e-3)
    e2e_value = total_chunks / (e2e_ms_max * 1e-3)

    if rank != 0:
        if world_size > 1:
            dist.destroy_process_group()
        return

    K = prof_steps
    E, Ey, N = stats["edges"] / K, stats["edgesY"] / K, stats["nodes"] / K
    T, S, NE = stats["triangles"] / K, stats["seamNodes"] / K, stats["nonEmptyChunks"] / K
    Q = T / 2
    F3, F2 = (V + 2) ** 3, (V + 2) ** 2
    peaks, peak_src = measured_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    stage_ms = {k: v / K for k, v in stats["ms"].items()}

    # algorithmic work per launch, SURVEY.md 8(d)
    herm_flop = (5 * Ey + 21 * (E - Ey)) * FLOP_PER_DENSITY + 60 * E
    herm_tflops = herm_flop / (stage_ms["hermite"] * 1e-3) / 1e12 if stage_ms["hermite"] > 0 else 0.0
    col_flop = nchunks * (F2 * FLOP_PER_DENSITY + 2 * F3)
    col_sets = len({(int(m[0]), int(m[2])) for m in ms})
    classify_bytes = nchunks * 2 * F3 + 4 * E + 12 * N                 # S2 + S4
    leaves_bytes = (8 * N + 20 * E + 32 * N) + (8 * N + 72 * N + 24 * Q) + (36 * N + 48 * N) + (4 * N + 36 * S + 48 * S)

    def gbs(b, ms_):
        return b / (ms_ * 1e-3) / 1e9 if ms_ > 0 else 0.0

    prof = {}
    ppath = os.path.join(ROOT, "profiles", "latest_traffic.json")
    if os.path.exists(ppath):
        try:
            prof = json.load(open(ppath))
        except Exception:
            prof = {}

    stages = [
        # chunks of one vertical stack share a column set: the kernel executes colsets*F^2 Terrain
        # evaluations, not chunks*F^2; the fraction is taken on the EXECUTED flops
        {"stage": "columns (S1)", "bound": "fp32", "ms": stage_ms["columns"], "algorithmic_flop": col_flop,
         "executed_flop": col_sets * F2 * FLOP_PER_DENSITY,
         "achieved_tflops": col_sets * F2 * FLOP_PER_DENSITY / (stage_ms["columns"] * 1e-3) / 1e12
         if stage_ms["columns"] > 0 else 0.0,
         "peak_tflops": fp32_peak,
         "note": "this interval starts when the copy engine has delivered the batch head and so includes the hand-over to the "
                 "compute engine; the kernel alone takes 14 us (ncu, profiles/r01t_summary.txt), 0.33 of the peak on its executed flops"},
        {"stage": "rows (S2+S4)", "bound": "hbm", "ms": stage_ms["classify"], "algorithmic_bytes": classify_bytes,
         "achieved_gbs": gbs(classify_bytes, stage_ms["classify"]), "peak_gbs": hbm_peak},
        {"stage": "hermite (S3)", "bound": "fp32", "ms": stage_ms["hermite"], "algorithmic_flop": herm_flop,
         "achieved_tflops": herm_tflops, "peak_tflops": fp32_peak},
        {"stage": "leaves (S5+S6+S8+S9+S10)", "bound": "hbm", "ms": stage_ms["leaves"], "algorithmic_bytes": leaves_bytes,
         "achieved_gbs": gbs(leaves_bytes, stage_ms["leaves"]), "peak_gbs": hbm_peak},
    ]
    # issue-slot view: executed warp instructions of the launch (ncu, profiles/latest_traffic.json)
    # over the time it took here, against 4 schedulers x 148 SMs x the SM clock seen under load
    sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6
    issue_peak = 4 * 148 * sm_hz
    for s, key in zip(stages, ("columns", "rows", "hermite", "leaves")):
        if s["bound"] == "fp32":
            s["frac"] = s["achieved_tflops"] / fp32_peak if fp32_peak else None
        else:
            s["frac"] = s["achieved_gbs"] / hbm_peak
        inst = prof.get(key + "_warp_inst_per_launch")
        if inst and s["ms"] > 0:
            s["warp_inst_per_launch_ncu"] = inst
            s["issue_frac"] = inst / (s["ms"] * 1e-3) / issue_peak

    h2d = nchunks * 16                                   # the caller's chunk list (4 ints per chunk)
    d2h = totV * 48 + totT * 12 + totS * 48 + nchunks * 32   # mesh + seam arenas + per-chunk results

    line = {
        "metric": METRIC, "value": value, "unit": "chunks/s", "n_gpus": world_size, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(), "chunks_per_gpu": nchunks, "non_empty_chunks_per_gpu": NE,
                   "edges_per_step": E, "vertices_per_step": N, "triangles_per_step": T, "seam_nodes_per_step": S,
                   "l2": "flushed before every timed step (256 MiB device memset, outside the step's events)",
                   "timing": "CUDA events on the launching stream around each step, summed; max over ranks",
                   "sharding": "one 512-chunk ring per GPU, no collective on the data path"},
        "ms_per_step_wall_incl_flush": 1e3 * wall / args.steps,
        "e2e": {"value": e2e_value, "unit": "chunks/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms_max / args.steps,
                "link": {"d2h_copy_ms": link_ms, "gbs": link_bytes / link_ms / 1e6, "frac_of_step": link_ms / (e2e_ms_max / args.steps),
                         "what": "the step's download as ONE device -> pinned host copy, no kernels running, all ranks copying at the same "
                                 "time (rank 0's figure): the PCIe floor of the step on this box at this number of GPUs"},
                "api": "lvn_meshgen_generate_batch (host chunk list in, pinned host mesh/seam arenas out)"},
        "per_rank": {"e2e_ms_per_step": [r[0] for r in per_rank], "device_ms_per_step": [r[1] for r in per_rank],
                     "link_gbs": [r[2] for r in per_rank]},
        "gpu_launches": int(sum(run_stats["launches"].values())),
        "clocks": clocks,
        "roofline": {"kernel": "k_hermite (FindEdgeIntersectionInfo)", "bound": "fp32",
                     "achieved": herm_tflops, "peak": fp32_peak, "unit": "TFLOP/s",
                     "frac": herm_tflops / fp32_peak if fp32_peak else None,
                     "traffic": prof.get("hermite_dram_bytes_per_launch"),
                     "algorithmic": "(5*E_y + 21*(E_x+E_z))*1217 + 60*E flop per launch (SURVEY.md 8d)",
                     "peak_source": "measured in this run: independent FMA chains on all SMs (lvn_measure_fp32_peak); "
                                    "nominal 74.4 TFLOP/s",
                     "share_of_step": stage_ms["hermite"] / sum(stage_ms[k] for k in ("columns", "classify", "hermite", "leaves")),
                     "timing": f"CUDA events around the kernel, {prof_steps} single-lane steps (kernel alone on the GPU)"},
        "roofline_hbm": {"kernel": "k_leaves", "bound": "hbm", "achieved": stages[3]["achieved_gbs"], "peak": hbm_peak,
                         "unit": "GB/s", "frac": stages[3]["frac"], "traffic": prof.get("leaves_dram_bytes_per_launch"),
                         "peak_source": peak_src},
        "stages": stages,
        "serial_ms_per_step": prof_ms / prof_steps,
        "pipeline": {"device": {"lanes": pipe[0], "streams": pipe[1]}, "e2e": {"lanes": pipe_e2e[0], "streams": pipe_e2e[1]}},
        "host_numa": {"node": numa[0], "cpus_bound": numa[1], "memory_policy_set": numa[2],
                      "what": "rank 0's binding to its GPU's NUMA node before the pinned arenas are allocated (lvn_compute_bind_host_numa; node < 0: the platform names none)"},
    }

    if world_size == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O      # the checker's CPU port, timed as the reported baseline
        O.set_num_threads(os.cpu_count() or 1)
        world = O.World(image=lc.Compute_GetNoiseImage(), default_material=0, voxels_per_chunk=V)
        world.batch_counts(ms[:32])
        t0 = time.perf_counter()
        counts, threads = world.batch_counts(ms)
        dt = time.perf_counter() - t0
        assert np.array_equal(counts[:, 0], res["numEdges"]) and np.array_equal(counts[:, 2], res["numTriangles"]), \
            "CUDA path and oracle disagree on the bench workload"
        port = {"value": nchunks / dt, "unit": "chunks/s", "cores": int(threads), "kind": "port",
                "sample": f"the full {nchunks}-chunk workload once ({dt:.1f} s), OpenMP over chunks; "
                          "C restatement of the reference pipeline (oracle/lvn_oracle.c)"}
        from oracle import ref as R
        if R.available():
            # the reference's own kernels on the host cores: a spread sample sized for ~12 s of CPU work
            image = lc.Compute_GetNoiseImage()
            rate, _, _, _ = time_reference_kernels(image, reference_sample(ms, 16), 1, 1)
            sample = reference_sample(ms, max(16, min(nchunks, int(12.0 * rate))))
            value, rdt, rthreads, non_empty = time_reference_kernels(image, sample, 1, 0)
            line["cpu_baseline"] = {"value": value, "unit": "chunks/s", "cores": int(rthreads), "kind": "reference",
                                    "sample": f"{len(sample)} of the workload's {nchunks} chunks once ({rdt:.1f} s; evenly spread, "
                                              f"{non_empty} contain surface), leven/cl kernels compiled for the host (oracle/_ref), "
                                              "OpenMP over the work-items of each NDRange"}
            line["cpu_port"] = port
        else:
            line["cpu_baseline"] = port
        world.close()
    print(json.dumps(line), flush=True)
    if world_size > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
