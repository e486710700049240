#!/usr/bin/env python
"""bench.py -- 64^3 DC chunks/sec (density -> Hermite -> active voxels -> QEF -> mesh).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Headline (`value`, `e2e`): BASELINE.json configs[1], the clipmap LOD0 ring of 8x8x8 = 512 chunks of
64^3 default noise terrain.  With N GPUs every rank meshes THE SAME 512-chunk ring (identical load
on every rank), no data-path collective: weak scaling, value = N * chunks / max-over-ranks device time.
A driver "step" is `batches_per_step` passes of the hot path over the ring, chosen so that each
timed region lasts at least half a second; every batch is bracketed by its own pair of CUDA events
(L2 flushed before it, outside the events) and the line carries min / median / max per batch.

  value   : kernels + the per-batch count read-back, results left resident in HBM
  e2e     : the same pass through the host-facing C ABI call (lvn_meshgen_generate_batch): chunk
            list uploaded, every mesh / seam arena copied back into pinned host memory, all
            inside the timed region
  roofline: the dominant kernel (Hermite, FP32-bound), algorithmic flops of SURVEY.md 8(d)
            over its CUDA-event time, against the FP32 peak measured in the same run
  stages  : every kernel of the step with its bound, algorithmic work and fraction of the roof
  configs : the other BASELINE configurations on the same line --
            sweep  (configs[4]): the 4096-chunk world sharded round-robin (chunk i -> GPU i mod N),
                   STRONG scaling: per-rank batch + the count gather over NCCL inside the timed region
            csg    (configs[2]): the 32-op edit script on the ring's fields, one op per step (N = 1)
            stress (configs[3]): the 64-chunk dense 3-D field (N = 1)
  cpu_baseline / --impl reference: the reference's own kernels (leven/cl/*.cl) compiled for the
            host cores through oracle/ref_shim (kind "reference"); the C restatement
            (oracle/lvn_oracle.c, kind "port") where that library was not built
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import leven_b200.workloads as W   # noqa: E402  (numpy only: the chunk lists of the five configs)

SEED, V, SIZE, CY0, RING = W.SEED, W.V, W.SIZE, W.CY0, W.RING
FLOP_PER_DENSITY = 1217   # one DensityFunc of the default terrain
# config 4's density (density.cuh: stress_density), counted like SURVEY.md 8(d) counts the terrain
# (FMA = 2, every other fp op / compare / select / floor = 1): snoise3 = 121 (skew 3, floor 6,
# unskew 9, simplex ordering 3 + 6 + 15, four corners 4 x 15 + 15 for their offsets, final sum 4);
# one octave = snoise3 + 9; four octaves + scale + threshold = 3 + 4 * 130 + 2
FLOP_PER_STRESS_DENSITY = 525
METRIC = "64^3 DC chunks/sec (density->QEF->mesh)"
MIN_REGION_S = 0.5        # every timed region lasts at least this long


def ring_chunks(rank=0):
    """configs[1]; every rank takes the same ring (rank is ignored: kept for the scripts under profiles/)"""
    return W.ring_chunks()


def workload_name():
    return (f"configs[1]: clipmap LOD0 ring, {RING}x{RING}x{RING}={RING ** 3} chunks of 64^3 default noise terrain "
            f"(seed {SEED}) per GPU, one batch per pass")


def config_block(arm):
    """`config` carries the same keys in both arms (the driver compares them)"""
    if arm == "ours":
        return {"workload": workload_name(),
                "l2": "flushed before every timed batch (256 MiB device memset, outside the batch's events)",
                "timing": "CUDA events on the launching stream around every batch, summed; max over ranks",
                "sharding": "the same 512-chunk ring on every GPU, no collective on the data path (weak scaling); "
                            "configs.sweep is the round-robin 4096-chunk split (strong scaling)"}
    return {"workload": workload_name(),
            "l2": "n/a (host cores)",
            "timing": "time.perf_counter around the steps, rank 0 only",
            "sharding": "none: the reference has one device and no multi-device path (compute.cpp:147)"}


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
                                          "--format=csv,noheader,nounits", "-lms", "50"],
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


# ---------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the only places that execute oracle/
# ---------------------------------------------------------------------------------------------
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
        "config": config_block("reference"),
        "cpu_baseline": {"value": value, "unit": "chunks/s", "cores": int(threads), "kind": kind, "sample": what},
        "e2e": {"value": value, "unit": "chunks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# stage accounting: algorithmic work per launch from the batch's actual counts
# ---------------------------------------------------------------------------------------------
def spread(xs):
    xs = np.asarray(xs, np.float64)
    return {"min": float(xs.min()), "median": float(np.median(xs)), "max": float(xs.max()), "n": int(len(xs))}


def stage_table(stats, K, nchunks, col_sets, kind, fp32_peak, hbm_peak, prof=None, sm_mhz=None):
    """kind: "terrain" (default terrain: columns / rows / hermite / leaves) or "stress" (3-D density:
    field / rows / hermite / leaves).  Counts are per launch (= per batch: K profiled batches)."""
    prof = prof or {}
    E, Ey, N = stats["edges"] / K, stats["edgesY"] / K, stats["nodes"] / K
    T, S, NE = stats["triangles"] / K, stats["seamNodes"] / K, stats["nonEmptyChunks"] / K
    Q = T / 2
    F, H, Vv = V + 2, V + 1, V
    F3, F2 = F ** 3, F ** 2
    ms = {k: v / K for k, v in stats["ms"].items()}

    def gbs(b, t):
        return b / (t * 1e-3) / 1e9 if t > 0 else 0.0

    def tfl(f, t):
        return f / (t * 1e-3) / 1e12 if t > 0 else 0.0

    stages = []
    if kind == "terrain":
        col_exec = col_sets * F2 * FLOP_PER_DENSITY
        stages.append({"stage": "columns (S1)", "kernel": "k_columns", "bound": "fp32", "ms": ms["columns"],
                       "algorithmic_flop": nchunks * (F2 * FLOP_PER_DENSITY + 2 * F3), "executed_flop": col_exec,
                       "achieved_tflops": tfl(col_exec, ms["columns"]), "peak_tflops": fp32_peak,
                       "note": "chunks of one vertical stack share a column set: the kernel executes colsets*F^2 Terrain "
                               "evaluations, not chunks*F^2; the fraction is taken on the EXECUTED flops"})
        dens_flop = FLOP_PER_DENSITY
        herm_flop = (5 * Ey + 21 * (E - Ey)) * dens_flop + 60 * E
        herm_note = "(5*E_y + 21*(E_x+E_z))*1217 + 60*E flop per launch (SURVEY.md 8d)"
        # S2+S4 as SURVEY.md 8(d) states it (u8 field read twice) and restated for this representation: the sign
        # rows are built from the column heights, so a chunk with a surface reads 4F^2 B of heights and writes
        # 12F^2 B of sign rows + 4H^2 + 3*4V^2 B of row offsets; a chunk without one costs two loads
        rows_survey = nchunks * 2 * F3 + 4 * E + 12 * N
        rows_restated = NE * (4 * F2 + 12 * F2 + 4 * H * H + 12 * Vv * Vv)
    else:
        dens_flop = FLOP_PER_STRESS_DENSITY
        stages.append({"stage": "field (S1, 3-D density)", "kernel": "k_field_density", "bound": "fp32", "ms": ms["field"],
                       "algorithmic_flop": nchunks * F3 * (dens_flop + 1), "achieved_tflops": tfl(nchunks * F3 * (dens_flop + 1), ms["field"]),
                       "peak_tflops": fp32_peak})
        herm_flop = 23 * E * dens_flop + 60 * E
        herm_note = f"23*E*C_d + 60*E flop per launch, C_d = {dens_flop} (SURVEY.md 8d: general 3-D density)"
        rows_survey = nchunks * 2 * F3 + 4 * E + 12 * N
        rows_restated = nchunks * (F3 + 12 * F2 + 4 * H * H + 12 * Vv * Vv)      # the u8 field is read once
    stages.append({"stage": "rows (S2+S4)", "kernel": "k_rows", "bound": "hbm", "ms": ms["classify"],
                   "algorithmic_bytes_survey": rows_survey, "algorithmic_bytes": rows_restated,
                   "achieved_gbs": gbs(rows_restated, ms["classify"]), "achieved_gbs_survey_count": gbs(rows_survey, ms["classify"]),
                   "peak_gbs": hbm_peak,
                   "note": "frac is taken on the bytes of THIS representation (bit-packed sign rows built from column heights / "
                           "the u8 field); the SURVEY.md 8(d) count (F^3 B read twice per chunk) is reported beside it"})
    stages.append({"stage": "hermite (S3)", "kernel": "k_hermite_terrain" if kind == "terrain" else "k_hermite_density",
                   "bound": "fp32", "ms": ms["hermite"], "algorithmic_flop": herm_flop, "algorithmic": herm_note,
                   "achieved_tflops": tfl(herm_flop, ms["hermite"]), "peak_tflops": fp32_peak})
    # S5 (unfused: 8N + 20E read, 80N written) + S8 + S9 + S10 for k_leaves; S6 for k_solve
    leaves_bytes = (8 * N + 20 * E + 80 * N) + (8 * N + 72 * N + 24 * Q) + (36 * N + 48 * N) + (4 * N + 36 * S + 48 * S)
    stages.append({"stage": "leaves (S5+S8+S9+S10)", "kernel": "k_leaves", "bound": "hbm", "ms": ms["leaves"],
                   "algorithmic_bytes": leaves_bytes, "achieved_gbs": gbs(leaves_bytes, ms["leaves"]), "peak_gbs": hbm_peak})
    stages.append({"stage": "solve (S6)", "kernel": "k_solve", "bound": "fp32", "ms": ms["solve"],
                   "algorithmic_flop": 1670 * N, "algorithmic": "1670 flop per node (SURVEY.md 8d); read 64N, write 16N B",
                   "achieved_tflops": tfl(1670 * N, ms["solve"]), "peak_tflops": fp32_peak,
                   "algorithmic_bytes": 80 * N, "achieved_gbs": gbs(80 * N, ms["solve"])})
    issue_peak = 4 * 148 * (sm_mhz or 1965.0) * 1e6
    for s in stages:
        if s["bound"] == "fp32":
            s["frac"] = s["achieved_tflops"] / fp32_peak if fp32_peak else None
        else:
            s["frac"] = s["achieved_gbs"] / hbm_peak
        key = {"k_columns": "columns", "k_rows": "rows", "k_hermite_terrain": "hermite", "k_leaves": "leaves", "k_solve": "solve"}.get(s["kernel"], "")
        inst = prof.get(key + "_warp_inst_per_launch")
        if inst and s["ms"] > 0 and kind == "terrain":
            s["warp_inst_per_launch_ncu"] = inst
            s["issue_frac"] = inst / (s["ms"] * 1e-3) / issue_peak
        if kind == "terrain" and prof.get(key + "_dram_bytes_per_launch") is not None:
            s["traffic_dram_bytes_ncu"] = prof[key + "_dram_bytes_per_launch"]
    counts = {"edges": E, "edges_y": Ey, "vertices": N, "triangles": T, "seam_nodes": S, "non_empty_chunks": NE}
    return stages, counts


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="headline only (skip the sweep / csg / stress blocks)")
    ap.add_argument("--region-seconds", type=float, default=MIN_REGION_S)
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
    from leven_b200 import sharding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world_size > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # one process per GPU: give every rank its own share of the host cores (the box's 8 GPUs hang off
    # one virtual socket with no NUMA placement to choose, profiles/r01r_notes.md; LVN_PIN_CORES=0 disables)
    pinned_cores = None
    if world_size > 1 and os.environ.get("LVN_PIN_CORES", "1") != "0":
        try:
            allowed = sorted(os.sched_getaffinity(0))
            per = max(1, len(allowed) // world_size)
            mine = allowed[local_rank * per:(local_rank + 1) * per] or allowed
            os.sched_setaffinity(0, mine)
            pinned_cores = [mine[0], mine[-1]]
        except Exception:
            pinned_cores = None

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

    def gather_ranks(vals):
        mine = torch.tensor([float(v) for v in vals], device=dev, dtype=torch.float64)
        out = [mine.clone() for _ in range(world_size)]
        if world_size > 1:
            dist.all_gather(out, mine)
        return [[round(float(v), 4) for v in t.tolist()] for t in out]

    assert lc.Compute_SetDevice(local_rank) == 0
    numa = (-1, 0, False)
    if os.environ.get("LVN_NUMA_BIND", "1") != "0" and pinned_cores is None:
        numa = lc.Compute_BindHostNuma(local_rank)
    rc = lc.Compute_Initialise(SEED, 0, 2)
    assert rc == 0, f"Compute_Initialise: {lc.GetCLErrorString(rc)} {lc.last_cuda_error()}"
    ctx = lc.Compute_MeshGenContext.create(V)
    assert ctx.privateCtx_
    stream = torch.cuda.current_stream()
    ctx.setStream(stream.cuda_stream)        # so that torch.cuda.Event brackets the kernels
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    peaks, peak_src = measured_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))

    def pinned(n, dtype):
        t = torch.empty(max(n, 1) * dtype.itemsize, dtype=torch.uint8, pin_memory=True)
        return t, t.numpy().view(dtype)

    def timed_batches(batch_fn, count):
        """`count` batches; L2 flushed before each; a CUDA event pair on the launching stream around each.
        Returns (per-batch device ms, wall seconds incl. the flushes)."""
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(count)]
        barrier()
        t0 = time.perf_counter()
        for a, b in evs:
            flush.zero_()
            a.record(stream)
            batch_fn()
            b.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        return np.array([a.elapsed_time(b) for a, b in evs]), wall

    def batches_for(batch_fn, steps, warm):
        """batches per step so that steps * batches last >= region-seconds (from a short untimed probe,
        agreed between the ranks)"""
        for _ in range(warm):
            flush.zero_()
            batch_fn()
        probe, _ = timed_batches(batch_fn, 3)
        est = max_over_ranks(float(np.median(probe)))
        return max(1, int(math.ceil(args.region_seconds * 1e3 / (max(steps, 1) * max(est, 1e-3)))))

    class Workload:
        """one chunk list on this rank: sizing pass, pinned host arenas, the two step functions"""

        def __init__(self, ms):
            self.ms = np.ascontiguousarray(ms, np.int32)
            rc, self.res, view = ctx.generateBatchDevice(self.ms)
            assert rc == 0, f"generateBatchDevice: {lc.GetCLErrorString(rc)} {lc.last_cuda_error()}"
            self.totV, self.totT, self.totS = int(view.totalVertices), int(view.totalTriangles), int(view.totalSeamNodes)
            self.keep = [pinned(self.totV + 1024, lc.MeshVertex), pinned(self.totT + 1024, lc.MeshTriangle),
                         pinned(self.totS + 1024, lc.SeamNodeInfo)]
            self.hostV, self.hostT, self.hostS = (k[1] for k in self.keep)
            self.h2d = len(self.ms) * 16                              # the caller's chunk list (4 ints per chunk)
            self.d2h = self.totV * 48 + self.totT * 12 + self.totS * 48 + len(self.ms) * 32   # arenas + per-chunk results

        def step_device(self):
            rc, res, _ = ctx.generateBatchDevice(self.ms)
            assert rc == 0
            return res

        def step_e2e(self):
            rc, res = ctx.generateBatch(self.ms, self.hostV, self.hostT, self.hostS)
            assert rc == 0
            return res

    # =========================================================================================
    # headline: configs[1], the same ring on every rank
    # =========================================================================================
    ms = ring_chunks()
    nchunks = len(ms)
    ring = Workload(ms)
    if rank == 0:
        origin = [i for i, m in enumerate(ms) if m[0] == 0 and m[1] == CY0 * SIZE and m[2] == 0][0]
        assert ring.res[origin]["numVertices"] > 0, "CY0 does not name the surface chunk above the origin"
    inner = batches_for(ring.step_device, args.steps, args.warmup)
    fp32_peak = lc.MeasureFP32Peak()

    ctx.getStats(reset=True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.mark()
    dev_batch_ms, wall = timed_batches(ring.step_device, args.steps * inner)
    run_stats = ctx.getStats(reset=True)
    pipe = ctx.getPipeline()
    for _ in range(2):
        ring.step_e2e()
    inner_e2e = batches_for(ring.step_e2e, args.steps, 0)
    e2e_batch_ms, e2e_wall = timed_batches(ring.step_e2e, args.steps * inner_e2e)
    pipe_e2e = ctx.getPipeline()
    sampler.mark()
    clocks = sampler.stop()

    # The same end-to-end step as a stream of batches: two contexts, two sets of host arenas, the counts-first
    # call (lvn_meshgen_generate_batch_async) for batch i while batch i - 1 still crosses PCIe, lvn_meshgen_wait
    # before a batch's arenas are touched.  Reported beside `e2e` (which stays one synchronous call per batch).
    pipelined = None
    try:
        ctx_b = lc.Compute_MeshGenContext.create(V)
        keep_b = [pinned(ring.totV + 1024, lc.MeshVertex), pinned(ring.totT + 1024, lc.MeshTriangle), pinned(ring.totS + 1024, lc.SeamNodeInfo)]
        pair = [(ctx, (ring.hostV, ring.hostT, ring.hostS)), (ctx_b, tuple(k[1] for k in keep_b))]
        for _ in range(3):
            assert ctx_b.generateBatch(ring.ms, *pair[1][1])[0] == 0

        def stream_of_batches(count):
            barrier()
            t0 = time.perf_counter()
            for i in range(count):
                c, a = pair[i % 2]
                assert c.generateBatchAsync(ring.ms, *a)[0] == 0
                if i:
                    assert pair[(i - 1) % 2][0].wait() == 0      # batch i - 1 is complete: its arenas may be read
            assert pair[(count - 1) % 2][0].wait() == 0
            barrier()
            return time.perf_counter() - t0

        stream_of_batches(20)
        n_pipe = max(20, int(0.5 / (float(np.median(e2e_batch_ms)) * 1e-3)))
        pipe_s = max_over_ranks(stream_of_batches(n_pipe))
        pipelined = {"value": world_size * nchunks * n_pipe / pipe_s, "unit": "chunks/s", "ms_per_batch": 1e3 * pipe_s / n_pipe,
                     "batches": n_pipe, "contexts_in_flight": 2, "timed_region_s": pipe_s,
                     "clock": "host wall clock between two device synchronisations (the batches of two contexts overlap, "
                              "so there is no per-batch event pair); no L2 flush between batches",
                     "api": "lvn_meshgen_generate_batch_async + lvn_meshgen_wait, alternating between two contexts"}
        ctx_b.destroy()
    except Exception as e:  # noqa: BLE001
        pipelined = {"error": repr(e)[:200]}
    # the link itself on this box: the step's download as one device -> pinned-host copy (outside the
    # timed regions; boxes of one pool differ here, and the end-to-end step is bound by it)
    link_bytes = ring.totV * 48 + ring.totT * 12 + ring.totS * 48
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
    # the same copy 24 times back to back on every rank at once: the SUSTAINED rate of this rank's link while the
    # others keep theirs busy too -- what the end-to-end step sees (a one-off copy hides shared uplinks)
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(24):
        link_host.copy_(flush[:link_bytes], non_blocking=True)
    b.record(stream)
    torch.cuda.synchronize()
    link_sustained_ms = a.elapsed_time(b) / 24.0
    barrier()
    # ---- per-kernel durations: the same batch with one lane on one stream, so that every kernel
    #      runs alone between its two CUDA events (in the timed regions above kernels of
    #      different lanes overlap and an event pair would also time its neighbours) ----
    ctx.setProfiling(True)
    ctx.getStats(reset=True)
    prof_batches = 50
    prof_ms, _ = timed_batches(ring.step_device, prof_batches)
    stats = ctx.getStats(reset=True)
    ctx.setProfiling(False)

    dev_ms = float(dev_batch_ms.sum())
    e2e_ms = float(e2e_batch_ms.sum())
    dev_ms_max = max_over_ranks(dev_ms)
    e2e_ms_max = max_over_ranks(e2e_ms)
    value = world_size * nchunks * args.steps * inner / (dev_ms_max * 1e-3)
    e2e_value = world_size * nchunks * args.steps * inner_e2e / (e2e_ms_max * 1e-3)
    per_rank = gather_ranks([e2e_ms / (args.steps * inner_e2e), dev_ms / (args.steps * inner), link_bytes / link_ms / 1e6,
                             link_bytes / link_sustained_ms / 1e6])

    # =========================================================================================
    # configs.sweep: configs[4], 4096 chunks, chunk i -> GPU i mod N, count gather inside the step
    # =========================================================================================
    blocks = {}
    if not args.no_configs:
        sweep = W.sweep_chunks()
        mine = sharding.shard_round_robin(len(sweep), rank, world_size)
        sw = Workload(sweep[mine])
        gather = sharding.CountGather(len(sweep), rank, world_size, device=dev if world_size > 1 else None)

        def sweep_device():
            # the counts are final after the classify kernels: the call returns with them while the Hermite / leaves /
            # solve kernels run, and the gather (on its own stream) crosses NVLink under them
            rc, res, _ = ctx.generateBatchDeviceAsync(sw.ms)
            assert rc == 0
            gather.gather(res["numVertices"], res["numTriangles"], res["numSeamNodes"])
            assert ctx.wait() == 0

        def sweep_e2e():
            # likewise: the call returns when every lane's counts are in and its copies are queued; the gather runs
            # while the last lanes compute and cross PCIe
            rc, res = ctx.generateBatchAsync(sw.ms, sw.hostV, sw.hostT, sw.hostS)
            assert rc == 0
            gather.gather(res["numVertices"], res["numTriangles"], res["numSeamNodes"])
            assert ctx.wait() == 0

        n_dev = batches_for(sweep_device, args.steps, 3) * args.steps
        sd_ms, _ = timed_batches(sweep_device, n_dev)
        n_e2e = batches_for(sweep_e2e, args.steps, 2) * args.steps
        se_ms, _ = timed_batches(sweep_e2e, n_e2e)
        counts, offsets, totals = gather.gather(sw.res["numVertices"], sw.res["numTriangles"], sw.res["numSeamNodes"])
        sd_max, se_max = max_over_ranks(float(sd_ms.sum())), max_over_ranks(float(se_ms.sum()))
        sweep_rank = gather_ranks([float(np.median(sd_ms)), float(np.median(se_ms)), float((sw.res["numEdges"] > 0).sum()),
                                   float(sw.d2h)])
        ctx.setProfiling(True)
        ctx.getStats(reset=True)
        timed_batches(sw.step_device, 20)
        sweep_stats = ctx.getStats(reset=True)
        ctx.setProfiling(False)
        if rank == 0:
            sweep_cols = len({(int(m[0]), int(m[2])) for m in sw.ms})
            sstages, scounts = stage_table(sweep_stats, 20, len(sw.ms), sweep_cols, "terrain", fp32_peak, hbm_peak)
            blocks["sweep"] = {
                "workload": "configs[4]: full-world sweep, 16x16x16 = 4096 LOD0 chunks, chunk linear index i -> GPU i mod N, one batch "
                            "per rank per pass + the (numVertices, numTriangles, numSeamNodes) gather that places every chunk in one "
                            "global mesh (12 B per chunk over NCCL, exclusive scan on the host)",
                "scaling": "strong", "chunks": len(sweep), "n_gpus": world_size,
                "value": len(sweep) * n_dev / (sd_max * 1e-3), "unit": "chunks/s", "ms_per_pass": sd_max / n_dev,
                "e2e": {"value": len(sweep) * n_e2e / (se_max * 1e-3), "unit": "chunks/s", "ms_per_pass": se_max / n_e2e,
                        "h2d_bytes_per_pass_per_rank": sw.h2d, "d2h_bytes_per_pass_per_rank": [r[3] for r in sweep_rank]},
                "per_batch_ms": {"device": spread(sd_ms), "e2e": spread(se_ms)},
                "per_rank": {"device_ms_median": [r[0] for r in sweep_rank], "e2e_ms_median": [r[1] for r in sweep_rank],
                             "non_empty_chunks": [int(r[2]) for r in sweep_rank]},
                "global_mesh": {"vertices": int(totals[0]), "triangles": int(totals[1]), "seam_nodes": int(totals[2]),
                                "non_empty_chunks": int((counts[:, 1] > 0).sum())},
                "passes_timed": {"device": n_dev, "e2e": n_e2e},
                "stages_rank0": sstages, "counts_rank0": scounts,
            }

    if rank != 0:
        if world_size > 1:
            dist.destroy_process_group()
        return

    # =========================================================================================
    # rank 0: the line
    # =========================================================================================
    prof = {}
    ppath = os.path.join(ROOT, "profiles", "latest_traffic.json")
    if os.path.exists(ppath):
        try:
            prof = json.load(open(ppath))
        except Exception:
            prof = {}
    col_sets = len({(int(m[0]), int(m[2])) for m in ms})
    stages, cnt = stage_table(stats, prof_batches, nchunks, col_sets, "terrain", fp32_peak, hbm_peak, prof, clocks.get("sm_mhz"))
    by = {s["kernel"]: s for s in stages}
    herm, leaves = by["k_hermite_terrain"], by["k_leaves"]
    stage_sum = sum(s["ms"] for s in stages)

    line = {
        "metric": METRIC, "value": value, "unit": "chunks/s", "n_gpus": world_size, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_block("ours"),
        "workload_detail": {"chunks_per_gpu": nchunks, "non_empty_chunks_per_gpu": cnt["non_empty_chunks"],
                            "edges_per_batch": cnt["edges"], "vertices_per_batch": cnt["vertices"],
                            "triangles_per_batch": cnt["triangles"], "seam_nodes_per_batch": cnt["seam_nodes"],
                            "batches_per_step": inner, "chunks_timed": world_size * nchunks * args.steps * inner,
                            "timed_region_s": dev_ms_max * 1e-3},
        "ms_per_batch": dev_ms_max / (args.steps * inner),
        "per_batch_ms": {"device": spread(dev_batch_ms), "e2e": spread(e2e_batch_ms)},
        "ms_per_batch_wall_incl_flush": 1e3 * wall / (args.steps * inner),
        "e2e": {"value": e2e_value, "unit": "chunks/s", "h2d_bytes_per_step": ring.h2d * inner_e2e, "d2h_bytes_per_step": ring.d2h * inner_e2e,
                "h2d_bytes_per_batch": ring.h2d, "d2h_bytes_per_batch": ring.d2h,
                "ms_per_step": e2e_ms_max / args.steps, "ms_per_batch": e2e_ms_max / (args.steps * inner_e2e),
                "batches_per_step": inner_e2e, "timed_region_s": e2e_ms_max * 1e-3,
                "link": {"d2h_copy_ms": link_ms, "gbs": link_bytes / link_ms / 1e6,
                         "frac_of_batch": link_ms / (e2e_ms_max / (args.steps * inner_e2e)),
                         "what": "the batch's download as ONE device -> pinned host copy, no kernels running, all ranks copying at the same "
                                 "time (rank 0's figure): the PCIe floor of the batch on this box at this number of GPUs"},
                "api": "lvn_meshgen_generate_batch (host chunk list in, pinned host mesh/seam arenas out)",
                "pipelined": pipelined},
        "per_rank": {"e2e_ms_per_batch": [r[0] for r in per_rank], "device_ms_per_batch": [r[1] for r in per_rank],
                     "link_gbs": [r[2] for r in per_rank], "link_gbs_sustained": [r[3] for r in per_rank],
                     "link_floor_ms_per_batch": [round(link_bytes / (r[3] * 1e6), 4) for r in per_rank],
                     "pinned_cores_rank0": pinned_cores},
        "gpu_launches": int(sum(run_stats["launches"].values())),
        "clocks": clocks,
        "roofline": {"kernel": "k_hermite_terrain (FindEdgeIntersectionInfo)", "bound": "fp32",
                     "achieved": herm["achieved_tflops"], "peak": fp32_peak, "unit": "TFLOP/s", "frac": herm["frac"],
                     "traffic": prof.get("hermite_dram_bytes_per_launch"),
                     "algorithmic": herm["algorithmic"],
                     "peak_source": "measured in this run: independent FMA chains on all SMs (lvn_measure_fp32_peak); "
                                    "nominal 74.4 TFLOP/s; MEASURED_PEAKS.json has no non-tensor FP32 figure",
                     "share_of_step": herm["ms"] / stage_sum if stage_sum > 0 else None,
                     "timing": f"CUDA events around the kernel, {prof_batches} single-lane batches (kernel alone on the GPU)"},
        "roofline_hbm": {"kernel": "k_leaves", "bound": "hbm", "achieved": leaves["achieved_gbs"], "peak": hbm_peak,
                         "unit": "GB/s", "frac": leaves["frac"], "traffic": prof.get("leaves_dram_bytes_per_launch"),
                         "peak_source": peak_src},
        "stages": stages,
        "serial_ms_per_batch": float(prof_ms.sum()) / prof_batches,
        "pipeline": {"device": {"lanes": pipe[0], "streams": pipe[1]}, "e2e": {"lanes": pipe_e2e[0], "streams": pipe_e2e[1]}},
        "host_numa": {"node": numa[0], "cpus_bound": numa[1], "memory_policy_set": numa[2]},
        "configs": blocks,
    }

    if world_size == 1 and not args.no_configs:
        try:
            blocks["csg"] = bench_csg(lc, ctx, torch, stream, flush, fp32_peak, hbm_peak)
        except Exception as e:      # a failed block must not cost the headline
            blocks["csg"] = {"error": repr(e)[:300]}
        try:
            blocks["stress"] = bench_stress(lc, ctx, torch, stream, fp32_peak, hbm_peak, timed_batches)
        except Exception as e:
            blocks["stress"] = {"error": repr(e)[:300]}

    if world_size == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O      # the checker's CPU port, timed as the reported baseline
        O.set_num_threads(os.cpu_count() or 1)
        world = O.World(image=lc.Compute_GetNoiseImage(), default_material=0, voxels_per_chunk=V)
        world.batch_counts(ms[:32])
        t0 = time.perf_counter()
        counts, threads = world.batch_counts(ms)
        dt = time.perf_counter() - t0
        assert np.array_equal(counts[:, 0], ring.res["numEdges"]) and np.array_equal(counts[:, 2], ring.res["numTriangles"]), \
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
            rvalue, rdt, rthreads, non_empty = time_reference_kernels(image, sample, 1, 0)
            line["cpu_baseline"] = {"value": rvalue, "unit": "chunks/s", "cores": int(rthreads), "kind": "reference",
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


def bench_csg(lc, ctx, torch, stream, flush, fp32_peak, hbm_peak):
    """configs[2]: the 32-op script on the ring's fields, one op per step: apply to the overlapping
    chunks (lvn_meshgen_apply_csg_operations_batch), re-mesh exactly those through the host-facing
    batch call, store the op.  Timed end to end per op with CUDA events on the context's stream
    (every call is synchronous, so the event interval is the call)."""
    ring = W.ring_chunks()
    keep = [torch.empty(n * sz, dtype=torch.uint8, pin_memory=True) for n, sz in ((400000, 48), (800000, 12), (100000, 48))]
    Vh, Th, Sh = keep[0].numpy().view(lc.MeshVertex), keep[1].numpy().view(lc.MeshTriangle), keep[2].numpy().view(lc.SeamNodeInfo)
    # pass 0: another script of the same kind (other centres and sizes) warms the pools, the context's buffers and
    # the fields of the region, and is cleared; pass 1, the fixed script, is reported -- every one of its ops is a
    # first application, i.e. a real edit of the fields (re-applying the same script would mostly change nothing)
    scripts = [[lc.CSGOperationInfo.make(*s) for s in W.csg_script(seed=777)], [lc.CSGOperationInfo.make(*s) for s in W.csg_script()]]
    apply_ms, mesh_ms, wall_ms, edits = [], [], [], 0
    ctx.getStats(reset=True)
    for rep, ops in enumerate(scripts):
        apply_ms, mesh_ms, wall_ms, edits = [], [], [], 0
        for op in ops:
            lo, hi = lc.CalcCSGOperationBounds(op)
            touched = W.touched_chunks(ring, lo, hi)
            if not len(touched):
                continue
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            t0 = time.perf_counter()
            e[0].record(stream)
            assert ctx.applyCSGOperationsBatch([op], touched) == 0
            e[1].record(stream)
            rc, r = ctx.generateBatch(touched, Vh, Th, Sh)
            assert rc == 0, lc.GetCLErrorString(rc)
            e[2].record(stream)
            torch.cuda.synchronize()
            wall_ms.append(1e3 * (time.perf_counter() - t0))
            assert lc.Compute_StoreCSGOperation(op, lo, hi) == 0
            apply_ms.append(e[0].elapsed_time(e[1])); mesh_ms.append(e[1].elapsed_time(e[2])); edits += len(touched)
        if rep == 0:
            lc.Compute_ClearCSGOperations()
            ctx.getStats(reset=True)
    st = ctx.getStats(reset=True)
    lc.Compute_ClearCSGOperations()
    F3 = (V + 2) ** 3
    total = float(np.sum(apply_ms) + np.sum(mesh_ms))
    # a16's compulsory traffic per edited chunk: the u8 field read + written once (2 F^3 B) and the touched-edge
    # bitmap (3 H^3 bits, set + read); its flops: K ops x F^3 x C_brush (25 sphere / 45 box)
    csg_bytes = edits * (2 * F3 + 2 * 3 * (V + 1) ** 3 / 8)
    return {"workload": "configs[2]: 32 scripted sphere/cube add/subtract ops on the 512-chunk ring's fields, one op per step; "
                        "re-mesh only the chunks whose AABB overlaps the op's bounds; every timed op is a first application "
                        "(a real edit) on a context warmed by another script of the same kind",
            "ops": len(apply_ms), "chunk_edits": edits,
            "e2e_ms_per_op": spread(np.asarray(apply_ms) + np.asarray(mesh_ms)), "wall_ms_per_op": spread(wall_ms),
            "apply_ms_per_op": spread(apply_ms), "remesh_ms_per_op": spread(mesh_ms),
            "value": edits / (total * 1e-3), "unit": "edited chunks/s (apply + re-mesh + download)",
            "ops_per_s": len(apply_ms) / (total * 1e-3),
            "launches": {k: int(v) for k, v in st["launches"].items() if v},
            "roofline": {"kernel": "k_csg_* (apply)", "bound": "hbm", "algorithmic_bytes": csg_bytes,
                         "achieved": csg_bytes / (float(np.sum(apply_ms)) * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": csg_bytes / (float(np.sum(apply_ms)) * 1e-3) / 1e9 / hbm_peak,
                         "note": "the interval is the whole apply call (uploads, kernels, hash-table rebuild, host waits), "
                                 "not a kernel alone: an op touches ~4 chunks, so the call is latency-bound, not bandwidth-bound"}}


def bench_stress(lc, terrain_ctx, torch, stream, fp32_peak, hbm_peak, timed_batches):
    """configs[3]: 64 chunks of the dense 3-D field (ridged fBm of snoise3, ~30 % active voxels).
    The density function is process-wide state (like the reference's program build options), so the
    terrain context is not used while it is switched; it is restored afterwards."""
    lc.Compute_SetDensityFunction(1, W.STRESS_THRESHOLD)
    try:
        ctx = lc.Compute_MeshGenContext.create(V)
        ctx.setStream(stream.cuda_stream)
        ms = W.stress_chunks()
        rc, res, view = ctx.generateBatchDevice(ms)
        assert rc == 0, lc.GetCLErrorString(rc)
        totV, totT, totS = int(view.totalVertices), int(view.totalTriangles), int(view.totalSeamNodes)
        keep = [torch.empty((totV + 1024) * 48, dtype=torch.uint8, pin_memory=True),
                torch.empty((totT + 1024) * 12, dtype=torch.uint8, pin_memory=True),
                torch.empty((totS + 1024) * 48, dtype=torch.uint8, pin_memory=True)]
        hv, ht, hs = keep[0].numpy().view(lc.MeshVertex), keep[1].numpy().view(lc.MeshTriangle), keep[2].numpy().view(lc.SeamNodeInfo)

        def step_device():
            assert ctx.generateBatchDevice(ms)[0] == 0

        def step_e2e():
            assert ctx.generateBatch(ms, hv, ht, hs)[0] == 0

        for _ in range(2):
            step_device()
        d_ms, _ = timed_batches(step_device, 10)
        step_e2e()
        e_ms, _ = timed_batches(step_e2e, 10)
        ctx.setProfiling(True); ctx.getStats(reset=True)
        timed_batches(step_device, 5)
        st = ctx.getStats(reset=True); ctx.setProfiling(False)
        stages, cnt = stage_table(st, 5, len(ms), 0, "stress", fp32_peak, hbm_peak)
        out = {"workload": "configs[3]: dense-surface stress field, 4x4x4 = 64 chunks, density = threshold - ridged 3-D fBm of snoise3 "
                           f"(4 octaves), threshold {W.STRESS_THRESHOLD}",
               "chunks": len(ms), "active_fraction": cnt["vertices"] / (len(ms) * 64.0 ** 3),
               "value": len(ms) / (float(np.median(d_ms)) * 1e-3), "unit": "chunks/s", "ms_per_batch": spread(d_ms),
               "e2e": {"value": len(ms) / (float(np.median(e_ms)) * 1e-3), "unit": "chunks/s", "ms_per_batch": spread(e_ms),
                       "d2h_bytes_per_batch": totV * 48 + totT * 12 + totS * 48 + len(ms) * 32, "h2d_bytes_per_batch": len(ms) * 16},
               "counts": cnt, "stages": stages}
        ctx.destroy()
        return out
    finally:
        lc.Compute_SetDensityFunction(0, 0.5)


if __name__ == "__main__":
    main()
