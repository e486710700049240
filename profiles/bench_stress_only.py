#!/usr/bin/env python
"""configs[3] alone (bench.py's stress block): python profiles/bench_stress_only.py   [LVN_LIB_VARIANT=<name> for an A/B]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import leven_b200.compute as lc
import leven_b200.workloads as W
import bench as B
assert lc.Compute_Initialise(W.SEED, 0, 2) == 0
ctx = lc.Compute_MeshGenContext.create(W.V)
import time
import numpy as np
stream = torch.cuda.current_stream()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed_batches(batch_fn, count):
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(count)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for a, b in evs:
        flush.zero_()
        a.record(stream)
        batch_fn()
        b.record(stream)
    torch.cuda.synchronize()
    return np.array([a.elapsed_time(b) for a, b in evs]), time.perf_counter() - t0


out = B.bench_stress(lc, ctx, torch, stream, 71.7, 6532.2, timed_batches)
print(json.dumps({"ms_per_batch": out["ms_per_batch"], "stages": [(s["kernel"], round(s["ms"], 4), round(s["frac"], 3)) for s in out["stages"]]}))
