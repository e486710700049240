#!/usr/bin/env python
"""Where a batch call's host time goes (LVN_TRACE=1): the ring (512 chunks) and one rank's share of the
sweep at 8 GPUs (512 chunks, ~48 with surface), device-resident and host path; plus the Python-side cost
of the call and of the count gather."""
import os, sys, time
os.environ["LVN_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import leven_b200.compute as lc
import leven_b200.workloads as W
from leven_b200 import sharding
assert lc.Compute_Initialise(W.SEED, 0, 2) == 0
ctx = lc.Compute_MeshGenContext.create(W.V)
for name, ms in (("ring", W.ring_chunks()), ("sweep shard 1/8", W.sweep_chunks()[::8].copy())):
    rc, res, view = ctx.generateBatchDevice(ms)
    t = [torch.empty((int(n) + 1024) * sz, dtype=torch.uint8, pin_memory=True) for n, sz in
         ((view.totalVertices, 48), (view.totalTriangles, 12), (view.totalSeamNodes, 48))]
    hv, ht, hs = t[0].numpy().view(lc.MeshVertex), t[1].numpy().view(lc.MeshTriangle), t[2].numpy().view(lc.SeamNodeInfo)
    for _ in range(3):
        ctx.generateBatchDevice(ms)
    print(f"[lvn trace] ==== {name}: device-resident", file=sys.stderr, flush=True)
    t0 = time.perf_counter()
    for _ in range(3):
        ctx.generateBatchDevice(ms)
    print(f"[lvn trace] python wall per call {1e6 * (time.perf_counter() - t0) / 3:.0f} us", file=sys.stderr, flush=True)
    ctx.generateBatch(ms, hv, ht, hs)
    print(f"[lvn trace] ==== {name}: host path", file=sys.stderr, flush=True)
    t0 = time.perf_counter()
    for _ in range(3):
        ctx.generateBatch(ms, hv, ht, hs)
    print(f"[lvn trace] python wall per call {1e6 * (time.perf_counter() - t0) / 3:.0f} us", file=sys.stderr, flush=True)
g = sharding.CountGather(4096, 0, 1)
t0 = time.perf_counter()
for _ in range(100):
    g.gather(res["numVertices"][:0], res["numTriangles"][:0], res["numSeamNodes"][:0]) if False else g.gather(np.zeros(4096, np.int32), np.zeros(4096, np.int32), np.zeros(4096, np.int32))
print(f"[lvn trace] CountGather (1 rank, host only) {1e4 * (time.perf_counter() - t0):.0f} us per call", file=sys.stderr)
