#!/usr/bin/env python
"""LVN_TRACE timeline of config 3's per-op work: apply to the touched chunks, re-mesh them (host path)."""
import os, sys, time
os.environ["LVN_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import leven_b200.compute as lc
import leven_b200.workloads as W
assert lc.Compute_Initialise(W.SEED, 0, 2) == 0
ctx = lc.Compute_MeshGenContext.create(W.V)
ring = W.ring_chunks()
keep = [torch.empty(n * sz, dtype=torch.uint8, pin_memory=True) for n, sz in ((400000, 48), (800000, 12), (100000, 48))]
Vh, Th, Sh = keep[0].numpy().view(lc.MeshVertex), keep[1].numpy().view(lc.MeshTriangle), keep[2].numpy().view(lc.SeamNodeInfo)
ops = [lc.CSGOperationInfo.make(*s) for s in W.csg_script()]
for rep in range(2):
    for i, op in enumerate(ops[:8]):
        lo, hi = lc.CalcCSGOperationBounds(op)
        touched = W.touched_chunks(ring, lo, hi)
        if not len(touched):
            continue
        t0 = time.perf_counter()
        assert ctx.applyCSGOperationsBatch([op], touched) == 0
        t1 = time.perf_counter()
        rc, r = ctx.generateBatch(touched, Vh, Th, Sh)
        t2 = time.perf_counter()
        if rep:
            print(f"[lvn trace] op {i}: {len(touched)} chunks, apply {1e6 * (t1 - t0):.0f} us, re-mesh {1e6 * (t2 - t1):.0f} us, "
                  f"{int(r['numVertices'].sum())} vertices", file=sys.stderr, flush=True)
        lc.Compute_StoreCSGOperation(op, lo, hi)
    lc.Compute_ClearCSGOperations()
