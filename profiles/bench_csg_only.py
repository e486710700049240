#!/usr/bin/env python
"""configs[2] alone (bench.py's csg block): python profiles/bench_csg_only.py   [LVN_LIB_VARIANT=<name> for an A/B]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import leven_b200.compute as lc
import leven_b200.workloads as W
import bench as B
assert lc.Compute_Initialise(W.SEED, 0, 2) == 0
ctx = lc.Compute_MeshGenContext.create(W.V)
out = B.bench_csg(lc, ctx, torch, torch.cuda.current_stream(), None, 71.7, 6532.2)
print(json.dumps({k: out[k] for k in ("ops", "chunk_edits", "e2e_ms_per_op", "apply_ms_per_op", "remesh_ms_per_op", "value", "launches")}))
# a third script (real edits again) under per-stage event timing: where an op's device time goes
ctx.setProfiling(True); ctx.getStats(reset=True)
ring = W.ring_chunks()
keep = [torch.empty(n * sz, dtype=torch.uint8, pin_memory=True) for n, sz in ((400000, 48), (800000, 12), (100000, 48))]
Vh, Th, Sh = keep[0].numpy().view(lc.MeshVertex), keep[1].numpy().view(lc.MeshTriangle), keep[2].numpy().view(lc.SeamNodeInfo)
nops = 0
for s in W.csg_script(seed=4242):
    op = lc.CSGOperationInfo.make(*s)
    lo, hi = lc.CalcCSGOperationBounds(op)
    touched = W.touched_chunks(ring, lo, hi)
    if not len(touched):
        continue
    assert ctx.applyCSGOperationsBatch([op], touched) == 0
    assert ctx.generateBatch(touched, Vh, Th, Sh)[0] == 0
    nops += 1
st = ctx.getStats(reset=True); ctx.setProfiling(False)
print(json.dumps({"profiled_ops": nops, "stage_us_per_op": {k: round(1e3 * v / nops, 1) for k, v in st["ms"].items() if v},
                  "launches_per_op": {k: round(v / nops, 1) for k, v in st["launches"].items() if v}}))
