#!/usr/bin/env python
"""Per-lane kernel timeline of one batch (LVN_TRACE=1) for a few lanes x streams settings."""
import os, sys
os.environ["LVN_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import leven_b200.compute as lc
import bench as B
assert lc.Compute_Initialise(B.SEED, 0, 2) == 0
ctx = lc.Compute_MeshGenContext.create(B.V)
ms = B.ring_chunks(0)
ctx.setPipeline(1, 1)
for _ in range(3):
    ctx.generateBatchDevice(ms)
import torch
rc, res, view = ctx.generateBatchDevice(ms)
def pinned(n, dtype):
    t = torch.empty(max(n, 1) * dtype.itemsize, dtype=torch.uint8, pin_memory=True)
    return t, t.numpy().view(dtype)
k1, hv = pinned(int(view.totalVertices) + 1024, lc.MeshVertex)
k2, ht = pinned(int(view.totalTriangles) + 1024, lc.MeshTriangle)
k3, hs = pinned(int(view.totalSeamNodes) + 1024, lc.SeamNodeInfo)
for lanes, streams in [(int(a), int(b)) for a, b in (x.split("x") for x in sys.argv[1:])]:
    ctx.setPipeline(lanes, streams)
    ctx.generateBatchDevice(ms)
    ctx.generateBatchDevice(ms)
    print("[lvn trace] host path", file=sys.stderr, flush=True)
    ctx.generateBatch(ms, hv, ht, hs)
    ctx.generateBatch(ms, hv, ht, hs)
