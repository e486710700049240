#!/usr/bin/env python
"""Sweep lanes x streams of the batch pipeline on the bench workload (wall clock per synchronous
call, L2 not flushed: relative numbers only)."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import leven_b200.compute as lc
import bench as B

def main():
    assert lc.Compute_Initialise(B.SEED, 0, 2) == 0
    ctx = lc.Compute_MeshGenContext.create(B.V)
    ms = B.ring_chunks(0)
    rc, res, view = ctx.generateBatchDevice(ms)
    assert rc == 0
    def pinned(n, dtype):
        t = torch.empty(max(n, 1) * dtype.itemsize, dtype=torch.uint8, pin_memory=True)
        return t, t.numpy().view(dtype)
    k1, hv = pinned(int(view.totalVertices) + 1024, lc.MeshVertex)
    k2, ht = pinned(int(view.totalTriangles) + 1024, lc.MeshTriangle)
    k3, hs = pinned(int(view.totalSeamNodes) + 1024, lc.SeamNodeInfo)
    rows = []
    for lanes in (1, 2, 4, 8, 16, 32):
        for streams in (1, 2, 3, 4):
            if lanes == 1 and streams > 1:
                continue
            ctx.setPipeline(lanes, streams)
            for _ in range(3):
                ctx.generateBatchDevice(ms); ctx.generateBatch(ms, hv, ht, hs)
            n = 30
            t0 = time.perf_counter()
            for _ in range(n):
                ctx.generateBatchDevice(ms)
            td = (time.perf_counter() - t0) / n
            t0 = time.perf_counter()
            for _ in range(n):
                rc, _ = ctx.generateBatch(ms, hv, ht, hs)
                assert rc == 0
            te = (time.perf_counter() - t0) / n
            rows.append({"lanes": lanes, "streams": streams, "device_ms": td * 1e3, "e2e_ms": te * 1e3})
            print(json.dumps(rows[-1]), flush=True)

if __name__ == "__main__":
    main()
