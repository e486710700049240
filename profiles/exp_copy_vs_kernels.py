#!/usr/bin/env python
"""Does a bulk copy in flight slow the chunk kernels down?  The device-resident ring batch, timed alone and while
a second stream copies 34 MB blocks back to back: device -> pinned host, pinned host -> device, device -> device."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import leven_b200.compute as lc
import leven_b200.workloads as W
assert lc.Compute_Initialise(W.SEED, 0, 2) == 0
ctx = lc.Compute_MeshGenContext.create(W.V)
ms = W.ring_chunks()
stream = torch.cuda.current_stream()
ctx.setStream(stream.cuda_stream)
side = torch.cuda.Stream()
src = torch.empty(34 << 20, dtype=torch.uint8, device="cuda")
dst = torch.empty(34 << 20, dtype=torch.uint8, pin_memory=True)
for lanes in ((2, 2), (1, 1), (8, 2)):
    ctx.setPipeline(*lanes)
    for _ in range(5):
        ctx.generateBatchDevice(ms)
    def run(n):
        evs = []
        for _ in range(n):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); ctx.generateBatchDevice(ms); b.record(stream); evs.append((a, b))
        torch.cuda.synchronize()
        return float(np.median([a.elapsed_time(b) for a, b in evs]))
    alone = run(30)
    src2 = torch.empty(34 << 20, dtype=torch.uint8, device="cuda")
    res = {}
    for kind, a, b, reps in (("D2H", dst, src, 400), ("H2D", src, dst, 400), ("D2D", src2, src, 4000)):
        with torch.cuda.stream(side):
            for _ in range(reps):
                a.copy_(b, non_blocking=True)        # ~0.6 ms each over PCIe: ~240 ms of continuous copying
        res[kind] = run(30)
        torch.cuda.synchronize()
    print(f"lanes x streams {lanes}: batch alone {alone:.3f} ms; with a copy in flight: " +
          ", ".join(f"{k} {v:.3f} ms" for k, v in res.items()))
