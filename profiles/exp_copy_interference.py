#!/usr/bin/env python
"""Do kernels slow down while an unrelated D2H copy streams over PCIe?  Per-stage event times of
single-lane batches with and without a concurrent 256 MB device->pinned-host copy loop."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import leven_b200.compute as lc
import bench as B
assert lc.Compute_Initialise(B.SEED, 0, 2) == 0
ctx = lc.Compute_MeshGenContext.create(B.V)
ms = B.ring_chunks(0)
for _ in range(3):
    ctx.generateBatchDevice(ms)
ctx.setProfiling(True)
dev = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
host = torch.empty(64 << 20, dtype=torch.uint8, pin_memory=True)
side = torch.cuda.Stream()

def run(label, copying, direction="d2h"):
    ctx.getStats(reset=True)
    n = 20
    for _ in range(n):
        if copying:
            with torch.cuda.stream(side):
                for _ in range(2):
                    if direction == "d2h":
                        host.copy_(dev, non_blocking=True)
                    else:
                        dev.copy_(host, non_blocking=True)
        ctx.generateBatchDevice(ms)
        torch.cuda.synchronize()
    st = ctx.getStats(reset=True)
    print(label, {k: round(v / n * 1e3, 1) for k, v in st["ms"].items() if v > 0}, flush=True)

run("no copy      ", False)
run("with d2h copy", True, "d2h")
run("with h2d copy", True, "h2d")
run("no copy      ", False)

# the same without any timing event between the kernels: wall clock per synchronous batch
ctx.setProfiling(False)
def run_wall(label, lanes, streams, copying):
    ctx.setPipeline(lanes, streams)
    for _ in range(3):
        ctx.generateBatchDevice(ms)
    torch.cuda.synchronize()
    n = 20
    tot = 0.0
    for _ in range(n):
        if copying:
            with torch.cuda.stream(side):
                for _ in range(2):
                    host.copy_(dev, non_blocking=True)
        t0 = time.perf_counter()
        ctx.generateBatchDevice(ms)
        tot += time.perf_counter() - t0
        torch.cuda.synchronize()
    print(f"{label} lanes={lanes} streams={streams} copy={copying}: {tot / n * 1e6:.1f} us per batch", flush=True)

for lanes, streams in ((1, 1), (4, 1), (4, 2), (8, 2)):
    run_wall("wall", lanes, streams, False)
    run_wall("wall", lanes, streams, True)
