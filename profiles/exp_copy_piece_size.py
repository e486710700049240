#!/usr/bin/env python
"""The device-resident ring batch while a second stream copies device -> pinned host continuously, in pieces of
different sizes: does the kernel chain's slow-down under a D2H copy depend on the size of the copy commands?"""
import os, sys
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
src = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
dst = torch.empty(64 << 20, dtype=torch.uint8, pin_memory=True)
for _ in range(5):
    ctx.generateBatchDevice(ms)


def run(n):
    evs = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); ctx.generateBatchDevice(ms); b.record(stream); evs.append((a, b))
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in evs]))


print(f"alone {run(30):.3f} ms")
for piece in (64 << 20, 8 << 20, 1 << 20, 256 << 10, 64 << 10):
    total = 12 << 30      # ~240 ms at 50 GB/s
    n = min(total // piece, 60000)
    with torch.cuda.stream(side):
        for i in range(n):
            o = (i * piece) % (64 << 20)
            dst[o:o + piece].copy_(src[o:o + piece], non_blocking=True)
    t = run(30)
    done = side.query()
    torch.cuda.synchronize()
    print(f"D2H in pieces of {piece >> 10} KB ({n} copies, still running at the end: {not done}): batch {t:.3f} ms")

# the same copies alone: what piece size costs in throughput
for piece in (64 << 20, 8 << 20, 4 << 20, 2 << 20, 1 << 20, 512 << 10, 256 << 10, 64 << 10):
    n = min((2 << 30) // piece, 20000)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        a.record(side)
        for i in range(n):
            o = (i * piece) % (64 << 20)
            dst[o:o + piece].copy_(src[o:o + piece], non_blocking=True)
        b.record(side)
    torch.cuda.synchronize()
    print(f"D2H alone in pieces of {piece >> 10} KB: {n * piece / a.elapsed_time(b) / 1e6:.1f} GB/s")
