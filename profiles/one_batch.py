#!/usr/bin/env python
"""Five device-resident passes of the bench workload (ncu target)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import leven_b200.compute as lc
import bench as B
assert lc.Compute_Initialise(B.SEED, 0, 2) == 0
ctx = lc.Compute_MeshGenContext.create(B.V)
ms = B.ring_chunks(0)
for _ in range(5):
    rc, res, view = ctx.generateBatchDevice(ms)
    assert rc == 0
