#!/usr/bin/env python
"""Three device-resident passes of config 4 (64 chunks of the dense 3-D stress field): ncu target for k_hermite / k_field_density."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import leven_b200.compute as lc
import leven_b200.workloads as W
assert lc.Compute_Initialise(W.SEED, 0, 2) == 0
assert lc.Compute_SetDensityFunction(1, W.STRESS_THRESHOLD) == 0
ctx = lc.Compute_MeshGenContext.create(W.V)
ms = W.stress_chunks()
for _ in range(3):
    rc, res, view = ctx.generateBatchDevice(ms)
    assert rc == 0
