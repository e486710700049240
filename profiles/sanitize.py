#!/usr/bin/env python
"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck): one surface
chunk through generateChunkMesh, a 27-chunk batch through both batch entry points, one CSG edit
with re-mesh, one stress-field chunk."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import leven_b200.compute as lc
import bench as B
assert lc.Compute_Initialise(B.SEED, 0, 2) == 0
ctx = lc.Compute_MeshGenContext.create(B.V)
mn = [0, B.CY0 * B.SIZE, 0]
mesh, seams = lc.MeshBuffer(), []
assert ctx.generateChunkMesh(mn, B.SIZE, mesh, seams) == 0
ms = np.array([[cx * 256, (B.CY0 + dy) * 256, cz * 256, 256] for dy in (-1, 0, 1) for cz in (-1, 0, 1) for cx in (-1, 0, 1)], np.int32)
ctx.setPipeline(3, 2)
rc, res, view = ctx.generateBatchDevice(ms)
assert rc == 0
V = np.zeros(int(view.totalVertices), lc.MeshVertex); T = np.zeros(int(view.totalTriangles), lc.MeshTriangle)
S = np.zeros(int(view.totalSeamNodes), lc.SeamNodeInfo)
rc, res = ctx.generateBatch(ms, V, T, S)
assert rc == 0
op = lc.CSGOperationInfo.make(1, 1, 201, [40.5, B.CY0 * 64 + 30.5, 40.5], [12.0, 12.0, 12.0], 0.0)
assert ctx.applyCSGOperationsBatch([op], ms[:8]) == 0
rc, res = ctx.generateBatch(ms, V, T, S)
assert rc == 0 or rc == lc.LVN_ERR_CAPACITY
ctx.destroy()
lc.Compute_SetDensityFunction(1, 0.735)
ctx = lc.Compute_MeshGenContext.create(16)
rc, res, view = ctx.generateBatchDevice(np.array([[0, 0, 0, 64]], np.int32))
assert rc == 0 and view.totalVertices > 0
ctx.destroy()
print("sanitize workload ok")
