#!/usr/bin/env python
"""compute-sanitizer workload for the widened rows (seam meshes, simplifier, fused batch, collision
feed, batched update): a 3 x 2 x 3 block of LOD0 nodes around the surface through
lvn_clipmap_update_batch, the collision batch on two 512 nodes, and the stand-alone simplifier on
hand-made meshes (closed, open, two materials, under the size floor)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import leven_b200.compute as lc
import simplify_scenarios as SS
import bench as B
assert lc.Compute_Initialise(B.SEED, 0, 2) == 0
ctx = lc.Compute_MeshGenContext.create(B.V)
cover = [((cx * 256, (B.CY0 + dy) * 256, cz * 256), 256) for cx in (-1, 0, 1) for dy in (-1, 0) for cz in (-1, 0, 1)]
nodes = np.zeros(len(cover), lc.ClipmapNode)
for k, (mn, size) in enumerate(cover):
    nodes[k]["min"] = mn; nodes[k]["size"] = size
V = np.zeros(200000, lc.MeshVertex); T = np.zeros(400000, lc.MeshTriangle); Sn = np.zeros(100000, lc.SeamNodeInfo)
rc, cres, upd, sres, tot = lc.ClipmapUpdateBatch(ctx, nodes, 0, Sn, 0, V, T)
assert rc == 0 and tot.numSeamUpdates > 0 and tot.nodeTriangles > 0
# pass 2 on its own, in two shares, the second reading the seam nodes from device memory (the form
# the multi-GPU update uses after the all-gather)
import torch
active = np.array([k for k, r in enumerate(cres) if r["numTriangles"] > 0 or r["numSeamNodes"] > 0], np.int32)
dev = torch.from_numpy(Sn[:tot.seamNodesUsed].view(np.uint8).reshape(-1).copy()).cuda()
for shard, arena in ((0, Sn), (1, int(dev.data_ptr()))):
    V2 = np.zeros(100000, lc.MeshVertex); T2s = np.zeros(100000, lc.MeshTriangle)
    rc, u2, s2, n_all = lc.ClipmapSeamUpdateBatch(64, nodes, active, active, arena, tot.seamNodesUsed, V2, T2s, shard, 2)
    assert rc == 0 and n_all == tot.numSeamUpdates and len(u2) > 0
y1 = (B.CY0 * 256 // 512) * 512
P = np.zeros((100000, 4), np.float32); T2 = np.zeros((200000, 3), np.int32)
rc, res, simp = ctx.generateCollisionBatch([[0, y1, 0, 512], [-512, y1, 0, 512]], P, T2, Sn)
assert rc == 0 and res["numTriangles"].sum() > 0
meshes, opts = [], []
for (v, t), off, o in ((SS.torus(), [0, 0, 0], dict(edgeFraction=0.5, minAngleCosine=0.2)), (SS.grid_plane(40, 4.0, 0.6), [78, 0, 78], dict(edgeFraction=1.0)),
                       (SS.torus(two_materials=True), [0, 0, 0], dict(edgeFraction=1.0, maxIterations=20)), (SS.grid_plane(7), [12, 0, 12], {}),
                       (SS.fan(), [0, 0, 0], dict(edgeFraction=1.0))):
    meshes.append((v, t, off))
    opts.append(lc.SimplifyOptions.make(maxError=50.0, maxEdgeSize=20.0, **o))
rc, out, sr = lc.ngMeshSimplifierBatch(meshes, opts)
assert rc == 0 and sr["numTriangles"][0] < len(meshes[0][1])
ctx.destroy()
print("sanitize widened workload ok")
