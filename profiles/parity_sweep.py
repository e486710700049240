#!/usr/bin/env python
"""Bit-exact parity of the CUDA path against the CPU checker over a whole world, outside the test
suite's time budget: BASELINE config 5's sweep (16 x 16 x 16 = 4096 LOD0 chunks of the default
terrain) plus the same footprint at LOD1 and LOD2 -- counts of every chunk, and vertices, index
topology and seam nodes of every chunk that has a surface.  Checker = oracle/ (test infrastructure).
    python profiles/parity_sweep.py  ->  one JSON line"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import leven_b200.compute as lc
from oracle import oracle as O
import bench as B

assert lc.Compute_Initialise(B.SEED, 0, 2) == 0
ctx = lc.Compute_MeshGenContext.create(B.V)
world = O.World(image=lc.Compute_GetNoiseImage(), default_material=0, voxels_per_chunk=B.V)
out = {"levels": []}
t0 = time.perf_counter()
for lod, size in enumerate((256, 512, 1024)):
    n = 16 >> lod
    ms = np.array([[cx * size, cy * size, cz * size, size] for cy in range(n) for cz in range(-n // 2, n // 2) for cx in range(-n // 2, n // 2)], np.int32)
    rc, res, view = ctx.generateBatchDevice(ms)
    assert rc == 0, lc.GetCLErrorString(rc)
    V = np.zeros(int(view.totalVertices) + 1, lc.MeshVertex); T = np.zeros(int(view.totalTriangles) + 1, lc.MeshTriangle)
    S = np.zeros(int(view.totalSeamNodes) + 1, lc.SeamNodeInfo)
    rc, res = ctx.generateBatch(ms, V, T, S)
    assert rc == 0
    counts, _ = world.batch_counts(ms)          # edges, nodes, triangles, seam nodes per chunk
    bad_counts = int(np.sum((counts[:, 0] != res["numEdges"]) | (counts[:, 2] != res["numTriangles"]) | (counts[:, 3] != res["numSeamNodes"])))
    surface = np.nonzero(counts[:, 0] > 0)[0]
    bad = 0
    for i in surface:
        r = res[i]
        ref = world.generate_chunk_mesh(list(ms[i][:3]), int(ms[i][3]))
        world.free_chunk_octree(list(ms[i][:3]), int(ms[i][3]))
        v = V[r["vertexOffset"]:r["vertexOffset"] + r["numVertices"]]
        t = T["indices_"][r["triangleOffset"]:r["triangleOffset"] + r["numTriangles"]]
        s = S[r["seamOffset"]:r["seamOffset"] + r["numSeamNodes"]]
        nv = ref["numNodes"] if ref["numTriangles"] > 0 else 0      # no quads -> the mesh buffer stays empty
        ok = (len(v) == nv and len(t) == ref["numTriangles"] and len(s) == ref["numSeamNodes"]
              and v.tobytes() == ref["vertices"][:nv].tobytes() and np.array_equal(t, ref["indices"]) and s.tobytes() == ref["seams"].tobytes())
        bad += 0 if ok else 1
    out["levels"].append({"lod": lod, "chunks": len(ms), "with_surface": int(len(surface)), "vertices": int(res["numVertices"].sum()),
                          "triangles": int(res["numTriangles"].sum()), "seam_nodes": int(res["numSeamNodes"].sum()),
                          "chunks_with_wrong_counts": bad_counts, "chunks_with_wrong_bytes": bad})
out["seconds"] = time.perf_counter() - t0
out["what"] = "vertices (48 B each), triangle indices in order, seam nodes (48 B each): byte-identical to oracle/lvn_oracle.c"
print(json.dumps(out))
