#!/usr/bin/env python
"""Seam meshes: GPU batch (lvn_seam_mesh_generate_batch, host arrays in and out,
copies included) against the reference's own octree.cpp compiled for the host (oracle/_ref), one
seam at a time as GenerateClipmapSeamMesh does, on the same inputs.  Workload: every node of a
16 x 3 x 16 LOD0 block around the surface (768 seams) and the mixed-LOD covers of the tests."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import leven_b200.compute as lc
import seam_scenarios as S
import bench as B
from oracle import oracle as O, ref as R

assert lc.Compute_Initialise(B.SEED, 0, 2) == 0
ctx = lc.Compute_MeshGenContext.create(64)
cy = B.CY0
active = [((cx * 256, (cy + dy) * 256, cz * 256), 256) for cx in range(-8, 8) for dy in (-1, 0, 1) for cz in range(-8, 8)]
ms = np.array([list(a[0]) + [256] for a in active], np.int32)
# the seam nodes come from the CUDA path itself (one batch), as they would in the application
rc, res, view = ctx.generateBatchDevice(ms)
V = np.zeros(int(view.totalVertices) + 1, lc.MeshVertex); T = np.zeros(int(view.totalTriangles) + 1, lc.MeshTriangle)
Sn = np.zeros(int(view.totalSeamNodes) + 1, lc.SeamNodeInfo)
rc, res = ctx.generateBatch(ms, V, T, Sn)
assert rc == 0
seams = {a[0]: Sn[r["seamOffset"]:r["seamOffset"] + r["numSeamNodes"]] for a, r in zip(active, res)}
jobs = S.build_jobs(active, lambda mn, size: seams[tuple(mn)])
pj, pn, pa = lc.PackSeamJobs(jobs)
cand = int(pn['numNodes'].sum())
Vb = np.zeros(cand, lc.MeshVertex); Tb = np.zeros(8 * cand, lc.MeshTriangle)
for _ in range(3):
    rc, _, _, sres = lc.GenerateClipmapSeamMeshesPacked(64, pj, pn, pa, Vb, Tb)
assert rc == 0
n = 20
t0 = time.perf_counter()
for _ in range(n):
    rc, _, _, sres = lc.GenerateClipmapSeamMeshesPacked(64, pj, pn, pa, Vb, Tb)
gpu_s = (time.perf_counter() - t0) / n
out = {"seams": len(jobs), "candidate_nodes": cand, "uploaded_nodes": int(len(pa)), "selected_nodes": int(sres["numSelectedNodes"].sum()),
       "selected_nodes_max": int(sres["numSelectedNodes"].max()), "candidate_nodes_max": int(max(pn["numNodes"][j["firstNeighbour"]:j["firstNeighbour"] + j["numNeighbours"]].sum() for j in pj)),
       "vertices": int(sres["numVertices"].sum()), "triangles": int(sres["numTriangles"].sum()),
       "gpu_ms_per_batch": gpu_s * 1e3, "gpu_seams_per_s": len(jobs) / gpu_s,
       "gpu_what": "one lvn_seam_mesh_generate_batch call: H2D of the seam nodes (pageable), one kernel, D2H of the meshes"}
if R.octree_available():
    conv = lambda a: np.ascontiguousarray(a).view(R.SEAM_DTYPE) if a.dtype != R.SEAM_DTYPE else a
    t0 = time.perf_counter()
    tv = tt = 0
    for host, size, nbs in jobs:
        rv, rt = R.seam_mesh(host, size, nbs)
        tv += len(rv); tt += len(rt)
    ref_s = time.perf_counter() - t0
    # the selection loop of oracle/ref.py is Python; time the compiled octree part alone as well
    sel = [R.select_seam_nodes(h, s, nb) for h, s, nb in jobs]
    import ctypes as C
    t0 = time.perf_counter()
    for (h, s, nb), (m, p, nr, mt) in zip(jobs, sel):
        k = len(m)
        if k == 0:
            continue
        verts = np.zeros(k, R.VERTEX_DTYPE); tris = np.zeros((12 * k, 3), np.int32); nv = C.c_int(0)
        R.octree_lib().ref_seam_octree_mesh(k, R._p(m), R._p(p), R._p(nr), R._p(mt), R._p(np.array(h, np.int32)), 2 * s,
                                            R._p(np.ones(3, np.float32)), R._p(verts), k, C.byref(nv), R._p(tris), len(tris))
    oct_s = time.perf_counter() - t0
    assert tv == out["vertices"] and tt == out["triangles"], (tv, tt, out)
    out.update(reference_octree_cpp_ms=oct_s * 1e3, reference_octree_seams_per_s=len(jobs) / oct_s,
               reference_with_python_selection_ms=ref_s * 1e3)
print(json.dumps(out))
