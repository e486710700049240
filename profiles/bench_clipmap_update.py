#!/usr/bin/env python
"""One Clipmap::update through lvn_clipmap_update_batch: construct a block of
LOD0 nodes around the surface (chunk pass + simplifier, fused) and regenerate every seam the
update invalidates, host arenas in and out.  Beside it: the same work as separate batch calls
with host round trips (generateBatch -> ngMeshSimplifierBatch -> seam batch), and, where oracle/_ref
is built, the reference's own simplifier and octree.cpp on one host core for the two CPU steps the
application runs today (its chunk kernels are timed by bench.py)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import leven_b200.compute as lc
import seam_scenarios as S
import simplify_scenarios as SS
import bench as B
from oracle import ref as R

assert lc.Compute_Initialise(B.SEED, 0, 2) == 0
ctx = lc.Compute_MeshGenContext.create(64)
cy = B.CY0
cover = [((cx * 256, (cy + dy) * 256, cz * 256), 256) for cx in range(-8, 8) for dy in (-1, 0, 1) for cz in range(-8, 8)]
keep = []
pin = lambda n, dt: (keep.append(lc.PinnedArray(n, dt)), keep[-1].array)[1]
V, T, Sn = pin(3000000, lc.MeshVertex), pin(6000000, lc.MeshTriangle), pin(1000000, lc.SeamNodeInfo)


def fresh_nodes():
    nodes = np.zeros(len(cover), lc.ClipmapNode)
    for k, (mn, size) in enumerate(cover):
        nodes[k]["min"] = mn; nodes[k]["size"] = size
    return nodes


times = []
for it in range(8):
    nodes = fresh_nodes()
    t0 = time.perf_counter()
    rc, cres, upd, sres, tot = lc.ClipmapUpdateBatch(ctx, nodes, 0, Sn, 0, V, T)
    times.append(time.perf_counter() - t0)
    assert rc == 0, (rc, lc.last_cuda_error())
upd_s = float(np.median(times[2:]))
out = {"nodes_constructed": len(cover), "nodes_active": int(tot.numConstructedActive), "seam_updates": int(tot.numSeamUpdates),
       "node_vertices": int(tot.nodeVertices), "node_triangles": int(tot.nodeTriangles), "seam_vertices": int(tot.seamVertices),
       "seam_triangles": int(tot.seamTriangles), "seam_nodes": int(tot.seamNodesUsed),
       "update_ms": upd_s * 1e3, "update_nodes_per_s": len(cover) / upd_s,
       "update_what": "one lvn_clipmap_update_batch call: node list in; simplified node meshes, seam nodes and seam meshes out (pinned arenas)"}

# the same work as separate calls with host round trips between them
ms = np.array([list(mn) + [size] for mn, size in cover], np.int32)
V2, T2, S2 = pin(3000000, lc.MeshVertex), pin(6000000, lc.MeshTriangle), pin(1000000, lc.SeamNodeInfo)
t_gen, t_simp, t_seam = [], [], []
for it in range(5):
    t0 = time.perf_counter()
    rc, res = ctx.generateBatch(ms, V2, T2, S2)
    t1 = time.perf_counter()
    assert rc == 0
    meshes = [(V2[r["vertexOffset"]:r["vertexOffset"] + r["numVertices"]], T2[r["triangleOffset"]:r["triangleOffset"] + r["numTriangles"]],
               [c[0] + 128.0, c[1] + 128.0, c[2] + 128.0]) for c, r in zip(ms, res) if r["numTriangles"]]
    jobs, Vp, Tp = lc.PackSimplifyMeshes(meshes)
    t2 = time.perf_counter()
    rc, sr = lc.ngMeshSimplifierPacked(jobs, lc.SimplifyOptions.for_clipmap_node(256), Vp, Tp)
    t3 = time.perf_counter()
    assert rc == 0
    seams = {a[0]: S2[r["seamOffset"]:r["seamOffset"] + r["numSeamNodes"]] for a, r in zip(cover, res)}
    sj = S.build_jobs([c for c, r in zip(cover, res) if r["numTriangles"] or r["numSeamNodes"]], lambda mn, size: seams[tuple(mn)])
    pj, pn, pa = lc.PackSeamJobs(sj)
    t4 = time.perf_counter()
    rc, _, _, sres2 = lc.GenerateClipmapSeamMeshesPacked(64, pj, pn, pa)
    t5 = time.perf_counter()
    assert rc == 0
    t_gen.append(t1 - t0); t_simp.append(t3 - t2); t_seam.append(t5 - t4)
assert int(sres2["numTriangles"].sum()) == out["seam_triangles"] and int(sr["numTriangles"].sum()) == out["node_triangles"]
out.update(separate_generate_ms=float(np.median(t_gen)) * 1e3, separate_simplify_ms=float(np.median(t_simp)) * 1e3,
           separate_seams_ms=float(np.median(t_seam)) * 1e3,
           separate_what="lvn_meshgen_generate_batch, lvn_mesh_simplify_batch, lvn_seam_mesh_generate_batch one after the other "
                         "(the python list building between them is not timed)")
if R.simplify_available() and R.octree_available():
    ropt = SS.clipmap_options(256)
    t0 = time.perf_counter()
    for v, t, off in meshes:
        R.simplify_mesh(SS.as_vertices(v), t["indices_"], off, ropt)
    ref_simp = time.perf_counter() - t0
    import ctypes as C
    sel = [R.select_seam_nodes(h, s, nb) for h, s, nb in sj]
    t0 = time.perf_counter()
    for (h, s, nb), (m, p, nr, mt) in zip(sj, sel):
        k = len(m)
        if k == 0:
            continue
        verts = np.zeros(k, R.VERTEX_DTYPE); tris = np.zeros((12 * k, 3), np.int32); nv = C.c_int(0)
        R.octree_lib().ref_seam_octree_mesh(k, R._p(m), R._p(p), R._p(nr), R._p(mt), R._p(np.array(h, np.int32)), 2 * s,
                                            R._p(np.ones(3, np.float32)), R._p(verts), k, C.byref(nv), R._p(tris), len(tris))
    ref_seam = time.perf_counter() - t0
    out.update(reference_simplify_ms=ref_simp * 1e3, reference_seam_octree_ms=ref_seam * 1e3,
               reference_what="ng_mesh_simplify.cpp and octree.cpp (oracle/_ref) on one host core, per node as the application calls them; "
                              "the chunk kernels themselves are bench.py's reference arm")
print(json.dumps(out))
