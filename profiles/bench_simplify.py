#!/usr/bin/env python
"""Mesh simplification: GPU batch (lvn_mesh_simplify_batch, host arrays in and out,
copies included) against the reference's own ng_mesh_simplify.cpp + qef_simd.h compiled for the host
(oracle/_ref), one mesh at a time as ConstructClipmapNodeData does, on the same inputs.
Workloads: the non-empty chunk meshes of BASELINE config 2 (512-chunk LOD0 ring) and of config 5
(the 4096-chunk world sweep), produced by the CUDA path itself."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import leven_b200.compute as lc
import simplify_scenarios as S
import bench as B
from oracle import ref as R

assert lc.Compute_Initialise(B.SEED, 0, 2) == 0
ctx = lc.Compute_MeshGenContext.create(64)


_keep = []


def pinned(n, dtype):
    a = lc.PinnedArray(n, dtype)
    _keep.append(a)
    return a.array


TOTALS = {}


def chunk_meshes(ms):
    rc, res, view = ctx.generateBatchDevice(ms)
    TOTALS.update(v=int(view.totalVertices) + 1, t=int(view.totalTriangles) + 1, s=int(view.totalSeamNodes) + 1)
    V = np.zeros(int(view.totalVertices) + 1, lc.MeshVertex); T = np.zeros(int(view.totalTriangles) + 1, lc.MeshTriangle)
    Sn = np.zeros(int(view.totalSeamNodes) + 1, lc.SeamNodeInfo)
    rc, res = ctx.generateBatch(ms, V, T, Sn)
    assert rc == 0
    out = []
    for c, r in zip(ms, res):
        if r["numTriangles"]:
            out.append((V[r["vertexOffset"]:r["vertexOffset"] + r["numVertices"]], T[r["triangleOffset"]:r["triangleOffset"] + r["numTriangles"]],
                        [c[0] + 128.0, c[1] + 128.0, c[2] + 128.0]))
    return out


def run(name, ms, check):
    meshes = chunk_meshes(ms)
    jobs, V0, T0 = lc.PackSimplifyMeshes(meshes)
    opt = lc.SimplifyOptions.for_clipmap_node(256)
    times = []
    V, T = pinned(len(V0), lc.MeshVertex), pinned(len(T0), lc.MeshTriangle)     # the caller's arrays: pinned, as MeshBuffer pools would be
    for it in range(8):
        V[:] = V0; T[:] = T0
        t0 = time.perf_counter()
        rc, res = lc.ngMeshSimplifierPacked(jobs, opt, V, T)
        times.append(time.perf_counter() - t0)
        assert rc == 0, lc.lib().lvn_mesh_simplify_last_error()
    gpu_s = float(np.median(times[2:]))
    out = {"workload": name, "chunks": len(ms), "meshes": len(meshes), "vertices_in": int(jobs["numVertices"].sum()),
           "triangles_in": int(jobs["numTriangles"].sum()), "vertices_out": int(res["numVertices"].sum()),
           "triangles_out": int(res["numTriangles"].sum()), "iterations_mean": float(res["iterations"].mean()),
           "gpu_ms_per_batch": gpu_s * 1e3, "gpu_meshes_per_s": len(meshes) / gpu_s,
           "gpu_what": "one lvn_mesh_simplify_batch call: H2D of the meshes (pinned), one kernel (a block per mesh), D2H"}
    # the fused route: generate + simplify in HBM, simplified meshes out (lvn_meshgen_generate_simplified_batch)
    Vf, Tf, Sf = pinned(TOTALS['v'], lc.MeshVertex), pinned(TOTALS['t'], lc.MeshTriangle), pinned(TOTALS['s'], lc.SeamNodeInfo)
    tf, tg = [], []
    for it in range(8):
        t0 = time.perf_counter()
        rc, fres, fsimp = ctx.generateSimplifiedBatch(ms, Vf, Tf, Sf)
        tf.append(time.perf_counter() - t0)
        assert rc == 0, lc.last_cuda_error()
        t0 = time.perf_counter()
        rc, gres = ctx.generateBatch(ms, Vf, Tf, Sf)
        tg.append(time.perf_counter() - t0)
        assert rc == 0
    out.update(fused_ms_per_batch=float(np.median(tf[2:])) * 1e3, fused_chunks_per_s=len(ms) / float(np.median(tf[2:])),
               generate_only_ms_per_batch=float(np.median(tg[2:])) * 1e3,
               fused_what="lvn_meshgen_generate_simplified_batch: chunk list in, simplified meshes + seam nodes out (pinned arenas)",
               fused_vertices_out=int(fres["numVertices"].sum()), fused_triangles_out=int(fres["numTriangles"].sum()))
    assert out["fused_vertices_out"] == out["vertices_out"] and out["fused_triangles_out"] == out["triangles_out"]
    if R.simplify_available():
        ropt = S.clipmap_options(256)
        ref_s, bad = 0.0, 0
        for k, ((v, t, off), j, r) in enumerate(zip(meshes, jobs, res)):
            vv = S.as_vertices(v)
            a = time.perf_counter()
            rv, rt = R.simplify_mesh(vv, t["indices_"], off, ropt)
            ref_s += time.perf_counter() - a
            if check:
                gv = S.as_vertices(V[j["vertexOffset"]:j["vertexOffset"] + r["numVertices"]])
                gt = T["indices_"][j["triangleOffset"]:j["triangleOffset"] + r["numTriangles"]]
                bad += not (gv.tobytes() == rv.tobytes() and np.array_equal(gt, rt))
        out.update(reference_ms=ref_s * 1e3, reference_meshes_per_s=len(meshes) / ref_s, reference_what="ngMeshSimplifier per mesh, 1 host thread "
                   "(python call overhead and a MeshBuffer copy included)", mismatching_meshes=bad if check else None)
    print(json.dumps(out), flush=True)


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("ring", "all"):
    run("config 2: 512-chunk LOD0 ring", B.ring_chunks(0), True)
if which in ("sweep", "all"):
    run("config 5: 4096-chunk world sweep", np.array([[(cx - 8) * B.SIZE, cy * B.SIZE, (cz - 8) * B.SIZE, B.SIZE] for cy in range(16) for cz in range(16) for cx in range(16)], np.int32), True)
