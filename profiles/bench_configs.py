#!/usr/bin/env python
"""Throughput of the other BASELINE.json configurations (bench.py measures configs[1]).

    python profiles/bench_configs.py [--out profiles/r01_configs.json]

configs[0] single chunk through generateChunkMesh (cold octree cache every call)
configs[2] CSG edit script: 32 ops, one per step; apply to the overlapping chunks, re-mesh them
configs[3] dense stress field (ridged 3-D fBm, ~30 % active voxels), 64 chunks
configs[4] full sweep of 4096 chunks (16 x 16 x 16), one batch
Wall clock around synchronous calls (every call ends in a stream synchronisation)."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import leven_b200.compute as lc
import bench as B


def timed(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    assert lc.Compute_Initialise(B.SEED, 0, 2) == 0
    out = {}

    # ---- configs[0]: one surface chunk, the reference's own call ----
    ctx = lc.Compute_MeshGenContext.create(B.V)
    mn = [0, B.CY0 * B.SIZE, 0]
    mesh, seams = lc.MeshBuffer(), []
    def one():
        assert ctx.generateChunkMesh(mn, B.SIZE, mesh, seams) == 0
        ctx.freeChunkOctree(mn, B.SIZE)
    dt = timed(one, 50)
    out["config0_single_chunk"] = {"ms_per_chunk": dt * 1e3, "chunks_per_s": 1.0 / dt,
                                   "vertices": int(mesh.numVertices), "triangles": int(mesh.numTriangles)}
    print(json.dumps({"config0": out["config0_single_chunk"]}), flush=True)

    # ---- configs[4]: 4096-chunk sweep, one batch ----
    sweep = np.array([[(cx - 8) * B.SIZE, cy * B.SIZE, (cz - 8) * B.SIZE, B.SIZE]
                      for cy in range(16) for cz in range(16) for cx in range(16)], np.int32)
    rc, res, view = ctx.generateBatchDevice(sweep)
    assert rc == 0
    dt = timed(lambda: ctx.generateBatchDevice(sweep), 10)
    import torch
    def pinned(n, dtype):
        t = torch.empty(max(n, 1) * dtype.itemsize, dtype=torch.uint8, pin_memory=True)
        return t, t.numpy().view(dtype)
    k1, hv = pinned(int(view.totalVertices) + 1024, lc.MeshVertex)
    k2, ht = pinned(int(view.totalTriangles) + 1024, lc.MeshTriangle)
    k3, hs = pinned(int(view.totalSeamNodes) + 1024, lc.SeamNodeInfo)
    def e2e():
        rc, _ = ctx.generateBatch(sweep, hv, ht, hs)
        assert rc == 0
    dte = timed(e2e, 10)
    out["config4_sweep_4096"] = {"chunks": len(sweep), "non_empty": int((res["numEdges"] > 0).sum()),
                                 "vertices": int(view.totalVertices), "triangles": int(view.totalTriangles),
                                 "device_ms": dt * 1e3, "device_chunks_per_s": len(sweep) / dt,
                                 "e2e_ms": dte * 1e3, "e2e_chunks_per_s": len(sweep) / dte}
    print(json.dumps({"config4": out["config4_sweep_4096"]}), flush=True)

    # ---- configs[2]: CSG edit script on the ring's fields ----
    ring = B.ring_chunks(0)
    rng = np.random.RandomState(12345)
    sy = B.CY0 * 64
    ops = []
    for step in range(32):
        shape = step % 2
        add = (step // 2) % 2 == 0
        origin = [float(rng.randint(-128, 128)) + 0.5, sy + float(rng.randint(-40, 60)) + 0.5, float(rng.randint(-128, 128)) + 0.5]
        half = float(rng.randint(1, 32))
        dims = [half, half, half] if shape == 1 else [float(rng.randint(1, 32)) for _ in range(3)]
        ops.append(lc.CSGOperationInfo.make(0 if add else 1, shape, int(rng.randint(1, 4)) if add else 201, origin, dims, 0.0))
    V = np.zeros(2000000, lc.MeshVertex); T = np.zeros(4000000, lc.MeshTriangle); S = np.zeros(400000, lc.SeamNodeInfo)
    t_apply = t_mesh = 0.0
    touched_total = 0
    for op in ops:
        lo, hi = lc.CalcCSGOperationBounds(op)
        touched = np.array([c for c in ring if not (c[0] + 256 < lo[0] or c[1] + 256 < lo[1] or c[2] + 256 < lo[2] or
                                                    c[0] > hi[0] or c[1] > hi[1] or c[2] > hi[2])], np.int32)
        if not len(touched):
            continue
        t0 = time.perf_counter()
        assert ctx.applyCSGOperationsBatch([op], touched) == 0
        t1 = time.perf_counter()
        rc, r = ctx.generateBatch(touched, V, T, S)
        assert rc == 0, lc.GetCLErrorString(rc)
        t2 = time.perf_counter()
        assert lc.Compute_StoreCSGOperation(op, lo, hi) == 0
        t_apply += t1 - t0; t_mesh += t2 - t1; touched_total += len(touched)
    out["config2_csg_script"] = {"ops": len(ops), "chunk_edits": touched_total, "apply_ms_total": t_apply * 1e3,
                                 "remesh_ms_total": t_mesh * 1e3,
                                 "edited_chunks_per_s": touched_total / (t_apply + t_mesh),
                                 "ops_per_s": len(ops) / (t_apply + t_mesh)}
    print(json.dumps({"config2": out["config2_csg_script"]}), flush=True)
    lc.Compute_ClearCSGOperations()
    ctx.destroy()

    # ---- configs[3]: dense stress field ----
    lc.Compute_SetDensityFunction(1, 0.735)
    ctx = lc.Compute_MeshGenContext.create(B.V)
    stress = np.array([[cx * B.SIZE, cy * B.SIZE, cz * B.SIZE, B.SIZE] for cy in range(4) for cz in range(4) for cx in range(4)], np.int32)
    rc, res, view = ctx.generateBatchDevice(stress)
    assert rc == 0, lc.GetCLErrorString(rc)
    dt = timed(lambda: ctx.generateBatchDevice(stress), 5, warm=1)
    ctx.setProfiling(True); ctx.getStats(reset=True)
    ctx.generateBatchDevice(stress)
    st = ctx.getStats(reset=True); ctx.setProfiling(False)
    out["config3_stress_64"] = {"chunks": len(stress), "vertices": int(view.totalVertices), "edges": int(view.totalEdges),
                                "active_fraction": float(view.totalVertices) / (len(stress) * 64 ** 3),
                                "device_ms": dt * 1e3, "chunks_per_s": len(stress) / dt,
                                "stage_ms": {k: v for k, v in st["ms"].items() if v > 0}}
    print(json.dumps({"config3": out["config3_stress_64"]}), flush=True)
    lc.Compute_SetDensityFunction(0, 0.5)
    ctx.destroy()
    if args.out:
        json.dump(out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
