"""BASELINE.json configs 3, 4 and 5 at FULL size against the oracle.

  config 5  the 4096-chunk LOD0 sweep + the same footprint at LOD1 (512) and LOD2 (64): counts of
            every chunk; vertices, index topology in order and seam nodes of every chunk with a
            surface, byte for byte (FNV-1a digests taken on both sides, OpenMP over chunks on the CPU)
  config 3  the 32-op CSG script on the 512-chunk ring, one op per step: after every op every touched
            chunk -- field, edge set with (normal, t), nodes, QEF vertices -- and the re-meshed batch
  config 4  all 64 chunks of the dense stress field
  + the batch entry points after Compute_StoreCSGOperation (stored ops must replay there too)

Bar: bit-exact (the arithmetic spec fixes every operation); the north-star tolerance
1e-4 relative is what a relaxation of the spec would have to meet.
"""
import numpy as np
import pytest

import leven_b200.workloads as W
from test_parity_gpu import compare_csg_field

pytestmark = pytest.mark.gpu


@pytest.fixture()
def ctx(lc):
    c = lc.Compute_MeshGenContext.create(64)
    assert c.privateCtx_
    yield c
    c.destroy()


@pytest.fixture()
def gpu_world(lc, oracle_mod):
    oracle_mod.set_num_threads(__import__("os").cpu_count() or 1)
    w = oracle_mod.World(image=lc.Compute_GetNoiseImage(), default_material=0, voxels_per_chunk=64)
    yield w
    w.close()


def batch_host(lc, ctx, ms):
    rc, res, view = ctx.generateBatchDevice(ms)
    assert rc == 0, lc.GetCLErrorString(rc)
    V = np.zeros(int(view.totalVertices) + 1, lc.MeshVertex)
    T = np.zeros(int(view.totalTriangles) + 1, lc.MeshTriangle)
    S = np.zeros(int(view.totalSeamNodes) + 1, lc.SeamNodeInfo)
    rc, res = ctx.generateBatch(ms, V, T, S)
    assert rc == 0, lc.GetCLErrorString(rc)
    return res, V, T, S


def gpu_digests(O, res, V, T, S):
    out = np.zeros((len(res), 3), np.uint64)
    for i, r in enumerate(res):
        out[i, 0] = O.fnv1a64(V[r["vertexOffset"]:r["vertexOffset"] + r["numVertices"]])
        out[i, 1] = O.fnv1a64(T[r["triangleOffset"]:r["triangleOffset"] + r["numTriangles"]])
        out[i, 2] = O.fnv1a64(S[r["seamOffset"]:r["seamOffset"] + r["numSeamNodes"]])
    return out


def assert_batch_equals_oracle(lc, O, ctx, world, ms, what):
    res, V, T, S = batch_host(lc, ctx, ms)
    counts, dig = world.batch_digests(ms)          # E, N, T, S + digests per chunk
    assert np.array_equal(res["numEdges"], counts[:, 0]), f"{what}: edge counts"
    assert np.array_equal(res["numTriangles"], counts[:, 2]), f"{what}: triangle counts"
    assert np.array_equal(res["numSeamNodes"], counts[:, 3]), f"{what}: seam-node counts"
    # no quads -> the mesh buffer stays empty (compute_octree.cpp:227-232)
    nv = np.where(counts[:, 2] > 0, counts[:, 1], 0)
    assert np.array_equal(res["numVertices"], nv), f"{what}: vertex counts"
    got = gpu_digests(O, res, V, T, S)
    has_mesh = counts[:, 2] > 0
    assert np.array_equal(got[has_mesh, 0], dig[has_mesh, 0]), f"{what}: vertex bytes"
    assert np.array_equal(got[:, 1], dig[:, 1]), f"{what}: triangle indices"
    assert np.array_equal(got[:, 2], dig[:, 2]), f"{what}: seam nodes"
    return counts


def test_config5_full_sweep_all_lods(lc, oracle_mod, ctx, gpu_world):
    """4096 LOD0 chunks (BASELINE configs[4]) + the same footprint at LOD1 and LOD2"""
    for lod, size in enumerate((256, 512, 1024)):
        n = 16 >> lod
        ms = np.array([[cx * size, cy * size, cz * size, size]
                       for cy in range(n) for cz in range(-n // 2, n // 2) for cx in range(-n // 2, n // 2)], np.int32)
        if lod == 0:
            assert np.array_equal(ms, W.sweep_chunks())
        counts = assert_batch_equals_oracle(lc, oracle_mod, ctx, gpu_world, ms, f"sweep LOD{lod}")
        if lod == 0:
            surface = int((counts[:, 0] > 0).sum())
            assert 300 < surface < 600, surface      # "about 10 % of the chunks are non-empty"


def test_config4_stress_all_64_chunks(lc, oracle_mod):
    """every chunk of the dense field, not only [0, 0, 0]; ~30 % active voxels each"""
    try:
        lc.Compute_SetDensityFunction(1, W.STRESS_THRESHOLD)
        ctx = lc.Compute_MeshGenContext.create(64)
        oracle_mod.set_num_threads(__import__("os").cpu_count() or 1)
        world = oracle_mod.World(image=lc.Compute_GetNoiseImage(), default_material=0, voxels_per_chunk=64)
        world.set_density(1, W.STRESS_THRESHOLD)
        ms = W.stress_chunks()
        counts = assert_batch_equals_oracle(lc, oracle_mod, ctx, world, ms, "stress")
        active = counts[:, 1].sum() / (len(ms) * 64.0 ** 3)
        assert 0.28 <= active <= 0.32, active
        assert counts[:, 1].max() > 14336            # beyond MeshBuffer's fixed capacity: arenas, not MeshBuffer
        ctx.destroy(); world.close()
    finally:
        lc.Compute_SetDensityFunction(0, 0.5)


def test_stress_hermite_with_and_without_lattice_densities(lc, monkeypatch):
    """the generic Hermite kernel takes steps 0 and 16 of an edge's search from the density values the field
    kernel of the same batch left behind; without them (batches beyond 8 GB of such values) it evaluates all 17
    steps itself: the same meshes, byte for byte"""
    try:
        lc.Compute_SetDensityFunction(1, W.STRESS_THRESHOLD)
        ms = W.stress_chunks()[:8]
        out = []
        for flag in ("0", "1"):
            monkeypatch.setenv("LVN_TEST_NO_LATTICE_DENSITY", flag)      # read when the context is created
            ctx = lc.Compute_MeshGenContext.create(64)
            rc, res, view = ctx.generateBatchDevice(ms)
            assert rc == 0
            V = np.zeros(int(view.totalVertices) + 16, lc.MeshVertex); T = np.zeros(int(view.totalTriangles) + 16, lc.MeshTriangle)
            S = np.zeros(int(view.totalSeamNodes) + 16, lc.SeamNodeInfo)
            rc, r = ctx.generateBatch(ms, V, T, S)
            assert rc == 0
            out.append((r.copy(), V.copy(), T.copy(), S.copy()))
            ctx.destroy()
        (ra, Va, Ta, Sa), (rb, Vb, Tb, Sb) = out
        assert ra["numVertices"].sum() > 100000
        for k in ("numEdges", "numVertices", "numTriangles", "numSeamNodes"):
            assert (ra[k] == rb[k]).all(), k
        for i in range(len(ms)):
            for arr_a, arr_b, off, cnt in ((Va, Vb, "vertexOffset", "numVertices"), (Ta, Tb, "triangleOffset", "numTriangles"),
                                           (Sa, Sb, "seamOffset", "numSeamNodes")):
                a = arr_a[ra[off][i]:ra[off][i] + ra[cnt][i]]
                b = arr_b[rb[off][i]:rb[off][i] + rb[cnt][i]]
                assert a.tobytes() == b.tobytes(), (i, cnt)
    finally:
        lc.Compute_SetDensityFunction(0, 0.5)


def test_config3_csg_script_on_ring(lc, oracle_mod, ctx, gpu_world):
    """the 32-op script, one op per step on the ring's fields; after each op the touched chunks are
    compared stage by stage and re-meshed through the batch call config 3 times"""
    world = gpu_world
    ring = W.ring_chunks()
    edits = 0
    try:
        for spec in W.csg_script():
            op, oop = lc.CSGOperationInfo.make(*spec), oracle_mod.make_csg_op(*spec)
            lo, hi = lc.CalcCSGOperationBounds(op)
            assert (lo, hi) == tuple(oracle_mod.csg_operation_bounds(oop))
            touched = W.touched_chunks(ring, lo, hi)
            if len(touched):
                assert ctx.applyCSGOperationsBatch([op], touched) == 0
                for c in touched:
                    world.apply_csg_operations([oop], [int(v) for v in c[:3]], int(c[3]))
                    world.free_chunk_octree([int(v) for v in c[:3]], int(c[3]))
            assert lc.Compute_StoreCSGOperation(op, lo, hi) == 0
            world.store_csg_operation(oop, lo, hi)
            if not len(touched):
                continue
            res, V, T, S = batch_host(lc, ctx, touched)
            for c, r in zip(touched, res):
                mn, size = [int(v) for v in c[:3]], int(c[3])
                ref = compare_csg_field(ctx, world, mn, size)        # field, edges, nodes, positions, normals
                edits += 1
                nv = ref["numNodes"] if ref["numTriangles"] > 0 else 0
                assert r["numVertices"] == nv and r["numTriangles"] == ref["numTriangles"] and r["numSeamNodes"] == ref["numSeamNodes"]
                ref = world.generate_chunk_mesh(mn, size)
                world.free_chunk_octree(mn, size)
                v = V[r["vertexOffset"]:r["vertexOffset"] + nv]
                assert v.tobytes() == ref["vertices"][:nv].tobytes(), f"CSG vertices {mn}"
                assert np.array_equal(T["indices_"][r["triangleOffset"]:r["triangleOffset"] + r["numTriangles"]], ref["indices"]), f"CSG indices {mn}"
                assert S[r["seamOffset"]:r["seamOffset"] + r["numSeamNodes"]].tobytes() == ref["seams"].tobytes(), f"CSG seams {mn}"
        assert edits >= 64, edits
    finally:
        lc.Compute_ClearCSGOperations()


def test_batch_entry_points_replay_stored_ops(lc, oracle_mod, surface_cy):
    """LoadDensityField's replay (compute_density_field.cpp:235-274) also runs in front of the batch
    entry points: a chunk that stored operations overlap is meshed from the edited field by
    generate_batch, generate_batch_device, generate_simplified_batch and lvn_clipmap_update_batch"""
    ctx = lc.Compute_MeshGenContext.create(64)
    world = oracle_mod.World(image=lc.Compute_GetNoiseImage(), default_material=0, voxels_per_chunk=64)
    try:
        yc = surface_cy * 64 + 30.5
        specs = [(1, 1, 201, [30.5, yc, 30.5], [9, 9, 9]), (0, 0, 2, [60.5, yc + 8, 40.5], [7, 3, 11]),
                 (0, 1, 3, [250.5, yc, 64.5], [12, 12, 12])]          # the last one straddles two chunks
        for s in specs:
            op, oop = lc.CSGOperationInfo.make(*s), oracle_mod.make_csg_op(*s)
            lo, hi = lc.CalcCSGOperationBounds(op)
            lc.Compute_StoreCSGOperation(op, lo, hi)
            world.store_csg_operation(oop, lo, hi)
        ms = np.array([[0, surface_cy * 256, 0, 256], [256, surface_cy * 256, 0, 256], [0, surface_cy * 256, 256, 256],
                       [-256, surface_cy * 256, 0, 256]], np.int32)
        refs = [world.generate_chunk_mesh([int(v) for v in c[:3]], 256) for c in ms]
        plain = oracle_mod.World(image=lc.Compute_GetNoiseImage(), default_material=0, voxels_per_chunk=64)
        unedited = [plain.generate_chunk_mesh([int(v) for v in c[:3]], 256)["numNodes"] for c in ms]
        plain.close()
        assert [r["numNodes"] for r in refs][:2] != unedited[:2]       # the edits are visible in the reference
        assert refs[3]["numNodes"] == unedited[3]                      # and this chunk is outside every op
        # host batch call
        res, V, T, S = batch_host(lc, ctx, ms)
        for r, ref in zip(res, refs):
            nv = ref["numNodes"] if ref["numTriangles"] > 0 else 0
            assert r["numVertices"] == nv and r["numTriangles"] == ref["numTriangles"]
            assert V[r["vertexOffset"]:r["vertexOffset"] + nv].tobytes() == ref["vertices"][:nv].tobytes()
            assert np.array_equal(T["indices_"][r["triangleOffset"]:r["triangleOffset"] + r["numTriangles"]], ref["indices"])
            assert S[r["seamOffset"]:r["seamOffset"] + r["numSeamNodes"]].tobytes() == ref["seams"].tobytes()
        # a fresh context: the device-resident call is the first one to see the stored ops
        ctx2 = lc.Compute_MeshGenContext.create(64)
        rc, res2, view = ctx2.generateBatchDevice(ms)
        assert rc == 0
        assert np.array_equal(res2["numTriangles"], [r["numTriangles"] for r in refs])
        assert np.array_equal(res2["numSeamNodes"], [r["numSeamNodes"] for r in refs])
        ctx2.destroy()
        # fused chunk + simplifier batch: the seam nodes come from the unsimplified octree of the EDITED field
        ctx3 = lc.Compute_MeshGenContext.create(64)
        V3 = np.zeros(200000, lc.MeshVertex); T3 = np.zeros(400000, lc.MeshTriangle); S3 = np.zeros(50000, lc.SeamNodeInfo)
        rc, res3, simp = ctx3.generateSimplifiedBatch(ms, V3, T3, S3)
        assert rc == 0
        for r, ref in zip(res3, refs):
            assert r["numSeamNodes"] == ref["numSeamNodes"]
            assert S3[r["seamOffset"]:r["seamOffset"] + r["numSeamNodes"]].tobytes() == ref["seams"].tobytes()
        ctx3.destroy()
        # the update driver (pass 1 = the fused call)
        ctx4 = lc.Compute_MeshGenContext.create(64)
        nodes = np.zeros(len(ms), lc.ClipmapNode)
        nodes["min"] = ms[:, :3]; nodes["size"] = ms[:, 3]
        S4 = np.zeros(50000, lc.SeamNodeInfo)
        out = lc.ClipmapUpdateBatch(ctx4, nodes, 0, S4, 0, V3, T3)
        assert out[0] == 0, lc.GetCLErrorString(out[0])
        for i, ref in enumerate(refs):
            n = nodes[i]                                                # updated in place
            assert n["numSeamNodes"] == ref["numSeamNodes"]
            assert S4[n["firstSeamNode"]:n["firstSeamNode"] + n["numSeamNodes"]].tobytes() == ref["seams"].tobytes()
        ctx4.destroy()
    finally:
        lc.Compute_ClearCSGOperations()
        ctx.destroy(); world.close()


def test_csg_deferred_hash_tables_are_validated(lc, oracle_mod, surface_cy, monkeypatch):
    """an edit queues its hash-table builds without waiting for the insert flags; the batch behind it looks at
    them before trusting what it meshed.  LVN_TEST_CUCKOO_RETRY=1 declares every first insertion failed: the
    tables are rehashed with fresh parameters and the batch runs again -- results unchanged"""
    monkeypatch.setenv("LVN_TEST_CUCKOO_RETRY", "1")
    ctx = lc.Compute_MeshGenContext.create(64)
    monkeypatch.delenv("LVN_TEST_CUCKOO_RETRY")
    world = oracle_mod.World(image=lc.Compute_GetNoiseImage(), default_material=0, voxels_per_chunk=64)
    try:
        chunks = np.array([[cx * 256, surface_cy * 256, 0, 256] for cx in (-1, 0)], np.int32)
        yc = surface_cy * 64 + 30.5
        for spec in [(1, 1, 201, [-2.5, yc, 30.5], [11, 11, 11]), (0, 0, 2, [10.5, yc + 6, 40.5], [9, 4, 12])]:
            op, oop = lc.CSGOperationInfo.make(*spec), oracle_mod.make_csg_op(*spec)
            assert ctx.applyCSGOperationsBatch([op], chunks) == 0
            for c in chunks:
                world.apply_csg_operations([oop], [int(v) for v in c[:3]], 256)
            res, V, T, S = batch_host(lc, ctx, chunks)
            for c, r in zip(chunks, res):
                mn = [int(v) for v in c[:3]]
                ref = world.generate_chunk_mesh(mn, 256)
                world.free_chunk_octree(mn, 256)
                nv = ref["numNodes"] if ref["numTriangles"] > 0 else 0
                assert r["numVertices"] == nv and r["numTriangles"] == ref["numTriangles"]
                assert V[r["vertexOffset"]:r["vertexOffset"] + nv].tobytes() == ref["vertices"][:nv].tobytes()
                assert np.array_equal(T["indices_"][r["triangleOffset"]:r["triangleOffset"] + r["numTriangles"]], ref["indices"])
    finally:
        ctx.destroy(); world.close()
