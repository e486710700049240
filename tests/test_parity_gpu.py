"""GPU parity tests proper: the CUDA path, called through the C ABI, against the oracle on the
same seeded inputs.  Bar (BASELINE.json north_star): integer / index outputs bit-exact; float
outputs within 1e-4 relative in voxel units -- here they are required bit-exact, because the
arithmetic spec fixes every operation; the 1e-4 tolerance is asserted as well so
that a future relaxation of the spec still has a stated bound."""
import numpy as np
import pytest

from conftest import SEED

pytestmark = pytest.mark.gpu

REL_TOL = 1e-4   # north_star: "within a stated tolerance (e.g. 1e-4 relative, in voxel units)"


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_float_parity(got, want, what):
    got, want = np.asarray(got, np.float32), np.asarray(want, np.float32)
    assert got.shape == want.shape, what
    scale = np.maximum(np.abs(want), 1.0)
    assert np.all(np.abs(got - want) <= REL_TOL * scale) or np.array_equal(bits(got), bits(want)), what
    assert np.array_equal(bits(got), bits(want)), f"{what}: not bit-exact (max abs diff {np.abs(got - want).max()})"


@pytest.fixture(scope="module")
def ctx(lc):
    c = lc.Compute_MeshGenContext.create(64)
    assert c.privateCtx_
    yield c
    c.destroy()


@pytest.fixture(scope="module")
def gpu_world(lc, oracle_mod):
    """oracle world fed the very image bytes the CUDA path uses"""
    w = oracle_mod.World(image=lc.Compute_GetNoiseImage(), default_material=0, voxels_per_chunk=64)
    yield w
    w.close()


def test_noise_image_matches_oracle(lc, oracle_mod):
    assert np.array_equal(lc.Compute_GetNoiseImage(), oracle_mod.noise_image(SEED))


def surface_chunks(world, cy):
    """a handful of chunks around the origin that contain surface, plus air / solid ones"""
    return [[0, cy * 256, 0], [256, cy * 256, 0], [-256, cy * 256, -256], [0, (cy - 1) * 256, 256],
            [512, (cy + 1) * 256, -512], [0, 15 * 256, 0], [0, 0, 0]]


def check_chunk_stages(ctx, world, mn, size=256):
    got = ctx.debugDumpChunk(mn, size)
    ref = world.generate_chunk_mesh(mn, size)
    world.free_chunk_octree(mn, size)
    # S1: materials, bit-exact (u8 vs the oracle's int)
    assert np.array_equal(got["materials"].astype(np.int32), ref["materials"]), f"materials {mn}"
    # S2: edge keys, count and canonical order
    assert got["numEdges"] == ref["numEdges"]
    assert np.array_equal(got["edgeKeys"], ref["edgeKeys"]), f"edge keys {mn}"
    # S3: Hermite normals + t
    assert_float_parity(got["edgeInfo"], ref["edgeInfo"], f"edge info {mn}")
    # S4: active voxels, edge masks, material words
    assert got["numNodes"] == ref["numNodes"]
    assert np.array_equal(got["nodeCodes"], ref["codes"]), f"node codes {mn}"
    assert np.array_equal(got["nodeEdgeMasks"], ref["edgeMasks"]), f"edge masks {mn}"
    assert np.array_equal(got["nodeMaterials"], ref["matWords"]), f"material words {mn}"
    # S5: leaf QEFs and averaged normals
    if ref["numNodes"]:
        q = ref["qefs"]
        ref_q = np.concatenate([q["ATA"], q["pad"], q["ATb"], q["masspoint"]], axis=1)
        assert_float_parity(got["nodeQEFs"], ref_q, f"QEF data {mn}")
        assert_float_parity(got["nodeNormals"], ref["normals"], f"node normals {mn}")
        # S6: solved positions
        assert_float_parity(got["nodePositions"], ref["positions"], f"QEF positions {mn}")
    assert got["numTriangles"] == ref["numTriangles"] and got["numSeamNodes"] == ref["numSeamNodes"]
    return ref


def test_config1_every_stage(ctx, gpu_world, surface_cy):
    """BASELINE config 1: the surface chunk above the world origin, every stage dumped"""
    ref = check_chunk_stages(ctx, gpu_world, [0, surface_cy * 256, 0])
    assert ref["numEdges"] > 1000 and ref["numNodes"] > 1000


def test_more_chunks_every_stage(ctx, gpu_world, surface_cy):
    for mn in surface_chunks(gpu_world, surface_cy)[1:]:
        check_chunk_stages(ctx, gpu_world, mn)


def test_lod_levels(ctx, gpu_world, surface_cy):
    """sampleScale 2 and 4 (clipmap node sizes 512, 1024; the collision generator uses 512,
    volume_constants.h:20-21)"""
    for size in (512, 1024):
        cy = int(900 * gpu_world.terrain(0.0, 0.0) * 4 // size)
        check_chunk_stages(ctx, gpu_world, [0, cy * size, 0], size)
        check_chunk_stages(ctx, gpu_world, [-size, cy * size, size], size)


def check_mesh(lc, ctx, world, mn, size=256):
    mesh, seams = lc.MeshBuffer(), []
    rc = ctx.generateChunkMesh(mn, size, mesh, seams)
    assert rc == 0, lc.GetCLErrorString(rc)
    ref = world.generate_chunk_mesh(mn, size)
    nV = ref["numNodes"] if ref["numTriangles"] > 0 else 0
    assert mesh.numVertices == nV and mesh.numTriangles == ref["numTriangles"]
    assert np.array_equal(mesh.triangles["indices_"][:mesh.numTriangles], ref["indices"]), f"indices {mn}"
    v = mesh.vertices[:mesh.numVertices]
    if nV:
        assert_float_parity(v["xyz"], ref["vertices"]["xyz"], "vertex xyz")
        assert_float_parity(v["normal"], ref["vertices"]["normal"], "vertex normal")
        assert_float_parity(v["colour"], ref["vertices"]["colour"], "vertex colour")
    s = seams[0]
    assert len(s) == ref["numSeamNodes"]
    assert np.array_equal(s["localspaceMin"], ref["seams"]["localspaceMin"])
    assert_float_parity(s["position"], ref["seams"]["position"], "seam position")
    assert_float_parity(s["normal"], ref["seams"]["normal"], "seam normal")
    return ref


def test_generate_chunk_mesh_api(lc, ctx, gpu_world, surface_cy):
    """generateChunkMesh (compute.h:68-72): MeshBuffer + seam vector contents"""
    for mn in surface_chunks(gpu_world, surface_cy):
        check_mesh(lc, ctx, gpu_world, mn)
        ctx.freeChunkOctree(mn, 256)
        gpu_world.free_chunk_octree(mn, 256)


def test_capacity_error(lc, ctx, surface_cy):
    """over-capacity is an error code, not the reference's __debugbreak (compute_octree.cpp:235-236)"""
    mesh, seams = lc.MeshBuffer(max_vertices=16), []
    rc = ctx.generateChunkMesh([0, surface_cy * 256, 0], 256, mesh, seams)
    assert rc == lc.LVN_ERR_CAPACITY
    ctx.freeChunkOctree([0, surface_cy * 256, 0], 256)


def ring(cy, r=2, ry=2):
    return np.array([[cx * 256, (cy + dy) * 256, cz * 256, 256]
                     for dy in range(-ry, ry) for cz in range(-r, r) for cx in range(-r, r)], np.int32)


def test_batch_equals_single_chunks(lc, ctx, gpu_world, surface_cy):
    """the batch entry point returns, per chunk, exactly what generateChunkMesh returns"""
    ms = ring(surface_cy, 2, 2)                      # 64 chunks
    V = np.zeros(400000, lc.MeshVertex); T = np.zeros(800000, lc.MeshTriangle); S = np.zeros(100000, lc.SeamNodeInfo)
    rc, res = ctx.generateBatch(ms, V, T, S)
    assert rc == 0, lc.GetCLErrorString(rc)
    counts, _ = gpu_world.batch_counts(ms)
    assert np.array_equal(res["numEdges"], counts[:, 0])
    assert np.array_equal(res["numTriangles"], counts[:, 2])
    assert np.array_equal(res["numSeamNodes"], counts[:, 3])
    nonempty = np.nonzero(counts[:, 1])[0]
    assert len(nonempty) >= 8
    for i in nonempty[:: max(1, len(nonempty) // 6)]:
        ref = gpu_world.generate_chunk_mesh(ms[i, :3], 256)
        gpu_world.free_chunk_octree(ms[i, :3], 256)
        r = res[i]
        v = V[r["vertexOffset"]: r["vertexOffset"] + r["numVertices"]]
        t = T[r["triangleOffset"]: r["triangleOffset"] + r["numTriangles"]]
        s = S[r["seamOffset"]: r["seamOffset"] + r["numSeamNodes"]]
        assert np.array_equal(t["indices_"], ref["indices"])
        assert_float_parity(v["xyz"], ref["vertices"]["xyz"], "batch vertex xyz")
        assert_float_parity(v["normal"], ref["vertices"]["normal"], "batch vertex normal")
        assert np.array_equal(s["localspaceMin"], ref["seams"]["localspaceMin"])
        assert_float_parity(s["position"], ref["seams"]["position"], "batch seam position")


def _chunk_bytes(res, V, T, S):
    """per chunk: the raw bytes of its mesh and seam slices (layout-independent comparison)"""
    out = []
    for r in res:
        v = V[r["vertexOffset"]: r["vertexOffset"] + r["numVertices"]]
        t = T[r["triangleOffset"]: r["triangleOffset"] + r["numTriangles"]]
        s = S[r["seamOffset"]: r["seamOffset"] + r["numSeamNodes"]]
        out.append((v.tobytes(), t.tobytes(), s.tobytes()))
    return out


def test_pipeline_lanes_and_streams(lc, surface_cy):
    """lanes x streams of the batch pipeline (lvn_meshgen_set_pipeline) change where a chunk's
    slices land, never what is in them: every setting returns, per chunk, the bytes of the
    one-lane run; host offsets tile the packed host arenas; a ragged last lane, a batch smaller
    than the lane count and a first call that must grow its arenas are covered"""
    c = lc.Compute_MeshGenContext.create(64)
    try:
        ms = ring(surface_cy, 3, 2)[:131]                  # 131 chunks: ragged lanes
        assert c.setPipeline(1, 1) == 0
        V = np.zeros(600000, lc.MeshVertex); T = np.zeros(1200000, lc.MeshTriangle); S = np.zeros(150000, lc.SeamNodeInfo)
        rc, res0 = c.generateBatch(ms, V, T, S)            # cold context: arenas grow inside this call
        assert rc == 0, lc.GetCLErrorString(rc)
        want = _chunk_bytes(res0, V, T, S)
        assert sum(len(w[0]) for w in want) > 0
        rc, _, view0 = c.generateBatchDevice(ms)
        assert rc == 0
        PV = lc.PinnedArray(int(view0.totalVertices), lc.MeshVertex)      # every node, also of chunks without quads
        PT = lc.PinnedArray(int(view0.totalTriangles), lc.MeshTriangle)
        PS = lc.PinnedArray(int(view0.totalSeamNodes), lc.SeamNodeInfo)
        for lanes, streams in ((2, 1), (2, 2), (4, 1), (5, 3), (8, 4), (32, 4), (0, 2)):
            assert c.setPipeline(lanes, streams) == 0
            V2 = np.zeros_like(V); T2 = np.zeros_like(T); S2 = np.zeros_like(S)
            rc, res = c.generateBatch(ms, V2, T2, S2)
            assert rc == 0, (lanes, streams, lc.GetCLErrorString(rc))
            got_lanes, got_streams = c.getPipeline()
            if lanes:
                assert got_lanes == min(lanes, len(ms)) and got_streams == min(streams, got_lanes)
            for k in ("numEdges", "numVertices", "numTriangles", "numSeamNodes", "status"):
                assert np.array_equal(res[k], res0[k]), (lanes, streams, k)
            assert _chunk_bytes(res, V2, T2, S2) == want, (lanes, streams)
            # the host arenas are packed: slices of non-empty chunks tile [0, total)
            ne = res[res["numSeamNodes"] > 0]
            iv = sorted((int(r["seamOffset"]), int(r["seamOffset"] + r["numSeamNodes"])) for r in ne)
            assert iv[0][0] == 0 and all(a[1] == b[0] for a, b in zip(iv, iv[1:]))
            assert iv[-1][1] == int(res["numSeamNodes"].sum())
            # pinned arenas sized exactly: same bytes per chunk
            rc, resp = c.generateBatch(ms, PV.array, PT.array, PS.array)
            assert rc == 0, (lanes, streams, lc.GetCLErrorString(rc))
            assert _chunk_bytes(resp, PV.array, PT.array, PS.array) == want, (lanes, streams, "pinned")
            # the device-resident pass agrees on the counts
            rc, resd, view = c.generateBatchDevice(ms)
            assert rc == 0
            assert np.array_equal(resd["numTriangles"], res0["numTriangles"])
            assert int(view.totalTriangles) == int(res0["numTriangles"].sum())
        # fewer chunks than lanes
        assert c.setPipeline(8, 2) == 0
        few = ms[res0["numTriangles"] > 0][:3]
        rc, resf = c.generateBatch(few, V, T, S)
        assert rc == 0 and c.getPipeline()[0] == 3
        idx = np.nonzero(res0["numTriangles"] > 0)[0][:3]
        assert _chunk_bytes(resf, V, T, S) == [want[i] for i in idx]
        # too-small host arenas: an error code and the counts the caller must provide
        assert c.setPipeline(4, 1) == 0
        rc, rese = c.generateBatch(ms, V[:100], T, S)
        assert rc == lc.LVN_ERR_CAPACITY
        assert np.array_equal(rese["numVertices"], res0["numVertices"])
        assert c.setPipeline(33, 1) < 0 and c.setPipeline(2, 0) < 0 and c.setPipeline(2, 5) < 0
        PV.close(); PT.close(); PS.close()
    finally:
        c.destroy()


def test_batch_full_size_properties(lc, ctx, surface_cy):
    """config 2 at full size (512 chunks): size-independent properties -- offsets tile the arenas,
    every index addresses its own chunk's vertices, quads are consistent, a second run is identical"""
    ms = ring(surface_cy, 4, 4)
    assert len(ms) == 512
    rc, res, view = ctx.generateBatchDevice(ms)
    assert rc == 0, lc.GetCLErrorString(rc)
    nV, nT, nS = int(view.totalVertices), int(view.totalTriangles), int(view.totalSeamNodes)
    V = np.zeros(nV, lc.MeshVertex); T = np.zeros(nT, lc.MeshTriangle); S = np.zeros(nS, lc.SeamNodeInfo)
    rc, res = ctx.generateBatch(ms, V, T, S)
    assert rc == 0
    ne = res[res["numTriangles"] > 0]
    assert ne["numVertices"].sum() <= nV and res["numTriangles"].sum() == nT and res["numSeamNodes"].sum() == nS
    # slices do not overlap
    iv = sorted((int(r["vertexOffset"]), int(r["vertexOffset"] + r["numVertices"])) for r in ne)
    assert all(a[1] <= b[0] for a, b in zip(iv, iv[1:]))
    digest = []
    for r in ne:
        t = T[r["triangleOffset"]: r["triangleOffset"] + r["numTriangles"]]["indices_"]
        assert t.min() >= 0 and t.max() < r["numVertices"]
        q = t.reshape(-1, 6)
        assert np.all(q[:, 0] == q[:, 3])            # both triangles of a quad start at the emitting node
        v = V[r["vertexOffset"]: r["vertexOffset"] + r["numVertices"]]
        assert np.all(v["xyz"][:, 3] == 1.0) and np.all(np.isfinite(v["xyz"])) and np.all(np.isfinite(v["normal"]))
        digest.append((int(r["numVertices"]), int(r["numTriangles"]), int(t.astype(np.int64).sum()),
                       float(v["xyz"].astype(np.float64).sum())))
    # idempotence: checksum of per-chunk checksums is identical on a second pass
    V2 = np.zeros_like(V); T2 = np.zeros_like(T); S2 = np.zeros_like(S)
    rc, res2 = ctx.generateBatch(ms, V2, T2, S2)
    assert rc == 0
    for k in ("numEdges", "numVertices", "numTriangles", "numSeamNodes"):
        assert np.array_equal(res[k], res2[k])
    digest2 = []
    for r in res2[res2["numTriangles"] > 0]:
        t = T2[r["triangleOffset"]: r["triangleOffset"] + r["numTriangles"]]["indices_"]
        v = V2[r["vertexOffset"]: r["vertexOffset"] + r["numVertices"]]
        digest2.append((int(r["numVertices"]), int(r["numTriangles"]), int(t.astype(np.int64).sum()),
                        float(v["xyz"].astype(np.float64).sum())))
    assert digest == digest2


def test_octree_cache_semantics(lc, oracle_mod, surface_cy):
    """stale until freeChunkOctree (compute_octree.cpp:154-181,379-387), same as the oracle twin"""
    ctx = lc.Compute_MeshGenContext.create(64)
    world = oracle_mod.World(image=lc.Compute_GetNoiseImage(), default_material=0, voxels_per_chunk=64)
    try:
        mn = [0, surface_cy * 256, 0]
        a = check_mesh(lc, ctx, world, mn)
        assert a["numNodes"] > 0
        yc = surface_cy * 64 + 32.5
        specs = [(1, 1, 201, [20.5, yc, 20.5], [6, 6, 6]), (0, 0, 3, [44.5, yc, 44.5], [5, 4, 3])]
        assert ctx.applyCSGOperations([lc.CSGOperationInfo.make(*s) for s in specs], mn, 256) == 0
        world.apply_csg_operations([oracle_mod.make_csg_op(*s) for s in specs], mn, 256)
        b = check_mesh(lc, ctx, world, mn)                 # both still serve the cached octree
        assert b["numNodes"] == a["numNodes"]
        ctx.freeChunkOctree(mn, 256)
        world.free_chunk_octree(mn, 256)
        c = check_mesh(lc, ctx, world, mn)                 # re-meshed from the edited field
        assert c["numNodes"] != a["numNodes"]
        # a chunk that was empty is not cached at all: an edit shows up without freeChunkOctree
        em = [0, 15 * 256, 0]
        assert check_mesh(lc, ctx, world, em)["numNodes"] == 0
        blob = (0, 1, 2, [30.5, 15 * 64 + 30.5, 30.5], [8, 8, 8])
        assert ctx.applyCSGOperations([lc.CSGOperationInfo.make(*blob)], em, 256) == 0
        world.apply_csg_operations([oracle_mod.make_csg_op(*blob)], em, 256)
        assert check_mesh(lc, ctx, world, em)["numNodes"] > 0
    finally:
        ctx.destroy(); world.close()


def compare_csg_field(ctx, world, mn, size=256):
    got = ctx.debugDumpChunk(mn, size)
    ref = world.generate_chunk_mesh(mn, size)
    world.free_chunk_octree(mn, size)
    assert np.array_equal(got["materials"].astype(np.int32), ref["materials"]), "CSG materials"
    # the edge list order is unobservable (hash lookups); compare as key -> (normal, t) maps
    assert got["numEdges"] == ref["numEdges"]
    go, ro = np.argsort(got["edgeKeys"]), np.argsort(ref["edgeKeys"])
    assert np.array_equal(got["edgeKeys"][go], ref["edgeKeys"][ro]), "CSG edge set"
    assert_float_parity(got["edgeInfo"][go], ref["edgeInfo"][ro], "CSG edge info")
    assert np.array_equal(got["nodeCodes"], ref["codes"])
    assert np.array_equal(got["nodeEdgeMasks"], ref["edgeMasks"])
    assert np.array_equal(got["nodeMaterials"], ref["matWords"])
    if ref["numNodes"]:
        assert_float_parity(got["nodePositions"], ref["positions"], "CSG positions")
        assert_float_parity(got["nodeNormals"], ref["normals"], "CSG normals")
    return ref


def test_csg_edit_sequence(lc, oracle_mod, gpu_world, surface_cy):
    """BASELINE config 3 in miniature: scripted sphere / cube add / subtract ops, applied one per
    step to the chunks they overlap, then stored for replay (clipmap.cpp:1647-1744)"""
    ctx = lc.Compute_MeshGenContext.create(64)
    world = oracle_mod.World(image=lc.Compute_GetNoiseImage(), default_material=0, voxels_per_chunk=64)
    rng = np.random.RandomState(12345)
    chunks = [[cx * 256, (surface_cy + dy) * 256, cz * 256] for dy in (-1, 0) for cz in (0, 1) for cx in (0, 1)]
    try:
        for step in range(6):
            shape = step % 2                      # alternate cube / sphere
            add = (step // 2) % 2 == 0
            origin = [float(rng.randint(20, 108)) + 0.5, surface_cy * 64 + float(rng.randint(-20, 40)) + 0.5,
                      float(rng.randint(20, 108)) + 0.5]
            dims = [float(rng.randint(1, 12)), float(rng.randint(1, 12)), float(rng.randint(1, 12))]
            mat = int(rng.randint(1, 4)) if add else 201
            rot = 0.0 if step < 4 else 0.5        # the app never rotates; the kernel supports it
            op = lc.CSGOperationInfo.make(0 if add else 1, shape, mat, origin, dims, rot)
            oop = oracle_mod.make_csg_op(0 if add else 1, shape, mat, origin, dims, rot)
            lo, hi = lc.CalcCSGOperationBounds(op)
            assert (lo, hi) == tuple(oracle_mod.csg_operation_bounds(oop))
            touched = [c for c in chunks if not (c[0] + 256 < lo[0] or c[1] + 256 < lo[1] or c[2] + 256 < lo[2] or
                                                 c[0] > hi[0] or c[1] > hi[1] or c[2] > hi[2])]
            if step % 2 == 0 and touched:
                # one batched call for every overlapping chunk (lvn_meshgen_apply_csg_operations_batch)
                assert ctx.applyCSGOperationsBatch([op], np.array([c + [256] for c in touched], np.int32)) == 0
            for c in touched:
                if step % 2 == 1:
                    assert ctx.applyCSGOperations([op], c, 256) == 0
                world.apply_csg_operations([oop], c, 256)
                ctx.freeChunkOctree(c, 256)
                world.free_chunk_octree(c, 256)
            assert lc.Compute_StoreCSGOperation(op, lo, hi) == 0
            world.store_csg_operation(oop, lo, hi)
            for c in touched:
                compare_csg_field(ctx, world, c)
                check_mesh(lc, ctx, world, c)
    finally:
        lc.Compute_ClearCSGOperations()
        ctx.destroy(); world.close()


def test_csg_replay_on_fresh_context(lc, oracle_mod, surface_cy):
    """stored ops are replayed lazily onto regenerated fields (LoadDensityField)"""
    ctx = lc.Compute_MeshGenContext.create(64)
    world = oracle_mod.World(image=lc.Compute_GetNoiseImage(), default_material=0, voxels_per_chunk=64)
    try:
        yc = surface_cy * 64 + 30.5
        specs = [(1, 1, 201, [30.5, yc, 30.5], [9, 9, 9]), (0, 0, 2, [60.5, yc + 8, 40.5], [7, 3, 11]),
                 (0, 1, 3, [64.5, yc, 64.5], [5, 5, 5])]
        for s in specs:
            op, oop = lc.CSGOperationInfo.make(*s), oracle_mod.make_csg_op(*s)
            lo, hi = lc.CalcCSGOperationBounds(op)
            lc.Compute_StoreCSGOperation(op, lo, hi)
            world.store_csg_operation(oop, lo, hi)
        for mn in ([0, surface_cy * 256, 0], [256, surface_cy * 256, 0], [0, surface_cy * 256, 256]):
            compare_csg_field(ctx, world, mn)
            check_mesh(lc, ctx, world, mn)
        rc, empty = ctx.isChunkEmpty([0, 15 * 256, 0], 256)
        assert rc == 0 and empty and world.is_chunk_empty([0, 15 * 256, 0], 256)
        rc, empty = ctx.isChunkEmpty([512, surface_cy * 256, 512], 256)
        assert rc == 0 and empty == world.is_chunk_empty([512, surface_cy * 256, 512], 256)
    finally:
        lc.Compute_ClearCSGOperations()
        ctx.destroy(); world.close()


def test_smaller_chunk_sizes(lc, oracle_mod):
    for V in (16, 32):
        ctx = lc.Compute_MeshGenContext.create(V)
        world = oracle_mod.World(image=lc.Compute_GetNoiseImage(), default_material=0, voxels_per_chunk=V)
        size = V * 4
        cy = int(900 * world.terrain(0.0, 0.0) * 4 // size)
        check_chunk_stages(ctx, world, [0, cy * size, 0], size)
        check_mesh(lc, ctx, world, [size, cy * size, 0], size)
        ctx.destroy(); world.close()
    assert lc.Compute_MeshGenContext.create(48).privateCtx_ is None


def test_default_material_and_seed(lc, oracle_mod):
    """another noise seed and a non-zero default material"""
    try:
        assert lc.Compute_Initialise(0x7d3af, 7, 2) == 0          # test_compute.cpp:24
        ctx = lc.Compute_MeshGenContext.create(64)
        world = oracle_mod.World(image=lc.Compute_GetNoiseImage(), default_material=7, voxels_per_chunk=64)
        assert np.array_equal(world.image, oracle_mod.noise_image(0x7d3af))
        cy = int(900 * world.terrain(0.0, 0.0) // 64)
        ref = check_chunk_stages(ctx, world, [0, cy * 256, 0])
        assert np.all((ref["matWords"] >> 8) == 7)
        ctx.destroy(); world.close()
        # a default material that sorts AFTER MATERIAL_AIR (201): FindDominantMaterial's sorted-run
        # scan (octree.cl:96-138) then starts on AIR and its tie rule can return AIR -- the quirk is
        # part of the restated behaviour
        assert lc.Compute_Initialise(0x7d3af, 250, 2) == 0
        ctx = lc.Compute_MeshGenContext.create(64)
        world = oracle_mod.World(image=lc.Compute_GetNoiseImage(), default_material=250, voxels_per_chunk=64)
        ref = check_chunk_stages(ctx, world, [0, cy * 256, 0])
        assert set(np.unique(ref["matWords"] >> 8)) <= {250, 201}
        ctx.destroy(); world.close()
    finally:
        assert lc.Compute_Initialise(SEED, 0, 2) == 0


def test_stress_field(lc, oracle_mod):
    """BASELINE config 4: ridged 3-D fBm from snoise3, ~30% active voxels, arenas > 14336 vertices"""
    thr = 0.735       # tuned once on the oracle: ~30% of the voxels active
    try:
        lc.Compute_SetDensityFunction(1, thr)
        ctx = lc.Compute_MeshGenContext.create(64)
        world = oracle_mod.World(image=lc.Compute_GetNoiseImage(), default_material=0, voxels_per_chunk=64)
        world.set_density(1, thr)
        ref = check_chunk_stages(ctx, world, [0, 0, 0])
        assert ref["numNodes"] > 14336
        mesh, seams = lc.MeshBuffer(), []
        assert ctx.generateChunkMesh([0, 0, 0], 256, mesh, seams) == lc.LVN_ERR_CAPACITY   # the reference asserts
        ctx.freeChunkOctree([0, 0, 0], 256)
        big = lc.MeshBuffer(max_vertices=1 << 18)
        assert ctx.generateChunkMesh([0, 0, 0], 256, big, seams) == 0
        assert big.numVertices == ref["numNodes"]
        assert np.array_equal(big.triangles["indices_"][:big.numTriangles], ref["indices"])
        ctx.destroy(); world.close()
    finally:
        lc.Compute_SetDensityFunction(0, 0.5)


# ---- a9 / a15 utilities through the C ABI -------------------------------------------------
def test_gpu_cuckoo_octree_keys(lc, golden_keys):
    """test_cuckoo.cpp:107-178 against the GPU table: every key inserts, every key is found, the
    stored value is the key's index (cuckoo.cl:35)"""
    for name, keys in sorted(golden_keys.items()):
        if name.endswith("duplicated"):
            continue
        t = lc.CuckooData()
        assert t.Cuckoo_InitialiseTable(len(keys)) == 0
        assert t.prime == lc.FindNextPrime(max(2048, 2 * len(keys)))
        assert t.Cuckoo_InsertKeys(keys) == 0
        vals = t.Cuckoo_Find(keys)
        assert np.array_equal(vals, np.arange(len(keys), dtype=np.uint32)), name
        missing = t.Cuckoo_Find(np.array([0xfffffff0, 0x7fffffff], np.uint32))
        assert np.all(missing == 0xffffffff)
        t.destroy()


def test_gpu_cuckoo_100_keys(lc):
    """Compute (Cuckoo), test_compute.cpp:70-88"""
    t = lc.CuckooData()
    assert t.Cuckoo_InitialiseTable(100) == 0
    assert t.Cuckoo_InsertKeys(np.arange(100, dtype=np.uint32)) == 0
    assert np.array_equal(t.Cuckoo_Find(np.arange(100, dtype=np.uint32)), np.arange(100, dtype=np.uint32))
    t.destroy()


def test_gpu_cuckoo_large_random(lc):
    """2^20 random unique keys (test_cuckoo.cpp:35-47 generates the same population size)"""
    rng = np.random.RandomState(7)
    keys = np.unique(rng.randint(0, 2 ** 32 - 2, size=(1 << 20) + 5000, dtype=np.uint64).astype(np.uint32))[:1 << 20]
    rng.shuffle(keys)
    t = lc.CuckooData()
    assert t.Cuckoo_InitialiseTable(len(keys)) == 0
    assert t.Cuckoo_InsertKeys(keys) == 0
    assert np.array_equal(t.Cuckoo_Find(keys), np.arange(len(keys), dtype=np.uint32))
    t.destroy()


def test_gpu_remove_duplicates(lc, golden_keys):
    """Compute (Remove Duplicates), test_compute.cpp:46-68: set equality after sorting"""
    out = lc.RemoveDuplicates(golden_keys["keys_3_duplicated"].view(np.int32))
    assert len(out) == len(golden_keys["keys_3"])
    assert np.array_equal(np.sort(out.view(np.uint32)), np.sort(golden_keys["keys_3"]))


def test_gpu_scan_compact(lc, oracle_mod):
    rng = np.random.RandomState(0)
    for n in [1, 2, 255, 256, 257, 4096, 4097, 262144, 823875, 5000000]:
        data = rng.randint(0, 3, size=n).astype(np.int32)
        total, scan = lc.ExclusiveScan(data)
        assert total == int(data.sum())
        assert np.array_equal(scan, np.cumsum(data) - data)
        vals = rng.randint(0, 1 << 30, size=n).astype(np.int32)
        valid = (data != 0).astype(np.int32)          # validity flags are 0 / 1 (compact.cl:4-16)
        assert np.array_equal(lc.CompactIndexArray(vals, valid), vals[valid != 0])
    assert lc.ExclusiveScan(np.zeros(0, np.int32))[0] == 0
