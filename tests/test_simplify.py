"""Mesh simplification on the GPU: ngMeshSimplifier, the pass the clipmap runs on
every chunk mesh right after export (clipmap.cpp:449-465,495-501).

Checker: the reference's own leven/src/ng_mesh_simplify.cpp + qef_simd.h, compiled for the host into
oracle/_ref/libleven_simplify_ref.so (oracle/ref_shim/ref_simplify.cpp says which two
platform-defined pieces the build has to fix: _mm_rsqrt_ps := 1 / sqrt, and libstdc++'s
std::uniform_int_distribution).  Where the library is absent the committed digests
(tests/golden/ref_simplify.npz, generated from it by tests/golden/gen_ref_vectors.py simplify)
stand in.

Bar: bit-identical -- same vertices in the same order, same triangles in the same order."""
import hashlib
import os

import numpy as np
import pytest

import simplify_scenarios as S
from conftest import ROOT, SEED

GOLDEN = os.path.join(ROOT, "tests", "golden", "ref_simplify.npz")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def cases(world, oracle_mod, surface_cy):
    w2 = oracle_mod.World(seed=SEED, default_material=0, voxels_per_chunk=64)
    try:
        return S.all_cases(world, oracle_mod, surface_cy, w2)
    finally:
        w2.close()


@pytest.fixture(scope="module")
def golden():
    g = np.load(GOLDEN)
    return {str(n): [str(x) for x in r] for n, r in zip(g["cases"], g["rows"])}, g


@pytest.fixture(scope="module")
def simplify_ref(built):
    from oracle import ref as R
    if not R.simplify_available():
        pytest.skip("oracle/_ref/libleven_simplify_ref.so not built (needs /root/reference)")
    return R


def test_inputs_match_golden(cases, golden):
    """the scenarios rebuild the very inputs the fixture was generated from"""
    rows, _ = golden
    assert [c[0] for c in cases] == list(rows)
    for name, v, t, off, opt in cases:
        r = rows[name]
        assert (str(len(v)), str(len(t)), sha(S.as_vertices(v)), sha(np.asarray(t, np.int32))) == tuple(r[:4]), name


def test_reference_simplifier_matches_golden(simplify_ref, cases, golden):
    """the shipped digests are what the reference code computes (guards the fixture itself)"""
    rows, _ = golden
    for name, v, t, off, opt in cases:
        rv, rt = simplify_ref.simplify_mesh(S.as_vertices(v), t, off, opt)
        assert (str(len(rv)), str(len(rt)), sha(rv), sha(rt)) == tuple(rows[name][4:]), name


def test_simplifier_invariants(simplify_ref, cases):
    """what the algorithm promises whatever the arithmetic: a mesh never grows, indices stay in
    range, no degenerate triangle survives, every kept vertex is used, small meshes are untouched,
    boundary vertices of an open mesh do not move"""
    for name, v, t, off, opt in cases:
        vin = S.as_vertices(v)
        rv, rt = simplify_ref.simplify_mesh(vin, t, off, opt)
        assert len(rv) <= len(vin) and len(rt) <= len(t)
        if len(t) < 100 or len(vin) < 100:
            assert rv.tobytes() == vin.tobytes() and np.array_equal(rt, t)
            continue
        assert rt.min() >= 0 and rt.max() < len(rv)
        assert not ((rt[:, 0] == rt[:, 1]) | (rt[:, 0] == rt[:, 2]) | (rt[:, 1] == rt[:, 2])).any()
        assert len(np.unique(rt)) == len(rv)
    v, t = S.grid_plane(24, 4.0, 0.0)
    rv, rt = simplify_ref.simplify_mesh(v, t, [46.0, 0.0, 46.0], dict(S.DEFAULTS, maxError=50.0, maxEdgeSize=20.0, minAngleCosine=0.5))
    rim = lambda a: a[(a[:, 0] == 0) | (a[:, 2] == 0) | (a[:, 0] == 92) | (a[:, 2] == 92)]
    a, b = rim(v["xyz"]), rim(rv["xyz"])
    assert len(a) == len(b) == 4 * 23 and {tuple(x) for x in a.tolist()} == {tuple(x) for x in b.tolist()}


def libstdcpp_uniform_int(raw, n_edges, count):
    """std::uniform_int_distribution<int>(0, n - 1) over a 32-bit engine in libstdc++ (GCC >= 11):
    product = raw * n; accept unless (uint32)product < (2^32 - n) % n; result = product >> 32.
    This is the statement the kernel implements over the precomputed mt19937(42) stream."""
    prod = raw.astype(np.uint64) * np.uint64(n_edges)
    ok = (prod & np.uint64(0xFFFFFFFF)) >= np.uint64(((1 << 32) - n_edges) % n_edges)
    return (prod >> np.uint64(32))[ok][:count].astype(np.int32)


def test_candidate_sampler_statement(golden):
    """ng_mesh_simplify.cpp:195-205 draws its candidate edges from std::mt19937(42): the engine is
    standardised (numpy's MT19937 gives the same raw words under init_genrand), the distribution is
    libstdc++'s; the committed samples came from the reference build"""
    _, g = golden
    mt = np.random.MT19937()
    mt._legacy_seeding(42)
    raw = mt.random_raw(8192).astype(np.uint32)
    assert raw[0] == 1608637542            # mt19937(42)'s first output
    for key in [k for k in g.files if k.startswith("random_edges/")]:
        n = int(key.split("/")[1])
        assert np.array_equal(libstdcpp_uniform_int(raw, n, 4096), g[key]), key


@pytest.mark.gpu
def test_gpu_simplify_matches_reference(lc, cases, golden, built):
    """every scenario in ONE batch launch, each mesh with its own options"""
    from oracle import ref as R
    rows, _ = golden
    meshes = [(S.as_vertices(v), t, off) for name, v, t, off, opt in cases]
    opts = [lc.SimplifyOptions.make(**opt) for name, v, t, off, opt in cases]
    rc, out, res = lc.ngMeshSimplifierBatch(meshes, opts)
    assert rc == 0, lc.lib().lvn_mesh_simplify_last_error()
    changed = 0
    for (name, v, t, off, opt), (gv, gt), r in zip(cases, out, res):
        gvv = S.as_vertices(gv)
        got = (str(len(gvv)), str(len(gt)), sha(gvv), sha(gt["indices_"]))
        assert got == tuple(rows[name][4:]), (name, "vs golden", got[:2], rows[name][4:6])
        if R.simplify_available():
            rv, rt = R.simplify_mesh(S.as_vertices(v), t, off, opt)
            for f in ("xyz", "normal", "colour"):
                assert gvv[f].tobytes() == rv[f].tobytes(), (name, f)
            assert np.array_equal(gt["indices_"], rt), name
        changed += len(gt) < len(t)
        assert (r["iterations"] == 0) == (len(t) < 100 or len(v) < 100), name
    assert changed >= len(cases) - 3


@pytest.mark.gpu
def test_gpu_simplify_shared_options_and_order(lc, cases, golden):
    """one options struct for the whole batch; the result of a mesh does not depend on its
    neighbours in the batch or its place in it"""
    rows, _ = golden
    sub = [c for c in cases if c[0].startswith("ring_")][:6]
    meshes = [(S.as_vertices(v), t, off) for name, v, t, off, opt in sub]
    rc, a, _ = lc.ngMeshSimplifierBatch(meshes, lc.SimplifyOptions.for_clipmap_node(256))
    assert rc == 0
    rc, b, _ = lc.ngMeshSimplifierBatch(meshes[::-1], lc.SimplifyOptions.for_clipmap_node(256))
    assert rc == 0
    for (name, *_), (av, at), (bv, bt) in zip(sub, a, b[::-1]):
        assert (sha(S.as_vertices(av)), sha(at["indices_"])) == tuple(rows[name][6:]), name
        assert av.tobytes() == bv.tobytes() and at.tobytes() == bt.tobytes()


@pytest.mark.gpu
def test_gpu_simplify_chunk_meshes_from_the_path(lc, surface_cy, built):
    """the call sequence of ConstructClipmapNodeData: generateChunkMesh -> ngMeshSimplifier, over a
    block of the ring, the meshes coming from the CUDA path's own export"""
    from oracle import ref as R
    ctx = lc.Compute_MeshGenContext.create(64)
    try:
        chunks = [[cx * 256, (surface_cy + dy) * 256, cz * 256, 256] for cx in range(-2, 2) for dy in (-1, 0) for cz in range(-2, 2)]
        rc, res, view = ctx.generateBatchDevice(chunks)
        V = np.zeros(int(view.totalVertices) + 1, lc.MeshVertex)
        T = np.zeros(int(view.totalTriangles) + 1, lc.MeshTriangle)
        Sn = np.zeros(int(view.totalSeamNodes) + 1, lc.SeamNodeInfo)
        rc, res = ctx.generateBatch(chunks, V, T, Sn)
        assert rc == 0
        meshes = []
        for c, r in zip(chunks, res):
            if r["numTriangles"]:
                meshes.append((V[r["vertexOffset"]:r["vertexOffset"] + r["numVertices"]],
                               T[r["triangleOffset"]:r["triangleOffset"] + r["numTriangles"]], [c[0] + 128.0, c[1] + 128.0, c[2] + 128.0]))
        assert len(meshes) >= 8
        rc, out, sres = lc.ngMeshSimplifierBatch(meshes, lc.SimplifyOptions.for_clipmap_node(256))
        assert rc == 0
        assert int(sres["numTriangles"].sum()) < 0.75 * sum(len(m[1]) for m in meshes)
        for (v, t, off), (gv, gt) in zip(meshes, out):
            assert gt["indices_"].max() < len(gv)
            if R.simplify_available():
                rv, rt = R.simplify_mesh(S.as_vertices(v), t["indices_"], off, S.clipmap_options(256))
                assert S.as_vertices(gv).tobytes() == rv.tobytes() and np.array_equal(gt["indices_"], rt)
    finally:
        ctx.destroy()


@pytest.mark.gpu
def test_gpu_simplify_edge_cases(lc):
    import ctypes as C
    c = lc
    opt = c.SimplifyOptions.make()
    assert c.ngMeshSimplifierBatch([], opt)[0] == 0
    # an empty mesh and a single triangle among real ones
    v, t = S.grid_plane(24, 4.0, 0.0)
    one = (S.as_vertices(v[:3]), np.array([[0, 1, 2]], np.int32), [0, 0, 0])
    empty = (S.as_vertices(v[:0]), np.zeros((0, 3), np.int32), [0, 0, 0])
    rc, out, res = c.ngMeshSimplifierBatch([empty, (v, t, [46.0, 0.0, 46.0]), one], c.SimplifyOptions.make(maxError=50.0, maxEdgeSize=20.0, minAngleCosine=0.5))
    assert rc == 0 and len(out[0][0]) == 0 and len(out[0][1]) == 0
    assert len(out[2][0]) == 3 and out[2][1]["indices_"].tolist() == [[0, 1, 2]]
    assert 0 < len(out[1][1]) < len(t)
    # maxIterations 0: nothing collapses, unused vertices are still dropped and indices stay valid
    rc, out, res = c.ngMeshSimplifierBatch([(v, t, [0, 0, 0])], c.SimplifyOptions.make(maxIterations=0))
    assert rc == 0 and len(out[0][1]) == len(t) and np.array_equal(out[0][1]["indices_"], t)
    # a wild triangle index: that mesh passes through, the others are simplified, the call says so
    tb = t.copy(); tb[17, 1] = len(v) + 5
    rc, out, res = c.ngMeshSimplifierBatch([(v, tb, [0, 0, 0]), (v, t, [0, 0, 0])], c.SimplifyOptions.make(maxError=50.0, maxEdgeSize=20.0, minAngleCosine=0.5))
    assert rc == c.LVN_ERR_INVALID_VALUE and res[0]["iterations"] == -1 and np.array_equal(out[0][1]["indices_"], tb)
    assert res[1]["iterations"] > 0 and len(out[1][1]) < len(t)
    # invalid arguments: a slice outside the arrays; an options count that is neither 1 nor numMeshes
    jobs = np.zeros(1, c.SimplifyJob); jobs[0]["numVertices"] = 10; jobs[0]["numTriangles"] = 5
    V = np.zeros(4, c.MeshVertex); T = np.zeros(5, c.MeshTriangle); r = np.zeros(1, c.SimplifyResult)
    o1 = (c.SimplifyOptions * 1)(opt)
    assert c.lib().lvn_mesh_simplify_batch(1, c._ptr(jobs), o1, 1, c._ptr(V), 4, c._ptr(T), 5, c._ptr(r)) == c.LVN_ERR_INVALID_VALUE
    jobs[0]["numVertices"] = 4
    assert c.lib().lvn_mesh_simplify_batch(1, c._ptr(jobs), o1, 2, c._ptr(V), 4, c._ptr(T), 5, c._ptr(r)) == c.LVN_ERR_INVALID_VALUE
    assert c.lib().lvn_mesh_simplify_batch(1, c._ptr(jobs), None, 1, c._ptr(V), 4, c._ptr(T), 5, c._ptr(r)) == c.LVN_ERR_INVALID_VALUE


@pytest.mark.gpu
def test_gpu_generate_simplified_batch(lc, world, surface_cy, golden, built):
    """lvn_meshgen_generate_simplified_batch = ConstructClipmapNodeData over a batch (mixed LODs, empty
    chunks among them): same meshes as generateChunkMesh followed by ngMeshSimplifier with the
    clipmap's per-node options -- against the committed digests of the reference simplifier, against
    the two-step GPU route, and against the reference itself where it is built"""
    from oracle import ref as R
    rows, _ = golden
    named = S.chunk_cases(world, surface_cy)
    chunks = [list(mn) + [size] for name, mn, size in named] + [[0, 15 * 256, 0, 256], [0, 0, 0, 256]]   # + air, solid
    ctx = lc.Compute_MeshGenContext.create(64)
    try:
        V = np.zeros(200000, lc.MeshVertex); T = np.zeros(400000, lc.MeshTriangle); Sn = np.zeros(100000, lc.SeamNodeInfo)
        rc, res, simp = ctx.generateSimplifiedBatch(chunks, V, T, Sn)
        assert rc == 0, lc.last_cuda_error()
        V2 = np.zeros(200000, lc.MeshVertex); T2 = np.zeros(400000, lc.MeshTriangle); Sn2 = np.zeros(100000, lc.SeamNodeInfo)
        rc, res2 = ctx.generateBatch(chunks, V2, T2, Sn2)
        assert rc == 0
        assert res[-1]["numVertices"] == 0 and res[-2]["numVertices"] == 0 and simp[-1]["iterations"] == 0
        # dense packing: the meshes' slices tile the front of the two arenas without a gap (chunk order
        # within the early and within the late group of a split simplifier launch)
        nz = sorted((r for r in res if r["numTriangles"]), key=lambda r: int(r["vertexOffset"]))
        assert nz[0]["vertexOffset"] == 0 and nz[0]["triangleOffset"] == 0
        for a, b in zip(nz, nz[1:]):
            assert b["vertexOffset"] == a["vertexOffset"] + a["numVertices"] and b["triangleOffset"] == a["triangleOffset"] + a["numTriangles"]
        seen = 0
        for (name, mn, size), r, r2, sr in zip(named, res, res2, simp):
            # seam nodes are those of the unsimplified octree
            assert r["numSeamNodes"] == r2["numSeamNodes"] and r["numEdges"] == r2["numEdges"]
            assert Sn[r["seamOffset"]:r["seamOffset"] + r["numSeamNodes"]].tobytes() == Sn2[r2["seamOffset"]:r2["seamOffset"] + r2["numSeamNodes"]].tobytes()
            if r2["numTriangles"] == 0:
                assert r["numTriangles"] == 0
                continue
            gv = S.as_vertices(V[r["vertexOffset"]:r["vertexOffset"] + r["numVertices"]])
            gt = T["indices_"][r["triangleOffset"]:r["triangleOffset"] + r["numTriangles"]]
            assert (str(len(gv)), str(len(gt)), sha(gv), sha(gt)) == tuple(rows[name][4:]), (name, "vs golden")
            assert sr["numVertices"] == len(gv) and sr["numTriangles"] == len(gt) and sr["iterations"] > 0
            v2 = V2[r2["vertexOffset"]:r2["vertexOffset"] + r2["numVertices"]]
            t2 = T2[r2["triangleOffset"]:r2["triangleOffset"] + r2["numTriangles"]]
            centre = [mn[0] + size / 2.0, mn[1] + size / 2.0, mn[2] + size / 2.0]
            rc, out, _ = lc.ngMeshSimplifierBatch([(v2, t2, centre)], lc.SimplifyOptions.for_clipmap_node(size))
            assert rc == 0 and S.as_vertices(out[0][0]).tobytes() == gv.tobytes() and np.array_equal(out[0][1]["indices_"], gt)
            if R.simplify_available():
                rv, rt = R.simplify_mesh(S.as_vertices(v2), t2["indices_"], centre, S.clipmap_options(size))
                assert rv.tobytes() == gv.tobytes() and np.array_equal(rt, gt), name
            seen += 1
        assert seen >= 15
        # arenas too small: LVN_ERR_CAPACITY, and the counts say what is needed
        rc, res3, _ = ctx.generateSimplifiedBatch(chunks, V[:100], T[:100], Sn)
        assert rc == lc.LVN_ERR_CAPACITY and np.array_equal(res3["numVertices"], res["numVertices"])
        rc, res3, _ = ctx.generateSimplifiedBatch(chunks, V, T, Sn[:10])
        assert rc == lc.LVN_ERR_CAPACITY and np.array_equal(res3["numSeamNodes"], res["numSeamNodes"])
        # an empty batch, a batch of empty chunks
        assert ctx.generateSimplifiedBatch(np.zeros((0, 4), np.int32), V, T, Sn)[0] == 0
        rc, res4, _ = ctx.generateSimplifiedBatch([[0, 15 * 256, 0, 256]] * 3, V, T, Sn)
        assert rc == 0 and res4["numTriangles"].sum() == 0
    finally:
        ctx.destroy()


@pytest.mark.gpu
def test_gpu_collision_batch(lc, world, surface_cy, golden, built):
    """lvn_meshgen_generate_collision_batch = Clipmap::loadCollisionNodes' per-node work: collision nodes
    are COLLISION_NODE_SIZE = 512 on a 64-voxel context (volume_constants.h:20-21), simplified, and
    handed to Bullet as vec4 positions relative to the node centre, scaled by PHYSICS_SCALE
    (physics.cpp:549-573) -- against the simplified-batch route + the conversion in numpy float32,
    and the committed digests of the reference simplifier for the two nodes the scenarios hold"""
    rows, _ = golden
    y1 = (surface_cy * 256 // 512) * 512
    nodes = [[x, y, z, 512] for x in (-512, 0) for y in (y1 - 512, y1, y1 + 512) for z in (0, 512)]
    ctx = lc.Compute_MeshGenContext.create(64)
    try:
        V = np.zeros(150000, lc.MeshVertex); T = np.zeros(300000, lc.MeshTriangle); Sn = np.zeros(100000, lc.SeamNodeInfo)
        rc, res, simp = ctx.generateSimplifiedBatch(nodes, V, T, Sn)
        assert rc == 0
        P = np.zeros((150000, 4), np.float32); T2 = np.zeros((300000, 3), np.int32); Sn2 = np.zeros(100000, lc.SeamNodeInfo)
        V2 = np.zeros(150000, lc.MeshVertex)
        rc, res2, simp2 = ctx.generateCollisionBatch(nodes, P, T2, Sn2, vertices=V2)
        assert rc == 0, lc.last_cuda_error()
        same = lambda a, b: all(np.array_equal(a[f], b[f]) for f in a.dtype.names if f != "seamOffset")
        assert same(res, res2) and simp.tobytes() == simp2.tobytes()      # (a chunk's place in the seam arena is not fixed)
        nv, nt = int(res["numVertices"].sum()), int(res["numTriangles"].sum())
        assert nt > 20000
        assert V[:nv].tobytes() == V2[:nv].tobytes() and np.array_equal(T["indices_"][:nt], T2[:nt])
        for n, r, r2 in zip(nodes, res, res2):
            assert Sn[r["seamOffset"]:r["seamOffset"] + r["numSeamNodes"]].tobytes() == Sn2[r2["seamOffset"]:r2["seamOffset"] + r2["numSeamNodes"]].tobytes()
            origin = np.array([n[0] + 256, n[1] + 256, n[2] + 256, 0], np.float32)
            want = (V["xyz"][r["vertexOffset"]:r["vertexOffset"] + r["numVertices"]] - origin) * np.float32(0.05)
            got = P[r["vertexOffset"]:r["vertexOffset"] + r["numVertices"]]
            assert got.tobytes() == want.astype(np.float32).tobytes()
            if r["numVertices"]:
                assert np.all(got[:, 3] == np.float32(0.05)) and np.abs(got[:, :3]).max() <= 256 * 0.05 * 1.5
        for name, mn in (("lod1_0", (0, y1, 0)), ("lod1_-1_1", (-512, y1, 512))):
            r = res[nodes.index(list(mn) + [512])]
            gv = S.as_vertices(V[r["vertexOffset"]:r["vertexOffset"] + r["numVertices"]])
            gt = T["indices_"][r["triangleOffset"]:r["triangleOffset"] + r["numTriangles"]]
            assert (str(len(gv)), str(len(gt)), sha(gv), sha(gt)) == tuple(rows[name][4:]), name
        # physics vertices only (no MeshVertex arena): same positions and triangles
        P3 = np.zeros((150000, 4), np.float32); T3 = np.zeros((300000, 3), np.int32)
        rc, res3, _ = ctx.generateCollisionBatch(nodes, P3, T3, Sn2)
        assert rc == 0 and P3.tobytes() == P.tobytes() and T3.tobytes() == T2.tobytes() and same(res3, res)
    finally:
        ctx.destroy()


@pytest.mark.gpu
def test_gpu_fallback_paths(lc, cases, golden, seams_of_world, surface_cy, monkeypatch):
    """the scratch a block keeps in shared memory falls back to global slices for meshes / seams too
    large for it; the switches force that path on inputs whose answers are known"""
    import seam_scenarios as SS
    from test_seam import seam_digest
    rows, _ = golden
    monkeypatch.setenv("LVN_SIMP_FORCE_GLOBAL", "1")
    monkeypatch.setenv("LVN_SEAM_FORCE_GLOBAL", "1")
    meshes = [(S.as_vertices(v), t, off) for name, v, t, off, opt in cases]
    opts = [lc.SimplifyOptions.make(**opt) for name, v, t, off, opt in cases]
    rc, out, res = lc.ngMeshSimplifierBatch(meshes, opts)
    assert rc == 0
    for (name, *_), (gv, gt) in zip(cases, out):
        gvv = S.as_vertices(gv)
        assert (str(len(gvv)), str(len(gt)), sha(gvv), sha(gt["indices_"])) == tuple(rows[name][4:]), name
    want = np.load(os.path.join(ROOT, "tests", "golden", "ref_seams.npz"))["mixed_lod01"]
    jobs = SS.build_jobs(SS.mixed_lod01(surface_cy), seams_of_world)
    rc, sm, sres = lc.GenerateClipmapSeamMeshes(64, jobs)
    assert rc == 0
    for (gv, gt), w in zip(sm, want):
        assert tuple(str(x) for x in seam_digest(gv, gt["indices_"])) == tuple(str(x) for x in w)


@pytest.fixture(scope="module")
def seams_of_world(world):
    cache = {}

    def get(mn, size):
        k = (tuple(mn), size)
        if k not in cache:
            r = world.generate_chunk_mesh(list(mn), size)
            world.free_chunk_octree(list(mn), size)
            cache[k] = r["seams"]
        return cache[k]
    return get
