"""The reference's OWN caller through the drop-in boundary (missing item 3).

oracle/_ref/libleven_clipmap_caller.so is leven/src/clipmap.cpp:329-504 -- GenerateMeshDataForNode,
ConstructClipmapNodeData, ConstructCollisionNodeData, verbatim -- compiled against
include/leven_compute.hpp (in place of compute.h / ng_mesh_simplify.h) and linked to
leven_b200/lib/libleven_b200.so (oracle/ref_shim/ref_clipmap_caller.cpp, built by `make -C oracle ref`
where /root/reference exists; the library travels to the GPU box).  Here that code requests a handful of
nodes, and what it ends up holding -- the simplified render mesh, the seam OctreeNodes, ClipmapNode::active_
-- is compared with the direct C-ABI calls and with the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def nodes_for(surface_cy):
    cy = surface_cy
    return [([0, cy * 256, 0], 256), ([256, cy * 256, -256], 256), ([0, 15 * 256, 0], 256),      # LOD0: surface, surface, air
            ([0, (cy // 2) * 512, 0], 512), ([-1024, (cy // 4) * 1024, 0], 1024)]                # LOD1, LOD2


def test_reference_caller_equals_direct_abi(lc, oracle_mod, surface_cy):
    from oracle import ref as R
    if not R.caller_available():
        pytest.skip("oracle/_ref/libleven_clipmap_caller.so not built (needs /root/reference at build time)")
    ctx = lc.Compute_MeshGenContext.create(64)
    world = oracle_mod.World(image=lc.Compute_GetNoiseImage(), default_material=0, voxels_per_chunk=64)
    V = np.zeros(200000, lc.MeshVertex); T = np.zeros(400000, lc.MeshTriangle); S = np.zeros(100000, lc.SeamNodeInfo)
    try:
        some_mesh = some_seams = 0
        for mn, size in nodes_for(surface_cy):
            got = R.caller_construct_node(mn, size)                      # the reference's ConstructClipmapNodeData
            ms = np.array([mn + [size]], np.int32)
            rc, res, simp = ctx.generateSimplifiedBatch(ms, V, T, S)     # the fused C-ABI call: same work in one pass
            assert rc == 0
            r = res[0]
            # the render mesh the reference keeps: generateChunkMesh + ngMeshSimplifier through the shim
            assert len(got["vertices"]) == r["numVertices"] and len(got["triangles"]) == r["numTriangles"]
            v = V[r["vertexOffset"]:r["vertexOffset"] + r["numVertices"]]
            assert got["vertices"].tobytes() == v.tobytes(), f"render mesh vertices {mn} {size}"
            assert np.array_equal(got["triangles"], T["indices_"][r["triangleOffset"]:r["triangleOffset"] + r["numTriangles"]])
            # the seam OctreeNodes it builds from SeamNodeInfo (clipmap.cpp:398-415) against the oracle's seam nodes
            ref = world.generate_chunk_mesh(mn, size)
            world.free_chunk_octree(mn, size)
            assert len(got["seamMinSize"]) == ref["numSeamNodes"] == r["numSeamNodes"]
            if ref["numSeamNodes"]:
                unit = size // 64
                lm = ref["seams"]["localspaceMin"]
                assert np.array_equal(got["seamMinSize"][:, :3], lm[:, :3] * unit + np.array(mn, np.int32))
                assert np.all(got["seamMinSize"][:, 3] == unit)
                assert np.array_equal(got["seamMaterial"], lm[:, 3])
                assert got["seamPosition"].tobytes() == np.ascontiguousarray(ref["seams"]["position"][:, :3]).tobytes()
                assert got["seamNormal"].tobytes() == np.ascontiguousarray(ref["seams"]["normal"][:, :3]).tobytes()
                # ColourForMinLeafSize(clipmapNodeSize): the reference passes the node size, not size / 64
                assert np.allclose(got["seamColour"], R.colour_for_min_leaf_size(size))
            assert got["active"] == (r["numTriangles"] > 0 or r["numSeamNodes"] > 0)
            some_mesh += int(r["numTriangles"] > 0); some_seams += int(r["numSeamNodes"] > 0)
        assert some_mesh >= 3 and some_seams >= 3
    finally:
        ctx.destroy(); world.close()


def test_reference_collision_caller(lc, oracle_mod, surface_cy):
    """ConstructCollisionNodeData (clipmap.cpp:474-504): COLLISION_NODE_SIZE = 512 on the 64-voxel physics context"""
    from oracle import ref as R
    if not R.caller_available():
        pytest.skip("oracle/_ref/libleven_clipmap_caller.so not built")
    ctx = lc.Compute_MeshGenContext.create(64)
    V = np.zeros(200000, lc.MeshVertex); T = np.zeros(400000, lc.MeshTriangle); S = np.zeros(100000, lc.SeamNodeInfo)
    try:
        mn, size = [0, (surface_cy // 2) * 512, 0], 512
        got = R.caller_construct_node(mn, size, collision=True)
        rc, res, simp = ctx.generateSimplifiedBatch(np.array([mn + [size]], np.int32), V, T, S)
        assert rc == 0 and res[0]["numTriangles"] > 0
        r = res[0]
        assert got["vertices"].tobytes() == V[r["vertexOffset"]:r["vertexOffset"] + r["numVertices"]].tobytes()
        assert np.array_equal(got["triangles"], T["indices_"][r["triangleOffset"]:r["triangleOffset"] + r["numTriangles"]])
        assert len(got["seamMinSize"]) == r["numSeamNodes"]
    finally:
        ctx.destroy()
