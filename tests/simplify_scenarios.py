"""Input meshes for the mesh-simplifier tests: what ConstructClipmapNodeData
(clipmap.cpp:432-468) hands to ngMeshSimplifier -- a chunk's exported MeshBuffer, the node centre
as worldSpaceOffset and options scaled by the node's leaf size -- plus hand-made meshes that hit
the branches terrain meshes do not.

A case is (name, vertices VERTEX_DTYPE[], triangles int32[n][3], worldSpaceOffset xyz, options dict).
The chunk meshes come from the CPU oracle (itself pinned to the reference's kernels,
tests/test_ref_pin.py), so this module needs no GPU and no /root/reference."""
import numpy as np

VERTEX_DTYPE = np.dtype([("xyz", np.float32, 4), ("normal", np.float32, 4), ("colour", np.float32, 4)])

# MeshSimplificationOptions defaults, ng_mesh_simplify.h:6-28
DEFAULTS = dict(edgeFraction=0.125, maxIterations=10, targetPercentage=0.05, maxError=5.0, maxEdgeSize=2.5, minAngleCosine=0.8)


def clipmap_options(node_size, **over):
    """clipmap.cpp:455-462 with options.h:14-16 (meshMaxError 5, meshMaxEdgeLen 2.5, meshMaxAngle 0.7)"""
    leaf = float(4 * (node_size // 256))
    o = dict(DEFAULTS)
    o.update(maxError=5.0 * leaf, maxEdgeSize=2.5 * leaf, minAngleCosine=0.7)
    o.update(over)
    return o


def _mesh(xyz, normal, tris, material=0.0):
    v = np.zeros(len(xyz), VERTEX_DTYPE)
    v["xyz"][:, :3] = xyz
    v["xyz"][:, 3] = 1.0
    v["normal"][:, :3] = normal
    v["colour"][:, :3] = 1.0
    v["colour"][:, 3] = material
    return v, np.ascontiguousarray(np.asarray(tris, np.int32).reshape(-1, 3))


def grid_plane(n=24, pitch=4.0, bump=0.0, seed=1):
    """an open n x n height-field patch: its rim vertices are boundary (edges seen once)"""
    rng = np.random.default_rng(seed)
    ii, jj = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    h = (bump * rng.standard_normal((n, n))).astype(np.float32)
    xyz = np.stack([ii * pitch, h, jj * pitch], -1).reshape(-1, 3).astype(np.float32)
    gx, gz = np.gradient(h, pitch)
    nr = np.stack([-gx, np.ones_like(h), -gz], -1).reshape(-1, 3)
    nr = (nr / np.linalg.norm(nr, axis=1, keepdims=True)).astype(np.float32)
    idx = lambda i, j: i * n + j
    tris = []
    for i in range(n - 1):
        for j in range(n - 1):
            tris += [[idx(i, j), idx(i, j + 1), idx(i + 1, j + 1)], [idx(i, j), idx(i + 1, j + 1), idx(i + 1, j)]]
    return _mesh(xyz, nr, tris)


def torus(nu=40, nv=20, R=40.0, r=12.0, two_materials=False):
    """a closed surface: no boundary vertex, every edge shared by two triangles"""
    u = np.arange(nu) * (2 * np.pi / nu)
    v = np.arange(nv) * (2 * np.pi / nv)
    uu, vv = np.meshgrid(u, v, indexing="ij")
    xyz = np.stack([(R + r * np.cos(vv)) * np.cos(uu), r * np.sin(vv), (R + r * np.cos(vv)) * np.sin(uu)], -1).reshape(-1, 3)
    nr = np.stack([np.cos(vv) * np.cos(uu), np.sin(vv), np.cos(vv) * np.sin(uu)], -1).reshape(-1, 3)
    idx = lambda i, j: (i % nu) * nv + (j % nv)
    tris = []
    for i in range(nu):
        for j in range(nv):
            tris += [[idx(i, j), idx(i + 1, j), idx(i + 1, j + 1)], [idx(i, j), idx(i + 1, j + 1), idx(i, j + 1)]]
    vtx, t = _mesh(xyz.astype(np.float32), nr.astype(np.float32), tris)
    if two_materials:       # collapses across the material border are refused (ng_mesh_simplify.cpp:228-232)
        vtx["colour"][:, 3] = (np.arange(len(vtx)) // nv >= nu // 2).astype(np.float32) * 3.0
    return vtx, t


def fan(spokes=120, rings=3):
    """a disc whose hub has `spokes` triangles: degree above COLLAPSE_MAX_DEGREE around the hub"""
    xyz = [[0.0, 0.0, 0.0]]
    for k in range(1, rings + 1):
        for s in range(spokes):
            a = 2 * np.pi * s / spokes
            xyz.append([2.0 * k * np.cos(a), 0.0, 2.0 * k * np.sin(a)])
    xyz = np.array(xyz, np.float32)
    nr = np.tile(np.array([[0.0, 1.0, 0.0]], np.float32), (len(xyz), 1))
    ring = lambda k, s: 1 + (k - 1) * spokes + (s % spokes)
    tris = [[0, ring(1, s + 1), ring(1, s)] for s in range(spokes)]
    for k in range(1, rings):
        for s in range(spokes):
            tris += [[ring(k, s), ring(k, s + 1), ring(k + 1, s + 1)], [ring(k, s), ring(k + 1, s + 1), ring(k + 1, s)]]
    return _mesh(xyz, nr, tris)


def chunk_cases(world, cy):
    """(name, min, node size) of the chunk meshes: the surface chunk above the origin (BASELINE
    config 1), a block of the LOD0 ring around it, LOD1 / LOD2 nodes"""
    out = [("origin", (0, cy * 256, 0), 256)]
    for cx in (-2, -1, 1):
        for dy in (-1, 0, 1):
            for cz in (-1, 0, 2):
                out.append((f"ring_{cx}_{dy}_{cz}", (cx * 256, (cy + dy) * 256, cz * 256), 256))
    y1 = (cy * 256 // 512) * 512
    out += [("lod1_0", (0, y1, 0), 512), ("lod1_-1_1", (-512, y1, 512), 512)]
    out += [("lod2_0", (0, (cy * 256 // 1024) * 1024, 0), 1024)]
    return out


def csg_chunk(world, oracle_mod, cy):
    """the origin chunk after an add / subtract script (voxel units, as CSGOperationInfo carries them):
    several materials in one mesh; leaves the world edited, so pass one made for this"""
    mn = [0, cy * 256, 0]
    h = lambda x, z: 900.0 * world.terrain(x, z)
    ops = [oracle_mod.make_csg_op(0, 1, 2, [20.5, h(20.0, 20.0), 20.5], [9, 9, 9]),
           oracle_mod.make_csg_op(1, 0, 201, [44.5, h(44.0, 44.0), 44.5], [8, 6, 7], 0.6),
           oracle_mod.make_csg_op(0, 0, 3, [12.5, h(12.0, 50.0) + 3, 50.5], [6, 10, 6], -0.4)]
    world.generate_chunk_mesh(mn, 256)
    world.apply_csg_operations(ops, mn, 256)
    world.free_chunk_octree(mn, 256)
    r = world.generate_chunk_mesh(mn, 256)
    world.free_chunk_octree(mn, 256)
    return r["vertices"], r["indices"]


def all_cases(world, oracle_mod, cy, csg_world=None):
    cases = []
    for name, mn, size in chunk_cases(world, cy):
        r = world.generate_chunk_mesh(list(mn), size)
        world.free_chunk_octree(list(mn), size)
        if r["numTriangles"] == 0:
            continue
        centre = [mn[0] + size / 2.0, mn[1] + size / 2.0, mn[2] + size / 2.0]
        cases.append((name, r["vertices"], r["indices"], centre, clipmap_options(size)))
    v0, t0, c0 = cases[0][1], cases[0][2], cases[0][3]
    # option variants on the origin chunk: one iteration; every edge sampled (the sample then exceeds
    # a block's worth of draws and revisits edges); a loose target; tight and slack thresholds
    cases.append(("origin_1iter", v0, t0, c0, clipmap_options(256, maxIterations=1)))
    cases.append(("origin_all_edges", v0, t0, c0, clipmap_options(256, edgeFraction=1.0)))
    cases.append(("origin_target50", v0, t0, c0, clipmap_options(256, targetPercentage=0.5, edgeFraction=0.5)))
    cases.append(("origin_tight", v0, t0, c0, clipmap_options(256, maxError=0.5, minAngleCosine=0.95)))
    cases.append(("origin_slack", v0, t0, c0, clipmap_options(256, maxError=500.0, maxEdgeSize=40.0, minAngleCosine=-1.0, maxIterations=25)))
    cases.append(("origin_defaults", v0, t0, c0, dict(DEFAULTS)))
    cases.append(("origin_no_offset", v0, t0, [0.0, 0.0, 0.0], clipmap_options(256)))
    if csg_world is not None:
        v, t = csg_chunk(csg_world, oracle_mod, cy)
        cases.append(("csg_materials", v, t, c0, clipmap_options(256)))
    slack = dict(DEFAULTS, maxError=50.0, maxEdgeSize=20.0, minAngleCosine=0.5)
    v, t = grid_plane(24, 4.0, 0.0)
    cases.append(("plane_flat", v, t, [46.0, 0.0, 46.0], slack))
    v, t = grid_plane(40, 4.0, 0.6)
    cases.append(("plane_bumpy", v, t, [78.0, 0.0, 78.0], dict(slack, edgeFraction=0.5)))
    v, t = torus()
    cases.append(("torus", v, t, [0.0, 0.0, 0.0], dict(slack, edgeFraction=0.5, minAngleCosine=0.2)))
    v, t = torus(two_materials=True)
    cases.append(("torus_two_materials", v, t, [0.0, 0.0, 0.0], dict(slack, edgeFraction=1.0, minAngleCosine=0.2, maxIterations=20)))
    v, t = fan()
    cases.append(("fan_high_degree", v, t, [0.0, 0.0, 0.0], dict(slack, edgeFraction=1.0)))
    # MeshBuffer's capacity (render_types.h:63-68: 14 336 vertices, 28 672 triangles), every edge sampled: the
    # largest mesh the reference can hold, the kernel's largest shared-memory footprint
    v, t = grid_plane(119, 4.0, 0.8, seed=7)
    cases.append(("plane_max_size", v, t, [236.0, 0.0, 236.0], dict(slack, edgeFraction=1.0, minAngleCosine=0.3)))
    v, t = grid_plane(7, 4.0, 0.0)          # 72 triangles: under the 100-triangle floor, returned untouched
    cases.append(("too_small", v, t, [12.0, 0.0, 12.0], slack))
    v, t = grid_plane(8, 4.0, 0.0)          # 98 triangles / 64 vertices
    cases.append(("too_small_98", v, t, [14.0, 0.0, 14.0], slack))
    v, t = grid_plane(9, 4.0, 0.0)          # 128 triangles but 81 vertices: the vertex floor
    cases.append(("too_few_vertices", v, t, [16.0, 0.0, 16.0], slack))
    return cases


def as_vertices(v):
    out = np.zeros(len(v), VERTEX_DTYPE)
    for f in ("xyz", "normal", "colour"):
        out[f] = v[f]
    return out
