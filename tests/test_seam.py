"""Seam meshes between clipmap nodes (the consumer of the path's SeamNodeInfo).

Checker: the reference's own leven/src/octree.cpp (Octree_ConstructUpwards, GenerateVertexIndices,
ContourCellProc / FaceProc / EdgeProc / ProcessEdge), compiled unmodified for the host into
oracle/_ref/libleven_octree_ref.so, behind the selection logic of clipmap.cpp restated in
oracle/ref.py.  Where the library is absent the committed digests (tests/golden/ref_seams.npz,
generated from it by tests/golden/gen_ref_vectors.py) stand in.

Bar: vertex arrays bit-identical, in the reference's order; triangles the same multiset of index
triples with the same winding (the CUDA path orders them by owner vertex and edge, the
reference by its recursion)."""
import hashlib
import os

import numpy as np
import pytest

import seam_scenarios as S
from conftest import ROOT, SEED

GOLDEN = os.path.join(ROOT, "tests", "golden", "ref_seams.npz")
VERTEX_DTYPE = np.dtype([("xyz", np.float32, 4), ("normal", np.float32, 4), ("colour", np.float32, 4)])


def digest(b):
    return hashlib.sha256(b).hexdigest()


def canon_tris(t):
    t = np.asarray(t, np.int32).reshape(-1, 3)
    return t[np.lexsort((t[:, 2], t[:, 1], t[:, 0]))] if len(t) else t


def seam_digest(verts, tris):
    v = np.zeros(len(verts), VERTEX_DTYPE)
    for f in ("xyz", "normal", "colour"):
        v[f] = verts[f]
    return len(v), len(np.asarray(tris).reshape(-1, 3)), digest(v.tobytes()), digest(canon_tris(tris).tobytes())


@pytest.fixture(scope="module")
def seams_of(world):
    cache = {}

    def get(mn, size):
        k = (tuple(mn), size)
        if k not in cache:
            r = world.generate_chunk_mesh(list(mn), size)
            world.free_chunk_octree(list(mn), size)
            cache[k] = r["seams"]
        return cache[k]
    return get


@pytest.fixture(scope="module")
def octree_ref(built):
    from oracle import ref as R
    if not R.octree_available():
        pytest.skip("oracle/_ref/libleven_octree_ref.so not built (needs /root/reference)")
    return R


def test_reference_octree_matches_golden(octree_ref, seams_of, surface_cy):
    """the shipped digests are what the reference code computes (guards the fixture itself)"""
    g = np.load(GOLDEN)
    for name, make in S.SCENARIOS.items():
        jobs = S.build_jobs(make(surface_cy), seams_of)
        want = g[name]
        assert len(want) == len(jobs)
        for (host, size, nbs), w in zip(jobs, want):
            v, t = octree_ref.seam_mesh(host, size, nbs)
            nv, nt, dv, dt = seam_digest(v, t)
            assert (str(nv), str(nt), dv, dt) == tuple(str(x) for x in w), (name, host, size)


def test_selection_and_duplicates(octree_ref, seams_of, surface_cy):
    """a coarser neighbour is listed once per slot it covers and its nodes can pass several slot
    filters: the reference links the same leaf twice (octree.cpp:63-70, assert commented out)"""
    active = S.mixed_lod01(surface_cy)
    seen_dup = seen_two_sizes = 0
    for host in [a for a in active if a[1] == 256]:
        nbs = S.neighbours_for(host[0], host[1], active, seams_of)
        if not any(n[2] == 512 for n in nbs):
            continue
        ms, pos, nrm, mat = octree_ref.select_seam_nodes(list(host[0]), host[1], nbs)
        keys = {tuple(r) for r in ms.tolist()}
        seen_dup += len(keys) < len(ms)
        seen_two_sizes += {int(r[3]) for r in ms} == {4, 8}      # leaves of two sizes in one seam octree
        v, t = octree_ref.seam_mesh(list(host[0]), host[1], nbs)
        assert len(v) in (0, len(keys))                          # each distinct leaf is one vertex
    assert seen_dup > 0 and seen_two_sizes > 0


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(S.SCENARIOS))
def test_gpu_seam_meshes(lc, seams_of, surface_cy, name, built):
    from oracle import ref as R
    jobs = S.build_jobs(S.SCENARIOS[name](surface_cy), seams_of)
    rc, meshes, res = lc.GenerateClipmapSeamMeshes(64, jobs)
    assert rc == 0, lc.lib().lvn_seam_last_error()
    want = np.load(GOLDEN)[name]
    total_t = 0
    for (host, size, nbs), (gv, gt), w, r in zip(jobs, meshes, want, res):
        got = seam_digest(gv, gt["indices_"])
        assert tuple(str(x) for x in got) == tuple(str(x) for x in w), (name, host, size, "vs golden")
        if R.octree_available():
            rv, rt = R.seam_mesh(host, size, nbs)
            assert gv["xyz"].tobytes() == rv["xyz"].tobytes() and gv["normal"].tobytes() == rv["normal"].tobytes()
            assert gv["colour"].tobytes() == rv["colour"].tobytes()
            assert np.array_equal(canon_tris(gt["indices_"]), canon_tris(rt)), (name, host, size)
            assert r["numSelectedNodes"] >= len(rv)
        total_t += len(gt)
    assert total_t > 1000


@pytest.mark.gpu
def test_gpu_seam_edge_cases(lc, seams_of, surface_cy):
    """no seams, a host without neighbours, a host whose neighbours hold no seam node, capacity"""
    assert lc.GenerateClipmapSeamMeshes(64, [])[0] == 0
    rc, meshes, res = lc.GenerateClipmapSeamMeshes(64, [([0, 0, 0], 256, [])])
    assert rc == 0 and res[0]["numVertices"] == 0 and res[0]["numTriangles"] == 0
    air = ([0, 15 * 256, 0], 256, S.neighbours_for((0, 15 * 256, 0), 256, [((0, 15 * 256, 0), 256)], seams_of))
    rc, meshes, res = lc.GenerateClipmapSeamMeshes(64, [air])
    assert rc == 0 and len(meshes[0][0]) == 0
    # a lone surface chunk: every selected node is on a face of the host but no edge has four cells
    mn = (0, surface_cy * 256, 0)
    lone = (list(mn), 256, S.neighbours_for(mn, 256, [(mn, 256)], seams_of))
    rc, meshes, res = lc.GenerateClipmapSeamMeshes(64, [lone])
    assert rc == 0 and res[0]["numSelectedNodes"] > 0 and res[0]["numTriangles"] == 0 and res[0]["numVertices"] == 0
    # invalid arguments
    import leven_b200.compute as c
    bad = np.zeros(1, c.SeamJob); bad[0]["hostSize"] = 300; bad[0]["numNeighbours"] = 0
    r = np.zeros(1, c.SeamResult)
    assert c.lib().lvn_seam_mesh_generate_batch(64, 1, c._ptr(bad), None, 0, None, 0, None, 0, None, 0, c._ptr(r)) == c.LVN_ERR_INVALID_VALUE


@pytest.mark.gpu
def test_gpu_seam_internal_triangle_scratch_retries(lc, seams_of, surface_cy, monkeypatch):
    """the triangle scratch of a seam is sized for 8 triangles per candidate node; a seam that needs more makes
    the call run again at the true bound of 24 instead of failing with a capacity error the caller cannot
    cure.  LVN_SEAM_TRIANGLES_PER_CANDIDATE=0 makes every seam with a triangle take that path."""
    jobs = S.build_jobs(S.SCENARIOS["uniform_lod0"](surface_cy), seams_of)
    rc0, meshes0, res0 = lc.GenerateClipmapSeamMeshes(64, jobs)
    monkeypatch.setenv("LVN_SEAM_TRIANGLES_PER_CANDIDATE", "0")
    rc1, meshes1, res1 = lc.GenerateClipmapSeamMeshes(64, jobs)
    assert rc0 == 0 and rc1 == 0, lc.lib().lvn_seam_last_error()
    assert sum(len(t) for _, t in meshes0) > 1000
    for (v0, t0), (v1, t1) in zip(meshes0, meshes1):
        assert v0.tobytes() == v1.tobytes() and t0.tobytes() == t1.tobytes()
