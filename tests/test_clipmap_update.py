"""One Clipmap::update as two batched passes: lvn_clipmap_update_batch =
ConstructClipmapNodeData for the nodes to load + the seam-update set + GenerateClipmapSeamMesh
for that set (clipmap.cpp:1253-1340).

Checked against: the separately tested pieces called one by one (generateSimplifiedBatch, the
seam batch fed by tests/seam_scenarios.py's restatement of findNode + findActiveNodes), the
committed digests of the reference's octree.cpp for the seams (tests/golden/ref_seams.npz) and a
brute-force statement of the seam-update rule."""
import os

import numpy as np
import pytest

import seam_scenarios as S
from conftest import ROOT
from test_seam import seam_digest

GOLDEN = os.path.join(ROOT, "tests", "golden", "ref_seams.npz")


def brute_force_updates(active, constructed):
    """clipmap.cpp:1306-1324 on a flat list: active nodes that contain the min of one of the 8 cells
    min - CHILD_MIN_OFFSETS[i] * size of a constructed node, or whose min lies inside such a cell"""
    out = set()
    for (cmn, csz) in constructed:
        for off in S.CHILD_MIN_OFFSETS:
            cell = [cmn[k] - off[k] * csz for k in range(3)]
            for (amn, asz) in active:
                if all(amn[k] <= cell[k] < amn[k] + asz for k in range(3)) or all(cell[k] <= amn[k] < cell[k] + csz for k in range(3)):
                    out.add((tuple(amn), asz))
    return out


def same_node_meshes(res_a, Va, Ta, res_b, Vb, Tb):
    """the two calls hold the same mesh for every node, each at its own offsets; both pack densely"""
    for a, b in zip(res_a, res_b):
        assert Va[a["vertexOffset"]:a["vertexOffset"] + a["numVertices"]].tobytes() == Vb[b["vertexOffset"]:b["vertexOffset"] + b["numVertices"]].tobytes()
        assert Ta[a["triangleOffset"]:a["triangleOffset"] + a["numTriangles"]].tobytes() == Tb[b["triangleOffset"]:b["triangleOffset"] + b["numTriangles"]].tobytes()
    for res in (res_a, res_b):
        nz = sorted((r for r in res if r["numTriangles"]), key=lambda r: int(r["vertexOffset"]))
        assert nz[0]["vertexOffset"] == 0 and nz[0]["triangleOffset"] == 0
        for x, y in zip(nz, nz[1:]):
            assert y["vertexOffset"] == x["vertexOffset"] + x["numVertices"] and y["triangleOffset"] == x["triangleOffset"] + x["numTriangles"]


def run_update(lc, ctx, nodes, num_active, seam_arena, used, V, T):
    rc, cres, upd, sres, tot = lc.ClipmapUpdateBatch(ctx, nodes, num_active, seam_arena, used, V, T)
    assert rc == 0, (rc, lc.last_cuda_error(), lc.lib().lvn_seam_last_error())
    return cres, upd, sres, tot


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["uniform_lod0", "mixed_lod01"])
def test_gpu_update_constructs_whole_cover(lc, surface_cy, name, built):
    cover = S.SCENARIOS[name](surface_cy)
    want = np.load(GOLDEN)[name]
    ctx = lc.Compute_MeshGenContext.create(64)
    try:
        nodes = np.zeros(len(cover), lc.ClipmapNode)
        for k, (mn, size) in enumerate(cover):
            nodes[k]["min"] = mn; nodes[k]["size"] = size
        V = np.zeros(400000, lc.MeshVertex); T = np.zeros(800000, lc.MeshTriangle); Sn = np.zeros(200000, lc.SeamNodeInfo)
        cres, upd, sres, tot = run_update(lc, ctx, nodes, 0, Sn, 0, V, T)
        # pass 1 = the fused construct call
        V1 = np.zeros(400000, lc.MeshVertex); T1 = np.zeros(800000, lc.MeshTriangle); S1 = np.zeros(200000, lc.SeamNodeInfo)
        rc, res1, _ = ctx.generateSimplifiedBatch([list(mn) + [size] for mn, size in cover], V1, T1, S1)
        assert rc == 0
        for f in ("numVertices", "numTriangles", "numSeamNodes"):
            assert np.array_equal(cres[f], res1[f]), f
        assert tot.nodeVertices == res1["numVertices"].sum() and tot.nodeTriangles == res1["numTriangles"].sum()
        # node by node (the fused call on its own packs the largest meshes last; the update does not)
        same_node_meshes(cres, V, T, res1, V1, T1)
        assert tot.seamNodesUsed == res1["numSeamNodes"].sum()
        for n, r, r1 in zip(nodes, cres, res1):
            assert n["firstSeamNode"] == r["seamOffset"] and n["numSeamNodes"] == r["numSeamNodes"]
            assert Sn[r["seamOffset"]:r["seamOffset"] + r["numSeamNodes"]].tobytes() == S1[r1["seamOffset"]:r1["seamOffset"] + r1["numSeamNodes"]].tobytes()
        # every node that became active is in its own update set (cell 0), nothing else is
        active = [k for k, r in enumerate(cres) if r["numTriangles"] > 0 or r["numSeamNodes"] > 0]
        assert upd.tolist() == active and tot.numConstructedActive == len(active) == tot.numSeamUpdates
        # pass 2 = the reference's seam meshes
        seen = 0
        for k, sr in zip(upd, sres):
            gv = V[sr["vertexOffset"]:sr["vertexOffset"] + sr["numVertices"]]
            gt = T["indices_"][sr["triangleOffset"]:sr["triangleOffset"] + sr["numTriangles"]]
            assert sr["vertexOffset"] >= tot.nodeVertices and sr["triangleOffset"] >= tot.nodeTriangles
            assert tuple(str(x) for x in seam_digest(gv, gt)) == tuple(str(x) for x in want[k]), (name, cover[k])
            seen += len(gt)
        assert seen > 1000 and tot.seamTriangles == seen
    finally:
        ctx.destroy()


@pytest.mark.gpu
def test_gpu_update_incremental(lc, surface_cy, built):
    """two updates: the LOD1 shell first, then the 8 LOD0 children of the centre; the second one
    regenerates exactly the seams the rule names, and they equal the reference's seams of the full cover"""
    cover = S.mixed_lod01(surface_cy)
    want = np.load(GOLDEN)["mixed_lod01"]
    first = [c for c in cover if c[1] == 512]
    second = [c for c in cover if c[1] == 256]
    assert len(second) == 8
    ctx = lc.Compute_MeshGenContext.create(64)
    try:
        V = np.zeros(400000, lc.MeshVertex); T = np.zeros(800000, lc.MeshTriangle); Sn = np.zeros(200000, lc.SeamNodeInfo)
        nodes = np.zeros(len(first), lc.ClipmapNode)
        for k, (mn, size) in enumerate(first):
            nodes[k]["min"] = mn; nodes[k]["size"] = size
        cres, upd, sres, tot = run_update(lc, ctx, nodes, 0, Sn, 0, V, T)
        keep = nodes[[k for k, r in enumerate(cres) if r["numTriangles"] > 0 or r["numSeamNodes"] > 0]]
        assert 0 < len(keep) < len(first)
        # the application keeps the active nodes and the arena; the next update appends
        nodes2 = np.zeros(len(keep) + 8, lc.ClipmapNode)
        nodes2[:len(keep)] = keep
        for k, (mn, size) in enumerate(second):
            nodes2[len(keep) + k]["min"] = mn; nodes2[len(keep) + k]["size"] = size
        used = int(tot.seamNodesUsed)
        before = Sn[:used].copy()
        cres2, upd2, sres2, tot2 = run_update(lc, ctx, nodes2, len(keep), Sn, used, V, T)
        assert Sn[:used].tobytes() == before.tobytes() and tot2.seamNodesUsed == used + cres2["numSeamNodes"].sum()
        assert all(r["seamOffset"] >= used for r in cres2 if r["numSeamNodes"])
        new_active = [k for k, r in enumerate(cres2) if r["numTriangles"] > 0 or r["numSeamNodes"] > 0]
        as_key = lambda n: (tuple(int(x) for x in n["min"]), int(n["size"]))
        active_all = [as_key(n) for n in keep] + [as_key(nodes2[len(keep) + k]) for k in new_active]
        rule = brute_force_updates(active_all, [as_key(nodes2[len(keep) + k]) for k in new_active])
        assert {as_key(nodes2[k]) for k in upd2} == rule and len(upd2) == len(rule)
        assert len(rule) < len(active_all)            # far nodes are not touched
        assert any(nodes2[k]["size"] == 512 for k in upd2) and any(nodes2[k]["size"] == 256 for k in upd2)
        order = {(tuple(mn), size): k for k, (mn, size) in enumerate(cover)}
        for k, sr in zip(upd2, sres2):
            gv = V[sr["vertexOffset"]:sr["vertexOffset"] + sr["numVertices"]]
            gt = T["indices_"][sr["triangleOffset"]:sr["triangleOffset"] + sr["numTriangles"]]
            assert tuple(str(x) for x in seam_digest(gv, gt)) == tuple(str(x) for x in want[order[as_key(nodes2[k])]]), as_key(nodes2[k])
        # nothing to construct: nothing to update
        cres3, upd3, sres3, tot3 = run_update(lc, ctx, nodes2[:len(keep)].copy(), len(keep), Sn, used, V, T)
        assert len(upd3) == 0 and tot3.numSeamUpdates == 0
        # invalid node (not aligned to its size), arenas too small
        bad = nodes2.copy(); bad[-1]["min"] = [100, 0, 0]
        assert lc.ClipmapUpdateBatch(ctx, bad, len(keep), Sn, used, V, T)[0] == lc.LVN_ERR_INVALID_VALUE
        rc = lc.ClipmapUpdateBatch(ctx, nodes2.copy(), len(keep), Sn, used, V[:10], T)[0]
        assert rc == lc.LVN_ERR_CAPACITY
    finally:
        ctx.destroy()


@pytest.mark.gpu
def test_gpu_sharded_update_pieces(lc, surface_cy, built):
    """the pieces the multi-GPU update is made of (leven_b200/sharding.py), on one GPU: pass 2 taken
    in two shares gives, together, exactly the seams of the one-call update; the single-process form
    of sharded_clipmap_update equals lvn_clipmap_update_batch; the seam nodes may sit in device memory"""
    import torch
    from leven_b200 import sharding
    cover = S.mixed_lod01(surface_cy)
    ms = np.array([list(mn) + [size] for mn, size in cover], np.int32)
    ctx = lc.Compute_MeshGenContext.create(64)
    try:
        nodes = np.zeros(len(cover), lc.ClipmapNode)
        nodes["min"] = ms[:, :3]; nodes["size"] = ms[:, 3]
        V = np.zeros(400000, lc.MeshVertex); T = np.zeros(800000, lc.MeshTriangle); Sn = np.zeros(200000, lc.SeamNodeInfo)
        cres, upd, sres, tot = run_update(lc, ctx, nodes, 0, Sn, 0, V, T)
        whole = {int(k): seam_digest(V[r["vertexOffset"]:r["vertexOffset"] + r["numVertices"]],
                                     T["indices_"][r["triangleOffset"]:r["triangleOffset"] + r["numTriangles"]]) for k, r in zip(upd, sres)}
        active = np.array([k for k, r in enumerate(cres) if r["numTriangles"] > 0 or r["numSeamNodes"] > 0], np.int32)
        dev = torch.from_numpy(Sn[:tot.seamNodesUsed].view(np.uint8).reshape(-1).copy()).cuda()      # shard 1 reads device memory
        got = {}
        for shard, arena in ((0, Sn), (1, int(dev.data_ptr()))):
            V2 = np.zeros(100000, lc.MeshVertex); T2 = np.zeros(100000, lc.MeshTriangle)
            rc, u2, s2, n_all = lc.ClipmapSeamUpdateBatch(64, nodes, active, active, arena, tot.seamNodesUsed, V2, T2, shard, 2)
            assert rc == 0 and n_all == len(upd) and len(u2) in (len(upd) // 2, (len(upd) + 1) // 2)
            assert u2.tolist() == upd[shard::2].tolist()
            for k, r in zip(u2, s2):
                got[int(k)] = seam_digest(V2[r["vertexOffset"]:r["vertexOffset"] + r["numVertices"]],
                                          T2["indices_"][r["triangleOffset"]:r["triangleOffset"] + r["numTriangles"]])
        assert got == whole and len(got) > 10
        # world size 1: the same answer through the sharded driver
        V3 = np.zeros(400000, lc.MeshVertex); T3 = np.zeros(800000, lc.MeshTriangle); S3 = np.zeros(200000, lc.SeamNodeInfo)
        out = sharding.sharded_clipmap_update(lc, ctx, ms, V3, T3, S3)
        assert out["seam_update_nodes"].tolist() == upd.tolist() and out["num_seam_updates_all"] == len(upd)
        assert out["node_totals"] == (tot.nodeVertices, tot.nodeTriangles)
        same_node_meshes(cres, V, T, out["results"], V3, T3)
        for k, r in zip(out["seam_update_nodes"], out["seam_results"]):
            assert seam_digest(V3[r["vertexOffset"]:r["vertexOffset"] + r["numVertices"]],
                               T3["indices_"][r["triangleOffset"]:r["triangleOffset"] + r["numTriangles"]]) == whole[int(k)]
    finally:
        ctx.destroy()


def _random_cover(rng, depth=2):
    """an octree cover of a 2048^3 cube: cells split at random down to 256 (aligned, non-overlapping)"""
    cells = [((0, 0, 0), 2048)]
    out = []
    while cells:
        mn, size = cells.pop()
        if size > 256 and (size > 1024 or rng.random() < 0.55):
            h = size // 2
            cells += [((mn[0] + dx * h, mn[1] + dy * h, mn[2] + dz * h), h) for dx in (0, 1) for dy in (0, 1) for dz in (0, 1)]
        else:
            out.append((mn, size))
    return out


def test_seam_update_set_host_logic_random_covers(built):
    """the seam-update set (clipmap.cpp:1306-1324) is host logic: on random mixed-LOD covers with a
    random part of the nodes newly constructed it equals the brute-force statement of the rule, is
    ascending, and its shares partition it (no device needed: the set is reported before the seam
    launch is attempted)"""
    import torch
    import leven_b200.compute as lc          # (not the `lc` fixture: that one needs a device)
    rng = np.random.default_rng(31)
    for trial in range(12):
        cover = [c for c in _random_cover(rng) if rng.random() < 0.8]        # some cells stay empty / inactive
        nodes = np.zeros(len(cover), lc.ClipmapNode)
        for k, (mn, size) in enumerate(cover):
            nodes[k]["min"] = mn; nodes[k]["size"] = size
        nodes["numSeamNodes"] = 1
        active = np.arange(len(cover), dtype=np.int32)
        constructed = np.sort(rng.choice(len(cover), size=max(1, len(cover) // 5), replace=False)).astype(np.int32)
        want = brute_force_updates(cover, [cover[i] for i in constructed])
        want_idx = sorted(k for k, c in enumerate(cover) if (tuple(c[0]), c[1]) in want)
        V = np.zeros(8, lc.MeshVertex); T = np.zeros(8, lc.MeshTriangle); arena = np.zeros(8, lc.SeamNodeInfo)
        got = []
        for shard in range(3):
            rc, upd, sres, n_all = lc.ClipmapSeamUpdateBatch(64, nodes, active, constructed, arena, 8, V, T, shard, 3)
            if not torch.cuda.is_available():
                assert rc == lc.LVN_ERR_NO_DEVICE or len(upd) == 0
            assert n_all == len(want_idx), (trial, n_all, len(want_idx))
            assert upd.tolist() == want_idx[shard::3]
            got += upd.tolist()
        assert sorted(got) == want_idx and len(want_idx) >= len(constructed)
