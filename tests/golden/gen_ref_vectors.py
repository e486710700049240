#!/usr/bin/env python
"""Generate tests/golden/ref_chunks.npz from the REFERENCE's own kernels.

    python tests/golden/gen_ref_vectors.py            (in the build container: needs /root/reference)

The kernels of leven/cl/*.cl are compiled for the host from where they lie (oracle/ref_shim,
`make -C oracle ref`) and driven through the reference's host sequence (oracle/ref.py).  Every
array below is therefore an output of the reference's kernel text, not of our restatement; the
fixture travels to the GPU box, /root/reference does not.

Cases (V = 64, seed 93923590 = leven/default.cfg:8, noise image from the documented generator):
  full arrays   "origin"   the surface chunk above the world origin, LOD0 (BASELINE config 1)
                "csg"      the same chunk after a 4-op CSG script (sphere/cuboid, add/subtract,
                           rotated, one brush on the chunk boundary), edge list sorted by key
  digests only  more LOD0 chunks of the ring, LOD1 / LOD2 chunks (sampleScale 2 and 4), an
                all-air and an all-solid chunk: sha256 of every stage array
The QEF records' two pad floats are uninitialised in the reference (qef.cl:7-14) and are zeroed
before hashing / storing.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

SEED = 93923590
STAGES = ("materials", "edgeKeys", "edgeInfo", "codes", "edgeMasks", "matWords", "qefs", "positions", "normals",
          "vertices", "indices", "seams")
COUNTS = ("numEdges", "numNodes", "numTriangles", "numSeamNodes")


def csg_script(terrain_height):
    """(type, shape, material, rotateY, origin, dimensions): type 0 add / 1 subtract, shape 0 cube / 1 sphere"""
    th = terrain_height
    return [(1, 1, 201, 0.0, [20.5, th(20.0, 20.0), 20.5], [6, 6, 6]),
            (0, 0, 3, 0.6, [44.5, th(44.0, 44.0), 44.5], [5, 4, 3]),
            (0, 1, 2, 0.0, [10.5, th(10.0, 50.0) + 3, 50.5], [7, 7, 7]),
            (1, 0, 201, -1.1, [63.5, th(63.0, 30.0), 30.5], [4, 9, 6])]


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def normalise(res):
    """zero the QEF pad floats; give absent stages an empty array"""
    from oracle import ref as R
    empty = dict(codes=np.zeros(0, np.uint32), edgeMasks=np.zeros(0, np.int32), matWords=np.zeros(0, np.int32),
                 qefs=np.zeros(0, R.QEF_DTYPE), positions=np.zeros((0, 4), np.float32), normals=np.zeros((0, 4), np.float32),
                 vertices=np.zeros(0, R.VERTEX_DTYPE), indices=np.zeros((0, 3), np.int32), seams=np.zeros(0, R.SEAM_DTYPE))
    out = dict(res)
    for k, v in empty.items():
        out.setdefault(k, v)
    q = out["qefs"].copy()
    if len(q):
        q["pad"] = 0
    out["qefs"] = q
    return out


def digest_cases(cy0):
    """(name, min, size)"""
    return [("ring_-4_-1_3", [-4 * 256, (cy0 - 1) * 256, 3 * 256], 256),
            ("ring_-1_0_-1", [-256, cy0 * 256, -256], 256),
            ("ring_2_0_-3", [2 * 256, cy0 * 256, -3 * 256], 256),       # in the ring, above the surface
            ("world_-8_9_-8", [-8 * 256, 9 * 256, -8 * 256], 256),      # the corner of the default world
            ("world_3_6_2", [3 * 256, 6 * 256, 2 * 256], 256),          # the densest chunk sampled
            ("world_3_5_2", [3 * 256, 5 * 256, 2 * 256], 256),          # the surface just clips the top
            ("world_7_4_-3", [7 * 256, 4 * 256, -3 * 256], 256),        # the lowest terrain
            ("world_2_7_-8", [2 * 256, 7 * 256, -8 * 256], 256),
            ("lod1_0", [0, (cy0 * 256 // 512) * 512, 0], 512),
            ("lod1_-2_1", [-1024, (cy0 * 256 // 512) * 512, 512], 512),
            ("lod2_0", [0, (cy0 * 256 // 1024) * 1024, 0], 1024),
            ("air", [0, 15 * 256, 0], 256),
            ("solid", [0, 0, 0], 256)]


def main():
    from oracle import oracle as O, ref as R
    assert R.build(), "the reference sources are needed to generate the vectors"
    image = O.noise_image(SEED)     # the documented generator ("Noise image"); an input of every implementation
    rw = R.RefWorld(image, default_material=0)
    height = lambda x, z: float(-rw.density(np.float32(x), np.float32(0.0), np.float32(z)))
    cy0 = int(height(0.0, 0.0) // 64)
    out = {"image_sha256": np.array(digest(image)), "cy0": np.array(cy0, np.int32), "seed": np.array(SEED, np.int64)}

    mn = [0, cy0 * 256, 0]
    res = normalise(rw.generate_chunk_mesh(mn, 256))
    out["origin/min_size"] = np.array(mn + [256], np.int32)
    for k in COUNTS:
        out[f"origin/{k}"] = np.array(res[k], np.int32)
    out["origin/materials_u8"] = res["materials"].astype(np.uint8)
    for k in STAGES[1:]:
        out[f"origin/{k}"] = res[k]
    print("origin", {k: res[k] for k in COUNTS})

    script = csg_script(height)
    ops = np.zeros(len(script), R.CSG_DTYPE)
    for i, (ty, sh, mat, rot, org, dim) in enumerate(script):
        ops[i] = (ty, sh, mat, rot, list(org) + [0.0], list(dim) + [0.0])
    m, keys, info = rw.apply_csg(mn, 256, ops, res["materials"], res["edgeKeys"], res["edgeInfo"])
    order = np.argsort(keys, kind="stable")
    oc = rw.construct_octree(mn, 256, m, keys, info)
    verts, tris = rw.generate_mesh(256, oc)
    seams = rw.gather_seam_nodes(oc)
    out["csg/ops"] = ops
    out["csg/materials_u8"] = m.astype(np.uint8)
    out["csg/edgeKeys_sorted"] = keys[order]
    out["csg/edgeInfo_sorted"] = info[order]
    out["csg/codes"] = oc["codes"]; out["csg/matWords"] = oc["matWords"]
    out["csg/positions"] = oc["positions"]; out["csg/normals"] = oc["normals"]
    out["csg/indices"] = tris; out["csg/seams"] = seams
    out["csg/vertices_sha256"] = np.array(digest(verts))
    print("csg", len(keys), "edges", oc["numNodes"], "nodes", len(tris), "triangles", len(seams), "seam nodes")

    names = []
    for name, cmn, size in digest_cases(cy0):
        r = normalise(rw.generate_chunk_mesh(cmn, size))
        names.append(name)
        out[f"d/{name}/min_size"] = np.array(list(cmn) + [size], np.int32)
        out[f"d/{name}/counts"] = np.array([r[k] for k in COUNTS], np.int32)
        out[f"d/{name}/sha256"] = np.array([digest(r[k]) for k in STAGES])
        print(name, cmn, size, {k: r[k] for k in COUNTS})
    out["digest_cases"] = np.array(names)
    out["stages"] = np.array(STAGES)
    path = os.path.join(ROOT, "tests", "golden", "ref_chunks.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


def main_seams():
    """tests/golden/ref_seams.npz: per scenario (tests/seam_scenarios.py) and host node, the
    vertex / triangle counts and sha256 digests of the seam mesh the reference's octree.cpp
    produces (oracle/_ref/libleven_octree_ref.so behind oracle/ref.py's selection)."""
    import hashlib
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import seam_scenarios as S
    from oracle import oracle as O, ref as R
    assert R.build() and R.octree_available()
    W = O.World(seed=SEED)
    cy = int(900 * W.terrain(0.0, 0.0) // 64)
    cache = {}

    def seams_of(mn, size):
        k = (tuple(mn), size)
        if k not in cache:
            r = W.generate_chunk_mesh(list(mn), size)
            W.free_chunk_octree(list(mn), size)
            cache[k] = r["seams"]
        return cache[k]

    def canon(t):
        t = np.asarray(t, np.int32).reshape(-1, 3)
        return t[np.lexsort((t[:, 2], t[:, 1], t[:, 0]))] if len(t) else t
    out = {}
    for name, make in S.SCENARIOS.items():
        rows = []
        for host, size, nbs in S.build_jobs(make(cy), seams_of):
            v, t = R.seam_mesh(host, size, nbs)
            rows.append([str(len(v)), str(len(t)), hashlib.sha256(v.tobytes()).hexdigest(), hashlib.sha256(canon(t).tobytes()).hexdigest()])
        out[name] = np.array(rows)
        print(name, len(rows), "seams", sum(int(r[0]) for r in rows), "vertices", sum(int(r[1]) for r in rows), "triangles")
    path = os.path.join(ROOT, "tests", "golden", "ref_seams.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


def main_simplify():
    """tests/golden/ref_simplify.npz: per case of tests/simplify_scenarios.py, the input's and the
    output's vertex / triangle counts and sha256 digests, the output being what the reference's
    ng_mesh_simplify.cpp + qef_simd.h produce (oracle/_ref/libleven_simplify_ref.so; see
    oracle/ref_shim/ref_simplify.cpp for the two platform-defined pieces it has to fix)."""
    import hashlib
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import simplify_scenarios as S
    from oracle import oracle as O, ref as R
    assert R.build() and R.simplify_available()
    W, W2 = O.World(seed=SEED), O.World(seed=SEED)
    cy = int(900 * W.terrain(0.0, 0.0) // 64)
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    names, rows = [], []
    for name, v, t, off, opt in S.all_cases(W, O, cy, W2):
        vin = S.as_vertices(v)
        rv, rt = R.simplify_mesh(vin, t, off, opt)
        names.append(name)
        rows.append([str(len(vin)), str(len(t)), sha(vin), sha(np.asarray(t, np.int32)), str(len(rv)), str(len(rt)), sha(rv), sha(rt)])
        print(name, len(vin), len(t), "->", len(rv), len(rt))
    # the candidate sample itself (libstdc++'s uniform_int_distribution over mt19937(42)) for a few ranges
    out = {"cases": np.array(names), "rows": np.array(rows)}
    for n in (7, 1000, 16944, 86016, (1 << 20) + 3):
        out[f"random_edges/{n}"] = R.random_edges(n, 4096)
    path = os.path.join(ROOT, "tests", "golden", "ref_simplify.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("chunks", "all"):
        main()
    if which in ("seams", "all"):
        main_seams()
    if which in ("simplify", "all"):
        main_simplify()
