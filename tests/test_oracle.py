"""CPU tests: pin the oracle against the reference's own fixtures and cross-check the
unpinned stages with invariants."""
import os

import numpy as np
import pytest

from conftest import ROOT, SEED


# ---- a9: cuckoo, pinned by leven/src/test_cuckoo.cpp:107-178 ---------------------------------
def test_cpu_cuckoo_octree_keys(oracle_mod, golden_keys):
    """CuckooHashTable (Octree data): every key inserts with value 42, every key is found, all 42."""
    import ctypes as C
    L = oracle_mod.lib()
    for name, keys in golden_keys.items():
        if name.endswith("duplicated"):
            continue
        # the reference sizes the table with sizeof(array)/sizeof(uint32_t), i.e. keys + terminator
        t = L.lvo_cpu_cuckoo_create(len(keys) + 1, 0x4be3f)
        inserted = sum(L.lvo_cpu_cuckoo_insert(t, int(k), 42) for k in keys)
        assert inserted == len(keys), name
        v = C.c_uint32(0)
        for k in keys:
            assert L.lvo_cpu_cuckoo_find(t, int(k), C.byref(v)) == 1
            assert v.value == 42
        L.lvo_cpu_cuckoo_destroy(t)


def test_cpu_cuckoo_insert_ratios(oracle_mod):
    """CuckooHashTable: leven/src/test_cuckoo.cpp:5-105, the nine (keys, table size, min ratio) cases
    (the three 2^20-key cases run with 2^17 keys to keep the CPU suite short)."""
    import ctypes as C
    L = oracle_mod.lib()
    rng = np.random.RandomState(1)
    small, medium, large = 1 << 5, 1 << 14, 1 << 17
    keys = np.unique(rng.randint(0, 2 ** 32 - 2, size=large * 2, dtype=np.uint64).astype(np.uint32))[:large]
    rng.shuffle(keys)
    cases = [(small, small, .98), (medium, medium, .98), (large, large, .98),
             (small, small, .98), (small, medium, .99), (small, large, 1.0),
             (medium, small, .01), (medium, medium, .98), (medium, large, .99)]
    for nkeys, tsize, ratio in cases:
        t = L.lvo_cpu_cuckoo_create(tsize, 0x4be3f)
        ins = [L.lvo_cpu_cuckoo_insert(t, int(k), 0) for k in keys[:nkeys]]
        assert sum(ins) / nkeys >= ratio
        if ratio >= .98:
            v = C.c_uint32(1)
            found = sum(L.lvo_cpu_cuckoo_find(t, int(k), C.byref(v)) for k in keys[:nkeys])
            assert found == sum(ins)
        L.lvo_cpu_cuckoo_destroy(t)


def test_kernel_cuckoo_restatement(oracle_mod, golden_keys):
    """Compute (Cuckoo), leven/src/test_compute.cpp:70-88 (100 keys insert), plus the octree key
    sets through the kernel's hash (32-bit a*key wrap, cuckoo.cl:21): every key is found with its
    index -- the only observable property of the table."""
    import ctypes as C
    L = oracle_mod.lib()
    sets = [np.arange(100, dtype=np.uint32)] + [v for k, v in sorted(golden_keys.items()) if not k.endswith("duplicated")]
    for keys in sets:
        keys = np.ascontiguousarray(keys, np.uint32)
        t = oracle_mod.Cuckoo()
        L.lvo_cuckoo_init(C.byref(t), len(keys))
        assert t.prime == L.lvo_find_next_prime(max(2048, 2 * len(keys)))
        assert L.lvo_cuckoo_insert_keys(C.byref(t), keys.ctypes.data, len(keys)) == 0
        for i in range(0, len(keys), 7):
            assert L.lvo_cuckoo_find(C.byref(t), int(keys[i])) == i
        assert L.lvo_cuckoo_find(C.byref(t), 0xfffffff0) == 0xffffffff
        L.lvo_cuckoo_free(C.byref(t))


def test_find_next_prime(oracle_mod):
    L = oracle_mod.lib()
    def is_prime(n):
        return n > 1 and all(n % p for p in range(2, int(n ** .5) + 1))
    for n in [0, 1, 2, 3, 4, 5, 6, 7, 24, 25, 2048, 4096, 8418, 23014, 100003, 1647750]:
        p = L.lvo_find_next_prime(n)
        assert p >= n and is_prime(p)
        assert not any(is_prime(q) for q in range(max(n, 2), p))


# ---- a15: dedupe / scan / compact, pinned by leven/src/test_compute.cpp:46-68 ---------------
def test_remove_duplicates_fixture(oracle_mod, golden_keys):
    L = oracle_mod.lib()
    dup = np.ascontiguousarray(golden_keys["keys_3_duplicated"].view(np.int32))
    out = np.zeros_like(dup)
    n = L.lvo_remove_duplicates(dup.ctypes.data, len(dup), out.ctypes.data)
    uniq = golden_keys["keys_3"]
    assert n == len(uniq)
    assert np.array_equal(np.sort(out[:n].view(np.uint32)), np.sort(uniq))


def test_scan_compact(oracle_mod):
    L = oracle_mod.lib()
    rng = np.random.RandomState(0)
    for n in [1, 2, 255, 256, 257, 823875]:
        data = rng.randint(0, 2, size=n).astype(np.int32)
        scan = np.zeros_like(data)
        total = L.lvo_exclusive_scan(data.ctypes.data, scan.ctypes.data, n)
        assert total == data.sum()
        assert np.array_equal(scan, np.cumsum(data) - data)
        vals = rng.randint(0, 1 << 30, size=n).astype(np.int32)
        out = np.zeros_like(vals)
        m = L.lvo_compact(vals.ctypes.data, data.ctypes.data, n, out.ctypes.data)
        assert np.array_equal(out[:m], vals[data != 0])


# ---- a1: noise table ---------------------------------------------------------------------
def test_noise_image(oracle_mod):
    img = oracle_mod.noise_image(SEED).reshape(256, 256, 4)
    # gradient bytes are grad3*64+64 (compute_density_field.cpp:107-110)
    assert set(np.unique(img[..., :3])) <= {0, 64, 128}
    # NoiseHash is Jenkins one-at-a-time over the little-endian key bytes
    def jenkins(x, y, seed):
        key = (((x << 24) | (y << 16)) ^ seed) & 0xffffffff
        h = 0
        for i in range(4):
            h = (h + ((key >> (8 * i)) & 0xff)) & 0xffffffff
            h = (h + (h << 10)) & 0xffffffff
            h ^= h >> 6
        h = (h + (h << 3)) & 0xffffffff
        h ^= h >> 11
        h = (h + (h << 15)) & 0xffffffff
        return h
    L = oracle_mod.lib()
    for (x, y) in [(0, 0), (1, 2), (255, 255), (128, 7)]:
        assert L.lvo_noise_hash(x, y, SEED) == jenkins(x, y, SEED)
    assert not np.array_equal(img, oracle_mod.noise_image(SEED + 1).reshape(256, 256, 4))


def test_snoise2_independent_restatement(world):
    """second, independent statement of simplex.cl:99-157 in numpy float32 (fma emulated in
    float64, exact for these magnitudes) agrees bit-for-bit with the C oracle"""
    f32 = np.float32
    img = world.image.reshape(256, 256, 4)

    def fma(a, b, c):
        return f32(np.float64(a) * np.float64(b) + np.float64(c))

    def grad(i, j):
        px = img[j & 255, i & 255]
        return [f32(f32(f32(px[k]) / f32(255.0)) * f32(4.0)) - f32(1.0) for k in range(2)]

    def snoise2(px, py):
        F2, G2 = f32(0.366025403784), f32(0.211324865405)
        s = f32(px + py) * F2
        ix, iy = np.floor(f32(px + s)), np.floor(f32(py + s))
        t = f32(ix + iy) * G2
        x0, y0 = f32(px - f32(ix - t)), f32(py - f32(iy - t))
        o1 = (f32(1), f32(0)) if x0 > y0 else (f32(0), f32(1))
        total = f32(0)
        pts = [(int(ix), int(iy), x0, y0),
               (int(ix) + int(o1[0]), int(iy) + int(o1[1]), f32(f32(x0 - o1[0]) + G2), f32(f32(y0 - o1[1]) + G2)),
               (int(ix) + 1, int(iy) + 1, f32(x0 - f32(f32(1) - f32(2) * G2)), f32(y0 - f32(f32(1) - f32(2) * G2)))]
        ns = []
        for (i, j, x, y) in pts:
            g = grad(i, j)
            t0 = f32(f32(0.5) - fma(y, y, f32(x * x)))
            if t0 < 0:
                ns.append(f32(0))
            else:
                t2 = f32(t0 * t0)
                ns.append(f32(f32(t2 * t2) * fma(g[1], y, f32(g[0] * x))))
        return f32(f32(70) * f32(f32(ns[0] + ns[1]) + ns[2]))

    rng = np.random.RandomState(5)
    for _ in range(300):
        x, y = f32(rng.uniform(-200, 200)), f32(rng.uniform(-200, 200))
        a, b = f32(world.snoise2(float(x), float(y))), snoise2(x, y)
        assert a.view(np.uint32) == b.view(np.uint32), (x, y, a, b)


# ---- unpinned stages: invariants on config 1 ---------------------------------
@pytest.fixture(scope="module")
def chunk1(world, surface_cy):
    return world.generate_chunk_mesh([0, surface_cy * 256, 0], 256)


def test_chunk_counts_consistent(world, chunk1):
    c = chunk1
    E, N, T = c["numEdges"], c["numNodes"], c["numTriangles"]
    assert E > 1000 and N > 1000 and T > 1000 and T % 2 == 0
    mats = c["materials"]
    assert set(np.unique(mats)) <= {0, 201}
    # edges == sign changes of the field, recounted with numpy
    f = (mats.reshape(66, 66, 66) != 201)            # [z][y][x]
    h = f[:65, :65, :65]
    cnt = (h != f[:65, :65, 1:66]).sum() + (h != f[:65, 1:66, :65]).sum() + (h != f[1:66, :65, :65]).sum()
    assert cnt == E
    # keys ascending in (x + 65y + 65^2 z)*3 + axis order
    k = c["edgeKeys"]
    idx = k >> 2
    order = ((idx & 127) + 65 * ((idx >> 7) & 127) + 65 * 65 * ((idx >> 14) & 127)) * 3 + (k & 3)
    assert np.all(np.diff(order) > 0)
    # active voxels == cells whose 8 corners differ
    s = f.astype(np.int32)
    corners = sum(s[dz:64 + dz, dy:64 + dy, dx:64 + dx] for dz in (0, 1) for dy in (0, 1) for dx in (0, 1))
    assert ((corners != 0) & (corners != 8)).sum() == N


def test_hermite_invariants(chunk1):
    info = chunk1["edgeInfo"]
    n = info[:, :3]
    assert np.allclose(np.linalg.norm(n, axis=1), 1.0, atol=1e-5)
    t = info[:, 3]
    assert np.all((t >= 0) & (t <= 1)) and np.all(np.abs(t * 16 - np.round(t * 16)) == 0)
    assert np.all(n[:, 1] > 0)          # a heightfield's gradient points up


def test_mesh_invariants(world, chunk1):
    c = chunk1
    N = c["numNodes"]
    idx = c["indices"]
    assert idx.min() >= 0 and idx.max() < N
    # node codes strictly ascending in x + 64y + 4096z order; decode MSB-first triples (x<<2|y<<1|z)
    codes = c["codes"]
    pos = np.zeros((N, 3), np.int64)
    for d in range(6):
        trip = (codes >> (3 * d)) & 7
        pos[:, 0] |= ((trip >> 2) & 1) << d
        pos[:, 1] |= ((trip >> 1) & 1) << d
        pos[:, 2] |= (trip & 1) << d
    lin = pos[:, 0] + 64 * pos[:, 1] + 4096 * pos[:, 2]
    assert np.all(np.diff(lin) > 0) and np.all(codes >> 18 == 1)
    # every quad (two consecutive triangles) joins the 4 voxels around one sign-changing edge:
    # their positions span a 2x2x1 block
    quads = idx.reshape(-1, 6)
    for q in quads[:: max(1, len(quads) // 400)]:
        p = pos[np.unique(q)]
        assert len(p) == 4
        ext = p.max(0) - p.min(0)
        assert sorted(ext.tolist()) == [0, 1, 1]
    # vertices stay near their voxel (QEF unclamped, so allow slack): |v/4 - chunk offset - cell| small
    v = c["positions"][:, :3] / 4.0
    cell = pos + np.array([0, c["positions"][0, 1] // 256 * 64, 0])
    rel = v - np.array([0.0, np.floor(c["positions"][0, 1] / 256.0) * 64.0, 0.0]) - pos
    assert np.percentile(np.abs(rel - 0.5), 99) < 2.0
    # seam nodes == nodes with a coordinate on a face
    seam = ((pos == 0) | (pos == 63)).any(1)
    assert seam.sum() == c["numSeamNodes"]
    assert np.array_equal(c["seams"]["localspaceMin"][:, :3], pos[seam])
    assert np.array_equal(c["seams"]["localspaceMin"][:, 3], c["matWords"][seam])
    # vertex buffer interleave
    assert np.array_equal(c["vertices"]["xyz"], c["positions"])
    assert np.allclose(c["vertices"]["colour"][:, :3], [0.3, 0.1, 0.0])
    assert np.all(c["vertices"]["colour"][:, 3] == (c["matWords"] >> 8))


def test_qef_solution_satisfies_planes(chunk1):
    """the solved vertex minimises the plane distances: residual small for well-conditioned leaves"""
    q = chunk1["qefs"]
    mp = q["masspoint"][:, :3]
    pos_local = (chunk1["positions"][:, :3] - np.array([0, np.floor(chunk1["positions"][0, 1] / 256) * 256, 0])) / 4.0
    assert np.percentile(np.linalg.norm(pos_local - mp, axis=1), 95) < 1.5


def test_octree_cache_and_csg_semantics(oracle_mod, surface_cy):
    """generateChunkMesh is served from the octree cache until freeChunkOctree
    (compute_octree.cpp:154-181,379-387); applyCSGOperations edits the cached field."""
    w = oracle_mod.World(seed=SEED)
    mn = [0, surface_cy * 256, 0]
    a = w.generate_chunk_mesh(mn, 256)
    yc = surface_cy * 64 + 32.5
    op = oracle_mod.make_csg_op(1, 1, 201, [20.5, yc, 20.5], [6, 6, 6])       # subtract a sphere
    op2 = oracle_mod.make_csg_op(0, 0, 3, [44.5, yc, 44.5], [5, 4, 3])        # add a cuboid of material 3
    w.apply_csg_operations([op, op2], mn, 256)
    b = w.generate_chunk_mesh(mn, 256)          # stale: octree not freed
    assert b["numNodes"] == a["numNodes"] and np.array_equal(a["indices"], b["indices"])
    w.free_chunk_octree(mn, 256)
    c = w.generate_chunk_mesh(mn, 256)
    assert c["numNodes"] != a["numNodes"]
    # edge set == sign changes of the edited field
    f = (c["materials"].reshape(66, 66, 66) != 201)
    hgrid = f[:65, :65, :65]
    cnt = (hgrid != f[:65, :65, 1:66]).sum() + (hgrid != f[:65, 1:66, :65]).sum() + (hgrid != f[1:66, :65, :65]).sum()
    assert cnt == c["numEdges"] and len(np.unique(c["edgeKeys"])) == c["numEdges"]
    # replay: a fresh world that stored the op reproduces the same field lazily
    w2 = oracle_mod.World(seed=SEED)
    for o in (op, op2):
        lo, hi = oracle_mod.csg_operation_bounds(o)
        w2.store_csg_operation(o, lo, hi)
    d = w2.generate_chunk_mesh(mn, 256)
    assert np.array_equal(d["materials"], c["materials"]) and d["numNodes"] == c["numNodes"]
    assert np.array_equal(d["indices"], c["indices"])
    w.close(); w2.close()


def test_empty_chunks(world):
    hi = world.generate_chunk_mesh([0, 15 * 256, 0], 256)
    lo = world.generate_chunk_mesh([0, 0, 0], 256)
    for c in (hi, lo):
        assert c["numEdges"] == 0 and c["numNodes"] == 0 and c["numTriangles"] == 0
    assert np.all(hi["materials"] == 201) and np.all(lo["materials"] == 0)
    assert world.is_chunk_empty([0, 15 * 256, 0], 256)


def test_small_chunk_sizes(oracle_mod):
    """V = 16 at sampleScale 4 covers the same world footprint: the generic-V path of the host code"""
    w = oracle_mod.World(seed=SEED, voxels_per_chunk=16)
    cy = int(900 * w.terrain(0.0, 0.0) // 64)
    c = w.generate_chunk_mesh([0, cy * 256, 0], 256)
    assert c["numNodes"] > 50 and c["numTriangles"] > 50
    w.close()
