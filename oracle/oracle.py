"""ctypes front-end of the CPU oracle (oracle/lvn_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(leven_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liblvn_oracle.so")

MATERIAL_AIR = 201
MATERIAL_NONE = 200
LEAF_SIZE_SCALE = 4


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("lvn_oracle.c", "lvn_oracle.h")]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in src)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", "-s"])
    return _LIB_PATH


class F4(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float), ("w", C.c_float)]


class CSGOp(C.Structure):
    _fields_ = [("type", C.c_int32), ("brushShape", C.c_int32), ("material", C.c_int32),
                ("rotateY", C.c_float), ("origin", C.c_float * 4), ("dimensions", C.c_float * 4)]


class Chunk(C.Structure):
    _fields_ = [("numEdges", C.c_int), ("numNodes", C.c_int), ("numTriangles", C.c_int),
                ("numSeamNodes", C.c_int),
                ("materials", C.c_void_p), ("edgeKeys", C.c_void_p), ("edgeInfo", C.c_void_p),
                ("codes", C.c_void_p), ("edgeMasks", C.c_void_p), ("matWords", C.c_void_p),
                ("qefs", C.c_void_p), ("positions", C.c_void_p), ("normals", C.c_void_p),
                ("vertices", C.c_void_p), ("indices", C.c_void_p), ("seams", C.c_void_p)]


class Cuckoo(C.Structure):
    _fields_ = [("table", C.c_void_p), ("stash", C.c_uint64 * 101), ("prime", C.c_uint32),
                ("params", C.c_uint32 * 10), ("stashUsed", C.c_int), ("insertedKeys", C.c_int),
                ("retries", C.c_int)]


SEAM_DTYPE = np.dtype([("localspaceMin", np.int32, 4), ("position", np.float32, 4),
                       ("normal", np.float32, 4)])
VERTEX_DTYPE = np.dtype([("xyz", np.float32, 4), ("normal", np.float32, 4), ("colour", np.float32, 4)])
QEF_DTYPE = np.dtype([("ATA", np.float32, 6), ("pad", np.float32, 2), ("ATb", np.float32, 4),
                      ("masspoint", np.float32, 4)])

_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        P, I, F = C.c_void_p, C.c_int, C.c_float
        L.lvo_noise_image.argtypes = [I, P]
        L.lvo_noise_hash.argtypes = [I, I, I]; L.lvo_noise_hash.restype = C.c_uint32
        L.lvo_world_create.argtypes = [P, I, I]; L.lvo_world_create.restype = P
        L.lvo_world_destroy.argtypes = [P]
        L.lvo_world_set_density.argtypes = [P, I, F]
        for name, n in (("lvo_snoise2", 2), ("lvo_snoise3", 3), ("lvo_terrain", 2), ("lvo_density", 3)):
            fn = getattr(L, name); fn.argtypes = [P] + [F] * n; fn.restype = F
        L.lvo_generate_field.argtypes = [P, P, I, P]
        L.lvo_find_edges.argtypes = [P, P, P]; L.lvo_find_edges.restype = I
        L.lvo_edge_info.argtypes = [P, P, I, P, I, P]
        L.lvo_find_active_voxels.argtypes = [P, P, P, P, P]; L.lvo_find_active_voxels.restype = I
        L.lvo_create_leaf_nodes.argtypes = [P, I, P, P, I, P, P, I, P, P]; L.lvo_create_leaf_nodes.restype = I
        L.lvo_solve_qefs.argtypes = [P, P, I, P]
        L.lvo_generate_mesh.argtypes = [P, P, P, I, P]; L.lvo_generate_mesh.restype = I
        L.lvo_find_next_prime.argtypes = [I]; L.lvo_find_next_prime.restype = I
        L.lvo_cuckoo_init.argtypes = [P, C.c_uint32]
        L.lvo_cuckoo_insert_keys.argtypes = [P, P, C.c_uint32]; L.lvo_cuckoo_insert_keys.restype = I
        L.lvo_cuckoo_find.argtypes = [P, C.c_uint32]; L.lvo_cuckoo_find.restype = C.c_uint32
        L.lvo_cuckoo_free.argtypes = [P]
        L.lvo_cpu_cuckoo_create.argtypes = [I, C.c_uint32]; L.lvo_cpu_cuckoo_create.restype = P
        L.lvo_cpu_cuckoo_insert.argtypes = [P, C.c_uint32, C.c_uint32]; L.lvo_cpu_cuckoo_insert.restype = I
        L.lvo_cpu_cuckoo_find.argtypes = [P, C.c_uint32, P]; L.lvo_cpu_cuckoo_find.restype = I
        L.lvo_cpu_cuckoo_destroy.argtypes = [P]
        L.lvo_exclusive_scan.argtypes = [P, P, I]; L.lvo_exclusive_scan.restype = I
        L.lvo_compact.argtypes = [P, P, I, P]; L.lvo_compact.restype = I
        L.lvo_remove_duplicates.argtypes = [P, I, P]; L.lvo_remove_duplicates.restype = I
        L.lvo_brush_density.argtypes = [F, F, F, P]; L.lvo_brush_density.restype = F
        L.lvo_store_csg_operation.argtypes = [P, P, P, P]
        L.lvo_clear_csg_operations.argtypes = [P]
        L.lvo_apply_csg_operations.argtypes = [P, P, I, P, I]
        L.lvo_free_chunk_octree.argtypes = [P, P, I]
        L.lvo_is_chunk_empty.argtypes = [P, P, I, P]
        L.lvo_generate_chunk_mesh.argtypes = [P, P, I, P]
        L.lvo_chunk_free.argtypes = [P]
        L.lvo_generate_batch_counts.argtypes = [P, I, P, P]; L.lvo_generate_batch_counts.restype = I
        L.lvo_set_num_threads.argtypes = [I]; L.lvo_set_num_threads.restype = I
        L.lvo_generate_batch_digests.argtypes = [P, I, P, P, P]; L.lvo_generate_batch_digests.restype = I
        L.lvo_fnv1a64.argtypes = [P, C.c_size_t]; L.lvo_fnv1a64.restype = C.c_uint64
        _lib = L
    return _lib


def set_num_threads(n):
    """OpenMP threads of World.batch_counts (torchrun exports OMP_NUM_THREADS=1); returns the count in effect"""
    return lib().lvo_set_num_threads(int(n))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _i3(v):
    return (C.c_int * 3)(int(v[0]), int(v[1]), int(v[2]))


def noise_image(seed):
    img = np.zeros(256 * 256 * 4, np.uint8)
    lib().lvo_noise_image(int(seed), _ptr(img))
    return img


def _copy(ptr, count, dtype):
    if not ptr or count <= 0:
        return np.zeros(0, dtype)
    nbytes = int(count) * np.dtype(dtype).itemsize
    buf = (C.c_uint8 * nbytes).from_address(ptr)
    return np.frombuffer(bytes(buf), dtype=dtype).copy()


def make_csg_op(type_, shape, material, origin, dimensions, rotate_y=0.0):
    op = CSGOp()
    op.type, op.brushShape, op.material, op.rotateY = int(type_), int(shape), int(material), float(rotate_y)
    for i in range(3):
        op.origin[i] = float(origin[i]); op.dimensions[i] = float(dimensions[i])
    op.origin[3] = 0.0; op.dimensions[3] = 0.0
    return op


class World:
    """Oracle twin of Compute_* + Compute_MeshGenContext (leven/src/compute.h:35-72)."""

    def __init__(self, image=None, seed=93923590, default_material=0, voxels_per_chunk=64):
        self.L = lib()
        self.image = noise_image(seed) if image is None else np.ascontiguousarray(image, np.uint8)
        self.V = voxels_per_chunk
        self.H, self.F = self.V + 1, self.V + 2
        self.h = self.L.lvo_world_create(_ptr(self.image), int(default_material), int(voxels_per_chunk))

    def close(self):
        if self.h:
            self.L.lvo_world_destroy(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_density(self, kind, threshold=0.5):
        self.L.lvo_world_set_density(self.h, int(kind), float(threshold))

    def snoise2(self, x, y): return self.L.lvo_snoise2(self.h, x, y)
    def snoise3(self, x, y, z): return self.L.lvo_snoise3(self.h, x, y, z)
    def terrain(self, x, z): return self.L.lvo_terrain(self.h, x, z)
    def density(self, x, y, z): return self.L.lvo_density(self.h, x, y, z)

    # ---- single stages -------------------------------------------------
    def generate_field(self, mn, size):
        out = np.zeros(self.F ** 3, np.int32)
        self.L.lvo_generate_field(self.h, _i3(mn), int(size), _ptr(out))
        return out

    def find_edges(self, materials):
        keys = np.zeros(3 * self.H ** 3, np.int32)
        n = self.L.lvo_find_edges(self.h, _ptr(materials), _ptr(keys))
        return keys[:n].copy()

    def edge_info(self, mn, size, keys):
        out = np.zeros((len(keys), 4), np.float32)
        if len(keys):
            keys = np.ascontiguousarray(keys, np.int32)
            self.L.lvo_edge_info(self.h, _i3(mn), int(size), _ptr(keys), len(keys), _ptr(out))
        return out

    def find_active_voxels(self, materials):
        n3 = self.V ** 3
        codes = np.zeros(n3, np.uint32); masks = np.zeros(n3, np.int32); mats = np.zeros(n3, np.int32)
        n = self.L.lvo_find_active_voxels(self.h, _ptr(materials), _ptr(codes), _ptr(masks), _ptr(mats))
        return codes[:n].copy(), masks[:n].copy(), mats[:n].copy()

    # ---- host API twin --------------------------------------------------
    def generate_chunk_mesh(self, mn, size):
        ch = Chunk()
        rc = self.L.lvo_generate_chunk_mesh(self.h, _i3(mn), int(size), C.byref(ch))
        assert rc == 0
        E, N, T, S = ch.numEdges, ch.numNodes, ch.numTriangles, ch.numSeamNodes
        out = dict(
            numEdges=E, numNodes=N, numTriangles=T, numSeamNodes=S,
            materials=_copy(ch.materials, self.F ** 3, np.int32),
            edgeKeys=_copy(ch.edgeKeys, E, np.int32),
            edgeInfo=_copy(ch.edgeInfo, E * 4, np.float32).reshape(-1, 4),
            codes=_copy(ch.codes, N, np.uint32),
            edgeMasks=_copy(ch.edgeMasks, N, np.int32),
            matWords=_copy(ch.matWords, N, np.int32),
            qefs=_copy(ch.qefs, N, QEF_DTYPE),
            positions=_copy(ch.positions, N * 4, np.float32).reshape(-1, 4),
            normals=_copy(ch.normals, N * 4, np.float32).reshape(-1, 4),
            vertices=_copy(ch.vertices, N, VERTEX_DTYPE),
            indices=_copy(ch.indices, T * 3, np.int32).reshape(-1, 3),
            seams=_copy(ch.seams, S, SEAM_DTYPE),
        )
        self.L.lvo_chunk_free(C.byref(ch))
        return out

    def apply_csg_operations(self, ops, mn, size):
        arr = (CSGOp * len(ops))(*ops)
        return self.L.lvo_apply_csg_operations(self.h, arr, len(ops), _i3(mn), int(size))

    def store_csg_operation(self, op, aabb_min, aabb_max):
        return self.L.lvo_store_csg_operation(self.h, C.byref(op), _i3(aabb_min), _i3(aabb_max))

    def clear_csg_operations(self):
        return self.L.lvo_clear_csg_operations(self.h)

    def free_chunk_octree(self, mn, size):
        return self.L.lvo_free_chunk_octree(self.h, _i3(mn), int(size))

    def is_chunk_empty(self, mn, size):
        e = C.c_int(0)
        self.L.lvo_is_chunk_empty(self.h, _i3(mn), int(size), C.byref(e))
        return bool(e.value)

    def batch_digests(self, min_size):
        """counts (E, N, T, S) and FNV-1a digests (vertices, indices, seam nodes) per chunk, OpenMP over chunks"""
        ms = np.ascontiguousarray(min_size, np.int32).reshape(-1, 4)
        counts = np.zeros((len(ms), 4), np.int32)
        digests = np.zeros((len(ms), 3), np.uint64)
        self.L.lvo_generate_batch_digests(self.h, len(ms), _ptr(ms), _ptr(counts), _ptr(digests))
        return counts, digests

    def batch_counts(self, min_size):
        ms = np.ascontiguousarray(min_size, np.int32).reshape(-1, 4)
        counts = np.zeros((len(ms), 4), np.int32)
        threads = self.L.lvo_generate_batch_counts(self.h, len(ms), _ptr(ms), _ptr(counts))
        return counts, threads


def fnv1a64(data):
    """FNV-1a 64 of an array's bytes: the digest lvo_generate_batch_digests takes of each chunk's
    arrays, applied by the tests to the CUDA path's output"""
    b = np.frombuffer(bytes(data), np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data).view(np.uint8).reshape(-1)
    return int(lib().lvo_fnv1a64(_ptr(b) if len(b) else None, len(b)))


def csg_operation_bounds(op):
    """CalcCSGOperationBounds, leven/src/clipmap.cpp:1638-1643 (CSG_OFFSET 0.5,
    CSG_BOUNDS_FUDGE 2 -- clipmap.cpp:31-32)."""
    half = [int(op.dimensions[i] * LEAF_SIZE_SCALE) + 2 for i in range(3)]
    org = [int((op.origin[i] - 0.5) * LEAF_SIZE_SCALE) for i in range(3)]
    return [org[i] - half[i] for i in range(3)], [org[i] + half[i] for i in range(3)]
