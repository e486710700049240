"""oracle/ref.py -- drive the reference's own OpenCL C kernels, compiled for the host CPU
(oracle/_ref/libleven_cl_ref.so, built by `make -C oracle ref` from /root/reference/leven/cl).

TEST INFRASTRUCTURE: used to pin the C restatement (oracle/lvn_oracle.c) and to generate the
golden vectors under tests/golden/ (tests/golden/gen_ref_vectors.py).  Only tests/, smoke() and
bench.py's CPU-baseline legs may import this module; the product (leven_b200/) never does.

The kernels are the reference's text.  What is restated HERE is the host orchestration that
strings them together -- which kernel, which arguments, which NDRange, which scan between --
following the reference host files line by line:
    GenerateDefaultDensityField / FindDefaultEdges   leven/src/compute_density_field.cpp:138-231
    ConstructOctreeFromField                          leven/src/compute_octree.cpp:25-150
    GenerateMeshFromOctree                            leven/src/compute_octree.cpp:189-271
    GatherSeamNodesFromOctree                         leven/src/compute_octree.cpp:275-322
    Cuckoo_InitialiseTable / Cuckoo_InsertKeys        leven/src/compute_cuckoo.cpp:49-138
    ExclusiveScan (scan.cl needs work-group barriers; it computes a plain exclusive prefix sum
    and returns data[n-1] + scan[n-1])                leven/src/compute.cpp:328-420
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libleven_cl_ref.so")
REFERENCE_CL = "/root/reference/leven/cl"

LEAF_SIZE_SCALE = 4
CLIPMAP_LEAF_SIZE = 256
CUCKOO_EMPTY = np.uint64(0xFFFFFFFFFFFFFFFF)
MIN_TABLE_SIZE = 2048            # compute_cuckoo.cpp:16

SEAM_DTYPE = np.dtype([("localspaceMin", np.int32, 4), ("position", np.float32, 4), ("normal", np.float32, 4)])
VERTEX_DTYPE = np.dtype([("xyz", np.float32, 4), ("normal", np.float32, 4), ("colour", np.float32, 4)])
QEF_DTYPE = np.dtype([("ATA", np.float32, 6), ("pad", np.float32, 2), ("ATb", np.float32, 4), ("masspoint", np.float32, 4)])
CSG_DTYPE = np.dtype([("type", np.int32), ("brushShape", np.int32), ("material", np.int32), ("rotateY", np.float32),
                      ("origin", np.float32, 4), ("dimensions", np.float32, 4)])

_libs = {}


def _lib_path(V):
    return LIB_PATH if V == 64 else os.path.join(_HERE, "_ref", f"libleven_cl_ref_v{V}.so")


def available(V=64):
    return os.path.exists(_lib_path(V))


def build(force=False):
    """compile the reference kernels where they lie; a no-op (returns False) when /root/reference is absent"""
    if not os.path.isdir(REFERENCE_CL):
        return available()
    deps = [os.path.join(_HERE, "ref_shim", f) for f in ("ref_kernels.cpp", "clc.hpp", "translate.py", "ref_clipmap_caller.cpp")] + \
           [os.path.join(_HERE, "Makefile"), os.path.join(os.path.dirname(_HERE), "include", "leven_compute.hpp")]
    libs = [_lib_path(V) for V in (64, 32, 16)]
    if os.path.exists(PRODUCT_LIB_PATH):
        libs.append(CALLER_LIB_PATH)      # links against the product library: built once that exists
    stale = any(not os.path.exists(l) for l in libs) or any(os.path.getmtime(d) > min(os.path.getmtime(l) for l in libs) for d in deps)
    if force or stale:
        env = dict(os.environ)
        env.pop("CC", None); env.pop("CXX", None)
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "ref"], env=env)
    return True


def lib(V=64):
    """the reference programs built for voxelsPerChunk = V (64, 32 or 16)"""
    if V not in _libs:
        if not available(V):
            raise ImportError(f"{_lib_path(V)} is missing: `make -C oracle ref` (needs /root/reference)")
        L = C.CDLL(_lib_path(V))
        L.ref_DensityFunc.restype = C.c_float
        L.ref_DensityFunc.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float]
        L.ref_snoise2.restype = C.c_float
        L.ref_snoise2.argtypes = [C.c_void_p, C.c_float, C.c_float]
        L.ref_snoise3.restype = C.c_float
        L.ref_snoise3.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float]
        L.ref_Cuckoo_Find.restype = C.c_uint32
        L.ref_Cuckoo_Hash.restype = C.c_uint32
        assert L.ref_voxels_per_chunk() == V
        _libs[V] = L
    return _libs[V]


def set_num_threads(n):
    """OpenMP threads over the work-items of each NDRange (torchrun exports OMP_NUM_THREADS=1)"""
    return int(lib().ref_set_num_threads(int(n)))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _i4(v):
    return np.array([int(v[0]), int(v[1]), int(v[2]), 0], np.int32)


def exclusive_scan(data):
    """compute.cpp:328-420: scan[i] = sum(data[:i]); returns (scan, data[n-1] + scan[n-1])"""
    data = np.asarray(data, np.int32)
    scan = np.zeros(len(data), np.int32)
    if len(data):
        np.cumsum(data[:-1], out=scan[1:])
    total = int(data[-1] + scan[-1]) if len(data) else 0
    return scan, total


def find_next_prime(n):
    """primes.cpp: smallest prime >= n"""
    def is_prime(k):
        if k < 2:
            return False
        i = 2
        while i * i <= k:
            if k % i == 0:
                return False
            i += 1
        return True
    while not is_prime(n):
        n += 1
    return n


def colour_for_min_leaf_size(min_leaf_size):
    """ColourForMinLeafSize, clipmap.cpp:329-352"""
    table = {1: (0.3, 0.1, 0.0), 2: (0.0, 0.0, 0.5), 4: (0.0, 0.5, 0.5), 8: (0.5, 0.0, 0.5), 16: (0.0, 0.5, 0.0)}
    return table.get(int(min_leaf_size), (0.5, 0.0, 0.0))


class Cuckoo:
    """CuckooData + Cuckoo_InitialiseTable + Cuckoo_InsertKeys (compute_cuckoo.cpp:49-138).  The hash
    parameters come from std::mt19937 + uniform_int_distribution there (implementation-defined,
    unobservable); here numpy's generator draws from the same range [2^15, 2^30]."""
    _rng = np.random.default_rng(5489)

    def __init__(self, count):
        self.prime = find_next_prime(max(MIN_TABLE_SIZE, 2 * int(count)))
        self.table = np.full(self.prime, CUCKOO_EMPTY, np.uint64)
        # the reference's stash has 101 entries but is indexed mod prime (cuckoo.cl:67-69): give it room
        self.stash = np.full(self.prime, CUCKOO_EMPTY, np.uint64)
        self.params = self._draw()
        self.stash_used = 0
        self.retries = 0

    def _draw(self):
        return self._rng.integers(1 << 15, (1 << 30) + 1, size=10, dtype=np.uint32)

    def insert_keys(self, keys):
        keys = np.ascontiguousarray(keys, np.uint32)
        n = len(keys)
        inserted = np.zeros(n, np.int32)
        stash_used = np.zeros(n, np.int32)
        inserted_count = 0
        while True:
            if inserted_count > 0:
                self.retries += 1
                self.params = self._draw()
                self.table[:] = CUCKOO_EMPTY
                self.stash[:] = CUCKOO_EMPTY
            lib().ref_Cuckoo_InsertKeys(n, _p(keys), _p(self.table), _p(self.stash), C.c_uint32(self.prime),
                                        _p(self.params), _p(inserted), _p(stash_used))
            _, inserted_count = exclusive_scan(inserted)
            _, stash_count = exclusive_scan(stash_used)
            if inserted_count >= n:
                break
            if self.retries > 64:
                raise RuntimeError("cuckoo insert does not converge")
        self.stash_used |= int(stash_count != 0)

    def find(self, key):
        return lib().ref_Cuckoo_Find(C.c_uint32(int(key)), _p(self.table), _p(self.stash), C.c_uint32(self.prime),
                                     _p(self.params), int(self.stash_used))


class RefWorld:
    """The reference pipeline for one chunk, kernels from the reference text."""

    def __init__(self, image, default_material=0, voxels_per_chunk=64):
        self.V = int(voxels_per_chunk)
        self.H, self.F = self.V + 1, self.V + 2
        self.shift = self.V.bit_length()          # VOXEL_INDEX_SHIFT = log2(V) + 1
        self.L = lib(self.V)
        self.image = np.ascontiguousarray(image, np.uint8)
        assert self.image.size == 256 * 256 * 4
        self.default_material = int(default_material)

    def density(self, x, y, z):
        return self.L.ref_DensityFunc(_p(self.image), x, y, z)

    def snoise2(self, x, y):
        return self.L.ref_snoise2(_p(self.image), x, y)

    def snoise3(self, x, y, z):
        return self.L.ref_snoise3(_p(self.image), x, y, z)

    # ---- compute_density_field.cpp:138-161 ----
    def generate_field(self, mn, size):
        off = _i4([m // LEAF_SIZE_SCALE for m in mn])          # LeafScaleVec(field->min)
        sample_scale = size // (self.V * LEAF_SIZE_SCALE)
        field = np.zeros(self.F ** 3, np.int32)
        self.L.ref_GenerateDefaultField(_p(self.image), _p(off), sample_scale, self.default_material, _p(field))
        return field

    # ---- compute_density_field.cpp:165-231 ----
    def find_edges(self, mn, size, materials):
        off = _i4([m // LEAF_SIZE_SCALE for m in mn])
        n = 3 * self.H ** 3
        occupancy = np.zeros(n, np.int32)
        indices = np.zeros(n, np.int32)
        self.L.ref_FindFieldEdges(_p(off), _p(materials), _p(occupancy), _p(indices))
        scan, num_edges = exclusive_scan(occupancy)
        if num_edges == 0:
            return np.zeros(0, np.int32), np.zeros((0, 4), np.float32)
        compact = np.zeros(num_edges, np.int32)
        self.L.ref_CompactEdges(n, _p(occupancy), _p(scan), _p(indices), _p(compact))
        info = np.zeros((num_edges, 4), np.float32)
        sample_scale = size // (self.V * LEAF_SIZE_SCALE)
        self.L.ref_FindEdgeIntersectionInfo(_p(self.image), _p(off), sample_scale, num_edges, _p(compact), _p(info))
        return compact, info

    # ---- compute_octree.cpp:25-150 ----
    def construct_octree(self, mn, size, materials, edge_keys, edge_info):
        n3 = self.V ** 3
        occ = np.zeros(n3, np.int32); einfo = np.zeros(n3, np.int32)
        codes = np.zeros(n3, np.int32); mats = np.zeros(n3, np.int32)
        self.L.ref_FindActiveVoxels(_p(materials), _p(occ), _p(einfo), _p(codes), _p(mats))
        scan, num_nodes = exclusive_scan(occ)
        if num_nodes <= 0:
            return None
        c_codes = np.zeros(num_nodes, np.int32); c_einfo = np.zeros(num_nodes, np.int32); c_mats = np.zeros(num_nodes, np.int32)
        self.L.ref_CompactVoxels(n3, _p(occ), _p(einfo), _p(codes), _p(mats), _p(scan), _p(c_codes), _p(c_einfo), _p(c_mats))
        edge_table = Cuckoo(len(edge_keys))
        edge_table.insert_keys(edge_keys)
        sample_scale = size // (self.V * LEAF_SIZE_SCALE)
        normals = np.zeros((num_nodes, 4), np.float32)
        qefs = np.zeros(num_nodes, QEF_DTYPE)
        edge_info = np.ascontiguousarray(edge_info, np.float32)
        self.L.ref_CreateLeafNodes(num_nodes, sample_scale, _p(c_codes), _p(c_einfo), _p(edge_info), _p(normals), _p(qefs),
                                   _p(edge_table.table), _p(edge_table.stash), C.c_uint32(edge_table.prime),
                                   _p(edge_table.params), int(edge_table.stash_used))
        positions = np.zeros((num_nodes, 4), np.float32)
        wso = np.array([mn[0], mn[1], mn[2], 0], np.float32)
        self.L.ref_SolveQEFs(num_nodes, _p(wso), _p(qefs), _p(positions))
        node_table = Cuckoo(num_nodes)
        node_table.insert_keys(c_codes.view(np.uint32))
        return dict(numNodes=num_nodes, codes=c_codes.view(np.uint32), edgeMasks=c_einfo, matWords=c_mats, qefs=qefs,
                    positions=positions, normals=normals, table=node_table, edgeTable=edge_table)

    # ---- compute_octree.cpp:189-271 ----
    def generate_mesh(self, size, octree):
        n = octree["numNodes"]
        index_buffer = np.zeros(n * 18, np.int32)
        tri_valid = np.zeros(n * 3, np.int32)
        t = octree["table"]
        self.L.ref_GenerateMesh(n, _p(octree["codes"]), _p(octree["matWords"]), _p(index_buffer), _p(tri_valid),
                                _p(t.table), _p(t.stash), C.c_uint32(t.prime), _p(t.params), int(t.stash_used))
        scan, num_quads = exclusive_scan(tri_valid)
        compact = np.zeros(max(num_quads, 0) * 6, np.int32)
        if num_quads > 0:
            self.L.ref_CompactMeshTriangles(n * 3, _p(tri_valid), _p(scan), _p(index_buffer), _p(compact))
        colour = np.array(list(colour_for_min_leaf_size(size // CLIPMAP_LEAF_SIZE)) + [0.0], np.float32)
        vertices = np.zeros(n, VERTEX_DTYPE)
        self.L.ref_GenerateMeshVertexBuffer(n, _p(octree["positions"]), _p(octree["normals"]), _p(octree["matWords"]),
                                            _p(colour), _p(vertices))
        return vertices, compact.reshape(-1, 3)

    # ---- compute_octree.cpp:275-322 ----
    def gather_seam_nodes(self, octree):
        n = octree["numNodes"]
        is_seam = np.zeros(n, np.int32)
        self.L.ref_FindSeamNodes(n, _p(octree["codes"]), _p(is_seam))
        scan, num = exclusive_scan(is_seam)
        out = np.zeros(max(num, 0), SEAM_DTYPE)
        if num > 0:
            self.L.ref_ExtractSeamNodeInfo(n, _p(is_seam), _p(scan), _p(octree["codes"]), _p(octree["matWords"]),
                                           _p(octree["positions"]), _p(octree["normals"]), _p(out))
        return out

    # ---- Compute_GenerateChunkMesh, compute_octree.cpp:351-375 (no caches: one fresh chunk) ----
    def generate_chunk_mesh(self, mn, size):
        materials = self.generate_field(mn, size)
        keys, info = self.find_edges(mn, size, materials)
        out = dict(materials=materials, edgeKeys=keys, edgeInfo=info, numEdges=len(keys), numNodes=0, numTriangles=0,
                   numSeamNodes=0)
        if len(keys) == 0:
            return out
        octree = self.construct_octree(mn, size, materials, keys, info)
        if octree is None:
            return out
        vertices, tris = self.generate_mesh(size, octree)
        seams = self.gather_seam_nodes(octree)
        out.update(numNodes=octree["numNodes"], numTriangles=len(tris), numSeamNodes=len(seams),
                   codes=octree["codes"], edgeMasks=octree["edgeMasks"], matWords=octree["matWords"], qefs=octree["qefs"],
                   positions=octree["positions"], normals=octree["normals"], vertices=vertices, indices=tris, seams=seams,
                   cuckooRetries=octree["table"].retries + octree["edgeTable"].retries)
        return out

    # ---- apply_csg_operation.cl, kernel by kernel (compute_csg.cpp:11-220 strings them together) ----
    def csg_materials(self, mn, size, ops, materials):
        """CSG_HermiteIndices over the whole field: (updated flags, positions, new materials)"""
        off = _i4([m // LEAF_SIZE_SCALE for m in mn])
        sample_scale = size // (self.V * LEAF_SIZE_SCALE)
        ops = np.ascontiguousarray(ops, CSG_DTYPE)
        n = self.F ** 3
        updated = np.zeros(n, np.int32); positions = np.zeros((n, 4), np.int32); new_mats = np.zeros(n, np.int32)
        self.L.ref_CSG_HermiteIndices(_p(off), len(ops), _p(ops), sample_scale, _p(np.ascontiguousarray(materials, np.int32)),
                                      _p(updated), _p(positions), _p(new_mats))
        return updated, positions, new_mats

    def csg_updated_edges(self, positions):
        positions = np.ascontiguousarray(positions, np.int32).reshape(-1, 4)
        out = np.zeros(len(positions) * 6, np.int32)
        self.L.ref_CSG_FindUpdatedEdges(len(positions), _p(positions), _p(out))
        return out

    def csg_filter_valid_edges(self, edge_indices, materials):
        edge_indices = np.ascontiguousarray(edge_indices, np.int32)
        valid = np.zeros(len(edge_indices), np.int32)
        self.L.ref_CSG_FilterValidEdges(len(edge_indices), _p(edge_indices), _p(np.ascontiguousarray(materials, np.int32)), _p(valid))
        return valid

    def csg_edge_info(self, mn, size, ops, edge_keys):
        off = _i4([m // LEAF_SIZE_SCALE for m in mn])
        sample_scale = size // (self.V * LEAF_SIZE_SCALE)
        ops = np.ascontiguousarray(ops, CSG_DTYPE)
        edge_keys = np.ascontiguousarray(edge_keys, np.int32)
        out = np.zeros((len(edge_keys), 4), np.float32)
        self.L.ref_CSG_FindEdgeIntersectionInfo(_p(off), len(ops), _p(ops), sample_scale, len(edge_keys), _p(edge_keys), _p(out))
        return out

    def apply_csg(self, mn, size, ops, materials, edge_keys, edge_info):
        """ApplyCSGOperations, compute_csg.cpp:11-220, on one field (materials, edge list):
        returns the edited (materials, edge_keys, edge_info).  The scans, compactions and
        RemoveDuplicates between the kernels are numpy (their results are defined: stable
        compaction, set of unique keys -- np.unique gives the set in ascending order, the
        reference's hash order is arbitrary); PruneFieldEdges' O(E*K) membership loop is np.isin."""
        materials = np.array(materials, np.int32)
        edge_keys = np.asarray(edge_keys, np.int32)
        edge_info = np.asarray(edge_info, np.float32).reshape(-1, 4)
        if len(ops) == 0:
            return materials, edge_keys, edge_info
        updated, positions, new_mats = self.csg_materials(mn, size, ops, materials)
        sel = np.nonzero(updated)[0]
        if len(sel) == 0:
            return materials, edge_keys, edge_info
        pts = positions[sel]                                   # CompactPoints
        F = self.F
        materials[pts[:, 0] + F * pts[:, 1] + F * F * pts[:, 2]] = new_mats[sel]   # UpdateFieldMaterials
        gen = self.csg_updated_edges(pts)
        gen = gen[gen != -1]                                   # RemoveInvalidIndices + CompactIndexArray
        invalidated = np.unique(gen)                           # RemoveDuplicates
        # FilterValidEdges reads one sample past the grid for edges that start at index F-1 (the
        # reference does the same read on the device): pad the field so the read stays in bounds
        padded = np.concatenate([materials, np.full(F * F + F + 2, 201, np.int32)])
        valid = self.csg_filter_valid_edges(invalidated, padded)
        created = invalidated[valid != 0]
        if len(invalidated) and len(edge_keys):                # PruneFieldEdges + CompactFieldEdges
            keep = ~np.isin(edge_keys, invalidated)
            if keep.sum() > 0:                                 # compute_csg.cpp:160: no swap when nothing survives
                edge_keys, edge_info = edge_keys[keep], edge_info[keep]
        if len(created):
            info = self.csg_edge_info(mn, size, ops, created)
            if len(edge_keys):
                edge_keys = np.concatenate([edge_keys, created]); edge_info = np.concatenate([edge_info, info])
            else:
                edge_keys, edge_info = created, info
        return materials, edge_keys, edge_info


# ---------------------------------------------------------------------------------------------
# Seam meshes: the reference's seam octree (leven/src/octree.cpp, compiled
# unmodified into oracle/_ref/libleven_octree_ref.so) behind the selection logic of clipmap.cpp.
# ---------------------------------------------------------------------------------------------
OCTREE_LIB_PATH = os.path.join(_HERE, "_ref", "libleven_octree_ref.so")
CHILD_MIN_OFFSETS = [(0, 0, 0), (0, 0, 1), (0, 1, 0), (0, 1, 1), (1, 0, 0), (1, 0, 1), (1, 1, 0), (1, 1, 1)]   # volume_constants.h:24-35
_octree_lib = None


def octree_available():
    return os.path.exists(OCTREE_LIB_PATH)


def octree_lib():
    global _octree_lib
    if _octree_lib is None:
        if not octree_available():
            raise ImportError(f"{OCTREE_LIB_PATH} is missing: `make -C oracle ref` (needs /root/reference)")
        _octree_lib = C.CDLL(OCTREE_LIB_PATH)
    return _octree_lib


def filter_seam_node(child_index, seam_bounds, mn, mx):
    """FilterSeamNode, clipmap.cpp:508-536"""
    b = seam_bounds
    if child_index == 0:
        return mx[0] == b[0] or mx[1] == b[1] or mx[2] == b[2]
    if child_index == 1:
        return mn[2] == b[2]
    if child_index == 2:
        return mn[1] == b[1]
    if child_index == 3:
        return mn[1] == b[1] or mn[2] == b[2]
    if child_index == 4:
        return mn[0] == b[0]
    if child_index == 5:
        return mn[0] == b[0] or mn[2] == b[2]
    if child_index == 6:
        return mn[0] == b[0] or mn[1] == b[1]
    if child_index == 7:
        return mn[0] == b[0] and mn[1] == b[1] and mn[2] == b[2]
    return False


def select_seam_nodes(host_min, host_size, neighbours, V=64):
    """GenerateMeshDataForNode's node construction (clipmap.cpp:398-415) + SelectSeamNodes
    (clipmap.cpp:542-569) over the neighbour list of GenerateClipmapSeamMesh (clipmap.cpp:573-611).
    neighbours: [(neighbourIndex 0..7, neighbourMin, neighbourSize, SeamNodeInfo array)].
    Returns (minSize int[n][4], positions float[n][3], normals float[n][3], materialInfo int[n])."""
    seam_bounds = [host_min[i] + host_size for i in range(3)]
    lo, hi = list(host_min), [host_min[i] + 2 * host_size for i in range(3)]       # AABB(min, hostNodeSize * 2)
    ms, pos, nrm, mat = [], [], [], []
    for index, nb_min, nb_size, nodes in neighbours:
        seam_node_size = nb_size // V                                              # clipmap.cpp:399
        size = nb_size // (V * LEAF_SIZE_SCALE)                                    # clipmap.cpp:555
        for nd in nodes:
            lm = nd["localspaceMin"]
            mn = [int(lm[i]) * seam_node_size + nb_min[i] for i in range(3)]
            mx = [mn[i] + size * LEAF_SIZE_SCALE for i in range(3)]
            inside = all(lo[i] <= mn[i] < hi[i] for i in range(3))                 # aabb.pointIsInside(node->min)
            if not filter_seam_node(index, seam_bounds, mn, mx) or not inside:
                continue
            ms.append(mn + [seam_node_size])
            pos.append([float(v) for v in nd["position"][:3]])
            nrm.append([float(v) for v in nd["normal"][:3]])
            mat.append(int(lm[3]))
    return (np.array(ms, np.int32).reshape(-1, 4), np.array(pos, np.float32).reshape(-1, 3),
            np.array(nrm, np.float32).reshape(-1, 3), np.array(mat, np.int32))


def seam_mesh(host_min, host_size, neighbours, colour=(1.0, 1.0, 1.0), V=64):
    """GenerateClipmapSeamMesh (clipmap.cpp:573-611): selection, Octree_ConstructUpwards over
    (host min, 2 * host size), Octree_GenerateMesh.  Returns (vertices VERTEX_DTYPE[], triangles int[n][3])."""
    ms, pos, nrm, mat = select_seam_nodes(host_min, host_size, neighbours, V)
    n = len(ms)
    verts = np.zeros(max(n, 1), VERTEX_DTYPE)
    tris = np.zeros((max(4 * n, 1) * 3, 3), np.int32)
    nv = C.c_int(0)
    root = np.array(list(host_min), np.int32)
    col = np.array(colour, np.float32)
    rc = octree_lib().ref_seam_octree_mesh(n, _p(ms), _p(pos), _p(nrm), _p(mat), _p(root), int(2 * host_size), _p(col),
                                           _p(verts), len(verts), C.byref(nv), _p(tris), len(tris))
    assert rc >= 0, "seam mesh buffers too small"
    return verts[:nv.value].copy() if rc > 0 else verts[:0].copy(), tris[:rc].copy()


# ---------------------------------------------------------------------------------------------
# Mesh simplification: the reference's ng_mesh_simplify.cpp + qef_simd.h compiled
# into oracle/_ref/libleven_simplify_ref.so (see ref_shim/ref_simplify.cpp for what the shim defines)
# ---------------------------------------------------------------------------------------------
SIMPLIFY_LIB_PATH = os.path.join(_HERE, "_ref", "libleven_simplify_ref.so")
_simplify_lib = None
# MeshSimplificationOptions defaults (ng_mesh_simplify.h:6-28) with the values the clipmap passes at
# LOD0 (clipmap.cpp:449-465, options.h:14-16): maxError 5 * leafSize, maxEdgeSize 2.5 * leafSize,
# minAngleCosine 0.7, leafSize = LEAF_SIZE_SCALE * size / CLIPMAP_LEAF_SIZE
SIMPLIFY_DEFAULTS = dict(edgeFraction=0.125, maxIterations=10, targetPercentage=0.05, maxError=5.0, maxEdgeSize=2.5,
                         minAngleCosine=0.8)


def simplify_available():
    return os.path.exists(SIMPLIFY_LIB_PATH)


# ---- the reference's own caller (clipmap.cpp:329-504) on top of the C ABI: ref_shim/ref_clipmap_caller.cpp ----
PRODUCT_LIB_PATH = os.path.join(os.path.dirname(_HERE), "leven_b200", "lib", "libleven_b200.so")
CALLER_LIB_PATH = os.path.join(_HERE, "_ref", "libleven_clipmap_caller.so")
_caller_lib = None


def caller_available():
    return os.path.exists(CALLER_LIB_PATH)


def caller_construct_node(mn, size, collision=False, voxels_per_chunk=64, mesh_max_error=5.0, mesh_max_edge_len=2.5,
                          mesh_max_angle=0.7):
    """ConstructClipmapNodeData / ConstructCollisionNodeData (clipmap.cpp:431-504) -- the reference's own code,
    compiled against include/leven_compute.hpp -- for one node.  The process must have initialised the product
    library (Compute_Initialise).  Returns dict(vertices, triangles, seamMinSize, seamPosition, seamNormal,
    seamColour, seamMaterial, active)."""
    global _caller_lib
    if _caller_lib is None:
        _caller_lib = C.CDLL(CALLER_LIB_PATH)
    v = np.zeros(14 * 1024, VERTEX_DTYPE); t = np.zeros((28 * 1024, 3), np.int32)
    cap = 64 * 64 * 6 + 16
    sm = np.zeros((cap, 4), np.int32); sf = np.zeros((cap, 9), np.float32); mat = np.zeros(cap, np.int32)
    nv, nt, ns, act = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
    f = C.c_float
    m3 = np.array(list(mn)[:3], np.int32)
    rc = _caller_lib.refcaller_construct_node(int(voxels_per_chunk), int(bool(collision)), _p(m3), int(size), f(mesh_max_error),
                                              f(mesh_max_edge_len), f(mesh_max_angle), _p(v), len(v), C.byref(nv), _p(t), len(t),
                                              C.byref(nt), _p(sm), _p(sf), _p(mat), cap, C.byref(ns), C.byref(act))
    assert rc == 0, rc
    n = ns.value
    return dict(vertices=v[:nv.value].copy(), triangles=t[:nt.value].copy(), seamMinSize=sm[:n].copy(), seamPosition=sf[:n, 0:3].copy(),
                seamNormal=sf[:n, 3:6].copy(), seamColour=sf[:n, 6:9].copy(), seamMaterial=mat[:n].copy(), active=bool(act.value))


def simplify_lib():
    global _simplify_lib
    if _simplify_lib is None:
        if not simplify_available():
            raise ImportError(f"{SIMPLIFY_LIB_PATH} is missing: `make -C oracle ref` (needs /root/reference)")
        _simplify_lib = C.CDLL(SIMPLIFY_LIB_PATH)
    return _simplify_lib


def clipmap_simplify_options(node_size):
    """ConstructClipmapNodeData, clipmap.cpp:449-465"""
    leaf = float(LEAF_SIZE_SCALE * (node_size // CLIPMAP_LEAF_SIZE))
    o = dict(SIMPLIFY_DEFAULTS)
    o.update(maxError=5.0 * leaf, maxEdgeSize=2.5 * leaf, minAngleCosine=0.7)
    return o


def simplify_mesh(vertices, triangles, world_space_offset, options):
    """ngMeshSimplifier (ng_mesh_simplify.cpp:441-540) -> (vertices VERTEX_DTYPE[], triangles int[n][3])"""
    v = np.array(vertices, VERTEX_DTYPE)
    t = np.ascontiguousarray(np.asarray(triangles, np.int32).reshape(-1, 3)).copy()
    nv, nt = C.c_int(len(v)), C.c_int(len(t))
    off = np.array(list(world_space_offset)[:3] + [0.0], np.float32)
    f = C.c_float
    rc = simplify_lib().ref_mesh_simplify(_p(v), C.byref(nv), _p(t), C.byref(nt), _p(off), f(options["edgeFraction"]),
                                          int(options["maxIterations"]), f(options["targetPercentage"]), f(options["maxError"]),
                                          f(options["maxEdgeSize"]), f(options["minAngleCosine"]))
    assert rc == 0, "mesh exceeds the reference's MeshBuffer capacity"
    return v[:nv.value].copy(), t[:nt.value].copy()


def random_edges(num_edges, count):
    out = np.zeros(count, np.int32)
    simplify_lib().ref_random_edges(int(num_edges), int(count), _p(out))
    return out
