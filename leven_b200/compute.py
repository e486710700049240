"""Host-side mirror of the reference compute interface (leven/src/compute.h:12-81)
over the C ABI of libleven_b200.so (include/leven_compute.h).

Same names, argument meaning and error behaviour as the reference: every call
returns an int, 0 = success, negative = failure (callers test ``< 0``,
clipmap.cpp:379).  There is no CPU fallback: importing works without a GPU (so
the symbol table can be checked), but every compute call fails with
LVN_ERR_NO_DEVICE unless a CUDA device runs the sm_100a kernels.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libleven_b200.so")

LVN_SUCCESS = 0
LVN_CL_ERROR = -99999
LVN_ERR_NO_DEVICE = -1
LVN_ERR_OUT_OF_MEMORY = -4
LVN_ERR_INVALID_VALUE = -30
LVN_ERR_NOT_INITIALISED = -34
LVN_ERR_CAPACITY = -61
LVN_ERR_CUDA = -9999

MATERIAL_NONE = 200      # volume_materials.h:7
MATERIAL_AIR = 201       # volume_materials.h:8
LEAF_SIZE_SCALE = 4      # volume_constants.h:7-8
MAX_MESH_VERTICES = 14 * 1024          # render_types.h:62-68 (LEVEN defined)
MAX_MESH_TRIANGLES = MAX_MESH_VERTICES * 2
RenderShape_Cube, RenderShape_Sphere = 0, 1   # render_types.h:11-20

LVN_STAGES = ("columns", "classify", "hermite", "leaves", "field", "csg", "cuckoo", "solve")

# POD layouts (static_asserted against the C header in tests/test_abi.py)
MeshVertex = np.dtype([("xyz", np.float32, 4), ("normal", np.float32, 4), ("colour", np.float32, 4)])
MeshTriangle = np.dtype([("indices_", np.int32, 3)])
SeamNodeInfo = np.dtype([("localspaceMin", np.int32, 4), ("position", np.float32, 4), ("normal", np.float32, 4)])
ChunkResult = np.dtype([("numEdges", np.int32), ("numVertices", np.int32), ("numTriangles", np.int32),
                        ("numSeamNodes", np.int32), ("vertexOffset", np.int32), ("triangleOffset", np.int32),
                        ("seamOffset", np.int32), ("status", np.int32)])


class CSGOperationInfo(C.Structure):
    """compute.h:16-24"""
    _fields_ = [("type", C.c_int32), ("brushShape", C.c_int32), ("material", C.c_int32),
                ("rotateY", C.c_float), ("origin", C.c_float * 4), ("dimensions", C.c_float * 4)]

    @classmethod
    def make(cls, type_, shape, material, origin, dimensions, rotate_y=0.0):
        op = cls()
        op.type, op.brushShape, op.material, op.rotateY = int(type_), int(shape), int(material), float(rotate_y)
        for i in range(3):
            op.origin[i] = float(origin[i])
            op.dimensions[i] = float(dimensions[i])
        op.origin[3] = 0.0
        op.dimensions[3] = 0.0
        return op


class AABB(C.Structure):
    """aabb.h:95-96"""
    _fields_ = [("min", C.c_int32 * 3), ("max", C.c_int32 * 3)]


class BatchDeviceView(C.Structure):
    _fields_ = [("vertices", C.c_void_p), ("triangles", C.c_void_p), ("seamNodes", C.c_void_p),
                ("totalVertices", C.c_int64), ("totalTriangles", C.c_int64), ("totalSeamNodes", C.c_int64),
                ("totalEdges", C.c_int64), ("nonEmptyChunks", C.c_int32)]


class StageDump(C.Structure):
    _fields_ = [("edgeCapacity", C.c_int32), ("nodeCapacity", C.c_int32),
                ("numEdges", C.c_int32), ("numNodes", C.c_int32), ("numTriangles", C.c_int32),
                ("numSeamNodes", C.c_int32),
                ("materials", C.c_void_p), ("edgeKeys", C.c_void_p), ("edgeInfo", C.c_void_p),
                ("nodeCodes", C.c_void_p), ("nodeEdgeMasks", C.c_void_p), ("nodeMaterials", C.c_void_p),
                ("nodeQEFs", C.c_void_p), ("nodePositions", C.c_void_p), ("nodeNormals", C.c_void_p)]


class StageStats(C.Structure):
    _fields_ = [("ms", C.c_double * len(LVN_STAGES)), ("launches", C.c_int64 * len(LVN_STAGES)),
                ("terrainEvals", C.c_int64), ("edges", C.c_int64), ("edgesY", C.c_int64), ("nodes", C.c_int64),
                ("triangles", C.c_int64), ("seamNodes", C.c_int64), ("chunks", C.c_int64),
                ("nonEmptyChunks", C.c_int64)]


# every symbol include/leven_compute.h declares: name -> (restype, argtypes)
_P, _I, _U, _F, _I64 = C.c_void_p, C.c_int, C.c_uint, C.c_float, C.c_int64
ABI = {
    "lvn_compute_set_device": (_I, [_I]),
    "lvn_compute_bind_host_numa": (_I, [_I, _P, _P]),
    "lvn_compute_initialise": (_I, [_I, _U, _I]),
    "lvn_compute_shutdown": (_I, []),
    "lvn_compute_set_noise_seed": (_I, [_I]),
    "lvn_compute_set_noise_image": (_I, [_P]),
    "lvn_compute_get_noise_image": (_I, [_P]),
    "lvn_compute_set_density_function": (_I, [_I, _F]),
    "lvn_compute_store_csg_operation": (_I, [_P, _P]),
    "lvn_compute_clear_csg_operations": (_I, []),
    "lvn_error_string": (C.c_char_p, [_I]),
    "lvn_last_cuda_error": (C.c_char_p, []),
    "lvn_meshgen_create": (_P, [_I]),
    "lvn_meshgen_destroy": (None, [_P]),
    "lvn_meshgen_voxels_per_chunk": (_I, [_P]),
    "lvn_meshgen_apply_csg_operations": (_I, [_P, _P, _I, _P, _I]),
    "lvn_meshgen_free_chunk_octree": (_I, [_P, _P, _I]),
    "lvn_meshgen_is_chunk_empty": (_I, [_P, _P, _I, _P]),
    "lvn_meshgen_generate_chunk_mesh": (_I, [_P, _P, _I, _P, _I, _P, _P, _I, _P, _P, _I, _P]),
    "lvn_meshgen_generate_batch": (_I, [_P, _I, _P, _P, _I64, _P, _I64, _P, _I64, _P]),
    "lvn_meshgen_generate_batch_device": (_I, [_P, _I, _P, _P, _P]),
    "lvn_meshgen_generate_batch_device_async": (_I, [_P, _I, _P, _P, _P]),
    "lvn_meshgen_wait": (_I, [_P]),
    "lvn_meshgen_generate_batch_async": (_I, [_P, _I, _P, _P, _I64, _P, _I64, _P, _I64, _P]),
    "lvn_meshgen_apply_csg_operations_batch": (_I, [_P, _P, _I, _I, _P]),
    "lvn_meshgen_debug_dump_chunk": (_I, [_P, _P, _I, _P]),
    "lvn_debug_solve_qefs": (_I, [_I, _I, _P, _P]),
    "lvn_meshgen_set_profiling": (_I, [_P, _I]),
    "lvn_meshgen_get_stats": (_I, [_P, _P, _I]),
    "lvn_meshgen_set_stream": (_I, [_P, _P]),
    "lvn_meshgen_set_pipeline": (_I, [_P, _I, _I]),
    "lvn_meshgen_get_pipeline": (_I, [_P, _P, _P]),
    "lvn_alloc_pinned": (_P, [C.c_size_t]),
    "lvn_free_pinned": (None, [_P]),
    "lvn_measure_fp32_peak": (_I, [_P]),
    "lvn_global_mesh_offsets": (_I, [_P, _I, _I, _I, _P, _P, _P]),
    "lvn_find_next_prime": (_I, [_I]),
    "lvn_exclusive_scan": (_I, [_P, _P, _I]),
    "lvn_compact_index_array": (_I, [_P, _P, _I, _P]),
    "lvn_remove_duplicates": (_I, [_P, _I, _P]),
    "lvn_cuckoo_create": (_P, [_U]),
    "lvn_cuckoo_insert_keys": (_I, [_P, _P, _U]),
    "lvn_cuckoo_find": (_I, [_P, _P, _U, _P]),
    "lvn_cuckoo_prime": (_I, [_P]),
    "lvn_cuckoo_retries": (_I, [_P]),
    "lvn_cuckoo_destroy": (None, [_P]),
    "lvn_seam_mesh_generate_batch": (_I, [_I, _I, _P, _P, _I, _P, _I, _P, _I64, _P, _I64, _P]),
    "lvn_seam_last_error": (C.c_char_p, []),
    "lvn_mesh_simplify_batch": (_I, [_I, _P, _P, _I, _P, _I64, _P, _I64, _P]),
    "lvn_mesh_simplify_last_error": (C.c_char_p, []),
    "lvn_meshgen_generate_simplified_batch": (_I, [_P, _I, _P, _P, _P, _I64, _P, _I64, _P, _I64, _P, _P]),
    "lvn_clipmap_seam_update_batch": (_I, [_I, _P, _I, _P, _I, _P, _I, _P, _I64, _I, _I, _P, _I64, _P, _I64, _P, _P, _P, _P, _P]),
    "lvn_clipmap_update_batch": (_I, [_P, _P, _I, _I, _P, _P, _I64, _I64, _P, _I64, _P, _I64, _P, _P, _P, _P, _P]),
    "lvn_meshgen_generate_collision_batch": (_I, [_P, _I, _P, _P, C.c_float, _P, _P, _I64, _P, _I64, _P, _I64, _P, _P]),
}

_lib = None


def lib():
    """Load libleven_b200.so; fails loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        path = LIB_PATH
        variant = os.environ.get("LVN_LIB_VARIANT")     # experiment builds: profiles/build_variant.sh <name> <nvcc flags>
        if variant:
            path = LIB_PATH[:-3] + "." + variant + ".so"
        if not os.path.exists(path):
            raise ImportError(
                f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "leven_b200 has no CPU fallback.")
        L = C.CDLL(path)
        for name, (res, args) in ABI.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _i3(v):
    return (C.c_int32 * 3)(int(v[0]), int(v[1]), int(v[2]))


def GetCLErrorString(error):
    """compute.h:81"""
    return lib().lvn_error_string(int(error)).decode()


def last_cuda_error():
    return lib().lvn_last_cuda_error().decode()


def Compute_SetDevice(device):
    return lib().lvn_compute_set_device(int(device))


def Compute_BindHostNuma(device):
    """lvn_compute_bind_host_numa -> (node or negative, cpus in the new affinity mask, memory policy set)"""
    cpus, mem = C.c_int(0), C.c_int(0)
    node = lib().lvn_compute_bind_host_numa(int(device), C.byref(cpus), C.byref(mem))
    return node, cpus.value, bool(mem.value)


def Compute_Initialise(noiseSeed, defaultMaterial, numCSGBrushes):
    """compute.h:35"""
    return lib().lvn_compute_initialise(int(noiseSeed), int(defaultMaterial), int(numCSGBrushes))


def Compute_Shutdown():
    """compute.h:36"""
    return lib().lvn_compute_shutdown()


def Compute_SetNoiseSeed(noiseSeed):
    """compute.h:38"""
    return lib().lvn_compute_set_noise_seed(int(noiseSeed))


def Compute_SetNoiseImage(rgba):
    rgba = np.ascontiguousarray(rgba, np.uint8).reshape(-1)
    assert rgba.size == 256 * 256 * 4
    return lib().lvn_compute_set_noise_image(_ptr(rgba))


def Compute_GetNoiseImage():
    out = np.zeros(256 * 256 * 4, np.uint8)
    rc = lib().lvn_compute_get_noise_image(_ptr(out))
    if rc < 0:
        raise RuntimeError(GetCLErrorString(rc))
    return out


def Compute_SetDensityFunction(kind, param=0.5):
    return lib().lvn_compute_set_density_function(int(kind), float(param))


def Compute_StoreCSGOperation(opInfo, aabb_min, aabb_max):
    """compute.h:39"""
    bb = AABB()
    for i in range(3):
        bb.min[i] = int(aabb_min[i])
        bb.max[i] = int(aabb_max[i])
    return lib().lvn_compute_store_csg_operation(C.byref(opInfo), C.byref(bb))


def Compute_ClearCSGOperations():
    """compute.h:40"""
    return lib().lvn_compute_clear_csg_operations()


def CalcCSGOperationBounds(opInfo):
    """clipmap.cpp:1638-1643 (CSG_OFFSET 0.5, CSG_BOUNDS_FUDGE 2, clipmap.cpp:1613-1614): the AABB
    the caller passes to Compute_StoreCSGOperation."""
    half = [int(np.float32(opInfo.dimensions[i]) * np.float32(LEAF_SIZE_SCALE)) + 2 for i in range(3)]
    org = [int((np.float32(opInfo.origin[i]) - np.float32(0.5)) * np.float32(LEAF_SIZE_SCALE)) for i in range(3)]
    return [org[i] - half[i] for i in range(3)], [org[i] + half[i] for i in range(3)]


class MeshBuffer:
    """render_types.h:70-90: fixed-capacity vertex / triangle arrays owned by the caller."""

    def __init__(self, max_vertices=MAX_MESH_VERTICES, max_triangles=None):
        self.vertices = np.zeros(max_vertices, MeshVertex)
        self.triangles = np.zeros(max_triangles if max_triangles is not None else 2 * max_vertices, MeshTriangle)
        self.numVertices = 0
        self.numTriangles = 0


class Compute_MeshGenContext:
    """compute.h:46-77"""

    def __init__(self, handle):
        self.privateCtx_ = handle
        self._L = lib()

    @staticmethod
    def create(voxelsPerChunk):
        """compute.h:50; privateCtx_ may be None on failure, as in the reference (compute.cpp:605-610)."""
        return Compute_MeshGenContext(lib().lvn_meshgen_create(int(voxelsPerChunk)))

    def destroy(self):
        if self.privateCtx_:
            self._L.lvn_meshgen_destroy(self.privateCtx_)
            self.privateCtx_ = None

    def voxelsPerChunk(self):
        return self._L.lvn_meshgen_voxels_per_chunk(self.privateCtx_)

    def applyCSGOperations(self, opInfo, clipmapNodeMin, clipmapNodeSize):
        arr = (CSGOperationInfo * len(opInfo))(*opInfo)
        return self._L.lvn_meshgen_apply_csg_operations(self.privateCtx_, arr, len(opInfo), _i3(clipmapNodeMin),
                                                        int(clipmapNodeSize))

    def freeChunkOctree(self, min_, size):
        return self._L.lvn_meshgen_free_chunk_octree(self.privateCtx_, _i3(min_), int(size))

    def isChunkEmpty(self, min_, size):
        """returns (error, isEmpty)"""
        e = C.c_int(0)
        rc = self._L.lvn_meshgen_is_chunk_empty(self.privateCtx_, _i3(min_), int(size), C.byref(e))
        return rc, bool(e.value)

    def generateChunkMesh(self, min_, clipmapNodeSize, meshBuffer, seamNodeBuffer):
        """compute.h:68-72.  seamNodeBuffer is a python list that is cleared and refilled with one
        SeamNodeInfo ndarray (the std::vector<SeamNodeInfo>& of the reference)."""
        del seamNodeBuffer[:]
        nV, nT, nS = C.c_int(0), C.c_int(0), C.c_int(0)
        cap = 4096
        while True:
            seams = np.zeros(cap, SeamNodeInfo)
            rc = self._L.lvn_meshgen_generate_chunk_mesh(
                self.privateCtx_, _i3(min_), int(clipmapNodeSize),
                _ptr(meshBuffer.vertices), len(meshBuffer.vertices), C.byref(nV),
                _ptr(meshBuffer.triangles), len(meshBuffer.triangles), C.byref(nT),
                _ptr(seams), cap, C.byref(nS))
            if rc == LVN_ERR_CAPACITY and nS.value > cap:
                cap = nS.value      # the vector grows; the MeshBuffer does not
                continue
            break
        if rc < 0:
            return rc
        meshBuffer.numVertices, meshBuffer.numTriangles = nV.value, nT.value
        seamNodeBuffer.append(seams[:nS.value].copy())
        return rc

    # ---- batch entry points (new) ---------------------------------------
    def generateBatch(self, chunkMinSize, vertices, triangles, seamNodes):
        """host arenas (numpy arrays of MeshVertex / MeshTriangle / SeamNodeInfo); returns (error, results)"""
        ms = np.ascontiguousarray(chunkMinSize, np.int32).reshape(-1, 4)
        results = np.zeros(len(ms), ChunkResult)
        rc = self._L.lvn_meshgen_generate_batch(self.privateCtx_, len(ms), _ptr(ms),
                                                _ptr(vertices), len(vertices), _ptr(triangles), len(triangles),
                                                _ptr(seamNodes), len(seamNodes), _ptr(results))
        return rc, results

    def generateBatchDevice(self, chunkMinSize):
        """results stay in HBM; returns (error, results, view)"""
        ms = np.ascontiguousarray(chunkMinSize, np.int32).reshape(-1, 4)
        results = np.zeros(len(ms), ChunkResult)
        view = BatchDeviceView()
        rc = self._L.lvn_meshgen_generate_batch_device(self.privateCtx_, len(ms), _ptr(ms), _ptr(results),
                                                       C.byref(view))
        return rc, results, view

    def generateBatchDeviceAsync(self, chunkMinSize):
        """generateBatchDevice that returns once the counts and offsets are final; the arenas are complete in
        stream order on the context's stream, or after wait().  Returns (error, results, view)"""
        ms = np.ascontiguousarray(chunkMinSize, np.int32).reshape(-1, 4)
        results = np.zeros(len(ms), ChunkResult)
        view = BatchDeviceView()
        rc = self._L.lvn_meshgen_generate_batch_device_async(self.privateCtx_, len(ms), _ptr(ms), _ptr(results),
                                                             C.byref(view))
        return rc, results, view

    def wait(self):
        return self._L.lvn_meshgen_wait(self.privateCtx_)

    def generateBatchAsync(self, chunkMinSize, vertices, triangles, seamNodes):
        """generateBatch that returns once the counts are final and every copy is queued; the host arenas are
        complete after wait().  Returns (error, results)"""
        ms = np.ascontiguousarray(chunkMinSize, np.int32).reshape(-1, 4)
        results = np.zeros(len(ms), ChunkResult)
        rc = self._L.lvn_meshgen_generate_batch_async(self.privateCtx_, len(ms), _ptr(ms),
                                                      _ptr(vertices), len(vertices), _ptr(triangles), len(triangles),
                                                      _ptr(seamNodes), len(seamNodes), _ptr(results))
        return rc, results

    def generateSimplifiedBatch(self, chunkMinSize, vertices, triangles, seamNodes, unitOptions=None):
        """ConstructClipmapNodeData (clipmap.cpp:432-468) over a batch: generateChunkMesh + ngMeshSimplifier,
        the meshes staying in HBM between the two; host arenas receive the simplified meshes.
        unitOptions: SimplifyOptions whose maxError / maxEdgeSize are per leaf size (options.h:14-16).
        Returns (error, results ChunkResult[], simplified SimplifyResult[])"""
        ms = np.ascontiguousarray(chunkMinSize, np.int32).reshape(-1, 4)
        results = np.zeros(len(ms), ChunkResult)
        simp = np.zeros(len(ms), SimplifyResult)
        opt = unitOptions if unitOptions is not None else SimplifyOptions.clipmap_unit()
        rc = self._L.lvn_meshgen_generate_simplified_batch(self.privateCtx_, len(ms), _ptr(ms), C.byref(opt),
                                                           _ptr(vertices), len(vertices), _ptr(triangles), len(triangles),
                                                           _ptr(seamNodes), len(seamNodes), _ptr(results), _ptr(simp))
        return rc, results, simp

    def generateCollisionBatch(self, nodeMinSize, physicsVertices, triangles, seamNodes, vertices=None, unitOptions=None,
                               physicsScale=0.05):
        """Clipmap::loadCollisionNodes' per-node work (clipmap.cpp:1346-1385) over a batch: ConstructCollisionNodeData
        + AddMeshToWorldImpl's conversion (physics.cpp:549-573).  physicsVertices: float32[n][4]; triangles int32[n][3].
        Returns (error, results, simplified)"""
        ms = np.ascontiguousarray(nodeMinSize, np.int32).reshape(-1, 4)
        results = np.zeros(len(ms), ChunkResult)
        simp = np.zeros(len(ms), SimplifyResult)
        opt = unitOptions if unitOptions is not None else SimplifyOptions.clipmap_unit()
        assert vertices is None or len(vertices) >= len(physicsVertices)
        rc = self._L.lvn_meshgen_generate_collision_batch(self.privateCtx_, len(ms), _ptr(ms), C.byref(opt), physicsScale,
                                                          _ptr(physicsVertices), _ptr(vertices), len(physicsVertices),
                                                          _ptr(triangles), len(triangles), _ptr(seamNodes), len(seamNodes),
                                                          _ptr(results), _ptr(simp))
        return rc, results, simp

    def applyCSGOperationsBatch(self, opInfo, chunkMinSize):
        ms = np.ascontiguousarray(chunkMinSize, np.int32).reshape(-1, 4)
        arr = (CSGOperationInfo * len(opInfo))(*opInfo)
        return self._L.lvn_meshgen_apply_csg_operations_batch(self.privateCtx_, arr, len(opInfo), len(ms), _ptr(ms))

    def debugDumpChunk(self, min_, size):
        """every stage of one chunk as numpy arrays (parity tests)"""
        V = self.voxelsPerChunk()
        F = V + 2
        ecap, ncap = 1 << 15, 1 << 15
        while True:
            d = StageDump()
            d.edgeCapacity, d.nodeCapacity = ecap, ncap
            bufs = dict(materials=np.zeros(F ** 3, np.uint8), edgeKeys=np.zeros(ecap, np.int32),
                        edgeInfo=np.zeros((ecap, 4), np.float32), nodeCodes=np.zeros(ncap, np.uint32),
                        nodeEdgeMasks=np.zeros(ncap, np.int32), nodeMaterials=np.zeros(ncap, np.int32),
                        nodeQEFs=np.zeros((ncap, 16), np.float32), nodePositions=np.zeros((ncap, 4), np.float32),
                        nodeNormals=np.zeros((ncap, 4), np.float32))
            for k, v in bufs.items():
                setattr(d, k, v.ctypes.data)
            rc = self._L.lvn_meshgen_debug_dump_chunk(self.privateCtx_, _i3(min_), int(size), C.byref(d))
            if rc == LVN_ERR_CAPACITY:
                ecap, ncap = max(ecap, d.numEdges), max(ncap, d.numNodes)
                continue
            break
        if rc < 0:
            raise RuntimeError(f"debug_dump_chunk: {GetCLErrorString(rc)} {last_cuda_error()}")
        E, N = d.numEdges, d.numNodes
        out = dict(numEdges=E, numNodes=N, numTriangles=d.numTriangles, numSeamNodes=d.numSeamNodes,
                   materials=bufs["materials"])
        for k in ("edgeKeys", "edgeInfo"):
            out[k] = bufs[k][:E].copy()
        for k in ("nodeCodes", "nodeEdgeMasks", "nodeMaterials", "nodeQEFs", "nodePositions", "nodeNormals"):
            out[k] = bufs[k][:N].copy()
        return out

    def setPipeline(self, lanes, streams):
        """lanes (0 = automatic) x streams of a batch call, see lvn_meshgen_set_pipeline"""
        return self._L.lvn_meshgen_set_pipeline(self.privateCtx_, int(lanes), int(streams))

    def getPipeline(self):
        """(lanes of the last batch, configured streams)"""
        a, b = C.c_int(0), C.c_int(0)
        self._L.lvn_meshgen_get_pipeline(self.privateCtx_, C.byref(a), C.byref(b))
        return a.value, b.value

    def setStream(self, cuda_stream):
        """cuda_stream: integer cudaStream_t handle (e.g. torch.cuda.current_stream().cuda_stream) or None"""
        return self._L.lvn_meshgen_set_stream(self.privateCtx_, C.c_void_p(cuda_stream))

    def setProfiling(self, enabled):
        return self._L.lvn_meshgen_set_profiling(self.privateCtx_, int(bool(enabled)))

    def getStats(self, reset=False):
        s = StageStats()
        self._L.lvn_meshgen_get_stats(self.privateCtx_, C.byref(s), int(bool(reset)))
        out = {f: getattr(s, f) for f in ("terrainEvals", "edges", "edgesY", "nodes", "triangles", "seamNodes",
                                          "chunks", "nonEmptyChunks")}
        out["ms"] = {n: s.ms[i] for i, n in enumerate(LVN_STAGES)}
        out["launches"] = {n: s.launches[i] for i, n in enumerate(LVN_STAGES)}
        return out


class PinnedArray:
    """numpy array in pinned host memory (lvn_alloc_pinned) for the arenas of generateBatch.
    Keep the object alive while the array is in use."""

    def __init__(self, n, dtype):
        dtype = np.dtype(dtype)
        self._L = lib()
        self.nbytes = max(int(n), 1) * dtype.itemsize
        self.ptr = self._L.lvn_alloc_pinned(self.nbytes)
        if not self.ptr:
            raise MemoryError("lvn_alloc_pinned failed")
        buf = (C.c_uint8 * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=max(int(n), 1))[:int(n)]

    def close(self):
        if self.ptr:
            self.array = None
            self._L.lvn_free_pinned(C.c_void_p(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def MeasureFP32Peak():
    """measured non-tensor FP32 peak in TFLOP/s (roofline denominator of the FP32-bound stages)"""
    v = C.c_double(0.0)
    rc = lib().lvn_measure_fp32_peak(C.byref(v))
    if rc < 0:
        raise RuntimeError(GetCLErrorString(rc))
    return v.value


# ---- utilities of the path (compute.cpp:328-543, compute_cuckoo.h) -------------
def GlobalMeshOffsets(gathered, world_size, per_rank, num_chunks, counts, offsets, totals):
    """lvn_global_mesh_offsets into caller-owned int64 arrays counts[n, 3], offsets[n, 3], totals[3]"""
    return lib().lvn_global_mesh_offsets(_ptr(gathered), int(world_size), int(per_rank), int(num_chunks),
                                         _ptr(counts), _ptr(offsets), _ptr(totals))


def DebugSolveQEFs(qefs16, packed=True):
    """lvn_debug_solve_qefs: qefs16 float32[n, 16] (QEFData layout) -> positions float32[n, 4]"""
    q = np.ascontiguousarray(qefs16, np.float32).reshape(-1, 16)
    out = np.zeros((len(q), 4), np.float32)
    rc = lib().lvn_debug_solve_qefs(int(bool(packed)), len(q), _ptr(q), _ptr(out))
    assert rc == 0, GetCLErrorString(rc)
    return out


def FindNextPrime(n):
    return lib().lvn_find_next_prime(int(n))


def ExclusiveScan(data):
    """returns (total, scan)"""
    data = np.ascontiguousarray(data, np.int32)
    scan = np.zeros_like(data)
    total = lib().lvn_exclusive_scan(_ptr(data), _ptr(scan), len(data))
    return total, scan


def CompactIndexArray(values, valid):
    values = np.ascontiguousarray(values, np.int32)
    valid = np.ascontiguousarray(valid, np.int32)
    out = np.zeros_like(values)
    n = lib().lvn_compact_index_array(_ptr(values), _ptr(valid), len(values), _ptr(out))
    if n < 0:
        raise RuntimeError(GetCLErrorString(n))
    return out[:n].copy()


def RemoveDuplicates(values):
    values = np.ascontiguousarray(values, np.int32)
    out = np.zeros_like(values)
    n = lib().lvn_remove_duplicates(_ptr(values), len(values), _ptr(out))
    if n < 0:
        raise RuntimeError(GetCLErrorString(n))
    return out[:n].copy()


class CuckooData:
    """compute_cuckoo.h:12-24"""

    def __init__(self):
        self.h = None

    def Cuckoo_InitialiseTable(self, tableSize):
        self.h = lib().lvn_cuckoo_create(int(tableSize))
        return LVN_SUCCESS if self.h else LVN_CL_ERROR

    def Cuckoo_InsertKeys(self, keys):
        keys = np.ascontiguousarray(keys, np.uint32)
        return lib().lvn_cuckoo_insert_keys(self.h, _ptr(keys), len(keys))

    def Cuckoo_Find(self, keys):
        keys = np.ascontiguousarray(keys, np.uint32)
        values = np.zeros_like(keys)
        rc = lib().lvn_cuckoo_find(self.h, _ptr(keys), len(keys), _ptr(values))
        if rc < 0:
            raise RuntimeError(GetCLErrorString(rc))
        return values

    @property
    def prime(self):
        return lib().lvn_cuckoo_prime(self.h)

    @property
    def retries(self):
        return lib().lvn_cuckoo_retries(self.h)

    def destroy(self):
        if self.h:
            lib().lvn_cuckoo_destroy(self.h)
            self.h = None


# ---------------------------------------------------------------------------------------------
# seam meshes between clipmap nodes (GenerateClipmapSeamMesh, clipmap.cpp:573-611)
# ---------------------------------------------------------------------------------------------
SeamNeighbour = np.dtype([("index", np.int32), ("min", np.int32, 3), ("size", np.int32), ("firstNode", np.int32),
                          ("numNodes", np.int32), ("pad", np.int32)])
SeamJob = np.dtype([("hostMin", np.int32, 3), ("hostSize", np.int32), ("firstNeighbour", np.int32), ("numNeighbours", np.int32),
                    ("colour", np.float32, 3), ("pad", np.int32)])
SeamResult = np.dtype([("numVertices", np.int32), ("numTriangles", np.int32), ("vertexOffset", np.int32),
                       ("triangleOffset", np.int32), ("numSelectedNodes", np.int32), ("status", np.int32)])


def PackSeamJobs(seams, colour=(1.0, 1.0, 1.0)):
    """seams: [(hostMin, hostSize, [(neighbourIndex, neighbourMin, neighbourSize, SeamNodeInfo array), ...]), ...]
    -> the three flat arrays of lvn_seam_mesh_generate_batch (jobs, neighbours, seam nodes)"""
    jobs = np.zeros(len(seams), SeamJob)
    nbs, nodes = [], []
    first_node = 0
    placed = {}     # a node's SeamNodeInfo array is uploaded once, however many seams it feeds
    for s, (host_min, host_size, neighbours) in enumerate(seams):
        jobs[s]["hostMin"] = host_min
        jobs[s]["hostSize"] = host_size
        jobs[s]["firstNeighbour"] = len(nbs)
        jobs[s]["numNeighbours"] = len(neighbours)
        jobs[s]["colour"] = colour
        for index, nb_min, nb_size, arr in neighbours:
            k = (tuple(nb_min), nb_size, id(arr))
            if k not in placed:
                a = np.ascontiguousarray(arr, SeamNodeInfo)
                placed[k] = (first_node, len(a))
                nodes.append(a)
                first_node += len(a)
            nbs.append((index, list(nb_min), nb_size, placed[k][0], placed[k][1], 0))
    nb_arr = np.array(nbs, SeamNeighbour) if nbs else np.zeros(0, SeamNeighbour)
    node_arr = np.concatenate(nodes) if nodes else np.zeros(0, SeamNodeInfo)
    return jobs, nb_arr, node_arr


def GenerateClipmapSeamMeshesPacked(voxelsPerChunk, jobs, nb_arr, node_arr, V=None, T=None):
    """one lvn_seam_mesh_generate_batch call on packed arrays -> (rc, V, T, results)"""
    cand = int(nb_arr["numNodes"].sum()) if len(nb_arr) else 0    # a node can be a vertex of several seams
    vcap, tcap = max(cand, 1), max(8 * cand, 1)
    V = np.zeros(vcap, MeshVertex) if V is None else V
    T = np.zeros(tcap, MeshTriangle) if T is None else T
    res = np.zeros(len(jobs), SeamResult)
    rc = lib().lvn_seam_mesh_generate_batch(int(voxelsPerChunk), len(jobs), _ptr(jobs), _ptr(nb_arr), len(nb_arr), _ptr(node_arr),
                                            len(node_arr), _ptr(V), len(V), _ptr(T), len(T), _ptr(res))
    return rc, V, T, res


def GenerateClipmapSeamMeshes(voxelsPerChunk, seams, colour=(1.0, 1.0, 1.0)):
    """GenerateClipmapSeamMesh (clipmap.cpp:573-611) for a list of host nodes
    -> (rc, [(vertices MeshVertex[], triangles MeshTriangle[]) per seam], results)"""
    jobs, nb_arr, node_arr = PackSeamJobs(seams, colour)
    rc, V, T, res = GenerateClipmapSeamMeshesPacked(voxelsPerChunk, jobs, nb_arr, node_arr)
    meshes = [(V[r["vertexOffset"]:r["vertexOffset"] + r["numVertices"]].copy(),
               T[r["triangleOffset"]:r["triangleOffset"] + r["numTriangles"]].copy()) for r in res]
    return rc, meshes, res


# ---------------------------------------------------------------------------------------------
# mesh simplification (ngMeshSimplifier, ng_mesh_simplify.cpp:441-540)
# ---------------------------------------------------------------------------------------------
SimplifyJob = np.dtype([("vertexOffset", np.int32), ("numVertices", np.int32), ("triangleOffset", np.int32),
                        ("numTriangles", np.int32), ("worldSpaceOffset", np.float32, 4)])
SimplifyResult = np.dtype([("numVertices", np.int32), ("numTriangles", np.int32), ("iterations", np.int32), ("numEdges", np.int32)])


class SimplifyOptions(C.Structure):
    """MeshSimplificationOptions, ng_mesh_simplify.h:6-28"""
    _fields_ = [("edgeFraction", C.c_float), ("maxIterations", C.c_int32), ("targetPercentage", C.c_float),
                ("maxError", C.c_float), ("maxEdgeSize", C.c_float), ("minAngleCosine", C.c_float)]

    @classmethod
    def for_clipmap_node(cls, node_size):
        """ConstructClipmapNodeData, clipmap.cpp:449-465 (options.h:14-16)"""
        leaf = float(4 * (node_size // 256))
        return cls(0.125, 10, 0.05, 5.0 * leaf, 2.5 * leaf, 0.7)


    @classmethod
    def clipmap_unit(cls):
        """Options::meshMaxError_ / meshMaxEdgeLen_ / meshMinCosAngle_ (options.h:14-16), before the leaf-size scaling"""
        return cls(0.125, 10, 0.05, 5.0, 2.5, 0.7)

    @classmethod
    def make(cls, **kw):
        """the struct's defaults (ng_mesh_simplify.h:6-28) with overrides"""
        d = dict(edgeFraction=0.125, maxIterations=10, targetPercentage=0.05, maxError=5.0, maxEdgeSize=2.5, minAngleCosine=0.8)
        d.update(kw)
        return cls(d["edgeFraction"], d["maxIterations"], d["targetPercentage"], d["maxError"], d["maxEdgeSize"], d["minAngleCosine"])


def PackSimplifyMeshes(meshes):
    """[(vertices, triangles MeshTriangle[] or int[n][3], worldSpaceOffset xyz)] -> (jobs SimplifyJob[],
    V MeshVertex[], T MeshTriangle[]): the packed arrays lvn_mesh_simplify_batch works on in place"""
    jobs = np.zeros(len(meshes), SimplifyJob)
    vo = to = 0
    for m, (v, t, off) in enumerate(meshes):
        nt = len(t) if isinstance(t, np.ndarray) and t.dtype == MeshTriangle else len(np.asarray(t).reshape(-1)) // 3
        jobs[m] = (vo, len(v), to, nt, list(off)[:3] + [0.0])
        vo += len(v); to += nt
    V = np.zeros(max(vo, 1), MeshVertex)
    T = np.zeros(max(to, 1), MeshTriangle)
    for j, (v, t, off) in zip(jobs, meshes):
        for f in ("xyz", "normal", "colour"):
            V[f][j["vertexOffset"]:j["vertexOffset"] + j["numVertices"]] = v[f]
        tt = t["indices_"] if isinstance(t, np.ndarray) and t.dtype == MeshTriangle else np.asarray(t, np.int32).reshape(-1, 3)
        T["indices_"][j["triangleOffset"]:j["triangleOffset"] + j["numTriangles"]] = tt
    return jobs, V, T


def ngMeshSimplifierPacked(jobs, options, V, T):
    """lvn_mesh_simplify_batch on packed arrays, in place -> (rc, results SimplifyResult[]);
    options: one SimplifyOptions for all, or a list with one per mesh"""
    res = np.zeros(len(jobs), SimplifyResult)
    if isinstance(options, SimplifyOptions):
        oarr, nopt = (SimplifyOptions * 1)(options), 1
    else:
        assert len(options) == len(jobs)
        oarr, nopt = (SimplifyOptions * max(len(options), 1))(*options), len(options)
    nv = int(jobs["vertexOffset"][-1] + jobs["numVertices"][-1]) if len(jobs) else 0
    nt = int(jobs["triangleOffset"][-1] + jobs["numTriangles"][-1]) if len(jobs) else 0
    rc = lib().lvn_mesh_simplify_batch(len(jobs), _ptr(jobs), oarr, nopt, _ptr(V), nv, _ptr(T), nt, _ptr(res))
    return rc, res


def ngMeshSimplifierBatch(meshes, options):
    """ngMeshSimplifier (ng_mesh_simplify.cpp:441-540) over many meshes in one launch.
    meshes: [(vertices MeshVertex[], triangles MeshTriangle[] or int[n][3], worldSpaceOffset xyz)];
    options: one SimplifyOptions for all, or a list with one per mesh (they scale with the node size)
    -> (rc, [(vertices, triangles) simplified], results)"""
    jobs, V, T = PackSimplifyMeshes(meshes)
    rc, res = ngMeshSimplifierPacked(jobs, options, V, T)
    out = [(V[j["vertexOffset"]:j["vertexOffset"] + r["numVertices"]].copy(),
            T[j["triangleOffset"]:j["triangleOffset"] + r["numTriangles"]].copy()) for j, r in zip(jobs, res)]
    return rc, out, res


# ---------------------------------------------------------------------------------------------
# one Clipmap::update as two batched passes (clipmap.cpp:1253-1340)
# ---------------------------------------------------------------------------------------------
ClipmapNode = np.dtype([("min", np.int32, 3), ("size", np.int32), ("firstSeamNode", np.int32), ("numSeamNodes", np.int32)])


class ClipmapUpdateTotals(C.Structure):
    _fields_ = [("nodeVertices", C.c_int64), ("nodeTriangles", C.c_int64), ("seamVertices", C.c_int64), ("seamTriangles", C.c_int64),
                ("seamNodesUsed", C.c_int64), ("numConstructedActive", C.c_int32), ("numSeamUpdates", C.c_int32)]


def ClipmapUpdateBatch(ctx, nodes, numActive, seamNodes, seamNodesUsed, vertices, triangles, unitOptions=None, colour=(1.0, 1.0, 1.0)):
    """lvn_clipmap_update_batch: nodes = ClipmapNode[numActive + numConstruct] (updated in place);
    -> (rc, constructResults, seamUpdateNodes, seamResults, totals)"""
    n_construct = len(nodes) - numActive
    cres = np.zeros(max(n_construct, 1), ChunkResult)
    upd = np.zeros(max(len(nodes), 1), np.int32)
    sres = np.zeros(max(len(nodes), 1), SeamResult)
    tot = ClipmapUpdateTotals()
    opt = unitOptions if unitOptions is not None else SimplifyOptions.clipmap_unit()
    col = (C.c_float * 3)(*colour)
    rc = lib().lvn_clipmap_update_batch(ctx.privateCtx_, _ptr(nodes), int(numActive), int(n_construct), C.byref(opt),
                                        _ptr(seamNodes), int(seamNodesUsed), len(seamNodes), _ptr(vertices), len(vertices),
                                        _ptr(triangles), len(triangles), _ptr(cres), _ptr(upd), _ptr(sres), col, C.byref(tot))
    return rc, cres[:n_construct], upd[:tot.numSeamUpdates], sres[:tot.numSeamUpdates], tot


def ClipmapSeamUpdateBatch(voxelsPerChunk, nodes, active, constructed, seamNodes, numSeamNodes, vertices, triangles,
                           shardIndex=0, shardCount=1, colour=(1.0, 1.0, 1.0)):
    """lvn_clipmap_seam_update_batch: pass 2 of an update for one shard of the seam-update set.
    seamNodes: a numpy SeamNodeInfo array, or an integer device address (the all-gathered arena).
    -> (rc, seamUpdateNodes, seamResults, numSeamUpdatesAll)"""
    active = np.ascontiguousarray(active, np.int32); constructed = np.ascontiguousarray(constructed, np.int32)
    upd = np.zeros(max(len(nodes), 1), np.int32)
    sres = np.zeros(max(len(nodes), 1), SeamResult)
    n_all, n_mine = C.c_int32(0), C.c_int32(0)
    col = (C.c_float * 3)(*colour)
    sn = C.c_void_p(int(seamNodes)) if isinstance(seamNodes, int) else _ptr(seamNodes)
    rc = lib().lvn_clipmap_seam_update_batch(int(voxelsPerChunk), _ptr(nodes), len(nodes), _ptr(active), len(active), _ptr(constructed),
                                             len(constructed), sn, int(numSeamNodes), int(shardIndex), int(shardCount),
                                             _ptr(vertices), len(vertices), _ptr(triangles), len(triangles), _ptr(upd), _ptr(sres), col,
                                             C.byref(n_all), C.byref(n_mine))
    return rc, upd[:n_mine.value], sres[:n_mine.value], n_all.value
