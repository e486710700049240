/*
 * leven_compute.h -- C ABI of the B200-native chunk-meshing path.
 *
 * This is the drop-in boundary: every entry point replaces one symbol of the
 * reference's compute interface, leven/src/compute.h:12-81 (cited per
 * function below; paths are relative to the reference tree).  Plain pointers
 * and sizes only; no C++, CUDA or torch types cross it.  The header-only C++
 * shim include/leven_compute.hpp rebuilds the reference's own
 * Compute_MeshGenContext class on top of it; INTEGRATION.md shows the
 * reference-side change.
 *
 * Conventions (same as the reference):
 *   - return value: 0 = success, negative = failure (LVN_* below; the
 *     reference returns OpenCL codes the same way and its callers test < 0,
 *     clipmap.cpp:379).
 *   - min / size are world units; a voxel is LEAF_SIZE_SCALE (4) world units
 *     at LOD0 (volume_constants.h:7-8); a 64^3 chunk has size 256.
 *   - not thread-safe: one caller at a time per process, as in the reference
 *     (compute.cpp:168-191, volume.cpp:68-113).
 *   - there is no CPU fallback: every call fails with LVN_ERR_NO_DEVICE when
 *     no CUDA device is usable.
 */
#ifndef LEVEN_COMPUTE_H
#define LEVEN_COMPUTE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LVN_SUCCESS              0
#define LVN_CL_ERROR             (-99999)  /* compute.h:12, non-OpenCL failure */
#define LVN_ERR_NO_DEVICE        (-1)      /* == CL_DEVICE_NOT_FOUND */
#define LVN_ERR_OUT_OF_MEMORY    (-4)      /* == CL_MEM_OBJECT_ALLOCATION_FAILURE */
#define LVN_ERR_INVALID_VALUE    (-30)     /* == CL_INVALID_VALUE */
#define LVN_ERR_NOT_INITIALISED  (-34)     /* == CL_INVALID_CONTEXT */
#define LVN_ERR_CAPACITY         (-61)     /* == CL_INVALID_BUFFER_SIZE: caller arena too small
                                              (the reference asserts, compute_octree.cpp:235-236) */
#define LVN_ERR_CUDA             (-9999)   /* a CUDA runtime call failed; see lvn_last_cuda_error */

#define LVN_MATERIAL_NONE 200              /* volume_materials.h:7 */
#define LVN_MATERIAL_AIR  201              /* volume_materials.h:8 */
#define LVN_LEAF_SIZE_SCALE 4              /* volume_constants.h:7-8 */

/* ---- PODs shared with the callers (layouts are the reference's) -------- */

/* CSGOperationInfo, compute.h:16-24 (== CSGOperation, apply_csg_operation.cl:5-14) */
typedef struct lvn_csg_operation_info {
    int32_t type;          /* 0 = add, 1 = subtract */
    int32_t brushShape;    /* RenderShape: 0 = cube, 1 = sphere (render_types.h:11-20) */
    int32_t material;
    float   rotateY;       /* radians */
    float   origin[4];     /* voxel units: world/4 + 0.5 (clipmap.cpp:1628) */
    float   dimensions[4]; /* half extents, voxel units; sphere radius = dimensions[0] */
} lvn_csg_operation_info;

/* SeamNodeInfo, compute.h:26-31 */
typedef struct lvn_seam_node_info {
    int32_t localspaceMin[4];   /* xyz in [0, V); w = (dominantMaterial << 8) | cornerMask */
    float   position[4];
    float   normal[4];
} lvn_seam_node_info;

/* MeshVertex, render_types.h:24-39 */
typedef struct lvn_mesh_vertex { float xyz[4], normal[4], colour[4]; } lvn_mesh_vertex;
/* MeshTriangle, render_types.h:42-58 */
typedef struct lvn_mesh_triangle { int32_t indices_[3]; } lvn_mesh_triangle;
/* AABB, aabb.h:5-96 (data members only) */
typedef struct lvn_aabb { int32_t min[3], max[3]; } lvn_aabb;

/* ---- process-wide state ------------------------------------------------- */

/* new: choose the CUDA device before lvn_compute_initialise (one process per GPU) */
int lvn_compute_set_device(int cudaDevice);
/* new: keep the calling thread and the pinned memory it allocates afterwards on the NUMA node the
 * device is attached to (one process per GPU on a multi-socket host).  Returns the node, negative
 * when the platform names none (nothing changed); *cpusBound = CPUs in the new affinity mask (0 =
 * left alone), *memoryBound = 1 when the preferred-node policy was set.  Optional, off by default. */
int lvn_compute_bind_host_numa(int cudaDevice, int *cpusBound, int *memoryBound);

/* Compute_Initialise, compute.h:35 (compute.cpp:195-234) */
int lvn_compute_initialise(int noiseSeed, unsigned int defaultMaterial, int numCSGBrushes);
/* Compute_Shutdown, compute.h:36 (a no-op in the reference; frees the device state here) */
int lvn_compute_shutdown(void);
/* Compute_SetNoiseSeed, compute.h:38 (compute_density_field.cpp:129-134).  The 512-entry
 * shuffle uses mt19937(seed) + Fisher-Yates (documented in DESIGN.md) because the reference's
 * std::shuffle(std::default_random_engine) is implementation-defined. */
int lvn_compute_set_noise_seed(int noiseSeed);
/* new: feed the exact 256x256 RGBA8 permutation image (row i, column j at ((i*256)+j)*4,
 * compute_density_field.cpp:101-113) produced by any other implementation */
int lvn_compute_set_noise_image(const uint8_t *rgba /* 262144 bytes */);
int lvn_compute_get_noise_image(uint8_t *rgba /* 262144 bytes */);
/* new: density function selector.  0 = the reference terrain (noise.cl:225-268),
 * 1 = BASELINE.json config 4 dense-stress field (ridged fBm over snoise3, simplex.cl:159-230) */
int lvn_compute_set_density_function(int kind, float param);
/* Compute_StoreCSGOperation / Compute_ClearCSGOperations, compute.h:39-40 */
int lvn_compute_store_csg_operation(const lvn_csg_operation_info *op, const lvn_aabb *aabb);
int lvn_compute_clear_csg_operations(void);
/* GetCLErrorString, compute.h:81 */
const char *lvn_error_string(int error);
const char *lvn_last_cuda_error(void);

/* ---- Compute_MeshGenContext, compute.h:46-77 ----------------------------- */

typedef struct lvn_meshgen lvn_meshgen;

/* Compute_MeshGenContext::create, compute.h:50; NULL on failure */
lvn_meshgen *lvn_meshgen_create(int voxelsPerChunk);
/* new: the reference never destroys contexts */
void lvn_meshgen_destroy(lvn_meshgen *ctx);
/* voxelsPerChunk, compute.h:52 */
int lvn_meshgen_voxels_per_chunk(const lvn_meshgen *ctx);

/* applyCSGOperations, compute.h:54-57 (compute_csg.cpp:224-242) */
int lvn_meshgen_apply_csg_operations(lvn_meshgen *ctx, const lvn_csg_operation_info *ops, int numOps,
                                     const int32_t clipmapNodeMin[3], int clipmapNodeSize);
/* freeChunkOctree, compute.h:59-61 (compute_octree.cpp:379-387) */
int lvn_meshgen_free_chunk_octree(lvn_meshgen *ctx, const int32_t min[3], int size);
/* isChunkEmpty, compute.h:63-66.  *isEmpty = 1 when the chunk has no surface crossing (the
 * reference's dead code returns the inverted flag, compute_density_field.cpp:285,296) */
int lvn_meshgen_is_chunk_empty(lvn_meshgen *ctx, const int32_t min[3], int size, int *isEmpty);

/* generateChunkMesh, compute.h:68-72 (compute_octree.cpp:351-375).  Caller-owned buffers
 * (MeshBuffer::vertices/triangles, render_types.h:70-90; the seam std::vector's storage).
 * Returns LVN_ERR_CAPACITY, with the needed counts in *numX, when a buffer is too small. */
int lvn_meshgen_generate_chunk_mesh(lvn_meshgen *ctx, const int32_t min[3], int clipmapNodeSize,
                                    lvn_mesh_vertex *vertices, int vertexCapacity, int *numVertices,
                                    lvn_mesh_triangle *triangles, int triangleCapacity, int *numTriangles,
                                    lvn_seam_node_info *seamNodes, int seamCapacity, int *numSeamNodes);

/* ---- new: batch entry points (the 512 / 4096-chunk configurations) ------- */

typedef struct lvn_chunk_result {
    int32_t numEdges, numVertices, numTriangles, numSeamNodes;
    int32_t vertexOffset, triangleOffset, seamOffset;   /* element offsets into the arenas (the
                                                          * slices of different chunks never overlap;
                                                          * the device arenas may have gaps) */
    int32_t status;                                     /* 0 or a negative LVN_* code */
} lvn_chunk_result;

/* One pass of the whole path over nChunks independent chunks (chunkMinSize = 4 ints per chunk:
 * min.x, min.y, min.z, size).  Same per-chunk results as generateChunkMesh on cold octree
 * caches; does not read or fill the octree cache, does honour CSG-edited density fields.
 * Host arenas (pinned memory recommended); results[i] addresses chunk i's slices. */
int lvn_meshgen_generate_batch(lvn_meshgen *ctx, int nChunks, const int32_t *chunkMinSize,
                               lvn_mesh_vertex *vertices, int64_t vertexCapacity,
                               lvn_mesh_triangle *triangles, int64_t triangleCapacity,
                               lvn_seam_node_info *seamNodes, int64_t seamCapacity,
                               lvn_chunk_result *results);

/* Same pass, results left resident in HBM (arenas owned by ctx, valid until the next call). */
typedef struct lvn_batch_device_view {
    const lvn_mesh_vertex    *vertices;    /* device pointers */
    const lvn_mesh_triangle  *triangles;
    const lvn_seam_node_info *seamNodes;
    int64_t totalVertices, totalTriangles, totalSeamNodes, totalEdges;
    int32_t nonEmptyChunks;
} lvn_batch_device_view;
int lvn_meshgen_generate_batch_device(lvn_meshgen *ctx, int nChunks, const int32_t *chunkMinSize,
                                      lvn_chunk_result *results, lvn_batch_device_view *view);

/* The same, returning as soon as every chunk's counts and arena offsets are final (after the classify kernels of
 * every lane) with the rest of the batch queued on the context's stream: `results` and the view's totals are
 * valid on return, the arenas in stream order on that stream or after lvn_meshgen_wait.  For callers that can
 * use the counts while the meshes are still being made: the count gather of a sharded sweep, the
 * sizing of a consumer's buffers.  A batch that has to grow its arenas or rebuild a hash table completes before
 * the call returns, like lvn_meshgen_generate_batch_device. */
int lvn_meshgen_generate_batch_device_async(lvn_meshgen *ctx, int nChunks, const int32_t *chunkMinSize,
                                            lvn_chunk_result *results, lvn_batch_device_view *view);
/* blocks until everything queued on the context's stream has run */
int lvn_meshgen_wait(lvn_meshgen *ctx);
/* lvn_meshgen_generate_batch the same way: returns once every lane's counts are published and its copies are
 * queued; `results` are final on return, the host arenas complete after lvn_meshgen_wait (the caller must not
 * free or reuse them before) */
int lvn_meshgen_generate_batch_async(lvn_meshgen *ctx, int nChunks, const int32_t *chunkMinSize,
                                     lvn_mesh_vertex *vertices, int64_t vertexCapacity,
                                     lvn_mesh_triangle *triangles, int64_t triangleCapacity,
                                     lvn_seam_node_info *seamNodes, int64_t seamCapacity,
                                     lvn_chunk_result *results);

/* applyCSGOperations over many chunks in one pass (same ops for every chunk) */
int lvn_meshgen_apply_csg_operations_batch(lvn_meshgen *ctx, const lvn_csg_operation_info *ops, int numOps,
                                           int nChunks, const int32_t *chunkMinSize);

/* ---- new: per-stage dump for parity tests -------------------------------- */

typedef struct lvn_stage_dump {
    /* capacities, set by the caller */
    int32_t edgeCapacity, nodeCapacity;
    /* counts, set by the callee */
    int32_t numEdges, numNodes, numTriangles, numSeamNodes;
    /* host buffers, any may be NULL */
    uint8_t  *materials;     /* F^3, index x + F*y + F*F*z (shared_constants.cl:24-27) */
    int32_t  *edgeKeys;      /* E: ((x | y<<s | z<<2s) << 2) | axis (density_field.cl:73) */
    float    *edgeInfo;      /* 4E: normal.xyz, t (density_field.cl:150) */
    uint32_t *nodeCodes;     /* N (octree.cl:34-47) */
    int32_t  *nodeEdgeMasks; /* N (octree.cl:177-191) */
    int32_t  *nodeMaterials; /* N: (dominant << 8) | cornerMask (octree.cl:199-200) */
    float    *nodeQEFs;      /* 16N floats: QEFData (qef.cl:7-14) */
    float    *nodePositions; /* 4N (octree.cl:327-330) */
    float    *nodeNormals;   /* 4N (octree.cl:303-311) */
} lvn_stage_dump;
/* qef_solve (qef.cl:239-256) on caller-supplied QEFData records (16 floats each: ATA[6], pad[2], ATb[4],
 * masspoint[4], qef.cl:7-14), positions in chunk-local units (worldSpaceOffset 0, scale 4): the unit test
 * of k_solve's arithmetic.  packed = 1 runs the production form (two nodes per thread, packed FP32 with
 * written-out division / square-root sequences), packed = 0 the plain scalar form. */
int lvn_debug_solve_qefs(int packed, int n, const float *qefs16, float *positions4);

/* runs the path for one chunk with cold octree cache semantics and copies every stage out */
int lvn_meshgen_debug_dump_chunk(lvn_meshgen *ctx, const int32_t min[3], int size, lvn_stage_dump *dump);

/* ---- new: measurement hooks ---------------------------------------------- */

enum {
    LVN_STAGE_COLUMNS = 0,   /* S1 density: Terrain per column (noise.cl:205-223) */
    LVN_STAGE_CLASSIFY,      /* S2+S4 edge scan, active voxels, prefix sums, compaction */
    LVN_STAGE_HERMITE,       /* S3 FindEdgeIntersectionInfo (density_field.cl:96-151) */
    LVN_STAGE_LEAVES,        /* S5+S8+S9+S10 leaf QEF accumulation + mesh + seams (k_leaves) */
    LVN_STAGE_FIELD,         /* u8 material field materialisation (CSG / 3-D density) */
    LVN_STAGE_CSG,           /* a16 kernels */
    LVN_STAGE_CUCKOO,        /* a9 kernels */
    LVN_STAGE_SOLVE,         /* S6 SolveQEFs (k_solve) */
    LVN_NUM_STAGES
};
typedef struct lvn_stage_stats {
    double  ms[LVN_NUM_STAGES];         /* CUDA-event time on the context's stream, accumulated */
    int64_t launches[LVN_NUM_STAGES];   /* kernel launches, accumulated */
    int64_t terrainEvals;               /* Terrain() evaluations issued, accumulated */
    int64_t edges, edgesY, nodes, triangles, seamNodes, chunks, nonEmptyChunks;
} lvn_stage_stats;
int lvn_meshgen_set_profiling(lvn_meshgen *ctx, int enabled);   /* per-stage events on/off */
/* launch on the caller's CUDA stream (a cudaStream_t passed as void*; NULL = the context's own) */
int lvn_meshgen_set_stream(lvn_meshgen *ctx, void *cudaStream);
/* A batch is cut into `lanes` independent slices (0 = chosen from the batch size) whose kernel
 * chains run on `streams` (1..4) CUDA streams forked from and joined back to the context's
 * stream; the host-path batch call drains finished lanes over the copy engine meanwhile. */
int lvn_meshgen_set_pipeline(lvn_meshgen *ctx, int lanes, int streams);
int lvn_meshgen_get_pipeline(const lvn_meshgen *ctx, int *lanesOfLastBatch, int *streamsOfLastBatch);
/* pinned host memory for the arenas of lvn_meshgen_generate_batch (NULL on failure): the lane
 * copies then run at PCIe rate instead of through the driver's staging buffer */
void *lvn_alloc_pinned(size_t bytes);
void lvn_free_pinned(void *p);
/* FP32 roofline denominator: independent FMA chains on every SM, CUDA-event timed (2 flop/FMA) */
int lvn_measure_fp32_peak(double *tflops);
int lvn_meshgen_get_stats(lvn_meshgen *ctx, lvn_stage_stats *out, int reset);

/* ---- seam meshes between clipmap nodes (the consumer of SeamNodeInfo) ------- */

/* One active clipmap node whose seam nodes may feed a host node's seam: what
 * GenerateClipmapSeamMesh (clipmap.cpp:573-611) hands to SelectSeamNodes (clipmap.cpp:542-569). */
typedef struct lvn_seam_neighbour {
    int32_t index;        /* 0..7: the slot of the host's 2x2x2 neighbourhood the node lies in
                             (CHILD_MIN_OFFSETS, volume_constants.h:24-35; 0 = the host node itself) */
    int32_t min[3];       /* the node's min, world units */
    int32_t size;         /* the node's size, world units (any LOD) */
    int32_t firstNode;    /* its SeamNodeInfo records, as generateChunkMesh returned them: */
    int32_t numNodes;     /* seamNodes[firstNode .. firstNode + numNodes) */
    int32_t pad;
} lvn_seam_neighbour;

typedef struct lvn_seam_job {
    int32_t hostMin[3];   /* ClipmapNode::min_ of the node that owns the seam */
    int32_t hostSize;     /* ClipmapNode::size_ */
    int32_t firstNeighbour, numNeighbours;   /* neighbours[firstNeighbour .. + numNeighbours) */
    float   colour[3];    /* vertex colour (GenerateClipmapSeamMesh's colour argument) */
    int32_t pad;
} lvn_seam_job;

typedef struct lvn_seam_result {
    int32_t numVertices, numTriangles;       /* 0 / 0 when the seam has no triangle (octree.cpp:536-540) */
    int32_t vertexOffset, triangleOffset;    /* into the caller's arenas; indices are seam-local */
    int32_t numSelectedNodes;                /* nodes that passed SelectSeamNodes */
    int32_t status;                          /* 0 or LVN_ERR_CAPACITY */
} lvn_seam_result;

/* GenerateClipmapSeamMesh for many host nodes in one launch: SelectSeamNodes +
 * Octree_ConstructUpwards(hostMin, 2 * hostSize) + Octree_GenerateMesh (octree.cpp:23-148,196-549)
 * as one thread block per seam.  Vertices come out in the reference's order (the DFS order of the
 * seam octree); triangles are the same index triples with the same winding, ordered by (owner
 * vertex, edge) instead of by the reference's recursion.  Host pointers in and out. */
int lvn_seam_mesh_generate_batch(int voxelsPerChunk, int numSeams, const lvn_seam_job *jobs,
                                 const lvn_seam_neighbour *neighbours, int numNeighbours,
                                 const lvn_seam_node_info *seamNodes, int numSeamNodes,
                                 lvn_mesh_vertex *vertices, int64_t vertexCapacity,
                                 lvn_mesh_triangle *triangles, int64_t triangleCapacity,
                                 lvn_seam_result *results);
const char *lvn_seam_last_error(void);

/* ---- mesh simplification (the pass over every chunk mesh right after export) --- */

/* MeshSimplificationOptions, ng_mesh_simplify.h:6-28 */
typedef struct lvn_simplify_options {
    float   edgeFraction;       /* 0.125 */
    int32_t maxIterations;      /* 10 */
    float   targetPercentage;   /* 0.05 */
    float   maxError;           /* 5 * leafSize at the clipmap's call site, clipmap.cpp:460 */
    float   maxEdgeSize;        /* 2.5 * leafSize, clipmap.cpp:461 */
    float   minAngleCosine;     /* options.h:16 = 0.7, clipmap.cpp:462 */
} lvn_simplify_options;

typedef struct lvn_simplify_job {
    int32_t vertexOffset, numVertices;       /* the mesh's slices of the two arrays */
    int32_t triangleOffset, numTriangles;    /* triangle indices are mesh-local */
    float   worldSpaceOffset[4];             /* ngMeshSimplifier's argument: the node centre, w = 0 (clipmap.cpp:451) */
} lvn_simplify_job;

typedef struct lvn_simplify_result { int32_t numVertices, numTriangles, iterations, numEdges; } lvn_simplify_result;

/* ngMeshSimplifier (ng_mesh_simplify.cpp:441-540) on many meshes in one launch, one thread block
 * per mesh, in place: each mesh's simplified vertices / triangles are left at the start of its
 * slices, results[m] holds the new counts.  Same vertices, same triangles, same order as the
 * reference built with libstdc++ (its candidate sampling is std::uniform_int_distribution over
 * std::mt19937(42), which is library-defined) and with _mm_rsqrt_ps taken as 1 / sqrt.
 * A mesh holding a triangle index outside its vertices is passed through untouched with
 * results[m].iterations = -1 and the call returns LVN_ERR_INVALID_VALUE (the other meshes are done). */
int lvn_mesh_simplify_batch(int numMeshes, const lvn_simplify_job *jobs,
                            const lvn_simplify_options *options, int numOptions,   /* 1 (shared) or numMeshes */
                            lvn_mesh_vertex *vertices, int64_t numVerticesTotal,
                            lvn_mesh_triangle *triangles, int64_t numTrianglesTotal,
                            lvn_simplify_result *results);
const char *lvn_mesh_simplify_last_error(void);

/* ConstructClipmapNodeData / ConstructCollisionNodeData (clipmap.cpp:432-504) for many nodes in one
 * pass: generateChunkMesh, then ngMeshSimplifier with worldSpaceOffset = min + size / 2 and
 * options scaled by the node's leaf size (leafSize = LEAF_SIZE_SCALE * (size / CLIPMAP_LEAF_SIZE);
 * maxError = unitOptions->maxError * leafSize, maxEdgeSize likewise; unitOptions carries
 * Options::meshMaxError_ / meshMaxEdgeLen_ / meshMinCosAngle_, options.h:14-16).  The meshes never
 * leave HBM between the two steps; the host arenas receive the SIMPLIFIED meshes, densely packed
 * (in chunk order, except that in a large batch the few largest meshes -- the ones the simplifier
 * works on longest -- come last, so that the others can cross PCIe meanwhile), and the
 * (unsimplified octree's) seam nodes.  results[i] addresses chunk i's slices
 * and holds the simplified counts (numEdges stays the chunk's Hermite edge count); simplified[i]
 * (may be NULL) adds the simplifier's iteration count (-2: the mesh was too large to simplify and
 * is returned as generated).  On LVN_ERR_CAPACITY the counts say what the caller must provide. */
int lvn_meshgen_generate_simplified_batch(lvn_meshgen *ctx, int nChunks, const int32_t *chunkMinSize,
                                          const lvn_simplify_options *unitOptions,
                                          lvn_mesh_vertex *vertices, int64_t vertexCapacity,
                                          lvn_mesh_triangle *triangles, int64_t triangleCapacity,
                                          lvn_seam_node_info *seamNodes, int64_t seamCapacity,
                                          lvn_chunk_result *results, lvn_simplify_result *simplified);

/* Clipmap::loadCollisionNodes' per-node work (clipmap.cpp:1346-1385) for many collision nodes in one
 * pass: ConstructCollisionNodeData (clipmap.cpp:472-504: generateChunkMesh on the physics context at
 * COLLISION_NODE_SIZE + ngMeshSimplifier) and the conversion AddMeshToWorldImpl makes for Bullet's
 * btIndexedMesh (physics.cpp:549-573): physicsVertices[i] = (vertex.xyz - vec4(origin, 0)) * physicsScale,
 * one vec4 (4 floats) per vertex, origin = min + size / 2; triangles = 3 ints each.  (The reference
 * leaves a TODO there: "vertices and triangles should be created on the GPU".)  `vertices` (may be
 * NULL) additionally receives the MeshVertex form the debug renderer is given (physics.cpp:706).
 * Everything else as lvn_meshgen_generate_simplified_batch. */
int lvn_meshgen_generate_collision_batch(lvn_meshgen *ctx, int nNodes, const int32_t *nodeMinSize,
                                         const lvn_simplify_options *unitOptions, float physicsScale /* PHYSICS_SCALE = 0.05f, physics.cpp:79 */,
                                         float *physicsVertices, lvn_mesh_vertex *vertices, int64_t vertexCapacity,
                                         int32_t *triangles, int64_t triangleCapacity,
                                         lvn_seam_node_info *seamNodes, int64_t seamCapacity,
                                         lvn_chunk_result *results, lvn_simplify_result *simplified);

/* ---- one Clipmap::update as two batched passes --------------------------- */

/* A clipmap node as the update sees it: ClipmapNode::min_ / size_ (clipmap.cpp) and where its
 * SeamNodeInfo records (ClipmapNode::seamNodes / numSeamNodes) lie in the caller's seam-node arena. */
typedef struct lvn_clipmap_node {
    int32_t min[3];
    int32_t size;
    int32_t firstSeamNode, numSeamNodes;
} lvn_clipmap_node;

typedef struct lvn_clipmap_update_totals {
    int64_t nodeVertices, nodeTriangles;     /* the constructed nodes' (simplified) meshes: arenas [0, ..) */
    int64_t seamVertices, seamTriangles;     /* the seam meshes: arenas [nodeVertices, ..), [nodeTriangles, ..) */
    int64_t seamNodesUsed;                   /* the seam-node arena's fill after the update */
    int32_t numConstructedActive;            /* constructed nodes with a mesh or seam nodes (the others are empty) */
    int32_t numSeamUpdates;                  /* seam meshes regenerated = entries of seamUpdateNodes / seamResults */
} lvn_clipmap_update_totals;

/* The GPU work of one Clipmap::update (clipmap.cpp:1253-1340) in two batched passes:
 *   nodes[0 .. numActive)                          the nodes active before the update, with their seam-node slices
 *   nodes[numActive .. numActive + numConstruct)   the nodes the update loads ("filteredNodes"); on return
 *                                                  their seam-node slices are filled in (appended to the arena)
 * Pass 1 = ConstructClipmapNodeData for every node to load (lvn_meshgen_generate_simplified_batch;
 * constructResults[i] addresses node numActive + i's mesh).  A node with a mesh or seam nodes
 * becomes active (clipmap.cpp:1269).  Pass 2 = the seam-update set -- every active node found
 * around the 8 cells min - CHILD_MIN_OFFSETS[i] * size of a newly active node (clipmap.cpp:1306-1324)
 * -- and GenerateClipmapSeamMesh (clipmap.cpp:573-611) for all of them in one launch:
 * seamUpdateNodes[u] = index into nodes (ascending), seamResults[u] its seam mesh (offsets into the
 * same two arenas, after the node meshes).  Which nodes exist, are active or get loaded stays
 * the caller's decision.  On LVN_ERR_CAPACITY totals says what pass 1 needs. */
int lvn_clipmap_update_batch(lvn_meshgen *ctx, lvn_clipmap_node *nodes, int numActive, int numConstruct,
                             const lvn_simplify_options *unitOptions,
                             lvn_seam_node_info *seamNodes, int64_t seamNodesUsed, int64_t seamCapacity,
                             lvn_mesh_vertex *vertices, int64_t vertexCapacity,
                             lvn_mesh_triangle *triangles, int64_t triangleCapacity,
                             lvn_chunk_result *constructResults,
                             int32_t *seamUpdateNodes, lvn_seam_result *seamResults, const float seamColour[3],
                             lvn_clipmap_update_totals *totals);

/* Pass 2 of the update on its own, shardable: the seam-update set (clipmap.cpp:1306-1324) over the
 * active nodes nodes[active[..]] around the newly active nodes nodes[constructed[..]], and
 * GenerateClipmapSeamMesh for the share of the set this caller takes -- entry u of the ascending set
 * belongs to shard u % shardCount (seams are independent of each other).  seamNodes may be host or
 * device memory (e.g. the all-gathered seam nodes of every GPU's nodes).  *numSeamUpdatesAll = size
 * of the whole set, *numSeamUpdatesMine = entries of seamUpdateNodes / seamResults filled. */
int lvn_clipmap_seam_update_batch(int voxelsPerChunk, const lvn_clipmap_node *nodes, int numNodes,
                                  const int32_t *active, int numActive, const int32_t *constructed, int numConstructed,
                                  const lvn_seam_node_info *seamNodes, int64_t numSeamNodes, int shardIndex, int shardCount,
                                  lvn_mesh_vertex *vertices, int64_t vertexCapacity,
                                  lvn_mesh_triangle *triangles, int64_t triangleCapacity,
                                  int32_t *seamUpdateNodes, lvn_seam_result *seamResults, const float seamColour[3],
                                  int32_t *numSeamUpdatesAll, int32_t *numSeamUpdatesMine);

/* new: placing the chunks of a sharded batch in one global mesh ("only a final count /
 * size gather where a global mesh is assembled").  gathered = the all-gathered per-chunk counts of a
 * round-robin split, [worldSize][3][perRank] (plane 0 vertices, 1 triangles, 2 seam nodes; chunk i of
 * the linear order sits at rank i % worldSize, slot i / worldSize).  Host arithmetic only (an
 * exclusive prefix sum over numChunks x 3 integers): counts / offsets are [numChunks][3] in linear
 * chunk order, totals[3] the sizes of the assembled arrays. */
int lvn_global_mesh_offsets(const int32_t *gathered, int worldSize, int perRank, int numChunks,
                            int64_t *counts, int64_t *offsets, int64_t totals[3]);

/* ---- utilities of the path (a9, a15), usable on their own ---------------- */

/* FindNextPrime, primes.h (primes.cpp:32-59) */
int lvn_find_next_prime(int n);
/* ExclusiveScan, compute.cpp:384-395: scan[i] = sum(data[0..i)); returns the total */
int lvn_exclusive_scan(const int32_t *data, int32_t *scan, int count);
/* CompactIndexArray, compute.cpp:424-442: stable; returns the compacted count */
int lvn_compact_index_array(const int32_t *values, const int32_t *valid, int count, int32_t *out);
/* RemoveDuplicates, compute.cpp:446-543: out = the distinct values (order unspecified, as in
 * the reference); returns their count */
int lvn_remove_duplicates(const int32_t *values, int count, int32_t *out);

/* CuckooData + Cuckoo_InitialiseTable + Cuckoo_InsertKeys, compute_cuckoo.h:12-24; values are
 * the key's index in the inserted array, as in cuckoo.cl:35 */
typedef struct lvn_cuckoo lvn_cuckoo;
lvn_cuckoo *lvn_cuckoo_create(unsigned int tableSize);
int  lvn_cuckoo_insert_keys(lvn_cuckoo *table, const uint32_t *keys, unsigned int count);
/* Cuckoo_Find, cuckoo.cl:73-104, for an array of keys; ~0u = not found */
int  lvn_cuckoo_find(const lvn_cuckoo *table, const uint32_t *keys, unsigned int count, uint32_t *values);
int  lvn_cuckoo_prime(const lvn_cuckoo *table);
int  lvn_cuckoo_retries(const lvn_cuckoo *table);
void lvn_cuckoo_destroy(lvn_cuckoo *table);

#ifdef __cplusplus
}
#endif
#endif /* LEVEN_COMPUTE_H */
