// Internal types shared by the CUDA translation units of libleven_b200.so.
// Nothing here crosses the C ABI (include/leven_compute.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/leven_compute.h"

namespace lvn {

// One requested chunk of a batch.  Filled on the host, read by every stage kernel.
struct ChunkDesc {
    int ox, oy, oz;        // field offset in voxel units = min / LEAF_SIZE_SCALE (compute.cpp:569-577)
    int scale;             // sampleScale = size / (V * 4) (compute_density_field.cpp:149)
    int minx, miny, minz;  // world-space chunk min (SolveQEFs' worldSpaceOffset, compute_octree.cpp:131)
    int size;
    int colSet;            // SRC_HEIGHTS: index of this chunk's column-height set
    int source;            // SRC_*
    int edgeMode;          // EDGES_*
    int cachedNumEdges;    // EDGES_CACHED
    const uint8_t *field;  // SRC_FIELD: u8 materials, F^3
    const float *latticeDensity;     // a 3-D density field made in this batch: its F^3 density values (k_field_density),
                                     // i.e. steps 0 and 16 of every edge's zero-crossing search; else null
    const int *cachedKeys;           // EDGES_CACHED: the field's edge list (arbitrary order)
    const float4 *cachedInfo;
    const unsigned long long *cuckooTable;   // edge key -> slot in cachedKeys (a9)
    unsigned int cuckooPrime;
    unsigned int cuckooParams[8];
    float colour[3];       // ColourForMinLeafSize(size / 256) (clipmap.cpp:329-352)
    int pad;
};

enum { SRC_HEIGHTS = 0, SRC_FIELD = 1 };
enum { EDGES_FRESH = 0, EDGES_CACHED = 1 };

// Per-chunk counts and arena placement, written by the classify kernel.
struct ChunkHdr {
    int E, N, Q, S;                          // edges, nodes (= vertices), quads, seam nodes
    int edgeBase, nodeBase, quadBase, seamBase;
    int status;                              // 0, or LVN_ERR_CAPACITY when an arena overflowed
    int Ey;                                  // y-axis edges among E (stage accounting)
    int pad[2];
};

struct ArenaCounters {
    unsigned int edges, nodes, quads, seams;
    unsigned int overflow;
    unsigned int nonEmpty;
    unsigned int edgeTiles, nodeTiles;   // entries of the tile directories
    unsigned int xzEdges;                // entries of the lane's x/z-edge search list (k_hermite_locate)
    unsigned int candidates;             // chunks of the lane that can contain surface (k_candidates): k_rows' list
};

// Tile directories: the Hermite and leaf kernels run one block per tile of LVN_TILE consecutive
// edges / nodes of one chunk, so their grids are flat over the surface of the whole lane and
// empty chunks cost nothing.
struct TileRef { int chunk, first; };
constexpr int LVN_TILE = 128;
// edges per block of the Hermite kernels.  Measured (ring): 128 edges 183.4 us, 256 edges 186.9 us -- although
// 256 edges fill the 256-thread block's search rounds better (720 items = 2.8 rounds against 360 = 1.4), the
// other resident blocks fill those gaps already and the larger tile only lengthens the tail.
#ifndef LVN_ETILE
#define LVN_ETILE 128
#endif

struct ArenaCaps { unsigned int edges, nodes, quads, seams; };

// one lane's slices of the arenas, tile directories and counters
struct LaneArenas {
    ArenaCaps caps, base;             // capacity of the slice and where it starts
    ArenaCounters *ctr;
    TileRef *edgeTiles, *nodeTiles;   // the lane's directories
    unsigned int tileCap;             // entries per directory
};

// Geometry of a mesh-generation context (compute.cpp:245-252,271).
struct Dims {
    int V, H, F;        // voxels per chunk, Hermite grid (V+1), field samples (V+2)
    int shift, mask;    // VOXEL_INDEX_SHIFT / MASK
    int depth;          // MAX_OCTREE_DEPTH
};

// k_rows works per z layer (one warp each); row offsets are relative to the layer, the layer
// records hold the layer's exclusive base.
constexpr int LVN_MAX_LAYERS = 66;   // F for V = 64

// Per-chunk scratch that links k_rows to the Hermite and leaf kernels.
struct ChunkScratch {
    unsigned long long *bitsLo;   // [n][F*F] solid bits x 0..63 of row (z*F + y)
    unsigned int *bitsHi;         // [n][F*F] solid bits x 64..
    unsigned int *rowE;           // [n][H*H] layer-relative exclusive edge offsets per Hermite row (z*H + y)
    unsigned int *rowN;           // [n][V*V] layer-relative exclusive node offsets per voxel row (z*V + y)
    unsigned int *rowQ;           // [n][V*V] quads
    unsigned int *rowS;           // [n][V*V] seam nodes
    uint4 *layer;                 // [n][LVN_MAX_LAYERS] exclusive (edge, node, quad, seam) base of each z layer
    unsigned int *layerEy;        // [n][LVN_MAX_LAYERS] y edges of each layer (stage accounting)
    unsigned int *ticket;         // [n] layers finished; zero between batches
};

struct DensityParams {
    const float2 *grad2;      // [257*257] snoise2 gradients (texel.xy * 4 - 1), wrap-padded
    const float *grad2x;      // [2][257*257] the same table as separate x and y planes (density.cuh: GradTables)
    const float4 *grad3;      // [256*256] snoise3 gradients xyz, w = bits of the perm column; then 65536 bytes: the columns alone
    int kind;                 // 0 terrain, 1 stress
    float param;              // stress threshold
    int defaultMaterial;
    float negZero;            // -0.f: the addend of every packed product (density.cuh, terrain_height_x2)
};

// Optional per-node stage outputs (parity dumps and the octree cache).
struct NodeDebug {
    unsigned int *codes;
    int *edgeMasks;
    int *matWords;
    float *qefs;       // 16 floats per node
    float4 *positions;
    float4 *normals;
};

// ---- launchers (kernels_chunk.cu) ----------------------------------------
// colMin / colMax: per column set, the ordered-int keys of the smallest / largest height
// (initialised to 0x7f7f7f7f / 0x80808080 before launch_columns: they arrive with the batch head)
void launch_columns(const DensityParams &dp, const Dims &d, const int4 *colSetOrigins, int numColSets,
                    float *heights, int *colMin, int *colMax, cudaStream_t s);
void launch_field_from_heights(const Dims &d, const ChunkDesc *descs, int n, const float *heights,
                               int defaultMaterial, uint8_t *const *fields, cudaStream_t s);
void launch_field_density(const DensityParams &dp, const Dims &d, const ChunkDesc *descs, int n,
                          uint8_t *const *fields, cudaStream_t s);
// the chunk kernels work on descs[first .. first + n): one lane of a batch (api.cu: run_batch)
// hostHdrs / hostCounters: mapped pinned mirrors written directly by the kernels
void launch_rows(const Dims &d, const ChunkDesc *descs, int first, int n, const float *heights,
                 const int *colMin, const int *colMax, ChunkHdr *hdrs, ChunkHdr *hostHdrs, ChunkScratch ws,
                 LaneArenas lane, int *candidateList /* [n chunks of the batch] */, cudaStream_t s);
// hostHdrs / hostCounters may be null: launch_publish then mirrors the lane's header slots
void launch_publish(const ChunkHdr *devHdrs, ChunkHdr *hostHdrs, int count, cudaStream_t s, bool evenIfEmpty = false);
void launch_hermite(const DensityParams &dp, const Dims &d, const ChunkDesc *descs, const ChunkHdr *hdrs,
                    ChunkScratch ws, LaneArenas lane, const float *heights, int *edgeKeys, float4 *edgeInfo,
                    int2 *xzList, cudaStream_t s);
void launch_leaves(const DensityParams &dp, const Dims &d, const ChunkDesc *descs, const ChunkHdr *hdrs,
                   ChunkScratch ws, LaneArenas lane, ArenaCounters *hostCounters, const float4 *edgeInfo,
                   void *qefScratch,   // 64 B per slot of the vertex arena: k_leaves -> k_solve
                   lvn_mesh_vertex *vertices, int *triIndices, lvn_seam_node_info *seams,
                   NodeDebug dbg, cudaStream_t s);
void launch_solve(const ChunkDesc *descs, LaneArenas lane, const void *qefScratch, lvn_mesh_vertex *vertices,
                  lvn_seam_node_info *seams, float4 *dbgPositions, cudaStream_t s);

void launch_solve_debug(int packed, int n, const float *qef16, float4 *out, cudaStream_t s);

// ---- launchers (kernels_csg.cu) -------------------------------------------
struct CsgOpDev {          // CSGOperation + host-computed cos/sin of rotateY
    int type, shape, material, pad;
    float ox, oy, oz, dx, dy, dz, c, s;
};
// one chunk of a batched CSG edit (device-side parameters, blockIdx.y)
struct CsgChunk {
    int ox, oy, oz, scale;        // field offset and sample scale (ChunkDesc)
    uint8_t *field;               // F^3 materials, updated in place
    unsigned int *touched;        // bitmap over the 3*H^3 Hermite edges: incident to a changed sample
    const int *oldKeys;           // the field's edge list before the edit
    const float4 *oldInfo;
    int numOld;
    int *newKeys;                 // emit pass: kept edges, then the created ones
    float4 *newInfo;
    int numKept;
    unsigned int *counts;         // [0] kept [1] created [2] changed samples [4] kept cursor [5] created cursor
    int opFirst, numOps;          // this chunk's slice of the op array
    int skip;                     // emit pass: nothing changed in this chunk
};
// ticket: one zeroed word; hostCounts: device address of the mapped mirror that receives the 8 counters of every chunk
void launch_csg_materials_count(const Dims &d, const CsgChunk *chunks, int n, const CsgOpDev *ops, unsigned int *ticket,
                                unsigned int *hostCounts, cudaStream_t s);
void launch_csg_emit(const Dims &d, const CsgChunk *chunks, int n, const CsgOpDev *ops, cudaStream_t s);

// ---- launchers (kernels_util.cu) -------------------------------------------
int  host_find_next_prime(int n);
void launch_fill_u64(unsigned long long *p, size_t n, unsigned long long v, cudaStream_t s);
void launch_cuckoo_insert(const unsigned int *keys, unsigned int count, unsigned long long *table,
                          unsigned int prime, const unsigned int *params8, unsigned int *failed, cudaStream_t s);
// Cuckoo_InitialiseTable + Cuckoo_InsertKeys of up to LVN_TABLE_JOBS tables in two launches (kernels_util.cu)
struct TableJob {
    const unsigned int *keys;
    unsigned long long *table;
    unsigned int *failed;          // set to 1 when the eviction chain of a key did not end (device or mapped host memory)
    unsigned int count, prime;
    unsigned int p[8];             // a0 b0 a1 b1 a2 b2 a3 b3
};
constexpr int LVN_TABLE_JOBS = 32;
struct TableJobs { TableJob job[LVN_TABLE_JOBS]; };
void launch_table_builds(const TableJobs &jobs, int numJobs, cudaStream_t s);
void launch_cuckoo_find(const unsigned int *keys, unsigned int count, const unsigned long long *table,
                        unsigned int prime, const unsigned int *params8, unsigned int *values, cudaStream_t s);
// device-wide exclusive scan of ints; total written to *total (device)
void launch_exclusive_scan(const int *data, int *scan, int count, int *blockSums, int *total, cudaStream_t s);
int  scan_block_sums_needed(int count);
void launch_compact(const int *values, const int *valid, const int *scan, int count, int *out, cudaStream_t s);
void launch_fma_peak(float *sink, int iters, int blocks, cudaStream_t s);
void launch_dedupe(const int *values, int count, unsigned int *table, unsigned int tableSize,
                   int *out, unsigned int *outCount, cudaStream_t s);

// ---- simplify.cu: ngMeshSimplifier on device-resident meshes ----------------
struct SimplifyMesh {
    int vertexOffset, numVertices;       // the mesh's slices of the arrays below
    int triangleOffset, numTriangles;
    float offset[4];                     // worldSpaceOffset
    lvn_simplify_options opt;
};
// in place, asynchronous on `st`; d_results[m] = (vertices, triangles, iterations, edges left);
// iterations -1: a wild triangle index, -2: too large; both leave the mesh untouched.
// With d_packT the simplified meshes are also gathered densely in mesh order (d_packV: MeshVertex,
// d_packP: the physics engine's vec4 = (xyz - offset) * physicsScale; either may be null).
// split (optional, packing only): the smaller meshes run and are packed on a second stream, at the
// front of the packed arrays, while the largest ones are still being simplified (simplify.cu).
struct SimplifySplit {
    cudaStream_t streamB;        // in: the second stream
    cudaEvent_t evFork, evEarly; // in: two events to order the streams with
    int2 *d_earlyTotals;         // in: device word that receives the end of the early region
    int numEarly;                // out: meshes in the early group; 0 = one launch, mesh order
};
int simplify_device(int n, const SimplifyMesh *meshes, lvn_mesh_vertex *d_V, int *d_T, int4 *d_results,
                    lvn_mesh_vertex *d_packV, int *d_packT, int2 *d_packOffsets, int2 *d_packTotals, cudaStream_t st,
                    float4 *d_packP = nullptr, float physicsScale = 0.f, SimplifySplit *split = nullptr);
const char *simplify_last_error();

// ---- api.cu: the fused chunk + simplifier batch, for clipmap_update.cu ----
int generate_simplified(lvn_meshgen *ctx, int nChunks, const int32_t *chunkMinSize, const lvn_simplify_options *unitOptions,
                        lvn_mesh_vertex *vertices, float *physicsVertices, float physicsScale, int64_t vertexCapacity,
                        lvn_mesh_triangle *triangles, int64_t triangleCapacity, lvn_seam_node_info *seamNodes, int64_t seamCapacity,
                        lvn_chunk_result *results, lvn_simplify_result *simplified, bool deferMeshCopies,
                        uint8_t *hadMesh);   // hadMesh (optional): chunk i had a mesh before the simplifier ran
int meshgen_wait(lvn_meshgen *ctx);

}  // namespace lvn
