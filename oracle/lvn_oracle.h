/*
 * lvn_oracle.h -- CPU ORACLE for the leven chunk-meshing hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under leven_b200/ may include, link or
 * call this.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker or as
 * the reported CPU baseline -- never as the thing shipped.
 *
 * What it is: a plain-C restatement, statement by statement, of the
 * reference's OpenCL kernels (the .cl files under leven/cl) and of the host orchestration
 * in leven/src/compute*.cpp.  Each function cites the reference file:line
 * it follows.  The arithmetic is the one fixed by DESIGN.md "arithmetic
 * spec": IEEE binary32, round-to-nearest, no contraction except the explicit
 * fmaf() calls in snoise2/snoise3's dot products.
 *
 * PARITY PIN STATUS
 *   - cuckoo hash + dedupe: PINNED by the reference's own fixtures
 *     (leven/src/testdata/octree_keys_*.cpp, test_cuckoo.cpp:120-178,
 *     test_compute.cpp:46-88) -> tests/golden/octree_keys.npz.
 *   - density, Hermite data, active voxels, QEF, mesh topology, seam nodes,
 *     CSG: PARITY UNPINNED.  The reference holds no test, golden file or
 *     known-answer vector for them, and its kernels
 *     cannot be executed here (OpenCL C, no ICD, MSVC-only host code).  The
 *     restatement is cross-checked by invariants only (tests/test_oracle.py).
 */
#ifndef LVN_ORACLE_H
#define LVN_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LVO_MATERIAL_NONE 200   /* leven/src/volume_materials.h:7 */
#define LVO_MATERIAL_AIR  201   /* leven/src/volume_materials.h:8 */
#define LVO_LEAF_SIZE_SCALE 4   /* leven/src/volume_constants.h:7-8 */

/* density function selector.  0 = the reference's live DensityFunc
 * (noise.cl:225-268, #else branch).  1 = BASELINE config 4 "dense stress":
 * ridged 3-D fBm built from the reference's snoise3 (simplex.cl:159-230). */
#define LVO_DENSITY_TERRAIN 0
#define LVO_DENSITY_STRESS  1

typedef struct { float x, y, z, w; } lvo_f4;
typedef struct { int32_t x, y, z, w; } lvo_i4;

/* qef.cl:7-14, 64 bytes */
typedef struct {
    float  ATA[6];
    float  pad[2];
    lvo_f4 ATb;
    lvo_f4 masspoint;
} lvo_qef;

/* leven/src/compute.h:16-24 == apply_csg_operation.cl:5-14, 48 bytes */
typedef struct {
    int32_t type;          /* 0 add, 1 subtract */
    int32_t brushShape;    /* 0 cube, 1 sphere (render_types.h:11-20) */
    int32_t material;
    float   rotateY;       /* radians (hg_sdf.glsl:460-463) */
    float   origin[4];
    float   dimensions[4];
} lvo_csg_op;

/* leven/src/compute.h:26-31, 48 bytes */
typedef struct {
    lvo_i4 localspaceMin;  /* w = material word */
    lvo_f4 position;
    lvo_f4 normal;
} lvo_seam_node;

/* leven/src/render_types.h:24-39, 48 bytes */
typedef struct { lvo_f4 xyz, normal, colour; } lvo_vertex;

/* ---- a1: noise table -------------------------------------------------- */
/* NoiseHash + pixel packing follow compute_density_field.cpp:69-113.  The
 * shuffle of perm[512] is OUR documented generator (mt19937(seed),
 * Fisher-Yates from the top, j = next() % (i+1)) because the reference's
 * std::shuffle(std::default_random_engine) is implementation-defined. */
void     lvo_noise_image(int seed, uint8_t *rgba /* 256*256*4 */);
uint32_t lvo_noise_hash(int x, int y, int seed);

/* ---- a2: density ------------------------------------------------------ */
typedef struct lvo_world lvo_world;   /* global compute state + caches */

lvo_world *lvo_world_create(const uint8_t *rgba, int defaultMaterial, int voxelsPerChunk);
void       lvo_world_destroy(lvo_world *);
void       lvo_world_set_density(lvo_world *, int kind, float stressThreshold);

float lvo_snoise2(const lvo_world *, float x, float y);
float lvo_snoise3(const lvo_world *, float x, float y, float z);
float lvo_terrain(const lvo_world *, float x, float z);
float lvo_density(const lvo_world *, float x, float y, float z);

/* ---- a3..a5: field, edges, Hermite ------------------------------------ */
void lvo_generate_field(const lvo_world *, const int min[3], int size, int32_t *materials /* F^3 */);
int  lvo_find_edges(const lvo_world *, const int32_t *materials, int32_t *edgeKeys /* cap 3*H^3 */);
void lvo_edge_info(const lvo_world *, const int min[3], int size,
                   const int32_t *edgeKeys, int numEdges, lvo_f4 *edgeInfo);

/* ---- a6..a8, a10: voxels, leaves, QEF --------------------------------- */
int  lvo_find_active_voxels(const lvo_world *, const int32_t *materials,
                            uint32_t *codes, int32_t *edgeMasks, int32_t *matWords /* cap V^3 each */);
int  lvo_create_leaf_nodes(const lvo_world *, int sampleScale,
                           const uint32_t *codes, const int32_t *edgeMasks, int numNodes,
                           const int32_t *edgeKeys, const lvo_f4 *edgeInfo, int numEdges,
                           lvo_qef *qefs, lvo_f4 *normals);
void lvo_solve_qefs(const int min[3], const lvo_qef *qefs, int numNodes, lvo_f4 *positions);

/* ---- a11, a12, a14: mesh + seams --------------------------------------- */
int  lvo_generate_mesh(const lvo_world *, const uint32_t *codes, const int32_t *matWords, int numNodes,
                       int32_t *indices /* cap 18*numNodes */);   /* returns numTriangles */
void lvo_vertex_buffer(const lvo_f4 *positions, const lvo_f4 *normals, const int32_t *matWords,
                       int numNodes, int size, lvo_vertex *out);
int  lvo_seam_nodes(const lvo_world *, const uint32_t *codes, const int32_t *matWords,
                    const lvo_f4 *positions, const lvo_f4 *normals, int numNodes, lvo_seam_node *out);
void lvo_colour_for_size(int size, float rgb[3]);

/* ---- a9: cuckoo (cuckoo.cl + compute_cuckoo.cpp) ----------------------- */
typedef struct {
    uint64_t *table;
    uint64_t  stash[101];
    uint32_t  prime;
    uint32_t  params[10];
    int       stashUsed;
    int       insertedKeys;
    int       retries;
} lvo_cuckoo;
int      lvo_find_next_prime(int n);                                  /* primes.cpp:32-59 */
int      lvo_cuckoo_init(lvo_cuckoo *, uint32_t tableSize);           /* compute_cuckoo.cpp:49-74 */
int      lvo_cuckoo_insert_keys(lvo_cuckoo *, const uint32_t *keys, uint32_t count); /* :78-138 */
uint32_t lvo_cuckoo_find(const lvo_cuckoo *, uint32_t key);           /* cuckoo.cl:73-104 */
void     lvo_cuckoo_free(lvo_cuckoo *);
/* the CPU table of leven/src/cuckoo.h (64-bit key*a), used by test_cuckoo.cpp */
typedef struct lvo_cpu_cuckoo lvo_cpu_cuckoo;
lvo_cpu_cuckoo *lvo_cpu_cuckoo_create(int size, uint32_t seed);
int             lvo_cpu_cuckoo_insert(lvo_cpu_cuckoo *, uint32_t key, uint32_t value);
int             lvo_cpu_cuckoo_find(const lvo_cpu_cuckoo *, uint32_t key, uint32_t *value);
void            lvo_cpu_cuckoo_destroy(lvo_cpu_cuckoo *);

/* ---- a15: scan / compact / dedupe -------------------------------------- */
int lvo_exclusive_scan(const int32_t *data, int32_t *scan, int count);          /* compute.cpp:384-395 */
int lvo_compact(const int32_t *values, const int32_t *valid, int count, int32_t *out);
int lvo_remove_duplicates(const int32_t *values, int count, int32_t *out);      /* compute.cpp:446-543 */

/* ---- a16: CSG ----------------------------------------------------------- */
float lvo_brush_density(float x, float y, float z, const lvo_csg_op *op);

/* ---- a17 + host orchestration: same call surface as compute.h:35-72 ---- */
int lvo_store_csg_operation(lvo_world *, const lvo_csg_op *op, const int aabbMin[3], const int aabbMax[3]);
int lvo_clear_csg_operations(lvo_world *);
int lvo_apply_csg_operations(lvo_world *, const lvo_csg_op *ops, int numOps, const int min[3], int size);
int lvo_free_chunk_octree(lvo_world *, const int min[3], int size);
int lvo_is_chunk_empty(lvo_world *, const int min[3], int size, int *isEmpty);

/* per-chunk dump of every stage (caller frees with lvo_chunk_free) */
typedef struct {
    int numEdges, numNodes, numTriangles, numSeamNodes;
    int32_t  *materials;   /* F^3 */
    int32_t  *edgeKeys;    /* E   */
    lvo_f4   *edgeInfo;    /* E   */
    uint32_t *codes;       /* N   */
    int32_t  *edgeMasks;   /* N   */
    int32_t  *matWords;    /* N   */
    lvo_qef  *qefs;        /* N   */
    lvo_f4   *positions;   /* N   */
    lvo_f4   *normals;     /* N   */
    lvo_vertex *vertices;  /* N   */
    int32_t  *indices;     /* 3*T */
    lvo_seam_node *seams;  /* S   */
} lvo_chunk;
/* Compute_GenerateChunkMesh (compute_octree.cpp:351-375) incl. both caches */
int  lvo_generate_chunk_mesh(lvo_world *, const int min[3], int size, lvo_chunk *out);
void lvo_chunk_free(lvo_chunk *);

/* batch helper for the CPU baseline: n independent chunks, OpenMP over
 * chunks, no caches; counts[4*i..] = E, N, T, S.  Returns threads used. */
int lvo_set_num_threads(int n);   /* OpenMP threads of the batch helper; returns the count in effect */
int lvo_generate_batch_counts(const lvo_world *, int n, const int *minSize /* 4n */, int32_t *counts);

#ifdef __cplusplus
}
#endif
#endif
