// ref_kernels.cpp -- the reference's own OpenCL C kernels (leven/cl/*.cl), compiled for the host
// CPU from where they lie under /root/reference, behind plain C launchers.
//
// TEST INFRASTRUCTURE (oracle/_ref): it pins the C restatement (oracle/lvn_oracle.c) and generates
// the golden vectors under tests/golden/; only tests/, smoke() and bench.py's CPU-baseline legs may
// load the library built from this file.  No reference source is copied into the repo: the build
// (oracle/Makefile, target ref) translates each .cl in place (ref_shim/translate.py: vector-literal
// syntax only) into oracle/_ref/gen/, #includes the result here, and deletes it again.
//
// Programs are assembled exactly as the reference host does (compute.cpp:223-226, 280-283,
// 291-294, 302-303: header files first, then the main file), each in its own namespace because
// shared_constants.cl is part of three of them.  The build options are compute.cpp:208-221 and
// :256-278 for voxelsPerChunk = 64.
//
// An NDRange is a loop nest; get_global_id() reads a thread-local.  Kernels that touch shared
// tables with atomics (Cuckoo_InsertKeys) run serially, in work-item order.
#include <cfloat>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "clc.hpp"

thread_local clc::work_item clc::clc_wi;

// ---- build options (compute.cpp:256-278; V = 64) ----
#ifndef VOXELS_PER_CHUNK
#define VOXELS_PER_CHUNK 64
#endif
#define LEAF_SIZE_SCALE 4                     // volume_constants.h:7-8
#define FIELD_DIM (VOXELS_PER_CHUNK + 2)
#define HERMITE_INDEX_SIZE (VOXELS_PER_CHUNK + 1)
#if VOXELS_PER_CHUNK == 64
#define VOXEL_INDEX_SHIFT 7                   // log2(V) + 1
#define MAX_OCTREE_DEPTH 6
#elif VOXELS_PER_CHUNK == 32
#define VOXEL_INDEX_SHIFT 6
#define MAX_OCTREE_DEPTH 5
#elif VOXELS_PER_CHUNK == 16
#define VOXEL_INDEX_SHIFT 5
#define MAX_OCTREE_DEPTH 4
#else
#error "VOXELS_PER_CHUNK must be 16, 32 or 64"
#endif
#define VOXEL_INDEX_MASK ((1 << VOXEL_INDEX_SHIFT) - 1)
#define MAX_TERRAIN_HEIGHT 900                // "const int maxTerrainHeight = 900.f" streamed as 900
#define MATERIAL_AIR 201                      // volume_materials.h:7-8
#define MATERIAL_NONE 200
#define FIND_EDGE_INFO_STEPS 16
#define FIND_EDGE_INFO_INCREMENT 0.0625       // (1.f/16.f) streamed through operator<<: a double literal
#define CUCKOO_EMPTY_VALUE 18446744073709551615ul   // compute_cuckoo.h:6-10
#define CUCKOO_STASH_HASH_INDEX 4
#define CUCKOO_HASH_FN_COUNT 5
#define CUCKOO_STASH_SIZE 101
#define CUCKOO_MAX_ITERATIONS 32
#define FIELD_BUFFER_SIZE (FIELD_DIM * FIELD_DIM * FIELD_DIM)
#define NUM_CSG_BRUSHES 2

namespace clc {
namespace density_prog {   // compute.cpp:280-283
#include "shared_constants.cl.inc"
#include "simplex.cl.inc"
#include "noise.cl.inc"
#include "density_field.cl.inc"
}
#undef HAS_SHARED_CONSTANTS_CL_BEEN_INCLUDED
namespace octree_prog {    // compute.cpp:291-294
#include "shared_constants.cl.inc"
#include "cuckoo.cl.inc"
#include "qef.cl.inc"
#include "octree.cl.inc"
}
#undef HAS_SHARED_CONSTANTS_CL_BEEN_INCLUDED
namespace csg_prog {       // compute.cpp:302-303
#include "shared_constants.cl.inc"
#include "apply_csg_operation.cl.inc"
}
namespace util_prog {      // compute.cpp:223-226 (scan.cl needs work-group barriers: not built; an
#include "duplicate.cl.inc"   // exclusive prefix sum is what it computes, compute.cpp:328-420)
#include "fill_buffer.cl.inc"
#include "compact.cl.inc"
}
}  // namespace clc

// the qualifier macros must not leak into the launchers below
#undef kernel
#undef global
#undef constant
#undef local

using namespace clc;

template <class F> static void launch3(size_t gx, size_t gy, size_t gz, F f)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (long z = 0; z < (long)gz; z++)
        for (long y = 0; y < (long)gy; y++)
            for (size_t x = 0; x < gx; x++) {
                clc_wi.gid[0] = x; clc_wi.gid[1] = (size_t)y; clc_wi.gid[2] = (size_t)z;
                clc_wi.gsize[0] = gx; clc_wi.gsize[1] = gy; clc_wi.gsize[2] = gz;
                f();
            }
}
template <class F> static void launch1(size_t n, F f)
{
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)n; i++) {
        clc_wi.gid[0] = (size_t)i; clc_wi.gid[1] = clc_wi.gid[2] = 0;
        clc_wi.gsize[0] = n; clc_wi.gsize[1] = clc_wi.gsize[2] = 1;
        f();
    }
}
template <class F> static void launch1_serial(size_t n, F f)
{
    for (size_t i = 0; i < n; i++) {
        clc_wi.gid[0] = i; clc_wi.gid[1] = clc_wi.gid[2] = 0;
        clc_wi.gsize[0] = n; clc_wi.gsize[1] = clc_wi.gsize[2] = 1;
        f();
    }
}

static_assert(sizeof(float4) == 16 && sizeof(int4) == 16, "vector layout");
static_assert(sizeof(octree_prog::QEFData) == 64, "QEFData (qef.cl:7-14)");
static_assert(sizeof(octree_prog::MeshVertex) == 48, "MeshVertex");
static_assert(sizeof(octree_prog::SeamNodeInfo) == 48, "SeamNodeInfo");
static_assert(sizeof(csg_prog::CSGOperation) == 48, "CSGOperation (apply_csg_operation.cl:5-14)");

extern "C" {

int ref_voxels_per_chunk(void) { return VOXELS_PER_CHUNK; }
// threads used for the work-items of every NDRange (a CPU OpenCL runtime would use all cores)
int ref_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// one DensityFunc evaluation (noise.cl:225-268), for spot checks
float ref_DensityFunc(const unsigned char *rgba, float x, float y, float z)
{
    const image2d img = {rgba, 256, 256};
    return density_prog::DensityFunc(float4(x, y, z, 0.f), &img);
}
float ref_snoise2(const unsigned char *rgba, float x, float y)
{
    const image2d img = {rgba, 256, 256};
    return density_prog::snoise2(float2(x, y), &img);
}
float ref_snoise3(const unsigned char *rgba, float x, float y, float z)
{
    const image2d img = {rgba, 256, 256};
    return density_prog::snoise3(float3(x, y, z), &img);
}

// ---- density_field.cl ----
void ref_GenerateDefaultField(const unsigned char *rgba, const int *offset4, int sampleScale, int defaultMaterial, int *field)
{
    const image2d img = {rgba, 256, 256};
    const int4 off(offset4[0], offset4[1], offset4[2], offset4[3]);
    launch3(FIELD_DIM, FIELD_DIM, FIELD_DIM, [&] { density_prog::GenerateDefaultField(&img, off, sampleScale, defaultMaterial, field); });
}
void ref_FindFieldEdges(const int *offset4, int *materials, int *edgeOccupancy, int *edgeIndices)
{
    const int4 off(offset4[0], offset4[1], offset4[2], offset4[3]);
    launch3(HERMITE_INDEX_SIZE, HERMITE_INDEX_SIZE, HERMITE_INDEX_SIZE,
            [&] { density_prog::FindFieldEdges(off, materials, edgeOccupancy, edgeIndices); });
}
void ref_CompactEdges(int n, int *edgeValid, int *edgeScan, int *edges, int *compact)
{
    launch1(n, [&] { density_prog::CompactEdges(edgeValid, edgeScan, edges, compact); });
}
void ref_FindEdgeIntersectionInfo(const unsigned char *rgba, const int *offset4, int sampleScale, int n, int *encodedEdges, float *edgeInfo)
{
    const image2d img = {rgba, 256, 256};
    const int4 off(offset4[0], offset4[1], offset4[2], offset4[3]);
    launch1(n, [&] { density_prog::FindEdgeIntersectionInfo(&img, off, sampleScale, encodedEdges, (float4 *)edgeInfo); });
}

// ---- cuckoo.cl ----  (stash must hold `prime` entries: the reference indexes it mod prime, cuckoo.cl:67-69)
void ref_Cuckoo_InsertKeys(int n, uint *keys, ulong *data, ulong *stash, uint prime, uint *hashParams, int *inserted, int *stashUsed)
{
    launch1_serial(n, [&] { octree_prog::Cuckoo_InsertKeys(keys, data, stash, prime, hashParams, inserted, stashUsed); });
}
uint ref_Cuckoo_Find(uint key, ulong *data, ulong *stash, uint prime, uint *hashParams, int stashUsed)
{
    return octree_prog::Cuckoo_Find(key, data, stash, prime, hashParams, stashUsed);
}
uint ref_Cuckoo_Hash(int whichHash, uint key, uint a, uint b, uint prime) { return octree_prog::Cuckoo_Hash(whichHash, key, a, b, prime); }

// ---- octree.cl ----
void ref_FindActiveVoxels(int *materials, int *voxelOccupancy, int *voxelEdgeInfo, int *voxelPositions, int *voxelMaterials)
{
    launch3(VOXELS_PER_CHUNK, VOXELS_PER_CHUNK, VOXELS_PER_CHUNK,
            [&] { octree_prog::FindActiveVoxels(materials, voxelOccupancy, voxelEdgeInfo, voxelPositions, voxelMaterials); });
}
void ref_CompactVoxels(int n, int *valid, int *edgeInfo, int *positions, int *materials, int *scan, int *cPositions, int *cEdgeInfo, int *cMaterials)
{
    launch1(n, [&] { octree_prog::CompactVoxels(valid, edgeInfo, positions, materials, scan, cPositions, cEdgeInfo, cMaterials); });
}
void ref_CreateLeafNodes(int n, int sampleScale, int *voxelPositions, int *voxelEdgeInfo, float *edgeDataTable, float *vertexNormals,
                         void *leafQEFs, ulong *table, ulong *stash, uint prime, uint *hashParams, int checkStash)
{
    launch1(n, [&] {
        octree_prog::CreateLeafNodes(sampleScale, voxelPositions, voxelEdgeInfo, (float4 *)edgeDataTable, (float4 *)vertexNormals,
                                     (octree_prog::QEFData *)leafQEFs, table, stash, prime, hashParams, checkStash);
    });
}
void ref_SolveQEFs(int n, const float *worldSpaceOffset4, void *qefs, float *solved)
{
    const float4 off(worldSpaceOffset4[0], worldSpaceOffset4[1], worldSpaceOffset4[2], worldSpaceOffset4[3]);
    launch1(n, [&] { octree_prog::SolveQEFs(off, (octree_prog::QEFData *)qefs, (float4 *)solved); });
}
void ref_GenerateMesh(int n, uint *codes, int *materials, int *meshIndexBuffer, int *trianglesValid, ulong *table, ulong *stash,
                      uint prime, uint *hashParams, int checkStash)
{
    launch1(n, [&] { octree_prog::GenerateMesh(codes, materials, meshIndexBuffer, trianglesValid, table, stash, prime, hashParams, checkStash); });
}
void ref_CompactMeshTriangles(int n, int *valid, int *scan, int *meshIndexBuffer, int *compact)
{
    launch1(n, [&] { octree_prog::CompactMeshTriangles(valid, scan, meshIndexBuffer, compact); });
}
void ref_GenerateMeshVertexBuffer(int n, float *positions, float *normals, int *materials, const float *colour4, void *meshVertices)
{
    const float4 colour(colour4[0], colour4[1], colour4[2], colour4[3]);
    launch1(n, [&] {
        octree_prog::GenerateMeshVertexBuffer((float4 *)positions, (float4 *)normals, materials, colour, (octree_prog::MeshVertex *)meshVertices);
    });
}
void ref_FindSeamNodes(int n, uint *codes, int *isSeamNode)
{
    launch1(n, [&] { octree_prog::FindSeamNodes(codes, isSeamNode); });
}
void ref_ExtractSeamNodeInfo(int n, int *isSeamNode, int *scan, uint *codes, int *materials, float *positions, float *normals, void *out)
{
    launch1(n, [&] {
        octree_prog::ExtractSeamNodeInfo(isSeamNode, scan, codes, materials, (float4 *)positions, (float4 *)normals, (octree_prog::SeamNodeInfo *)out);
    });
}

// ---- apply_csg_operation.cl ----
void ref_CSG_HermiteIndices(const int *offset4, int numOps, const void *ops, int sampleScale, const int *fieldMaterials,
                            int *updatedIndices, int *updatedPositions, int *updatedMaterials)
{
    const int4 off(offset4[0], offset4[1], offset4[2], offset4[3]);
    launch3(FIELD_DIM, FIELD_DIM, FIELD_DIM, [&] {
        csg_prog::CSG_HermiteIndices(off, numOps, (const csg_prog::CSGOperation *)ops, sampleScale, fieldMaterials, updatedIndices,
                                     (int4 *)updatedPositions, updatedMaterials);
    });
}
void ref_CSG_FindUpdatedEdges(int n, int *updatedPositions4, int *updatedEdgeIndices)
{
    launch1(n, [&] { csg_prog::FindUpdatedEdges((int4 *)updatedPositions4, updatedEdgeIndices); });
}
void ref_CSG_FilterValidEdges(int n, int *generatedEdgeIndices, int *materials, int *edgeValid)
{
    launch1(n, [&] { csg_prog::FilterValidEdges(generatedEdgeIndices, materials, edgeValid); });
}
void ref_CSG_FindEdgeIntersectionInfo(const int *offset4, int numOps, const void *ops, int sampleScale, int n, const int *compactEdges, float *normals)
{
    const int4 off(offset4[0], offset4[1], offset4[2], offset4[3]);
    launch1(n, [&] {
        csg_prog::FindEdgeIntersectionInfo(off, numOps, (const csg_prog::CSGOperation *)ops, sampleScale, compactEdges, (float4 *)normals);
    });
}

}  // extern "C"
