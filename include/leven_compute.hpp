// leven_compute.hpp -- header-only C++ shim: the reference's own compute interface
// (leven/src/compute.h:12-81) re-created on top of the C ABI in leven_compute.h, so that
// clipmap.cpp / main.cpp / volume.cpp compile against it unchanged.  See INTEGRATION.md.
//
// glm is not required: when <glm/glm.hpp> was included first its ivec3/vec4 are used,
// otherwise minimal stand-ins with the same layout are defined.
#ifndef LEVEN_COMPUTE_HPP
#define LEVEN_COMPUTE_HPP

#include <vector>
#include <stdint.h>

#include "leven_compute.h"

#ifdef GLM_VERSION
namespace lvn_shim { using ivec3 = glm::ivec3; using ivec4 = glm::ivec4; using vec4 = glm::vec4; }
#else
namespace lvn_shim {
struct ivec3 { int x, y, z; ivec3(int a = 0, int b = 0, int c = 0) : x(a), y(b), z(c) {} };
struct ivec4 { int x, y, z, w; };
struct vec4 { float x, y, z, w; };
}
#endif

#ifndef LVN_CL_ERROR
#define LVN_CL_ERROR (-99999)
#endif

#ifndef HAS_RENDER_TYPES_H_BEEN_INCLUDED
// render_types.h:11-20
enum RenderShape { RenderShape_Cube, RenderShape_Sphere, RenderShape_Line, RenderShape_SIZE, RenderShape_None };
#endif

// compute.h:16-24
struct CSGOperationInfo {
    int             type = 0;
    RenderShape     brushShape = RenderShape_Cube;
    int             material = 0;
    float           rotateY = 0.f;
    lvn_shim::vec4  origin;
    lvn_shim::vec4  dimensions;
};

// compute.h:26-31
struct SeamNodeInfo {
    lvn_shim::ivec4 localspaceMin;
    lvn_shim::vec4  position;
    lvn_shim::vec4  normal;
};

#ifndef HAS_RENDER_TYPES_H_BEEN_INCLUDED
// render_types.h:24-90 (LEVEN defined in every build configuration, leven.vcxproj:185)
struct MeshVertex { lvn_shim::vec4 xyz, normal, colour; };
struct MeshTriangle { int indices_[3]; };
const int MAX_MESH_VERTICES = 14 * 1024;
const int MAX_MESH_TRIANGLES = MAX_MESH_VERTICES * 2;
class MeshBuffer {
public:
    MeshBuffer() : tag(nullptr), numVertices(0), numTriangles(0) {}
    const char*  tag;
    MeshVertex   vertices[MAX_MESH_VERTICES];
    int          numVertices;
    MeshTriangle triangles[MAX_MESH_TRIANGLES];
    int          numTriangles;
};
#endif

#ifndef HAS_AABB_BEEN_INCLUDED_H
struct AABB { lvn_shim::ivec3 min, max; };   // aabb.h:95-96 (data members)
#endif

static_assert(sizeof(CSGOperationInfo) == sizeof(lvn_csg_operation_info), "CSGOperationInfo layout");
static_assert(sizeof(SeamNodeInfo) == sizeof(lvn_seam_node_info), "SeamNodeInfo layout");
static_assert(sizeof(MeshVertex) == sizeof(lvn_mesh_vertex), "MeshVertex layout");
static_assert(sizeof(MeshTriangle) == sizeof(lvn_mesh_triangle), "MeshTriangle layout");

// compute.h:35-40
inline int Compute_Initialise(const int noiseSeed, const unsigned int defaultMaterial, const int numCSGBrushes)
{ return lvn_compute_initialise(noiseSeed, defaultMaterial, numCSGBrushes); }
inline int Compute_Shutdown() { return lvn_compute_shutdown(); }
inline int Compute_SetNoiseSeed(const int noiseSeed) { return lvn_compute_set_noise_seed(noiseSeed); }
inline int Compute_StoreCSGOperation(const CSGOperationInfo& opInfo, const AABB& aabb)
{
    lvn_aabb bb = {{aabb.min.x, aabb.min.y, aabb.min.z}, {aabb.max.x, aabb.max.y, aabb.max.z}};
    return lvn_compute_store_csg_operation(reinterpret_cast<const lvn_csg_operation_info*>(&opInfo), &bb);
}
inline int Compute_ClearCSGOperations() { return lvn_compute_clear_csg_operations(); }

// compute.h:46-77
class Compute_MeshGenContext
{
public:
    static Compute_MeshGenContext* create(const int voxelsPerChunk)
    {
        Compute_MeshGenContext* ctx = new Compute_MeshGenContext;
        ctx->privateCtx_ = lvn_meshgen_create(voxelsPerChunk);   // may be null, as in compute.cpp:605-610
        return ctx;
    }

    int voxelsPerChunk() const { return lvn_meshgen_voxels_per_chunk(privateCtx_); }

    int applyCSGOperations(const std::vector<CSGOperationInfo>& opInfo,
                           const lvn_shim::ivec3& clipmapNodeMin, const int clipmapNodeSize)
    {
        const int32_t mn[3] = { clipmapNodeMin.x, clipmapNodeMin.y, clipmapNodeMin.z };
        return lvn_meshgen_apply_csg_operations(privateCtx_,
            reinterpret_cast<const lvn_csg_operation_info*>(opInfo.data()), (int)opInfo.size(), mn, clipmapNodeSize);
    }

    int freeChunkOctree(const lvn_shim::ivec3& min, const int size)
    {
        const int32_t mn[3] = { min.x, min.y, min.z };
        return lvn_meshgen_free_chunk_octree(privateCtx_, mn, size);
    }

    int isChunkEmpty(const lvn_shim::ivec3& min, const int size, bool& isEmpty)
    {
        const int32_t mn[3] = { min.x, min.y, min.z };
        int e = 0;
        const int rc = lvn_meshgen_is_chunk_empty(privateCtx_, mn, size, &e);
        isEmpty = e != 0;
        return rc;
    }

    int generateChunkMesh(const lvn_shim::ivec3& min, const int clipmapNodeSize,
                          MeshBuffer* meshBuffer, std::vector<SeamNodeInfo>& seamNodeBuffer)
    {
        const int32_t mn[3] = { min.x, min.y, min.z };
        seamNodeBuffer.clear();                       // compute_octree.cpp:359
        seamNodeBuffer.resize(4096);
        int nV = 0, nT = 0, nS = 0;
        int rc = LVN_ERR_CAPACITY;
        for (int attempt = 0; attempt < 2 && rc == LVN_ERR_CAPACITY; attempt++) {
            rc = lvn_meshgen_generate_chunk_mesh(privateCtx_, mn, clipmapNodeSize,
                reinterpret_cast<lvn_mesh_vertex*>(meshBuffer->vertices), MAX_MESH_VERTICES, &nV,
                reinterpret_cast<lvn_mesh_triangle*>(meshBuffer->triangles), MAX_MESH_TRIANGLES, &nT,
                reinterpret_cast<lvn_seam_node_info*>(seamNodeBuffer.data()), (int)seamNodeBuffer.size(), &nS);
            if (rc == LVN_ERR_CAPACITY && nS > (int)seamNodeBuffer.size()) seamNodeBuffer.resize(nS);
            else break;
        }
        if (rc < 0) { seamNodeBuffer.clear(); return rc; }
        meshBuffer->numVertices = nV;
        meshBuffer->numTriangles = nT;
        seamNodeBuffer.resize(nS);
        return rc;
    }

private:
    lvn_meshgen* privateCtx_;
};

// compute.h:81
inline const char* GetCLErrorString(int error) { return lvn_error_string(error); }

// ng_mesh_simplify.h:6-36: the simplifier clipmap.cpp runs on every exported mesh (clipmap.cpp:449-465,495-501)
struct MeshSimplificationOptions
{
    float edgeFraction = 0.125f;
    int maxIterations = 10;
    float targetPercentage = 0.05f;
    float maxError = 5.f;
    float maxEdgeSize = 2.5f;
    float minAngleCosine = 0.8f;
};

inline void ngMeshSimplifier(MeshBuffer* mesh, const lvn_shim::vec4& worldSpaceOffset, const MeshSimplificationOptions& options)
{
    lvn_simplify_job job = { 0, mesh->numVertices, 0, mesh->numTriangles,
                             { worldSpaceOffset.x, worldSpaceOffset.y, worldSpaceOffset.z, worldSpaceOffset.w } };
    const lvn_simplify_options opt = { options.edgeFraction, options.maxIterations, options.targetPercentage,
                                       options.maxError, options.maxEdgeSize, options.minAngleCosine };
    lvn_simplify_result r = { mesh->numVertices, mesh->numTriangles, 0, 0 };
    if (lvn_mesh_simplify_batch(1, &job, &opt, 1, reinterpret_cast<lvn_mesh_vertex*>(mesh->vertices), mesh->numVertices,
                                reinterpret_cast<lvn_mesh_triangle*>(mesh->triangles), mesh->numTriangles, &r) < 0)
        return;                                   // the mesh is left as it was (the reference returns void)
    mesh->numVertices = r.numVertices;
    mesh->numTriangles = r.numTriangles;
}

#endif  // LEVEN_COMPUTE_HPP
