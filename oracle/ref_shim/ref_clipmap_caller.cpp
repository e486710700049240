// ref_clipmap_caller.cpp -- the reference's OWN caller of the compute interface, compiled from where it
// lies against this repository's drop-in boundary: leven/src/clipmap.cpp:329-504 (ColourForMinLeafSize,
// GenerateMeshDataForNode, ConstructClipmapNodeData, ConstructCollisionNodeData) and the struct it needs
// from :86-100 (ClipmapCollisionNode), verbatim (ref_shim/translate.py --lines), with
//     #include "compute.h" + "ng_mesh_simplify.h"   ->   include/leven_compute.hpp
// and linked against leven_b200/lib/libleven_b200.so.  This is the drop-in proof of SURVEY.md 8(b): the
// very code that requests chunk meshes in the application (Clipmap::update -> ConstructClipmapNodeData,
// clipmap.cpp:1259; loadCollisionNodes -> ConstructCollisionNodeData, :1369) runs unmodified on top of
// the C ABI.  tests/test_ref_caller_gpu.py drives it and compares with the direct C-ABI calls.
//
// TEST INFRASTRUCTURE (oracle/_ref; built only where /root/reference exists).  What this file supplies
// for the headers that cannot be included here (renderer, physics, profiler):
//   rmt_ScopedCPUSample            Remotery.h: a profiler scope -> nothing
//   LVN_ASSERT                     force_include.h:21-27 (MSVC __debugbreak -> abort)
//   SlabAllocator<T, N>            slab_allocator.h (octree.h includes it; MSVC-only template syntax)
//   RenderMesh, Render_Alloc*      render.h / render_mesh.h: GL vertex buffers there; here the RenderMesh
//                                  keeps the MeshBuffer it was made from so that the test can read it
//   class Frustum                  frustum.h: only named in Clipmap's declarations (clipmap.h)
// The reference's own render_types.h, aabb.h, volume_constants.h, octree.h and clipmap.h (minus its
// frustum.h / physics.h includes) are used as they are; GLM is ref_shim/miniglm.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <memory>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#include <stdint.h>

#include <glm/glm.hpp>

#define LVN_ASSERT(x) { if (!(x)) { std::fprintf(stderr, "LVN_ASSERT failed: %s\n", #x); std::abort(); } }
#define LVN_ALWAYS_ASSERT(msg, x) { if (!(x)) { std::fprintf(stderr, "%s\n", msg); std::abort(); } }
#define rmt_ScopedCPUSample(name)

#define HAS_SLAB_ALLOCATOR_BEEN_INCLUDED
template <typename T, int N> class SlabAllocator {
public:
    T *alloc() { store_.emplace_back(); return &store_.back(); }
    void clear() { store_.clear(); }
    size_t size() const { return store_.size(); }
private:
    std::deque<T> store_;
};

// the reference's own headers, where they lie (-I $(REF)/src)
#include "render_types.h"
#include "aabb.h"
#include "volume_constants.h"
#include "octree.h"

class Frustum;
#include "clipmap.h.inc"          // clipmap.h without its frustum.h / physics.h includes

// the drop-in boundary: replaces compute.h and ng_mesh_simplify.h
#include "leven_compute.hpp"

using glm::ivec3;
using glm::ivec4;
using glm::vec3;
using glm::vec4;

class RenderMesh {
public:
    MeshBuffer *buffer = nullptr;
    glm::vec3 position;
};
static MeshBuffer *Render_AllocMeshBuffer(const char *tag) { MeshBuffer *b = new MeshBuffer; b->tag = tag; return b; }
static void Render_FreeMeshBuffer(MeshBuffer *b) { delete b; }
static RenderMesh *Render_AllocRenderMesh(const char *const, MeshBuffer *buffer, const glm::vec3 &position)
{
    RenderMesh *m = new RenderMesh;
    m->buffer = buffer;
    m->position = position;
    return m;
}

#include "clipmap_slice.cpp.inc"  // clipmap.cpp:86-100 and :329-504, verbatim

extern "C" {

// ConstructClipmapNodeData (collision = 0) or ConstructCollisionNodeData (collision = 1) for one node
// through the reference's own code.  Outputs: the node's (simplified) mesh, its seam nodes as the
// OctreeNodes GenerateMeshDataForNode builds (min xyz + size; position, averageNormal, colour: 3 floats
// each; materialInfo), and ClipmapNode::active_.  Returns 0, or -1 when an output array is too small.
int refcaller_construct_node(int voxelsPerChunk, int collision, const int *min3, int size,
                             float meshMaxError, float meshMaxEdgeLen, float meshMaxAngle,
                             float *vertices12, int vertexCap, int *numVertices, int *triangles3, int triangleCap, int *numTriangles,
                             int *seamMinSize4, float *seamPosNrmCol9, int *seamMaterial, int seamCap, int *numSeamNodes, int *active)
{
    static std::unordered_map<int, Compute_MeshGenContext *> contexts;   // never destroyed, as in the reference
    Compute_MeshGenContext *&meshGen = contexts[voxelsPerChunk];
    if (!meshGen) meshGen = Compute_MeshGenContext::create(voxelsPerChunk);
    *numVertices = *numTriangles = *numSeamNodes = *active = 0;
    MeshBuffer *mesh = nullptr;
    OctreeNode *seamNodes = nullptr;
    int nSeams = 0;
    RenderMesh *renderMesh = nullptr;
    if (collision) {
        LVN_ASSERT(size == COLLISION_NODE_SIZE);
        ClipmapCollisionNode node(ivec3(min3[0], min3[1], min3[2]));
        mesh = ConstructCollisionNodeData(meshGen, &node, meshMaxError, meshMaxEdgeLen, meshMaxAngle);
        seamNodes = node.seamNodes; nSeams = node.numSeamNodes;
        *active = mesh != nullptr || nSeams != 0;
    } else {
        ClipmapNode node;
        node.min_ = ivec3(min3[0], min3[1], min3[2]);
        node.size_ = size;
        if (ConstructClipmapNodeData(meshGen, &node, meshMaxError, meshMaxEdgeLen, meshMaxAngle) != LVN_SUCCESS) return -2;
        renderMesh = node.renderMesh;
        mesh = renderMesh ? renderMesh->buffer : nullptr;
        seamNodes = node.seamNodes; nSeams = node.numSeamNodes;
        *active = node.active_ ? 1 : 0;
    }
    int rc = 0;
    if (mesh) {
        *numVertices = mesh->numVertices; *numTriangles = mesh->numTriangles;
        if (mesh->numVertices > vertexCap || mesh->numTriangles > triangleCap) rc = -1;
        else {
            std::memcpy(vertices12, mesh->vertices, sizeof(MeshVertex) * (size_t)mesh->numVertices);
            std::memcpy(triangles3, mesh->triangles, sizeof(MeshTriangle) * (size_t)mesh->numTriangles);
        }
    }
    *numSeamNodes = nSeams;
    if (nSeams > seamCap) rc = -1;
    else for (int i = 0; i < nSeams; i++) {
        const OctreeNode &n = seamNodes[i];
        seamMinSize4[4 * i] = n.min.x; seamMinSize4[4 * i + 1] = n.min.y; seamMinSize4[4 * i + 2] = n.min.z; seamMinSize4[4 * i + 3] = n.size;
        const OctreeDrawInfo &di = *n.drawInfo;
        float *o = seamPosNrmCol9 + 9 * (size_t)i;
        o[0] = di.position.x; o[1] = di.position.y; o[2] = di.position.z;
        o[3] = di.averageNormal.x; o[4] = di.averageNormal.y; o[5] = di.averageNormal.z;
        o[6] = di.colour.x; o[7] = di.colour.y; o[8] = di.colour.z;
        seamMaterial[i] = di.materialInfo;
    }
    // the application keeps these alive (ReleaseClipmapNodeData frees them later); the test does not
    for (int i = 0; i < nSeams; i++) delete seamNodes[i].drawInfo;
    delete[] seamNodes;
    if (mesh) Render_FreeMeshBuffer(mesh);
    delete renderMesh;
    // every call is a fresh request: the reference would evict the chunk's octree when the node is
    // released (Compute_FreeChunkOctree via ReleaseClipmapNodeData, clipmap.cpp:644-680)
    meshGen->freeChunkOctree(ivec3(min3[0], min3[1], min3[2]), size);
    return rc;
}

}  // extern "C"
