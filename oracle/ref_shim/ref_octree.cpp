// ref_octree.cpp -- the reference's seam-octree code (leven/src/octree.cpp: Octree_ConstructUpwards,
// GenerateVertexIndices, ContourCellProc / FaceProc / EdgeProc / ProcessEdge, Octree_GenerateMesh),
// compiled for the host from where it lies, behind one plain C entry point.
//
// TEST INFRASTRUCTURE (oracle/_ref): the checker of the GPU seam-mesh path.
// octree.cpp is compiled unmodified (ref_shim/translate.py only drops the #include lines of
// headers that need the renderer / window system); GLM, which the reference does not vendor, is
// stood in for by ref_shim/miniglm.  What this file supplies in place of the dropped headers:
//   LVN_ASSERT / LVN_ALWAYS_ASSERT   force_include.h:21-27 (MSVC __debugbreak -> abort)
//   ChunkMinForPosition              volume.cpp:30-41, restated: p & ~(CLIPMAP_LEAF_SIZE - 1)
//   Render_AllocMeshBuffer / Free    render.h: a pool allocator there, new / delete here
//   SlabAllocator<T, N>              slab_allocator.h re-declares its template parameter as a member
//                                    (accepted by MSVC only): same interface, std::deque storage
// LEVEN is defined as in the reference's project file (MAX_MESH_VERTICES = 14 * 1024).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#include <climits>
#include <stdint.h>

#include <glm/glm.hpp>

#define LVN_ASSERT(x) { if (!(x)) { std::fprintf(stderr, "LVN_ASSERT failed: %s\n", #x); std::abort(); } }
#define LVN_ALWAYS_ASSERT(msg, x) { if (!(x)) { std::fprintf(stderr, "%s\n", msg); std::abort(); } }

#include <deque>
#define HAS_SLAB_ALLOCATOR_BEEN_INCLUDED
template <typename T, int N> class SlabAllocator {
public:
    T *alloc() { store_.emplace_back(); return &store_.back(); }
    void clear() { store_.clear(); }
    size_t size() const { return store_.size(); }
private:
    std::deque<T> store_;
};

#include "render_types.h"
#include "volume_constants.h"

using glm::ivec3;
const ivec3 ChunkMinForPosition(const ivec3 &p)
{
    const unsigned int mask = ~(CLIPMAP_LEAF_SIZE - 1);
    return ivec3(p.x & mask, p.y & mask, p.z & mask);
}
const ivec3 ChunkMinForPosition(const int x, const int y, const int z)
{
    const unsigned int mask = ~(CLIPMAP_LEAF_SIZE - 1);
    return ivec3(x & mask, y & mask, z & mask);
}
static MeshBuffer *Render_AllocMeshBuffer(const char *tag) { MeshBuffer *b = new MeshBuffer; b->tag = tag; return b; }
static void Render_FreeMeshBuffer(MeshBuffer *b) { delete b; }

#include "octree.cpp.inc"

extern "C" {

// One seam octree: `n` leaf nodes (min xyz + size, world units; position / averageNormal / the
// material word, as GenerateMeshDataForNode fills them from SeamNodeInfo, clipmap.cpp:398-415)
// -> Octree_ConstructUpwards(rootMin, rootSize) -> Octree_GenerateMesh(colour).
// Returns the triangle count (0: no mesh), or -1 when the buffers are too small.
int ref_seam_octree_mesh(int n, const int *minSize, const float *positions3, const float *normals3, const int *materialInfo,
                         const int *rootMin, int rootSize, const float *colour3,
                         float *vertices12, int vertexCap, int *numVertices, int *triangles3, int triangleCap)
{
    *numVertices = 0;
    if (n <= 0) return 0;
    std::vector<OctreeNode> nodes(n);
    std::vector<OctreeDrawInfo> infos(n);
    std::vector<OctreeNode *> input(n);
    for (int i = 0; i < n; i++) {
        OctreeNode &node = nodes[i];
        node.size = minSize[4 * i + 3];
        node.min = ivec3(minSize[4 * i], minSize[4 * i + 1], minSize[4 * i + 2]);
        node.type = Node_Leaf;
        node.drawInfo = &infos[i];
        infos[i].position = glm::vec3(positions3[3 * i], positions3[3 * i + 1], positions3[3 * i + 2]);
        infos[i].averageNormal = glm::vec3(normals3[3 * i], normals3[3 * i + 1], normals3[3 * i + 2]);
        infos[i].materialInfo = materialInfo[i];
        input[i] = &node;
    }
    Octree octree;
    OctreeNode *root = Octree_ConstructUpwards(&octree, input, ivec3(rootMin[0], rootMin[1], rootMin[2]), rootSize);
    MeshBuffer *mesh = Octree_GenerateMesh(root, glm::vec3(colour3[0], colour3[1], colour3[2]));
    if (!mesh) return 0;
    int rc = mesh->numTriangles;
    if (mesh->numVertices > vertexCap || mesh->numTriangles > triangleCap) rc = -1;
    else {
        *numVertices = mesh->numVertices;
        std::memcpy(vertices12, mesh->vertices, sizeof(MeshVertex) * mesh->numVertices);
        std::memcpy(triangles3, mesh->triangles, sizeof(MeshTriangle) * mesh->numTriangles);
    }
    Render_FreeMeshBuffer(mesh);
    return rc;
}

}  // extern "C"
