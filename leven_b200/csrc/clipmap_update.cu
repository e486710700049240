// clipmap_update.cu -- the GPU work of one Clipmap::update (leven/src/clipmap.cpp:1253-1340) as two
// batched passes instead of one blocking call per node:
//   1. ConstructClipmapNodeData (clipmap.cpp:432-468) for every node the update decided to load:
//      one lvn_meshgen_generate_simplified_batch call (chunk pass + simplifier, meshes stay in HBM);
//   2. "setting a node will invalidate its seam, so need to tell neighbours to update too"
//      (clipmap.cpp:1306-1324): the seam-update set, then GenerateClipmapSeamMesh (clipmap.cpp:573-611)
//      for every node of the set in one lvn_seam_mesh_generate_batch call.
// Host code only: which nodes exist, are active or get loaded stays the application's decision
// (LOD selection is control plane); this file only restates the two neighbour walks the update
// makes over the ACTIVE nodes -- Clipmap::findNode + findActiveNodes (clipmap.cpp:1011-1014,
// 1449-1481) -- on a flat node list: an active node belongs to a candidate cell when it contains
// the cell's min or its min lies inside the cell (AABB::pointIsInside, aabb.h:37-43).  Octree cells
// are aligned, so the reference's tree descent prunes nothing this flat test keeps.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unordered_map>
#include <chrono>
#include <vector>

#include "lvn_internal.h"

namespace {

struct CellKey {
    int x, y, z, size;
    bool operator==(const CellKey &o) const { return x == o.x && y == o.y && z == o.z && size == o.size; }
};
struct CellKeyHash {
    size_t operator()(const CellKey &k) const
    {
        unsigned long long h = 1469598103934665603ull;
        for (int v : {k.x, k.y, k.z, k.size}) { h ^= (unsigned int)v; h *= 1099511628211ull; }
        return (size_t)h;
    }
};

const int kChildMinOffsets[8][3] = {{0, 0, 0}, {0, 0, 1}, {0, 1, 0}, {0, 1, 1}, {1, 0, 0}, {1, 0, 1}, {1, 1, 0}, {1, 1, 1}};   // volume_constants.h:24-35

int floor_align(int v, int s) { const int r = v % s; return r < 0 ? v - r - s : v - r; }

// the active nodes, findable by cell
struct ActiveIndex {
    const lvn_clipmap_node *nodes = nullptr;
    std::unordered_map<CellKey, int, CellKeyHash> exact;                  // (min, size) -> node
    std::unordered_map<CellKey, std::vector<int>, CellKeyHash> inside;    // (cell min, cell size) -> smaller nodes inside that cell
    std::vector<int> sizes;                                               // node sizes present, ascending

    void build(const lvn_clipmap_node *n, const std::vector<int> &active)
    {
        nodes = n;
        for (int i : active) if (std::find(sizes.begin(), sizes.end(), n[i].size) == sizes.end()) sizes.push_back(n[i].size);
        std::sort(sizes.begin(), sizes.end());
        const int maxSize = sizes.empty() ? 0 : sizes.back();
        for (int i : active) {
            exact[CellKey{n[i].min[0], n[i].min[1], n[i].min[2], n[i].size}] = i;
            for (int s = n[i].size * 2; s > 0 && s <= maxSize; s *= 2)
                inside[CellKey{floor_align(n[i].min[0], s), floor_align(n[i].min[1], s), floor_align(n[i].min[2], s), s}].push_back(i);
        }
    }
    // FindActiveNodes(root, cell): nodes that contain the cell's min, then nodes whose min lies inside the cell
    void query(const int cmin[3], int csize, std::vector<int> &out) const
    {
        for (int s : sizes) {
            if (s < csize) continue;       // a smaller node containing the cell's min has its min inside the cell: found below
            const auto it = exact.find(CellKey{floor_align(cmin[0], s), floor_align(cmin[1], s), floor_align(cmin[2], s), s});
            if (it != exact.end()) out.push_back(it->second);
        }
        const auto it = inside.find(CellKey{cmin[0], cmin[1], cmin[2], csize});
        if (it != inside.end()) out.insert(out.end(), it->second.begin(), it->second.end());
    }
};

}  // namespace

// LVN_UPDATE_TIMING=1: host wall-clock of the update's phases on stderr (no extra synchronisation)
static bool update_timing() { static const bool on = getenv("LVN_UPDATE_TIMING") && atoi(getenv("LVN_UPDATE_TIMING")) != 0; return on; }
static double now_us() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }


// Pass 2 on its own: the seam-update set of an update (clipmap.cpp:1306-1324) over the nodes listed in
// `active`, and GenerateClipmapSeamMesh (clipmap.cpp:573-611) for the share of that set this caller
// takes: entry u of the (ascending) set belongs to shard u % shardCount.  With one GPU the share is
// everything; with G GPUs every rank holds all nodes' seam nodes after the exchange (host or device
// memory: the seam batch copies from either) and contours the seams of its share.
extern "C" int lvn_clipmap_seam_update_batch(int voxelsPerChunk, const lvn_clipmap_node *nodes, int numNodes,
                                             const int32_t *active, int numActive, const int32_t *constructed, int numConstructed,
                                             const lvn_seam_node_info *seamNodes, int64_t numSeamNodes, int shardIndex, int shardCount,
                                             lvn_mesh_vertex *vertices, int64_t vertexCapacity,
                                             lvn_mesh_triangle *triangles, int64_t triangleCapacity,
                                             int32_t *seamUpdateNodes, lvn_seam_result *seamResults, const float seamColour[3],
                                             int32_t *numSeamUpdatesAll, int32_t *numSeamUpdatesMine)
{
    if (numNodes < 0 || numActive < 0 || numConstructed < 0 || shardCount < 1 || shardIndex < 0 || shardIndex >= shardCount ||
        !numSeamUpdatesAll || !numSeamUpdatesMine || (numActive > 0 && (!nodes || !active)) || (numConstructed > 0 && !constructed))
        return LVN_ERR_INVALID_VALUE;
    *numSeamUpdatesAll = 0; *numSeamUpdatesMine = 0;
    const double tEnter = now_us();
    std::vector<int> act(active, active + numActive);
    for (int a : act) if (a < 0 || a >= numNodes) return LVN_ERR_INVALID_VALUE;
    for (int c = 0; c < numConstructed; c++) if (constructed[c] < 0 || constructed[c] >= numNodes) return LVN_ERR_INVALID_VALUE;
    ActiveIndex index;
    index.build(nodes, act);
    std::vector<char> marked((size_t)numNodes, 0);
    std::vector<int> found;
    for (int k = 0; k < numConstructed; k++) {
        const int c = constructed[k];
        for (int i = 0; i < 8; i++) {
            int cmin[3];
            for (int a = 0; a < 3; a++) cmin[a] = nodes[c].min[a] - kChildMinOffsets[i][a] * nodes[c].size;
            found.clear();
            index.query(cmin, nodes[c].size, found);
            for (int f : found) marked[f] = 1;
        }
    }
    std::vector<int> updates;
    int all = 0;
    for (int i = 0; i < numNodes; i++)
        if (marked[i]) { if (all % shardCount == shardIndex) updates.push_back(i); all++; }   // (the reference iterates an unordered_set)
    *numSeamUpdatesAll = all;
    *numSeamUpdatesMine = (int32_t)updates.size();
    if (updates.empty()) return LVN_SUCCESS;
    if (!seamUpdateNodes || !seamResults) return LVN_ERR_INVALID_VALUE;

    std::vector<lvn_seam_job> jobs(updates.size());
    std::vector<lvn_seam_neighbour> nbs;
    for (size_t u = 0; u < updates.size(); u++) {
        const lvn_clipmap_node &h = nodes[updates[u]];
        seamUpdateNodes[u] = updates[u];
        lvn_seam_job &j = jobs[u];
        memset(&j, 0, sizeof(j));
        memcpy(j.hostMin, h.min, sizeof(j.hostMin));
        j.hostSize = h.size;
        j.firstNeighbour = (int32_t)nbs.size();
        for (int a = 0; a < 3; a++) j.colour[a] = seamColour ? seamColour[a] : 1.f;
        for (int i = 0; i < 8; i++) {
            int cmin[3];
            for (int a = 0; a < 3; a++) cmin[a] = h.min[a] + kChildMinOffsets[i][a] * h.size;
            found.clear();
            index.query(cmin, h.size, found);
            for (int f : found) {
                const lvn_clipmap_node &nb = nodes[f];
                if (nb.numSeamNodes <= 0) continue;
                lvn_seam_neighbour sn;
                memset(&sn, 0, sizeof(sn));
                sn.index = i;
                memcpy(sn.min, nb.min, sizeof(sn.min));
                sn.size = nb.size;
                sn.firstNode = nb.firstSeamNode;
                sn.numNodes = nb.numSeamNodes;
                nbs.push_back(sn);
            }
        }
        j.numNeighbours = (int32_t)nbs.size() - j.firstNeighbour;
    }
    const double tJobs = now_us();
    const int rc = lvn_seam_mesh_generate_batch(voxelsPerChunk, (int)jobs.size(), jobs.data(), nbs.data(), (int)nbs.size(), seamNodes, (int)numSeamNodes,
                                                vertices, vertexCapacity, triangles, triangleCapacity, seamResults);
    if (update_timing())
        fprintf(stderr, "[lvn update]   pass 2: seam-update set + job lists on the host %.0f us, seam batch %.0f us\n", tJobs - tEnter, now_us() - tJobs);
    return rc;
}

extern "C" int lvn_clipmap_update_batch(lvn_meshgen *ctx, lvn_clipmap_node *nodes, int numActive, int numConstruct,
                                        const lvn_simplify_options *unitOptions,
                                        lvn_seam_node_info *seamNodes, int64_t seamNodesUsed, int64_t seamCapacity,
                                        lvn_mesh_vertex *vertices, int64_t vertexCapacity,
                                        lvn_mesh_triangle *triangles, int64_t triangleCapacity,
                                        lvn_chunk_result *constructResults,
                                        int32_t *seamUpdateNodes, lvn_seam_result *seamResults, const float seamColour[3],
                                        lvn_clipmap_update_totals *totals)
{
    if (!ctx || !totals || numActive < 0 || numConstruct < 0 || seamNodesUsed < 0 || seamNodesUsed > seamCapacity ||
        ((numActive + numConstruct) > 0 && (!nodes || !seamUpdateNodes || !seamResults)) || (numConstruct > 0 && !constructResults) || !unitOptions)
        return LVN_ERR_INVALID_VALUE;
    memset(totals, 0, sizeof(*totals));
    totals->seamNodesUsed = seamNodesUsed;
    const int V = lvn_meshgen_voxels_per_chunk(ctx);
    for (int i = 0; i < numActive + numConstruct; i++) {
        const lvn_clipmap_node &n = nodes[i];
        if (n.size <= 0 || n.size % (V * LVN_LEAF_SIZE_SCALE) != 0 || (n.size / (V * LVN_LEAF_SIZE_SCALE) & (n.size / (V * LVN_LEAF_SIZE_SCALE) - 1)) != 0 ||
            floor_align(n.min[0], n.size) != n.min[0] || floor_align(n.min[1], n.size) != n.min[1] || floor_align(n.min[2], n.size) != n.min[2])
            return LVN_ERR_INVALID_VALUE;      // octree cells: power-of-two multiples of the leaf size, aligned to their size
        if (i < numActive && (n.firstSeamNode < 0 || n.numSeamNodes < 0 || (int64_t)n.firstSeamNode + n.numSeamNodes > seamNodesUsed))
            return LVN_ERR_INVALID_VALUE;
    }

    const double tStart = now_us();
    // ---- 1. construct (clipmap.cpp:1253-1282) ----
    lvn_clipmap_node *construct = nodes + numActive;
    std::vector<uint8_t> hadMesh((size_t)std::max(numConstruct, 1), 0);
    if (numConstruct > 0) {
        std::vector<int32_t> minSize(4 * (size_t)numConstruct);
        for (int i = 0; i < numConstruct; i++) {
            memcpy(&minSize[4 * (size_t)i], construct[i].min, 3 * sizeof(int32_t));
            minSize[4 * (size_t)i + 3] = construct[i].size;
        }
        // (the node meshes are still on their way to the host arenas while pass 2 runs: it needs the seam nodes only)
        const int rc = lvn::generate_simplified(ctx, numConstruct, minSize.data(), unitOptions, vertices, nullptr, 0.f, vertexCapacity,
                                                triangles, triangleCapacity, seamNodes ? seamNodes + seamNodesUsed : nullptr,
                                                seamCapacity - seamNodesUsed, constructResults, nullptr, true, hadMesh.data());
        int64_t sn = 0;
        for (int i = 0; i < numConstruct; i++) {
            const lvn_chunk_result &r = constructResults[i];
            totals->nodeVertices += r.numVertices; totals->nodeTriangles += r.numTriangles; sn += r.numSeamNodes;
        }
        totals->seamNodesUsed = seamNodesUsed + sn;      // on LVN_ERR_CAPACITY: what the caller must provide
        if (rc < 0) return rc;
        for (int i = 0; i < numConstruct; i++) {
            lvn_chunk_result &r = constructResults[i];
            r.seamOffset += (int32_t)seamNodesUsed;
            construct[i].firstSeamNode = r.seamOffset;
            construct[i].numSeamNodes = r.numSeamNodes;
        }
    }
    // a node with a mesh buffer (decided before the simplifier runs, clipmap.cpp:446-466) or seam nodes
    // becomes active, the others are empty (clipmap.cpp:1269-1281).  Active nodes never nest: the
    // reference's FindActiveNodes stops at the first active node on its way down, the flat cell test
    // would report both an active ancestor and its active descendant.
    std::vector<int> active, constructed;
    for (int i = 0; i < numActive; i++) active.push_back(i);
    for (int i = 0; i < numConstruct; i++)
        if (hadMesh[i] || constructResults[i].numSeamNodes > 0) {
            active.push_back(numActive + i);
            constructed.push_back(numActive + i);
        }
    totals->numConstructedActive = (int32_t)constructed.size();

    const double tPass1 = now_us();
    // ---- 2 + 3. the seam-update set and its seam meshes, after the node meshes in the two arenas ----
    int32_t numMine = 0;
    const int rc = lvn_clipmap_seam_update_batch(V, nodes, numActive + numConstruct, active.data(), (int)active.size(),
                                                 constructed.data(), (int)constructed.size(), seamNodes, totals->seamNodesUsed, 0, 1,
                                                 vertices ? vertices + totals->nodeVertices : nullptr, vertexCapacity - totals->nodeVertices,
                                                 triangles ? triangles + totals->nodeTriangles : nullptr, triangleCapacity - totals->nodeTriangles,
                                                 seamUpdateNodes, seamResults, seamColour, &totals->numSeamUpdates, &numMine);
    for (int u = 0; u < numMine; u++) {
        lvn_seam_result &r = seamResults[u];
        totals->seamVertices += r.numVertices; totals->seamTriangles += r.numTriangles;
        r.vertexOffset += (int32_t)totals->nodeVertices;
        r.triangleOffset += (int32_t)totals->nodeTriangles;
    }
    const double tPass2 = now_us();
    const int rcWait = lvn::meshgen_wait(ctx);
    if (update_timing())
        fprintf(stderr, "[lvn update] pass 1 (construct, node meshes still in flight) %.0f us, pass 2 (seam set + seam meshes) %.0f us, wait for the node meshes %.0f us\n",
                tPass1 - tStart, tPass2 - tPass1, now_us() - tPass2);
    return rc < 0 ? rc : rcWait;
}
