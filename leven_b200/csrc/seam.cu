// seam.cu -- seam meshes between clipmap nodes on the GPU (the step right after
// the chunk-meshing path, consuming the SeamNodeInfo arrays generateChunkMesh returns).
//
// What the reference does on the CPU (one seam at a time, clipmap.cpp:573-611):
//   SelectSeamNodes      clipmap.cpp:508-569   filter each neighbour's seam nodes by the face / edge /
//                                              corner of the host node they touch
//   Octree_ConstructUpwards  octree.cpp:23-148 hash the leaves' parents level by level up to the root
//   Octree_GenerateMesh      octree.cpp:196-549 DFS vertex numbering, then the recursive
//                                              ContourCellProc / FaceProc / EdgeProc / ProcessEdge
//
// Here: one thread block per seam, many seams per launch, no recursion and no pointer tree.
//   1. select        the same filter, candidates -> compacted leaf list
//   2. order         a leaf's DFS position is the Morton code of its min (child index = x<<2|y<<1|z,
//                    volume_constants.h:24-35), so the reference's vertex order is a sort by that key:
//                    rank by counting (a seam has a few hundred leaves)
//   3. tree          "construct upwards" = insert every leaf and every ancestor cell into an
//                    open-addressing table keyed by (min, log2 size)
//   4. contour       every leaf looks at its own 12 edges.  It owns an edge when no smaller leaf
//                    touches it and no equal-sized leaf precedes it in the reference's node order
//                    (ContourProcessEdge takes the FIRST minimal node, octree.cpp:255-262); the other
//                    three cells around the edge are found by probing the table at growing sizes.
//                    The recursion's "same chunk" cut-offs (octree.cpp:292-303,358-361) are evaluated
//                    along the path the recursion would have taken to reach this edge: the face
//                    levels from the smallest cell that holds the edge in its interior down to the
//                    cross of the face, then the edge levels down to the smallest leaf.
//   5. emit          per-leaf quad counts, block scan, triangles in (vertex, edge) order
// Vertices come out in the reference's order (bit-identical arrays); triangles are the same set of
// index triples with the same winding, in (owner vertex, edge) order instead of recursion order.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "lvn_internal.h"

namespace lvn {

constexpr int SEAM_BLOCK = 1024;      // one block per seam; the launch lasts as long as its largest seam, so a seam gets a whole SM
constexpr unsigned int SEAM_INTERNAL = 0xfffffffeu;
constexpr unsigned int SEAM_NONE = 0xffffffffu;
// A seam of the default terrain has a few hundred leaves: its cell table, its sort keys and its
// quad flags fit the block's shared memory, and only larger seams use the global scratch slices.
constexpr int SEAM_SMEM_TABLE = 16384;       // entries (8 B key + 4 B value): a seam owns its SM (1024 threads at 64 registers), so it may as well
                                             // own its shared memory: 208 KB; seams of up to 1024 leaves x 8 levels keep their table on chip
constexpr int SEAM_SMEM_KEYS = SEAM_SMEM_TABLE;   // selected leaves whose sort keys fit (they lie where the table's keys will)
constexpr int SEAM_SMEM_WORDS = 1024;        // quad flag words: 2730 leaves * 12 edges / 32
constexpr size_t SEAM_SMEM_BYTES = (size_t)SEAM_SMEM_TABLE * 12 + (size_t)SEAM_SMEM_WORDS * 16;

struct SeamJobDev {
    int hostMin[3], hostSize;
    int firstNeighbour, numNeighbours;
    int firstCandidate, numCandidates;     // this job's slice of the per-candidate scratch
    unsigned long long tableOffset;        // this job's slice of the hash table
    unsigned int tableMask;                // capacity - 1 (power of two)
    int vertexBase, triangleBase;          // output arena slices (capacity = numCandidates, 4 * numCandidates * ... )
    int vertexCap, triangleCap;
    float colour[3];
    int rootLog2;                          // log2(2 * hostSize / LEAF_SIZE_SCALE)
};

struct SeamScratch {
    // per candidate (indexed within the job's slice)
    int4 *leaf;              // selected: local min xyz (units of LEAF_SIZE_SCALE), w = log2 size | corners << 8
    unsigned long long *key; // Morton DFS key of the selected leaf
    int *src;                // candidate index (into the seam node array) of the selected leaf
    int4 *sorted;            // leaves in DFS order
    int *quadCount;          // per sorted leaf
    int4 *quads;             // [12 per sorted leaf]: the four vertex indices, x < 0: none; flip in bit 30 of w
    unsigned long long *tableKeys;
    unsigned int *tableVals;
};

__device__ __forceinline__ unsigned long long seam_cell_key(int x, int y, int z, int lg)
{
    return ((unsigned long long)(unsigned int)x | ((unsigned long long)(unsigned int)y << 13) | ((unsigned long long)(unsigned int)z << 26) |
            ((unsigned long long)(unsigned int)lg << 39)) + 1ull;
}
__device__ __forceinline__ unsigned int seam_hash(unsigned long long k)
{
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return (unsigned int)k;
}
__device__ __forceinline__ void seam_insert(unsigned long long *keys, unsigned int *vals, unsigned int mask, unsigned long long k, unsigned int v)
{
    unsigned int h = seam_hash(k) & mask;
    for (;;) {
        const unsigned long long prev = atomicCAS(&keys[h], 0ull, k);
        if (prev == 0ull || prev == k) { if (v != SEAM_INTERNAL || prev == 0ull) vals[h] = v; return; }
        h = (h + 1) & mask;
    }
}
__device__ __forceinline__ unsigned int seam_find(const unsigned long long *keys, const unsigned int *vals, unsigned int mask, unsigned long long k)
{
    unsigned int h = seam_hash(k) & mask;
    for (;;) {
        const unsigned long long cur = keys[h];
        if (cur == k) return vals[h];
        if (cur == 0ull) return SEAM_NONE;
        h = (h + 1) & mask;
    }
}
// Morton code with x as the most significant bit of each triple: the DFS order of the octree
__device__ __forceinline__ unsigned long long seam_dfs_key(int x, int y, int z)
{
    unsigned long long k = 0;
#pragma unroll
    for (int b = 0; b < 13; b++)
        k |= ((unsigned long long)((x >> b) & 1) << (3 * b + 2)) | ((unsigned long long)((y >> b) & 1) << (3 * b + 1)) |
             ((unsigned long long)((z >> b) & 1) << (3 * b));
    return k;
}
__device__ __forceinline__ bool seam_filter(int index, int bx, int by, int bz, int mnx, int mny, int mnz, int mxx, int mxy, int mxz)
{   // FilterSeamNode, clipmap.cpp:508-536
    switch (index) {
    case 0: return mxx == bx || mxy == by || mxz == bz;
    case 1: return mnz == bz;
    case 2: return mny == by;
    case 3: return mny == by || mnz == bz;
    case 4: return mnx == bx;
    case 5: return mnx == bx || mnz == bz;
    case 6: return mnx == bx || mny == by;
    case 7: return mnx == bx && mny == by && mnz == bz;
    }
    return false;
}
__device__ __forceinline__ int seam_lowbit_log2(int v, int cap) { return v == 0 ? cap : min(__ffs(v) - 1, cap); }

// the node the recursion holds at cell size 2^lg for a quadrant whose leaf is L: the leaf itself
// once reached (its size >= the level), else the aligned cell around the quadrant point
__device__ __forceinline__ int3 seam_current_min(int4 L, int px, int py, int pz, int lg)
{
    const int llg = L.w & 0xff;
    if (llg >= lg) return make_int3(L.x, L.y, L.z);
    const int m = ~((1 << lg) - 1);
    return make_int3(px & m, py & m, pz & m);
}
__device__ __forceinline__ bool seam_same_chunk(int3 a, int3 b)
{   // ChunkMinForPosition (volume.cpp:30-41) on world = root + 4 * units: 64 units per chunk
    return (a.x >> 6) == (b.x >> 6) && (a.y >> 6) == (b.y >> 6) && (a.z >> 6) == (b.z >> 6);
}

// exclusive prefix of the popcounts of flags[0 .. nwords) into pref; returns the total (same in every thread)
__device__ __forceinline__ int seam_words_prefix(const unsigned int *flags, int *pref, int nwords, int *s_warp)
{
    const int tid = threadIdx.x;
    int carry = 0;
    for (int base = 0; base < nwords; base += SEAM_BLOCK) {
        const int w = base + tid;
        const int c = w < nwords ? __popc(flags[w]) : 0;
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if ((tid & 31) >= o) incl += t; }
        if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
        __syncthreads();
        int woff = 0, tot = 0;
#pragma unroll
        for (int q = 0; q < SEAM_BLOCK / 32; q++) { const int v = s_warp[q]; if (q < (tid >> 5)) woff += v; tot += v; }
        if (w < nwords) pref[w] = carry + woff + incl - c;
        carry += tot;
        __syncthreads();
    }
    return carry;
}

#ifdef LVN_SEAM_TIMING
__device__ long long g_seamTiming[16];      // [0..8) sums over blocks, [8..16) maxima
#define SPHASE(k) do { __syncthreads(); if (threadIdx.x == 0) { const long long now_ = clock64(); atomicAdd((unsigned long long *)&g_seamTiming[k], (unsigned long long)(now_ - t_)); atomicMax(&g_seamTiming[8 + (k)], now_ - t_); t_ = now_; } } while (0)
#else
#define SPHASE(k) do { } while (0)
#endif

__global__ void __launch_bounds__(SEAM_BLOCK)
k_seam(const SeamJobDev *__restrict__ jobs, const int *__restrict__ launchOrder, const lvn_seam_neighbour *__restrict__ neighbours,
       const lvn_seam_node_info *__restrict__ nodes, int voxelsPerChunk, int forceGlobal, SeamScratch ws,
       lvn_mesh_vertex *__restrict__ vertices, int *__restrict__ triangles, int4 *__restrict__ results)
{
    __shared__ int s_n, s_total, s_warp[SEAM_BLOCK / 32];
    extern __shared__ unsigned long long s_seam[];
    unsigned long long *s_tkeys = s_seam;                                             // [SEAM_SMEM_TABLE]
    unsigned long long *s_keys = s_seam;                                              // [SEAM_SMEM_KEYS]: phase 2 only
    unsigned int *s_tvals = reinterpret_cast<unsigned int *>(s_seam + SEAM_SMEM_TABLE);  // [SEAM_SMEM_TABLE]
    unsigned int *s_flags = s_tvals + SEAM_SMEM_TABLE;                                // [SEAM_SMEM_WORDS]
    int *s_fpref = reinterpret_cast<int *>(s_flags + SEAM_SMEM_WORDS);                // [SEAM_SMEM_WORDS]
    unsigned int *s_flagsB = s_flags + 2 * SEAM_SMEM_WORDS;                           // the same pair for the quads
    int *s_fprefB = reinterpret_cast<int *>(s_flags + 3 * SEAM_SMEM_WORDS);
    const int jobIndex = launchOrder[blockIdx.x];     // largest seams first
    const SeamJobDev job = jobs[jobIndex];
    const int tid = threadIdx.x;
    int4 *leaf = ws.leaf + job.firstCandidate;
    unsigned long long *key = ws.key + job.firstCandidate;
    int *src = ws.src + job.firstCandidate;
    int4 *sorted = ws.sorted + job.firstCandidate;
    int *quadCount = ws.quadCount + job.firstCandidate;
    int4 *quads = ws.quads + (size_t)job.firstCandidate * 12;
    if (tid == 0) { s_n = 0; s_total = 0; }
    __syncthreads();
#ifdef LVN_SEAM_TIMING
    long long t_ = clock64();
#endif

    // ---- 1. select (SelectSeamNodes, clipmap.cpp:542-569) ----
    {
        const int bx = job.hostMin[0] + job.hostSize, by = job.hostMin[1] + job.hostSize, bz = job.hostMin[2] + job.hostSize;
        int base = 0;
        for (int k = 0; k < job.numNeighbours; k++) {
            const lvn_seam_neighbour nb = neighbours[job.firstNeighbour + k];
            const int nodeSize = nb.size / voxelsPerChunk;                  // seamNodeSize, clipmap.cpp:399
            const int lg = 31 - __clz(max(nodeSize / LVN_LEAF_SIZE_SCALE, 1));
            for (int j = tid; j < nb.numNodes; j += SEAM_BLOCK) {
                const lvn_seam_node_info nd = nodes[nb.firstNode + j];
                const int mnx = nd.localspaceMin[0] * nodeSize + nb.min[0], mny = nd.localspaceMin[1] * nodeSize + nb.min[1],
                          mnz = nd.localspaceMin[2] * nodeSize + nb.min[2];
                const bool inside = mnx >= job.hostMin[0] && mnx < job.hostMin[0] + 2 * job.hostSize &&
                                    mny >= job.hostMin[1] && mny < job.hostMin[1] + 2 * job.hostSize &&
                                    mnz >= job.hostMin[2] && mnz < job.hostMin[2] + 2 * job.hostSize;
                if (!inside || !seam_filter(nb.index, bx, by, bz, mnx, mny, mnz, mnx + nodeSize, mny + nodeSize, mnz + nodeSize)) continue;
                const int slot = atomicAdd(&s_n, 1);
                const int lx = (mnx - job.hostMin[0]) / LVN_LEAF_SIZE_SCALE, ly = (mny - job.hostMin[1]) / LVN_LEAF_SIZE_SCALE,
                          lz = (mnz - job.hostMin[2]) / LVN_LEAF_SIZE_SCALE;
                leaf[slot] = make_int4(lx, ly, lz, lg | ((nd.localspaceMin[3] & 0xff) << 8));
                key[slot] = seam_dfs_key(lx, ly, lz);
                src[slot] = nb.firstNode + j;
            }
            base += nb.numNodes;
        }
    }
    __syncthreads();
    const int nSel = s_n;
    SPHASE(0);

    // ---- 2. DFS order by counting + vertices (GenerateVertexIndices, octree.cpp:196-233) ----
    // A coarser neighbour is returned for several of the host's slots (findActiveNodes,
    // clipmap.cpp:1449-1481) and its nodes can pass the filter of more than one: the reference
    // then links the same leaf twice, which changes nothing.  Keep the first of equal keys.
    // Up to SEAM_SMEM_KEYS leaves: a bitonic sort of (key << 16 | selection index) in shared memory; the
    // first of a run of equal keys is then the earliest selected, and a leaf's rank the number of run
    // heads before it.  Larger seams rank by counting through the global scratch.
    int n;
    if (nSel <= SEAM_SMEM_KEYS && !forceGlobal) {
        int P = 32;
        while (P < nSel) P <<= 1;
        for (int i = tid; i < P; i += SEAM_BLOCK) s_keys[i] = i < nSel ? ((key[i] << 16) | (unsigned long long)i) : ~0ull;
        __syncthreads();
        for (int k = 2; k <= P; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = tid; i < P; i += SEAM_BLOCK) {
                    const int l = i ^ j;
                    if (l > i) {
                        const unsigned long long a = s_keys[i], b = s_keys[l];
                        if (((i & k) == 0) ? (a > b) : (a < b)) { s_keys[i] = b; s_keys[l] = a; }
                    }
                }
                __syncthreads();
            }
        for (int p = tid; p < P; p += SEAM_BLOCK) {      // run heads as ballot words (P is a multiple of 32)
            const bool head = p < nSel && (p == 0 || (s_keys[p - 1] >> 16) != (s_keys[p] >> 16));
            const unsigned int m = __ballot_sync(0xffffffffu, head);
            if ((tid & 31) == 0) s_flags[p >> 5] = m;
        }
        __syncthreads();
        n = seam_words_prefix(s_flags, s_fpref, P >> 5, s_warp);
        for (int p = tid; p < nSel; p += SEAM_BLOCK) {
            const unsigned int m = s_flags[p >> 5], bit = 1u << (p & 31);
            if (!(m & bit)) continue;
            const int rank = s_fpref[p >> 5] + __popc(m & (bit - 1u));
            const int i = (int)(s_keys[p] & 0xffffull);
            sorted[rank] = leaf[i];
            if (rank < job.vertexCap) {
                const lvn_seam_node_info nd = nodes[src[i]];
                float4 *vp = reinterpret_cast<float4 *>(&vertices[job.vertexBase + rank]);
                vp[0] = make_float4(nd.position[0], nd.position[1], nd.position[2], 0.f);
                vp[1] = make_float4(nd.normal[0], nd.normal[1], nd.normal[2], 0.f);
                vp[2] = make_float4(job.colour[0], job.colour[1], job.colour[2], (float)(nd.localspaceMin[3] >> 8));
            }
        }
        __syncthreads();
    } else {
        for (int i = tid; i < nSel; i += SEAM_BLOCK) {
            const unsigned long long ki = key[i];
            int first = 1;
            for (int j = 0; j < i; j++) first &= key[j] != ki;
            quadCount[i] = first;
            if (first) atomicAdd(&s_total, 1);
        }
        __syncthreads();
        n = s_total;
        for (int i = tid; i < nSel; i += SEAM_BLOCK) {
            if (!quadCount[i]) continue;
            const unsigned long long ki = key[i];
            int rank = 0;
            for (int j = 0; j < nSel; j++) rank += (key[j] < ki) & quadCount[j];
            sorted[rank] = leaf[i];
            if (rank < job.vertexCap) {
                const lvn_seam_node_info nd = nodes[src[i]];
                float4 *vp = reinterpret_cast<float4 *>(&vertices[job.vertexBase + rank]);
                vp[0] = make_float4(nd.position[0], nd.position[1], nd.position[2], 0.f);
                vp[1] = make_float4(nd.normal[0], nd.normal[1], nd.normal[2], 0.f);
                vp[2] = make_float4(job.colour[0], job.colour[1], job.colour[2], (float)(nd.localspaceMin[3] >> 8));
            }
        }
        __syncthreads();
    }

    SPHASE(1);
    // ---- 3. leaves and all their ancestors into the table (Octree_ConstructUpwards) ----
    // table size from the seam's own leaf count (load <= 1/2), not from its candidates' (the host's bound, 3.3 x
    // as many): the zero fill of candidate-sized global tables was 63 MB of the launch's 66 MB of DRAM writes
    unsigned int tsize = 64u;
    while ((long long)tsize < (long long)n * (job.rootLog2 + 1) * 2) tsize <<= 1;
    const bool tableInSmem = !forceGlobal && tsize <= (unsigned int)SEAM_SMEM_TABLE;
    unsigned long long *tkeys = tableInSmem ? s_tkeys : ws.tableKeys + job.tableOffset;
    unsigned int *tvals = tableInSmem ? s_tvals : ws.tableVals + job.tableOffset;
    const unsigned int tmask = tableInSmem ? tsize - 1u : min(job.tableMask, tsize - 1u);
    for (unsigned int i = tid; i <= tmask; i += SEAM_BLOCK) tkeys[i] = 0ull;
    __syncthreads();
    for (int i = tid; i < n; i += SEAM_BLOCK) {
        const int4 L = sorted[i];
        const int lg = L.w & 0xff;
        seam_insert(tkeys, tvals, tmask, seam_cell_key(L.x, L.y, L.z, lg), (unsigned int)i);
        for (int g = lg + 1; g <= job.rootLog2; g++) {
            const int m = ~((1 << g) - 1);
            seam_insert(tkeys, tvals, tmask, seam_cell_key(L.x & m, L.y & m, L.z & m, g), SEAM_INTERNAL);
        }
    }
    __syncthreads();

    // ---- 4. contour: every (leaf, edge) ----
    SPHASE(2);
    // ---- 4. contour.  Of a leaf's 12 edges only those with a sign change inside the root can yield a
    //      quad: that cheap test runs on every (leaf, edge) item, the survivors are numbered by ballot
    //      words + prefix, and the table walks below run on full warps of survivors, each finding its
    //      item by a select over the words.  Which survivors yield a quad is a second set of words. ----
    const int rootUnits = 1 << job.rootLog2;
    const int nItems = n * 12, nWords = (nItems + 31) >> 5;
    const bool flagsInSmem = !forceGlobal && nWords <= SEAM_SMEM_WORDS;
    unsigned int *flagsA = flagsInSmem ? s_flags : reinterpret_cast<unsigned int *>(quadCount);
    int *prefA = flagsInSmem ? s_fpref : src;
    unsigned int *flagsB = flagsInSmem ? s_flagsB : reinterpret_cast<unsigned int *>(key);
    int *prefB = flagsInSmem ? s_fprefB : reinterpret_cast<int *>(key) + job.numCandidates;
    for (int item = tid; item < nWords * 32; item += SEAM_BLOCK) {
        bool cand = false;
        if (item < nItems) {
            const int r = item / 12, e = item - r * 12;
            const int4 L = sorted[r];
            const int s = 1 << (L.w & 0xff), corners = (L.w >> 8) & 0xff, dir = e >> 2;
            const int own = dir == 0 ? 3 - e : (dir == 1 ? (e == 7 ? 0 : (e == 5 ? 1 : (e == 6 ? 2 : 3))) : 11 - e);
            const int mp = dir == 0 ? L.y : (dir == 1 ? L.z : L.x), mq = dir == 0 ? L.z : (dir == 1 ? L.x : L.y);
            const int lp = mp + (((own >> 1) & 1) ? 0 : s), lq = mq + ((own & 1) ? 0 : s);
            const int c0 = dir == 0 ? (e & 3) : (dir == 1 ? ((e & 1) | ((e & 2) << 1)) : ((e & 3) << 1));
            const int c1 = c0 | (dir == 0 ? 4 : (dir == 1 ? 2 : 1));
            cand = (((corners >> c0) ^ (corners >> c1)) & 1) && lp > 0 && lq > 0 && lp < rootUnits && lq < rootUnits;
        }
        const unsigned int m = __ballot_sync(0xffffffffu, cand);
        if ((tid & 31) == 0) flagsA[item >> 5] = m;
    }
    __syncthreads();
    const int nLive = seam_words_prefix(flagsA, prefA, nWords, s_warp);
    const int nWordsB = (nLive + 31) >> 5;
    for (int d = tid; d < nWordsB * 32; d += SEAM_BLOCK) {
        bool ok = d < nLive;
        if (ok) {
        int lo = 0, hi = nWords - 1;            // the d-th survivor: its word by binary search over the prefix, its bit by fns
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (prefA[mid] <= d) lo = mid; else hi = mid - 1; }
        const int item = (lo << 5) + (int)__fns(flagsA[lo], 0, d - prefA[lo] + 1);
        const int r = item / 12, e = item - r * 12;
        const int4 L = sorted[r];
        const int lg = L.w & 0xff, s = 1 << lg, corners = (L.w >> 8) & 0xff;
        const int dir = e >> 2;
        // this leaf's place among the four nodes of its edge: processEdgeMask, contour_constants.h
        const int own = dir == 0 ? 3 - e : (dir == 1 ? (e == 7 ? 0 : (e == 5 ? 1 : (e == 6 ? 2 : 3))) : 11 - e);
        // perpendicular axes: bit 1 of the node index is the side along axis p, bit 0 along axis q
        const int a = dir, p = dir == 0 ? 1 : (dir == 1 ? 2 : 0), qa = dir == 0 ? 2 : (dir == 1 ? 0 : 1);
        const int mn[3] = {L.x, L.y, L.z};
        int line[3];
        line[a] = mn[a];
        line[p] = mn[p] + (((own >> 1) & 1) ? 0 : s);
        line[qa] = mn[qa] + ((own & 1) ? 0 : s);
        // sign change and winding come from this leaf (the first minimal node)
        const int c0 = dir == 0 ? (e & 3) : (dir == 1 ? ((e & 1) | ((e & 2) << 1)) : ((e & 3) << 1));   // edgevmap[e][0]
        const int c1 = c0 | (dir == 0 ? 4 : (dir == 1 ? 2 : 1));
        const int m1 = (corners >> c1) & 1;
        int4 Lq[4];
        int idx[4], pt[4][3];
        for (int i = 0; i < 4 && ok; i++) {
            pt[i][a] = line[a];
            pt[i][p] = line[p] - (((i >> 1) & 1) ? 0 : 1);
            pt[i][qa] = line[qa] - ((i & 1) ? 0 : 1);
            if (i == own) { Lq[i] = L; idx[i] = r; continue; }
            unsigned int found = SEAM_NONE;
            for (int g = lg; g < job.rootLog2; g++) {
                const int m = ~((1 << g) - 1);
                const unsigned int v = seam_find(tkeys, tvals, tmask, seam_cell_key(pt[i][0] & m, pt[i][1] & m, pt[i][2] & m, g));
                if (v == SEAM_NONE) continue;
                if (v == SEAM_INTERNAL) { if (g == lg) { ok = false; break; } continue; }   // smaller leaves own this edge
                found = v;
                break;
            }
            if (!ok) break;
            if (found == SEAM_NONE) { ok = false; break; }
            Lq[i] = sorted[found];
            idx[i] = (int)found;
            if ((Lq[i].w & 0xff) == lg && i < own) ok = false;                              // an equal-sized leaf comes first
        }
        if (ok) {
            // ---- the "same chunk" cut-offs along the recursion's path to this edge ----
            const int ap = seam_lowbit_log2(line[p], job.rootLog2), aq = seam_lowbit_log2(line[qa], job.rootLog2);
            int edgeTop = min(ap, aq);                       // first EdgeProc level (cell size 2^edgeTop)
            if (max(ap, aq) + 1 > job.rootLog2) ok = false;  // the edge lies on the root's boundary
            if (ok && ap != aq) {
                // ContourFaceProc levels, octree.cpp:350-361: the two cells across the face
                const bool bigIsP = ap > aq;
                const int big = bigIsP ? ap : aq;
                // the two quadrants on the small axis's negative side stand for the two face sides
                const int iNeg = 0, iPos = bigIsP ? 2 : 1;
                for (int g = big; g > min(ap, aq) && ok; g--) {
                    const int3 n0 = seam_current_min(Lq[iNeg], pt[iNeg][0], pt[iNeg][1], pt[iNeg][2], g);
                    const int3 n1 = seam_current_min(Lq[iPos], pt[iPos][0], pt[iPos][1], pt[iPos][2], g);
                    if (seam_same_chunk(n0, n1)) ok = false;
                }
            }
            for (int g = edgeTop; g >= lg && ok; g--) {      // ContourEdgeProc levels, octree.cpp:286-303
                const int3 n0 = seam_current_min(Lq[0], pt[0][0], pt[0][1], pt[0][2], g);
                bool all = true;
                for (int i = 1; i < 4; i++) all = all && seam_same_chunk(n0, seam_current_min(Lq[i], pt[i][0], pt[i][1], pt[i][2], g));
                if (all) ok = false;
            }
        }
        if (ok) quads[d] = make_int4(idx[0], idx[1], idx[2], idx[3] | ((m1 != 1) ? (1 << 30) : 0));   // flip = m1 != 1
        }
        const unsigned int m = __ballot_sync(0xffffffffu, ok);
        if ((tid & 31) == 0) flagsB[d >> 5] = m;
    }
    __syncthreads();

    SPHASE(3);
    // ---- 5. emit: triangles in (leaf, edge) order (ContourProcessEdge, octree.cpp:270-283) ----
    const int quadsTotal = seam_words_prefix(flagsB, prefB, nWordsB, s_warp);
    for (int d = tid; d < nLive; d += SEAM_BLOCK) {
        const unsigned int m = flagsB[d >> 5], bit = 1u << (d & 31);
        if (!(m & bit)) continue;
        const int off = prefB[d >> 5] + __popc(m & (bit - 1u));
        if (2 * off + 2 > job.triangleCap) continue;
        const int4 q = quads[d];
        const int i0 = q.x, i1 = q.y, i2 = q.z, i3 = q.w & ~(1 << 30);
        int *t = triangles + ((size_t)job.triangleBase + 2 * (size_t)off) * 3;
        if (!(q.w & (1 << 30))) { t[0] = i0; t[1] = i1; t[2] = i3; t[3] = i0; t[4] = i3; t[5] = i2; }
        else                    { t[0] = i0; t[1] = i3; t[2] = i1; t[3] = i0; t[4] = i2; t[5] = i3; }
    }
    SPHASE(4);
    if (tid == 0) {
        // Octree_GenerateMesh returns no mesh at all when there is no triangle (octree.cpp:536-540)
        results[jobIndex] = make_int4(quadsTotal > 0 ? n : 0, 2 * quadsTotal, n, (n > job.vertexCap || 2 * quadsTotal > job.triangleCap) ? 1 : 0);
    }
}

// the meshes of all seams, from their per-job slices into two dense arrays (host order = job order)
__global__ void k_seam_pack(const SeamJobDev *__restrict__ jobs, const int4 *__restrict__ packOffsets,
                            const lvn_mesh_vertex *__restrict__ vertices, const int *__restrict__ triangles,
                            lvn_mesh_vertex *__restrict__ outV, int *__restrict__ outT)
{
    const SeamJobDev job = jobs[blockIdx.x];
    const int4 po = packOffsets[blockIdx.x];   // x: vertex offset, y: vertex count, z: triangle offset, w: triangle count
    const uint4 *sv = reinterpret_cast<const uint4 *>(vertices + job.vertexBase);
    uint4 *dv = reinterpret_cast<uint4 *>(outV + po.x);
    for (int i = threadIdx.x; i < po.y * 3; i += blockDim.x) dv[i] = sv[i];
    const int *st = triangles + (size_t)job.triangleBase * 3;
    int *dt = outT + (size_t)po.z * 3;
    for (int i = threadIdx.x; i < po.w * 3; i += blockDim.x) dt[i] = st[i];
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct SeamState {
    cudaStream_t stream = nullptr;
    void *d_blob = nullptr; size_t blobCap = 0;
};
static SeamState g_seam;
static const char *g_seamError = "";

static int seam_fail(cudaError_t e) { g_seamError = cudaGetErrorString(e); cudaGetLastError(); return LVN_ERR_CUDA; }
#define SCU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return seam_fail(e_); } while (0)

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace lvn

using namespace lvn;

extern "C" const char *lvn_seam_last_error(void) { return g_seamError; }

// trianglesPerCandidate: the internal triangle scratch of a seam.  A leaf owns at most 12 quads = 24 triangles;
// the default meshes average under one triangle per candidate node, so the scratch is sized for 8 and a seam
// that needs more makes the whole call run once more at the true bound (the limit was internal,
// so the caller could not fix a capacity error by passing larger arenas).
static int seam_mesh_generate(int voxelsPerChunk, int numSeams, const lvn_seam_job *jobs,
                              const lvn_seam_neighbour *neighbours, int numNeighbours,
                              const lvn_seam_node_info *seamNodes, int numSeamNodes,
                              lvn_mesh_vertex *vertices, int64_t vertexCapacity,
                              lvn_mesh_triangle *triangles, int64_t triangleCapacity,
                              lvn_seam_result *results, int trianglesPerCandidate);

extern "C" int lvn_seam_mesh_generate_batch(int voxelsPerChunk, int numSeams, const lvn_seam_job *jobs,
                                            const lvn_seam_neighbour *neighbours, int numNeighbours,
                                            const lvn_seam_node_info *seamNodes, int numSeamNodes,
                                            lvn_mesh_vertex *vertices, int64_t vertexCapacity,
                                            lvn_mesh_triangle *triangles, int64_t triangleCapacity,
                                            lvn_seam_result *results)
{
    int first = 8;
    if (const char *e = getenv("LVN_SEAM_TRIANGLES_PER_CANDIDATE")) first = std::max(0, std::min(24, atoi(e)));   // tests: force the retry
    return seam_mesh_generate(voxelsPerChunk, numSeams, jobs, neighbours, numNeighbours, seamNodes, numSeamNodes, vertices,
                              vertexCapacity, triangles, triangleCapacity, results, first);
}

static int seam_mesh_generate(int voxelsPerChunk, int numSeams, const lvn_seam_job *jobs,
                              const lvn_seam_neighbour *neighbours, int numNeighbours,
                              const lvn_seam_node_info *seamNodes, int numSeamNodes,
                              lvn_mesh_vertex *vertices, int64_t vertexCapacity,
                              lvn_mesh_triangle *triangles, int64_t triangleCapacity,
                              lvn_seam_result *results, int trianglesPerCandidate)
{
    const size_t TPC = (size_t)trianglesPerCandidate;
    if (numSeams < 0 || voxelsPerChunk <= 0 || (numSeams > 0 && (!jobs || !results))) return LVN_ERR_INVALID_VALUE;
    if (numSeams == 0) return LVN_SUCCESS;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return LVN_ERR_NO_DEVICE; }
    if (!g_seam.stream) SCU(cudaStreamCreateWithFlags(&g_seam.stream, cudaStreamNonBlocking));

    // ---- plan: per-job slices of the scratch, the table and the output arenas ----
    std::vector<SeamJobDev> jd(numSeams);
    size_t totalCand = 0, totalTable = 0;
    for (int s = 0; s < numSeams; s++) {
        const lvn_seam_job &j = jobs[s];
        if (j.numNeighbours < 0 || j.firstNeighbour < 0 || j.firstNeighbour + j.numNeighbours > numNeighbours || j.hostSize <= 0)
            return LVN_ERR_INVALID_VALUE;
        SeamJobDev &d = jd[s];
        memcpy(d.hostMin, j.hostMin, sizeof(d.hostMin));
        d.hostSize = j.hostSize;
        d.firstNeighbour = j.firstNeighbour; d.numNeighbours = j.numNeighbours;
        memcpy(d.colour, j.colour, sizeof(d.colour));
        int cand = 0;
        for (int k = 0; k < j.numNeighbours; k++) {
            const lvn_seam_neighbour &nb = neighbours[j.firstNeighbour + k];
            if (nb.numNodes < 0 || nb.firstNode < 0 || nb.firstNode + nb.numNodes > numSeamNodes || nb.size < voxelsPerChunk * LVN_LEAF_SIZE_SCALE ||
                nb.index < 0 || nb.index > 7)
                return LVN_ERR_INVALID_VALUE;
            cand += nb.numNodes;
        }
        int rootUnits = 2 * j.hostSize / LVN_LEAF_SIZE_SCALE, lg = 0;
        while ((1 << lg) < rootUnits) lg++;
        if ((1 << lg) != rootUnits || lg > 12) return LVN_ERR_INVALID_VALUE;     // 13-bit cell coordinates
        d.rootLog2 = lg;
        d.firstCandidate = (int)totalCand; d.numCandidates = cand;
        size_t cap = 64;
        while (cap < (size_t)cand * (size_t)(lg + 1) * 2) cap <<= 1;
        d.tableOffset = totalTable; d.tableMask = (unsigned int)(cap - 1);
        d.vertexBase = (int)totalCand; d.vertexCap = cand;
        d.triangleBase = (int)(totalCand * TPC); d.triangleCap = cand * (int)TPC;
        totalCand += (size_t)cand;
        totalTable += cap;
    }
    const size_t candPad = std::max<size_t>(totalCand, 1);
    // device blob layout
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += align256(bytes); return o; };
    const size_t oJobs = take(sizeof(SeamJobDev) * numSeams), oNb = take(sizeof(lvn_seam_neighbour) * std::max(numNeighbours, 1)),
                 oNodes = take(sizeof(lvn_seam_node_info) * std::max(numSeamNodes, 1)), oLeaf = take(sizeof(int4) * candPad),
                 oKey = take(8 * candPad), oSrc = take(4 * candPad), oSorted = take(sizeof(int4) * candPad), oCnt = take(4 * candPad),
                 oQuads = take(sizeof(int4) * 12 * candPad), oTK = take(8 * std::max<size_t>(totalTable, 1)),
                 oTV = take(4 * std::max<size_t>(totalTable, 1)), oV = take(sizeof(lvn_mesh_vertex) * candPad),
                 oT = take(sizeof(int) * 3 * TPC * candPad), oRes = take(sizeof(int4) * numSeams),
                 oPack = take(sizeof(int4) * numSeams), oOrder = take(sizeof(int) * numSeams), oPV = take(sizeof(lvn_mesh_vertex) * candPad), oPT = take(sizeof(int) * 3 * TPC * candPad);
    if (off > g_seam.blobCap) {
        if (g_seam.d_blob) SCU(cudaFree(g_seam.d_blob));
        g_seam.d_blob = nullptr; g_seam.blobCap = 0;
        SCU(cudaMalloc(&g_seam.d_blob, off + off / 4));
        g_seam.blobCap = off + off / 4;
    }
    char *B = (char *)g_seam.d_blob;
    cudaStream_t st = g_seam.stream;
    SCU(cudaMemcpyAsync(B + oJobs, jd.data(), sizeof(SeamJobDev) * numSeams, cudaMemcpyHostToDevice, st));
    std::vector<int> order(numSeams);
    for (int i = 0; i < numSeams; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return jd[a].numCandidates > jd[b].numCandidates; });
    SCU(cudaMemcpyAsync(B + oOrder, order.data(), sizeof(int) * numSeams, cudaMemcpyHostToDevice, st));
    if (numNeighbours) SCU(cudaMemcpyAsync(B + oNb, neighbours, sizeof(lvn_seam_neighbour) * numNeighbours, cudaMemcpyHostToDevice, st));
    // (cudaMemcpyDefault: the seam nodes may also be the device arena of a batch view)
    if (numSeamNodes) SCU(cudaMemcpyAsync(B + oNodes, seamNodes, sizeof(lvn_seam_node_info) * numSeamNodes, cudaMemcpyDefault, st));
    SeamScratch ws;
    ws.leaf = (int4 *)(B + oLeaf); ws.key = (unsigned long long *)(B + oKey); ws.src = (int *)(B + oSrc);
    ws.sorted = (int4 *)(B + oSorted); ws.quadCount = (int *)(B + oCnt); ws.quads = (int4 *)(B + oQuads);
    ws.tableKeys = (unsigned long long *)(B + oTK); ws.tableVals = (unsigned int *)(B + oTV);
    static bool smemSet = false;
    if (!smemSet) { SCU(cudaFuncSetAttribute(k_seam, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SEAM_SMEM_BYTES)); smemSet = true; }
    k_seam<<<numSeams, SEAM_BLOCK, SEAM_SMEM_BYTES, st>>>((const SeamJobDev *)(B + oJobs), (const int *)(B + oOrder), (const lvn_seam_neighbour *)(B + oNb),
                                            (const lvn_seam_node_info *)(B + oNodes), voxelsPerChunk,
                                            getenv("LVN_SEAM_FORCE_GLOBAL") ? 1 : 0 /* test switch: the large-seam fallbacks on every seam */, ws,
                                            (lvn_mesh_vertex *)(B + oV), (int *)(B + oT), (int4 *)(B + oRes));
    SCU(cudaGetLastError());
#ifdef LVN_SEAM_TIMING
    {
        long long t[16];
        SCU(cudaStreamSynchronize(st));
        SCU(cudaMemcpyFromSymbol(t, g_seamTiming, sizeof(t)));
        fprintf(stderr, "[seam timing] %d seams, kilocycles summed over blocks: select %.0f  order %.0f  table %.0f  contour %.0f  emit %.0f\n",
                numSeams, t[0] / 1e3, t[1] / 1e3, t[2] / 1e3, t[3] / 1e3, t[4] / 1e3);
        fprintf(stderr, "[seam timing]           slowest block of each phase:   select %.0f  order %.0f  table %.0f  contour %.0f  emit %.0f\n",
                t[8] / 1e3, t[9] / 1e3, t[10] / 1e3, t[11] / 1e3, t[12] / 1e3);
        memset(t, 0, sizeof(t));
        SCU(cudaMemcpyToSymbol(g_seamTiming, t, sizeof(t)));
    }
#endif
    std::vector<int4> res(numSeams);
    SCU(cudaMemcpyAsync(res.data(), B + oRes, sizeof(int4) * numSeams, cudaMemcpyDeviceToHost, st));
    SCU(cudaStreamSynchronize(st));

    // ---- pack the meshes into the caller's arenas, in job order: offsets on the host, one pack
    //      kernel, two copies ----
    if (trianglesPerCandidate < 24) {
        // a seam whose triangles did not fit the internal scratch (its vertices always do: one per candidate at most)
        bool internal = false;
        for (int s = 0; s < numSeams; s++) internal = internal || (res[s].w && res[s].y > jd[s].triangleCap);
        if (internal)
            return seam_mesh_generate(voxelsPerChunk, numSeams, jobs, neighbours, numNeighbours, seamNodes, numSeamNodes, vertices,
                                      vertexCapacity, triangles, triangleCapacity, results, 24);
    }
    std::vector<int4> pack(numSeams);
    int64_t hv = 0, ht = 0;
    int rc = LVN_SUCCESS;
    for (int s = 0; s < numSeams; s++) {
        lvn_seam_result &r = results[s];
        r.numVertices = res[s].x; r.numTriangles = res[s].y; r.numSelectedNodes = res[s].z;
        r.vertexOffset = (int32_t)hv; r.triangleOffset = (int32_t)ht;
        r.status = 0;
        if (res[s].w || hv + r.numVertices > vertexCapacity || ht + r.numTriangles > triangleCapacity) {
            rc = LVN_ERR_CAPACITY; r.status = LVN_ERR_CAPACITY; r.numVertices = r.numTriangles = 0;
        }
        pack[s] = make_int4((int)hv, r.numVertices, (int)ht, r.numTriangles);
        hv += r.numVertices; ht += r.numTriangles;
    }
    if (hv + ht > 0) {
        const size_t needV = sizeof(lvn_mesh_vertex) * (size_t)hv, needT = sizeof(int) * 3 * (size_t)ht;
        SCU(cudaMemcpyAsync(B + oPack, pack.data(), sizeof(int4) * numSeams, cudaMemcpyHostToDevice, st));
        k_seam_pack<<<numSeams, 256, 0, st>>>((const SeamJobDev *)(B + oJobs), (const int4 *)(B + oPack), (const lvn_mesh_vertex *)(B + oV),
                                              (const int *)(B + oT), (lvn_mesh_vertex *)(B + oPV), (int *)(B + oPT));
        SCU(cudaGetLastError());
        if (hv) SCU(cudaMemcpyAsync(vertices, B + oPV, needV, cudaMemcpyDeviceToHost, st));
        if (ht) SCU(cudaMemcpyAsync(triangles, B + oPT, needT, cudaMemcpyDeviceToHost, st));
        SCU(cudaStreamSynchronize(st));
    }
    return rc;
}
